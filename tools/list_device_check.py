"""Development helper (under gpurun): K3 on device-resident SoA records (one launch), kernel time only."""
import ctypes as C, json, sys, hashlib
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pabi_b200 as pb
from pabi_b200 import workloads as W, _native
e = pb.Engine(0)
L = _native.lib()
fx = json.loads((ROOT / "tests/golden/workload_fixtures.json").read_text())["playouts"]
positions = W.playout_positions(e, 1_000_000)
n = len(positions)
rec = torch.empty(8 * n, dtype=torch.int64, device="cuda")
off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
mv = torch.empty(40 * n, dtype=torch.int16, device="cuda")
assert L.pabi_cuda_pack_records(e._ctx, positions.ctypes.data, n, rec.data_ptr(), n) == 0
tot, t = C.c_uint64(), _native.Timing()
best = 1e9
for _ in range(8):
    rc = L.pabi_cuda_generate_moves_device(e._ctx, rec.data_ptr(), n, n, off.data_ptr(), mv.data_ptr(), 40 * n, C.byref(tot), C.byref(t))
    assert rc == 0, rc
    best = min(best, t.device_ms)
o = off.cpu().numpy().view(np.uint32); m = mv[:tot.value].cpu().numpy().view(np.uint16)
ok = hashlib.sha256(o.tobytes()).hexdigest() == fx["offsets_sha256"] and hashlib.sha256(m.tobytes()).hexdigest() == fx["moves_sha256"]
print(json.dumps({"device_ms": round(best, 4), "launches": t.kernel_launches, "moves": tot.value, "exact": ok}))
