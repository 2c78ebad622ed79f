"""List the local-memory (spill) instructions of a kernel with their source lines.
    python tools/spill_lines.py <lib.so> <mangled-kernel-substring>"""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

so, kern = sys.argv[1], sys.argv[2]
tmp = Path(tempfile.mkdtemp(dir=Path(__file__).resolve().parent.parent / "gpurun_out"))
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(so).resolve())], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = max(tmp.glob("*.cubin"), key=lambda p: p.stat().st_size)
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
chain, fresh = [], True
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    found = re.findall(r'File "([^"]+)", line (\d+)', l)
    if l.lstrip().startswith("//## File") and found:
        if fresh:
            chain, fresh = [], False
        chain += [f"{Path(f).name}:{n}" for f, n in found]
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        fresh = True
        if re.search(r"\b(STL|LDL)\b", m.group(2)):
            print(m.group(1), m.group(2).strip()[:60].ljust(60), " <- ".join(dict.fromkeys(chain)))
for f in tmp.iterdir():
    f.unlink()
tmp.rmdir()
