"""Join an ncu `--page source --csv` SASS table with nvdisasm line info and
aggregate executed instructions / stall samples per source line.

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top]
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# optional 5th argument "file:first-last": an extra table by the innermost frame inside that line range
# (the lines of one function: where the time inside it goes, helpers folded into their call sites)
span = sys.argv[5] if len(sys.argv) > 5 else None
# optional 6th argument "file:line": restrict that extra table to instructions inlined below this call site
within = sys.argv[6] if len(sys.argv) > 6 else None
tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(so).resolve())], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = max(tmp.glob("*.cubin"), key=lambda p: p.stat().st_size)
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's text section
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
lines = []  # (offset, file:line, text)
chains = []  # per instruction: inline chain, innermost first
cur, chain, fresh = "?", [], True
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    found = re.findall(r'File "([^"]+)", line (\d+)', l)
    if l.lstrip().startswith("//## File") and found:
        if fresh:
            chain, fresh = [], False
        for f, n in found:
            chain.append(f"{Path(f).name}:{n}")
        cur = chain[0]
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
        chains.append(list(dict.fromkeys(chain)))
        fresh = True
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
body = rows[hdr_i + 1:]
assert len(body) == len(lines), (len(body), len(lines))
agg = defaultdict(lambda: [0, 0, 0])
ops = defaultdict(int)
tot_inst = tot_samp = 0
for (off, where, text), r in zip(lines, body):
    inst = int(r[col["Instructions Executed"]])
    samp = int(r[col["# Samples"]])
    agg[where][0] += inst
    agg[where][1] += samp
    agg[where][2] += 1
    ops[text.split()[0] if not text.startswith("@") else text.split()[1]] += inst
    tot_inst += inst
    tot_samp += samp
print(f"total warp instructions {tot_inst}, samples {tot_samp}, SASS lines {len(lines)}")
src_cache = {}
def src(where):
    f, n = where.split(":")
    for d in ("pabi_b200/csrc",):
        p = Path(d) / f
        if p.exists():
            if p not in src_cache:
                src_cache[p] = p.read_text().splitlines()
            return src_cache[p][int(n) - 1].strip()[:90]
    return ""
print("--- by source line (instructions executed) ---")
for where, (inst, samp, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*inst/tot_inst:5.1f}% inst {100*samp/max(1,tot_samp):5.1f}% samp {n:4d} sass  {where:18s} {src(where)}")
def outer(chain, fname):
    """outermost frame of the chain that lies in `fname`"""
    for fr in reversed(chain):
        if fr.startswith(fname + ":"):
            return fr
    return chain[-1] if chain else "?"
for fname, title in (("chess.cuh", "by outermost chess.cuh line (callers of the helpers)"),
                     ("incremental.cuh", "by outermost incremental.cuh line (fast path and split enumeration)"), ("kernels.cuh", "by kernel line")):
    agg2 = defaultdict(lambda: [0, 0, 0])
    tcol = col.get("Thread Instructions Executed")
    for ch, r in zip(chains, body):
        k = outer(ch, fname)
        agg2[k][0] += int(r[col["Instructions Executed"]])
        agg2[k][1] += int(r[col["# Samples"]])
        if tcol is not None:
            agg2[k][2] += int(r[tcol])
    print(f"--- {title} (lanes = active threads per warp instruction) ---")
    for where, (inst, samp, thr) in sorted(agg2.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*inst/tot_inst:5.1f}% inst {100*samp/max(1,tot_samp):5.1f}% samp {thr/max(1,inst):5.1f} lanes  {where:18s} {src(where)}")
if span:
    fname, rng = span.split(":")
    first, last = (int(x) for x in rng.split("-"))
    agg3 = defaultdict(lambda: [0, 0])
    for ch, r in zip(chains, body):
        if within and within not in ch:
            continue
        k = next((fr for fr in ch if fr.startswith(fname + ":") and first <= int(fr.split(":")[1]) <= last), None)
        if k:
            agg3[k][0] += int(r[col["Instructions Executed"]])
            agg3[k][1] += int(r[col["# Samples"]])
    print(f"--- inside {span}{' below ' + within if within else ''}, by its own lines ---")
    for where, (inst, samp) in sorted(agg3.items(), key=lambda kv: int(kv[0].split(":")[1])):
        if inst * 1000 >= tot_inst:
            print(f"{100*inst/tot_inst:5.1f}% inst {100*samp/max(1,tot_samp):5.1f}% samp  {where:18s} {src(where)}")
print("--- by opcode ---")
for op, inst in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{100*inst/tot_inst:5.1f}%  {op}")
