"""Attribute an ncu --page source --csv SASS listing to CUDA source lines / device functions.

usage: python tools/ncu_by_line.py <prof_src.csv> <lib.so> <kernel-substring> [topN]
Joins the per-SASS-instruction samples with `nvdisasm -g` line info of the cubin embedded in the .so.
"""
import bisect
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
col = sys.argv[5] if len(sys.argv) > 5 else "# Samples"   # e.g. stall_long_sb, stall_short_sb, stall_wait
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate kernel section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines = []
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # outermost "inlined at" is not given here; keep the innermost location in our own file
        if m.group(1).endswith("mcb_engine.cu"):
            cur = int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
print(f"SASS instructions: disasm {len(lines)}, ncu {len(data)}")
n = min(len(lines), len(data))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mycobotgym_b200", "csrc", "mcb_engine.cu")).read().splitlines()
# device function ranges by scanning for definitions
fstart = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:template.*)?__device__ .*?(\w+)\(", l) or re.match(r"\s*__global__ void .*?(\w+)\(", l) or re.match(r"\s*__device__ .*?(\w+)\(", l)
    if m and not l.strip().startswith("//"):
        fstart.append((i, m.group(1)))
starts = [f[0] for f in fstart]


def func_of(line):
    if line is None:
        return "?"
    k = bisect.bisect_right(starts, line) - 1
    return fstart[k][1] if k >= 0 else "?"


per_line, per_func = {}, {}
tot_s = tot_i = 0
for k in range(n):
    r = data[k]
    s = float(r[ci[col]] or 0)
    ie = float(r[ci["Instructions Executed"]] or 0)
    ln = lines[k]
    a = per_line.setdefault(ln, [0.0, 0.0])
    a[0] += s; a[1] += ie
    f = per_func.setdefault(func_of(ln), [0.0, 0.0])
    f[0] += s; f[1] += ie
    tot_s += s; tot_i += ie
print(f"column {col}: total {tot_s:.0f}, warp instructions executed {tot_i:.3e}")
print("\n== by device function (samples %, instr %)")
for f, (s, ie) in sorted(per_func.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} {100 * s / tot_s:6.2f}%  {100 * ie / tot_i:6.2f}%")
print(f"\n== top {topn} source lines by samples")
for ln, (s, ie) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = src[ln - 1].strip()[:110] if ln else "?"
    print(f"  L{ln}: {100 * s / tot_s:5.2f}% smp {100 * ie / tot_i:5.2f}% ins | {text}")

# ---- opcode-class mix per device function (instructions executed) ----
CLASSES = [("fp64", r"^(DFMA|DMUL|DADD|DSETP|DMNMX)"), ("lds", r"^LDS"), ("sts", r"^STS"), ("ldst_other", r"^(LD|ST|ATOM|RED)"),
           ("branch", r"^(BRA|BSSY|BSYNC|BREAK|CALL|RET|EXIT|WARPSYNC|BAR|JMP)"), ("setp", r"^(ISETP|FSETP|PLOP3|PSETP)"),
           ("imad", r"^IMAD"), ("shfl", r"^(SHFL|VOTE|MATCH|REDUX)"), ("mufu", r"^MUFU"),
           ("int_other", r"^(LOP3|VIADD|IADD3|LEA|SHF|SEL|FSEL|MOV|UMOV|CS2R|PRMT|BREV|FLO|ULEA|S2UR|S2R|UIADD|ULOP|USHF|UIMAD|R2UR|POPC|VIADDMNMX|IABS|I2F|F2I|I2FP|F2F|NOP)")]
mix = {}
for k in range(n):
    r = data[k]
    ie = float(r[ci["Instructions Executed"]] or 0)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
    op = m.group(2) if m else "?"
    cls = next((c for c, pat in CLASSES if re.match(pat, op)), "other")
    d = mix.setdefault(func_of(lines[k]), {})
    d[cls] = d.get(cls, 0.0) + ie
names = [c for c, _ in CLASSES] + ["other"]
print("\n== opcode-class mix per device function (% of ALL executed warp instructions)")
print("  " + " " * 24 + "".join(f"{c:>11s}" for c in names) + "      total")
for f, d in sorted(mix.items(), key=lambda kv: -sum(kv[1].values()))[:24]:
    print(f"  {f:24s}" + "".join(f"{100 * d.get(c, 0) / tot_i:10.2f}%" for c in names) + f"{100 * sum(d.values()) / tot_i:10.2f}%")
allc = {c: sum(d.get(c, 0) for d in mix.values()) for c in names}
print(f"  {'ALL':24s}" + "".join(f"{100 * allc[c] / tot_i:10.2f}%" for c in names))
