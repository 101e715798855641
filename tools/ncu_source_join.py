#!/usr/bin/env python
"""Join ncu's per-SASS-instruction execution counts with nvdisasm line info for one kernel of the
in-tree library, and print (1) the dynamic opcode mix per evaluated pair-warp and (2) the hottest
source lines.  Usage: tools/ncu_source_join.py REPORT.ncu-rep MANGLED_SUBSTR LAUNCH_SKIP N_WARPS"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, sub, skip, warps = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "noa_b200", "libnoa_dcs_b200.so")],
               cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
seq, cur, on = [], None, False
for l in dis.splitlines():
    if l.startswith("\t.section") or ".text." in l and l.strip().startswith(".section"):
        on = sub in l
    if re.match(r"^\s*\.section\s+\.text\.", l):
        on = sub in l
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2).strip(), cur))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                      "--kernel-name", "regex:vmap_kernel|table_kernel", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:90])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 5]
ia, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(data) >= len(seq), (len(seq), len(data))
data = data[:len(seq)]
byline, byfp, byop, smp = (collections.Counter() for _ in range(4))
tot = 0
for (txt, loc), r in zip(seq, data):
    cnt = int(r[ia]); tot += cnt
    t = txt.split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += cnt; byline[loc] += cnt; smp[loc] += int(r[ismp])
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        byfp[loc] += cnt
fp = sum(byop[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"warp instructions per warp of evaluations: {tot / warps:.1f}  (FP64 {fp / warps:.1f} = {100 * fp / tot:.1f} %)")
print("  ".join(f"{op} {c / warps:.1f}" for op, c in byop.most_common(26)))
src = {}
for f in ("dcs_math.cuh", "glibm.cuh", "dcs_kernels.cu"):
    src[f] = open(os.path.join(root, "noa_b200", "csrc", f)).read().splitlines()
print("--- hottest source lines: total, fp64, samples")
for loc, c in byline.most_common(int(os.environ.get("TOP", "40"))):
    f, ln = loc if loc else ("?", 0)
    text = src[f][ln - 1].strip()[:84] if f in src and 0 < ln <= len(src[f]) else ""
    print(f"{c / warps:7.1f} {byfp[loc] / warps:7.1f} {smp[loc]:6d}  {f}:{ln}  {text}")
