"""Per-source-line dynamic instruction profile of the NUTS kernel from an `ncu --page source --csv` export (SASS view with
executed counts) joined with `nvdisasm -c -gi` line / inline info of the SAME build.
Usage: python scripts/k3_line_profile.py <source.csv> <nvdisasm -gi listing of the kernel> <kernel .cu> <device .cuh> <leaves>
(The r01 capture predates the template split: rebuild that commit's ahmc_nuts.cu to a cubin first; see profiles/README.md.)"""
import collections
import csv
import re
import sys

csv.field_size_limit(10 ** 9)
path_csv, path_sass, path_cu, path_dev, leaves = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
rows = list(csv.reader(open(path_csv)))
s = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address" and "Source" in r][0]
hdr = rows[s]
body = [r for r in rows[s + 1:] if len(r) == len(hdr)]
ix = {n: i for i, n in enumerate(hdr)}
ex = [float(r[ix["Instructions Executed"]] or 0) for r in body]
samp = [float(r[ix["# Samples"]] or 0) for r in body]
cu_name = path_cu.split("/")[-1]
dev_name = path_dev.split("/")[-1]
ann, pending, cur = [], [], None
for ln in open(path_sass).read().splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/\s+\S", ln):
        if pending:
            cur = (pending[-1][1] if pending[-1][0] == cu_name else None, pending[0])
            pending = []
        ann.append(cur)
assert len(ann) == len(ex), (len(ann), len(ex))
src, dev = open(path_cu).read().splitlines(), open(path_dev).read().splitlines()


def func_of(lines, l):
    for i in range(l - 1, -1, -1):
        m = re.search(r"(?:__device__|__global__)[^;{]*?\b([A-Za-z_0-9]+)\s*\(", lines[i])
        if m and not lines[i].strip().startswith("//"):
            return m.group(1)
    return "?"


tot, tots = sum(ex), sum(samp)
byk, byks, byf, byfs = (collections.Counter() for _ in range(4))
for c, e, sm in zip(ann, ex, samp):
    if c is None:
        continue
    kl, inner = c
    byk[kl] += e
    byks[kl] += sm
    fn = func_of(dev, inner[1]) if inner[0] == dev_name else ("[kernel body / libdevice]" if inner[0] == cu_name else inner[0])
    byf[fn] += e
    byfs[fn] += sm
print(f"{tot:.0f} warp-instructions over {leaves:.0f} leaves = {tot / leaves:.0f} per leaf; {tots:.0f} stall samples")
print("\ninnermost function: warp-instructions per leaf | % of instructions | % of stall samples")
for fn, e in byf.most_common(18):
    print(f"  {fn:28s} {e / leaves:7.1f} {100 * e / tot:6.1f}% {100 * byfs[fn] / tots:6.1f}%")
print("\nkernel source line (outermost frame): per leaf | % instr | % stalls | source")
for kl, e in byk.most_common(32):
    print(f"  L{kl}: {e / leaves:6.1f} {100 * e / tot:5.1f}% {100 * byks[kl] / tots:5.1f}%  {src[kl - 1].strip()[:100] if kl else ''}")
