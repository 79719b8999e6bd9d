#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.
usage: ncu_by_line.py <report.ncu-rep> <cubin> <mangled kernel name> [top]
(the cubin comes from `cuobjdump -xelf all libauvi.so`; needs -lineinfo at compile time)"""
import collections, csv, io, re, subprocess, sys
rep, cubin, fun = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-fun", fun, cubin], capture_output=True, text=True).stdout
if not dis.strip():
    full = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    i = full.index(".text." + fun + ":")
    j = full.find("//--------------------- .text.", i)
    dis = full[i:j if j > 0 else None]
cur, seq = None, []
for ln in dis.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
assert len(data) == len(seq), (len(data), len(seq))
agg, tot = collections.defaultdict(lambda: [0, 0, 0]), [0, 0, 0]
for d, (li, ins) in zip(data, seq):
    v = (int(d[ii]), int(d[ti]), int(d[si]))
    for k in range(3):
        agg[li][k] += v[k]; tot[k] += v[k]
print("total: warp-instr %d thread-instr %d samples %d  (avg active lanes %.1f)" % (tot[0], tot[1], tot[2], tot[1] / tot[0]))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(k):34s} inst {100*a[0]/tot[0]:5.1f}%  thread-inst {100*a[1]/tot[1]:5.1f}%  lanes {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/tot[2]:5.1f}%")
