#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by kernel PHASE: the `// ---- ` banner comments of the
source file delimit the phases; every SASS instruction is charged to the phase of the last source line of that
file seen in program order (inlined helpers are charged to their caller's phase).
usage: ncu_by_phase.py <report.ncu-rep> <cubin> <mangled kernel name> <source.cu>"""
import collections, csv, io, re, subprocess, sys
rep, cubin, fun, src = sys.argv[1:5]
base = src.split("/")[-1]
marks = [(1, "(before first banner)")]
for no, ln in enumerate(open(src), 1):
    m = re.match(r"\s*// ---- (.*?)[- ]*$", ln)
    if m: marks.append((no, m.group(1).strip()[:60]))
def phase(line):
    cur = marks[0][1]
    for no, name in marks:
        if no <= line: cur = name
    return cur
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
    if m: seq.append((cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
assert len(data) == len(seq), (len(data), len(seq))
agg, tot, ops = collections.OrderedDict(), [0, 0, 0], {}
last = marks[0][1]
for d, (li, ins) in zip(data, seq):
    if li and li[0] == base: last = phase(li[1])
    v = (int(d[ii]), int(d[ti]), int(d[si]))
    a = agg.setdefault(last, [0, 0, 0])
    for k in range(3): a[k] += v[k]; tot[k] += v[k]
    t = ins.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops.setdefault(last, collections.Counter())[op] += v[0]
print("total: warp-instr %d thread-instr %d samples %d (avg active lanes %.1f)" % (tot[0], tot[1], tot[2], tot[1] / tot[0]))
for n, a in agg.items():
    if a[0] == 0 and a[2] == 0: continue
    print(f"{n:62s} inst {100*a[0]/tot[0]:5.1f}%  lanes {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/tot[2]:5.1f}%  | " +
          ", ".join(f"{o} {100*c/max(a[0],1):.0f}%" for o, c in ops[n].most_common(6)))
