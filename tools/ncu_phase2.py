"""Aggregate an ncu source page (--print-source cuda,sass) of a fill kernel into its phases: samples and warp instructions.
usage: python tools/ncu_phase2.py <rep> [fill.cu as profiled]   -- phases are found by comment markers in the profiled source."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []; cur = None; src = {}
for r in csv.reader(txt.splitlines()):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] == "Line No": continue
    if r[2] == "-":
        try: rows.append((cur, int(r[0]), int(r[4]), int(r[7]))); src[(cur, int(r[0]))] = r[1]
        except ValueError: pass
# phase boundaries in fill.cu from comment markers
marks = [("once per CTA", "setup"), ("validity bitmasks", "bitmask"), ("compact masked cells", "compact/pass"), ("pass valid cells through", "compact/pass"),
         ("BILINEAR: no search", "bilinear"), ("phase A1", "A1"), ("phase A2", "A2"), ("order the records by bin", "sort"),
         ("phase B, general path", "general"), ("phase B: warps draw", "phaseB dispatch"), ("replay of the near-path", "replay"),
         ("kriging of the near-path", "kriging phase"), ("queries the bitmask paths handed back", "literal"), ("the finished tile leaves", "output pass"),
         ("every generic access", "loop end"), ("helper: selection in registers", "select"), ("helper: the method's value", "finish"),
         ("helper: one near-path query", "near_query"), ("__global__ void", "kernel head"), ("helper: fewer than four", "finish_few"),
         ("helper: ordinary kriging", "kriging_four"), ("per-axis query tables", "axis tables")]
srcfile = sys.argv[2] if len(sys.argv) > 2 else "auv-real-time-interpolation_b200/csrc/fill.cu"
lines = [(k + 1, t) for k, t in enumerate(open(srcfile).read().split("\n"))]
bounds = []
for l, s in lines:
    for key, name in marks:
        if key in s: bounds.append((l, name))
bounds.sort()
def phase(l):
    name = "head"
    for b, n in bounds:
        if l >= b: name = n
    return name
agg = collections.Counter(); ins = collections.Counter()
for f, l, sm, i in rows:
    k = phase(l) if f == "fill.cu" else f
    agg[k] += sm; ins[k] += i
ts, ti = sum(agg.values()), sum(ins.values())
print(f"total samples {ts}  warp instructions {ti}")
for k, v in agg.most_common():
    print(f"{k:22s} samples {v:7d} {100*v/ts:5.1f}%   inst {ins[k]:11d} {100*ins[k]/ti:5.1f}%")
