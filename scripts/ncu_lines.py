"""Aggregates an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line:
warp instructions executed and stall samples, per kernel.  usage: ncu_lines.py src.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
kern = None; hdr = None
agg = {}
opagg = {}
for r in rows:
    if len(r) == 2 and r[0] == "Function Name":
        kern = r[1]; hdr = None; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or kern is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    # columns: first "Source" is the CUDA line text, second is SASS (dict keeps the last) -> use indices
    line_no = r[0]; cuda_src = r[1]; sass = r[3]
    try:
        ie = int(d["Instructions Executed"]); smp = int(d["# Samples"])
    except (ValueError, KeyError):
        continue
    a = agg.setdefault(kern, collections.OrderedDict())
    key = (line_no, cuda_src.strip()[:110])
    v = a.setdefault(key, [0, 0])
    v[0] += ie; v[1] += smp
    op = sass.split()[0] if sass.split() else "?"
    if op.startswith("@"):
        op = sass.split()[1] if len(sass.split()) > 1 else op
    op = op.split(".")[0]
    o = opagg.setdefault(kern, collections.Counter()); o[op] += ie
for kern, a in agg.items():
    tot = sum(v[0] for v in a.values()); tots = sum(v[1] for v in a.values())
    print("=" * 100); print(kern, "warp-inst", tot, "samples", tots)
    for (ln, src), v in sorted(a.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6s %6.2f%% inst %6.2f%% stall | %s" % (ln, 100.0 * v[0] / tot, 100.0 * v[1] / max(tots, 1), src))
    print("  opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * c / tot) for k, c in opagg[kern].most_common(22)))
