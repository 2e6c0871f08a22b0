"""Splits an `ncu --page source --csv --print-source sass` dump at BAR.SYNC instructions: warp
instructions executed and stall samples per barrier-delimited phase, plus opcode mix per phase.
usage: ncu_phases.py src.csv kernel_substring"""
import csv, sys, collections
path, want = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
kern = None; hdr = None; seg = []; cur = dict(n=0, smp=0, ops=collections.Counter(), first=None)
total = 0
for r in rows:
    if len(r) >= 2 and r[0] in ("Function Name", "Kernel Name"):
        kern = r[1]; hdr = None; continue
    if len(r) > 5 and r[0] in ("Line No", "Address", "#"):
        hdr = r; continue
    if hdr is None or kern is None or want not in kern or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    sass = d.get("Source", "")
    if "Address" not in d or not d["Address"]:
        continue
    try:
        ie = int(d["Instructions Executed"]); smp = int(d["# Samples"])
    except (ValueError, KeyError):
        continue
    toks = sass.split()
    op = toks[0] if toks else "?"
    if op.startswith("@") and len(toks) > 1:
        op = toks[1]
    op = op.split(".")[0].rstrip(";")
    cur["n"] += ie; cur["smp"] += smp; cur["ops"][op] += ie; total += ie
    if cur["first"] is None: cur["first"] = d["Address"]
    if op == "BAR":
        seg.append(cur); cur = dict(n=0, smp=0, ops=collections.Counter(), first=None)
seg.append(cur)
print("total warp-inst", total)
for i, s in enumerate(seg):
    if s["n"] == 0: continue
    print("phase %2d  %6.2f%% inst  samples %6d | %s" % (i, 100.0 * s["n"] / total, s["smp"],
          ", ".join("%s %.1f" % (k, 100.0 * c / s["n"]) for k, c in s["ops"].most_common(14))))
