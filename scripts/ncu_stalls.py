"""Top SASS instructions of an .ncu-rep by stall samples, with the dominant reason; usage: ncu_stalls.py rep [top]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
want = sys.argv[3] if len(sys.argv) > 3 else None      # substring of the kernel name (default: first kernel)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []; take = want is None; nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        take = (want in r[1]) if want else (nk == 1)
        continue
    if r and r[0] == "Address": hdr = r; continue
    if take and hdr and len(r) == len(hdr) and re.fullmatch(r"0x[0-9a-f]+", r[0]): data.append(r)
ix = {k: i for i, k in enumerate(hdr)}
reasons = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[ix["# Samples"]]) for r in data)
agg = {k: sum(int(r[ix[k]]) for r in data) for k in reasons}
print("total samples", tot, {k[6:]: "%.1f%%" % (100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
base = int(data[0][0], 16)
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]; n = int(r[ix["# Samples"]])
    rs = sorted(((int(r[ix[k]]), k[6:]) for k in reasons), reverse=True)[:2]
    print("%5x %6d %5.2f%% %-28s | %s" % (int(r[0], 16) - base, n, 100.0 * n / tot, ",".join("%s:%d" % (k, v) for v, k in rs if v), r[1].strip()[:90]))
