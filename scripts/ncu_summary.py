"""Summary of an .ncu-rep: key metrics, stall reasons per issue, executed-instruction mix per opcode.
usage: ncu_summary.py report.ncu-rep [rows_for_per_row_normalisation]"""
import csv, subprocess, sys, re, collections, io
rep = sys.argv[1]
rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[0]
for v in r[2:]:
    d = dict(zip(h, v))
    print("==", d.get("Kernel Name"), d.get("gpu__time_duration.sum"), "us")
    keys = ["smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    for k in keys:
        if k in d: print("  %-75s %s" % (k, d[k]))
    st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[k])) for k in h
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and d[k]]
    print("  stalls/issue:", ", ".join("%s %.2f" % kv for kv in sorted(st, key=lambda kv: -kv[1]) if kv[1] >= 0.03))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
seen = {}
for row in csv.reader(io.StringIO(src)):
    if len(row) < 8: continue
    # sass-only view: Address, Source, ..., find columns heuristically
    addr = None
    for i, c in enumerate(row[:4]):
        if re.fullmatch(r"(0x)?[0-9a-f]{6,16}", c): addr = c; ai = i; break
    if addr is None or addr in seen: continue
    sass = row[ai + 1]
    nums = [c for c in row[ai + 2:ai + 8]]
    try:
        smp = int(nums[2]); ie = int(nums[3])
    except Exception:
        continue
    seen[addr] = (sass, ie, smp)
tot = sum(v[1] for v in seen.values()); tots = sum(v[2] for v in seen.values())
op = collections.Counter(); st = collections.Counter()
for sass, ie, smp in seen.values():
    t = sass.split()
    o = t[1] if t[0].startswith("@") else t[0]
    o = o.split(".")[0]; op[o] += ie; st[o] += smp
print("unique SASS %d, warp-inst %d" % (len(seen), tot))
for o, c in op.most_common(28):
    print("  %-8s %6.2f%%  %s stall-samples %5.2f%%" % (o, 100.0 * c / tot, ("per-row %6.1f " % (c / rows_n)) if rows_n else "", 100.0 * st[o] / max(tots, 1)))
