"""Diagnostic: per-chunk start / end times of the streaming kernel (library built with -DPPEA_STREAM_CLOCKS, loaded through
PPEA_LIB).  Prints the distribution of the chunks' finishing times relative to the launch and the per-SM spread."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ppea_depth_b200 import _cabi as C

wl = dict(batch=12, height=192, width=640, num_scales=4)
tset = bench.make_sets(wl, 1, 0, False)[0]
plan = bench.build_plan(tset, wl, torch.device("cuda"), False, True)
plan.capture()
for _ in range(5):
    plan.replay()
torch.cuda.synchronize()
n = 1776
clk = (ctypes.c_ulonglong * (2 * n))()
sm = (ctypes.c_uint * n)()
rows = (ctypes.c_int * (2 * n))()
lib = C.lib()
assert lib.ppea_debug_stream_clocks(clk, sm, rows, n) == 0
t = np.array(clk[:], dtype=np.int64).reshape(n, 2)
r = np.array(rows[:], dtype=np.int64).reshape(n, 2)
ok = (t[:, 1] > 0) & (r[:, 1] > r[:, 0])
t, r, s = t[ok], r[ok], np.array(sm[:])[ok]
n = len(t)
t0 = t[:, 0].min()
start = (t[:, 0] - t0) / 1e3
end = (t[:, 1] - t0) / 1e3
dur = end - start
print("chunks", n, "rows per chunk min/max", (r[:, 1] - r[:, 0]).min(), (r[:, 1] - r[:, 0]).max(), "kernel span us %.1f" % end.max())
print("start us pct 0/50/90/100: %s" % np.percentile(start, [0, 50, 90, 100]).round(1))
print("end   us pct 0/5/25/50/75/95/100: %s" % np.percentile(end, [0, 5, 25, 50, 75, 95, 100]).round(1))
print("mean end / max end = %.3f  (1 - this = share of warp-slot time idle at the end)" % (end.mean() / end.max()))
per_sm_end = np.array([end[s == k].max() for k in np.unique(s)])
print("per-SM last end us pct 0/25/50/75/100: %s" % np.percentile(per_sm_end, [0, 25, 50, 75, 100]).round(1))
# least squares: duration ~ sum_s rows_s * c_s + (pieces - 1) * c_p
H, S = wl["height"], wl["num_scales"]
A = np.zeros((n, S + 1))
for i, (a, b) in enumerate(r):
    g = np.arange(a, b)
    sc = (g // H) % S
    for k in range(S):
        A[i, k] = (sc == k).sum()
    A[i, S] = len(np.unique(g // H)) - 1
coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
print("us per row of scale 0..%d: %s   relative to the coarsest: %s   extra piece: %.2f us = %.1f coarse rows"
      % (S - 1, coef[:S].round(3), (coef[:S] / coef[S - 1]).round(3), coef[S], coef[S] / coef[S - 1]))
