import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppea_depth_b200 as P
from oracle import matching_oracle as M
B, Fr, C, h, w, D = 12, 1, 64, 48, 160, 96
cur, look, poses, K, invK, bins = M.synthetic_case(B=B, Fr=Fr, C=C, h=h, w=w, D=D, seed=0, min_bin=0.3, max_bin=30.0)
g = [t.cuda() for t in (cur, look, poses, K, invK)]
cost, missing = P.match_features(*g, bins, True)
want = M.cost_volume_tail(cost.cpu(), missing.cpu())
vol = cost.clone()
conf, mins, argmin = P.cost_volume_tail(vol, missing)
print("tail equal to oracle:", torch.equal(conf.cpu(), want[0]), torch.equal(mins.cpu(), want[1]), torch.equal(argmin.cpu(), want[2]), torch.equal(vol.cpu(), want[3]), float(want[0].mean()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
vols = [cost.clone() for _ in range(20)]
for v in vols[:3]: P.cost_volume_tail(v, missing)
torch.cuda.synchronize(); a.record()
for v in vols[3:]: P.cost_volume_tail(v, missing)
b.record(); torch.cuda.synchronize()
print("tail ms", a.elapsed_time(b) / 17)
