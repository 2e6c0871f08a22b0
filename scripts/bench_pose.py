"""transformation_from_parameters forward + backward at B=12 on the GPU: the fused kernels vs the reference's op sequence in
PyTorch eager (wall time per call, host + device, which is what sits on the step's critical path)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ppea_depth_b200 as P
from ppea_depth_b200 import layers as L

def eager(aa, tr, inv):
    R = L.rot_from_axisangle(aa)
    t = tr.clone()
    if inv:
        return torch.matmul(R.transpose(1, 2), L.get_translation_matrix(-t))
    return torch.matmul(L.get_translation_matrix(t), R)

B = 12
aa = (0.02 * torch.randn(B, 1, 3)).cuda(); tr = (0.05 * torch.randn(B, 1, 3)).cuda(); w = torch.randn(B, 4, 4).cuda()
def run(fn, n=300):
    for _ in range(20):
        a, t = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
        (fn(a, t, True) * w).sum().backward()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        a, t = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
        (fn(a, t, True) * w).sum().backward()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
print(json.dumps({"op": "transformation_from_parameters fwd+bwd", "batch": B, "fused_us_per_call": run(P.transformation_from_parameters),
                  "torch_eager_us_per_call": run(eager)}))
