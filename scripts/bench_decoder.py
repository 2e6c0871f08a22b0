"""Disparity head of DepthDecoderV2 at the KITTI head shape B=12, C=32, 192x640, forward and backward:
the fused kernels (C ABI, CUDA events around the launches) vs the reference's module sequence (ReflectionPad2d + Conv2d + Sigmoid,
cuDNN) in PyTorch eager on the same GPU.  Algorithmic bytes: forward reads x once (B*C*H*W*4) and writes disp; backward writes
grad_x once and reads x once."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
from ppea_depth_b200 import _cabi as C

B, Cn, H, W = 12, 32, 192, 640
torch.manual_seed(0)
xs = [torch.randn(B, Cn, H, W, device="cuda") for _ in range(3)]       # 3 x 189 MB > L2: every call reads from HBM
w = (torch.randn(1, Cn, 3, 3, device="cuda") * 0.1); b = torch.zeros(1, device="cuda")
disp = torch.empty(B, 1, H, W, device="cuda"); g = torch.randn(B, 1, H, W, device="cuda")
gx = torch.empty_like(xs[0]); gw = torch.empty_like(w); gb = torch.empty(1, device="cuda")
ws = torch.empty(C.lib().ppea_disp_head_workspace_bytes(B, Cn, H, W) // 4, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def timed(fn, n=30):
    for i in range(5): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

fwd = lambda i: C.check(C.lib().ppea_disp_head_forward(xs[i % 3].data_ptr(), w.data_ptr(), b.data_ptr(), disp.data_ptr(), None, B, Cn, H, W, 0.0, 0.0, st))
bwd = lambda i: C.check(C.lib().ppea_disp_head_backward(xs[i % 3].data_ptr(), w.data_ptr(), disp.data_ptr(), g.data_ptr(), gx.data_ptr(), gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), B, Cn, H, W, st))
bwd_dx = lambda i: C.check(C.lib().ppea_disp_head_backward(None, w.data_ptr(), disp.data_ptr(), g.data_ptr(), gx.data_ptr(), None, None, None, B, Cn, H, W, st))
bwd_dw = lambda i: C.check(C.lib().ppea_disp_head_backward(xs[i % 3].data_ptr(), w.data_ptr(), disp.data_ptr(), g.data_ptr(), None, gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), B, Cn, H, W, st))
t_f, t_b, t_dx, t_dw = timed(fwd), timed(bwd), timed(bwd_dx), timed(bwd_dw)

pad, conv, sig = nn.ReflectionPad2d(1), nn.Conv2d(Cn, 1, 3).cuda(), nn.Sigmoid()
def ref_f(i):
    with torch.no_grad():
        sig(conv(pad(xs[i % 3])))
def ref_fb(i):
    x = xs[i % 3].detach().requires_grad_(True)
    (sig(conv(pad(x)))).backward(g)
r_f, r_fb = timed(ref_f), timed(ref_fb, 10)
bytes_f = B * Cn * H * W * 4 + B * H * W * 4
bytes_b = 2 * B * Cn * H * W * 4 + 3 * B * H * W * 4
peak = 6550.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"op": "disp head sigmoid(Conv3x3(x))", "shape": [B, Cn, H, W], "fused_forward_ms": t_f, "fused_backward_ms": t_b, "grad_x_ms": t_dx, "grad_weight_bias_ms": t_dw,
                  "forward_gb_s": bytes_f / t_f / 1e6, "backward_gb_s": bytes_b / t_b / 1e6, "forward_frac_of_hbm_peak": bytes_f / t_f / 1e6 / peak,
                  "backward_frac_of_hbm_peak": bytes_b / t_b / 1e6 / peak, "hbm_peak_gb_s": peak,
                  "torch_eager_forward_ms": r_f, "torch_eager_forward_backward_ms": r_fb}))
