import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.argv = ["bench.py", "--steps", "2", "--warmup", "3", "--no-e2e", "--no-cpu-baseline"]
import bench
bench.main()
from ppea_depth_b200 import _cabi
lib = _cabi.lib()
out = (ctypes.c_ulonglong * 4)()
lib.ppea_win_stats(out)
print("cells", out[0], "miss0 %.4f miss1 %.4f warps-with-any-miss %.4f" % (out[1] / out[0], out[2] / out[0], out[3] * 32 / out[0]))
