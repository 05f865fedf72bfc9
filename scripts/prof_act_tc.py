"""ncu driver / timing lab for the acting kernels: encode pre-pass + iqn_act_tc_kernel in sampling mode, 65 536 envs.
   python scripts/prof_act_tc.py [--time]   (--time: CUDA-event timing of 20 calls, prints us per call)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributional_rl_navigation_b200 import iqn_ops
flat = torch.randn(35785, device="cuda") * 0.1
ptc = torch.empty(iqn_ops.packed_tc_bytes(), dtype=torch.uint8, device="cuda"); iqn_ops.pack_tc(flat, ptc)
E = 65536
obs = torch.randn(E, 26, device="cuda")
for i in range(3):
    iqn_ops.act_tc_sample(flat, ptc, obs, 0.05, 1, i)
torch.cuda.synchronize()
if "--time" in sys.argv:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        iqn_ops.act_tc_sample(flat, ptc, obs, 0.05, 1, 3 + i)
    e1.record(); torch.cuda.synchronize()
    print(f"act_tc_sample: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us per 65536-env call")
