"""ncu driver for the acting kernels: encode pre-pass + iqn_act_tc_kernel in sampling mode, 65 536 envs."""
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
