"""lab: act call (encode + tcgen05 kernel) time of the library named by MNV_LIB, 65 536 envs, sampling mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from distributional_rl_navigation_b200 import iqn_ops
flat = torch.randn(35785, device="cuda") * 0.1
ptc = torch.empty(iqn_ops.packed_tc_bytes(), dtype=torch.uint8, device="cuda"); iqn_ops.pack_tc(flat, ptc)
E = 65536
obs = torch.randn(E, 26, device="cuda") * 3
for _ in range(3):
    iqn_ops.act_tc_sample(flat, ptc, obs, 0.05, 1, 1)
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        iqn_ops.act_tc_sample(flat, ptc, obs, 0.05, 1, i)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print(os.environ.get("MNV_LIB", "default"), f"act {best * 1e3:.1f} us per 65 536 envs")
