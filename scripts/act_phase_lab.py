"""Lab: where a tile of iqn_act_tc_kernel spends its time.  The "act_timing" option makes the kernel write clock64 stamps of
CTA 0's phase boundaries (both tile groups, first 64 tiles each) into the debug buffer instead of the accumulators.
    python scripts/act_phase_lab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from distributional_rl_navigation_b200 import _lib, iqn_ops

flat = torch.randn(35785, device="cuda") * 0.1
ptc = torch.empty(iqn_ops.packed_tc_bytes(), dtype=torch.uint8, device="cuda"); iqn_ops.pack_tc(flat, ptc)
E = 65536
obs = torch.randn(E, 26, device="cuda"); taus = torch.rand(E, 32, device="cuda")
for _ in range(2):
    iqn_ops.act_tc(flat, ptc, obs, taus, 1.0)
_lib.set_option("act_timing", 1)
dbg = torch.zeros(128 * (208 + 64 + 64 + 16), dtype=torch.float32, device="cuda")
iqn_ops.act_tc(flat, ptc, obs, taus, 1.0, debug=dbg)
torch.cuda.synchronize()
_lib.set_option("act_timing", 0)
st = dbg.view(torch.int64).cpu().numpy()[:2 * 64 * 12].reshape(2, 64, 12)
names = ["top->D1 ready (feat store, sync, wait bar_a)", "epilogue 1 (+fence, sync)", "issue layer 2", "A0(next) + sync + issue layer 1a",
         "wait layer 2", "epilogue 2 (+sync)", "issue + wait layer 3", "epilogue 3 (+sync)", "issue + wait layer 4 (+1b)", "epilogue 4 (mean, argmax)"]
ends = dbg.view(torch.int64).cpu().numpy()[2 * 64 * 12:2 * 64 * 12 + 16].reshape(4, 4)
for b in range(4):
    print(f"CTA {b}: entry -> loop end {ends[b, 2] - ends[b, 0]} / {ends[b, 3] - ends[b, 1]} cycles (group 0 / 1)")
print(f"CTA 0 group 0: entry -> first tile top {st[0, 0, 0] - ends[0, 0]} cycles")
for g in range(2):
    n_it = int((st[g, :, 0] != 0).sum())
    s = st[g, 2:n_it].astype(np.float64)
    d = np.diff(s[:, :11], axis=1)
    per_tile = np.diff(s[:, 0]).mean()
    print(f"group {g}: {per_tile:.0f} cycles per tile over {n_it} tiles")
    for k, n in enumerate(names):
        print(f"   {d[:, k].mean():7.0f}  {n}")
