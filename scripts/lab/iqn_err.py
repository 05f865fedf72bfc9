"""lab: numerical error of the IQN kernels against the reference fixtures (tests/golden/iqn_kat.npz)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from distributional_rl_navigation_b200 import iqn_ops
from oracle import iqn_oracle as io
kat = np.load(os.path.join(ROOT, "tests/golden/iqn_kat.npz")); w = np.load(os.path.join(ROOT, "tests/golden/iqn_weights.npz"))
weights = {k: w[k] for k in w.files}
dev = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dt)
def packed_of(flat):
    p = torch.empty(iqn_ops.N_PACKED, dtype=torch.float32, device="cuda"); iqn_ops.pack(flat, p); return p
flat = dev(io.flatten(weights))
for K in (8, 32):
    for cvar in (1.0, 0.37):
        tag = f"K{K}_cvar{str(cvar).replace('.', 'p')}"
        q, _, _ = iqn_ops.forward(flat, packed_of(flat), dev(kat["fwd_x"]), dev(kat[f"fwd_taus_{tag}"]), cvar)
        ref = kat[f"fwd_q_{tag}"]
        print(tag, "forward err / max:", np.abs(q.cpu().numpy() - ref).max() / max(1.0, np.abs(ref).max()))
for B in (32, 1024):
    target = dev(kat[f"target_flat_B{B}"])
    st, ac, rw, ns, dn = [kat[f"{n}_B{B}"] for n in ("states", "actions", "rewards", "next_states", "dones")]
    taus = kat[f"taus_B{B}"]
    scratch = torch.empty(iqn_ops.train_scratch_floats(B), dtype=torch.float32, device="cuda")
    loss = torch.zeros(1, device="cuda"); grad = torch.zeros(iqn_ops.N_PARAMS, device="cuda")
    iqn_ops.loss_grad(flat, packed_of(flat), target, packed_of(target), dev(st), dev(ac.reshape(B), torch.int64), dev(rw.reshape(B)), dev(ns),
                      dev(dn.reshape(B)), dev(taus[0]), dev(taus[1]), 0.99, scratch, loss, grad)
    ref_g = kat[f"grad_B{B}"]
    print(f"B={B}: loss rel err {abs(loss.item() - float(kat[f'losses_B{B}'][0])) / abs(float(kat[f'losses_B{B}'][0])):.2e}, "
          f"grad err / max {np.abs(grad.cpu().numpy() - ref_g).max() / np.abs(ref_g).max():.2e}")
