"""IQN golden fixtures from the UNMODIFIED reference (thirdparty/IQN) -- see make_golden.py.

  iqn_weights.npz    the reference's pretrained IQN weights (pretrained_models/IQN/seed_3/network_params.pth), 14 tensors
  iqn_kat.npz        forward / get_qvals / adjust_cvar / train() known answers with INJECTED taus (torch.rand patched):
                     B=32 and B=1024 batches of SURVEY.md 8(c): loss, flat gradient (pre-clip), total grad norm,
                     parameters + Adam moments after 1 and 3 train() calls, and the two loss KATs under
                     torch.manual_seed(1234) (266.52407837 / 307.13052368).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402
from oracle.iqn_oracle import PARAM_SPECS  # noqa: E402


def kat_batch(B):
    import torch
    idx = torch.arange(B * 26, dtype=torch.float64).view(B, 26)
    states = (5 * torch.sin(0.37 * idx)).float()
    next_states = (5 * torch.cos(0.11 * idx + 1)).float()
    b = torch.arange(B)
    actions = (b % 9).long().view(B, 1)
    rewards = (-1 + 0.2 * torch.sin(0.5 * b.double())).float().view(B, 1)
    dones = (b % 7 == 0).float().view(B, 1)
    return states, actions, rewards, next_states, dones


def make_iqn_goldens():
    import torch
    torch.set_num_threads(1)
    _, ref_agent, ref_model = ref_import.load_reference()
    pm = os.path.join(ref_import.REF_ROOT, "pretrained_models", "IQN", "seed_3")
    sd = torch.load(os.path.join(pm, "network_params.pth"), map_location="cpu")
    assert list(sd.keys()) == [n for n, _ in PARAM_SPECS]
    np.savez_compressed(os.path.join(HERE, "iqn_weights.npz"), **{k: v.numpy() for k, v in sd.items()})

    out = {}
    real_rand = torch.rand
    for B in (32, 1024):
        agent = ref_agent.IQNAgent(26, 9, BATCH_SIZE=B, seed=0)
        agent.load_model(pm)
        batch = kat_batch(B)
        # --- KAT with the CPU generator ---
        torch.manual_seed(1234)
        loss = agent.train(batch)
        out[f"kat_loss_seed1234_B{B}"] = np.float32(loss)
        # --- injected taus ---
        agent = ref_agent.IQNAgent(26, 9, BATCH_SIZE=B, seed=0)
        agent.load_model(pm)
        # make the target differ from the local network (as it does during training)
        with torch.no_grad():
            g = torch.Generator().manual_seed(5)
            for p in agent.qnetwork_target.parameters():
                p.add_(0.02 * torch.randn(p.shape, generator=g))
        out[f"target_flat_B{B}"] = np.concatenate([p.detach().numpy().ravel() for p in agent.qnetwork_target.parameters()])
        rs = np.random.RandomState(100 + B)
        tau_seq = [rs.rand(B, 8).astype(np.float32) for _ in range(6)]
        out[f"taus_B{B}"] = np.stack(tau_seq)                      # order of use: target, local, target, local, ...
        it = iter(tau_seq)
        torch.rand = lambda *shape, **kw: torch.from_numpy(next(it))
        try:
            # first call: also capture the raw gradient before clipping
            agent.optimizer.zero_grad()
            losses = []
            for k in range(3):
                if k == 0:
                    import torch.nn.utils as U
                    real_clip = U.clip_grad_norm_
                    grabbed = {}

                    def grab(params, max_norm):
                        params = list(params)
                        grabbed["grad"] = np.concatenate([p.grad.detach().numpy().ravel() for p in params]).copy()
                        tn = real_clip(params, max_norm)
                        grabbed["norm"] = float(tn)
                        return tn
                    U.clip_grad_norm_ = grab
                    torch.nn.utils.clip_grad_norm_ = grab
                losses.append(float(agent.train(batch)))
                if k == 0:
                    U.clip_grad_norm_ = real_clip
                    torch.nn.utils.clip_grad_norm_ = real_clip
                    out[f"grad_B{B}"] = grabbed["grad"]; out[f"gradnorm_B{B}"] = np.float32(grabbed["norm"])
                    out[f"params_after1_B{B}"] = np.concatenate([p.detach().numpy().ravel() for p in agent.qnetwork_local.parameters()])
            out[f"losses_B{B}"] = np.asarray(losses, np.float32)
            out[f"params_after3_B{B}"] = np.concatenate([p.detach().numpy().ravel() for p in agent.qnetwork_local.parameters()])
            st = agent.optimizer.state_dict()["state"]
            out[f"adam_m_after3_B{B}"] = np.concatenate([st[i]["exp_avg"].numpy().ravel() for i in range(14)])
            out[f"adam_v_after3_B{B}"] = np.concatenate([st[i]["exp_avg_sq"].numpy().ravel() for i in range(14)])
        finally:
            torch.rand = real_rand
        for k, name in zip(batch, ("states", "actions", "rewards", "next_states", "dones")):
            out[f"{name}_B{B}"] = k.numpy()

    # --- forward / get_qvals with injected taus, K = 32 and N = 8, cvar 1 and 0.37 ---
    net = ref_model.ObsEncoder.load(pm)
    rs = np.random.RandomState(9)
    x = (rs.randn(64, 26) * 3).astype(np.float32)
    x[:, 4:] *= (rs.rand(64, 22) > 0.5)
    for K in (8, 32):
        taus = rs.rand(64, K).astype(np.float32)
        for cvar in (1.0, 0.37):
            torch.rand = lambda *shape, **kw: torch.from_numpy(taus)
            try:
                with torch.no_grad():
                    q, t = net.forward(torch.from_numpy(x), K, cvar)
            finally:
                torch.rand = real_rand
            tag = f"K{K}_cvar{str(cvar).replace('.', 'p')}"
            out[f"fwd_taus_{tag}"] = taus; out[f"fwd_q_{tag}"] = q.numpy(); out[f"fwd_taus_out_{tag}"] = t.numpy()
    out["fwd_x"] = x
    # --- adjust_cvar ---
    agent = ref_agent.IQNAgent(26, 9, seed=0)
    obs = (rs.randn(200, 26) * 4)
    obs[:, 4:] *= (rs.rand(200, 22) > 0.6)
    obs[::5, 4:] = 0.0
    obs[1::7, 4:] *= 1e-4
    out["cvar_obs"] = obs
    out["cvar_val"] = np.array([agent.adjust_cvar(o) for o in obs])
    np.savez_compressed(os.path.join(HERE, "iqn_kat.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "kat_loss" in k or "losses" in k},
          out["kat_loss_seed1234_B32"], out["kat_loss_seed1234_B1024"])


if __name__ == "__main__":
    make_iqn_goldens()
