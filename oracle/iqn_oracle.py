"""TEST INFRASTRUCTURE ONLY -- numpy (fp32) restatement of the reference IQN path; never on the product path.

Follows RobustFieldAutonomyLab/Distributional_RL_Navigation @ e77bbbf:
    thirdparty/IQN/model.py:111-191   ObsEncoder (calc_cos :141-158, forward :160-186, get_qvals :188-191)
    thirdparty/IQN/agent.py:186-205   IQNAgent.act      :249-267 adjust_cvar
    thirdparty/IQN/agent.py:269-304   IQNAgent.train    :401-407 calculate_huber_loss
    torch.nn.utils.clip_grad_norm_ (max_norm 0.5, agent.py:299) and torch.optim.Adam defaults (agent.py:66)
The taus are INJECTED (the reference draws them with torch.rand on the CPU generator, model.py:149; Q9: the target
network draws first).  The backward pass is written out by hand -- it is the specification of the CUDA kernel.

PARITY PINNED: tests/test_iqn_oracle.py checks this file against (1) the reference's own PyTorch modules imported
in-process with torch.rand patched to the injected taus (forward, loss, every gradient, clipped Adam step) and (2) the
committed fixtures tests/golden/iqn_*.npz generated from the reference (pretrained weights network_params.pth, loss KATs
266.52407837 @B=32 / 307.13052368 @B=1024 of SURVEY.md 8(c)).
"""
import numpy as np

F = np.float32

# state_dict key order of ObsEncoder (model.py:125-136) and the shapes; flat parameter vector = concatenation
PARAM_SPECS = [
    ("velocity_encoder.weight", (16, 2)), ("velocity_encoder.bias", (16,)),
    ("goal_encoder.weight", (16, 2)), ("goal_encoder.bias", (16,)),
    ("sensor_encoder.weight", (176, 22)), ("sensor_encoder.bias", (176,)),
    ("cos_embedding.weight", (208, 64)), ("cos_embedding.bias", (208,)),
    ("hidden_layer.weight", (64, 208)), ("hidden_layer.bias", (64,)),
    ("hidden_layer_2.weight", (64, 64)), ("hidden_layer_2.bias", (64,)),
    ("output_layer.weight", (9, 64)), ("output_layer.bias", (9,)),
]
N_PARAMS = sum(int(np.prod(s)) for _, s in PARAM_SPECS)     # 35 785
PIS = np.array([np.pi * i for i in range(64)], dtype=np.float32)   # model.py:130 (FloatTensor of float64 pi*i)


def unflatten(flat):
    out, o = {}, 0
    for name, shape in PARAM_SPECS:
        n = int(np.prod(shape))
        out[name] = np.asarray(flat[o:o + n], F).reshape(shape)
        o += n
    assert o == N_PARAMS
    return out


def flatten(params):
    return np.concatenate([np.asarray(params[name], F).ravel() for name, _ in PARAM_SPECS])


def forward(P, x, taus, cvar=1.0, keep=False):
    """ObsEncoder.forward with injected taus. x f32 [B,26], taus f32 [B,N] (pre-distortion) -> quantiles [B,N,9]."""
    x = np.asarray(x, F); B = x.shape[0]
    taus = (np.asarray(taus, F) * F(cvar)).astype(F)                           # model.py:153
    N = taus.shape[1]
    feat = np.concatenate([x[:, 0:2] @ P["velocity_encoder.weight"].T + P["velocity_encoder.bias"],
                           x[:, 2:4] @ P["goal_encoder.weight"].T + P["goal_encoder.bias"],
                           x[:, 4:26] @ P["sensor_encoder.weight"].T + P["sensor_encoder.bias"]], axis=1).astype(F)   # :169-172
    cos = np.cos((taus[:, :, None] * PIS[None, None, :]).astype(F)).astype(F)    # :155, fp32 product then fp32 cos
    cosr = cos.reshape(B * N, 64)
    zc = (cosr @ P["cos_embedding.weight"].T + P["cos_embedding.bias"]).astype(F)
    c = np.maximum(zc, 0).reshape(B, N, 208)                                     # :177
    h0 = (feat[:, None, :] * c).reshape(B * N, 208).astype(F)                    # :180
    h1 = np.maximum(h0 @ P["hidden_layer.weight"].T + P["hidden_layer.bias"], 0).astype(F)
    h2 = np.maximum(h1 @ P["hidden_layer_2.weight"].T + P["hidden_layer_2.bias"], 0).astype(F)
    q = (h2 @ P["output_layer.weight"].T + P["output_layer.bias"]).astype(F).reshape(B, N, 9)
    if keep:
        return q, dict(x=x, taus=taus, feat=feat, cos=cosr, c=c.reshape(B * N, 208), h0=h0, h1=h1, h2=h2)
    return q


def get_qvals(P, x, taus, cvar=1.0):
    """ObsEncoder.get_qvals (model.py:188-191): mean over the K quantile samples."""
    return forward(P, x, taus, cvar).mean(axis=1)


def adjust_cvar(state):
    """IQNAgent.adjust_cvar (agent.py:249-267)."""
    s = np.asarray(state)[4:]
    closest = np.inf
    for i in range(0, len(s), 2):
        if abs(s[i]) < 1e-3 and abs(s[i + 1]) < 1e-3:
            continue
        closest = min(closest, float(np.linalg.norm(s[i:i + 2])))
    return closest / 10.0 if closest < 10.0 else 1.0


def loss_and_grad(P_local, P_target, states, actions, rewards, next_states, dones, taus_target, taus_local, gamma=0.99, n_step=1):
    """IQNAgent.train up to loss.backward() (agent.py:276-298). Returns loss (f32) and the flat gradient (f32 [35785])."""
    B = states.shape[0]
    N = taus_local.shape[1]
    assert N == 8 and taus_target.shape[1] == 8                                   # hard-coded 8 (agent.py:286,290)
    qt = forward(P_target, next_states, taus_target)                              # [B,8,9] (drawn FIRST: Q9)
    q_next = qt.max(axis=2)                                                       # max over actions per quantile, :280
    r = np.asarray(rewards, F).reshape(B, 1); d = np.asarray(dones, F).reshape(B, 1)
    T = (r + F(gamma ** n_step) * q_next * (F(1.0) - d)).astype(F)                # [B,8] indexed by j, :283
    ql, A = forward(P_local, states, taus_local, keep=True)
    a = np.asarray(actions).reshape(B).astype(np.int64)
    Eq = ql[np.arange(B), :, a]                                                   # [B,8] indexed by i, :286
    td = (T[:, None, :] - Eq[:, :, None]).astype(F)                               # td[b,i,j] = T_j - E_i, :289
    absd = np.abs(td)
    huber = np.where(absd <= 1.0, F(0.5) * td * td, absd - F(0.5)).astype(F)      # :401-407, k = 1
    tau = A["taus"][:, :, None]                                                   # [B,8,1]: tau_i
    w = np.abs(tau - (td < 0).astype(F)).astype(F)
    ql_ = (w * huber).astype(F)
    loss = ql_.sum(axis=1).mean(axis=1).mean()                                    # sum_i, mean_j, mean_b  :294-295
    # ---- backward ----
    dE = (-(w * np.clip(td, -1.0, 1.0)).sum(axis=2) / F(8.0 * B)).astype(F)       # [B,8]
    dq = np.zeros((B, N, 9), F)
    dq[np.arange(B), :, a] = dE
    dq = dq.reshape(B * N, 9)
    G = {}
    G["output_layer.weight"] = dq.T @ A["h2"]; G["output_layer.bias"] = dq.sum(0)
    dz2 = (dq @ P_local["output_layer.weight"]) * (A["h2"] > 0)
    G["hidden_layer_2.weight"] = dz2.T @ A["h1"]; G["hidden_layer_2.bias"] = dz2.sum(0)
    dz1 = (dz2 @ P_local["hidden_layer_2.weight"]) * (A["h1"] > 0)
    G["hidden_layer.weight"] = dz1.T @ A["h0"]; G["hidden_layer.bias"] = dz1.sum(0)
    dh0 = dz1 @ P_local["hidden_layer.weight"]                                    # [BN,208]
    feat_rep = np.repeat(A["feat"], N, axis=0)
    dzc = dh0 * feat_rep * (A["c"] > 0)
    G["cos_embedding.weight"] = dzc.T @ A["cos"]; G["cos_embedding.bias"] = dzc.sum(0)
    dfeat = (dh0 * A["c"]).reshape(B, N, 208).sum(axis=1)
    x = A["x"]
    G["velocity_encoder.weight"] = dfeat[:, 0:16].T @ x[:, 0:2]; G["velocity_encoder.bias"] = dfeat[:, 0:16].sum(0)
    G["goal_encoder.weight"] = dfeat[:, 16:32].T @ x[:, 2:4]; G["goal_encoder.bias"] = dfeat[:, 16:32].sum(0)
    G["sensor_encoder.weight"] = dfeat[:, 32:208].T @ x[:, 4:26]; G["sensor_encoder.bias"] = dfeat[:, 32:208].sum(0)
    return F(loss), flatten(G)


def clip_adam(flat_params, flat_grad, m, v, step, lr=1e-4, max_norm=0.5, beta1=0.9, beta2=0.999, eps=1e-8):
    """clip_grad_norm_(params, 0.5) (agent.py:299) + Adam.step (agent.py:66,300; torch defaults, no weight decay).
    step = number of optimizer steps INCLUDING this one.  Returns (params, m, v, total_norm)."""
    g = np.asarray(flat_grad, F)
    norms = []
    o = 0
    for _, shape in PARAM_SPECS:                       # torch: norm of the per-tensor norms
        n = int(np.prod(shape)); norms.append(np.sqrt(np.sum(g[o:o + n].astype(F) ** 2, dtype=F))); o += n
    total = np.sqrt(np.sum(np.asarray(norms, F) ** 2, dtype=F))
    coef = min(F(max_norm) / (total + F(1e-6)), F(1.0))
    g = (g * F(coef)).astype(F)
    m = (F(beta1) * m + F(1 - beta1) * g).astype(F)
    v = (F(beta2) * v + F(1 - beta2) * g * g).astype(F)
    bc1 = 1 - beta1 ** step; bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (np.sqrt(v) / F(np.sqrt(bc2)) + F(eps)).astype(F)
    p = (np.asarray(flat_params, F) - F(step_size) * (m / denom)).astype(F)
    return p, m, v, total
