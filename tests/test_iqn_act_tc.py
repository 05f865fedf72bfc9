"""GPU: the tcgen05 acting kernel (iqn_act_tc) against a bf16-emulating reference (stage by stage) and the fp32 kernel."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from distributional_rl_navigation_b200 import iqn_ops  # noqa: E402
from oracle import iqn_oracle as io  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module")
def weights(golden_dir):
    w = np.load(os.path.join(golden_dir, "iqn_weights.npz"))
    return {k: w[k] for k in w.files}


def bf(a):
    """round-to-nearest-even bf16 rounding of a float32 array (emulation of the kernel's operand precision)."""
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def emulate(P, x, taus, cvar=1.0):
    """The kernel's arithmetic: fp32 encoders, bf16 operands for the four GEMMs, fp32 accumulation."""
    B, K = taus.shape
    t = (taus * np.float32(cvar)).astype(np.float32)
    feat = np.concatenate([x[:, 0:2] @ P["velocity_encoder.weight"].T + P["velocity_encoder.bias"],
                           x[:, 2:4] @ P["goal_encoder.weight"].T + P["goal_encoder.bias"],
                           x[:, 4:26] @ P["sensor_encoder.weight"].T + P["sensor_encoder.bias"]], axis=1).astype(np.float32)
    cos = np.cos(np.pi * np.arange(64)[None, None, :] * t[:, :, None].astype(np.float64)).astype(np.float32).reshape(B * K, 64)
    # the biases ride inside the GEMMs as one extra reduction column (bf16-rounded like the weights); layer 1's bias is
    # folded into the weight of cos_0 = 1 (bf16 sum of the two bf16 values, iqn_act_tc.cu)
    wc = bf(P["cos_embedding.weight"]).copy()
    wc[:, 0] = bf(wc[:, 0] + bf(P["cos_embedding.bias"]))
    d1 = bf(cos) @ wc.T
    h0 = bf(np.maximum(d1, 0)) * bf(np.repeat(feat, K, axis=0))       # relu -> bf16, bf16 features, bf16 product (mul.bf16x2)
    d2 = bf(h0) @ bf(P["hidden_layer.weight"]).T + bf(P["hidden_layer.bias"])
    d3 = bf(np.maximum(d2, 0)) @ bf(P["hidden_layer_2.weight"]).T + bf(P["hidden_layer_2.bias"])
    d4 = bf(np.maximum(d3, 0)) @ bf(P["output_layer.weight"]).T + bf(P["output_layer.bias"])
    q = d4.reshape(B, K, 9).mean(axis=1)
    return d1, d2, d3, d4, q


def setup(weights):
    flat = torch.from_numpy(io.flatten(weights)).to(DEV)
    ptc = torch.empty(iqn_ops.packed_tc_bytes(), dtype=torch.uint8, device=DEV)
    iqn_ops.pack_tc(flat, ptc)
    return flat, ptc


def test_stage_accumulators_match_bf16_emulation(weights):
    rs = np.random.RandomState(0)
    B = 8
    x = (rs.randn(B, 26) * 3).astype(np.float32); x[:, 4:] *= (rs.rand(B, 22) > 0.5)
    taus = rs.rand(B, 32).astype(np.float32)
    flat, ptc = setup(weights)
    debug = torch.zeros(128 * (208 + 64 + 64 + 16), dtype=torch.float32, device=DEV)
    qm, gr = iqn_ops.act_tc(flat, ptc, torch.from_numpy(x).to(DEV), torch.from_numpy(taus).to(DEV), 1.0, want_qmean=True, debug=debug)
    d = debug.cpu().numpy()
    g1 = d[:128 * 208].reshape(128, 208); g2 = d[128 * 208:128 * 272].reshape(128, 64)
    g3 = d[128 * 272:128 * 336].reshape(128, 64); g4 = d[128 * 336:].reshape(128, 16)
    d1, d2, d3, d4, q = emulate(weights, x, taus)
    for name, got, want in (("D1", g1, d1[:128]), ("D2", g2, d2[:128]), ("D3", g3, d3[:128]), ("D4", g4[:, :9], d4[:128])):
        scale = max(1.0, np.abs(want).max())
        err = np.abs(got - want).max() / scale
        assert err < 2e-2, (name, err)          # each stage re-rounds its inputs to bf16: small drift accumulates
    assert np.abs(g4[:, 9:]).max() == 0.0        # padded output columns
    assert np.abs(qm.cpu().numpy() - q).max() < 2e-2 * max(1.0, np.abs(q).max())


@pytest.mark.parametrize("B", [1, 37, 4096])
def test_qmean_and_argmax_vs_fp32_kernel(weights, B):
    rs = np.random.RandomState(B)
    x = (rs.randn(B, 26) * 3).astype(np.float32); x[:, 4:] *= (rs.rand(B, 22) > 0.5)
    x[:, 2:4] = rs.uniform(-40, 40, size=(B, 2))
    taus = rs.rand(B, 32).astype(np.float32)
    cvar = rs.uniform(0.1, 1.0, size=B).astype(np.float32)
    flat, ptc = setup(weights)
    packed = torch.empty(iqn_ops.N_PACKED, dtype=torch.float32, device=DEV)
    iqn_ops.pack(flat, packed)
    xd, td, cd = torch.from_numpy(x).to(DEV), torch.from_numpy(taus).to(DEV), torch.from_numpy(cvar).to(DEV)
    for cv in (1.0, cd):
        qm, gr = iqn_ops.act_tc(flat, ptc, xd, td, cv, want_qmean=True)
        _, q32, g32 = iqn_ops.forward(flat, packed, xd, td, cv, want_quantiles=False, want_qmean=True, want_greedy=True)
        q32 = q32.cpu().numpy(); qm = qm.cpu().numpy()
        assert np.abs(qm - q32).max() < 3e-2 * max(1.0, np.abs(q32).max())
        top2 = np.sort(q32, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 5e-2 * np.maximum(1.0, np.abs(top2[:, 1]))
        assert np.array_equal(gr.cpu().numpy()[clear], g32.cpu().numpy()[clear])
        assert (gr.cpu().numpy() == g32.cpu().numpy()).mean() > 0.9
        assert np.array_equal(gr.cpu().numpy(), qm.argmax(axis=1))


def _philox(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    x, y, z, w = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * x, M1 * z
        x, y, z, w = (p1 >> 32) ^ y ^ k0, p1 & 0xffffffff, (p0 >> 32) ^ w ^ k1, p0 & 0xffffffff
        k0, k1 = (k0 + W0) & 0xffffffff, (k1 + W1) & 0xffffffff
    return x, y, z, w


def model_draws(seed, step, B):
    """The kernel's random streams (csrc/iqn_act_tc.cu act_draw): taus [B, 32], epsilon coin u [B], random action [B]."""
    key = (seed & 0xffffffff, seed >> 32)
    taus = np.zeros((B, 32), np.float32); coin = np.zeros(B, np.float32); rnd = np.zeros(B, np.int64)
    for b in range(B):
        for sub in range(8):
            r = _philox((b & 0xffffffff, (b >> 32) ^ (sub << 24), step & 0xffffffff, step >> 32), key)
            taus[b, 4 * sub:4 * sub + 4] = [np.float32(v >> 8) * np.float32(1.0 / 16777216.0) for v in r]
        r = _philox((b & 0xffffffff, (b >> 32) ^ (8 << 24), step & 0xffffffff, step >> 32), key)
        coin[b] = np.float32(r[0] >> 8) * np.float32(1.0 / 16777216.0)
        rnd[b] = (r[1] * 9) >> 32
    return taus, coin, rnd


def test_sampling_mode_follows_the_philox_model(weights):
    """iqn_act_tc_sample == iqn_act_tc fed with the model's taus; epsilon-greedy: greedy iff coin > eps (agent.py:200),
    else the model's random action; adaptive CVaR = adjust_cvar (agent.py:249-267) of every observation."""
    from distributional_rl_navigation_b200.iqn_agent import IQNAgent
    rs = np.random.RandomState(5)
    B, seed, step = 300, 0x1234ABCD5678, 77
    x = (rs.randn(B, 26) * 3).astype(np.float32); x[:, 4:] *= (rs.rand(B, 22) > 0.6); x[::7, 4:] = 0.0
    flat, ptc = setup(weights)
    xd = torch.from_numpy(x).to(DEV)
    taus, coin, rnd = model_draws(seed, step, B)
    assert 0.0 <= taus.min() and taus.max() < 1.0 and abs(taus.mean() - 0.5) < 0.01
    for adaptive in (False, True):
        cv_ref = np.array([IQNAgent.adjust_cvar(None, o) for o in x], np.float32) if adaptive else None
        cvar_arg = 0.7
        if adaptive:                                      # the pre-pass's own CVaR levels (fp32 norm; checked against adjust_cvar below)
            cvar_arg = torch.zeros(B, device=DEV)
            iqn_ops.act_tc_sample(flat, ptc, xd, 0.0, seed, step, adaptive=True, cvar_out=cvar_arg)
        qm_ref, gr_ref = iqn_ops.act_tc(flat, ptc, xd, torch.from_numpy(taus).to(DEV), cvar_arg, want_qmean=True)
        for eps in (0.0, 0.3, 1.0):
            cvar_out = torch.zeros(B, device=DEV)
            act, gr, qm = iqn_ops.act_tc_sample(flat, ptc, xd, eps, seed, step, cvar=0.7, adaptive=adaptive, want_greedy=True,
                                                want_qmean=True, cvar_out=cvar_out)
            if adaptive:
                np.testing.assert_allclose(cvar_out.cpu().numpy(), cv_ref, rtol=1e-6)
            assert torch.equal(gr, gr_ref) and torch.equal(qm, qm_ref)            # same taus -> bit-identical forward
            want = np.where(coin > eps, gr_ref.cpu().numpy(), rnd) if eps > 0 else gr_ref.cpu().numpy()
            np.testing.assert_array_equal(act.cpu().numpy(), want)
    a2, _, _ = iqn_ops.act_tc_sample(flat, ptc, xd, 1.0, seed, step + 1)
    assert not np.array_equal(a2.cpu().numpy(), rnd) and set(a2.cpu().numpy().tolist()) <= set(range(9))
