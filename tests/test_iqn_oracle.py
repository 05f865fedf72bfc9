"""CPU: the numpy IQN oracle (oracle/iqn_oracle.py) against fixtures recorded from the reference's PyTorch path
(tests/golden/make_golden_iqn.py): forward, get_qvals, adjust_cvar, loss, every gradient, clipped Adam steps."""
import os

import numpy as np
import pytest

from oracle import iqn_oracle as io


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "iqn_kat.npz"))


@pytest.fixture(scope="module")
def weights(golden_dir):
    w = np.load(os.path.join(golden_dir, "iqn_weights.npz"))
    return {k: w[k] for k in w.files}


def test_param_layout(weights):
    assert [n for n, _ in io.PARAM_SPECS] == list(weights.keys())
    assert io.N_PARAMS == 35785 == sum(v.size for v in weights.values())
    flat = io.flatten(weights)
    back = io.unflatten(flat)
    assert all(np.array_equal(back[k], weights[k]) for k in weights)


@pytest.mark.parametrize("K", [8, 32])
@pytest.mark.parametrize("cvar", [1.0, 0.37])
def test_forward_matches_reference(kat, weights, K, cvar):
    tag = f"K{K}_cvar{str(cvar).replace('.', 'p')}"
    q = io.forward(weights, kat["fwd_x"], kat[f"fwd_taus_{tag}"], cvar)
    ref = kat[f"fwd_q_{tag}"]
    assert np.abs(q - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    qv = io.get_qvals(weights, kat["fwd_x"], kat[f"fwd_taus_{tag}"], cvar)
    np.testing.assert_allclose(qv, ref.mean(axis=1), rtol=1e-5, atol=1e-4)


def test_adjust_cvar(kat):
    got = np.array([io.adjust_cvar(o) for o in kat["cvar_obs"]])
    np.testing.assert_allclose(got, kat["cvar_val"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("B", [32, 1024])
def test_train_loss_grad_adam(kat, weights, B):
    P_local = {k: v.copy() for k, v in weights.items()}
    P_target = io.unflatten(kat[f"target_flat_B{B}"])
    batch = [kat[f"{n}_B{B}"] for n in ("states", "actions", "rewards", "next_states", "dones")]
    taus = kat[f"taus_B{B}"]
    flat = io.flatten(P_local)
    m = np.zeros_like(flat); v = np.zeros_like(flat)
    for k in range(3):
        loss, grad = io.loss_and_grad(io.unflatten(flat), P_target, batch[0], batch[1], batch[2], batch[3], batch[4],
                                      taus_target=taus[2 * k], taus_local=taus[2 * k + 1])
        ref_loss = kat[f"losses_B{B}"][k]
        assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss), (loss, ref_loss)          # north_star asks 1e-4
        if k == 0:
            ref_g = kat[f"grad_B{B}"]
            assert np.abs(grad - ref_g).max() <= 2e-5 * np.abs(ref_g).max()
        flat, m, v, total = io.clip_adam(flat, grad, m, v, step=k + 1)
        if k == 0:
            assert abs(total - kat[f"gradnorm_B{B}"]) <= 1e-5 * kat[f"gradnorm_B{B}"]
            np.testing.assert_allclose(flat, kat[f"params_after1_B{B}"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(flat, kat[f"params_after3_B{B}"], rtol=0, atol=5e-7)
    np.testing.assert_allclose(m, kat[f"adam_m_after3_B{B}"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(v, kat[f"adam_v_after3_B{B}"], rtol=1e-4, atol=1e-12)


def test_loss_kat_values(kat):
    """the reference's train() under torch.manual_seed(1234) on the deterministic KAT batch."""
    assert abs(float(kat["kat_loss_seed1234_B32"]) - 266.52407837) < 1e-3
    assert abs(float(kat["kat_loss_seed1234_B1024"]) - 307.13052368) < 1e-3
