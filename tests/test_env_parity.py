"""GPU parity: the CUDA env kernels (through the C-ABI) against the reference's golden vectors and the CPU oracle.

Tolerances (BASELINE.json north_star): done / info flags bit-exact; float observations and rewards within 1e-5
(|a-b| <= 1e-5 * max(1, |b|); observations leave the kernel as float32).  fp64 state within 1e-11.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from distributional_rl_navigation_b200 import _lib, env_ops  # noqa: E402
from oracle import marinenav_oracle as mo  # noqa: E402

DEV = "cuda:0"
TOL = 1e-5


def close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, f"max scaled err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


def nohit(obs):
    s = np.asarray(obs)[:, 4:].reshape(obs.shape[0], -1, 2)
    return (np.abs(s[..., 0]) < 1e-12) & (np.abs(s[..., 1]) < 1e-12)


def soa(a):
    return np.ascontiguousarray(np.asarray(a, np.float64).T)


def to_buf(state, velocity, goal, cores, obstacles, n_beams, action=None, episode_step=None):
    """numpy SoA arrays ([k][E]) -> device buffers."""
    E = state.shape[1]
    buf = env_ops.alloc_env_buffers(E, cores.shape[0] // 3, obstacles.shape[0] // 3, n_beams, DEV)
    for k, v in (("state", state), ("velocity", velocity), ("goal", goal), ("cores", cores), ("obstacles", obstacles)):
        buf[k].copy_(torch.from_numpy(np.ascontiguousarray(v)))
    if action is not None:
        buf["action"].copy_(torch.from_numpy(np.asarray(action, np.int32)))
    if episode_step is not None:
        buf["episode_step"].copy_(torch.from_numpy(np.asarray(episode_step, np.int32)))
    return buf


def mnv_params_like(orc_p):
    p = _lib.default_params(orc_p.n_beams)
    for f, _ in _lib.MnvParams._fields_:
        if not hasattr(orc_p, f):                              # launch-mode fields have no counterpart in the oracle
            continue
        v = getattr(orc_p, f)
        if hasattr(v, "__len__"):
            for i in range(len(v)):
                getattr(p, f)[i] = v[i]
        else:
            setattr(p, f, v)
    return p


def run_step_vectors(d, n_beams, pdl_prefetch=0):
    for boundary in (0, 1):
        sel = np.where(d["set_boundary"] == boundary)[0]
        if sel.size == 0:
            continue
        p = _lib.default_params(n_beams)
        p.set_boundary = boundary
        p.pdl_prefetch = pdl_prefetch
        buf = to_buf(soa(d["state"][sel]), soa(d["velocity"][sel]), soa(d["goal"][sel]), soa(d["cores"][sel]),
                     soa(d["obstacles"][sel]), n_beams, d["action"][sel], d["episode_step"][sel])
        env_ops.step(buf, p)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(buf["done"].cpu().numpy(), d["done"][sel])
        np.testing.assert_array_equal(buf["info"].cpu().numpy(), d["info"][sel])
        obs = buf["obs"].cpu().numpy()
        close(obs, d["obs"][sel])
        assert np.array_equal(nohit(obs), nohit(d["obs"][sel]))
        close(buf["reward"].cpu().numpy(), d["reward"][sel])
        np.testing.assert_allclose(buf["state"].cpu().numpy().T, d["state_out"][sel], rtol=0, atol=1e-11)
        np.testing.assert_allclose(buf["velocity"].cpu().numpy().T, d["velocity_out"][sel], rtol=0, atol=1e-11)
        np.testing.assert_array_equal(buf["episode_step"].cpu().numpy(), d["episode_step"][sel] + 1)


def test_step_golden_vectors(golden_dir):
    run_step_vectors(np.load(os.path.join(golden_dir, "step_vectors.npz")), 11)


@pytest.mark.parametrize("opts", [dict(tma=1, pdl=1), dict(tma=0, pdl=0), dict(tma=1, pdl=0)])
def test_step_golden_vectors_every_kernel_option(golden_dir, opts):
    """mnv_set_option switches (TMA bulk-copy staging of the obstacle rows, programmatic dependent launch) never change a result."""
    L = _lib.load()
    old = {k: L.mnv_get_option(k.encode()) for k in opts}
    try:
        for k, v in opts.items():
            assert L.mnv_set_option(k.encode(), v) == 0 and L.mnv_get_option(k.encode()) == v
        run_step_vectors(np.load(os.path.join(golden_dir, "step_vectors.npz")), 11)
    finally:
        for k, v in old.items():
            L.mnv_set_option(k.encode(), v)


def test_step_golden_vectors_per_call_pdl_prefetch(golden_dir):
    """mnv_params.pdl_prefetch (the per-call launch mode VecMarineNavEnv uses) changes no result and no process-wide switch."""
    L = _lib.load()
    before = L.mnv_get_option(b"pdl")
    run_step_vectors(np.load(os.path.join(golden_dir, "step_vectors.npz")), 11, pdl_prefetch=1)
    assert L.mnv_get_option(b"pdl") == before


def test_step_dense_golden_vectors(golden_dir):
    run_step_vectors(np.load(os.path.join(golden_dir, "dense_vectors.npz")), 64)


def test_observe_golden_vectors(golden_dir):
    d = np.load(os.path.join(golden_dir, "observe_vectors.npz"))
    p = _lib.default_params(11)
    buf = to_buf(soa(d["state"]), soa(d["velocity"]), soa(d["goal"]), soa(d["cores"]), soa(d["obstacles"]), 11)
    env_ops.observe(buf, p)
    obs = buf["obs"].cpu().numpy()
    close(obs, d["obs"])
    assert np.array_equal(nohit(obs), nohit(d["obs"]))


def test_observe_masked_leaves_other_rows(golden_dir):
    d = np.load(os.path.join(golden_dir, "observe_vectors.npz"))
    p = _lib.default_params(11)
    buf = to_buf(soa(d["state"]), soa(d["velocity"]), soa(d["goal"]), soa(d["cores"]), soa(d["obstacles"]), 11)
    buf["obs"].fill_(7.0)
    mask = torch.from_numpy((np.arange(d["state"].shape[0]) % 3 == 0).astype(np.uint8)).to(DEV)
    env_ops.observe(buf, p, mask=mask)
    obs = buf["obs"].cpu().numpy()
    m = mask.cpu().numpy().astype(bool)
    close(obs[m], d["obs"][m])
    assert np.all(obs[~m] == 7.0)


@pytest.mark.parametrize("E,nc,no,n_beams,max_c,max_o", [
    (1, 4, 8, 11, 4, 8), (63, 4, 8, 11, 4, 8), (4097, 8, 10, 11, 8, 10), (65536, 4, 8, 11, 4, 8),
    (3000, 0, 0, 11, 4, 8), (2048, 4, 16, 21, 8, 16), (16384, 4, 32, 64, 4, 32), (1023, 8, 20, 33, 8, 24), (5, 0, 3, 7, 1, 17),
    (700, 8, 16, 128, 8, 16), (600, 8, 32, 128, 8, 32)])     # compiled capacities: 8 cores, 16 / 32 obstacles, 128 beams
def test_step_vs_oracle_teacher_forced(E, nc, no, n_beams, max_c, max_o):
    """Seeded maps from the oracle's reset, random actions, every step compared from identical pre-states."""
    op = mo.default_params(n_beams)
    p = mnv_params_like(op)
    seeds = np.arange(E, dtype=np.uint32) + 1000
    w = mo.reset_batch(seeds, nc, no, 30.0, max_c, max_o, op, n_threads=8)
    rng = np.random.RandomState(E)
    state, vel, ep = w["state"], w["velocity"], np.zeros(E, np.int32)
    buf = to_buf(state, vel, w["goal"], w["cores"], w["obstacles"], n_beams)
    n_steps = 40 if E <= 5000 else 6
    flags = {k: 0 for k in range(5)}
    for t in range(n_steps):
        action = rng.randint(0, 9, size=E).astype(np.int32)
        if t == 3:
            ep[::7] = 1000                                           # timeout priority (Q5)
        buf["state"].copy_(torch.from_numpy(state)); buf["velocity"].copy_(torch.from_numpy(vel))
        buf["episode_step"].copy_(torch.from_numpy(ep)); buf["action"].copy_(torch.from_numpy(action))
        env_ops.step(buf, p)
        obs, reward, done, info = mo.step_batch(state, vel, w["goal"], w["cores"], w["obstacles"], action, ep, op, n_threads=8)
        np.testing.assert_array_equal(buf["done"].cpu().numpy(), done)
        np.testing.assert_array_equal(buf["info"].cpu().numpy(), info)
        g_obs = buf["obs"].cpu().numpy()
        close(g_obs, obs)
        assert np.array_equal(nohit(g_obs), nohit(obs))
        close(buf["reward"].cpu().numpy(), reward)
        np.testing.assert_allclose(buf["state"].cpu().numpy(), state, rtol=0, atol=1e-11)
        np.testing.assert_array_equal(buf["episode_step"].cpu().numpy(), ep)
        for k in range(5):
            flags[k] += int((info == k).sum())
        ep[done != 0] = 0     # keep going from the oracle's post-state (episodes that ended simply continue)
    assert flags[0] > 0


def test_full_size_permutation_property():
    """65 536 envs: stepping a permuted batch gives the permuted result (no cross-env coupling, any CTA placement)."""
    E, n_beams = 65536, 11
    op = mo.default_params(n_beams)
    p = mnv_params_like(op)
    w = mo.reset_batch(np.arange(E, dtype=np.uint32), 4, 8, 30.0, 4, 8, op, n_threads=8)
    rng = np.random.RandomState(5)
    action = rng.randint(0, 9, size=E).astype(np.int32)
    perm = rng.permutation(E)
    a = to_buf(w["state"], w["velocity"], w["goal"], w["cores"], w["obstacles"], n_beams, action)
    b = to_buf(w["state"][:, perm], w["velocity"][:, perm], w["goal"][:, perm], w["cores"][:, perm], w["obstacles"][:, perm],
               n_beams, action[perm])
    for _ in range(5):
        env_ops.step(a, p); env_ops.step(b, p)
    pt = torch.from_numpy(perm).to(DEV)
    assert torch.equal(a["obs"][pt], b["obs"])
    assert torch.equal(a["state"][:, pt], b["state"])
    assert torch.equal(a["reward"][pt], b["reward"]) and torch.equal(a["done"][pt], b["done"])


def test_reset_bit_exact_vs_reference_vectors(golden_dir):
    """mnv_seed + mnv_reset reproduce MarineNavEnv(seed).reset() of the reference bit-for-bit (MT19937 draw order,
    check_core / check_obstacle), then mnv_observe gives the first observation."""
    d = np.load(os.path.join(golden_dir, "reset_vectors.npz"))
    E = d["seed"].shape[0]
    buf = env_ops.alloc_env_buffers(E, 8, 10, 11, DEV)
    key = torch.zeros(E, 624, dtype=torch.int32, device=DEV); pos = torch.zeros(E, dtype=torch.int32, device=DEV)
    env_ops.seed(key, pos, torch.from_numpy(d["seed"].astype(np.uint32).view(np.int32)).to(DEV))
    rp = _lib.default_reset_params()
    rp.num_cores, rp.num_obs, rp.min_start_goal_dis = 4, 8, 30.0
    env_ops.reset(buf, key, pos, rp)
    p = _lib.default_params(11)
    env_ops.observe(buf, p, velocity_from_state=True)
    np.testing.assert_array_equal(buf["cores"].cpu().numpy().T, d["cores"])
    np.testing.assert_array_equal(buf["obstacles"].cpu().numpy().T, d["obstacles"])
    np.testing.assert_array_equal(buf["goal"].cpu().numpy().T, d["goal"])
    np.testing.assert_array_equal(buf["state"].cpu().numpy().T, d["state"])
    np.testing.assert_array_equal(buf["start_pose"].cpu().numpy()[:2].T, d["start"])
    np.testing.assert_array_equal(buf["n_placed"].cpu().numpy()[0], d["n_cores"])
    np.testing.assert_array_equal(buf["n_placed"].cpu().numpy()[1], d["n_obs"])
    np.testing.assert_allclose(buf["velocity"].cpu().numpy().T, d["velocity"], rtol=0, atol=1e-11)
    close(buf["obs"].cpu().numpy(), d["obs"])


def test_reset_eval_config_kat(golden_dir):
    """create_eval_configs(MarineNavEnv(seed=348)) (train_IQN_model.py:123-148): 30 consecutive resets of ONE stream
    regenerate the reference's eval_config.json bit-for-bit on the device."""
    cfg = json.load(open(os.path.join(golden_dir, "eval_config.json")))
    buf = env_ops.alloc_env_buffers(1, 8, 10, 11, DEV)
    key = torch.zeros(1, 624, dtype=torch.int32, device=DEV); pos = torch.zeros(1, dtype=torch.int32, device=DEV)
    env_ops.seed(key, pos, torch.tensor([348], dtype=torch.int32, device=DEV))
    rp = _lib.default_reset_params()
    rp.reset_start_and_goal = 0
    rp.start[0] = rp.start[1] = 5.0
    rp.goal[0] = rp.goal[1] = 45.0
    k = 0
    for nc, no in ((4, 6), (6, 8), (8, 10)):
        for _ in range(10):
            rp.num_cores, rp.num_obs = nc, no
            env_ops.reset(buf, key, pos, rp)
            c = cfg[f"env_{k}"]; k += 1
            cores = buf["cores"].cpu().numpy()[:, 0]; obst = buf["obstacles"].cpu().numpy()[:, 0]
            st = buf["state"].cpu().numpy()[:, 0]
            assert [[cores[i], cores[8 + i]] for i in range(nc)] == c["env"]["cores"]["positions"]
            assert [abs(cores[16 + i]) for i in range(nc)] == c["env"]["cores"]["Gamma"]
            assert [int(cores[16 + i] > 0) for i in range(nc)] == c["env"]["cores"]["clockwise"]
            assert [[obst[i], obst[10 + i]] for i in range(no)] == c["env"]["obstacles"]["positions"]
            assert [obst[20 + i] for i in range(no)] == c["env"]["obstacles"]["r"]
            assert st[2] == c["robot"]["init_theta"] and st[3] == c["robot"]["init_speed"]
            assert list(st[:2]) == c["env"]["start"]


def test_reset_curriculum_streams_bit_exact_on_device(golden_dir):
    """The reference's reset() streams with the training curriculum (train_IQN_model.py:86-90), recorded from the reference
    in reset_vectors.npz `stream_*`: 8 seeds x 6 consecutive resets at total_timesteps 0 ... 2 999 999, i.e. all three
    stages INCLUDING stages 1-2 with RANDOM start / goal at min_start_goal_dis 35 / 40 and 6-8 cores / 8-10 obstacles.
    mnv_reset on the device, one stream per environment, must reproduce every map, start pose and goal bit for bit."""
    d = np.load(os.path.join(golden_dir, "reset_vectors.npz"))
    sched = dict(timesteps=[0, 1000000, 2000000], num_cores=[4, 6, 8], num_obstacles=[6, 8, 10], min_start_goal_dis=[30.0, 35.0, 40.0])
    S = d["stream_state"].shape[0]
    buf = env_ops.alloc_env_buffers(S, 8, 10, 11, DEV)
    key = torch.zeros(S, 624, dtype=torch.int32, device=DEV); pos = torch.zeros(S, dtype=torch.int32, device=DEV)
    env_ops.seed(key, pos, torch.arange(S, dtype=torch.int32, device=DEV))
    p = _lib.default_params(11)
    stages = set()
    for k, tt in enumerate(d["stream_totals"]):
        idx = int((np.array(sched["timesteps"]) - int(tt) <= 0).sum()) - 1            # marinenav_env.py:89-98
        stages.add(idx)
        rp = _lib.default_reset_params()
        rp.num_cores, rp.num_obs, rp.min_start_goal_dis = sched["num_cores"][idx], sched["num_obstacles"][idx], sched["min_start_goal_dis"][idx]
        env_ops.reset(buf, key, pos, rp)
        env_ops.observe(buf, p, velocity_from_state=True)
        np.testing.assert_array_equal(buf["state"].cpu().numpy().T, d["stream_state"][:, k])
        np.testing.assert_array_equal(buf["goal"].cpu().numpy().T, d["stream_goal"][:, k])
        np.testing.assert_array_equal(buf["cores"].cpu().numpy().T, d["stream_cores"][:, k])
        np.testing.assert_array_equal(buf["obstacles"].cpu().numpy().T, d["stream_obstacles"][:, k])
        np.testing.assert_array_equal(buf["n_placed"].cpu().numpy()[0], d["stream_n_cores"][:, k])
        np.testing.assert_array_equal(buf["n_placed"].cpu().numpy()[1], d["stream_n_obs"][:, k])
        close(buf["obs"].cpu().numpy(), d["stream_obs"][:, k])
    assert stages == {0, 1, 2}


def test_reset_masked_and_stream_continuation():
    """Masked reset only touches the selected environments and continues each stream like consecutive reset() calls."""
    E = 512
    op = mo.default_params(11)
    buf = env_ops.alloc_env_buffers(E, 4, 8, 11, DEV)
    key = torch.zeros(E, 624, dtype=torch.int32, device=DEV); pos = torch.zeros(E, dtype=torch.int32, device=DEV)
    seeds = np.arange(E, dtype=np.uint32) + 77
    env_ops.seed(key, pos, torch.from_numpy(seeds.view(np.int32)).to(DEV))
    rp = _lib.default_reset_params()
    rp.num_cores, rp.num_obs, rp.min_start_goal_dis = 4, 8, 30.0
    env_ops.reset(buf, key, pos, rp)
    first = {k: v.clone() for k, v in buf.items() if torch.is_tensor(v)}
    mask = torch.from_numpy((np.arange(E) % 5 == 0).astype(np.uint8)).to(DEV)
    env_ops.reset(buf, key, pos, rp, mask=mask)
    m = mask.bool()
    assert torch.equal(buf["cores"][:, ~m], first["cores"][:, ~m]) and torch.equal(buf["state"][:, ~m], first["state"][:, ~m])
    # second reset of the masked envs == second reset() of the oracle's stateful env
    for e in np.where(m.cpu().numpy())[0][:40]:
        env = mo.OracleEnv(seed=int(seeds[e]))
        env.e.num_cores, env.e.num_obs, env.e.min_start_goal_dis = 4, 8, 30.0
        env.reset(); env.reset()
        got = buf["cores"][:, e].cpu().numpy()
        want = [env.e.cores[i].x for i in range(4)] + [env.e.cores[i].y for i in range(4)] + \
               [env.e.cores[i].Gamma if env.e.cores[i].clockwise else -env.e.cores[i].Gamma for i in range(4)]
        assert list(got) == want
        assert list(buf["state"][:, e].cpu().numpy()) == [env.e.x, env.e.y, env.e.theta, env.e.speed]
    assert op.n_beams == 11


@pytest.mark.parametrize("name", ["greedy", "adaptive", "dqn"])
def test_recorded_episodes_full_replay(golden_dir, name):
    """All 9 000 recorded evaluation episodes of one file replayed as ONE batch (free-running, <= 1000 launches):
    success flags / episode times / discounted returns against the reference's recorded values (agent.py:345-357).
    Vortex-trapped episodes are chaotic: a 1e-16 rounding difference grows to metres within ~800 steps,
    so a handful of flag flips are tolerated and counted; everything else must match."""
    cfg = json.load(open(os.path.join(golden_dir, "eval_config.json")))
    d = np.load(os.path.join(golden_dir, f"episodes_{name}.npz"))
    lengths = d["lengths"]                                     # [300, 30]
    n_ev, n_map = lengths.shape
    E = n_ev * n_map
    cfgs = [cfg[f"env_{m}"] for m in range(n_map)]
    st, goal, cores, obst = mo.tables_from_eval_config(cfgs, 8, 10)
    tile = lambda a: np.ascontiguousarray(np.tile(a, (1, n_ev)))   # env index = ev * 30 + map
    p = _lib.default_params(11)
    buf = to_buf(tile(st), np.zeros((2, E)), tile(goal), tile(cores), tile(obst), 11)
    # padded action matrix [T][E]
    T = int(lengths.max())
    flat, L = d["actions_flat"], lengths.ravel()
    offs = np.concatenate([[0], np.cumsum(L)])
    acts = np.zeros((T, E), np.int32)
    for k in range(E):
        acts[:L[k], k] = flat[offs[k]:offs[k + 1]]
    acts_d = torch.from_numpy(acts).to(DEV)
    L_d = torch.from_numpy(L.astype(np.int64)).to(DEV)
    ret = torch.zeros(E, dtype=torch.float64, device=DEV)
    last_info = torch.zeros(E, dtype=torch.uint8, device=DEV)
    early_done = torch.zeros(E, dtype=torch.bool, device=DEV)
    finished = torch.zeros(E, dtype=torch.bool, device=DEV)
    for t in range(T):
        env_ops.step(buf, p, action=acts_d[t])
        active = L_d > t
        early_done |= active & finished                         # stepped again after done: the recording would not
        ret += torch.where(active, (0.99 ** t) * buf["reward"].double(), torch.zeros_like(ret))
        last_info = torch.where(active, buf["info"], last_info)
        finished |= active & (buf["done"] != 0)
    torch.cuda.synchronize()
    success = (last_info == 3).cpu().numpy().reshape(n_ev, n_map)
    ended = (finished.cpu().numpy() | (L == 1000)).reshape(n_ev, n_map)
    bad = (success != d["successes"]) | ~ended | early_done.cpu().numpy().reshape(n_ev, n_map)
    n_bad = int(bad.sum())
    err = np.abs(ret.cpu().numpy().reshape(n_ev, n_map) - d["rewards"])
    ok_err = err[~bad]
    print(f"{name}: {E} episodes, {int(L.sum())} steps, flag flips {n_bad}, return err max {ok_err.max():.3e}, "
          f"frac > 1e-3: {(ok_err > 1e-3).mean():.2e}")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):                                  # keep the observed counts (copied to profiles/ per round)
        with open(os.path.join(out_dir, "free_run_replay_counts.txt"), "a") as f:
            f.write(f"{name}: episodes {E} steps {int(L.sum())} flag_flips {n_bad} return_err_max {ok_err.max():.3e} "
                    f"frac_err_gt_1e-3 {(ok_err > 1e-3).mean():.3e} frac_err_gt_1e-2 {(ok_err > 1e-2).mean():.3e}\n")
    np.testing.assert_allclose(0.1 * 10 * lengths, d["times"], rtol=0, atol=1e-9)
    # observed (profiles/r2_free_run_replay_counts.txt, step kernel v8): greedy 9, adaptive 12, dqn 3 flips of 9 000 -- the band is
    # 1.5x the worst observed count (chaotic vortex-trapped episodes; the CPU oracle, same operation order as the reference: 0)
    assert n_bad <= 20, n_bad
    assert (ok_err > 1e-2).mean() <= 2e-3
