"""CPU: the C oracle (oracle/marinenav_oracle.c) against the committed golden vectors recorded from the reference.

Tolerances: the oracle keeps the reference's formulation, so floats agree to ~1e-13 teacher-forced; flags are exact.
"""
import json
import os

import numpy as np
import pytest

from oracle import marinenav_oracle as mo

ATOL = 1e-11


def _nohit(obs):
    """per-beam 'no return' pattern: both coordinates exactly zero (marinenav_env.py:314-315)"""
    s = obs[:, 4:].reshape(obs.shape[0], -1, 2)
    return (s[..., 0] == 0.0) & (s[..., 1] == 0.0)


def _soa(a):
    return np.ascontiguousarray(np.asarray(a, np.float64).T)


def _run_step_vectors(d, n_beams, n_threads=2):
    max_c, max_o = d["cores"].shape[1] // 3, d["obstacles"].shape[1] // 3
    for boundary in (0, 1):
        sel = np.where(d["set_boundary"] == boundary)[0]
        if sel.size == 0:
            continue
        p = mo.default_params(n_beams)
        p.set_boundary = boundary
        state, vel, goal = _soa(d["state"][sel]), _soa(d["velocity"][sel]), _soa(d["goal"][sel])
        cores, obst = _soa(d["cores"][sel]), _soa(d["obstacles"][sel])
        action = d["action"][sel].astype(np.int32)
        ep = d["episode_step"][sel].astype(np.int32)
        obs, reward, done, info = mo.step_batch(state, vel, goal, cores, obst, action, ep, p, n_threads)
        assert cores.shape[0] == 3 * max_c and obst.shape[0] == 3 * max_o
        np.testing.assert_array_equal(done, d["done"][sel])
        np.testing.assert_array_equal(info, d["info"][sel])
        np.testing.assert_allclose(obs[:, :4], d["obs"][sel][:, :4], rtol=0, atol=ATOL)
        # sonar entries: the reference's tan-slope quadratic is ill-conditioned near vertical beams (libm ulps get amplified)
        np.testing.assert_allclose(obs[:, 4:], d["obs"][sel][:, 4:], rtol=0, atol=1e-8)
        assert np.array_equal(_nohit(obs), _nohit(d["obs"][sel]))
        np.testing.assert_allclose(reward, d["reward"][sel], rtol=0, atol=ATOL)
        np.testing.assert_allclose(state.T, d["state_out"][sel], rtol=0, atol=ATOL)
        np.testing.assert_allclose(vel.T, d["velocity_out"][sel], rtol=0, atol=ATOL)
        np.testing.assert_array_equal(ep, d["episode_step"][sel] + 1)


def test_step_vectors(golden_dir):
    d = np.load(os.path.join(golden_dir, "step_vectors.npz"))
    assert set(np.unique(d["info"])) == {0, 1, 2, 3, 4}      # every termination kind is covered
    _run_step_vectors(d, 11)


def test_dense_vectors(golden_dir):
    _run_step_vectors(np.load(os.path.join(golden_dir, "dense_vectors.npz")), 64)


def test_observe_vectors(golden_dir):
    d = np.load(os.path.join(golden_dir, "observe_vectors.npz"))
    p = mo.default_params(11)
    obs = mo.observe_batch(_soa(d["state"]), _soa(d["velocity"]), _soa(d["goal"]), _soa(d["cores"]), _soa(d["obstacles"]), p)
    # the slope form loses digits next to vertical beams (tan ~ 1e3): both sides do, identically up to libm ulps
    np.testing.assert_allclose(obs, d["obs"], rtol=0, atol=1e-8)
    assert np.array_equal(_nohit(obs), _nohit(d["obs"]))


def test_reset_vectors_bit_exact(golden_dir):
    d = np.load(os.path.join(golden_dir, "reset_vectors.npz"))
    p = mo.default_params(11)
    out = mo.reset_batch(d["seed"], 4, 8, 30.0, 8, 10, p, n_threads=2)
    np.testing.assert_array_equal(out["cores"].T, d["cores"])          # MT19937 stream + rejection sampling: bit-exact
    np.testing.assert_array_equal(out["obstacles"].T, d["obstacles"])
    np.testing.assert_array_equal(out["goal"].T, d["goal"])
    np.testing.assert_array_equal(out["start_pose"][:2].T, d["start"])
    np.testing.assert_array_equal(out["state"].T, d["state"])
    np.testing.assert_array_equal(out["n_cores"], d["n_cores"])
    np.testing.assert_array_equal(out["n_obs"], d["n_obs"])
    np.testing.assert_allclose(out["velocity"].T, d["velocity"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["obs"], d["obs"], rtol=0, atol=ATOL)


def test_reset_stream_with_curriculum(golden_dir):
    d = np.load(os.path.join(golden_dir, "reset_vectors.npz"))
    sched = dict(timesteps=[0, 1000000, 2000000], num_cores=[4, 6, 8], num_obstacles=[6, 8, 10],
                 min_start_goal_dis=[30.0, 35.0, 40.0])
    for seed in range(8):
        env = mo.OracleEnv(seed=seed, schedule=sched)
        for k, tt in enumerate(d["stream_totals"]):
            env.e.total_timesteps = int(tt)
            obs = env.reset()
            e = env.e
            assert e.n_cores_placed == d["stream_n_cores"][seed, k] and e.n_obs_placed == d["stream_n_obs"][seed, k]
            got = np.array([e.x, e.y, e.theta, e.speed])
            np.testing.assert_array_equal(got, d["stream_state"][seed, k])
            cores = d["stream_cores"][seed, k]
            for c in range(e.n_cores_placed):
                assert (e.cores[c].x, e.cores[c].y) == (cores[c], cores[8 + c])
                assert (e.cores[c].Gamma if e.cores[c].clockwise else -e.cores[c].Gamma) == cores[16 + c]
            np.testing.assert_allclose(obs, d["stream_obs"][seed, k], rtol=0, atol=ATOL)


def test_eval_config_reset_kat(golden_dir):
    """create_eval_configs(MarineNavEnv(seed=348)) (train_IQN_model.py:123-148) regenerates the reference's eval_config.json."""
    cfg = json.load(open(os.path.join(golden_dir, "eval_config.json")))
    env = mo.OracleEnv(seed=348)
    env.e.reset_start_and_goal = 0
    env.e.start[0] = env.e.start[1] = 5.0
    env.e.goal[0] = env.e.goal[1] = 45.0
    k = 0
    for nc, no in ((4, 6), (6, 8), (8, 10)):
        for _ in range(10):
            env.e.num_cores, env.e.num_obs = nc, no
            env.reset()
            c, e = cfg[f"env_{k}"], env.e
            k += 1
            assert [[e.cores[j].x, e.cores[j].y] for j in range(e.n_cores_placed)] == c["env"]["cores"]["positions"]
            assert [e.cores[j].Gamma for j in range(e.n_cores_placed)] == c["env"]["cores"]["Gamma"]
            assert [e.cores[j].clockwise for j in range(e.n_cores_placed)] == c["env"]["cores"]["clockwise"]
            assert [[e.obstacles[j].x, e.obstacles[j].y] for j in range(e.n_obs_placed)] == c["env"]["obstacles"]["positions"]
            assert [e.obstacles[j].r for j in range(e.n_obs_placed)] == c["env"]["obstacles"]["r"]
            assert e.robot_init_theta == c["robot"]["init_theta"] and e.robot_init_speed == c["robot"]["init_speed"]


@pytest.mark.parametrize("name", ["greedy", "adaptive", "dqn"])
def test_recorded_episode_replay(golden_dir, name):
    """Free-running replay of the reference's recorded evaluation episodes (agent.py:345-357 definitions).
    A strided sample on CPU (every 10th evaluation x 30 maps = 900 episodes per file); the full 27 000 run on the GPU.
    The dynamics are chaotic in vortex-trapped episodes (SURVEY 8(c) chaos probe), so returns get a loose tolerance
    and a tiny number of flag flips is tolerated and counted."""
    cfg = json.load(open(os.path.join(golden_dir, "eval_config.json")))
    d = np.load(os.path.join(golden_dir, f"episodes_{name}.npz"))
    lengths = d["lengths"]
    offs = np.concatenate([[0], np.cumsum(lengths.ravel())])
    env = mo.OracleEnv(seed=0)
    bad_flag, n, worst = 0, 0, 0.0
    for ev in range(0, lengths.shape[0], 10):
        for m in range(30):
            k = ev * 30 + m
            acts = d["actions_flat"][offs[k]:offs[k + 1]]
            env.reset_with_eval_config(cfg[f"env_{m}"])
            ret, info, done = 0.0, {"state": "normal"}, False
            for t, a in enumerate(acts):
                assert not done
                _, r, done, info = env.step(int(a))
                ret += 0.99 ** t * r
            ok = (info["state"] == "reach goal") == bool(d["successes"][ev, m])
            ok = ok and abs(0.1 * 10 * len(acts) - d["times"][ev, m]) < 1e-9
            ok = ok and (done or len(acts) == 1000)
            bad_flag += not ok
            if ok:
                worst = max(worst, abs(ret - d["rewards"][ev, m]))
            n += 1
    assert n == 900
    assert bad_flag <= 1, bad_flag
    assert worst < 5e-3, worst
