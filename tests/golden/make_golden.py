"""Generates the committed golden fixtures in this directory from the UNMODIFIED reference
(/root/reference, RobustFieldAutonomyLab/Distributional_RL_Navigation @ e77bbbf) imported in-process.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py [env|iqn|all]

Fixtures (all small, committed; the GPU box has no /root/reference):
  eval_config.json            the reference's own 30 evaluation maps (pretrained_models/IQN/seed_3/eval_config.json)
  episodes_{greedy,adaptive,dqn}.npz   the reference's 27 000 recorded evaluation episodes, repacked (ragged object
                              arrays -> flat u8 actions + lengths) with their recorded returns / successes / times
  step_vectors.npz            teacher-forced MarineNavEnv.step I/O recorded from reference rollouts (several map sizes)
  observe_vectors.npz         MarineNavEnv.get_observation on crafted states (vertical-beam snap Q10, robot inside /
                              abutting obstacles, ordered-break Q3 stress)
  dense_vectors.npz           32 obstacles / 64 beams step I/O (BASELINE config 5 shape)
  reset_vectors.npz           MarineNavEnv(seed).reset() maps + first observation for seeds 0..255 and multi-reset streams
  iqn_*.npz                   see make_golden_iqn() below
"""
import contextlib
import io
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402

REF = ref_import.REF_ROOT
MAX_C, MAX_O = 8, 10


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def tables(env, max_c, max_o):
    cores = np.zeros(3 * max_c); obst = np.zeros(3 * max_o)
    for c, core in enumerate(env.cores):
        cores[c], cores[max_c + c] = core.x, core.y
        cores[2 * max_c + c] = core.Gamma if core.clockwise else -core.Gamma
    for o, ob in enumerate(env.obstacles):
        obst[o], obst[max_o + o], obst[2 * max_o + o] = ob.x, ob.y, ob.r
    return cores, obst


def robot_state(env):
    r = env.robot
    return np.array([r.x, r.y, r.theta, r.speed]), np.array(r.velocity, dtype=np.float64)


INFO = {"normal": 0, "too long episode": 1, "collision": 2, "reach goal": 3, "out of boundary": 4}


def record_steps(ref_env, configs, n_per, max_c, max_o, num_beams=11, seed0=0):
    rec = {k: [] for k in ("state", "velocity", "goal", "cores", "obstacles", "action", "episode_step", "set_boundary",
                           "state_out", "velocity_out", "obs", "reward", "done", "info")}
    rng = np.random.RandomState(1234 + seed0)
    for ci, (nc, no, boundary) in enumerate(configs):
        with quiet():
            env = ref_env.MarineNavEnv(seed=seed0 + ci)
            env.num_cores, env.num_obs, env.set_boundary = nc, no, boundary
            if num_beams != 11:
                env.robot.sonar.num_beams = num_beams
                env.robot.sonar.compute_phi(); env.robot.sonar.compute_beam_angles()
            env.reset()
        n = 0
        while n < n_per:
            # biased-forward random policy so that episodes also reach the goal / collide, not only time out
            a = int(rng.randint(9)) if rng.rand() < 0.7 else int(rng.choice([6, 7, 8]))
            if rng.rand() < 0.02:
                env.episode_timesteps = 999 + int(rng.randint(3))      # exercise the Q5 timeout priority
            s, v = robot_state(env)
            c, o = tables(env, max_c, max_o)
            rec["state"].append(s); rec["velocity"].append(v); rec["goal"].append(np.array(env.goal, dtype=np.float64))
            rec["cores"].append(c); rec["obstacles"].append(o); rec["action"].append(a)
            rec["episode_step"].append(env.episode_timesteps); rec["set_boundary"].append(int(boundary))
            obs, r, d, info = env.step(a)
            s2, v2 = robot_state(env)
            rec["state_out"].append(s2); rec["velocity_out"].append(v2); rec["obs"].append(np.asarray(obs, np.float64))
            rec["reward"].append(r); rec["done"].append(int(d)); rec["info"].append(INFO[info["state"]])
            n += 1
            if d:
                with quiet():
                    env.reset()
    return {k: np.asarray(v) for k, v in rec.items()}


def record_episode_tails(ref_env, pm, evals, tail):
    cfg = json.load(open(os.path.join(pm, "IQN", "seed_3", "eval_config.json")))
    d = np.load(os.path.join(pm, "IQN", "seed_3", "greedy_evaluations.npz"), allow_pickle=True)
    rec = {k: [] for k in ("state", "velocity", "goal", "cores", "obstacles", "action", "episode_step", "set_boundary",
                           "state_out", "velocity_out", "obs", "reward", "done", "info")}
    with quiet():
        env = ref_env.MarineNavEnv(seed=0)
    for ev in evals:
        for m in range(30):
            acts = d["actions"][ev, m]
            env.reset_with_eval_config(cfg[f"env_{m}"])
            for t, a in enumerate(acts):
                keep = t >= len(acts) - tail
                if keep:
                    s, v = robot_state(env); c, o = tables(env, MAX_C, MAX_O)
                    rec["state"].append(s); rec["velocity"].append(v); rec["goal"].append(np.array(env.goal, dtype=np.float64))
                    rec["cores"].append(c); rec["obstacles"].append(o); rec["action"].append(int(a))
                    rec["episode_step"].append(env.episode_timesteps); rec["set_boundary"].append(0)
                obs, r, dn, info = env.step(int(a))
                if keep:
                    s2, v2 = robot_state(env)
                    rec["state_out"].append(s2); rec["velocity_out"].append(v2); rec["obs"].append(np.asarray(obs, np.float64))
                    rec["reward"].append(r); rec["done"].append(int(dn)); rec["info"].append(INFO[info["state"]])
    return {k: np.asarray(v) for k, v in rec.items()}


def make_env_goldens():
    ref_env, _, _ = ref_import.load_reference()
    pm = os.path.join(REF, "pretrained_models")
    shutil.copyfile(os.path.join(pm, "IQN", "seed_3", "eval_config.json"), os.path.join(HERE, "eval_config.json"))

    for name, path in (("greedy", "IQN/seed_3/greedy_evaluations.npz"), ("adaptive", "IQN/seed_3/adaptive_evaluations.npz"),
                       ("dqn", "DQN/seed_3/evaluations.npz")):
        d = np.load(os.path.join(pm, path), allow_pickle=True)
        acts = d["actions"]
        lengths = np.array([[len(acts[i, j]) for j in range(acts.shape[1])] for i in range(acts.shape[0])], np.int32)
        flat = np.concatenate([np.asarray(acts[i, j], np.uint8) for i in range(acts.shape[0]) for j in range(acts.shape[1])])
        np.savez_compressed(os.path.join(HERE, f"episodes_{name}.npz"), actions_flat=flat, lengths=lengths,
                            rewards=d["rewards"], successes=d["successes"], times=d["times"], energies=d["energies"],
                            timesteps=d["timesteps"])
        print(name, flat.shape, lengths.sum())

    cfgs = [(4, 8, False), (8, 10, False), (6, 8, False), (4, 6, False), (0, 0, False), (0, 5, False), (3, 0, False),
            (4, 8, True), (1, 1, False), (8, 10, True)]
    steps = record_steps(ref_env, cfgs, 400, MAX_C, MAX_O)
    # tails of recorded evaluation episodes replayed on the reference (reach-goal / collision terminations)
    tails = record_episode_tails(ref_env, pm, evals=(100, 200, 299), tail=8)
    np.savez_compressed(os.path.join(HERE, "step_vectors.npz"),
                        **{k: np.concatenate([steps[k], tails[k]]) for k in steps})
    np.savez_compressed(os.path.join(HERE, "dense_vectors.npz"),
                        **record_steps(ref_env, [(4, 32, False), (4, 32, False)], 150, 4, 32, num_beams=64, seed0=50))

    # ---- crafted get_observation cases -----------------------------------------------------------------------
    rng = np.random.RandomState(99)
    rec = {k: [] for k in ("state", "velocity", "goal", "cores", "obstacles", "obs")}
    with quiet():
        env = ref_env.MarineNavEnv(seed=5)
        env.num_cores, env.num_obs = 4, 10
    beam_angles = np.array(env.robot.sonar.beam_angles)
    for it in range(1200):
        if it % 40 == 0:
            with quiet():
                env.reset()
        kind = it % 4
        ob = env.obstacles[int(rng.randint(len(env.obstacles)))]
        if kind == 0:    # near an obstacle, random heading (ordered-break stress: several obstacles in view)
            ang = rng.uniform(0, 2 * np.pi); dist = ob.r + rng.uniform(0.0, 9.0)
            x, y = ob.x + dist * np.cos(ang), ob.y + dist * np.sin(ang); th = rng.uniform(0, 2 * np.pi)
        elif kind == 1:  # Q10: some beam within / just outside 1e-3 rad of +-vertical
            ang = rng.uniform(0, 2 * np.pi); dist = ob.r + rng.uniform(0.5, 8.0)
            x, y = ob.x + dist * np.cos(ang), ob.y + dist * np.sin(ang)
            b = beam_angles[int(rng.randint(len(beam_angles)))]
            target = np.pi / 2 if rng.rand() < 0.5 else 3 * np.pi / 2
            th = (target - b + rng.choice([0.0, 5e-4, -5e-4, 9.9e-4, -9.9e-4, 1.01e-3, -1.01e-3, 1e-5])) % (2 * np.pi)
        elif kind == 2:  # robot inside / abutting the obstacle (nearer-root-first quirk)
            ang = rng.uniform(0, 2 * np.pi); dist = ob.r * rng.uniform(0.0, 1.05)
            x, y = ob.x + dist * np.cos(ang), ob.y + dist * np.sin(ang); th = rng.uniform(0, 2 * np.pi)
        else:            # anywhere
            x, y = rng.uniform(0, 50), rng.uniform(0, 50); th = rng.uniform(0, 2 * np.pi)
        env.robot.x, env.robot.y, env.robot.theta, env.robot.speed = float(x), float(y), float(th), float(rng.uniform(0, 2))
        env.robot.velocity = rng.uniform(-3, 3, size=2)
        s, v = robot_state(env); c, o = tables(env, MAX_C, MAX_O)
        obs = env.get_observation()
        for k, val in zip(rec.keys(), (s, v, np.array(env.goal, dtype=np.float64), c, o, np.asarray(obs, np.float64))):
            rec[k].append(val)
    np.savez_compressed(os.path.join(HERE, "observe_vectors.npz"), **{k: np.asarray(v) for k, v in rec.items()})

    # ---- reset(): seeds 0..255, first reset, config-2 counts (4 cores, 8 obstacles, min dist 30) ---------------
    rec = {k: [] for k in ("seed", "state", "velocity", "goal", "start", "cores", "obstacles", "n_cores", "n_obs", "obs")}
    for seed in range(256):
        with quiet():
            env = ref_env.MarineNavEnv(seed=seed)
            env.num_cores, env.num_obs, env.min_start_goal_dis = 4, 8, 30.0
            obs = env.reset()
        s, v = robot_state(env); c, o = tables(env, MAX_C, MAX_O)
        for k, val in zip(rec.keys(), (seed, s, v, np.array(env.goal, dtype=np.float64), np.array(env.start, dtype=np.float64), c, o,
                                       len(env.cores), len(env.obstacles), np.asarray(obs, np.float64))):
            rec[k].append(val)
    out = {k: np.asarray(v) for k, v in rec.items()}
    # multi-reset streams with the training curriculum (train_IQN_model.py:86-90): seeds 0..7, 6 consecutive resets,
    # total_timesteps forced across the stage boundaries
    sched = dict(timesteps=[0, 1000000, 2000000], num_cores=[4, 6, 8], num_obstacles=[6, 8, 10],
                 min_start_goal_dis=[30.0, 35.0, 40.0])
    totals = [0, 500, 1000000, 1500000, 2000000, 2999999]
    srec = {k: [] for k in ("state", "goal", "cores", "obstacles", "n_cores", "n_obs", "obs")}
    for seed in range(8):
        with quiet():
            env = ref_env.MarineNavEnv(seed=seed, schedule=sched)
        for tt in totals:
            env.total_timesteps = tt
            with quiet():
                obs = env.reset()
            s, _ = robot_state(env); c, o = tables(env, MAX_C, MAX_O)
            for k, val in zip(srec.keys(), (s, np.array(env.goal, dtype=np.float64), c, o, len(env.cores), len(env.obstacles),
                                            np.asarray(obs, np.float64))):
                srec[k].append(val)
    for k, v in srec.items():
        out["stream_" + k] = np.asarray(v).reshape((8, len(totals)) + np.asarray(v).shape[1:])
    out["stream_totals"] = np.asarray(totals)
    np.savez_compressed(os.path.join(HERE, "reset_vectors.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("env", "all"):
        make_env_goldens()
    if what in ("iqn", "all"):
        from make_golden_iqn import make_iqn_goldens
        make_iqn_goldens()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
