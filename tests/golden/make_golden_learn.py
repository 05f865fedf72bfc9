"""Golden trace of the UNMODIFIED reference's training loop, IQNAgent.learn (thirdparty/IQN/agent.py:94-173).

    python tests/golden/make_golden_learn.py      ->  tests/golden/learn_trace.npz

A seeded short run on the reference (MarineNavEnv + IQNAgent on the CPU, imported through oracle/ref_import.py):
env seed 13 (default map rules; two episodes end inside the run), agent seed 5, learning_starts 40, batch 32, UPDATE_EVERY 4, target interval 16,
total_timesteps 240 (epsilon 1.0 -> 0.05 over the first 24 steps), one evaluation map (env_0 of eval_config.json) evaluated
greedy + adaptive at learning step 0.  Recorded: the action / reward / done of every training step, the replay picks
(indices into the buffer, in the order random.sample returned them) and the loss of every train() call, and the two
evaluation episodes.  All random choices come from the streams the reference uses (python `random` seeded by the replay
buffer, torch's CPU generator seeded by the network constructors, numpy RandomState of the env), so an implementation that
mirrors the loop reproduces the trace: actions / picks / flags exactly, rewards and losses to float tolerance.
"""
import contextlib
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402

CFG = dict(env_seed=13, agent_seed=5, learning_starts=40, batch_size=32, update_every=4, target_update_interval=16,
           total_timesteps=240, eval_freq=10 ** 9)


def record_learn(make_env, make_agent, eval_cfg, quiet=True):
    """Runs learn() with the hooks that capture the trace; shared by the generator (reference classes) and the GPU test
    (drop-in classes), so both sides are observed through the same instrumentation."""
    import tempfile
    train_env, eval_env = make_env(CFG["env_seed"]), make_env(348)
    agent = make_agent()
    trace = dict(actions=[], rewards=[], dones=[], losses=[], picks=[])
    real_step, real_train, real_sample = train_env.step, agent.train, random.sample

    def step(a):
        o, r, d, i = real_step(a)
        trace["actions"].append(int(a)); trace["rewards"].append(float(r)); trace["dones"].append(bool(d))
        return o, r, d, i

    def train(exp, *a, **k):
        loss = real_train(exp, *a, **k)
        trace["losses"].append(float(loss))
        return loss

    def sample(population, k, **kw):
        if isinstance(population, range):                       # the drop-in buffer samples indices directly
            picks = real_sample(population, k, **kw)
            trace["picks"].append([int(p) for p in picks])
            return picks
        res = real_sample(population, k, **kw)                  # the reference samples the deque's elements
        where = {id(e): n for n, e in enumerate(population)}
        trace["picks"].append([where[id(e)] for e in res])
        return res

    train_env.step, agent.train, random.sample = step, train, sample
    try:
        with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
            agent.learn(total_timesteps=CFG["total_timesteps"], train_env=train_env, eval_env=eval_env,
                        eval_config={"env_0": eval_cfg}, eval_freq=CFG["eval_freq"], eval_log_path=tmp)
    finally:
        random.sample = real_sample
    out = {k: np.asarray(v) for k, v in trace.items()}
    for pol in ("greedy", "adaptive"):
        out[f"eval_{pol}_actions"] = np.asarray(agent.eval_actions[pol][0][0], np.int64)
        out[f"eval_{pol}_reward"] = np.float64(agent.eval_rewards[pol][0][0])
        out[f"eval_{pol}_success"] = np.bool_(agent.eval_successes[pol][0][0])
        out[f"eval_{pol}_timestep"] = np.int64(agent.eval_timesteps[pol][0])
    return out


def main():
    import torch
    torch.set_num_threads(1)
    ref_env, ref_agent, _ = ref_import.load_reference()
    with open(os.path.join(HERE, "eval_config.json")) as f:
        eval_cfg = json.load(f)["env_0"]
    out = record_learn(lambda s: ref_env.MarineNavEnv(seed=s),
                       lambda: ref_agent.IQNAgent(26, 9, seed=CFG["agent_seed"], BATCH_SIZE=CFG["batch_size"],
                                                  UPDATE_EVERY=CFG["update_every"], learning_starts=CFG["learning_starts"],
                                                  target_update_interval=CFG["target_update_interval"]), eval_cfg)
    out["config_json"] = np.asarray(json.dumps(CFG))
    np.savez_compressed(os.path.join(HERE, "learn_trace.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})
    print("losses", out["losses"][:5], "n_train", len(out["losses"]), "dones", int(out["dones"].sum()),
          "eval", out["eval_greedy_actions"].shape, out["eval_adaptive_actions"].shape, out["eval_greedy_reward"])


if __name__ == "__main__":
    main()
