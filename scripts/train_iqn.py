#!/usr/bin/env python
"""Training driver with the reference's config format (config/config_IQN.json: agent, seed, total_timesteps, eval_freq,
save_dir; list-valued fields expand to a cartesian product of trials, train_IQN_model.py:52-65) on the B200 path.

    python scripts/train_iqn.py -C config.json [-D cuda:0] [--num-envs 65536] [--batch-size 1024]

--num-envs 0 (default) runs the reference's single-env loop (IQNAgent.learn on the gym-style MarineNavEnv facade, what
train_IQN_model.py does); --num-envs N > 0 runs the vectorised trainer (IQNAgent.learn_vec on a VecMarineNavEnv of N
environments, evaluation of the 30 fixed maps as one batch).  Under torchrun every rank trains on its own env shard
and the gradient is all-reduced once per update.
"""
import argparse
import itertools
import json
import os
import sys
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

TRAINING_SCHEDULE = dict(timesteps=[0, 1000000, 2000000], num_cores=[4, 6, 8], num_obstacles=[6, 8, 10],
                         min_start_goal_dis=[30.0, 35.0, 40.0])          # train_IQN_model.py:86-90


def trial_params(params):
    """Cartesian product over list-valued fields (train_IQN_model.py:52-65)."""
    if isinstance(params, (str, int, float)):
        return [params]
    if isinstance(params, list):
        return params
    if isinstance(params, dict):
        keys, vals = zip(*params.items())
        return [dict(zip(keys, mix)) for mix in itertools.product(*[trial_params(v) for v in vals])]
    raise TypeError("Parameter type is incorrect.")


def create_eval_configs(eval_env):
    """30 evaluation maps of increasing difficulty from the eval env's stream (train_IQN_model.py:123-148)."""
    eval_config, count = {}, 0
    eval_env.obs_r_range = [1, 3]
    eval_env.reset_start_and_goal = False
    eval_env.start, eval_env.goal = np.array([5.0, 5.0]), np.array([45.0, 45.0])
    for num_c, num_o in ((4, 6), (6, 8), (8, 10)):
        for _ in range(10):
            eval_env.num_cores, eval_env.num_obs = num_c, num_o
            eval_env.reset()
            eval_config[f"env_{count}"] = eval_env.episode_data()
            count += 1
    return eval_config


def run_trial(device, params, num_envs, batch_size, updates_per_step=1, graph=False):
    import marinenav_env  # noqa: F401  (registers 'marinenav_env-v0')
    from distributional_rl_navigation_b200 import distributed as mdist
    from distributional_rl_navigation_b200 import marinenav_env as impl
    from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv
    from thirdparty import IQNAgent

    rank, world = mdist.init_from_env()
    exp_dir = os.path.join(params["save_dir"], "training_" + params["training_time"], "seed_" + str(params["seed"]))
    if rank == 0:
        os.makedirs(exp_dir, exist_ok=True)
        with open(os.path.join(exp_dir, "trial_config.json"), "w+") as f:
            json.dump(params, f)
        with open(os.path.join(exp_dir, "training_schedule.json"), "w+") as f:
            json.dump(TRAINING_SCHEDULE, f)
    eval_env = impl._gym.make('marinenav_env:marinenav_env-v0', seed=348)
    eval_env.verbose_schedule = False
    eval_config = create_eval_configs(eval_env)
    if rank == 0:
        with open(os.path.join(exp_dir, "eval_config.json"), "w+") as f:
            json.dump(eval_config, f)
    log_dir = exp_dir if rank == 0 else None
    if num_envs <= 0:
        train_env = impl._gym.make('marinenav_env:marinenav_env-v0', seed=params["seed"], schedule=TRAINING_SCHEDULE)
        train_env.verbose_schedule = False
        model = IQNAgent(train_env.get_state_space_dimension(), train_env.get_action_space_dimension(), device=device,
                         seed=params["seed"] + 100)
        model.learn(total_timesteps=params["total_timesteps"], train_env=train_env, eval_env=eval_env, eval_config=eval_config,
                    eval_freq=params["eval_freq"], eval_log_path=log_dir, verbose=params.get("verbose", False))
    else:
        lo, hi = mdist.shard_range(num_envs * world, rank, world)
        train_env = VecMarineNavEnv(hi - lo, seed=params["seed"] + lo, schedule=TRAINING_SCHEDULE, device=device)
        model = IQNAgent(train_env.get_state_space_dimension(), train_env.get_action_space_dimension(), device=device,
                         seed=params["seed"] + 100, BATCH_SIZE=batch_size)
        model.learn_vec(total_timesteps=params["total_timesteps"], train_env=train_env, eval_config=eval_config,
                        eval_freq=params["eval_freq"], eval_log_path=log_dir, batch_size=batch_size,
                        updates_per_step=updates_per_step, sample_without_replacement=True, graph=graph)
    train_env.close(); eval_env.close()
    return exp_dir, model


def main():
    ap = argparse.ArgumentParser(description="Train IQN model")
    ap.add_argument("-C", "--config-file", dest="config_file", type=open, required=True, help="training config json file")
    ap.add_argument("-D", "--device", dest="device", type=str, default="cuda:0", help="device to run all trials")
    ap.add_argument("--num-envs", type=int, default=0, help="environments per GPU of the vectorised trainer (0: single-env loop)")
    ap.add_argument("--batch-size", type=int, default=1024)
    ap.add_argument("--updates-per-step", default="1", help="IQN updates per vector step, or 'reference' = the reference's replay "
                    "ratio (one update of 32 per 4 transitions, agent.py:127-136) in vector form")
    ap.add_argument("--graph", action="store_true", help="vectorised trainer: replay the whole vector step as ONE CUDA graph "
                    "(learn_vec(graph=True): pipelined learner, device control block)")
    args = ap.parse_args()
    ups = args.updates_per_step if args.updates_per_step == "reference" else int(args.updates_per_step)
    params = json.load(args.config_file)
    trials = trial_params(params)
    stamp = datetime.now().strftime("%Y-%m-%d-%H-%M-%S")
    if "LOCAL_RANK" in os.environ:
        args.device = "cuda:%d" % int(os.environ["LOCAL_RANK"])
    for p in trials:
        p["training_time"] = stamp
        exp_dir, _ = run_trial(args.device, p, args.num_envs, args.batch_size, ups, graph=args.graph)
        print("trial done:", exp_dir)


if __name__ == "__main__":
    main()
