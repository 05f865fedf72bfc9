"""MarineNavEnv: the gym-style single-environment facade over the CUDA kernels (E = 1 VecMarineNavEnv).

Keeps the reference's public surface (marinenav_env/envs/marinenav_env.py:25-627, utils/robot.py): the constructor
MarineNavEnv(seed, schedule), reset / step / seed / close / reset_with_eval_config / episode_data / save_episode /
get_state_space_dimension / get_action_space_dimension / get_velocity / get_observation, action_space /
observation_space, the mutable attributes callers poke (discount, obs_r_range, reset_start_and_goal, start, goal,
num_cores, num_obs, random_reset_state, set_boundary, schedule, total_timesteps, episode_timesteps, cores, obstacles)
and env.robot.* (dt, N, a, w, r, max_speed, k, actions, init_theta, init_speed, x, y, theta, speed, velocity,
sonar.{range, angle, num_beams, beam_angles, phi, reflections}, action_history, trajectory,
compute_action_energy_cost ...).  step() returns (np.float64[26] obs, float reward, bool done, {"state": str}) like the
reference; all arithmetic happens in mnv_step / mnv_reset / mnv_observe on the GPU (one launch + one device->host read
per call).  The random stream is the reference's: MarineNavEnv(seed=s) here generates the same maps as the reference.
"""
import copy
import json

import numpy as np
import torch

from . import _lib, env_ops
from .vec_env import VecMarineNavEnv

try:                                    # the real gym when it is installed, the stand-ins otherwise
    import gym as _gym
    if getattr(_gym, "_mnv_oracle_stub", False):
        raise ImportError
except ImportError:
    from . import compat_gym as _gym

_CAP_CORES, _CAP_OBS = 8, 32


class _TableItem:
    """A map primitive whose field edits reach the device table it was read from (env.cores[0].x = ... works in place,
    like editing the reference's Python objects)."""
    _owner = None

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        owner = object.__getattribute__(self, "_owner")
        if owner is not None and not name.startswith("_"):
            owner._push()


class Core(_TableItem):
    """marinenav_env.py:8-15"""

    def __init__(self, x, y, clockwise, Gamma):
        self.x, self.y, self.clockwise, self.Gamma = x, y, clockwise, Gamma


class Obstacle(_TableItem):
    """marinenav_env.py:17-23"""

    def __init__(self, x, y, r):
        self.x, self.y, self.r = x, y, r


class _TableList(list):
    """What env.cores / env.obstacles return: a list of the device table's entries that WRITES BACK on mutation, so the
    reference-style in-place edits (env.cores.clear(), env.obstacles.append(...), run_experiments.py:131-180) reach the
    device tables instead of being lost on a temporary."""

    def __init__(self, items, env, attr):
        super().__init__(items)
        self._env, self._attr = env, attr
        for it in self:
            object.__setattr__(it, "_owner", self)

    def _push(self):
        for it in self:
            if isinstance(it, _TableItem):
                object.__setattr__(it, "_owner", self)
        setattr(self._env, self._attr, list(self))


def _mutator(name):
    def method(self, *a, **k):
        res = getattr(list, name)(self, *a, **k)
        self._push()
        return res
    method.__name__ = name
    return method


for _m in ("append", "extend", "insert", "remove", "pop", "clear", "sort", "reverse", "__setitem__", "__delitem__", "__iadd__", "__imul__"):
    setattr(_TableList, _m, _mutator(_m))


class Sonar:
    """robot.py:3-21"""

    def __init__(self):
        self.range = 10.0
        self.angle = 2 * np.pi / 3
        self.num_beams = 11
        self.compute_phi()
        self.compute_beam_angles()
        self.reflections = []

    def compute_phi(self):
        self.phi = self.angle / (self.num_beams - 1)

    def compute_beam_angles(self):
        angle = -self.angle / 2
        self.beam_angles = [angle + i * self.phi for i in range(self.num_beams)]


class Robot:
    """robot.py:23-123: parameters live here; the dynamic state (x, y, theta, speed, velocity) is read from the device."""

    def __init__(self, env):
        self._env = env
        self.dt, self.N = 0.1, 10
        self.sonar = Sonar()
        self.length, self.width = 1.0, 0.5
        self.r = 0.8
        self.max_speed = 2.0
        self.a = np.array([-0.4, 0.0, 0.4])
        self.w = np.array([-np.pi / 6, 0.0, np.pi / 6])
        self.compute_k()
        self.compute_actions()
        self.init_theta, self.init_speed = 0.0, 0.0
        self.action_history, self.trajectory = [], []

    def compute_k(self):
        self.k = np.max(self.a) / self.max_speed

    def compute_actions(self):
        self.actions = [(acc, ang_v) for acc in self.a for ang_v in self.w]

    def compute_actions_dimension(self):
        return len(self.actions)

    def compute_dist_reward_scale(self):
        return 1 / (self.max_speed * self.N * self.dt)

    def compute_action_energy_cost(self, action):
        a, w = self.actions[action]
        return np.abs(a / np.max(self.a)) + np.abs(w / np.max(self.w))

    def _state(self):
        return self._env._vec.buf["state"][:, 0].cpu().numpy()

    x = property(lambda self: float(self._state()[0]), lambda self, v: self._env._poke_state(0, v))
    y = property(lambda self: float(self._state()[1]), lambda self, v: self._env._poke_state(1, v))
    theta = property(lambda self: float(self._state()[2]), lambda self, v: self._env._poke_state(2, v))
    speed = property(lambda self: float(self._state()[3]), lambda self, v: self._env._poke_state(3, v))

    @property
    def velocity(self):
        return self._env._vec.buf["velocity"][:, 0].cpu().numpy()

    @velocity.setter
    def velocity(self, v):
        self._env._vec.buf["velocity"][:, 0] = torch.as_tensor(np.asarray(v, np.float64), device=self._env._vec.device)

    def reset_state(self, x, y, current_velocity=np.zeros(2)):
        """robot.py:79-87"""
        self.action_history.clear(); self.trajectory.clear()
        st = torch.tensor([x, y, self.init_theta, self.init_speed], dtype=torch.float64, device=self._env._vec.device)
        self._env._vec.buf["state"][:, 0] = st
        steer = self.init_speed * np.array([np.cos(self.init_theta), np.sin(self.init_theta)])
        self.velocity = steer + np.asarray(current_velocity)

    def get_robot_transform(self):
        """robot.py:89-93"""
        x, y, th, _ = self._state()
        R_wr = np.matrix([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        return R_wr, np.matrix([[x], [y]])


class MarineNavEnv(_gym.Env):
    def __init__(self, seed: int = 0, schedule: dict = None, device="cuda:0"):
        self._vec = VecMarineNavEnv(1, seed=seed, schedule=None, device=device, num_cores=8, num_obs=5,
                                    max_cores=_CAP_CORES, max_obstacles=_CAP_OBS)
        self._beams_allocated = 11
        self.robot = Robot(self)
        self.sd = seed
        self.action_space = _gym.spaces.Discrete(self.robot.compute_actions_dimension())
        self._set_spaces()
        v = self._vec
        # marinenav_env.py:40-73
        self.width, self.height, self.r = 50, 50, 0.5
        self.v_rel_max, self.p = 1.0, 0.8
        self.v_range, self.obs_r_range, self.clear_r = [5, 10], [1, 3], 10.0
        self.reset_start_and_goal = True
        self.start = np.array([5.0, 5.0])
        self.random_reset_state = True
        self.init_speed, self.init_theta = 0.0, np.pi / 4
        self.goal = np.array([45.0, 45.0])
        self.goal_dis = 2.0
        self.timestep_penalty, self.collision_penalty, self.goal_reward = -1.0, -50.0, 100.0
        self.discount = 0.99
        self.num_cores, self.num_obs, self.min_start_goal_dis = 8, 5, 25.0
        self.schedule = schedule
        self.episode_timesteps = 0
        self.total_timesteps = 0
        self.set_boundary = False
        self.verbose_schedule = True           # the reference prints the schedule on every reset (marinenav_env.py:100-104)
        assert v.num_envs == 1

    # ---- plumbing ------------------------------------------------------------------------------------------------------
    def _set_spaces(self):
        obs_len = 2 + 2 + 2 * self.robot.sonar.num_beams
        self.observation_space = _gym.spaces.Box(low=-np.inf * np.ones(obs_len), high=np.inf * np.ones(obs_len), dtype=np.float32)

    def _poke_state(self, i, value):
        self._vec.buf["state"][i, 0] = float(value)

    def _sync(self):
        """Push the facade's (mutable) attributes into the vector env before a launch."""
        v, rb = self._vec, self.robot
        nb = rb.sonar.num_beams
        if nb != v.num_beams:                                # sonar.num_beams was changed by the caller: re-allocate rows
            v.num_beams = nb
            v.buf["obs"] = torch.zeros(1, 4 + 2 * nb, dtype=torch.float32, device=v.device)
            v.buf["next_obs"] = torch.zeros_like(v.buf["obs"])
            v._pinned = None
            self._set_spaces()
        for name in ("width", "height", "r", "v_rel_max", "p", "v_range", "obs_r_range", "clear_r", "reset_start_and_goal",
                     "start", "goal", "random_reset_state", "init_speed", "init_theta", "goal_dis", "timestep_penalty",
                     "collision_penalty", "goal_reward", "discount", "num_cores", "num_obs", "min_start_goal_dis", "set_boundary"):
            setattr(v, name, getattr(self, name))
        v.dt, v.N, v.robot_r, v.max_speed, v.a, v.w = rb.dt, rb.N, rb.r, rb.max_speed, np.asarray(rb.a), np.asarray(rb.w)
        v.sonar_range, v.sonar_angle = rb.sonar.range, rb.sonar.angle
        goal = (float(self.goal[0]), float(self.goal[1]))      # callers assign env.goal directly (run_experiments.py:134,196)
        if getattr(self, "_goal_pushed", None) != goal:
            v.buf["goal"][:, 0] = torch.tensor(goal, dtype=torch.float64, device=v.device)
            v.tables_written()
            self._goal_pushed = goal

    def _obs_out(self, obs_t):
        return obs_t[0].double().cpu().numpy()

    # ---- gym surface -----------------------------------------------------------------------------------------------------
    def seed(self, seed):
        """marinenav_env.py:75-78"""
        self.sd = seed
        self._vec.seed(seed)
        return [seed]

    def get_state_space_dimension(self):
        return 2 + 2 + 2 * self.robot.sonar.num_beams

    def get_action_space_dimension(self):
        return self.robot.compute_actions_dimension()

    def reset(self):
        """marinenav_env.py:86-186"""
        if self.schedule is not None:
            steps = self.schedule["timesteps"]
            diffs = np.array(steps) - self.total_timesteps
            idx = len(diffs[diffs <= 0]) - 1
            self.num_cores = self.schedule["num_cores"][idx]
            self.num_obs = self.schedule["num_obstacles"][idx]
            self.min_start_goal_dis = self.schedule["min_start_goal_dis"][idx]
            if self.verbose_schedule:
                print("======== training schedule ========")
                print("num of cores: ", self.num_cores)
                print("num of obstacles: ", self.num_obs)
                print("min start goal dis: ", self.min_start_goal_dis)
                print("======== training schedule ========\n")
        self.episode_timesteps = 0
        self._sync()
        obs = self._vec.reset()
        b = self._vec.buf
        sp = b["start_pose"][:, 0].cpu().numpy(); goal = b["goal"][:, 0].cpu().numpy()
        self.start, self.goal = np.array(sp[:2]), np.array(goal)
        self._goal_pushed = (float(goal[0]), float(goal[1]))
        self.robot.init_theta, self.robot.init_speed = float(sp[2]), float(sp[3])
        self.robot.action_history.clear(); self.robot.trajectory.clear()
        return self._obs_out(obs)

    def step(self, action):
        """marinenav_env.py:199-262"""
        a = int(action)
        self.robot.actions[a]                                   # IndexError for an out-of-range action, like robot.py:110
        self.robot.action_history.append(action)
        self._sync()
        v = self._vec
        b = v.buf
        pin = self._step_pins()
        pin["in"][0], pin["in"][1] = a, int(self.episode_timesteps)
        with torch.cuda.device(v.device):
            self._dev_in.copy_(pin["in"], non_blocking=True)    # action + episode_timesteps in one 8-byte copy
            b["action"].copy_(self._dev_in[0:1]); b["episode_step"].copy_(self._dev_in[1:2])
            env_ops.step(b, v.params(), action=b["action"], obs=b["next_obs"], trajectory=self._traj)
            b["obs"].copy_(b["next_obs"])
            v.total_timesteps += 1
            pin["obs"].copy_(b["next_obs"], non_blocking=True); pin["reward"].copy_(b["reward"], non_blocking=True)
            pin["done"].copy_(b["done"], non_blocking=True); pin["info"].copy_(b["info"], non_blocking=True)
            pin["traj"].copy_(self._traj, non_blocking=True)
            torch.cuda.current_stream().synchronize()           # ONE host sync per step
        obs = pin["obs"][0].double().numpy().copy()
        reward = float(pin["reward"][0])
        done = bool(pin["done"][0])
        info = {"state": _lib.INFO_STRINGS[int(pin["info"][0])]}
        self.robot.trajectory.extend(pin["traj"][:, :, 0].tolist())
        self.episode_timesteps += 1
        self.total_timesteps += 1
        return obs, reward, done, info

    def _step_pins(self):
        """Pinned staging buffers of the single-env step (re-made when N or the number of beams changes)."""
        v = self._vec
        key = (int(self.robot.N), v.obs_dim)
        if getattr(self, "_pins_key", None) != key:
            self._pins_key = key
            self._pins = dict(obs=torch.zeros(1, v.obs_dim, dtype=torch.float32).pin_memory(),
                              reward=torch.zeros(1, dtype=torch.float32).pin_memory(), done=torch.zeros(1, dtype=torch.uint8).pin_memory(),
                              info=torch.zeros(1, dtype=torch.uint8).pin_memory(),
                              traj=torch.zeros(key[0], 2, 1, dtype=torch.float64).pin_memory(),
                              **{"in": torch.zeros(2, dtype=torch.int32).pin_memory()})
            self._traj = torch.zeros(key[0], 2, 1, dtype=torch.float64, device=v.device)
            self._dev_in = torch.zeros(2, dtype=torch.int32, device=v.device)
        return self._pins

    def close(self):
        self._vec.close()

    # ---- map / state access ------------------------------------------------------------------------------------------------
    @property
    def cores(self):
        b = self._vec.buf
        n, mc = int(b["n_placed"][0, 0]), self._vec.max_cores
        t = b["cores"][:, 0].cpu().numpy()
        return _TableList([Core(float(t[k]), float(t[mc + k]), int(t[2 * mc + k] > 0), float(abs(t[2 * mc + k]))) for k in range(n)],
                          self, "cores")

    @cores.setter
    def cores(self, cores):
        b, mc = self._vec.buf, self._vec.max_cores
        t = np.zeros(3 * mc)
        for k, c in enumerate(cores):
            t[k], t[mc + k], t[2 * mc + k] = c.x, c.y, (c.Gamma if c.clockwise else -c.Gamma)
        b["cores"][:, 0] = torch.from_numpy(t).to(self._vec.device)
        b["n_placed"][0, 0] = len(cores)
        self._vec.tables_written()

    @property
    def obstacles(self):
        b = self._vec.buf
        n, mo = int(b["n_placed"][1, 0]), self._vec.max_obstacles
        t = b["obstacles"][:, 0].cpu().numpy()
        return _TableList([Obstacle(float(t[k]), float(t[mo + k]), float(t[2 * mo + k])) for k in range(n)], self, "obstacles")

    @obstacles.setter
    def obstacles(self, obstacles):
        b, mo = self._vec.buf, self._vec.max_obstacles
        t = np.zeros(3 * mo)
        for k, o in enumerate(obstacles):
            t[k], t[mo + k], t[2 * mo + k] = o.x, o.y, o.r
        b["obstacles"][:, 0] = torch.from_numpy(t).to(self._vec.device)
        b["n_placed"][1, 0] = len(obstacles)
        self._vec.tables_written()

    # The reference keeps scipy KDTrees of the centres (marinenav_env.py:155,181) and callers re-assign them after editing
    # the lists (run_experiments.py:160,178).  The kernels do not need them: reading builds one on demand, writing is accepted
    # and ignored (the tables written by the cores / obstacles setters are the source of truth).
    @property
    def core_centers(self):
        import scipy.spatial
        cs = self.cores
        return scipy.spatial.KDTree(np.array([[c.x, c.y] for c in cs])) if cs else None

    @core_centers.setter
    def core_centers(self, tree):
        pass

    @property
    def obs_centers(self):
        import scipy.spatial
        os_ = self.obstacles
        return scipy.spatial.KDTree(np.array([[o.x, o.y] for o in os_])) if os_ else None

    @obs_centers.setter
    def obs_centers(self, tree):
        pass

    def get_velocity(self, x: float, y: float):
        """marinenav_env.py:422-455: current at (x, y) -- evaluated by the device kernel on a scratch copy of the map."""
        v = self._vec
        self._sync()
        tmp = {k: (t.clone() if k in ("state", "velocity", "obs") else t) for k, t in v.buf.items()}
        tmp["state"][:, 0] = torch.tensor([x, y, 0.0, 0.0], dtype=torch.float64, device=v.device)
        with torch.cuda.device(v.device):
            env_ops.observe(tmp, v.params(), velocity_from_state=True)
        return tmp["velocity"][:, 0].cpu().numpy()

    def get_observation(self, for_visualize=False):
        """marinenav_env.py:273-326"""
        v = self._vec
        self._sync()
        with torch.cuda.device(v.device):
            env_ops.observe(v.buf, v.params(), velocity_from_state=False)
        obs = self._obs_out(v.buf["obs"])
        if not for_visualize:
            return obs
        nb, rng = self.robot.sonar.num_beams, self.robot.sonar.range
        pts = obs[4:].reshape(nb, 2)
        hit = ~((pts[:, 0] == 0.0) & (pts[:, 1] == 0.0))
        ba = np.asarray(self.robot.sonar.beam_angles)
        far = 2.0 * rng * np.stack([np.cos(ba), np.sin(ba)], axis=1)
        out = np.vstack([np.where(hit[:, None], pts, far).T, hit.astype(np.float64)[None, :]])
        return obs[:2], np.matrix(out), obs[2:4]

    def out_of_boundary(self):
        x, y = self.robot.x, self.robot.y
        return (x < 0.0 or x > self.width) or (y < 0.0 or y > self.height)

    def dist_to_goal(self):
        return float(np.linalg.norm(self.goal - np.array([self.robot.x, self.robot.y])))

    # ---- evaluation configs ------------------------------------------------------------------------------------------------
    def reset_with_eval_config(self, eval_config):
        """marinenav_env.py:467-555 (does not touch the random stream)."""
        self.episode_timesteps = 0
        e, r = eval_config["env"], eval_config["robot"]
        self.sd = e["seed"]
        for k in ("width", "height", "r", "v_rel_max", "p", "clear_r", "goal_dis", "timestep_penalty", "collision_penalty",
                  "goal_reward", "discount"):
            setattr(self, k, e[k])
        self.v_range, self.obs_r_range = copy.deepcopy(e["v_range"]), copy.deepcopy(e["obs_r_range"])
        self.start, self.goal = np.array(e["start"]), np.array(e["goal"])
        rb = self.robot
        rb.dt, rb.N, rb.length, rb.width, rb.r, rb.max_speed = r["dt"], r["N"], r["length"], r["width"], r["r"], r["max_speed"]
        rb.a, rb.w = np.array(r["a"]), np.array(r["w"])
        rb.compute_k(); rb.compute_actions()
        rb.init_theta, rb.init_speed = r["init_theta"], r["init_speed"]
        rb.sonar.range, rb.sonar.angle, rb.sonar.num_beams = r["sonar"]["range"], r["sonar"]["angle"], r["sonar"]["num_beams"]
        rb.sonar.compute_phi(); rb.sonar.compute_beam_angles()
        self.action_space = _gym.spaces.Discrete(rb.compute_actions_dimension())
        self._set_spaces()
        self._sync()
        self._vec.load_eval_configs([eval_config])
        self._goal_pushed = (float(self.goal[0]), float(self.goal[1]))
        rb.action_history.clear(); rb.trajectory.clear()
        return self._obs_out(self._vec.observe_all())

    def episode_data(self):
        """marinenav_env.py:557-622"""
        self._sync()
        ep = self._vec.episode_data(0, self.robot.action_history, self.robot.trajectory)
        ep["env"]["seed"] = self.sd
        ep["env"]["start"], ep["env"]["goal"] = list(self.start), list(self.goal)
        ep["robot"]["length"], ep["robot"]["width"] = self.robot.length, self.robot.width
        ep["robot"]["init_theta"], ep["robot"]["init_speed"] = self.robot.init_theta, self.robot.init_speed
        ep["robot"]["action_history"] = copy.deepcopy(self.robot.action_history)
        return ep

    def save_episode(self, filename):
        with open(filename, "w") as f:
            json.dump(self.episode_data(), f, default=lambda o: o.item() if hasattr(o, "item") else o)
