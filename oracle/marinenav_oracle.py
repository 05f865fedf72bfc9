"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libmarinenav_oracle.so (the CPU parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package never does (tests/test_abi_cpu.py::test_product_never_imports_oracle enforces it).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmarinenav_oracle.so")

MAX_CORES, MAX_OBS, MAX_BEAMS, MAX_SCHED = 32, 64, 128, 8
INFO_STRINGS = ["normal", "too long episode", "collision", "reach goal", "out of boundary"]


class Core(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("clockwise", C.c_int), ("Gamma", C.c_double)]


class Obstacle(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("r", C.c_double)]


class MT19937(C.Structure):
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int)]


class Env(C.Structure):
    _fields_ = [
        ("width", C.c_double), ("height", C.c_double), ("r", C.c_double), ("v_rel_max", C.c_double), ("p", C.c_double),
        ("v_range", C.c_double * 2), ("obs_r_range", C.c_double * 2),
        ("clear_r", C.c_double),
        ("reset_start_and_goal", C.c_int),
        ("start", C.c_double * 2), ("goal", C.c_double * 2),
        ("random_reset_state", C.c_int),
        ("init_speed", C.c_double), ("init_theta", C.c_double),
        ("goal_dis", C.c_double), ("timestep_penalty", C.c_double), ("collision_penalty", C.c_double),
        ("goal_reward", C.c_double), ("discount", C.c_double),
        ("num_cores", C.c_int), ("num_obs", C.c_int),
        ("min_start_goal_dis", C.c_double),
        ("set_boundary", C.c_int),
        ("n_sched", C.c_int),
        ("sched_timesteps", C.c_int64 * MAX_SCHED),
        ("sched_num_cores", C.c_int * MAX_SCHED), ("sched_num_obs", C.c_int * MAX_SCHED),
        ("sched_min_start_goal_dis", C.c_double * MAX_SCHED),
        ("dt", C.c_double), ("N", C.c_int),
        ("robot_r", C.c_double), ("max_speed", C.c_double), ("a", C.c_double * 3), ("w", C.c_double * 3), ("k", C.c_double),
        ("sonar_range", C.c_double), ("sonar_angle", C.c_double), ("num_beams", C.c_int),
        ("beam_angles", C.c_double * MAX_BEAMS),
        ("x", C.c_double), ("y", C.c_double), ("theta", C.c_double), ("speed", C.c_double),
        ("vx", C.c_double), ("vy", C.c_double),
        ("robot_init_theta", C.c_double), ("robot_init_speed", C.c_double),
        ("episode_timesteps", C.c_int),
        ("total_timesteps", C.c_int64),
        ("n_cores_placed", C.c_int), ("n_obs_placed", C.c_int),
        ("cores", Core * MAX_CORES),
        ("obstacles", Obstacle * MAX_OBS),
        ("refl_x", C.c_double * MAX_BEAMS), ("refl_y", C.c_double * MAX_BEAMS), ("refl_hit", C.c_int * MAX_BEAMS),
        ("rd", MT19937),
    ]


class Params(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("n_substeps", C.c_int),
        ("accel", C.c_double * 3), ("yaw_rate", C.c_double * 3), ("k_drag", C.c_double), ("max_speed", C.c_double),
        ("robot_r", C.c_double), ("core_r", C.c_double), ("goal_dis", C.c_double),
        ("timestep_penalty", C.c_double), ("collision_penalty", C.c_double), ("goal_reward", C.c_double),
        ("sonar_range", C.c_double), ("sonar_angle", C.c_double), ("n_beams", C.c_int),
        ("max_episode_steps", C.c_int), ("set_boundary", C.c_int), ("width", C.c_double), ("height", C.c_double),
    ]


def build(force=False):
    """Compile the oracle (gcc). Building the checker is not using it."""
    src = os.path.join(_HERE, "marinenav_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "marinenav_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        pd = C.POINTER(C.c_double)
        L.orc_sizeof_env.restype = C.c_ulong
        assert L.orc_sizeof_env() == C.sizeof(Env), (L.orc_sizeof_env(), C.sizeof(Env))
        L.orc_mt_random_sample.restype = C.c_double
        L.orc_step.restype = C.c_int
        L.orc_step.argtypes = [C.POINTER(Env), C.c_int, pd, pd, C.POINTER(C.c_int)]
        L.orc_reset.argtypes = [C.POINTER(Env), pd]
        L.orc_restart_episode.argtypes = [C.POINTER(Env), pd]
        L.orc_get_observation.argtypes = [C.POINTER(Env), pd]
        L.orc_get_velocity.argtypes = [C.POINTER(Env), C.c_double, C.c_double, pd]
        L.orc_env_init.argtypes = [C.POINTER(Env), C.c_uint32]
        L.orc_set_num_beams.argtypes = [C.POINTER(Env), C.c_int]
        _lib = L
    return _lib


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def default_params(n_beams=11):
    p = Params()
    lib().orc_default_params(C.byref(p))
    p.n_beams = n_beams
    return p


class OracleEnv:
    """Stateful single environment mirroring MarineNavEnv(seed, schedule) of the reference."""

    def __init__(self, seed=0, schedule=None):
        self.L = lib()
        self.e = Env()
        self.L.orc_env_init(C.byref(self.e), C.c_uint32(seed))
        if schedule is not None:
            n = len(schedule["timesteps"])
            self.e.n_sched = n
            for i in range(n):
                self.e.sched_timesteps[i] = int(schedule["timesteps"][i])
                self.e.sched_num_cores[i] = int(schedule["num_cores"][i])
                self.e.sched_num_obs[i] = int(schedule["num_obstacles"][i])
                self.e.sched_min_start_goal_dis[i] = float(schedule["min_start_goal_dis"][i])

    @property
    def obs_dim(self):
        return 4 + 2 * self.e.num_beams

    def set_num_beams(self, n):
        self.L.orc_set_num_beams(C.byref(self.e), n)

    def reset(self):
        obs = np.zeros(self.obs_dim)
        self.L.orc_reset(C.byref(self.e), _p(obs))
        return obs

    def step(self, action):
        obs = np.zeros(self.obs_dim)
        r = C.c_double()
        info = C.c_int()
        done = self.L.orc_step(C.byref(self.e), int(action), _p(obs), C.byref(r), C.byref(info))
        return obs, r.value, bool(done), {"state": INFO_STRINGS[info.value]}

    def get_observation(self):
        obs = np.zeros(self.obs_dim)
        self.L.orc_get_observation(C.byref(self.e), _p(obs))
        return obs

    def get_velocity(self, x, y):
        out = np.zeros(2)
        self.L.orc_get_velocity(C.byref(self.e), float(x), float(y), _p(out))
        return out

    def reset_with_eval_config(self, cfg):
        """marinenav_env.py:467-555"""
        e, env, rob = self.e, cfg["env"], cfg["robot"]
        e.episode_timesteps = 0
        e.width, e.height, e.r = env["width"], env["height"], env["r"]
        e.v_rel_max, e.p = env["v_rel_max"], env["p"]
        e.v_range[0], e.v_range[1] = env["v_range"]
        e.obs_r_range[0], e.obs_r_range[1] = env["obs_r_range"]
        e.clear_r = env["clear_r"]
        e.start[0], e.start[1] = env["start"]
        e.goal[0], e.goal[1] = env["goal"]
        e.goal_dis = env["goal_dis"]
        e.timestep_penalty, e.collision_penalty = env["timestep_penalty"], env["collision_penalty"]
        e.goal_reward, e.discount = env["goal_reward"], env["discount"]
        e.n_cores_placed = len(env["cores"]["positions"])
        for i, (pos, cw, G) in enumerate(zip(env["cores"]["positions"], env["cores"]["clockwise"], env["cores"]["Gamma"])):
            e.cores[i].x, e.cores[i].y, e.cores[i].clockwise, e.cores[i].Gamma = pos[0], pos[1], int(bool(cw)), G
        e.n_obs_placed = len(env["obstacles"]["positions"])
        for i, (pos, r) in enumerate(zip(env["obstacles"]["positions"], env["obstacles"]["r"])):
            e.obstacles[i].x, e.obstacles[i].y, e.obstacles[i].r = pos[0], pos[1], r
        e.dt, e.N, e.robot_r, e.max_speed = rob["dt"], rob["N"], rob["r"], rob["max_speed"]
        for i in range(3):
            e.a[i], e.w[i] = rob["a"][i], rob["w"][i]
        e.k = max(rob["a"]) / rob["max_speed"]
        e.robot_init_theta, e.robot_init_speed = rob["init_theta"], rob["init_speed"]
        e.sonar_range, e.sonar_angle = rob["sonar"]["range"], rob["sonar"]["angle"]
        self.set_num_beams(rob["sonar"]["num_beams"])
        obs = np.zeros(self.obs_dim)
        self.L.orc_restart_episode(C.byref(e), _p(obs))
        return obs


# ---- batch forms over the SoA buffers shared with the CUDA C-ABI ----------------------------------------------

def step_batch(state, velocity, goal, cores, obstacles, action, episode_step, params, n_threads=1):
    """In-place on state/velocity/episode_step; returns obs f64[E,D], reward f64[E], done u8[E], info u8[E]."""
    E = state.shape[1]
    max_c, max_o = cores.shape[0] // 3, obstacles.shape[0] // 3
    D = 4 + 2 * params.n_beams
    obs = np.zeros((E, D)); reward = np.zeros(E)
    done = np.zeros(E, np.uint8); info = np.zeros(E, np.uint8)
    for a in (state, velocity, goal, cores, obstacles):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    assert action.dtype == np.int32 and episode_step.dtype == np.int32
    lib().orc_step_batch(_p(state), _p(velocity), _p(goal), _p(cores), _p(obstacles),
                         _p(action, C.c_int32), _p(episode_step, C.c_int32),
                         _p(obs), _p(reward), _p(done, C.c_uint8), _p(info, C.c_uint8),
                         C.c_int64(E), max_c, max_o, C.byref(params), int(n_threads))
    return obs, reward, done, info


def observe_batch(state, velocity, goal, cores, obstacles, params, velocity_from_state=False, n_threads=1):
    E = state.shape[1]
    max_c, max_o = cores.shape[0] // 3, obstacles.shape[0] // 3
    obs = np.zeros((E, 4 + 2 * params.n_beams))
    lib().orc_observe_batch(_p(state), _p(velocity), _p(goal), _p(cores), _p(obstacles), _p(obs),
                            C.c_int64(E), max_c, max_o, C.byref(params), int(bool(velocity_from_state)), int(n_threads))
    return obs


def reset_batch(seeds, num_cores, num_obs, min_start_goal_dis, max_c, max_o, params, n_threads=1):
    """Env i == MarineNavEnv(seed=seeds[i]) with the given counts, first reset(). Returns a dict of SoA arrays."""
    seeds = np.ascontiguousarray(seeds, np.uint32)
    E = seeds.shape[0]
    out = dict(state=np.zeros((4, E)), velocity=np.zeros((2, E)), goal=np.zeros((2, E)),
               cores=np.zeros((3 * max_c, E)), obstacles=np.zeros((3 * max_o, E)), start_pose=np.zeros((4, E)),
               n_cores=np.zeros(E, np.uint8), n_obs=np.zeros(E, np.uint8), obs=np.zeros((E, 4 + 2 * params.n_beams)))
    assert num_cores <= max_c and num_obs <= max_o
    lib().orc_reset_batch(_p(seeds, C.c_uint32), int(num_cores), int(num_obs), C.c_double(min_start_goal_dis),
                          _p(out["state"]), _p(out["velocity"]), _p(out["goal"]), _p(out["cores"]), _p(out["obstacles"]),
                          _p(out["start_pose"]), _p(out["n_cores"], C.c_uint8), _p(out["n_obs"], C.c_uint8),
                          _p(out["obs"]), C.c_int64(E), max_c, max_o, C.byref(params), int(n_threads))
    return out


def tables_from_eval_config(cfgs, max_c, max_o):
    """SoA map tables for a list of eval-config dicts (schema of pretrained_models/IQN/seed_3/eval_config.json)."""
    E = len(cfgs)
    state = np.zeros((4, E)); goal = np.zeros((2, E))
    cores = np.zeros((3 * max_c, E)); obst = np.zeros((3 * max_o, E))
    for i, cfg in enumerate(cfgs):
        env, rob = cfg["env"], cfg["robot"]
        state[:, i] = [env["start"][0], env["start"][1], rob["init_theta"], rob["init_speed"]]
        goal[:, i] = env["goal"]
        for c, (pos, cw, G) in enumerate(zip(env["cores"]["positions"], env["cores"]["clockwise"], env["cores"]["Gamma"])):
            cores[c, i], cores[max_c + c, i], cores[2 * max_c + c, i] = pos[0], pos[1], (G if cw else -G)
        for o, (pos, r) in enumerate(zip(env["obstacles"]["positions"], env["obstacles"]["r"])):
            obst[o, i], obst[max_o + o, i], obst[2 * max_o + o, i] = pos[0], pos[1], r
    return state, goal, cores, obst
