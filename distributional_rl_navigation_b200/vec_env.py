"""VecMarineNavEnv: E independent marinenav environments resident on one B200, stepped by one fused kernel.

The attribute names and defaults mirror the reference's MarineNavEnv / Robot / Sonar (marinenav_env.py:27-73,
robot.py:5-49) so that code written against the reference env finds the same knobs; the per-environment state lives in
device tensors laid out as include/marinenav_b200.h describes.  Environment e uses its own numpy-compatible MT19937
stream seeded with ``seed + e`` -> its maps are the ones ``MarineNavEnv(seed=seed + e)`` of the reference generates.
"""
import math
import os
import time

import numpy as np
import torch

from . import _lib, env_ops

INFO_STRINGS = _lib.INFO_STRINGS


class VecMarineNavEnv:
    def __init__(self, num_envs, seed=0, schedule=None, device="cuda:0", num_cores=8, num_obs=5, min_start_goal_dis=25.0,
                 num_beams=11, max_cores=None, max_obstacles=None, pdl_prefetch=True, host_transport="auto"):
        if not torch.cuda.is_available():
            raise _lib.MarinenavError("VecMarineNavEnv needs a CUDA device (there is no CPU fallback)")
        _lib.load()
        # Launch mode of THIS env's step launches (mnv_params.pdl_prefetch, per call -- no process-wide state is touched):
        # a step fetches the map tables (goal, cores, obstacles) while the launch before it on the stream still drains
        # (include/marinenav_b200.h).  Its contract -- the launch right before mnv_step does not write those tables -- holds
        # for every sequence of this class: mnv_reset is always followed by mnv_observe, table uploads are host copies, and
        # tables_written() fences device-side edits (the facade's setters).  MNV_PDL in the environment overrides (lab use).
        self.pdl_prefetch = bool(pdl_prefetch) and os.environ.get("MNV_PDL") is None
        # How step_host ships the observation block to the host (same results either way):
        #   "dense"    one 6.8 MB device -> host copy of the f32 [E, obs_dim] block (the DMA engine writes the rows; no host
        #              work): ~200 us per 65 536-env step on a single-GPU box (the copy alone is 149 us at 48 GB/s);
        #   "compact"  ~2 MB packet (head + list of sonar returns, mnv_pack_obs) expanded by host threads (libmnv_host.so):
        #              3.5x fewer bytes over PCIe, paid with host CPU time: 170 us with 8 expander threads, 212 us with 4;
        #   "hybrid"   the first 62.5 % of the environments (MNV_HOST_HYBRID_FRACTION) travel compact, the rest dense, in that
        #              order: the host threads expand the packet WHILE the DMA engine still writes the dense rows, so the link
        #              and the host cores work at the same time (the packet's arrival is signalled through a sequence number
        #              in pinned memory): 170 - 184 us with 8 / 4 threads -- the choice when a rank has few cores;
        #   "auto"     MEASURED: after 64 calls, 33 step_host calls run 11 steps on each transport (identical results, whichever
        #              carries them), the fastest median stays -- until the next measurement 4096 calls later.  Which one wins depends on the host: cores per rank, how many
        #              GPUs share the memory system.  MNV_HOST_TRANSPORT overrides.
        t = os.environ.get("MNV_HOST_TRANSPORT", host_transport)
        self._auto_cal = None
        if t == "auto":
            t = "hybrid"
            order = ("compact", "hybrid", "dense")
            # measured after the first 64 calls (right after a reset of all environments the robots see ~2x the sonar returns
            # of a running rollout, which is not what the choice should be made on) and again every 4096 calls
            self._auto_cal = dict(order=order, per=11, calls=-64, every=4096, times={k: [] for k in order})
        if t not in ("dense", "compact", "hybrid"):
            raise ValueError(f"host_transport must be 'auto', 'dense', 'compact' or 'hybrid' (got {t!r})")
        self._transport = t
        self.host_transport_calibration = None      # {"dense": median seconds per step, "compact": ...} once "auto" has measured
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.sd = seed
        self.schedule = schedule
        # --- MarineNavEnv.__init__ (marinenav_env.py:40-73) ---
        self.width, self.height, self.r = 50.0, 50.0, 0.5
        self.v_rel_max, self.p = 1.0, 0.8
        self.v_range, self.obs_r_range, self.clear_r = [5, 10], [1, 3], 10.0
        self.reset_start_and_goal = True
        self.start, self.goal = np.array([5.0, 5.0]), np.array([45.0, 45.0])
        self.random_reset_state = True
        self.init_speed, self.init_theta = 0.0, math.pi / 4
        self.goal_dis = 2.0
        self.timestep_penalty, self.collision_penalty, self.goal_reward = -1.0, -50.0, 100.0
        self.discount = 0.99
        self.num_cores, self.num_obs, self.min_start_goal_dis = num_cores, num_obs, min_start_goal_dis
        self.total_timesteps = 0          # env steps taken over ALL environments (curriculum key, marinenav_env.py:89-98)
        self.global_step_multiplier = 1   # data-parallel training: world size, so the curriculum follows GLOBAL steps
        self.set_boundary = False
        # --- Robot / Sonar (robot.py:7-9,28-37) ---
        self.dt, self.N = 0.1, 10
        self.robot_r, self.max_speed = 0.8, 2.0
        self.a = np.array([-0.4, 0.0, 0.4])
        self.w = np.array([-np.pi / 6, 0.0, np.pi / 6])
        self.sonar_range, self.sonar_angle, self.num_beams = 10.0, 2 * np.pi / 3, int(num_beams)

        if schedule is not None:
            max_cores = max_cores or max(schedule["num_cores"])
            max_obstacles = max_obstacles or max(schedule["num_obstacles"])
        self.max_cores = int(max_cores if max_cores is not None else num_cores)
        self.max_obstacles = int(max_obstacles if max_obstacles is not None else num_obs)
        E = self.num_envs
        with torch.cuda.device(self.device):
            self.buf = env_ops.alloc_env_buffers(E, self.max_cores, self.max_obstacles, self.num_beams, self.device)
            self.buf["next_obs"] = torch.zeros_like(self.buf["obs"])
            self.rng_key = torch.zeros(E, 624, dtype=torch.int32, device=self.device)
            self.rng_pos = torch.zeros(E, dtype=torch.int32, device=self.device)
        self._pinned = None
        self.seed(seed)

    @property
    def host_transport(self):
        """The transport step_host uses now ("dense" | "compact" | "hybrid").  Assigning one of them fixes it (and ends the
        measurements of "auto"); assigning "auto" starts measuring again."""
        return self._transport

    @host_transport.setter
    def host_transport(self, t):
        if t == "auto":
            order = ("compact", "hybrid", "dense")
            self._auto_cal = dict(order=order, per=11, calls=-64, every=4096, times={k: [] for k in order})
            return
        if t not in ("dense", "compact", "hybrid"):
            raise ValueError(f"host_transport must be 'auto', 'dense', 'compact' or 'hybrid' (got {t!r})")
        self._transport, self._auto_cal = t, None

    # ---- gym-style surface -------------------------------------------------------------------------------------
    @property
    def obs_dim(self):
        return 4 + 2 * self.num_beams

    def get_state_space_dimension(self):
        return self.obs_dim

    def get_action_space_dimension(self):
        return 9

    def seed(self, seed):
        """MarineNavEnv.seed per environment: stream e = RandomState(seed + e)."""
        self.sd = seed
        seeds = (np.arange(self.num_envs, dtype=np.int64) + int(seed)) % (1 << 32)
        with torch.cuda.device(self.device):
            env_ops.seed(self.rng_key, self.rng_pos,
                         torch.from_numpy(seeds.astype(np.uint32).view(np.int32)).to(self.device))
        return [seed]

    def params(self):
        """mnv_params of the current attribute values (cached: rebuilt only when an attribute changed)."""
        key = (self.num_beams, self.dt, self.N, tuple(float(v) for v in self.a), tuple(float(v) for v in self.w), self.max_speed,
               self.robot_r, self.r, self.goal_dis, self.timestep_penalty, self.collision_penalty, self.goal_reward,
               self.sonar_range, self.sonar_angle, bool(self.set_boundary), self.width, self.height, self.pdl_prefetch)
        if getattr(self, "_params_key", None) == key:
            return self._params_cache
        p = self._build_params()
        self._params_key, self._params_cache = key, p
        return p

    def _build_params(self):
        p = _lib.default_params(self.num_beams)
        p.dt, p.n_substeps = self.dt, self.N
        for i in range(3):
            p.accel[i], p.yaw_rate[i] = float(self.a[i]), float(self.w[i])
        p.max_speed = self.max_speed
        p.k_drag = float(np.max(self.a)) / self.max_speed                      # Robot.compute_k, robot.py:51-52
        p.robot_r, p.core_r, p.goal_dis = self.robot_r, self.r, self.goal_dis
        p.timestep_penalty, p.collision_penalty, p.goal_reward = self.timestep_penalty, self.collision_penalty, self.goal_reward
        p.sonar_range, p.sonar_angle = self.sonar_range, self.sonar_angle
        p.set_boundary, p.width, p.height = int(self.set_boundary), self.width, self.height
        p.pdl_prefetch = int(self.pdl_prefetch)
        return p

    def reset_params(self):
        if self.schedule is not None:                                           # marinenav_env.py:89-98
            steps = np.array(self.schedule["timesteps"])
            idx = int((steps - self.total_timesteps <= 0).sum()) - 1
            self.num_cores = self.schedule["num_cores"][idx]
            self.num_obs = self.schedule["num_obstacles"][idx]
            self.min_start_goal_dis = self.schedule["min_start_goal_dis"][idx]
        key = (self.width, self.height, self.r, self.v_rel_max, self.p, tuple(self.v_range), tuple(self.obs_r_range), self.clear_r,
               bool(self.reset_start_and_goal), float(self.start[0]), float(self.start[1]), float(self.goal[0]), float(self.goal[1]),
               bool(self.random_reset_state), self.init_theta, self.init_speed, self.max_speed, int(self.num_cores),
               int(self.num_obs), float(self.min_start_goal_dis))
        if getattr(self, "_reset_key", None) == key:
            return self._reset_cache
        self._reset_key, self._reset_cache = key, self._build_reset_params()
        return self._reset_cache

    def _build_reset_params(self):
        r = _lib.default_reset_params()
        r.width, r.height, r.core_r, r.v_rel_max, r.p = self.width, self.height, self.r, self.v_rel_max, self.p
        r.v_range[0], r.v_range[1] = self.v_range
        r.obs_r_range[0], r.obs_r_range[1] = self.obs_r_range
        r.clear_r = self.clear_r
        r.reset_start_and_goal = int(self.reset_start_and_goal)
        r.start[0], r.start[1] = float(self.start[0]), float(self.start[1])
        r.goal[0], r.goal[1] = float(self.goal[0]), float(self.goal[1])
        r.random_reset_state, r.init_theta, r.init_speed = int(self.random_reset_state), self.init_theta, self.init_speed
        r.max_speed = self.max_speed
        r.num_cores, r.num_obs, r.min_start_goal_dis = int(self.num_cores), int(self.num_obs), float(self.min_start_goal_dis)
        return r

    def reset(self, mask=None):
        """MarineNavEnv.reset for all (or the masked) environments; returns the observation tensor f32 [E, obs_dim]."""
        with torch.cuda.device(self.device):
            env_ops.reset(self.buf, self.rng_key, self.rng_pos, self.reset_params(), mask=mask)
            env_ops.observe(self.buf, self.params(), mask=mask, velocity_from_state=True)
        return self.buf["obs"]

    def step(self, actions, auto_reset=True, trajectory=None):
        """actions: CUDA int32 [E].  Returns (obs, reward, done, info) device tensors.

        With auto_reset (the VecEnv convention) environments that finished are reset and ``obs`` holds the first
        observation of their next episode, while ``self.buf['next_obs']`` keeps the step's own (terminal) observation --
        what IQNAgent.learn stores as next_state before it calls reset() (agent.py:122-124,170)."""
        self.step_begin(actions, trajectory=trajectory)
        return self.step_finish(auto_reset=auto_reset)

    def step_begin(self, actions, trajectory=None):
        """First half of step(): the fused step kernel alone.  buf['next_obs'|'reward'|'done'|'info'] are final when it ends,
        so a learner may consume them on another stream while step_finish() resets the finished environments."""
        b = self.buf
        with torch.cuda.device(self.device):
            env_ops.step(b, self.params(), action=actions, obs=b["next_obs"], trajectory=trajectory)
        self.total_timesteps += self.num_envs * self.global_step_multiplier
        return b["next_obs"], b["reward"], b["done"], b["info"]

    def step_finish(self, auto_reset=True):
        """Second half of step(): obs <- next_obs, then (auto_reset) masked reset + re-observe of the finished environments."""
        b = self.buf
        with torch.cuda.device(self.device):
            b["obs"].copy_(b["next_obs"])
            if auto_reset:
                env_ops.reset(b, self.rng_key, self.rng_pos, self.reset_params(), mask=b["done"])
                env_ops.observe(b, self.params(), mask=b["done"], velocity_from_state=True)
        return b["obs"], b["reward"], b["done"], b["info"]

    # ---- host-buffer surface (numpy in / numpy out, pinned staging) ------------------------------------------------
    def _pin(self):
        if self._pinned is None:
            from . import _hostlib
            E, D = self.num_envs, self.obs_dim
            off = self.buf["packet_offsets"]
            o_r, o_d, o_i = self.buf["rdi_offsets"].tolist()
            # first tier of the packet copy: reward | done | info | head | count | hits of 6.25 % of the beam slots
            tier1 = min(off["hit_cap"], max(512, E * self.num_beams // 16))
            pk = torch.zeros(self.buf["host_packet"].numel(), dtype=torch.uint8).pin_memory()
            act = torch.zeros(E, dtype=torch.int32).pin_memory()
            rdi = pk[:self.buf["rdi_pack"].numel()]
            self._pinned = dict(action=act, action_np=act.numpy(),
                                obs=torch.zeros(E, D, dtype=torch.float32).pin_memory(),
                                packet=pk, tier1_bytes=off["vals"] + 8 * tier1, tier1=tier1,
                                head=pk[off["head"]:off["mask"]], mask=pk[off["mask"]:off["dir"]], dir=pk[off["dir"]:off["count"]],
                                count_np=pk[off["count"]:off["vals"]].view(torch.int32).numpy(), vals=pk[off["vals"]:],
                                rdi_pack=rdi, reward=rdi[o_r:o_r + 4 * E].view(torch.float32), done=rdi[o_d:o_d + E],
                                info=rdi[o_i:o_i + E])
            self._host_graphs = {}
            self._expander = _hostlib.Expander(E, D)
            self._host_dirty = False                  # True: the pinned dense block was written outside the expander
            self._hyb = None
        return self._pinned

    def _hybrid(self):
        """Buffers of the hybrid transport: a packet (mnv_pack_obs layout) for the first Ec environments, its pinned mirror,
        an expander for those rows and the pinned copy of the packet's count block the host polls."""
        if self._hyb is None:
            from . import _hostlib
            E, D, nb = self.num_envs, self.obs_dim, self.num_beams
            frac = min(1.0, max(0.0, float(os.environ.get("MNV_HOST_HYBRID_FRACTION", "0.625"))))
            Ec = E if frac >= 1.0 else min(E, max(32, int(E * frac) // 32 * 32))
            W, Gc = (nb + 31) // 32, (Ec + 31) // 32
            seg = lambda n: (n + 15) // 16 * 16
            o_mask = 16 * Ec
            o_dir = o_mask + seg(4 * W * Ec)
            o_cnt = o_dir + seg(4 * Gc)
            o_vals = o_cnt + 16
            cap = max(1024, Ec * nb // 8)
            tier1 = min(cap, max(512, Ec * nb // 16))
            dev = torch.zeros(o_vals + 8 * cap, dtype=torch.uint8, device=self.device)
            pk = torch.zeros(dev.numel(), dtype=torch.uint8).pin_memory()
            seq = torch.zeros(16, dtype=torch.uint8).pin_memory()
            self._hyb = dict(
                Ec=Ec, dev=dev, pin=pk, cap=cap, tier1=tier1, tier1_bytes=o_vals + 8 * tier1, o_vals=o_vals,
                head=dev[:o_mask].view(torch.float32).view(Ec, 4), mask=dev[o_mask:o_mask + 4 * W * Ec].view(torch.int32).view(Ec, W),
                dir=dev[o_dir:o_dir + 4 * Gc].view(torch.int32), count=dev[o_cnt:o_vals].view(torch.int32),
                vals=dev[o_vals:].view(torch.float32).view(cap, 2),
                p_head=pk[:o_mask], p_mask=pk[o_mask:o_dir], p_dir=pk[o_dir:o_cnt], p_vals=pk[o_vals:],
                count_np=pk[o_cnt:o_vals].view(torch.int32).numpy(), seq_pin=seq, seq_np=seq.view(torch.int32).numpy().view(np.uint32),
                expander=_hostlib.Expander(Ec, D), next_seq=1, valid=False)
        return self._hyb

    def _capture_host_step(self, auto_reset):
        """One CUDA graph for the whole host-boundary step.  Stream A: fused step (actions read zero-copy from the pinned host
        buffer) -> (auto-reset: masked reset -> masked re-observe); stream B, forked right behind the step kernel: the
        step's own observation block goes through mnv_pack_obs (head of every row + the list of sonar returns, ~4 % of the
        beam slots) and ships together with reward | done | info in ONE device -> host copy of ~2 MB instead of the dense
        7.2 MB.  A joins B and overwrites the rows of the re-observed environments directly in the pinned dense array
        (mnv_scatter_rows_host, zero-copy stores); after the graph the native expander (libmnv_host.so) brings the dense
        array up to date from the packet, skipping those rows.  The copy runs under the reset, the host pays one graph
        launch and touches only the slots that change."""
        pin, b = self._pin(), self.buf
        if self.host_transport == "hybrid":
            self._hybrid()                          # allocate OUTSIDE the capture (a captured torch.zeros would re-zero it at every replay)
        params = self.params()
        rp = self.reset_params() if auto_reset else None
        cur = torch.cuda.current_stream()
        sa, sb = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
        g = torch.cuda.CUDAGraph()
        n1 = pin["tier1_bytes"]
        sa.wait_stream(cur)
        with torch.cuda.stream(sa):
            with torch.cuda.graph(g, stream=sa):
                env_ops.step(b, params, action=pin["action"], obs=b["next_obs"])    # actions read zero-copy from the pinned buffer
                sb.wait_stream(sa)
                with torch.cuda.stream(sb):
                    if self.host_transport == "compact":
                        env_ops.pack_obs(b["next_obs"], b["packet_head"], b["packet_mask"], b["packet_dir"], b["packet_count"], b["packet_vals"])
                        pin["packet"][:n1].copy_(b["host_packet"][:n1], non_blocking=True)
                    elif self.host_transport == "hybrid":
                        # compact part first (reward | done | info, then the packet of environments [0, Ec)), then a copy of the
                        # packet's count block -- its sequence number tells the polling host that the packet has landed --
                        # and only then the dense rows of environments [Ec, E): the host expands under that copy
                        h = self._hybrid()
                        Ec, nh = h["Ec"], h["tier1_bytes"]
                        env_ops.pack_obs(b["next_obs"][:Ec], h["head"], h["mask"], h["dir"], h["count"], h["vals"])
                        pin["rdi_pack"].copy_(b["rdi_pack"], non_blocking=True)
                        h["pin"][:nh].copy_(h["dev"][:nh], non_blocking=True)
                        h["seq_pin"].copy_(h["count"].view(torch.uint8), non_blocking=True)
                        if Ec < self.num_envs:
                            pin["obs"][Ec:].copy_(b["next_obs"][Ec:], non_blocking=True)
                    else:
                        pin["obs"].copy_(b["next_obs"], non_blocking=True); pin["rdi_pack"].copy_(b["rdi_pack"], non_blocking=True)
                b["obs"].copy_(b["next_obs"])
                if auto_reset:
                    env_ops.reset(b, self.rng_key, self.rng_pos, rp, mask=b["done"])
                    env_ops.observe(b, params, mask=b["done"], velocity_from_state=True)
                sa.wait_stream(sb)
                if auto_reset:
                    env_ops.scatter_rows_host(b["done"], b["obs"], pin["obs"])      # dense rows of the re-observed environments
        cur.wait_stream(sa)
        return g, (sa, sb)

    def _refresh_host_dense(self):
        """Dense device -> host copy of the current observation block + rebuild of the expander's bookkeeping (the slow
        path: first use after a dense write, hit-list overflow)."""
        pin = self._pin()
        pin["obs"].copy_(self.buf["obs"])
        torch.cuda.current_stream(self.device).synchronize()
        self._expander.rescan(pin["obs"].data_ptr())
        self._host_dirty = False
        if self._hyb is not None:
            self._hyb["valid"] = False

    def step_host(self, actions, auto_reset=True, graph=True):
        """numpy int actions [E] -> (obs f32 [E,D], reward f32 [E], done bool [E], info u8 [E]) numpy views of pinned buffers
        (valid until the next call; read-only: the observation array is updated in place from step to step).  obs holds the
        first observation of the next episode for finished environments, like step().  graph=False runs the same operations
        eagerly on one stream with a dense copy (the parity reference of the graph path)."""
        cal = self._auto_cal
        if cal is None or not graph:
            return self._step_host(actions, auto_reset, graph)
        # host_transport="auto" with several ranks on the node: time both transports under the load the ranks produce together
        k, per = cal["calls"], cal["per"]
        if k < 0:                                                  # not yet (or between two measurements): the current choice
            cal["calls"] = k + 1
            return self._step_host(actions, auto_reset, graph)
        self._transport = cal["order"][k // per]
        t0 = time.perf_counter()
        out = self._step_host(actions, auto_reset, graph)
        if k % per:                                                # the first call of a block captures the graph / rescans
            cal["times"][self.host_transport].append(time.perf_counter() - t0)
        cal["calls"] = k + 1
        if cal["calls"] == per * len(cal["order"]):
            med = {t: sorted(v)[len(v) // 2] for t, v in cal["times"].items()}
            self.host_transport_calibration = med
            self._transport = min(med, key=med.get)
            cal["calls"], cal["times"] = -cal["every"], {t: [] for t in cal["order"]}
        return out

    def _step_host(self, actions, auto_reset, graph):
        pin = self._pin()
        np.copyto(pin["action_np"], np.asarray(actions), casting="unsafe")       # 9 us; torch's CPU copy_ costs 14 - 500 us here
        with torch.cuda.device(self.device):
            if not graph:
                self.buf["action"].copy_(pin["action"], non_blocking=True)
                obs, reward, done, info = self.step(self.buf["action"], auto_reset=auto_reset)
                pin["obs"].copy_(obs, non_blocking=True); pin["rdi_pack"].copy_(self.buf["rdi_pack"], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                self._host_dirty = True
                if self._hyb is not None:
                    self._hyb["valid"] = False
                return pin["obs"].numpy(), pin["reward"].numpy(), pin["done"].numpy().view(np.bool_), pin["info"].numpy()
            self.params()
            if auto_reset:
                self.reset_params()
            key = (self._params_key, self._reset_key if auto_reset else None, bool(auto_reset), self.host_transport)
            entry = self._host_graphs.get(key)
            if entry is None:
                self._host_graphs.clear()                         # parameters changed: the old graph holds stale constants
                entry = self._host_graphs[key] = self._capture_host_step(auto_reset)
            if self.host_transport == "hybrid":
                self._step_host_hybrid(entry[0], pin, auto_reset)
                return pin["obs"].numpy(), pin["reward"].numpy(), pin["done"].numpy().view(np.bool_), pin["info"].numpy()
            compact = self.host_transport == "compact"
            if compact and self._host_dirty:
                self._expander.rescan(pin["obs"].data_ptr())
                self._host_dirty = False
            if self._hyb is not None:
                self._hyb["valid"] = False                        # the rows of the hybrid expander are rewritten behind its back
            entry[0].replay()
            self.total_timesteps += self.num_envs * self.global_step_multiplier
            torch.cuda.current_stream().synchronize()
            if not compact:
                self._host_dirty = True
                return pin["obs"].numpy(), pin["reward"].numpy(), pin["done"].numpy().view(np.bool_), pin["info"].numpy()
            n_hits = int(pin["count_np"][0])
            if n_hits > self.buf["packet_offsets"]["hit_cap"]:
                self._refresh_host_dense()                        # more returns than the list holds: dense block this once
            else:
                if n_hits > pin["tier1"]:                         # the rest of the list (rare: > 6.25 % of the beam slots)
                    lo, hi = pin["tier1_bytes"], self.buf["packet_offsets"]["vals"] + 8 * n_hits
                    pin["packet"][lo:hi].copy_(self.buf["host_packet"][lo:hi], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                self._expander.expand(pin["obs"].data_ptr(), pin["head"].data_ptr(), pin["done"].data_ptr() if auto_reset else None,
                                      pin["mask"].data_ptr(), pin["dir"].data_ptr(), pin["vals"].data_ptr())
        return pin["obs"].numpy(), pin["reward"].numpy(), pin["done"].numpy().view(np.bool_), pin["info"].numpy()

    def _step_host_hybrid(self, graph, pin, auto_reset):
        """Replay + host half of the hybrid transport: wait (polling the sequence number in pinned memory) until the packet of
        environments [0, Ec) has landed, expand it while the dense rows of [Ec, E) are still in flight, then wait for the
        stream and pick up the rows the GPU re-observed."""
        h = self._hybrid()
        obs_ptr = pin["obs"].data_ptr()
        if not h["valid"]:                                        # first use / another transport wrote the rows: rebuild the bookkeeping
            h["expander"].rescan(obs_ptr)
            h["valid"] = True
        self._host_dirty = True                                   # (the full-size expander of "compact" loses track of the rows)
        expect = h["next_seq"]
        trace = h.get("trace")                                    # lab (scripts/lab/auto_transport_trace.py): phase stamps of this call
        t_a = time.perf_counter()
        graph.replay()
        self.total_timesteps += self.num_envs * self.global_step_multiplier
        seq_np, spins = h["seq_np"], 0
        while int(seq_np[3]) != expect:
            spins += 1
            if spins > 200000:                                    # ~0.1 s: something else is wrong -- let the stream tell
                torch.cuda.current_stream().synchronize()
                if int(seq_np[3]) != expect:
                    raise _lib.MarinenavError(f"step_host(hybrid): packet sequence {int(seq_np[3])}, expected {expect}")
        h["next_seq"] = (expect + 1) & 0xFFFFFFFF
        t_b = time.perf_counter()
        n_hits = int(h["count_np"][0])
        skip = pin["done"].data_ptr() if auto_reset else None
        if n_hits > h["cap"]:                                     # more returns than the list holds: dense block this once
            torch.cuda.current_stream().synchronize()
            self._refresh_host_dense()
            return
        if n_hits > h["tier1"]:                                   # the rest of the list (rare: > 6.25 % of the beam slots)
            lo, hi = h["tier1_bytes"], h["o_vals"] + 8 * n_hits
            h["pin"][lo:hi].copy_(h["dev"][lo:hi], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h["expander"].expand_early(obs_ptr, h["p_head"].data_ptr(), skip, h["p_mask"].data_ptr(), h["p_dir"].data_ptr(),
                                   h["p_vals"].data_ptr())
        t_c = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        t_d = time.perf_counter()
        if auto_reset:
            h["expander"].rescan_skipped(obs_ptr, skip)
        if trace is not None:
            trace.append((spins, n_hits, t_b - t_a, t_c - t_b, t_d - t_c, time.perf_counter() - t_d))

    def tables_written(self):
        """Call after writing buf['goal'|'cores'|'obstacles'] with a device-side kernel (e.g. indexed assignment) if a step
        may follow directly: the step kernel reads these tables ahead of its stream dependency ("pdl" = 2)."""
        torch.cuda.current_stream(self.device).synchronize()

    def reset_host(self):
        pin = self._pin()
        self.reset()
        with torch.cuda.device(self.device):
            self._refresh_host_dense()
        return pin["obs"].numpy()

    def host_api_description(self):
        nthr = self._expander.n_threads if self._pinned else "?"
        how = {"compact": "the observation block travels as head + list of sonar returns (mnv_pack_obs) and is expanded into the dense "
                          f"numpy array by libmnv_host.so on {nthr} host threads",
               "hybrid": f"environments [0, {self._hybrid()['Ec'] if self._pinned else '?'}) travel as head + list of sonar returns and are "
                         f"expanded by libmnv_host.so on {nthr} host threads WHILE the dense rows of the other environments are still "
                         "being copied",
               "dense": "dense device -> host copy of the observation block under the masked reset"}[self.host_transport]
        return ("VecMarineNavEnv.step_host (numpy in/out, pinned staging, auto-reset; one two-stream CUDA graph per step; "
                f"host_transport={self.host_transport}: {how})")

    def h2d_bytes_per_step(self):
        return self.num_envs * 4

    def d2h_bytes_per_step(self):
        self._pin()
        # the packet copy (reward | done | info | head | count | first tier of the hit list); the zero-copy rows of the
        # re-observed environments (a few hundred x obs_dim x 4 bytes) come on top
        if self.host_transport == "compact":
            return self._pinned["tier1_bytes"]
        if self.host_transport == "hybrid":
            h = self._hybrid()
            return h["tier1_bytes"] + 16 + (self.num_envs - h["Ec"]) * self.obs_dim * 4 + self._pinned["rdi_pack"].numel()
        return self.num_envs * self.obs_dim * 4 + self._pinned["rdi_pack"].numel()

    def close(self):
        pass

    # ---- Robot helpers the callers reach through env.robot.* ---------------------------------------------------------
    def compute_action_energy_cost(self, action):
        """Robot.compute_action_energy_cost (robot.py:72-77): |a| / max(a) + |w| / max(w)."""
        a, w = self.a[int(action) // 3], self.w[int(action) % 3]
        return float(abs(a / np.max(self.a)) + abs(w / np.max(self.w)))

    # ---- evaluation maps (reset_with_eval_config, marinenav_env.py:467-555) ------------------------------------------
    @classmethod
    def from_eval_configs(cls, cfgs, device="cuda:0", seed=0):
        """One environment per eval-config dict (schema of eval_config.json), restarted at its recorded start pose."""
        num_beams = cfgs[0]["robot"]["sonar"]["num_beams"]
        max_c = max(1, max(len(c["env"]["cores"]["positions"]) for c in cfgs))
        max_o = max(1, max(len(c["env"]["obstacles"]["positions"]) for c in cfgs))
        env = cls(len(cfgs), seed=seed, device=device, num_cores=max_c, num_obs=max_o, num_beams=num_beams,
                  max_cores=max_c, max_obstacles=max_o)
        env.load_eval_configs(cfgs)
        return env

    def load_eval_configs(self, cfgs):
        """Per-env maps / start poses from eval-config dicts; the scalar env / robot / sonar constants come from cfgs[0]
        (they are identical across the reference's evaluation maps)."""
        assert len(cfgs) == self.num_envs
        c0, E = cfgs[0], self.num_envs
        e, r = c0["env"], c0["robot"]
        for i, c in enumerate(cfgs[1:], 1):                    # one kernel launch = one set of scalar constants
            for k in ("width", "height", "r", "goal_dis", "timestep_penalty", "collision_penalty", "goal_reward", "discount"):
                if c["env"][k] != e[k]:
                    raise ValueError(f"load_eval_configs: env constant '{k}' of config {i} differs from config 0 ({c['env'][k]} != {e[k]})")
            for k in ("dt", "N", "r", "max_speed", "a", "w"):
                if c["robot"][k] != r[k]:
                    raise ValueError(f"load_eval_configs: robot constant '{k}' of config {i} differs from config 0")
            if c["robot"]["sonar"] != r["sonar"]:
                raise ValueError(f"load_eval_configs: sonar constants of config {i} differ from config 0")
        self.width, self.height, self.r = e["width"], e["height"], e["r"]
        self.v_rel_max, self.p = e["v_rel_max"], e["p"]
        self.v_range, self.obs_r_range, self.clear_r = list(e["v_range"]), list(e["obs_r_range"]), e["clear_r"]
        self.goal_dis = e["goal_dis"]
        self.timestep_penalty, self.collision_penalty, self.goal_reward = e["timestep_penalty"], e["collision_penalty"], e["goal_reward"]
        self.discount = e["discount"]
        self.dt, self.N, self.robot_r, self.max_speed = r["dt"], r["N"], r["r"], r["max_speed"]
        self.a, self.w = np.array(r["a"]), np.array(r["w"])
        self.sonar_range, self.sonar_angle = r["sonar"]["range"], r["sonar"]["angle"]
        assert r["sonar"]["num_beams"] == self.num_beams, "allocate the env with the config's num_beams"
        mc, mo = self.max_cores, self.max_obstacles
        state = np.zeros((4, E)); goal = np.zeros((2, E)); cores = np.zeros((3 * mc, E)); obst = np.zeros((3 * mo, E))
        for i, cfg in enumerate(cfgs):
            ce, cr = cfg["env"], cfg["robot"]
            state[:, i] = [ce["start"][0], ce["start"][1], cr["init_theta"], cr["init_speed"]]
            goal[:, i] = ce["goal"]
            pos, cw, G = ce["cores"]["positions"], ce["cores"]["clockwise"], ce["cores"]["Gamma"]
            assert len(pos) <= mc and len(ce["obstacles"]["positions"]) <= mo
            for k in range(len(pos)):
                cores[k, i], cores[mc + k, i], cores[2 * mc + k, i] = pos[k][0], pos[k][1], (G[k] if cw[k] else -G[k])
            for k, (pp, rr) in enumerate(zip(ce["obstacles"]["positions"], ce["obstacles"]["r"])):
                obst[k, i], obst[mo + k, i], obst[2 * mo + k, i] = pp[0], pp[1], rr
        b = self.buf
        for name, arr in (("state", state), ("goal", goal), ("cores", cores), ("obstacles", obst), ("start_pose", state)):
            b[name].copy_(torch.from_numpy(arr))
        b["episode_step"].zero_()
        b["n_placed"][0].copy_(torch.tensor([len(c["env"]["cores"]["positions"]) for c in cfgs], dtype=torch.uint8))
        b["n_placed"][1].copy_(torch.tensor([len(c["env"]["obstacles"]["positions"]) for c in cfgs], dtype=torch.uint8))

    def observe_all(self):
        """robot.reset_state + get_observation after load_eval_configs (marinenav_env.py:551-555)."""
        with torch.cuda.device(self.device):
            env_ops.observe(self.buf, self.params(), velocity_from_state=True)
        return self.buf["obs"]

    def restart_episodes(self, mask=None):
        """Put (masked) environments back on their recorded start pose without drawing a new map."""
        b = self.buf
        if mask is None:
            b["state"].copy_(b["start_pose"]); b["episode_step"].zero_()
        else:
            m = mask.bool()
            b["state"][:, m] = b["start_pose"][:, m]; b["episode_step"][m] = 0
        with torch.cuda.device(self.device):
            env_ops.observe(b, self.params(), mask=mask, velocity_from_state=True)
        return b["obs"]

    def episode_data(self, i=0, action_history=(), trajectory=()):
        """MarineNavEnv.episode_data (marinenav_env.py:557-622) of environment i, the eval_config.json schema."""
        b = self.buf
        nc, no = int(b["n_placed"][0, i]), int(b["n_placed"][1, i])
        cores = b["cores"][:, i].cpu().numpy(); obst = b["obstacles"][:, i].cpu().numpy()
        sp = b["start_pose"][:, i].cpu().numpy(); goal = b["goal"][:, i].cpu().numpy()
        mc, mo = self.max_cores, self.max_obstacles
        ep = {"env": {"seed": int(self.sd) + int(i), "width": self.width, "height": self.height, "r": self.r, "v_rel_max": self.v_rel_max,
                      "p": self.p, "v_range": list(self.v_range), "obs_r_range": list(self.obs_r_range), "clear_r": self.clear_r,
                      "start": [float(sp[0]), float(sp[1])], "goal": [float(goal[0]), float(goal[1])], "goal_dis": self.goal_dis,
                      "timestep_penalty": self.timestep_penalty, "collision_penalty": self.collision_penalty,
                      "goal_reward": self.goal_reward, "discount": self.discount,
                      "cores": {"positions": [[float(cores[k]), float(cores[mc + k])] for k in range(nc)],
                                "clockwise": [int(cores[2 * mc + k] > 0) for k in range(nc)],
                                "Gamma": [float(abs(cores[2 * mc + k])) for k in range(nc)]},
                      "obstacles": {"positions": [[float(obst[k]), float(obst[mo + k])] for k in range(no)],
                                    "r": [float(obst[2 * mo + k]) for k in range(no)]}},
              "robot": {"dt": self.dt, "N": self.N, "length": 1.0, "width": 0.5, "r": self.robot_r, "max_speed": self.max_speed,
                        "a": [float(x) for x in self.a], "w": [float(x) for x in self.w],
                        "init_theta": float(sp[2]), "init_speed": float(sp[3]),
                        "sonar": {"range": self.sonar_range, "angle": self.sonar_angle, "num_beams": self.num_beams},
                        "action_history": [int(a) for a in action_history],
                        "trajectory": [[float(p[0]), float(p[1])] for p in trajectory]}}
        return ep
