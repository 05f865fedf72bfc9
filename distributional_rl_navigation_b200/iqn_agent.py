"""IQNAgent: host-side mirror of thirdparty/IQN/agent.py whose act / train run in the CUDA kernels.

Drop-in surface (same names, argument meaning and defaults as agent.py:12-29,86-398): IQNAgent(...), learn, act,
act_adaptive, act_eval, act_adaptive_eval, adjust_cvar, train, soft_update, evaluation, load_model, linear_eps and the
attributes qnetwork_local / qnetwork_target / memory / current_timestep / learning_timestep / eval_*.  The single-env
loop follows the reference's random choices (python `random` for epsilon-greedy and replay sampling, torch CPU generator
for the taus, target taus drawn before local taus).  On top of it: act_batch / learn_vec drive a VecMarineNavEnv with
everything (observations, replay, taus, epsilon-greedy) resident on the GPU, and train() all-reduces the flat gradient
when torch.distributed is initialised (one NCCL all-reduce of 35 785 floats per update).
"""
import os
import random
import warnings

import numpy as np
import torch

from . import _lib, distributed as mdist, env_ops, iqn_ops
from .iqn_model import ObsEncoder
from .replay_buffer import DeviceReplayBuffer, ReplayBuffer


def _resolve_device(device):
    dev = torch.device(device)
    if dev.type == "cpu":
        # the reference's default is device="cpu" (train_IQN_model.py -D); this implementation only computes on the GPU
        warnings.warn("IQNAgent(device='cpu'): this implementation has no CPU path, using cuda:%d" % torch.cuda.current_device()
                      if torch.cuda.is_available() else "IQNAgent needs a CUDA device")
        if not torch.cuda.is_available():
            raise _lib.MarinenavError("IQNAgent needs a CUDA device (there is no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class _AdamState:
    """Stand-in for torch.optim.Adam(qnetwork_local.parameters(), lr) (agent.py:66): moments as flat device vectors."""

    def __init__(self, n, lr, device):
        self.lr, self.betas, self.eps = lr, (0.9, 0.999), 1e-8
        self.m = torch.zeros(n, dtype=torch.float32, device=device)
        self.v = torch.zeros(n, dtype=torch.float32, device=device)
        self.step_count = 0

    def zero_grad(self):
        pass

    def state_dict(self):
        return dict(lr=self.lr, betas=self.betas, eps=self.eps, step=self.step_count, exp_avg=self.m.clone(), exp_avg_sq=self.v.clone())

    def load_state_dict(self, sd):
        self.lr, self.betas, self.eps, self.step_count = sd["lr"], tuple(sd["betas"]), sd["eps"], int(sd["step"])
        self.m.copy_(sd["exp_avg"]); self.v.copy_(sd["exp_avg_sq"])


class IQNAgent:
    def __init__(self, state_size, action_size, layer_size=64, n_step=1, BATCH_SIZE=32, BUFFER_SIZE=1_000_000, LR=1e-4,
                 TAU=1.0, GAMMA=0.99, UPDATE_EVERY=4, learning_starts=10000, target_update_interval=10000,
                 exploration_fraction=0.1, initial_eps=1.0, final_eps=0.05, device="cpu", seed=0):
        self.state_size, self.action_size = state_size, action_size
        self.device = _resolve_device(device)
        self.LR, self.TAU, self.GAMMA = LR, TAU, GAMMA
        self.UPDATE_EVERY, self.BATCH_SIZE, self.BUFFER_SIZE, self.n_step = UPDATE_EVERY, BATCH_SIZE, BUFFER_SIZE, n_step
        self.learning_starts, self.target_update_interval = learning_starts, target_update_interval
        self.exploration_fraction, self.initial_eps, self.final_eps = exploration_fraction, initial_eps, final_eps
        self.seed = seed

        self.qnetwork_local = ObsEncoder(state_size, action_size, seed, self.device)       # agent.py:63-64: same seed twice
        self.qnetwork_target = ObsEncoder(state_size, action_size, seed, self.device)
        self.optimizer = _AdamState(iqn_ops.N_PARAMS, LR, self.device)
        self.memory = ReplayBuffer(BUFFER_SIZE, BATCH_SIZE, self.device, seed, GAMMA, n_step, state_size)   # agent.py:70
        self.current_timestep = 0
        self.learning_timestep = 0

        self.eval_timesteps = dict(greedy=[], adaptive=[])
        self.eval_actions = dict(greedy=[], adaptive=[])
        self.eval_rewards = dict(greedy=[], adaptive=[])
        self.eval_successes = dict(greedy=[], adaptive=[])
        self.eval_times = dict(greedy=[], adaptive=[])
        self.eval_energies = dict(greedy=[], adaptive=[])

        n = iqn_ops.N_PARAMS
        self._grad = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._grad_norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._scratch = None
        self.fused_tail = os.environ.get("MNV_FUSED_TAIL", "1") != "0"   # iqn_update_tail (one launch behind the backward) vs reduce + all_reduce + clip_adam
        self.keep_grad = False                 # True: the fused tail also leaves the averaged gradient in self._grad
        self._tail = None
        self.device_memory = None              # DeviceReplayBuffer of the vectorised trainer
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed) + 12345)

    # ---- persistence ---------------------------------------------------------------------------------------------------
    def load_model(self, path, device=None):
        """agent.py:86-92: both networks from network_params.pth, fresh Adam."""
        dev = self.device if device is None else _resolve_device(device)
        self.qnetwork_local = ObsEncoder.load(path, dev)
        self.qnetwork_target = ObsEncoder.load(path, dev)
        self.optimizer = _AdamState(iqn_ops.N_PARAMS, self.LR, dev)

    # ---- schedules -------------------------------------------------------------------------------------------------------
    def linear_eps(self, total_timesteps):
        """agent.py:176-183"""
        progress = self.current_timestep / total_timesteps
        if progress < self.exploration_fraction:
            r = progress / self.exploration_fraction
            return self.initial_eps + r * (self.final_eps - self.initial_eps)
        return self.final_eps

    # ---- acting: single observation (the reference's interface) ---------------------------------------------------------
    def _qvals(self, state, cvar):
        x = torch.from_numpy(np.asarray(state)).float().unsqueeze(0)
        self.qnetwork_local.eval()
        with torch.no_grad():
            action_values = self.qnetwork_local.get_qvals(x, cvar)
        self.qnetwork_local.train()
        return action_values

    def act(self, state, eps, cvar=1.0):
        """agent.py:186-205: K = 32 quantile samples, epsilon-greedy with python `random`."""
        action_values = self._qvals(state, cvar)
        if random.random() > eps:
            action = np.argmax(action_values.cpu().data.numpy())
        else:
            action = random.choice(np.arange(self.action_size))
        return action

    def act_adaptive(self, state, eps):
        """agent.py:207-215"""
        cvar = self.adjust_cvar(state)
        return self.act(state, eps, cvar), cvar

    def act_eval(self, state, eps=0.0, cvar=1.0):
        """agent.py:217-237: also returns the quantiles [1,32,9] and taus [1,32,1]."""
        x = torch.from_numpy(np.asarray(state)).float().unsqueeze(0)
        self.qnetwork_local.eval()
        with torch.no_grad():
            quantiles, taus = self.qnetwork_local.forward(x, self.qnetwork_local.K, cvar)
            action_values = quantiles.mean(dim=1)
        self.qnetwork_local.train()
        if random.random() > eps:
            action = np.argmax(action_values.cpu().data.numpy())
        else:
            action = random.choice(np.arange(self.action_size))
        return action, quantiles.cpu().data.numpy(), taus.cpu().data.numpy()

    def act_adaptive_eval(self, state, eps=0.0):
        """agent.py:239-247"""
        cvar = self.adjust_cvar(state)
        return self.act_eval(state, eps, cvar), cvar

    def adjust_cvar(self, state):
        """agent.py:249-267: CVaR = min(1, closest sonar return / 10); a beam with |x|,|y| < 1e-3 is 'no return'."""
        pts = np.asarray(state)[4:].reshape(-1, 2)
        seen = ~((np.abs(pts[:, 0]) < 1e-3) & (np.abs(pts[:, 1]) < 1e-3))
        closest = np.inf
        for p in pts[seen]:
            closest = min(closest, np.linalg.norm(p))
        return closest / 10.0 if closest < 10.0 else 1.0

    # ---- acting: whole env batch on the device -----------------------------------------------------------------------------
    def adjust_cvar_batch(self, obs):
        """adjust_cvar for obs f32 [E, 26] on the device -> f32 [E]."""
        pts = obs[:, 4:].reshape(obs.shape[0], -1, 2)
        seen = ~((pts[..., 0].abs() < 1e-3) & (pts[..., 1].abs() < 1e-3))
        d = torch.linalg.vector_norm(pts, dim=-1)
        closest = torch.where(seen, d, torch.full_like(d, float("inf"))).min(dim=1).values
        return torch.where(closest < 10.0, closest / 10.0, torch.ones_like(closest)).contiguous()

    def act_batch(self, obs, eps, cvar=1.0, adaptive=False, tensor_cores=True):
        """Epsilon-greedy actions (int32 [E]) for a device batch of observations; K = 32 taus per env drawn on the device.
        tensor_cores=True: ONE call of iqn_act_tc_sample (tcgen05 kernel, bf16 operands -- only the argmax is consumed; the
        taus, the epsilon-greedy coin and the random action come from the kernel's Philox stream keyed by the agent seed
        and an act counter, the adaptive CVaR level from its pre-pass); False: the fp32 parity kernel with torch-drawn taus."""
        net = self.qnetwork_local
        E = obs.shape[0]
        with torch.cuda.device(self.device):
            if tensor_cores:
                self._act_calls = getattr(self, "_act_calls", 0) + 1
                action, _, _ = iqn_ops.act_tc_sample(net.flat, net.packed_tc, obs, eps, int(self.seed) + 0x5EED0000, self._act_calls,
                                                     cvar=cvar, adaptive=adaptive)
                return action
            taus = torch.rand(E, net.K, device=self.device, generator=self.gen)
            cv = self.adjust_cvar_batch(obs) if adaptive else cvar
            _, _, greedy = iqn_ops.forward(net.flat, net.packed, obs, taus, cv, want_quantiles=False, want_greedy=True)
            if eps <= 0.0:
                return greedy
            explore = torch.rand(E, device=self.device, generator=self.gen) <= eps           # agent.py:200: greedy iff random() > eps
            rand_a = torch.randint(0, self.action_size, (E,), device=self.device, generator=self.gen, dtype=torch.int32)
            return torch.where(explore, rand_a, greedy)

    def _act_with(self, enc_params, packed_tc, obs, eps, cvar=1.0, adaptive=False):
        """act_batch's tensor-core path on an explicit weight snapshot (encoder parameters + bf16 tiles)."""
        self._act_calls = getattr(self, "_act_calls", 0) + 1
        with torch.cuda.device(self.device):
            action, _, _ = iqn_ops.act_tc_sample(enc_params, packed_tc, obs, eps, int(self.seed) + 0x5EED0000, self._act_calls,
                                                 cvar=cvar, adaptive=adaptive)
        return action

    # ---- learning --------------------------------------------------------------------------------------------------------
    def train(self, experiences, taus=None):
        """agent.py:269-304.  experiences = (states [B,26], actions [B,1] int64, rewards [B,1], next_states, dones [B,1]).
        taus (optional) = (taus_target, taus_local) f32 [B,8]; by default they are drawn like the reference does (target
        first).  Returns the loss as a numpy scalar.  With torch.distributed initialised the flat gradient is summed over
        ranks and divided by the world size before the clip + Adam step (identical on every rank)."""
        states, actions, rewards, next_states, dones = experiences
        dev = self.device
        B = states.shape[0]
        f = lambda t: t.to(dev, torch.float32).reshape(B, -1).contiguous()
        states, next_states = f(states), f(next_states)
        rewards, dones = f(rewards).reshape(B), f(dones).reshape(B)
        actions = actions.to(dev, torch.int64).reshape(B).contiguous()
        self.optimizer.zero_grad()
        if taus is None:
            taus_t = self.qnetwork_target.draw_taus(B, 8)          # Q9: the target forward draws first (agent.py:279)
            taus_l = self.qnetwork_local.draw_taus(B, 8)
        else:
            taus_t, taus_l = (t.to(dev, torch.float32).contiguous() for t in taus)
        with torch.cuda.device(dev):
            self._update((states, actions, rewards, next_states, dones), (taus_t, taus_l))
        return self._loss.detach().cpu().numpy()[0]

    def train_async(self, experiences, taus):
        """train() without the device->host read of the loss (for CUDA-graph / benchmark loops). Returns the loss tensor."""
        self._update(experiences, taus)
        return self._loss

    def _update(self, experiences, taus, ctl=None):
        """loss.backward() ... optimizer.step() (agent.py:296-300) in two launches: the fused per-tile kernel (iqn_loss_partials)
        and the fused tail (iqn_update_tail: fixed-order sum of the tile partials, one-shot all-reduce of the gradient over peer
        memory when torch.distributed is initialised, clip_grad_norm_ 0.5, Adam, refresh of the kernel-side weight copies).
        `fused_tail = False` (or GPUs without peer access) keeps the three-launch path with torch.distributed's all-reduce."""
        states, actions, rewards, next_states, dones = experiences
        B = states.shape[0]
        need = iqn_ops.train_scratch_floats(B)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.float32, device=self.device)
        L, T, opt = self.qnetwork_local, self.qnetwork_target, self.optimizer
        gamma_n = float(self.GAMMA ** self.n_step)
        if self.fused_tail and self._tail is None:
            self._tail = iqn_ops.UpdateTail(self.device)
            if self._tail.peer_error is not None:
                warnings.warn("IQNAgent: peer-memory gradient exchange unavailable (%s); using torch.distributed all_reduce" % self._tail.peer_error)
        if ctl is not None:
            # CUDA-graph replays (learn_vec(graph=True)): Adam's bias corrections come from the device control block and the
            # caller advances opt.step_count per replay
            if not (self.fused_tail and self._tail.world == mdist.world_size()):
                raise _lib.MarinenavError("the captured update needs the fused tail (iqn_update_tail) on every rank")
            iqn_ops.loss_partials(L.flat, L.packed, T.flat, T.packed, states, actions, rewards, next_states, dones, taus[0], taus[1],
                                  gamma_n, self._scratch)
            self._tail.step(L.flat, opt.m, opt.v, L.packed, L.packed_tc, self._scratch, B, 1, loss=self._loss,
                            grad=self._grad if self.keep_grad else None, grad_norm=self._grad_norm, max_norm=0.5,
                            beta1=opt.betas[0], beta2=opt.betas[1], eps=opt.eps, ctl=ctl)
            return
        if self.fused_tail and self._tail.world == mdist.world_size():
            iqn_ops.loss_partials(L.flat, L.packed, T.flat, T.packed, states, actions, rewards, next_states, dones, taus[0], taus[1],
                                  gamma_n, self._scratch)
            opt.step_count += 1
            self._tail.step(L.flat, opt.m, opt.v, L.packed, L.packed_tc, self._scratch, B, opt.step_count, loss=self._loss,
                            grad=self._grad if self.keep_grad else None, grad_norm=self._grad_norm, lr=opt.lr, max_norm=0.5,
                            beta1=opt.betas[0], beta2=opt.betas[1], eps=opt.eps)
            return
        iqn_ops.loss_grad(L.flat, L.packed, T.flat, T.packed, states, actions, rewards, next_states, dones, taus[0], taus[1],
                          gamma_n, self._scratch, self._loss, self._grad)
        world = mdist.all_reduce_sum_(self._grad)                 # no-op (returns 1) unless torch.distributed is initialised
        opt.step_count += 1
        iqn_ops.clip_adam(L.flat, self._grad, opt.m, opt.v, L.packed, step=opt.step_count, lr=opt.lr, max_norm=0.5,
                          grad_scale=1.0 / world, beta1=opt.betas[0], beta2=opt.betas[1], eps=opt.eps, grad_norm=self._grad_norm,
                          packed_tc=L.packed_tc)

    def soft_update(self, local_model, target_model):
        """agent.py:307-317: theta_target = TAU * theta_local + (1 - TAU) * theta_target (TAU = 1: hard copy)."""
        target_model.flat.copy_(self.TAU * local_model.flat + (1.0 - self.TAU) * target_model.flat)
        target_model.repack()

    # ---- the reference's single-env training loop (agent.py:94-173) -------------------------------------------------------
    def learn(self, total_timesteps, train_env, eval_env, eval_config, eval_freq, eval_log_path, verbose=True):
        state = train_env.reset()
        ep_reward, ep_length, ep_num = 0.0, 0, 0
        while self.current_timestep <= total_timesteps:
            eps = self.linear_eps(total_timesteps)
            action = self.act(state, eps)
            next_state, reward, done, info = train_env.step(action)
            ep_reward += train_env.discount ** ep_length * reward
            ep_length += 1
            self.memory.add(state, action, reward, next_state, done)
            state = next_state
            if self.current_timestep >= self.learning_starts:
                if self.learning_timestep % self.UPDATE_EVERY == 0 and len(self.memory) > self.BATCH_SIZE:
                    self.train(self.memory.sample())
                if self.learning_timestep % self.target_update_interval == 0:
                    self.soft_update(self.qnetwork_local, self.qnetwork_target)
                if self.learning_timestep % eval_freq == 0:
                    self.evaluation(eval_env, eval_config=eval_config, eval_log_path=eval_log_path)
                    self.evaluation(eval_env, eval_config=eval_config, greedy=False, eval_log_path=eval_log_path)
                    if eval_log_path is not None:
                        self.qnetwork_local.save(eval_log_path)
                self.learning_timestep += 1
            if done:
                ep_num += 1
                if verbose:
                    print("======== training info ========")
                    print("current ep_length: ", ep_length)
                    print("current ep_reward: ", ep_reward)
                    print("current ep_result: ", info["state"])
                    print("episodes_num: ", ep_num)
                    print("exploration_rate: ", eps)
                    print("current_timesteps: ", self.current_timestep)
                    print("total_timesteps: ", total_timesteps)
                    print("======== training info ========\n")
                ep_reward, ep_length = 0.0, 0
                state = train_env.reset()
            self.current_timestep += 1

    def evaluation(self, eval_env, eval_config, greedy=True, eval_log_path=None, verbose=True):
        """agent.py:319-398: one episode per evaluation map (<= 1000 steps), .npz log with the reference's schema."""
        action_data, reward_data, success_data, time_data, energy_data = [], [], [], [], []
        for idx, config in enumerate(eval_config.values()):
            if verbose:
                print(f"Evaluating episode {idx}")
            observation = eval_env.reset_with_eval_config(config)
            actions, cumulative_reward, length, energy, done = [], 0.0, 0, 0.0, False
            info = {"state": "normal"}
            while not done and length < 1000:
                if greedy:
                    action = self.act(observation, eps=0.0)
                else:
                    action, _ = self.act_adaptive(observation, eps=0.0)
                observation, reward, done, info = eval_env.step(action)
                cumulative_reward += eval_env.discount ** length * reward
                length += 1
                energy += eval_env.robot.compute_action_energy_cost(int(action))
                actions.append(int(action))
            action_data.append(actions); reward_data.append(cumulative_reward)
            success_data.append(info["state"] == "reach goal")
            time_data.append(eval_env.robot.dt * eval_env.robot.N * length); energy_data.append(energy)
        self._log_evaluation(greedy, action_data, reward_data, success_data, time_data, energy_data, eval_log_path, verbose)

    def _log_evaluation(self, greedy, action_data, reward_data, success_data, time_data, energy_data, eval_log_path, verbose=True):
        avg_r = np.mean(reward_data)
        success_rate = np.sum(success_data) / len(success_data)
        idx = np.where(np.array(success_data) == 1)[0]
        avg_t = np.mean(np.array(time_data)[idx]) if len(idx) else float("nan")
        avg_e = np.mean(np.array(energy_data)[idx]) if len(idx) else float("nan")
        policy = "greedy" if greedy else "adaptive"
        if verbose:
            print(f"++++++++ Evaluation info ({policy} IQN) ++++++++")
            print(f"Avg cumulative reward: {avg_r:.2f}")
            print(f"Success rate: {success_rate:.2f}")
            print(f"Avg time: {avg_t:.2f}")
            print(f"Avg energy: {avg_e:.2f}")
            print(f"++++++++ Evaluation info ({policy} IQN) ++++++++\n")
        self.eval_timesteps[policy].append(self.current_timestep)
        self.eval_actions[policy].append(action_data)
        self.eval_rewards[policy].append(reward_data)
        self.eval_successes[policy].append(success_data)
        self.eval_times[policy].append(time_data)
        self.eval_energies[policy].append(energy_data)
        if eval_log_path is not None:
            filename = "greedy_evaluations.npz" if greedy else "adaptive_evaluations.npz"
            np.savez(os.path.join(eval_log_path, filename),
                     timesteps=np.array(self.eval_timesteps[policy]),
                     actions=np.array(self.eval_actions[policy], dtype=object),
                     rewards=np.array(self.eval_rewards[policy]), successes=np.array(self.eval_successes[policy]),
                     times=np.array(self.eval_times[policy]), energies=np.array(self.eval_energies[policy]))

    # ---- vectorised training: everything stays on the device ----------------------------------------------------------------
    def evaluation_vec(self, eval_config, greedy=True, eval_log_path=None, verbose=False, tensor_cores=False):
        """evaluation() with all maps of eval_config stepped as ONE env batch (same episode definitions).  Acts with the
        fp32 parity kernel (iqn_forward) by default: the logged evaluations are the fp32 policy's, like the reference's;
        tensor_cores=True selects the bf16 tcgen05 kernel (q-values within ~1e-2 relative: near-ties may flip an action)."""
        from .vec_env import VecMarineNavEnv
        cfgs = list(eval_config.values())
        env = VecMarineNavEnv.from_eval_configs(cfgs, device=self.device)
        obs = env.observe_all()
        n = len(cfgs)
        ret = torch.zeros(n, dtype=torch.float64, device=self.device)
        alive = torch.ones(n, dtype=torch.bool, device=self.device)
        last_info = torch.zeros(n, dtype=torch.uint8, device=self.device)
        length = torch.zeros(n, dtype=torch.int64, device=self.device)
        acts = []
        for t in range(1000):
            a = self.act_batch(obs, 0.0, adaptive=not greedy, tensor_cores=tensor_cores)
            obs, reward, done, info = env.step(a, auto_reset=False)
            ret += torch.where(alive, (env.discount ** t) * reward.double(), torch.zeros_like(ret))
            last_info = torch.where(alive, info, last_info)
            length += alive.long()
            acts.append(torch.where(alive, a, torch.full_like(a, -1)))
            alive = alive & (done == 0)
            if t % 50 == 49 and not bool(alive.any()):
                break
        acts = torch.stack(acts).cpu().numpy()
        length_h = length.cpu().numpy()
        action_data = [[int(x) for x in acts[:length_h[i], i]] for i in range(n)]
        energy = [float(sum(env.compute_action_energy_cost(a) for a in action_data[i])) for i in range(n)]
        self._log_evaluation(greedy, action_data, list(ret.cpu().numpy()), list((last_info == 3).cpu().numpy()),
                             list(env.dt * env.N * length_h.astype(np.float64)), energy, eval_log_path, verbose)

    def capture_updates(self, batches, taus):
        """n = len(batches) consecutive updates (train_async on each (experiences, (taus_target, taus_local)) pair, in order) as
        ONE CUDA graph: returns replay(), which runs the n updates again on whatever the batch tensors hold then and advances
        the optimizer's step count by n.  Adam's bias corrections change with every step, so they come from a table of n
        mnv_vstep_ctl blocks the graph copies in first (include/marinenav_b200.h).  With data-parallel replicas this takes the
        host out of the lock-step between the ranks: the gradient exchange inside iqn_update_tail makes every update wait for
        the slowest rank, and a rank that launches from Python is late whenever its interpreter is."""
        import ctypes as C
        n, dev, opt = len(batches), self.device, self.optimizer
        B = batches[0][0].shape[0]
        need = iqn_ops.train_scratch_floats(B)
        with torch.cuda.device(dev):
            if self._scratch is None or self._scratch.numel() < need:
                self._scratch = torch.empty(need, dtype=torch.float32, device=dev)
            if self.fused_tail and self._tail is None:
                self._tail = iqn_ops.UpdateTail(dev)
            if not (self.fused_tail and self._tail.world == mdist.world_size()):
                raise _lib.MarinenavError("capture_updates needs the fused update tail (iqn_update_tail) on every rank")
            size = C.sizeof(_lib.MnvVstepCtl)
            pin = torch.zeros(n, size, dtype=torch.uint8).pin_memory()
            host = (_lib.MnvVstepCtl * n).from_buffer(pin.numpy())
            ctl = torch.zeros(n, size, dtype=torch.uint8, device=dev)
            done = torch.cuda.Event()
            cur, side = torch.cuda.current_stream(dev), torch.cuda.Stream(device=dev)
            g = torch.cuda.CUDAGraph()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    ctl.copy_(pin, non_blocking=True)
                    for u in range(n):
                        self._update(batches[u], taus[u], ctl=ctl.data_ptr() + u * size)
            cur.wait_stream(side)

        def replay():
            done.synchronize()                                   # the previous replay has read the pinned table
            for u in range(n):
                host[u].adam_step_size, host[u].adam_inv_sqrt_bc2 = iqn_ops.adam_ctl_fields(opt.lr, opt.betas[0], opt.betas[1],
                                                                                             opt.step_count + 1 + u)
            g.replay()
            done.record(torch.cuda.current_stream(dev))
            opt.step_count += n
            return self._loss
        replay.graph, replay.keep = g, (pin, ctl, batches, taus)
        return replay

    def reference_updates_per_step(self, num_envs, batch_size=None):
        """The reference's replay ratio in vector form: agent.py:127-136 trains once on BATCH_SIZE = 32 samples every
        UPDATE_EVERY = 4 transitions, i.e. 8 sampled transitions per collected one; a vector step collects num_envs (x world
        size) transitions, so the same ratio needs num_envs * 32 / (UPDATE_EVERY * batch_size) updates of batch_size."""
        B = batch_size or self.BATCH_SIZE
        return max(1, int(round(num_envs * 32.0 / (self.UPDATE_EVERY * B))))

    def learn_vec(self, total_timesteps, train_env, eval_config=None, eval_freq=None, eval_log_path=None, batch_size=None,
                  updates_per_step=1, learning_starts=None, target_update_interval=None, buffer_size=None, verbose=False,
                  on_step=None, sample_without_replacement=False, policy_lag=1, graph=False):
        """Vectorised counterpart of learn(): E transitions per env.step, device replay buffer, `updates_per_step` IQN
        updates of `batch_size` per vector step once `learning_starts` transitions were collected.

        Cadence (UPDATE_EVERY, agent.py:127-136).  The reference interleaves ONE update with every UPDATE_EVERY-th
        transition; a vector step yields E transitions at once, so the unit here is updates per vector step:
        `updates_per_step=1` (default, the throughput configuration of BASELINE configs[2]) or
        `updates_per_step="reference"` = reference_updates_per_step(E x world, batch_size), which keeps the reference's
        8 sampled transitions per collected transition.  The learning-step counters (target update every
        `target_update_interval` updates... see below) count UPDATES x UPDATE_EVERY, so that `target_update_interval` and
        `eval_freq` keep the reference's meaning ("learning timesteps" = transitions since learning started).
        Schedules (epsilon, curriculum, termination) are keyed on GLOBAL transitions: E x world size per vector step.
        Order inside a learning step as in the reference: train, then soft_update, then evaluation (agent.py:129-148).

        policy_lag=1 (default): the acting forward of vector step t uses the weights after the updates of step t-2 -- a
        snapshot (encoder weights + bf16 tensor-core tiles, two alternating buffers) the learner stream leaves behind after
        every vector step -- so that act(t) never waits for an update in flight and the learner runs entirely beside the
        env stream (act -> step -> reset).  Two vector steps of lag are far inside what the replay buffer's off-policy data
        already tolerates.  policy_lag=0: act waits for the latest weights, like the reference's loop.

        graph=True: the whole vector step is ONE CUDA graph (two alternating captures), replayed with a 64-byte device control
        block for what changes from step to step -- see _learn_vec_pipelined.  graph="eager" launches the same pipelined
        sequence without capturing it (the parity reference of the graph path)."""
        if graph:
            return self._learn_vec_pipelined(total_timesteps, train_env, eval_config, eval_freq, eval_log_path, batch_size,
                                             updates_per_step, learning_starts, target_update_interval, buffer_size, verbose, on_step,
                                             sample_without_replacement, capture=(graph != "eager"))
        E = train_env.num_envs
        world = mdist.world_size()
        if world > 1:                                           # every rank must run the same number of all-reduces
            lo_hi = torch.tensor([E, -E], dtype=torch.int64, device=self.device)
            torch.distributed.all_reduce(lo_hi, op=torch.distributed.ReduceOp.MIN)
            if int(lo_hi[0]) != E or int(-lo_hi[1]) != E:
                raise ValueError(f"learn_vec: every rank needs the same number of environments (this rank: {E}, "
                                 f"min {int(lo_hi[0])}, max {int(-lo_hi[1])}); shard num_envs * world evenly")
        train_env.global_step_multiplier = world                 # curriculum keyed on global transitions (marinenav_env.py:89-98)
        B = batch_size or self.BATCH_SIZE
        if updates_per_step == "reference":
            updates_per_step = self.reference_updates_per_step(E * world, B)
        learning_starts = self.learning_starts if learning_starts is None else learning_starts
        target_update_interval = self.target_update_interval if target_update_interval is None else target_update_interval
        if self.device_memory is None:
            self.device_memory = DeviceReplayBuffer(buffer_size or self.BUFFER_SIZE, B, self.device, seed=self.seed,
                                                    gamma=self.GAMMA, n_step=self.n_step, num_envs=E)
        mem = self.device_memory
        if mem.n_step != self.n_step:
            raise ValueError("learn_vec: the device replay buffer was built for n_step=%d, the agent uses %d" % (mem.n_step, self.n_step))
        obs = train_env.reset().clone()
        losses = []
        next_eval = 0
        steps_per_update = (E * world) / float(updates_per_step)   # transitions one update stands for (reference: UPDATE_EVERY)
        # Two streams: the env stream (act -> fused step -> masked reset + re-observe) and the learner stream (replay append
        # -> sample -> IQN update), forked right behind the step kernel.  The update does not depend on the reset and the
        # reset does not depend on the update, so the two halves of a vector step overlap; act waits for the new weights.
        env_stream = torch.cuda.current_stream(self.device)
        if getattr(self, "_learn_stream", None) is None:
            self._learn_stream = torch.cuda.Stream(device=self.device)
        learn_stream = self._learn_stream
        ev_step, ev_added, ev_update = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        net = self.qnetwork_local
        n_enc = iqn_ops.N_ENCODER_PARAMS
        snaps = [(net.flat[:n_enc].clone(), net.packed_tc.clone(), torch.cuda.Event()) for _ in range(2)] if policy_lag else None
        vstep = 0
        learn_stream.wait_stream(env_stream)
        while self.current_timestep <= total_timesteps:
            eps = self.linear_eps(total_timesteps)
            if policy_lag:
                # the learner leaves snaps[k % 2] behind at the end of vector step k; while step vstep acts, the learner may
                # still be writing snaps[(vstep - 1) % 2], so act reads the other one, snaps[vstep % 2] = the weights after
                # step vstep - 2 (complete long ago: no waiting in steady state), and the learner of THIS step overwrites
                # it only after the step kernel, i.e. after this act has finished reading
                w_flat, w_tc, w_ev = snaps[vstep % 2]
                env_stream.wait_event(w_ev)
                action = self._act_with(w_flat, w_tc, obs, eps)
            else:
                env_stream.wait_event(ev_update)                  # the weights of the last update (no-op before the first)
                action = self.act_batch(obs, eps)
            next_obs, reward, done, _ = train_env.step_begin(action)
            ev_step.record(env_stream)
            self.current_timestep += E * world
            learn_stream.wait_event(ev_step)
            action.record_stream(learn_stream)
            with torch.cuda.stream(learn_stream):
                # next_obs is the step's own (terminal) observation; the env's obs buffer gets the post-auto-reset one
                # (agent.py:122-124,170)
                mem.add_batch(obs, action, reward, next_obs, done)
                ev_added.record(learn_stream)
                do_eval = False
                if self.current_timestep >= learning_starts and len(mem) > B:
                    for _ in range(updates_per_step):
                        batch = mem.sample(B, without_replacement=sample_without_replacement)
                        taus_t = torch.rand(B, 8, device=self.device, generator=self.gen)      # Q9: target taus first
                        taus_l = torch.rand(B, 8, device=self.device, generator=self.gen)
                        loss = self.train_async(batch, (taus_t, taus_l))
                        # agent.py:135-137: the target update follows the training step of the same learning timestep;
                        # learning timesteps advance by the transitions this update stands for
                        before = self.learning_timestep
                        self.learning_timestep += steps_per_update
                        if before == 0 or int(before // target_update_interval) != int(self.learning_timestep // target_update_interval):
                            self.soft_update(self.qnetwork_local, self.qnetwork_target)
                    if verbose:
                        losses.append(loss.clone())
                    do_eval = eval_config is not None and bool(eval_freq) and self.learning_timestep >= next_eval
                if policy_lag:                                    # leave the weights of this vector step behind for act(vstep + 2)
                    s_flat, s_tc, s_ev = snaps[vstep % 2]
                    s_flat.copy_(net.flat[:n_enc]); s_tc.copy_(net.packed_tc)
                    s_ev.record(learn_stream)
                ev_update.record(learn_stream)
            vstep += 1
            train_env.step_finish(auto_reset=True)                # env stream: overlaps the update
            env_stream.wait_event(ev_added)                       # the replay append has read the previous obs
            obs.copy_(train_env.buf["obs"])
            if do_eval:
                env_stream.wait_event(ev_update)
                if mdist.rank() == 0:                             # replicas are bit-identical: one rank evaluates and logs
                    self.evaluation_vec(eval_config, greedy=True, eval_log_path=eval_log_path)
                    self.evaluation_vec(eval_config, greedy=False, eval_log_path=eval_log_path)
                    if eval_log_path is not None:
                        self.qnetwork_local.save(eval_log_path)
                mdist.barrier()
                next_eval += eval_freq
                learn_stream.wait_stream(env_stream)
            if on_step is not None:
                on_step(self)
        env_stream.wait_stream(learn_stream)
        return [float(x.item()) for x in losses[::max(1, len(losses) // 200)]] if verbose else losses

    # ---- the vector step as one CUDA graph ---------------------------------------------------------------------------------
    def _learn_vec_pipelined(self, total_timesteps, train_env, eval_config, eval_freq, eval_log_path, batch_size, updates_per_step,
                             learning_starts, target_update_interval, buffer_size, verbose, on_step, sample_without_replacement,
                             capture=True):
        """learn_vec with the vector step captured ONCE as a CUDA graph and replayed (a vector step is ~12 launches of 5 - 250 us:
        launch-bound from Python).  One replay =

            env branch:      control block H2D -> act (encode + tcgen05 kernel) -> fused env step -> replay append -> obs <- next_obs
                             -> masked reset -> masked re-observe
            learner branch:  (forked behind the control-block copy) sample -> taus -> loss/backward -> fused tail (all-reduce over
                             peer memory, clip, Adam) -> weight snapshot for the NEXT step's act; joins at the end.

        The learner of step k samples the ring as it was BEFORE step k's append (the append waits for the gather, so a full ring
        is never overwritten under it) and runs entirely beside act(k); act(k) reads the snapshot the learner of step k-1 left
        (two alternating snapshots = two alternating captures).  What changes between replays -- eps, the Philox counters, the
        ring position, Adam's bias corrections -- lives in mnv_vstep_ctl blocks in device memory (include/marinenav_b200.h); the
        host fills a pinned copy before every replay (at most two steps ahead of the GPU).  Differences to the eager loop: the
        update of a step does not see that step's own transitions, learning starts one vector step later, and with several
        updates per vector step the target network is synchronised at the end of the step in which the interval was crossed.
        capture=False launches the identical sequence eagerly."""
        import ctypes as C
        E, dev = train_env.num_envs, self.device
        world = mdist.world_size()
        if world > 1:
            lo_hi = torch.tensor([E, -E], dtype=torch.int64, device=dev)
            torch.distributed.all_reduce(lo_hi, op=torch.distributed.ReduceOp.MIN)
            if int(lo_hi[0]) != E or int(-lo_hi[1]) != E:
                raise ValueError("learn_vec: every rank needs the same number of environments")
        train_env.global_step_multiplier = world
        B = batch_size or self.BATCH_SIZE
        if updates_per_step == "reference":
            updates_per_step = self.reference_updates_per_step(E * world, B)
        U = int(updates_per_step)
        learning_starts = self.learning_starts if learning_starts is None else learning_starts
        target_update_interval = self.target_update_interval if target_update_interval is None else target_update_interval
        if self.device_memory is None:
            self.device_memory = DeviceReplayBuffer(buffer_size or self.BUFFER_SIZE, B, dev, seed=self.seed, gamma=self.GAMMA,
                                                    n_step=self.n_step, num_envs=E)
        mem = self.device_memory
        if mem.n_step != self.n_step:
            raise ValueError("learn_vec: the device replay buffer was built for n_step=%d, the agent uses %d" % (mem.n_step, self.n_step))
        steps_per_update = (E * world) / float(U)
        net, opt = self.qnetwork_local, self.optimizer
        n_enc = iqn_ops.N_ENCODER_PARAMS
        if self.fused_tail and self._tail is None:
            self._tail = iqn_ops.UpdateTail(dev)
        if not (self.fused_tail and self._tail.world == world):
            raise _lib.MarinenavError("learn_vec(graph=True) needs the fused update tail on every rank; use graph=False")
        # everything a capture bakes in by address: env buffers, both networks, the Adam moments, the replay ring
        pipe_key = (id(train_env), E, B, U, train_env.buf["state"].data_ptr(), net.flat.data_ptr(), net.packed_tc.data_ptr(),
                    self.qnetwork_target.flat.data_ptr(), opt.m.data_ptr(), mem.states.data_ptr())
        st = getattr(self, "_pipe", None)
        if st is None or st["key"] != pipe_key:
            with torch.cuda.device(dev):
                pin = torch.zeros(2, U, C.sizeof(_lib.MnvVstepCtl), dtype=torch.uint8).pin_memory()
                st = self._pipe = dict(
                    key=pipe_key, pin=pin,
                    host=[(_lib.MnvVstepCtl * U).from_buffer(pin[k].numpy()) for k in range(2)],
                    ctl=torch.zeros(2, U, C.sizeof(_lib.MnvVstepCtl), dtype=torch.uint8, device=dev),
                    taus=torch.zeros(2, U, 2, B, 8, dtype=torch.float32, device=dev),
                    action=torch.zeros(E, dtype=torch.int32, device=dev),
                    snaps=[(net.flat[:n_enc].clone(), net.packed_tc.clone()) for _ in range(2)],
                    done=[torch.cuda.Event(), torch.cuda.Event()], graphs={}, seen={}, vstep=0,
                    side=torch.cuda.Stream(device=dev), learner=torch.cuda.Stream(device=dev))
                need = iqn_ops.train_scratch_floats(B)
                if self._scratch is None or self._scratch.numel() < need:
                    self._scratch = torch.empty(need, dtype=torch.float32, device=dev)
                iqn_ops._scratch_for(E, dev)
        else:
            for k in range(2):                                    # the weights may have changed since the last call (load_model, eager updates)
                st["snaps"][k][0].copy_(net.flat[:n_enc]); st["snaps"][k][1].copy_(net.packed_tc)
        b = train_env.buf
        train_env.reset()
        losses, next_eval = [], 0
        seed_act, seed_tau = int(self.seed) + 0x5EED0000, int(self.seed) + 0x7A050000
        ev_gathered = torch.cuda.Event()

        def launch(par, update, p, rp, env_stream, learn_stream):
            """The launches of one vector step (captured, or issued eagerly): `env_stream` is the current stream."""
            ctl0 = st["ctl"][par].data_ptr()
            st["ctl"][par].copy_(st["pin"][par], non_blocking=True)
            learn_stream.wait_stream(env_stream)
            if update:
                with torch.cuda.stream(learn_stream):
                    for u in range(U):
                        cu = ctl0 + u * C.sizeof(_lib.MnvVstepCtl)
                        batch = mem.sample(B, without_replacement=sample_without_replacement, ctl=cu, advance=False)
                        taus = iqn_ops.draw_taus(st["taus"][par, u], seed_tau, ctl=cu)
                        if u == U - 1:
                            ev_gathered.record(learn_stream)
                        self._update(batch, (taus[0], taus[1]), ctl=cu)
                    nxt = st["snaps"][par ^ 1]                    # act of the NEXT step reads it (after the join)
                    nxt[0].copy_(net.flat[:n_enc]); nxt[1].copy_(net.packed_tc)
            w_flat, w_tc = st["snaps"][par]
            iqn_ops.act_tc_sample(w_flat, w_tc, b["obs"], 0.0, seed_act, 0, action=st["action"], ctl=ctl0)
            env_ops.step(b, p, action=st["action"], obs=b["next_obs"])
            if update:
                env_stream.wait_event(ev_gathered)                # a full ring: the append must not overwrite rows still being gathered
            mem.add_batch(b["obs"], st["action"], b["reward"], b["next_obs"], b["done"], ctl=ctl0, advance=False)
            b["obs"].copy_(b["next_obs"])
            env_ops.reset(b, train_env.rng_key, train_env.rng_pos, rp, mask=b["done"])
            env_ops.observe(b, p, mask=b["done"], velocity_from_state=True)
            env_stream.wait_stream(learn_stream)

        cur = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            while self.current_timestep <= total_timesteps:
                vstep = st["vstep"]
                par = vstep & 1
                update = bool(self.current_timestep >= learning_starts and mem.size > B)
                p, rp = train_env.params(), train_env.reset_params()
                st["done"][par].synchronize()                     # the replay that last read this pinned block has finished
                eps = self.linear_eps(total_timesteps)
                self._act_calls = getattr(self, "_act_calls", 0) + 1
                for u in range(U):
                    c = st["host"][par][u]
                    c.act_eps, c.act_step = eps, self._act_calls
                    mem.fill_ctl(c)
                    c.rpl_call = mem.calls + u
                    if update:
                        c.adam_step_size, c.adam_inv_sqrt_bc2 = iqn_ops.adam_ctl_fields(opt.lr, opt.betas[0], opt.betas[1], opt.step_count + 1 + u)
                key = (par, update, train_env._params_key, train_env._reset_key, bool(sample_without_replacement))
                if not capture:
                    launch(par, update, p, rp, cur, st["learner"])
                else:
                    seen = st["seen"].get(key, 0)
                    st["seen"][key] = seen + 1
                    if seen == 0:                                 # first encounter: eager (warm-up, every buffer gets allocated)
                        launch(par, update, p, rp, cur, st["learner"])
                    else:
                        g = st["graphs"].get(key)
                        if g is None:
                            g = torch.cuda.CUDAGraph()
                            side = st["side"]
                            side.wait_stream(cur)
                            with torch.cuda.stream(side):
                                with torch.cuda.graph(g, stream=side):
                                    launch(par, update, p, rp, side, st["learner"])
                            cur.wait_stream(side)
                            st["graphs"][key] = g
                        g.replay()
                st["done"][par].record(cur)
                # host-side bookkeeping of what the launches did
                st["vstep"] = vstep + 1
                train_env.total_timesteps += E * train_env.global_step_multiplier
                self.current_timestep += E * world
                mem.advance_append(E)
                do_eval = False
                if update:
                    mem.calls += U
                    opt.step_count += U
                    before = self.learning_timestep
                    self.learning_timestep += steps_per_update * U
                    if before == 0 or int(before // target_update_interval) != int(self.learning_timestep // target_update_interval):
                        self.soft_update(self.qnetwork_local, self.qnetwork_target)      # agent.py:135-137
                    if verbose:
                        losses.append(self._loss.clone())
                    do_eval = eval_config is not None and bool(eval_freq) and self.learning_timestep >= next_eval
                if do_eval:
                    if mdist.rank() == 0:
                        self.evaluation_vec(eval_config, greedy=True, eval_log_path=eval_log_path)
                        self.evaluation_vec(eval_config, greedy=False, eval_log_path=eval_log_path)
                        if eval_log_path is not None:
                            self.qnetwork_local.save(eval_log_path)
                    mdist.barrier()
                    next_eval += eval_freq
                if on_step is not None:
                    on_step(self)
        return [float(x.item()) for x in losses[::max(1, len(losses) // 200)]] if verbose else losses
