"""The `_ctl` entry points (scalars from the mnv_vstep_ctl device block, include/marinenav_b200.h) against their by-value
counterparts, and learn_vec(graph=True) -- the rollout + learn vector step (agent.py:94-173 vectorised) replayed as ONE CUDA
graph -- against the same pipelined launch sequence issued eagerly: bit-identical weights, replay ring and env state."""
import ctypes as C

import numpy as np
import pytest
import torch

from test_replay import philox4x32_10

DEV = "cuda:0"


def _ctl_block(**fields):
    from distributional_rl_navigation_b200 import _lib
    c = _lib.MnvVstepCtl()
    for k, v in fields.items():
        setattr(c, k, v)
    t = torch.frombuffer(bytearray(bytes(c)), dtype=torch.uint8).to(DEV)
    return t


def test_ctl_struct_layout():
    """64 bytes, the field offsets the header documents (the kernels read it through the C struct)."""
    from distributional_rl_navigation_b200 import _lib
    S = _lib.MnvVstepCtl
    assert C.sizeof(S) == 64
    assert [getattr(S, f).offset for f in ("act_eps", "adam_step_size", "adam_inv_sqrt_bc2", "act_step", "rpl_pos", "rpl_t", "rpl_head",
                                           "rpl_size", "rpl_call")] == [0, 4, 8, 16, 24, 32, 40, 48, 56]


def test_adam_ctl_fields_match_by_value_arithmetic():
    from distributional_rl_navigation_b200 import iqn_ops
    for step in (1, 2, 10, 1000, 123456):
        ss, isb = iqn_ops.adam_ctl_fields(1e-4, 0.9, 0.999, step)
        lr, b1, b2 = float(np.float32(1e-4)), float(np.float32(0.9)), float(np.float32(0.999))     # float arguments of the C-ABI
        assert ss == lr / (1.0 - b1 ** step) and isb == 1.0 / np.sqrt(1.0 - b2 ** step)


@pytest.mark.gpu
def test_draw_taus_follows_the_philox_model():
    from distributional_rl_navigation_b200 import iqn_ops
    seed, call, B = 0x1234ABCD5, 77, 37
    out = torch.zeros(2, B, 8, device=DEV)
    iqn_ops.draw_taus(out, seed, call)
    got = out.cpu().numpy().reshape(-1)
    key = seed ^ 0x7A75
    want = []
    for q in range((got.size + 3) // 4):
        r = philox4x32_10((q, 0x7A, call & 0xffffffff, call >> 32), (key & 0xffffffff, key >> 32))
        want += [np.float32(x >> 8) * np.float32(1.0 / 16777216.0) for x in r]
    assert np.array_equal(got, np.asarray(want[:got.size], np.float32))
    assert got.min() >= 0.0 and got.max() < 1.0
    out2 = torch.zeros(2, B, 8, device=DEV)
    blk = _ctl_block(rpl_call=call)
    iqn_ops.draw_taus(out2, seed, 0, ctl=blk.data_ptr())                            # the call counter from the control block
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.gpu
def test_replay_ctl_variants_match_by_value():
    from distributional_rl_navigation_b200.replay_buffer import DeviceReplayBuffer
    g = torch.Generator(device=DEV); g.manual_seed(5)
    E, cap, B = 96, 400, 64
    for n_step in (1, 3):
        a = DeviceReplayBuffer(cap, B, DEV, seed=11, n_step=n_step, num_envs=E)
        b = DeviceReplayBuffer(cap, B, DEV, seed=11, n_step=n_step, num_envs=E)
        for t in range(9):                                   # wraps the ring (9 x 96 > 400)
            s, n = torch.randn(E, 26, device=DEV, generator=g), torch.randn(E, 26, device=DEV, generator=g)
            act = torch.randint(0, 9, (E,), device=DEV, generator=g, dtype=torch.int32)
            r = torch.randn(E, device=DEV, generator=g); d = (torch.rand(E, device=DEV, generator=g) < 0.1).to(torch.uint8)
            a.add_batch(s, act, r, n, d)
            blk = _ctl_block(rpl_pos=b.pos, rpl_t=b.t)
            b.add_batch(s, act, r, n, d, ctl=blk.data_ptr(), advance=False)
            b.advance_append(E)
            torch.cuda.synchronize()
            assert (a.pos, a.size, a.t) == (b.pos, b.size, b.t)
            if a.size >= B:
                for wo in (False, True):
                    ba = [x.clone() for x in a.sample(B, without_replacement=wo)]
                    blk = _ctl_block(rpl_head=b.head, rpl_size=b.size, rpl_call=b.calls)
                    bb = b.sample(B, without_replacement=wo, ctl=blk.data_ptr(), advance=False)
                    b.calls += 1
                    torch.cuda.synchronize()
                    assert torch.equal(a.last_indices, b.last_indices)
                    for x, y in zip(ba, bb):
                        assert torch.equal(x, y)
        for f in ("states", "next_states", "actions", "rewards", "dones"):
            assert torch.equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.gpu
def test_act_and_update_ctl_variants_match_by_value():
    from distributional_rl_navigation_b200 import iqn_ops
    from distributional_rl_navigation_b200.iqn_agent import IQNAgent
    g = torch.Generator(device=DEV); g.manual_seed(3)
    # act: eps / step from the block
    ag = IQNAgent(26, 9, seed=4, device=DEV, BATCH_SIZE=64)
    net = ag.qnetwork_local
    obs = torch.randn(1000, 26, device=DEV, generator=g) * 3
    a1, _, q1 = iqn_ops.act_tc_sample(net.flat, net.packed_tc, obs, 0.3, 99, 12345678901, want_qmean=True)
    blk = _ctl_block(act_eps=0.3, act_step=12345678901)
    a2, _, q2 = iqn_ops.act_tc_sample(net.flat, net.packed_tc, obs, 0.0, 99, 0, want_qmean=True, ctl=blk.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(a1, a2) and torch.equal(q1, q2)
    greedy, _, _ = iqn_ops.act_tc_sample(net.flat, net.packed_tc, obs, 0.0, 99, 12345678901)
    assert 0.15 < float((a1 != greedy).float().mean()) < 0.35          # eps really came from the block (8/9 of 30 % explore off-greedy)
    # update: Adam's bias corrections from the block, three consecutive steps
    B = 64
    agents = [IQNAgent(26, 9, seed=4, device=DEV, BATCH_SIZE=B) for _ in range(2)]
    for step in range(1, 4):
        batch = (torch.randn(B, 26, device=DEV, generator=g), torch.randint(0, 9, (B,), device=DEV, generator=g),
                 torch.randn(B, device=DEV, generator=g), torch.randn(B, 26, device=DEV, generator=g),
                 (torch.rand(B, device=DEV, generator=g) < 0.1).float())
        taus = (torch.rand(B, 8, device=DEV, generator=g), torch.rand(B, 8, device=DEV, generator=g))
        agents[0]._update(batch, taus)
        ss, isb = iqn_ops.adam_ctl_fields(1e-4, 0.9, 0.999, step)
        blk = _ctl_block(adam_step_size=ss, adam_inv_sqrt_bc2=isb)
        if agents[1]._tail is None:
            agents[1]._tail = iqn_ops.UpdateTail(torch.device(DEV))
        agents[1]._update(batch, taus, ctl=blk.data_ptr())
        agents[1].optimizer.step_count += 1
        torch.cuda.synchronize()
        for x, y in ((agents[0].qnetwork_local.flat, agents[1].qnetwork_local.flat), (agents[0].optimizer.m, agents[1].optimizer.m),
                     (agents[0].optimizer.v, agents[1].optimizer.v), (agents[0]._loss, agents[1]._loss),
                     (agents[0].qnetwork_local.packed_tc, agents[1].qnetwork_local.packed_tc)):
            assert torch.equal(x, y)


def _run_pipeline(graph, E=2048, B=256, steps=14, updates_per_step=1, n_step=1):
    from distributional_rl_navigation_b200.iqn_agent import IQNAgent
    from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv
    env = VecMarineNavEnv(E, seed=21, device=DEV)
    agent = IQNAgent(26, 9, seed=5, device=DEV, BATCH_SIZE=B, BUFFER_SIZE=5 * E, n_step=n_step)    # the ring wraps inside the run
    trace = []
    agent.learn_vec(total_timesteps=E * (steps - 1), train_env=env, batch_size=B, learning_starts=2 * E, target_update_interval=4 * E,
                    updates_per_step=updates_per_step, sample_without_replacement=True, graph=graph,
                    on_step=lambda a: trace.append(a._pipe["action"].clone()))
    # a second call continues on the cached graphs (what bench.py does: warm-up call, then the timed call)
    agent.learn_vec(total_timesteps=agent.current_timestep + E * 3, train_env=env, batch_size=B, learning_starts=2 * E,
                    target_update_interval=4 * E, updates_per_step=updates_per_step, sample_without_replacement=True, graph=graph)
    torch.cuda.synchronize()
    return agent, env, trace


@pytest.mark.gpu
@pytest.mark.parametrize("updates_per_step,n_step", [(1, 1), (3, 2)])
def test_learn_vec_graph_matches_the_eager_pipeline(updates_per_step, n_step):
    ag_g, env_g, tr_g = _run_pipeline(True, updates_per_step=updates_per_step, n_step=n_step)
    ag_e, env_e, tr_e = _run_pipeline("eager", updates_per_step=updates_per_step, n_step=n_step)
    assert len(tr_g) == len(tr_e) == 14
    for k, (x, y) in enumerate(zip(tr_g, tr_e)):
        assert torch.equal(x, y), f"actions of vector step {k} differ"
    assert len(ag_g._pipe["graphs"]) >= 2                                        # the graph path really replayed captures
    assert ag_g.optimizer.step_count == ag_e.optimizer.step_count > 0
    assert ag_g.current_timestep == ag_e.current_timestep and ag_g.learning_timestep == ag_e.learning_timestep
    assert torch.equal(ag_g.qnetwork_local.flat, ag_e.qnetwork_local.flat)
    assert torch.equal(ag_g.qnetwork_target.flat, ag_e.qnetwork_target.flat)
    fresh = type(ag_g)(26, 9, seed=5, device=DEV).qnetwork_local.flat
    assert not torch.equal(ag_g.qnetwork_local.flat, fresh) and not torch.equal(ag_g.qnetwork_target.flat, fresh)   # trained, target synchronised
    assert torch.isfinite(ag_g.qnetwork_local.flat).all()
    mg, me = ag_g.device_memory, ag_e.device_memory
    assert (mg.pos, mg.size, mg.t, mg.calls) == (me.pos, me.size, me.t, me.calls)
    for f in ("states", "next_states", "actions", "rewards", "dones"):
        assert torch.equal(getattr(mg, f), getattr(me, f)), f
    for k in ("state", "obs", "goal", "obstacles", "episode_step"):
        assert torch.equal(env_g.buf[k], env_e.buf[k]), k
    assert env_g.total_timesteps == env_e.total_timesteps


@pytest.mark.gpu
def test_learn_vec_graph_follows_the_schedules():
    """eps and the update gate change between replays of the SAME graph: exploration falls, updates start late."""
    from distributional_rl_navigation_b200.iqn_agent import IQNAgent
    from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv
    E = 1024
    env = VecMarineNavEnv(E, seed=2, device=DEV)
    agent = IQNAgent(26, 9, seed=6, device=DEV, BATCH_SIZE=128, BUFFER_SIZE=8 * E, exploration_fraction=0.5)
    eps_seen, upd = [], []
    def on_step(a):
        eps_seen.append(float(a._pipe["host"][(a._pipe["vstep"] - 1) & 1][0].act_eps)); upd.append(a.optimizer.step_count)
    agent.learn_vec(total_timesteps=E * 19, train_env=env, batch_size=128, learning_starts=6 * E, target_update_interval=4 * E, graph=True,
                    on_step=on_step)
    assert eps_seen[0] == 1.0 and eps_seen[-1] == pytest.approx(0.05) and all(a >= b for a, b in zip(eps_seen, eps_seen[1:]))
    assert upd[5] == 0 and upd[-1] == 20 - 6                                       # no update before learning_starts, one per step after


@pytest.mark.gpu
def test_capture_updates_matches_eager_updates():
    """IQNAgent.capture_updates: n updates as one CUDA graph (Adam's bias corrections from the control-block table), replayed
    twice == the same 2 n updates launched one by one."""
    from distributional_rl_navigation_b200.iqn_agent import IQNAgent
    g = torch.Generator(device=DEV); g.manual_seed(8)
    B, n = 128, 5
    batches = [(torch.randn(B, 26, device=DEV, generator=g), torch.randint(0, 9, (B,), device=DEV, generator=g),
                torch.randn(B, device=DEV, generator=g), torch.randn(B, 26, device=DEV, generator=g),
                (torch.rand(B, device=DEV, generator=g) < 0.1).float()) for _ in range(n)]
    taus = [(torch.rand(B, 8, device=DEV, generator=g), torch.rand(B, 8, device=DEV, generator=g)) for _ in range(n)]
    a, b = IQNAgent(26, 9, seed=4, device=DEV, BATCH_SIZE=B), IQNAgent(26, 9, seed=4, device=DEV, BATCH_SIZE=B)
    replay = a.capture_updates(batches, taus)
    replay(); replay()
    for _ in range(2):
        for u in range(n):
            b.train_async(batches[u], taus[u])
    torch.cuda.synchronize()
    assert a.optimizer.step_count == b.optimizer.step_count == 2 * n
    assert torch.equal(a.qnetwork_local.flat, b.qnetwork_local.flat) and torch.equal(a.optimizer.v, b.optimizer.v)
    assert torch.equal(a.qnetwork_local.packed_tc, b.qnetwork_local.packed_tc) and torch.equal(a._loss, b._loss)
