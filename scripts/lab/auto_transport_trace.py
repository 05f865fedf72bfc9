"""lab: per-call wall time of step_host while host_transport="auto" measures the transports, then in steady state."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E = 65536
acts = np.random.RandomState(0).randint(0, 9, size=(64, E)).astype(np.int32)
env = VecMarineNavEnv(E, seed=0, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0)
env.reset_host()
env._hybrid()["trace"] = []
rows = []
for i in range(110):
    c = env._auto_cal
    tr = c["order"][c["calls"] // c["per"]] if c and c["calls"] >= 0 else env.host_transport
    t0 = time.perf_counter(); env.step_host(acts[i % 64]); rows.append((tr, (time.perf_counter() - t0) * 1e6))
print(" ".join(f"{t[0]}{us:.0f}" for t, us in rows))
print("calibration:", env.host_transport_calibration, "->", env.host_transport)
tr_ = env._hybrid()["trace"]
print("hybrid calls while measuring (spins, hits, us: replay->packet seen, expand, stream sync, rescan_skipped):")
for r in tr_[:11]:
    print("   ", r[0], r[1], " ".join(f"{x * 1e6:.0f}" for x in r[2:]))
del tr_[:]
for tr in ("compact", "hybrid", "dense"):
    env.host_transport = tr
    for i in range(5):
        env.step_host(acts[i])
    ts = []
    for i in range(50):
        t0 = time.perf_counter(); env.step_host(acts[i % 64]); ts.append((time.perf_counter() - t0) * 1e6)
    print(f"steady {tr}: median {sorted(ts)[25]:.0f} us, min {min(ts):.0f}, max {max(ts):.0f}")
print("hybrid calls in steady state:")
for r in env._hybrid()["trace"][-6:]:
    print("   ", r[0], r[1], " ".join(f"{x * 1e6:.0f}" for x in r[2:]))
