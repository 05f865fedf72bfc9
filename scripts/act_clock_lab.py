"""Lab: SM clock / power while the acting kernels run back to back (nvidia-smi sampled from a thread), and the kernel's own
cycle count (clock64) against its wall time (globaltimer is not needed: CUDA events around N calls)."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributional_rl_navigation_b200 import iqn_ops
flat = torch.randn(35785, device="cuda") * 0.1
ptc = torch.empty(iqn_ops.packed_tc_bytes(), dtype=torch.uint8, device="cuda"); iqn_ops.pack_tc(flat, ptc)
E = 65536
obs = torch.randn(E, 26, device="cuda")
rows, stop = [], threading.Event()
def poll():
    while not stop.is_set():
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_event_reasons.active",
                              "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        rows.append(out); stop.wait(0.1)
th = threading.Thread(target=poll); th.start()
for phase, secs in (("idle", 0.5), ("act loop", 3.0)):
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        if phase != "idle":
            for i in range(50):
                iqn_ops.act_tc_sample(flat, ptc, obs, 0.05, 1, n + i)
            n += 50
            torch.cuda.synchronize()
        else:
            time.sleep(0.05)
    e1.record(); torch.cuda.synchronize()
    if n:
        print(f"{phase}: {e0.elapsed_time(e1) * 1e3 / n:.1f} us per call over {n} calls")
    print(phase, rows[-3:])
stop.set(); th.join()
