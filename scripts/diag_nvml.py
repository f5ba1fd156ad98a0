"""which NVML query stalls CUDA launches? log the duration of every query next to the step times"""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np, torch, pynvml
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
w = workloads.sipm8x8_scint(num_photon=12_500_000)
g = w["geom"]
sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=ph.MODE_MINIMAL, **w["config"])
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
sim.set_stream(stream.cuda_stream)
d_gs = torch.from_numpy(w["gensteps"]).to(dev)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
log = []
stop = threading.Event()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def loop():
    while not stop.is_set():
        for name, fn in (("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                         ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)),
                         ("power", lambda: pynvml.nvmlDeviceGetPowerUsage(h))):
            if which == "sleep":
                continue
            if which != "all" and which != name:
                continue
            t0 = time.perf_counter(); fn(); log.append((name, t0, time.perf_counter() - t0))
        stop.wait(0.1)
th = threading.Thread(target=loop, daemon=True)
if which != "none":
    th.start()
steps = []
for k in range(40):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.simulate_device(d_gs.data_ptr(), len(w["gensteps"]), 0, 0, k, 0)
    torch.cuda.synchronize()
    steps.append((t0, time.perf_counter() - t0))
    if k == 5:
        time.sleep(0.5)
stop.set()
if which != "none":
    th.join()
T0 = steps[0][0]
print(which, "slow steps:", [(i, round(d * 1e3)) for i, (t, d) in enumerate(steps) if d > 0.066], "median %.1f" % (1e3 * float(np.median([d for t, d in steps]))))
slow = [(n, (t - T0) * 1e3, d * 1e3) for n, t, d in log if d > 0.005]
print("nvml calls > 5 ms:", [(n, round(a), round(b, 1)) for n, a, b in slow])
print("nvml median ms:", {n: round(1e3 * float(np.median([d for m, t, d in log if m == n])), 3) for n in ("clock", "reasons", "power") if any(m == n for m, _, _ in log)})
