"""rays traced per bounce of one event (live photons at each bounce): num_ray(max_bounce = b + 1) - num_ray(max_bounce = b),
and how many of them their home cell settled (the rest is what k_wf_trace walks the BVH for).
Used to turn the DRAM bytes of one ncu-captured launch into bytes per ray / per photon (profiles/traffic_r2.json)."""
import json, sys
sys.path.insert(0, ".")
import eic_opticks_b200 as ph
from eic_opticks_b200 import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
event_id = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = workloads.sipm8x8_scint(num_photon=n)
g = w["geom"]
sim = ph.Simulator.Create(g["foundry"], g["bnd"], g["optical"], g["icdf"], event_mode=ph.MODE_MINIMAL, **w["config"])
prev, out, prevh, outh = 0, [], 0, []
for b in range(1, w["config"]["max_bounce"] + 1):
    sim.set_config(max_bounce=b)
    sim.simulate_np(w["gensteps"], event_id)
    r = sim.stats()["num_ray"]
    out.append(int(r - prev)); prev = r
    h = sim.stats()["num_home_ray"]
    outh.append(int(h - prevh)); prevh = h
print(json.dumps({"photons": n, "event_id": event_id, "rays_per_bounce": out, "home_rays_per_bounce": outh}))
