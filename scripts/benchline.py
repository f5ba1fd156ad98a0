"""one-line summary of bench.py JSON lines: python scripts/benchline.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r = j["roofline"]; s = r.get("second_kernel") or {}
        print("%s: %.1f M/s, e2e %.1f M/s, %s %.4f ms (frac %.3f), %s %.4f ms, home %.3f, %.1f bounces/photon" % (
            f.split("/")[-1], j["value"] / 1e6, j["e2e"]["value"] / 1e6, r["kernel"], r["kernel_ms"], r["frac"], s.get("kernel"), s.get("kernel_ms", 0.0),
            j.get("home_ray_fraction", 0.0), j["bounces_per_photon"]))
    except Exception as e:
        print(f, "ERR", e)
