import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith("{"): continue
    d=json.loads(line)
    print("%.1f Mph/s  %.2f Gray/s  e2e %.1f  kernel %.2f ms  share %.2f bounces %.1f" % (d["value"]/1e6, d["rays_per_s"]/1e9, d["e2e"]["value"]/1e6, d["roofline"]["kernel_ms"], d["roofline"].get("kernel_share_of_bounce_loop", 0.0), d["bounces_per_photon"]))
