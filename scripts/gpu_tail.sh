# tail threshold sweep (PHOX_TAIL_PHOTONS) on the workloads with long tails
for t in 0 65536 200000 500000; do
  for wl in scintillator_tank boolean_zoo_torch pfrich_photons pmt_wall_torch; do
    PHOX_TAIL_PHOTONS=$t python bench.py --no-cpu-baseline --no-clocks --workload $wl --photons 4000000 --steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tail $t $wl %.1f M/s' % (d['value']/1e6))"
  done
done
