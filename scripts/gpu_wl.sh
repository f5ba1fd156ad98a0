# bench lines of the default workload and the short-history workloads: bash scripts/gpu_wl.sh <outdir>
O=gpurun_out/${1:-wl}; mkdir -p $O
python bench.py --no-cpu-baseline > $O/bench_default.json 2>/dev/null
for wl in raindrop_cerenkov pmt_wall_torch scintillator_tank boolean_zoo_torch; do python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 5 > $O/bench_$wl.json 2>/dev/null; done
python bench.py --no-cpu-baseline --workload raindrop_cerenkov --photons 10000000 > $O/cfg2_raindrop_10M.json 2>/dev/null
python scripts/benchline.py $O/*.json
python - <<PY
import json
d=json.loads(open("$O/bench_default.json").read().strip().splitlines()[-1]); print("step_ms", d["step_ms"], "launches", d["gpu_launches"])
PY
