# bench lines of every workload and BASELINE config size on the final build (no ncu): bash scripts/gpu_final_wl.sh <outdir>
O=gpurun_out/${1:-final_wl}; mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -1 | tee $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err
for wl in raindrop_cerenkov sphere_leak_torch pmt_wall_torch boolean_zoo_torch scintillator_tank pfrich_photons box_maze_photons; do python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 5 > $O/bench_$wl.json 2>/dev/null; done
python bench.py --no-cpu-baseline --workload sphere_leak_torch --photons 1000000 > $O/cfg1_sphere_leak_1M.json 2>/dev/null
python bench.py --no-cpu-baseline --workload raindrop_cerenkov --photons 10000000 > $O/cfg2_raindrop_10M.json 2>/dev/null
python bench.py --no-cpu-baseline --workload boolean_zoo_torch --photons 1000000 > $O/cfg5_boolean_zoo_1M.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/benchline.py $O/bench_*.json $O/cfg*.json
