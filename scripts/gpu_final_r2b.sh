# final GPU pass of round 2 (second session) on one B200: suite, smoke, both bench arms, A/B without home cells, launch list, ncu captures,
# per-line attribution, traffic, parity reports, other workloads.  Results under gpurun_out/r2b_final/ (copied into profiles/ by hand).
O=gpurun_out/${OUT:-r2b_final}; mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python scripts/benchline.py $O/bench_default.json
python bench.py --no-cpu-baseline --accel nohome > $O/bench_nohome.json 2> $O/bench_nohome.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 10 -c 1 -o $O/ncu_wf_trace -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_propagate -s 10 -c 1 -o $O/ncu_wf_propagate -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
python scripts/live_counts.py 4000000 1 > $O/live_counts.json 2> $O/live_counts.err
for wl in raindrop_cerenkov sphere_leak_torch pmt_wall_torch boolean_zoo_torch scintillator_tank pfrich_photons box_maze_photons; do python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 5 > $O/bench_$wl.json 2>/dev/null; done
python bench.py --no-cpu-baseline --workload sphere_leak_torch --photons 1000000 > $O/cfg1_sphere_leak_1M.json 2>/dev/null
python bench.py --no-cpu-baseline --workload raindrop_cerenkov --photons 10000000 > $O/cfg2_raindrop_10M.json 2>/dev/null
python bench.py --no-cpu-baseline --workload boolean_zoo_torch --photons 1000000 > $O/cfg5_boolean_zoo_1M.json 2>/dev/null
python tests/_parity.py --build default --out $O/parity_r2b.json > $O/parity_default.txt 2>&1
python tests/_parity.py --build nofma --out $O/parity_r2b_nofma.json > $O/parity_nofma.txt 2>&1
tail -3 $O/parity_default.txt $O/parity_nofma.txt
python scripts/form_consistency.py > $O/form_consistency.txt 2>&1
python scripts/benchline.py $O/bench_*.json $O/cfg*.json
ls $O
