# full GPU check of round 2: parity suite, smoke, both bench arms, launch list, ncu captures of the two bounce kernels, other workloads
# SKIP_REF=1 leaves out the reference arm (the CPU oracle timing does not change with the kernels); OUT names the output directory
O=gpurun_out/${OUT:-r2}; mkdir -p $O
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
if [ -z "$SKIP_REF" ]; then python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; fi
python bench.py ${SKIP_REF:+--no-cpu-baseline} > $O/bench_default.json 2> $O/bench_default.err
tail -c 900 $O/bench_default.json
python bench.py --no-cpu-baseline --accel nohome > $O/bench_nohome.json 2> $O/bench_nohome.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 10 -c 1 -o $O/ncu_wf_trace -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_propagate -s 10 -c 1 -o $O/ncu_wf_propagate -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
python scripts/live_counts.py 4000000 1 > $O/live_counts.txt 2>&1
for wl in raindrop_cerenkov sphere_leak_torch pmt_wall_torch boolean_zoo_torch scintillator_tank; do python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 5 > $O/bench_$wl.json 2>/dev/null; done
ls -la $O
