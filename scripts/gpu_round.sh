# full GPU check of the round: parity suite, both bench arms, launch list, ncu captures of the two bounce kernels
# SKIP_REF=1 leaves out the reference arm (the CPU oracle timing does not change with the kernels)
mkdir -p gpurun_out/r1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ -z "$SKIP_REF" ]; then python bench.py --impl reference > gpurun_out/r1/bench_reference.json 2> gpurun_out/r1/bench_reference.err; fi
python bench.py > gpurun_out/r1/bench_default.json 2> gpurun_out/r1/bench_default.err
tail -c 600 gpurun_out/r1/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 40 -c 1 -o gpurun_out/r1/ncu_wf_trace -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_propagate -s 40 -c 1 -o gpurun_out/r1/ncu_wf_propagate -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
python scripts/live_counts.py > gpurun_out/r1/live_counts.txt 2>&1
for wl in raindrop_cerenkov sphere_leak_torch pmt_wall_torch boolean_zoo_torch scintillator_tank; do python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 5 > gpurun_out/r1/bench_$wl.json 2>/dev/null; done
ls -la gpurun_out/r1
