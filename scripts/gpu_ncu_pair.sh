# ncu captures of the two bounce-loop kernels (bounce 10 of a 4 M-photon event) + live counts: bash scripts/gpu_ncu_pair.sh <outdir>
O=gpurun_out/${1:-ncu_pair}; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 10 -c 1 -o $O/ncu_wf_trace -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_propagate -s 10 -c 1 -o $O/ncu_wf_propagate -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
python scripts/live_counts.py 4000000 1 > $O/live_counts.json 2> $O/live_counts.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --no-cpu-baseline --accel nohome > $O/bench_nohome.json 2>/dev/null
ls $O
