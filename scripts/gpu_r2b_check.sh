# re-entry check of the round-2 build on one B200: suite, bench line, ncu capture (with source) of the physics kernel
O=gpurun_out/${OUT:-r2b_check}; mkdir -p $O
( time python -m pytest tests -q -m gpu -x ) 2>&1 | tail -6 | tee $O/pytest_gpu.txt
python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
python scripts/benchline.py $O/bench_default.json
ncu --set full --clock-control none --import-source on -k regex:k_wf_propagate -s 10 -c 1 -o $O/ncu_wf_propagate -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
ls -la $O
