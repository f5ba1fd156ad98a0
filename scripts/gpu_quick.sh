# quick GPU iteration: suite + default bench line (+ optional extra command in $EXTRA)
O=gpurun_out/${OUT:-quick}; mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | grep -E "^(FAILED|E  +Assert)|passed|failed" | cut -c1-300 | tee $O/pytest_gpu.txt
python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
python scripts/benchline.py $O/bench_default.json
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
