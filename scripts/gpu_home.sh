# home-cell check: identity test, bench A (home cells) / B (plain BVH), ncu of k_wf_home
O=gpurun_out/${1:-r2b}; mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "home_cells" -s 2>&1 | grep -v "^$" | tail -12 | tee $O/pytest_home.txt
for a in ${ACCELS:-bvh}; do
  python bench.py --no-cpu-baseline --accel $a --steps 3 > $O/bench_$a.json 2> $O/bench_$a.err
  python - <<PY
import json; j=json.loads(open("$O/bench_$a.json").read().strip().splitlines()[-1]); r=j["roofline"]; print("$a", j["value"]/1e6, "M/s trace", r["kernel_ms"], "prop", r["propagate_kernel_ms"], "home", j.get("home_ray_fraction"))
PY
done
if [ -n "$NCU" ]; then ncu --set full --clock-control none --import-source on -k regex:$NCU -s 10 -c 1 -o $O/ncu_$NCU -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1; fi
ls $O
