# bench every tuning build under tune/*.so (PHOX_LIB override) on the default workload; results to gpurun_out/tune
mkdir -p gpurun_out/tune
python bench.py --no-cpu-baseline --steps 3 > gpurun_out/tune/v_base.json 2> gpurun_out/tune/v_base.err
for f in tune/*.so; do n=$(basename $f .so)
  PHOX_LIB=/root/repo/$f timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wavefront_form or intersect_bvh or edge_cases or lite" 2>&1 | tail -1
  PHOX_LIB=/root/repo/$f timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/tune/v_$n.json 2> gpurun_out/tune/v_$n.err
  if [ -n "$EXTRA_WL" ]; then for wl in $EXTRA_WL; do PHOX_LIB=/root/repo/$f timeout 600 python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 3 > gpurun_out/tune/v_${n}_$wl.json 2>/dev/null; done; fi
done
if [ -n "$EXTRA_WL" ]; then for wl in $EXTRA_WL; do timeout 600 python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 3 > gpurun_out/tune/v_base_$wl.json 2>/dev/null; done; fi
