# bench every tuning build under tune/*.so (PHOX_LIB override) on the default workload; results to gpurun_out/tune
mkdir -p gpurun_out/tune
python bench.py --no-cpu-baseline --steps 3 > gpurun_out/tune/v_base.json 2> gpurun_out/tune/v_base.err
for f in tune/*.so; do n=$(basename $f .so)
  PHOX_LIB=/root/repo/$f python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wavefront_form_is_bit or intersect_bvh" 2>&1 | tail -1
  PHOX_LIB=/root/repo/$f python bench.py --no-cpu-baseline --steps 3 > gpurun_out/tune/v_$n.json 2> gpurun_out/tune/v_$n.err
  if [ -n "$EXTRA_WL" ]; then for wl in $EXTRA_WL; do PHOX_LIB=/root/repo/$f python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 3 > gpurun_out/tune/v_${n}_$wl.json 2>/dev/null; done; fi
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune/v_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.1f M/s e2e %.1f kern_ms %.2f'%(j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['kernel_ms']))
    except Exception as e: print(f, 'ERR', e)
PY
