mkdir -p gpurun_out/tune
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
EXTRA_WL="raindrop_cerenkov pmt_wall_torch boolean_zoo_torch scintillator_tank" bash scripts/tune_variants.sh
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune/v_*_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.1f M/s'%(j['value']/1e6))
    except Exception as e: print(f,'ERR',e)
PY
for wl in raindrop_cerenkov pmt_wall_torch boolean_zoo_torch scintillator_tank; do python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --steps 3 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base', '$wl', '%.1f M/s'%(j['value']/1e6))"; done
