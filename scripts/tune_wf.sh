mkdir -p gpurun_out/tune
set -x
python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wavefront or event_modes or slicing" 2>&1 | tail -15
for km in persistent wavefront; do python bench.py --no-cpu-baseline --kernel-mode $km --steps 3 > gpurun_out/tune/b_$km.json 2> gpurun_out/tune/b_$km.err; done
for n in 3 4 5; do PHOX_LIB=/root/repo/tune_wt$n.so python bench.py --no-cpu-baseline --kernel-mode wavefront --steps 3 > gpurun_out/tune/b_wt$n.json 2> gpurun_out/tune/b_wt$n.err; done
for wl in boolean_zoo_torch pmt_wall_torch scintillator_tank; do for km in persistent wavefront; do python bench.py --no-cpu-baseline --workload $wl --photons 4000000 --kernel-mode $km --steps 3 > gpurun_out/tune/b_${wl}_$km.json 2>&1; done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/tune/launches_wf.csv python bench.py --no-cpu-baseline --kernel-mode wavefront --steps 1 --warmup 1 --photons 4000000 > /dev/null 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune/b_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.1f M/s e2e %.1f kern_ms %.2f'%(j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['kernel_ms']))
    except Exception as e: print(f, 'ERR', e)
PY
