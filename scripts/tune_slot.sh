mkdir -p gpurun_out/tune
for ms in 0 250000 500000 1000000 2000000 4000000; do python bench.py --no-cpu-baseline --steps 3 --max-slot $ms > gpurun_out/tune/s_$ms.json 2>gpurun_out/tune/s_$ms.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune/s_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j['roofline']
        print('%-45s %7.1f M/s e2e %7.1f ms/step %.2f loop_ms %7.2f'%(f[16:], j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], r['bounce_loop_ms']))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
