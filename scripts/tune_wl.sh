# A/B of tuning builds over several workloads: bash scripts/tune_wl.sh "<variants>" "<workload:photons> ..."
V="$1"; WL="$2"
O=gpurun_out/tune_wl; mkdir -p $O
for v in default $V; do
  if [ $v = default ]; then unset PHOX_LIB; else export PHOX_LIB=/root/repo/tune/$v.so; fi
  for wl in $WL; do
    timeout 300 python bench.py --no-cpu-baseline --steps 3 --workload ${wl%%:*} --photons ${wl#*:} > $O/${v}_${wl%%:*}.json 2> $O/${v}_${wl%%:*}.err
  done
done
unset PHOX_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/tune_wl/*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j.get('roofline', {})
        print(f.split('/')[-1], '%.1f M/s' % (j['value'] / 1e6), r.get('kernel'), '%.4f ms' % r.get('kernel_ms', 0), r.get('second_kernel', {}).get('kernel'), '%.4f ms' % r.get('second_kernel', {}).get('kernel_ms', 0))
    except Exception as e:
        print(f, 'ERR', e)
PY
