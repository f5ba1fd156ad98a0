# bench-only A/B of tuning builds (tune/<name>.so via PHOX_LIB) against the in-tree library: bash scripts/tune_quick.sh "<variants>" [steps]
V="$1"; S="${2:-4}"
O=gpurun_out/tune_quick; mkdir -p $O
for v in default $V; do
  if [ $v = default ]; then unset PHOX_LIB; else export PHOX_LIB=/root/repo/tune/$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --steps $S > $O/$v.json 2> $O/$v.err
done
unset PHOX_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/tune_quick/*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j.get('roofline', {})
        print(f.split('/')[-1], '%.1f M/s' % (j['value'] / 1e6), 'min-step %.1f M/s' % (12.5e3 / min(j['step_ms'])), 'prop %.4f ms trace %.4f ms' % (r.get('kernel_ms', 0), r.get('second_kernel', {}).get('kernel_ms', 0)))
    except Exception as e:
        print(f, 'ERR', e)
PY
