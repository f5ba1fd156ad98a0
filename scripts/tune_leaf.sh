mkdir -p gpurun_out/tune8
O=gpurun_out/tune8
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_default.txt
PHOX_LIB=/root/repo/tune/leafinl.so timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu 2>&1 | tail -8 > $O/pytest_leafinl.txt; tail -1 $O/pytest_leafinl.txt
for v in default prev leafinl; do
  if [ $v = default ]; then unset PHOX_LIB; else export PHOX_LIB=/root/repo/tune/$v.so; fi
  for wl in sipm8x8_scint:12500000 scintillator_tank:4000000 pmt_wall_torch:4000000 sphere_leak_torch:4000000 boolean_zoo_torch:4000000; do
    timeout 300 python bench.py --no-cpu-baseline --steps 3 --workload ${wl%%:*} --photons ${wl#*:} > $O/${v}_${wl%%:*}.json 2> $O/${v}_${wl%%:*}.err
  done
done
unset PHOX_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune8/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j.get('roofline',{})
        print(f.split('/')[-1], '%.1f M/s'%(j['value']/1e6), 'trace %.4f ms prop %.4f ms'%(r.get('kernel_ms',0), r.get('propagate_kernel_ms',0)))
    except Exception as e: print(f,'ERR',e)
PY
