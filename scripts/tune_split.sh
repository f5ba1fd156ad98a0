# A/B of the traversal-state split (default build) against tuning builds under tune/*.so ; results to gpurun_out/tune2
mkdir -p gpurun_out/tune2
O=gpurun_out/tune2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_default.txt
b() { # name lib workload photons
  if [ -n "$2" ]; then export PHOX_LIB=/root/repo/$2; else unset PHOX_LIB; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 3 --workload $3 --photons $4 > $O/$1_$3.json 2> $O/$1_$3.err
}
for wl in sipm8x8_scint; do
  b default "" $wl 12500000
  for v in split0 propinl r80 r48; do b $v tune/$v.so $wl 12500000; done
done
for wl in pmt_wall_torch boolean_zoo_torch scintillator_tank sphere_leak_torch; do
  b default "" $wl 4000000
  b split0 tune/split0.so $wl 4000000
  b propinl tune/propinl.so $wl 4000000
done
PHOX_LIB=/root/repo/tune/propinl.so timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_propinl.txt
unset PHOX_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune2/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j.get('roofline',{})
        print(f.split('/')[-1], '%.1f M/s'%(j['value']/1e6), 'trace %.4f ms prop %.4f ms'%(r.get('kernel_ms',0), r.get('propagate_kernel_ms',0)))
    except Exception as e: print(f,'ERR',e)
PY
