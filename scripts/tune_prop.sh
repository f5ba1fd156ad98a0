# physics-body variants (tune/*.so): parity suite + bench each ; results to gpurun_out/tune3
mkdir -p gpurun_out/tune3
O=gpurun_out/tune3
for v in ssa ssa_inl ssa_inl_all inl_all; do
  export PHOX_LIB=/root/repo/tune/$v.so
  timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu 2>&1 | tail -8 > $O/pytest_$v.txt
  tail -1 $O/pytest_$v.txt
  for wl in sipm8x8_scint:12500000 scintillator_tank:4000000 pmt_wall_torch:4000000; do
    timeout 300 python bench.py --no-cpu-baseline --steps 3 --workload ${wl%%:*} --photons ${wl#*:} > $O/${v}_${wl%%:*}.json 2> $O/${v}_${wl%%:*}.err
  done
done
unset PHOX_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune3/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j.get('roofline',{})
        print(f.split('/')[-1], '%.1f M/s'%(j['value']/1e6), 'trace %.4f ms prop %.4f ms'%(r.get('kernel_ms',0), r.get('propagate_kernel_ms',0)))
    except Exception as e: print(f,'ERR',e)
PY
