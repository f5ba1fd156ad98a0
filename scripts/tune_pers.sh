mkdir -p gpurun_out/tune4
O=gpurun_out/tune4
for v in default ssa_inl ssa_inl_all; do
  if [ $v = default ]; then unset PHOX_LIB; else export PHOX_LIB=/root/repo/tune/$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 3 --kernel-mode persistent --photons 4000000 > $O/${v}_pers.json 2> $O/${v}_pers.err
  timeout 300 python bench.py --no-cpu-baseline --steps 3 --kernel-mode persistent --workload scintillator_tank --photons 4000000 > $O/${v}_pers_tank.json 2> $O/${v}_pers_tank.err
  timeout 300 python scripts/small_events.py > $O/${v}_small.txt 2>&1
  echo == $v; cat $O/${v}_small.txt | tail -12
done
unset PHOX_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tune4/*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], '%.1f M/s'%(j['value']/1e6))
    except Exception as e: print(f,'ERR',e)
PY
