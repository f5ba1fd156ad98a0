# A/B of tuning builds (scripts/build_variants.sh -> tune/<name>.so, selected through PHOX_LIB) against the in-tree library.
#   bash scripts/tune_ab.sh "<variants>" ["<workload:photons> ..."] ["<variants that also run the parity suite>"]
# e.g. (the second pass of profiles/r1_summary.md):
#   bash scripts/build_variants.sh split0 "-DPHOX_TRAV_SPLIT=0" inl "-DPHOX_PROP_SSA=0"
#   gpurun -- 'bash scripts/tune_ab.sh "split0 inl" "sipm8x8_scint:12500000 scintillator_tank:4000000" "inl"'
# The full GPU suite runs first on the in-tree build; results land in gpurun_out/tune_ab/.
V="$1"; WL="${2:-sipm8x8_scint:12500000}"; PV="$3"
O=gpurun_out/tune_ab; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_default.txt
for v in $PV; do
  PHOX_LIB=/root/repo/tune/$v.so timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu 2>&1 | tail -8 > $O/pytest_$v.txt; echo "$v: $(tail -1 $O/pytest_$v.txt)"
done
for v in default $V; do
  if [ $v = default ]; then unset PHOX_LIB; else export PHOX_LIB=/root/repo/tune/$v.so; fi
  for wl in $WL; do
    timeout 300 python bench.py --no-cpu-baseline --steps 3 --workload ${wl%%:*} --photons ${wl#*:} > $O/${v}_${wl%%:*}.json 2> $O/${v}_${wl%%:*}.err
  done
done
unset PHOX_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/tune_ab/*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j.get('roofline', {})
        print(f.split('/')[-1], '%.1f M/s' % (j['value'] / 1e6), 'trace %.4f ms prop %.4f ms' % (r.get('kernel_ms', 0), r.get('propagate_kernel_ms', 0)))
    except Exception as e:
        print(f, 'ERR', e)
PY
