out=gpurun_out/$1; mkdir -p $out
( time timeout 1500 python bench.py --steps 20 --warmup 5 ) > $out/bench_full.json 2> $out/bench_full.err; echo "rc=$?" >> $out/bench_full.err
tail -6 $out/bench_full.err
python - <<PY
import json
d=json.loads(open('$out/bench_full.json').read().strip().splitlines()[-1])
print('value', d['value']/1e9, 'job', d['job']['seconds'], d['job']['inserts_by_kind'], 'fp', d['fingerprint'], 'parity', d['parity'], 'steady', d['steady_state']['value']/1e9, 'e2e', d['e2e']['value']/1e9)
print({k:(round(v['ms_total'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
print(d['roofline']['frac'], d['roofline']['step_frac'], d['parity_full_geometry'])
PY
timeout 2400 python -m pytest tests -m gpu -x -q > $out/pytest_all.log 2>&1; echo "rc=$?" >> $out/pytest_all.log
tail -5 $out/pytest_all.log
