out=gpurun_out/$1; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "binned_exchange or asynchronous_host or pipelined" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 900 python bench.py --fraction 0.1 --steps 4 --warmup 1 --e2e-steps 12 --no-cpu-baseline > $out/bench_f10.json 2> $out/bench_f10.err; echo "rc=$?" >> $out/bench_f10.err
tail -5 $out/bench_f10.err
python - <<PY
import json
d=json.loads(open('$out/bench_f10.json').read().strip().splitlines()[-1])
print('value', d['value']/1e9, 'job', d['job']['seconds'], d['job']['inserts_by_kind'], d['job']['window_ms'], 'e2e', d['e2e']['value']/1e9, 'parity', d['parity'], d['parity_detail'])
print({k:(round(v['ms_total'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
