out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "streamed or random_inputs or hot_kmers or binned_exchange or pipelined or asynchronous or nearly_full or table_full" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 900 python bench.py --fraction 0.1 --steps 4 --warmup 1 --no-e2e --no-cpu-baseline > $out/bench_f10.json 2> $out/bench_f10.err; echo "rc=$?" >> $out/bench_f10.err
tail -3 $out/bench_f10.err
python - <<PY
import json
d=json.loads(open('$out/bench_f10.json').read().strip().splitlines()[-1])
print('value', d['value']/1e9, 'job', d['job']['seconds'], d['job']['inserts_by_kind'], 'fp', d['fingerprint'], 'parity', d['parity'], 'steady', d['steady_state']['value']/1e9)
print({k:(round(v['ms_total'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
