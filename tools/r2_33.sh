out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "streamed or asynchronous or human_scale or random_inputs or hot_kmers or file_loaders" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > $out/bench_full.json 2> $out/bench_full.err; echo "rc=$?" >> $out/bench_full.err
tail -3 $out/bench_full.err
python - <<PY
import json
d=json.loads(open('$out/bench_full.json').read().strip().splitlines()[-1])
print('value', d['value']/1e9, 'job', d['job']['seconds'], d['job']['inserts_by_kind'], 'fp', d['fingerprint'], 'parity', d['parity'], 'steady', d['steady_state']['value']/1e9, 'e2e', d['e2e']['value']/1e9)
print({k:(round(v['ms_total'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
