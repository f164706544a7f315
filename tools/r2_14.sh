out=gpurun_out/$1; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "asynchronous_host or host_buffer_path or streamed_insert_with or many_small_batches" > $out/pytest_async.log 2>&1; echo "rc=$?" >> $out/pytest_async.log
tail -5 $out/pytest_async.log
timeout 900 python bench.py --fraction 0.1 --steps 4 --warmup 1 --e2e-steps 20 --no-cpu-baseline > $out/bench_f10.json 2> $out/bench_f10.err; echo "rc=$?" >> $out/bench_f10.err
tail -5 $out/bench_f10.err
python - <<PY
import json
d=json.loads(open('$out/bench_f10.json').read().strip().splitlines()[-1])
print('value', d['value']/1e9, 'e2e', d['e2e'], 'parity', d['parity'])
PY
