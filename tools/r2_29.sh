out=gpurun_out/$1; mkdir -p $out
for k in 7 20 3; do
timeout 900 python bench.py --fraction 0.15 --steps $k --warmup 1 --no-e2e --no-cpu-baseline --no-parity > $out/bench_k$k.json 2> $out/bench_k$k.err; echo "rc=$?" >> $out/bench_k$k.err
tail -2 $out/bench_k$k.err
python - <<PY
import json
d=json.loads(open('$out/bench_k$k.json').read().strip().splitlines()[-1])
print('K', d['steps'], 'value', d['value']/1e9, 'job', d['job']['seconds'], d['job']['read_pairs'], d['job']['inserts_by_kind'], 'fp', d['fingerprint'], len(d['job']['window_ms']), d['job']['batches_per_window'])
PY
done
