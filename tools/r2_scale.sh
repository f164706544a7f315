# usage: r2_scale.sh <outdir> <ngpus>   -- the driver's command line at N GPUs, full job
out=gpurun_out/$1; n=$2; mkdir -p $out
if [ $n = 1 ]; then launcher="python"; else launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517"; fi
( time timeout 1500 $launcher bench.py --gpus $n --steps 20 --warmup 5 ) > $out/bench_cfg4_N$n.json 2> $out/bench_cfg4_N$n.err; echo "rc=$?" >> $out/bench_cfg4_N$n.err
grep -v "^\*\|OMP_NUM" $out/bench_cfg4_N$n.err | tail -6
python - <<PY
import json
d=json.loads(open('$out/bench_cfg4_N$n.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value']/1e9, 'job s', d['job']['seconds'], 'fp', d['fingerprint'], 'parity', d['parity'], 'steady', d['steady_state'] and d['steady_state']['value']/1e9, 'e2e', d['e2e'] and d['e2e']['value']/1e9)
print(d['job']['inserts_by_kind'], d['job']['window_ms'], d['job']['steps_per_window'], d['job']['batches_accumulated_per_exchange'], d['clocks'])
PY
