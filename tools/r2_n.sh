# usage: r2_n.sh <outdir> <ngpus> <fraction> [extra bench args]
out=gpurun_out/$1; n=$2; f=$3; shift 3; mkdir -p $out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --fraction $f --steps 4 --warmup 1 --no-cpu-baseline "$@" > $out/bench_n${n}.json 2> $out/bench_n${n}.err; echo "rc=$?" >> $out/bench_n${n}.err
tail -4 $out/bench_n${n}.err
python - <<PY
import json
d=json.loads(open('$out/bench_n${n}.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value']/1e9, 'job s', d['job']['seconds'], 'fp', d['fingerprint'], 'parity', d['parity'], 'steady', d['steady_state'] and d['steady_state']['value']/1e9)
print('e2e', d['e2e'] and d['e2e']['value']/1e9, d['job']['inserts_by_kind'], d['job']['window_ms'], d['job']['steps_per_window'], d['job']['batches_accumulated_per_exchange'])
PY
