out=gpurun_out/$1; mkdir -p $out
for sl in 32 1 32 1; do
LASV_B200_ACC_SLOTS=$sl timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --steady-steps 0 > $out/bench_s$sl.json 2> $out/bench.err
python - <<PY
import json
d=json.loads(open('$out/bench_s$sl.json').read().strip().splitlines()[-1])
print('slots $sl', 'value', round(d['value']/1e9,3), 'job', round(d['job']['seconds'],4), d['job']['inserts_by_kind'], 'e2e', round(d['e2e']['value']/1e9,2), d['e2e']['inserts_by_kind_total'], {k:(round(v['ms_total'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
done
