out=gpurun_out/$1; shift; mkdir -p $out
for v in "$@"; do
lib=""; [ "$v" != base ] && lib=lasv_b200/variants/liblasv_b200_$v.so
LASV_B200_LIBRARY=$lib timeout 900 python bench.py --fraction 0.3 --steps 6 --warmup 2 --no-cpu-baseline --no-parity --steady-steps 0 --no-e2e > $out/bench_$v.json 2> $out/bench.err
python - <<PY
import json
d=json.loads(open('$out/bench_$v.json').read().strip().splitlines()[-1])
print('$v', 'value', round(d['value']/1e9,3), 'job', round(d['job']['seconds'],4), d['job']['inserts_by_kind'], d['fingerprint'], {k:(round(v['ms_total'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
done
