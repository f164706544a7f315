out=gpurun_out/$1; mkdir -p $out
for cfg in "16 1" "8 0" "16 0" "8 1"; do set -- $cfg
LASV_B200_ACC_FILL=$2 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --steady-steps 0 --pool-batches $1 > $out/bench_p$1_a$2.json 2> $out/bench.err
python - <<PY
import json
d=json.loads(open('$out/bench_p$1_a$2.json').read().strip().splitlines()[-1])
print('pool $1 adapt $2', 'value', round(d['value']/1e9,3), 'job', round(d['job']['seconds'],4), d['job']['inserts_by_kind'], {k:(round(v['ms_total'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
PY
done
