out=gpurun_out/$1; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -x -q > $out/pytest_all.log 2>&1; echo "rc=$?" >> $out/pytest_all.log
tail -5 $out/pytest_all.log
timeout 900 python bench.py --fraction 0.05 --steps 4 --warmup 1 --no-e2e --no-cpu-baseline > $out/bench_f05.json 2> $out/bench_f05.err; echo "rc=$?" >> $out/bench_f05.err
tail -3 $out/bench_f05.err
python - <<PY
import json
d=json.loads(open('$out/bench_f05.json').read().strip().splitlines()[-1])
print('value', d['value']/1e9, 'parity', d['parity'], d['parity_full_geometry'])
PY
