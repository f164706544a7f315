out=gpurun_out/$1; mkdir -p $out
for wl in cfg0 cfg1 cfg2; do
timeout 600 python bench.py --workload $wl --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_$wl.json 2> $out/bench_$wl.err; echo "rc=$?" >> $out/bench_$wl.err
tail -2 $out/bench_$wl.err
python - <<PY
import json
d=json.loads(open('$out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', 'value', round(d['value']/1e9,2), 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), d.get('parity'), d['roofline']['kernel'], round(d['roofline']['frac'],3))
PY
done
timeout 2400 python -m pytest tests -m gpu -x -q > $out/pytest_all.log 2>&1; echo "rc=$?" >> $out/pytest_all.log
tail -4 $out/pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "rc=$?" >> $out/smoke.log; tail -2 $out/smoke.log
