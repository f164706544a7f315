out=gpurun_out/$1; mkdir -p $out
for wl in cfg2 cfg0; do
timeout 600 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $out/bench_$wl.json 2> $out/bench_$wl.err; echo "rc=$?" >> $out/bench_$wl.err
python - <<PY
import json
d=json.loads(open('$out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', 'value', round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('partition_kernel_ms'))
PY
done
timeout 1200 python -m pytest tests -m gpu -x -q -k "streamed or asynchronous or random_inputs or hot_kmers or device_resident or pipelined or host_buffer" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
