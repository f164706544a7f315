out=gpurun_out/san; mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "device_resident_path or binned_exchange_matches or pipelined_inserts or host_buffer_path" > $out/memcheck.log 2>&1; echo "memcheck rc=$?" >> $out/memcheck.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_clocks.log 2>&1
