out=gpurun_out/p; mkdir -p $out
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/bench_pipe.log 2> $out/bench_pipe.err
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pipeline --no-e2e > $out/bench_nopipe.log 2> $out/bench_nopipe.err
LASV_B200_PIPELINE=1 timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_pipe.log 2>&1; echo "rc=$?" >> $out/pytest_pipe.log
