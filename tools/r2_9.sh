out=gpurun_out/$1; mkdir -p $out
timeout 900 python bench.py --fraction 0.1 --steps 4 --warmup 1 > $out/bench_f10.json 2> $out/bench_f10.err; echo "rc=$?" >> $out/bench_f10.err
tail -c 3000 $out/bench_f10.json; tail -5 $out/bench_f10.err
( time timeout 1500 python bench.py --steps 20 --warmup 5 ) > $out/bench_full.json 2> $out/bench_full.err; echo "rc=$?" >> $out/bench_full.err
tail -5 $out/bench_full.err
timeout 2400 python -m pytest tests -m gpu -x -q > $out/pytest_all.log 2>&1; echo "rc=$?" >> $out/pytest_all.log
tail -5 $out/pytest_all.log
