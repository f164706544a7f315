out=gpurun_out/n2t; mkdir -p $out
LASV_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_n2_trace.log 2> $out/bench_n2_trace.err
