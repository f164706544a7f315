out=gpurun_out/n2b; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_n2_p2p.log 2> $out/bench_n2_p2p.err
LASV_B200_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_n2_nccl.log 2> $out/bench_n2_nccl.err
