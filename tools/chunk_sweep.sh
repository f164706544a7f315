mkdir -p gpurun_out/kk
for c in 24 40 64; do LASV_B200_CHUNK_MB=$c python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/kk/bench_chunk$c.log 2>&1; done
