mkdir -p gpurun_out/ls
for t in 1 2 4 8; do LASV_B200_PARSE_THREADS=$t python tests/perf/loader_speed.py 1000000 > gpurun_out/ls/loader_t$t.json 2>&1; done
python tests/perf/loader_speed.py 1000000 > gpurun_out/ls/loader_default.json 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loaders or file or cli" > gpurun_out/ls/pytest_loaders.log 2>&1
LASV_B200_PARSE_THREADS=3 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loaders or file or cli or dropin" > gpurun_out/ls/pytest_loaders_t3.log 2>&1
