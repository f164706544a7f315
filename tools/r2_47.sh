#!/bin/bash
# after the cheap type sniff: gzip case of the device-parse test, then the gzip loader speed (8 threads per file)
export TMPDIR=/dev/shm
timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -k "goes_to_the_device_parse and gzip" > gpurun_out/pgz_test4.log 2>&1; tail -1 gpurun_out/pgz_test4.log
LASV_B200_GZ_THREADS=8 timeout 40 python tests/perf/loader_speed.py 1000000 24 --gzip > gpurun_out/loader_speed_gz4.jsonl 2> gpurun_out/ls_gz4.err
cat gpurun_out/loader_speed_gz4.jsonl
