#!/bin/bash
# after the decoder rewrite: the compressed-input GPU tests, then the gzip loader speed again (8 threads per file)
export TMPDIR=/dev/shm
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "goes_to_the_device_parse" > gpurun_out/pgz_test2.log 2>&1; tail -2 gpurun_out/pgz_test2.log
out=gpurun_out/loader_speed_gzip2.jsonl; : > $out
LASV_B200_GZ_THREADS=8 timeout 100 python tests/perf/loader_speed.py 1000000 24 --gzip >> $out 2> gpurun_out/ls_gzip2.err
cat $out
