#!/bin/bash
# own decoder for BGZF blocks + CLMUL CRC: compressed-input GPU tests, then loader speed for BGZF and gzip (8 threads per file)
export TMPDIR=/dev/shm
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "goes_to_the_device_parse" > gpurun_out/pgz_test3.log 2>&1; tail -2 gpurun_out/pgz_test3.log
out=gpurun_out/loader_speed_gz3.jsonl; : > $out
LASV_B200_GZ_THREADS=8 timeout 60 python tests/perf/loader_speed.py 1000000 24 --gzip >> $out 2> gpurun_out/ls_gz3.err
LASV_B200_GZ_THREADS=8 timeout 60 python tests/perf/loader_speed.py 1000000 24 --bgzf >> $out 2>> gpurun_out/ls_gz3.err
cat $out
