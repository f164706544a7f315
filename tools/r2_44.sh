#!/bin/bash
# ordinary gzip through the file loaders: parallel inflate (PgzStream) + device parse, against zlib's single stream
export TMPDIR=/dev/shm
out=gpurun_out/loader_speed_gzip.jsonl; : > $out
nproc > gpurun_out/nproc.txt
LASV_B200_GZ_THREADS=8 timeout 100 python tests/perf/loader_speed.py 1000000 24 --gzip >> $out 2> gpurun_out/ls_gzip.err
LASV_B200_GZ_THREADS=4 timeout 60 python tests/perf/loader_speed.py 1000000 24 --gzip >> $out 2>> gpurun_out/ls_gzip.err
LASV_B200_GZ_THREADS=0 timeout 60 python tests/perf/loader_speed.py 1000000 24 --gzip >> $out 2>> gpurun_out/ls_gzip.err
cat $out
