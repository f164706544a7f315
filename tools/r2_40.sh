out=gpurun_out/$1; mkdir -p $out
for cfg in "1 4" "1 8" "0 8" "1 0"; do set -- $cfg
LASV_B200_DEVICE_PARSE=$1 LASV_B200_GZ_THREADS=$2 timeout 900 python tests/perf/loader_speed.py 1000000 23 --bgzf 2>> $out/loader.err | sed "s/}$/, \"device_parse\": $1, \"gz_threads\": $2}/" >> $out/loader_speed_bgzf.jsonl
done
cut -c60-500 $out/loader_speed_bgzf.jsonl; tail -3 $out/loader.err
