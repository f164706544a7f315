out=gpurun_out/$1; mkdir -p $out
for dp in 1 0; do for th in 2 4 8; do
LASV_B200_DEVICE_PARSE=$dp LASV_B200_PARSE_THREADS=$th timeout 600 python tests/perf/loader_speed.py 4000000 24 2>> $out/loader.err | sed "s/}$/, \"device_parse\": $dp, \"parser_threads_per_file\": $th}/" >> $out/loader_speed.jsonl
done; done
LASV_B200_DEVICE_PARSE=1 timeout 600 python tests/perf/loader_speed.py 4000000 24 2>> $out/loader.err | sed "s/}$/, \"device_parse\": 1, \"parser_threads_per_file\": \"default\"}/" >> $out/loader_speed.jsonl
cat $out/loader_speed.jsonl | cut -c60-400; tail -3 $out/loader.err
