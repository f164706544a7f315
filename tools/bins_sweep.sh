mkdir -p gpurun_out/s
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
for nb in 2 16 128 1024 8192 65536; do
LASV_B200_BINS=$nb LASV_B200_INSERT=binned timeout 300 ncu --metrics $MET --clock-control none --profile-from-start off --csv --log-file gpurun_out/s/launches_bins$nb.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s/ncu_$nb.log 2>&1
done
