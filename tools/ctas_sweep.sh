mkdir -p gpurun_out/c2
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
for c in 1 2 3 4; do
LASV_B200_INSERT_CTAS=$c LASV_B200_INSERT=binned timeout 300 ncu --metrics $MET --clock-control none --profile-from-start off --csv --log-file gpurun_out/c2/launches_ctas$c.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c2/ncu_$c.log 2>&1
done
