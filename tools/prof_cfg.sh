# usage: bash tools/prof_cfg.sh <tag> <workload>
out=gpurun_out/$1; mkdir -p $out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
timeout 600 ncu --metrics $MET --clock-control none --profile-from-start off --csv --log-file $out/launches_$2.csv python bench.py --workload $2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_$2.log 2>&1
