out=gpurun_out/r2e; mkdir -p $out
for sb in 262144 1048576 4194304; do
LASV_B200_SLAB_BUCKETS=$sb timeout 600 python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 --timing --prefill > $out/job_streamed_f05_prefill_sb$sb.json 2>> $out/job.err
done
timeout 600 python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 --timing > $out/job_streamed_f05.json 2>> $out/job.err
cat $out/job_*.json; tail -3 $out/job.err
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
timeout 900 ncu --metrics $MET --clock-control none -k regex:"sg_" --launch-skip 40 -c 12 --csv --log-file $out/launches_streamed.csv python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 > $out/ncu_launches.log 2>&1
LASV_B200_SLAB_BUCKETS=1048576 timeout 900 ncu --metrics $MET --clock-control none -k regex:"sg_" --launch-skip 12 -c 12 --csv --log-file $out/launches_streamed_sb1M.csv python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 > $out/ncu_launches2.log 2>&1
