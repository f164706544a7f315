out=gpurun_out/r2b; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "streamed or random_inputs or hot_kmers or binned_exchange or pipelined" > $out/pytest_streamed.log 2>&1; echo "rc=$?" >> $out/pytest_streamed.log
tail -3 $out/pytest_streamed.log
(cd tools/ubench && timeout 300 ./ubench groups 12.5 && timeout 300 ./ubench groups 100) > $out/ubench_groups.txt 2>&1
cat $out/ubench_groups.txt
for bins in 2048 16384; do LASV_B200_BINS=$bins timeout 600 python tools/full_job.py --fraction 0.05 --insert random --timing > $out/job_random_bins$bins.json 2>> $out/job.err; done
LASV_B200_BINS=16384 timeout 600 python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 --timing > $out/job_streamed_bins16384.json 2>> $out/job.err
cat $out/job_*.json
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
timeout 900 ncu --metrics $MET --clock-control none -k regex:"sg_|build_kernel" --launch-skip 14 -c 40 --csv --log-file $out/launches_streamed.csv python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 > $out/ncu_launches.log 2>&1
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"sg_" --launch-skip 12 -c 4 -o $out/prof_sg python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 > $out/ncu_full.log 2>&1
tail -5 $out/ncu_full.log
ls -la $out
