out=gpurun_out/r2d; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "streamed or random_inputs or hot_kmers or binned_exchange or pipelined" > $out/pytest_streamed.log 2>&1; echo "rc=$?" >> $out/pytest_streamed.log
tail -3 $out/pytest_streamed.log
timeout 600 python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 --timing > $out/job_streamed_f05.json 2>> $out/job.err
timeout 600 python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 --timing --prefill > $out/job_streamed_f05_prefill.json 2>> $out/job.err
timeout 600 python tools/full_job.py --fraction 0.05 --insert random --timing --prefill > $out/job_random_f05_prefill.json 2>> $out/job.err
cat $out/job_*.json; tail -3 $out/job.err
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
timeout 900 ncu --metrics $MET --clock-control none -k regex:"sg_|build_kernel" --launch-skip 900 -c 24 --csv --log-file $out/launches_streamed_prefill.csv python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 --prefill > $out/ncu_launches.log 2>&1
