# round 2, first device run of the streamed insert: focused parity tests, then the job at three settings
out=gpurun_out/r2a; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "streamed or random_inputs or hot_kmers or binned_exchange or pipelined" > $out/pytest_streamed.log 2>&1; echo "rc=$?" >> $out/pytest_streamed.log
timeout 600 python tools/full_job.py --fraction 0.1 --insert random --timing > $out/job_random_f10.json 2> $out/job_random_f10.err
timeout 600 python tools/full_job.py --fraction 0.1 --accumulate-gb -1 --insert streamed --timing > $out/job_streamed_f10.json 2> $out/job_streamed_f10.err
timeout 600 python tools/full_job.py --fraction 0.1 --accumulate-gb 5 --insert streamed --timing > $out/job_streamed5_f10.json 2> $out/job_streamed5_f10.err
tail -3 $out/pytest_streamed.log; cat $out/job_*.json; tail -2 $out/*.err
