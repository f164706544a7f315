# round 2 evidence run: bench line, ncu launch list of the same command, --set full captures of the three hot kernels
out=gpurun_out/$1; mkdir -p $out
( time timeout 1500 python bench.py --steps 20 --warmup 5 ) > $out/bench_cfg4_N1.json 2> $out/bench_cfg4_N1.err; echo "rc=$?" >> $out/bench_cfg4_N1.err
tail -4 $out/bench_cfg4_N1.err
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 2400 ncu --metrics $MET --clock-control none --profile-from-start off -c 4000 --csv --log-file $out/launches_cfg4_N1.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
tail -2 $out/ncu_launches.log | cut -c1-300
for kn in sg_insert_kernel sg_sort_kernel build_kernel; do
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"$kn" --launch-skip 20 -c 1 -o $out/prof_$kn python tools/full_job.py --fraction 0.08 --insert streamed --accumulate-gb -1 --prefill > $out/ncu_full_$kn.log 2>&1; echo "ncu $kn rc=$?"
done
ls -la $out
