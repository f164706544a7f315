out=gpurun_out/$1; mkdir -p $out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $MET --clock-control none --profile-from-start off -c 4000 --csv --log-file $out/launches_cfg4_N1.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --steady-steps 0 > $out/ncu_launches.log 2>&1; echo "ncu rc=$?"
tail -c 600 $out/ncu_launches.log
