out=gpurun_out/$1; mkdir -p $out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,lts__t_sector_hit_rate.pct,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum
for bins in 4096 65536; do
LASV_B200_BINS=$bins timeout 900 ncu --metrics $MET --clock-control none -k regex:"build_kernel" --launch-skip 20 -c 3 --csv --log-file $out/launches_build_$bins.csv python tools/full_job.py --fraction 0.03 --insert streamed --accumulate-gb -1 --prefill > $out/ncu_$bins.log 2>&1
done
LASV_B200_BINS=65536 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"build_kernel" --launch-skip 20 -c 1 -o $out/prof_build_65536 python tools/full_job.py --fraction 0.03 --insert streamed --accumulate-gb -1 --prefill > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
