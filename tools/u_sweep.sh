mkdir -p gpurun_out/us
MET=gpu__time_duration.sum
for u in 1 2 4; do for c in 2 3 4 6; do
LASV_B200_INSERT_U=$u LASV_B200_INSERT_CTAS=$c LASV_B200_INSERT=binned timeout 300 ncu --metrics $MET --clock-control none --profile-from-start off -k regex:insert_bins --csv --log-file gpurun_out/us/launches_u${u}_c$c.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/us/ncu_u${u}_c$c.log 2>&1
done; done
