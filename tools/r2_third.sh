out=gpurun_out/r2c; mkdir -p $out
for bins in 65536; do LASV_B200_BINS=$bins timeout 600 python tools/full_job.py --fraction 0.05 --insert random --timing > $out/job_random_bins$bins.json 2>> $out/job.err; done
cat $out/job_*.json
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"sg_insert" --launch-skip 3 -c 1 -o $out/prof_sg_insert python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 > $out/ncu_full.log 2>&1
tail -5 $out/ncu_full.log
ls -la $out
