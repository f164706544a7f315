out=gpurun_out/r2f; mkdir -p $out
LASV_B200_SLAB_BUCKETS=1048576 timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"sg_insert" --launch-skip 36 -c 1 -o $out/prof_sg_insert_steady python tools/full_job.py --fraction 0.05 --insert streamed --accumulate-gb -1 --prefill > $out/ncu_full.log 2>&1
tail -3 $out/ncu_full.log
