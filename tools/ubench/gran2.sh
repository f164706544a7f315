mkdir -p gpurun_out/u
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for g in 128 64 32; do
ncu --metrics $M --clock-control none -k regex:gran_kernel --csv --log-file gpurun_out/u/gran_l2f$g.csv tools/ubench/ubench gran $g short > gpurun_out/u/gran_l2f$g.log 2>&1
done
