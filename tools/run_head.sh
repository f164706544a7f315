set -x
mkdir -p gpurun_out/h
python -m pytest tests -m gpu -x -q > gpurun_out/h/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h/pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/h/bench_default.log 2> gpurun_out/h/bench_default.err
LASV_B200_INSERT=direct python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/h/bench_direct.log 2>&1
LASV_B200_INSERT=binned python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/h/bench_binned.log 2>&1
LASV_B200_INSERT=direct ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:build_kernel -s 8 -c 4 --csv --log-file gpurun_out/h/launches_direct.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/h/ncu_direct.log 2>&1
LASV_B200_INSERT=binned ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:"build_kernel|insert_bins" -s 8 -c 4 --csv --log-file gpurun_out/h/launches_binned.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/h/ncu_binned.log 2>&1
LASV_B200_INSERT=binned ncu --set full --import-source on --clock-control none -k regex:"build_kernel|insert_bins" -s 8 -c 2 -o gpurun_out/h/prof_binned python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/h/ncu_binned_full.log 2>&1
LASV_B200_INSERT=direct ncu --set full --import-source on --clock-control none -k regex:build_kernel -s 8 -c 1 -o gpurun_out/h/prof_direct python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/h/ncu_direct_full.log 2>&1
ls -la gpurun_out/h
