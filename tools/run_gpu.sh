# usage: bash tools/run_gpu.sh <tag> [steps...]   (runs on the GPU box via gpurun)
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed
for step in "$@"; do
case $step in
  pytest) timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log;;
  bench) timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_default.log 2> $out/bench_default.err;;
  direct) LASV_B200_INSERT=direct timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $out/bench_direct.log 2>&1;;
  binned) LASV_B200_INSERT=binned timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $out/bench_binned.log 2>&1;;
  cfg0) timeout 600 python bench.py --workload cfg0 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_cfg0.log 2>&1;;
  launches) timeout 900 ncu --metrics $MET --clock-control none --profile-from-start off --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_launches.log 2>&1;;
  launches_direct) LASV_B200_INSERT=direct timeout 900 ncu --metrics $MET --clock-control none --profile-from-start off --csv --log-file $out/launches_direct.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_launches_direct.log 2>&1;;
  full) timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"build_kernel|insert_" -c 2 -o $out/prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_full.log 2>&1;;
  cfg1) timeout 600 python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_cfg1.log 2>&1;;
  cfg2) timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_cfg2.log 2>&1;;
  big2) timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --pairs-per-step 2097152 > $out/bench_big2.log 2>&1;;
  big4) timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --pairs-per-step 4194304 > $out/bench_big4.log 2>&1;;
  n2) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_n2.log 2> $out/bench_n2.err;;
  n4) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_n4.log 2> $out/bench_n4.err;;
  n8) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_n8.log 2> $out/bench_n8.err;;
  first_run) LASV_B200_TEST_BULK=1 timeout 900 python -m pytest tests -m gpu -q -rxX -k "branches or build_mode or health_check or grouped_insert or ranked" > $out/pytest_first_run.log 2>&1; echo "rc=$?" >> $out/pytest_first_run.log;;
  groups) (cd tools/ubench && ./ubench groups 12.5 && ./ubench groups 100) > $out/ubench_groups.txt 2>&1;;
  grouped) for bulk in 0 1; do for b in 1 2 4 8; do LASV_B200_GROUPS_BULK=$bulk timeout 600 python tools/grouped_insert_speed.py 1048576 $b 21; done; done > $out/grouped_insert_speed.jsonl 2> $out/grouped_insert_speed.err;;
  ranked_speed) for r in 0 1; do LASV_B200_SN_RANKING=$r timeout 900 python tests/perf/supernodes_speed.py 2000000 18 --clean 2; LASV_B200_SN_RANKING=$r timeout 900 python tests/perf/supernodes_speed.py 2000000 18 --no-ref; done > $out/ranked_speed.jsonl 2> $out/ranked_speed.err;;
  build_mode_2M) timeout 900 python tests/perf/build_mode_speed.py 2000000 18 2 > $out/build_mode_2Mbp.json 2> $out/build_mode_2Mbp.err;;
  build_mode_16M) timeout 1800 python tests/perf/build_mode_speed.py 16000000 21 2 > $out/build_mode_16Mbp.json 2> $out/build_mode_16Mbp.err;;
  build_mode_cfg0) timeout 1800 python tests/perf/build_mode_speed.py 63025520 22 2 --no-ref > $out/build_mode_cfg0.json 2> $out/build_mode_cfg0.err;;
  *) echo "unknown step $step";;
esac
done
ls -la $out
