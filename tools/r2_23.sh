out=gpurun_out/$1; mkdir -p $out
( time timeout 1500 python tests/perf/supernodes_speed.py 63025520 22 ) > $out/supernodes_cfg0_ref.json 2> $out/supernodes_cfg0_ref.err; echo "rc=$?" >> $out/supernodes_cfg0_ref.err
tail -c 1500 $out/supernodes_cfg0_ref.json; tail -5 $out/supernodes_cfg0_ref.err
