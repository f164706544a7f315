out=gpurun_out/$1; mkdir -p $out
for a in "20000 6000 5 --branching" "30000 12000 6 --branching" "8000 4000 7 --branching"; do timeout 900 python tests/perf/assembly_mode_check.py $a >> $out/assembly.jsonl 2>> $out/assembly.err; done
cat $out/assembly.jsonl; tail -5 $out/assembly.err
