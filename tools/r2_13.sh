out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "streamed or random_inputs or hot_kmers or binned_exchange or pipelined" > $out/pytest_streamed.log 2>&1; echo "rc=$?" >> $out/pytest_streamed.log
tail -3 $out/pytest_streamed.log
for cfg in "65536 0" "65536 2" "65536 3" "65536 4" "16384 2"; do set -- $cfg
LASV_B200_BINS=$1 LASV_B200_RUN_SHIFT=$2 timeout 600 python tools/full_job.py --fraction 0.1 --insert streamed --accumulate-gb -1 --timing --prefill > $out/job_bins$1_rs$2.json 2>> $out/job.err
python - <<EOF
import json
d=json.loads(open('$out/job_bins$1_rs$2.json').read().strip().splitlines()[-1])
k=d['kernel_ms_by_kind']
print($1, $2, round(d['seconds'],3), d['fingerprint'], {a:(round(b[0],1),b[1]) for a,b in k.items()})
EOF
done
tail -3 $out/job.err
