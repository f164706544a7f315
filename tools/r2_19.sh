out=gpurun_out/$1; mkdir -p $out
for c in 0 148 222; do
LASV_B200_SORT_CTAS=$c timeout 600 python tools/full_job.py --fraction 0.1 --insert streamed --accumulate-gb -1 --timing --prefill > $out/job_prefill_$c.json 2>> $out/job.err
python - <<PY
import json
d=json.loads(open('$out/job_prefill_$c.json').read().strip().splitlines()[-1])
k=d['kernel_ms_by_kind']
print($c, round(d['seconds'],3), d['fingerprint'], {a:(round(b[0],1),b[1]) for a,b in k.items()})
PY
done
tail -3 $out/job.err
