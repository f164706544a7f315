out=gpurun_out/$1; shift; mkdir -p $out
for v in "$@"; do
lib=""; [ "$v" != base ] && lib=lasv_b200/variants/liblasv_b200_$v.so
LASV_B200_LIBRARY=$lib timeout 600 python tools/full_job.py --fraction 0.1 --insert streamed --accumulate-gb -1 --timing --prefill > $out/job_$v.json 2>> $out/job.err
python - <<PY
import json
d=json.loads(open('$out/job_$v.json').read().strip().splitlines()[-1])
k=d['kernel_ms_by_kind']
print('$v', round(d['seconds'],3), d['fingerprint'], {a:(round(b[0],1),b[1]) for a,b in k.items()})
PY
done
tail -3 $out/job.err
