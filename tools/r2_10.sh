out=gpurun_out/$1; mkdir -p $out
for bins in 4096 8192 16384 32768 65536; do
LASV_B200_BINS=$bins timeout 600 python tools/full_job.py --fraction 0.1 --insert streamed --accumulate-gb -1 --timing --prefill > $out/job_bins$bins.json 2>> $out/job.err
python - <<EOF
import json
d=json.loads(open('$out/job_bins$bins.json').read().strip().splitlines()[-1])
k=d['kernel_ms_by_kind']
print($bins, round(d['seconds'],3), {a:(round(b[0],1),b[1]) for a,b in k.items()})
EOF
done
tail -3 $out/job.err
