out=gpurun_out/$1; mkdir -p $out
LASV_B200_LIBRARY=lasv_b200/variants/liblasv_b200_sifast.so timeout 900 python -m pytest tests -m gpu -x -q -k "streamed or hot_kmers or nearly_full or random_inputs" > $out/pytest_sifast.log 2>&1; echo "rc=$?" >> $out/pytest_sifast.log
tail -3 $out/pytest_sifast.log
bash tools/r2_20.sh $1 base sifast
for v in base sifast; do
lib=""; [ "$v" != base ] && lib=lasv_b200/variants/liblasv_b200_$v.so
LASV_B200_LIBRARY=$lib timeout 600 python tools/full_job.py --fraction 0.1 --insert streamed --accumulate-gb -1 --timing > $out/jobempty_$v.json 2>> $out/job.err
python - <<PY
import json
d=json.loads(open('$out/jobempty_$v.json').read().strip().splitlines()[-1])
k=d['kernel_ms_by_kind']
print('from empty', '$v', round(d['seconds'],3), d['fingerprint'], {a:(round(b[0],1),b[1]) for a,b in k.items()})
PY
done
