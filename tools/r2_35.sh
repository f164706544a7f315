out=gpurun_out/$1; mkdir -p $out
LASV_B200_ACC_TRACE=1 timeout 600 python -m pytest tests -m gpu -x -q -s -k "asynchronous_host" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
grep -a "\[acc\]" $out/pytest.log | head -20; tail -3 $out/pytest.log
