out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "streamed or random_inputs or hot_kmers or binned_exchange or pipelined or asynchronous or nearly_full or table_full" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
bash tools/r2_17.sh $1
