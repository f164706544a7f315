out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "walk_limit or assembly_mode or reference_cli or clean" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -30 $out/pytest.log
