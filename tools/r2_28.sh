out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "import_keeps" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -12 $out/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; echo "rc=$?" >> $out/smoke.log; tail -3 $out/smoke.log
