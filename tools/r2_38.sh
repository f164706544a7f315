out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "sequence_file_loaders or file_loaders or filelists or reference_cli_with_gpu_loader" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
