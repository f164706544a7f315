out=gpurun_out/$1; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -k "bgzf or sequence_file_loaders or parse_fastq_text" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
