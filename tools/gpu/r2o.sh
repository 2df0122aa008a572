mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2o_pytest.log; tail -6 gpurun_out/r2o_pytest.log
timeout 300 python tools/fullsize_parity.py c4 c1 > gpurun_out/r2o_fullsize_parity.jsonl 2> gpurun_out/r2o_fullsize_parity.err
cut -c1-500 gpurun_out/r2o_fullsize_parity.jsonl
