mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2k_pytest.log; tail -8 gpurun_out/r2k_pytest.log
bash tools/gpu/prof_r2.sh
