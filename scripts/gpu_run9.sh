set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "matrix_free" > gpurun_out/r02i_pytest.log 2>&1; tail -15 gpurun_out/r02i_pytest.log
