set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "duct_standin" > gpurun_out/r02zd_pytest.log 2>&1; tail -5 gpurun_out/r02zd_pytest.log; grep -E "^E  " gpurun_out/r02zd_pytest.log | head -5
