# round 2: whole GPU suite + smoke after the device step0 / k_pcg_res changes
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02v_pytest.log 2>&1; tail -8 gpurun_out/r02v_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02v_smoke.log 2>&1; echo smoke rc=$?; tail -4 gpurun_out/r02v_smoke.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "step0_on_the_device" 2>&1 | tail -1; done
