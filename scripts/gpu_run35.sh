# round 2: the restart tests on the GPU after the host-side restart change (last seconds of the budget)
mkdir -p gpurun_out
timeout 40 python -m pytest tests -m gpu -q -k "restart" > gpurun_out/r02_restart_pytest.log 2>&1; tail -3 gpurun_out/r02_restart_pytest.log
