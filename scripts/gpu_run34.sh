# round 2: the whole GPU suite at the final commit
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; tail -4 gpurun_out/r02_final_pytest.log
