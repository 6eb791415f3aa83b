set -x
mkdir -p gpurun_out
CFD_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02n_bench_ne27000e.json 2> gpurun_out/r02n_bench_ne27000e.err
grep -a "pressure CG\|Error\|error" gpurun_out/r02n_bench_ne27000e.err
CFD_VERBOSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02n_bench_ne27000i.json 2> gpurun_out/r02n_bench_ne27000i.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; tail -4 gpurun_out/r02n_pytest.log
