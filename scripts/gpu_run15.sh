set -x
mkdir -p gpurun_out
for v in 0 1; do
CFD_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --cg-variant $v --no-cpu-baseline > gpurun_out/r02o_bench_ne27000e_v$v.json 2> gpurun_out/r02o_bench_ne27000e_v$v.err
grep -a "pressure CG" gpurun_out/r02o_bench_ne27000e_v$v.err
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py -m gpu -q > gpurun_out/r02o_pytest.log 2>&1; tail -4 gpurun_out/r02o_pytest.log
