set -x
mkdir -p gpurun_out
for v in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --cg-variant $v --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02d_bench_ne27000e_v$v.json 2> gpurun_out/r02d_bench_ne27000e_v$v.err
timeout 600 python bench.py --steps 10 --warmup 3 --cg-variant $v --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02d_bench_ne27000i_v$v.json 2> gpurun_out/r02d_bench_ne27000i_v$v.err
timeout 600 python bench.py --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02d_bench128_n1_v$v.json 2> gpurun_out/r02d_bench128_n1_v$v.err
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "momentum or both_cg or spmv_and_pcg" > gpurun_out/r02d_pytest.log 2>&1; tail -3 gpurun_out/r02d_pytest.log
