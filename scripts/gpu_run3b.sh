set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "time_loop_parity or momentum or both_cg or spmv_and_pcg" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
for w in ne27000_explicit ne27000_implicit q1_128; do timeout 300 python scripts/krylov_phases.py pcg $w >> gpurun_out/r02c_pcg_phases.txt 2>&1; done
cat gpurun_out/r02c_pcg_phases.txt
for v in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --cg-variant $v --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02c_bench_ne27000e_v$v.json 2> gpurun_out/r02c_bench_ne27000e_v$v.err
timeout 600 python bench.py --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02c_bench128_n1_v$v.json 2> gpurun_out/r02c_bench128_n1_v$v.err
done
