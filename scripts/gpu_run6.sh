set -x
mkdir -p gpurun_out
for occ in 2 3; do
CFD_PCG_OCC=$occ timeout 300 python bench.py --steps 5 --warmup 2 --cg-variant 2 --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02f_bench128_v2_occ$occ.json 2> gpurun_out/r02f_bench128_v2_occ$occ.err
CFD_PCG_OCC=$occ timeout 300 python bench.py --steps 20 --warmup 3 --cg-variant 2 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02f_bench_ne27000e_v2_occ$occ.json 2> gpurun_out/r02f_bench_ne27000e_v2_occ$occ.err
CFD_PCG_OCC=$occ timeout 300 python bench.py --steps 20 --warmup 3 --cg-variant 1 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02f_bench_ne27000e_v1_occ$occ.json 2> gpurun_out/r02f_bench_ne27000e_v1_occ$occ.err
done
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_headline_config_fields_against_oracle > gpurun_out/r02f_pytest.log 2>&1; tail -4 gpurun_out/r02f_pytest.log
