set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt 2>&1
nproc >> gpurun_out/r02a_gpus.txt
timeout 2700 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
tail -6 gpurun_out/r02a_smoke.log
timeout 900 python scripts/asis_spread.py > gpurun_out/r02a_asis_spread.txt 2>&1
timeout 900 python bench.py --workload ne27000_explicit --steps 20 --warmup 3 > gpurun_out/r02a_bench_ne27000_explicit.json 2> gpurun_out/r02a_bench_ne27000_explicit.err
timeout 900 python bench.py --workload ne27000_implicit --steps 10 --warmup 3 > gpurun_out/r02a_bench_ne27000_implicit.json 2> gpurun_out/r02a_bench_ne27000_implicit.err
timeout 1500 python bench.py > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err
tail -c 600 gpurun_out/r02a_bench_default.json
