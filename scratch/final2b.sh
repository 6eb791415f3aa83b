python -m pytest tests -m gpu -q > gpurun_out/r01_gpu_suite.log 2>&1; tail -2 gpurun_out/r01_gpu_suite.log
python -m pytest tests -m gpu -q 2>&1 | tail -1
python __graft_entry__.py smoke > gpurun_out/r01_smoke.log 2>&1; tail -2 gpurun_out/r01_smoke.log | cut -c1-200
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench256.err | tail -1 > gpurun_out/r01_bench_cavity256_q1_implicit.json; python scratch/pj.py < gpurun_out/r01_bench_cavity256_q1_implicit.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/r01_bench_reference_arm.json; cut -c1-300 gpurun_out/r01_bench_reference_arm.json
