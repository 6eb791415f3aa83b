set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; tail -6 gpurun_out/r02m_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02m_smoke.log 2>&1; tail -4 gpurun_out/r02m_smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02m_bench_ne27000e.json 2> gpurun_out/r02m_bench_ne27000e.err
CFD_NO_PCG_SMEM=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02m_bench_ne27000e_stream.json 2> gpurun_out/r02m_bench_ne27000e_stream.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02m_bench_ne27000i.json 2> gpurun_out/r02m_bench_ne27000i.err
