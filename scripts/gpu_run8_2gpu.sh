set -x
mkdir -p gpurun_out
for v in 1 2; do
BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2973$v bench.py --gpus 2 --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit > gpurun_out/r02h_bench128_n2_v$v.json 2> gpurun_out/r02h_bench128_n2_v$v.err
done
CFD_PCG_PHASES=1 CFD_LIBPATH=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 2 --warmup 1 --cg-variant 2 --workload cavity128_q1_implicit > gpurun_out/r02h_bench128_n2_timers_v2.json 2> gpurun_out/r02h_bench128_n2_timers_v2.err
grep -a "pcg.*phases" gpurun_out/r02h_bench128_n2_timers_v2.err | tail -2
for occ in 2 3; do
CFD_PCG_OCC=$occ timeout 300 python bench.py --steps 5 --warmup 2 --cg-variant 2 --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02h_bench128_v2_occ$occ.json 2> gpurun_out/r02h_bench128_v2_occ$occ.err
CFD_PCG_OCC=$occ timeout 300 python bench.py --steps 20 --warmup 3 --cg-variant 2 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02h_bench_ne27000e_v2_occ$occ.json 2> gpurun_out/r02h_bench_ne27000e_v2_occ$occ.err
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest.log 2>&1; tail -6 gpurun_out/r02h_pytest.log
