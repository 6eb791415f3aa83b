set -x
mkdir -p gpurun_out
for v in 1 2; do
CFD_PCG_PHASES=1 CFD_LIBPATH=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$v bench.py --gpus 2 --steps 2 --warmup 1 --cg-variant $v --workload cavity128_q1_implicit > gpurun_out/r02g_bench128_n2_timers_v$v.json 2> gpurun_out/r02g_bench128_n2_timers_v$v.err
grep -a "phases" gpurun_out/r02g_bench128_n2_timers_v$v.err | tail -5
BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2973$v bench.py --gpus 2 --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit > gpurun_out/r02g_bench128_n2_v$v.json 2> gpurun_out/r02g_bench128_n2_v$v.err
CFD_PCG_PHASES=1 CFD_LIBPATH=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so timeout 200 python bench.py --steps 2 --warmup 1 --cg-variant $v --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02g_bench128_n1_timers_v$v.json 2> gpurun_out/r02g_bench128_n1_timers_v$v.err
grep -a "pcg.*phases" gpurun_out/r02g_bench128_n1_timers_v$v.err | tail -2
done
CG_VARIANTS=1,2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py > gpurun_out/r02g_mgpu2_p2p.log 2>&1
grep -a "relU\|CfdError" gpurun_out/r02g_mgpu2_p2p.log | head -20
