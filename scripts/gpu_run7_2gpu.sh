set -x
mkdir -p gpurun_out
CFD_PCG_PHASES=1 CFD_LIBPATH=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29725 bench.py --gpus 2 --steps 2 --warmup 1 --cg-variant 1 --workload cavity128_q1_implicit > gpurun_out/r02g_bench128_n2_timers_v1.json 2> gpurun_out/r02g_bench128_n2_timers_v1.err
grep -a "phases" gpurun_out/r02g_bench128_n2_timers_v1.err | tail -8
