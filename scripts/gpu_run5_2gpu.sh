set -x
mkdir -p gpurun_out
for v in 1 2; do
BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$v bench.py --gpus 2 --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit > gpurun_out/r02e_bench128_n2_v$v.json 2> gpurun_out/r02e_bench128_n2_v$v.err
done
CFD_PCG_PHASES=1 CFD_LIBPATH=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29725 bench.py --gpus 2 --steps 2 --warmup 1 --cg-variant 2 --workload cavity128_q1_implicit > gpurun_out/r02e_bench128_n2_timers.json 2> gpurun_out/r02e_bench128_n2_timers.err
grep -a "phases" gpurun_out/r02e_bench128_n2_timers.err | tail -6
CFD_PCG_PHASES=1 CFD_LIBPATH=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so timeout 200 python bench.py --steps 2 --warmup 1 --cg-variant 2 --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02e_bench128_n1_timers.json 2> gpurun_out/r02e_bench128_n1_timers.err
grep -a "phases" gpurun_out/r02e_bench128_n1_timers.err | tail -3
for v in 1 2; do
timeout 300 python bench.py --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02e_bench128_n1_v$v.json 2> gpurun_out/r02e_bench128_n1_v$v.err
timeout 300 python bench.py --steps 20 --warmup 3 --cg-variant $v --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02e_bench_ne27000e_v$v.json 2> gpurun_out/r02e_bench_ne27000e_v$v.err
done
