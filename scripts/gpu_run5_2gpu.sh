set -x
mkdir -p gpurun_out
BENCH_VERBOSE=1 BENCH_WATCHDOG_S=150 CFD_VERBOSE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 3 --warmup 1 --cg-variant 2 --workload cavity128_q1_implicit > gpurun_out/r02e_bench128_n2_v2.json 2> gpurun_out/r02e_bench128_n2_v2.err
grep -a "bench rank\|Error\|File\|cfdb200" gpurun_out/r02e_bench128_n2_v2.err | tail -40
