set -x
mkdir -p gpurun_out
BENCH_WATCHDOG_S=280 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02l_bench256_n8.json 2> gpurun_out/r02l_bench256_n8.err
CASES=0,1,2 CG_VARIANTS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29742 tests/mgpu_check.py > gpurun_out/r02l_mgpu8_p2p.log 2>&1
grep -a "relU\|CfdError" gpurun_out/r02l_mgpu8_p2p.log
