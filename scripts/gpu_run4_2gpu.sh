set -x
mkdir -p gpurun_out
CG_VARIANTS=1,2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py > gpurun_out/r02d_mgpu2_p2p.log 2>&1
grep -a "relU\|rror" gpurun_out/r02d_mgpu2_p2p.log | head -20
NO_P2P=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 tests/mgpu_check.py > gpurun_out/r02d_mgpu2_nccl.log 2>&1
grep -a "relU\|rror" gpurun_out/r02d_mgpu2_nccl.log | head -20
for v in 1 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$v bench.py --gpus 2 --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit > gpurun_out/r02d_bench128_n2_v$v.json 2> gpurun_out/r02d_bench128_n2_v$v.err
tail -c 400 gpurun_out/r02d_bench128_n2_v$v.json
done
