set -x
mkdir -p gpurun_out
CASES=1,2 CG_VARIANTS=1,2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py > gpurun_out/r02e_mgpu2_p2p.log 2>&1
grep -a "relU\|CfdError" gpurun_out/r02e_mgpu2_p2p.log | head -20
