# round 2, final state on a 2-GPU box: the two-GPU tests, partitioned vs single-GPU check (all four cases), configs[2] stand-in and the shipped mesh on 2 GPUs
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpu" > gpurun_out/r02z_pytest_2gpu.log 2>&1; tail -4 gpurun_out/r02z_pytest_2gpu.log
CG_VARIANTS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tests/mgpu_check.py > gpurun_out/r02z_mgpu2.log 2>&1; grep -a "relU\|Error\|error" gpurun_out/r02z_mgpu2.log | tail -8
BENCH_WATCHDOG_S=100 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 5 --warmup 3 --workload cavity64_q2_implicit > gpurun_out/r02z_bench_cavity64_q2_implicit_n2.json 2> gpurun_out/r02z_bench_cavity64_q2_implicit_n2.err
BENCH_WATCHDOG_S=100 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29743 bench.py --gpus 2 --steps 10 --warmup 3 --workload cavity64_q2_explicit > gpurun_out/r02z_bench_cavity64_q2_explicit_n2.json 2> gpurun_out/r02z_bench_cavity64_q2_explicit_n2.err
for f in gpurun_out/r02z_bench_*_n2.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"normU",j["fields"]["normU"],"normP",j["fields"]["normP"],"cg",j["fields"]["cg_iterations"][:5])
except Exception as e: print("ERR",sys.argv[1],e)
PY
done
