# round 2: configs[2] stand-in (duct, prescribed outflow) on 2 GPUs
set -x
mkdir -p gpurun_out
BENCH_WATCHDOG_S=100 timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 10 --warmup 3 --workload duct_48x32x32_implicit > gpurun_out/r02_bench_duct_48x32x32_implicit_n2.json 2> gpurun_out/r02ze_bench_duct_n2.err
python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_duct_48x32x32_implicit_n2.json').read().strip().splitlines()[-1]); print(j['value'], j['e2e']['value'], j['fields']['normU'], j['fields']['normP'], j['fields']['cg_iterations'], j['config']['bicgstab_iterations_last_step'], j['fields']['krylov_converged'], j['kernels'])"
tail -3 gpurun_out/r02ze_bench_duct_n2.err
