# round 2: the duct stand-in with prescribed outflow (configs[2]) -- parity test and bench line on one GPU
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "duct_standin or channel_standin" > gpurun_out/r02zc_pytest.log 2>&1; tail -5 gpurun_out/r02zc_pytest.log
timeout 200 python bench.py --workload duct_48x32x32_implicit --steps 10 --warmup 3 > gpurun_out/r02_bench_duct_48x32x32_implicit.json 2> gpurun_out/r02zc_bench_duct.err
python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_duct_48x32x32_implicit.json').read().strip().splitlines()[-1]); r=j['roofline']; print(j['value'], j['e2e']['value'], r['kernel'], r['frac'], r['share_of_step'], j['fields']['normU'], j['fields']['normP'], j['fields']['cg_iterations'], j['config']['bicgstab_iterations_last_step'], j['fields']['krylov_converged'], j.get('cpu_baseline',{}).get('value'), j['config']['setup_s'])"
