# round 2: alternating sweep direction in k_bicgstab3 (L2 reuse of a matrix somewhat larger than L2) -- implicit tests and the shipped implicit case
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x -k "implicit or gq or momentum or tight or ne27000" > gpurun_out/r02za_pytest.log 2>&1; tail -4 gpurun_out/r02za_pytest.log
for rev in 1 0; do
  if [ $rev = 0 ]; then export CFD_NO_BICG_REVERSE=1; fi
  timeout 100 python bench.py --workload ne27000_implicit --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02za_bench_ne27000i_rev$rev.json 2> gpurun_out/r02za_bench_ne27000i_rev$rev.err
  python -c "
import json; j=json.loads(open('gpurun_out/r02za_bench_ne27000i_rev$rev.json').read().strip().splitlines()[-1]); print('rev $rev', j['value'], j['e2e']['value'], j['roofline']['frac'], j['kernels']['bicgstab'], j['fields']['normU'], j['config']['bicgstab_iterations_last_step'])"
done
