# round 2: last sanity pass after the host-only changes (OpenMP setting scoped to set-up, resident note in the bench line)
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x -k "golden or time_loop_parity or resident or step0_on_the_device or solverGPU or refhost or host_program" > gpurun_out/r02zz_pytest.log 2>&1; tail -4 gpurun_out/r02zz_pytest.log
timeout 100 python bench.py --workload ne27000_explicit --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02zz_bench_ne27000e.json 2> gpurun_out/r02zz_bench_ne27000e.err; python -c "
import json; j=json.loads(open('gpurun_out/r02zz_bench_ne27000e.json').read().strip().splitlines()[-1]); print(j['value'], j['roofline'].get('resident'), j['config']['l2'])"
