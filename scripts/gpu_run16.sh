# round 2, re-entry: whole GPU suite + smoke + bench lines at the current code state
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02p_pytest.log 2>&1; tail -6 gpurun_out/r02p_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02p_smoke.log 2>&1; echo smoke rc=$?; tail -4 gpurun_out/r02p_smoke.log
CFD_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02p_bench_ne27000e.json 2> gpurun_out/r02p_bench_ne27000e.err
CFD_NO_PCG_SMEM=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02p_bench_ne27000e_stream.json 2> gpurun_out/r02p_bench_ne27000e_stream.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02p_bench_ne27000i.json 2> gpurun_out/r02p_bench_ne27000i.err
BENCH_VERBOSE=1 timeout 600 python bench.py > gpurun_out/r02p_bench_default.json 2> gpurun_out/r02p_bench_default.err
for f in gpurun_out/r02p_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(j["config"]["workload"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"setup",j["config"]["setup_s"],"roofline",j["roofline"]["kernel"],j["roofline"]["frac"],{k:(v["ms"],v["launches"],v["gbs"]) for k,v in j["kernels"].items()})
except Exception as e: print("ERR",e)
PY
done
