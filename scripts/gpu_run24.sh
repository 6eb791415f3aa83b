# round 2: pipelined k_pcg_res, device-built element map, bisection fix -- suite, phases, bench lines (tight time-outs)
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r02x_pytest.log 2>&1; tail -6 gpurun_out/r02x_pytest.log
T=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so
for pipe in 1 0; do CFD_PCG_RES_PIPE=$pipe timeout 120 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02x_bench_ne27000e_pipe$pipe.json 2> gpurun_out/r02x_bench_ne27000e_pipe$pipe.err; done
CFD_PCG_RES_PIPE=0 CFD_LIBPATH=$T CFD_PCG_PHASES=1 CFD_VERBOSE=1 timeout 120 python bench.py --steps 3 --warmup 2 --workload ne27000_explicit --no-cpu-baseline 2>&1 >/dev/null | grep -a "pcg (\|inside the sweep\|k_pcg_res" | tail -3
CFD_VERBOSE=1 BENCH_VERBOSE=1 timeout 200 python bench.py --step0 device --no-cpu-baseline > gpurun_out/r02x_bench_default_dev0.json 2> gpurun_out/r02x_bench_default_dev0.err; grep -a "cfdb200\|bench rank" gpurun_out/r02x_bench_default_dev0.err
for f in gpurun_out/r02x_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"setup",j["config"]["setup_s"],"normU",j["fields"]["normU"],"normP",j["fields"]["normP"],j["fields"]["cg_iterations"][:8], {k:(v["ms"],v["launches"]) for k,v in j["kernels"].items() if k in ("pcg","bicgstab")})
except Exception as e: print("ERR",sys.argv[1],e)
PY
done
