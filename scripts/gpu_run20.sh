# round 2: device step0 (tests, set-up breakdown at 256^3), sub-phase clock of the k_pcg_res sweep
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step0_on_the_device or resident" > gpurun_out/r02t_pytest.log 2>&1; tail -15 gpurun_out/r02t_pytest.log
T=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so
CFD_LIBPATH=$T CFD_PCG_PHASES=1 CFD_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 2 --workload ne27000_explicit --no-cpu-baseline > /dev/null 2> gpurun_out/r02t_phases_ne27000e.err; grep -a "pcg (\|inside the sweep\|k_pcg_res" gpurun_out/r02t_phases_ne27000e.err | tail -3
CFD_PCG_RES_T=8 CFD_LIBPATH=$T CFD_PCG_PHASES=1 CFD_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 2 --workload ne27000_explicit --no-cpu-baseline 2>&1 >/dev/null | grep -a "pcg (\|inside the sweep\|k_pcg_res" | tail -3
CFD_VERBOSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --workload ne27000_explicit --step0 device --no-cpu-baseline > gpurun_out/r02t_bench_ne27000e_dev0.json 2> gpurun_out/r02t_bench_ne27000e_dev0.err; grep -a "cfdb200" gpurun_out/r02t_bench_ne27000e_dev0.err
CFD_VERBOSE=1 BENCH_VERBOSE=1 timeout 600 python bench.py --step0 device --no-cpu-baseline > gpurun_out/r02t_bench_default_dev0.json 2> gpurun_out/r02t_bench_default_dev0.err; grep -a "cfdb200\|bench rank" gpurun_out/r02t_bench_default_dev0.err
for f in gpurun_out/r02t_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(j["config"]["workload"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"setup",j["config"]["setup_s"],"normU",j["fields"]["normU"],"normP",j["fields"]["normP"],j["fields"]["cg_iterations"])
except Exception as e: print("ERR",e)
PY
done
