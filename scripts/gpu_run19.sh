# round 2: k_pcg_res -- parity tests, phase clock, bench lines on the shipped mesh
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident or spmv_and_pcg or library_solver or time_loop_parity or both_cg or ne27000 or golden" > gpurun_out/r02s_pytest.log 2>&1; tail -15 gpurun_out/r02s_pytest.log
T=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so
CFD_LIBPATH=$T CFD_PCG_PHASES=1 CFD_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 2 --workload ne27000_explicit --no-cpu-baseline > /dev/null 2> gpurun_out/r02s_phases_ne27000e.err; grep -a "pcg (\|k_pcg_res" gpurun_out/r02s_phases_ne27000e.err | tail -4
for t in 4 16; do CFD_PCG_RES_T=$t CFD_LIBPATH=$T CFD_PCG_PHASES=1 timeout 300 python bench.py --steps 3 --warmup 2 --workload ne27000_explicit --no-cpu-baseline 2>&1 >/dev/null | grep -a "pcg (" | tail -1; done
CFD_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02s_bench_ne27000e.json 2> gpurun_out/r02s_bench_ne27000e.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02s_bench_ne27000i.json 2> gpurun_out/r02s_bench_ne27000i.err
for f in gpurun_out/r02s_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(j["config"]["workload"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"cg/step",j["config"]["cg_iterations_per_step"],{k:(v["ms"],v["launches"]) for k,v in j["kernels"].items()})
except Exception as e: print("ERR",e)
PY
done
