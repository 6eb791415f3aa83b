# round 2: k_pcg_res with cached column ids (512 threads), G^T shortcut, device step0 at 256^3
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step0_on_the_device or resident or spmv_and_pcg or library_solver or time_loop_parity or golden or matrix_free_op" > gpurun_out/r02u_pytest.log 2>&1; tail -15 gpurun_out/r02u_pytest.log
T=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so
for t in 0 4; do CFD_PCG_RES_T=$t CFD_LIBPATH=$T CFD_PCG_PHASES=1 CFD_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 2 --workload ne27000_explicit --no-cpu-baseline 2>&1 >/dev/null | grep -a "pcg (\|inside the sweep\|k_pcg_res:" | tail -3; done
CFD_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02u_bench_ne27000e.json 2> gpurun_out/r02u_bench_ne27000e.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02u_bench_ne27000i.json 2> gpurun_out/r02u_bench_ne27000i.err
CFD_VERBOSE=1 BENCH_VERBOSE=1 timeout 600 python bench.py --step0 device --no-cpu-baseline > gpurun_out/r02u_bench_default_dev0.json 2> gpurun_out/r02u_bench_default_dev0.err; grep -a "cfdb200\|bench rank" gpurun_out/r02u_bench_default_dev0.err
CFD_VERBOSE=1 BENCH_VERBOSE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02u_bench_default_host0.json 2> gpurun_out/r02u_bench_default_host0.err; grep -a "cfdb200\|bench rank" gpurun_out/r02u_bench_default_host0.err
for f in gpurun_out/r02u_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(j["config"]["workload"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"setup",j["config"]["setup_s"],"normU",j["fields"]["normU"],"normP",j["fields"]["normP"],j["fields"]["cg_iterations"][:6], {k:(v["ms"],v["launches"]) for k,v in j["kernels"].items() if k in ("pcg","bicgstab")})
except Exception as e: print("ERR",e)
PY
done
