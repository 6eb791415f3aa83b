# round 2, final state: suite, smoke, default bench line (+ reference arm), ncu launch list, 128^3 Q2/Q1 pair of lines (tight time-outs)
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02y_pytest.log 2>&1; tail -6 gpurun_out/r02y_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02y_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r02y_smoke.log
CFD_VERBOSE=1 timeout 300 python bench.py > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y_bench_default.err
timeout 200 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit > gpurun_out/r02y_bench_ne27000_explicit.json 2> gpurun_out/r02y_bench_ne27000_explicit.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_cavity256_q1_implicit.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02y_ncu_launches.log 2>&1
CFD_VERBOSE=1 timeout 170 python bench.py --workload cavity128_q2_explicit --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02y_bench_cavity128_q2_explicit.json 2> gpurun_out/r02y_bench_cavity128_q2_explicit.err; grep -a "cfdb200" gpurun_out/r02y_bench_cavity128_q2_explicit.err | tail -12
timeout 170 python bench.py --workload cavity128_q2_explicit --steps 2 --warmup 1 --no-cpu-baseline --path ebe > gpurun_out/r02y_bench_cavity128_q2_explicit_ebe.json 2> gpurun_out/r02y_bench_cavity128_q2_explicit_ebe.err
for f in gpurun_out/r02y_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=j["roofline"]
    print(sys.argv[1], j["config"]["operator_path"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"setup",j["config"]["setup_s"],"roofline",r["kernel"],r["frac"],r["share_of_step"],"normU",j["fields"]["normU"],"cg",j["fields"]["cg_iterations"][:5],"cpu",j.get("cpu_baseline",{}).get("value"))
    print("    ", {kk:(vv["ms"],vv["launches"],vv["gbs"]) for kk,vv in j["kernels"].items()})
except Exception as e: print("ERR",sys.argv[1],e)
PY
done
