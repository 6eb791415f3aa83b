# round 2: suite after the swap-based rotations, then one bench line per workload (C1..C5) for profiles/
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02w_pytest.log 2>&1; tail -6 gpurun_out/r02w_pytest.log
run() { w=$1; shift; timeout 900 python bench.py --workload $w "$@" > gpurun_out/r02w_bench_$w$SFX.json 2> gpurun_out/r02w_bench_$w$SFX.err; }
SFX= run ne27000_explicit --steps 20 --warmup 3
SFX= run ne27000_implicit --steps 10 --warmup 3
SFX= run channel_48x32x32_implicit --steps 5 --warmup 3 --no-cpu-baseline
SFX= run cavity128_q1_implicit --steps 10 --warmup 3 --no-cpu-baseline
SFX= run cavity128_q1_explicit --steps 5 --warmup 3 --no-cpu-baseline
SFX= run cavity64_q2_explicit --steps 10 --warmup 3 --no-cpu-baseline
SFX=_ebe run cavity64_q2_explicit --steps 10 --warmup 3 --no-cpu-baseline --path ebe
SFX= run cavity64_q2_implicit --steps 5 --warmup 3 --no-cpu-baseline
SFX= CFD_VERBOSE=1 run cavity128_q2_explicit --steps 5 --warmup 3 --no-cpu-baseline --step0 device
SFX=_ebe run cavity128_q2_explicit --steps 5 --warmup 3 --no-cpu-baseline --path ebe --step0 device
for f in gpurun_out/r02w_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=j["roofline"]
    print(j["config"]["workload"], j["config"]["operator_path"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"setup",j["config"]["setup_s"],"roofline",r["kernel"],r["frac"],r["share_of_step"],"cpu",j.get("cpu_baseline",{}).get("value"))
except Exception as e: print("ERR",sys.argv[1],e)
PY
done
