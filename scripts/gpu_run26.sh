# round 2: configs[3] Q2/Q1 at 128^3, assembled vs matrix-free operators (dt below the explicit viscous limit)
set -x
mkdir -p gpurun_out
timeout 170 python bench.py --workload cavity128_q2_explicit --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02z_bench_cavity128_q2_explicit.json 2> gpurun_out/r02z_bench_cavity128_q2_explicit.err
timeout 170 python bench.py --workload cavity128_q2_explicit --steps 3 --warmup 1 --no-cpu-baseline --path ebe > gpurun_out/r02z_bench_cavity128_q2_explicit_ebe.json 2> gpurun_out/r02z_bench_cavity128_q2_explicit_ebe.err
for f in gpurun_out/r02z_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=j["roofline"]
    print(sys.argv[1], j["config"]["operator_path"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),"setup",j["config"]["setup_s"],"roofline",r["kernel"],r["frac"],r["share_of_step"],"normU",j["fields"]["normU"],"maxAcc",j["fields"]["maxAcc"],"cg",j["fields"]["cg_iterations"][:5])
    print("    ", {kk:(vv["ms"],vv["launches"],vv["gbs"]) for kk,vv in j["kernels"].items()})
except Exception as e: print("ERR",sys.argv[1],e)
PY
done
