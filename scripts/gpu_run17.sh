# round 2: phase clock of the resident PCG, matrix-free path on the shipped mesh, ncu --set full of the non-Krylov kernels
set -x
mkdir -p gpurun_out
T=$PWD/cfd-with-cuda_b200/libcfdb200_timers.so
CFD_LIBPATH=$T CFD_PCG_PHASES=1 timeout 300 python bench.py --steps 3 --warmup 2 --workload ne27000_explicit --no-cpu-baseline > /dev/null 2> gpurun_out/r02q_phases_ne27000e.err; grep -a "pcg (" gpurun_out/r02q_phases_ne27000e.err | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --path ebe --no-cpu-baseline > gpurun_out/r02q_bench_ne27000e_ebe.json 2> gpurun_out/r02q_bench_ne27000e_ebe.err
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:'k_ebe|k_step1|k_step3|k_gt_rhs|k_spmv_multi' --launch-skip 35 --launch-count 7 -o gpurun_out/r02q_explicit_kernels python bench.py --steps 2 --warmup 1 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02q_ncu1.log 2>&1
timeout 600 $NCU -k regex:'k_pcg_smem' --launch-skip 6 --launch-count 1 -o gpurun_out/r02q_pcg_smem python bench.py --steps 2 --warmup 1 --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02q_ncu2.log 2>&1
timeout 600 $NCU -k regex:'k_ebe|k_gq_rhs' --launch-skip 4 --launch-count 3 -o gpurun_out/r02q_implicit_kernels python bench.py --steps 2 --warmup 1 --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02q_ncu3.log 2>&1
timeout 600 $NCU -k regex:'k_ebe' --launch-skip 30 --launch-count 5 -o gpurun_out/r02q_matrixfree_kernels python bench.py --steps 2 --warmup 1 --workload ne27000_explicit --path ebe --no-cpu-baseline > gpurun_out/r02q_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
python - gpurun_out/r02q_bench_ne27000e_ebe.json <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j["config"]["workload"], "value",round(j["value"],3),"e2e",round(j["e2e"]["value"],3),{k:(v["ms"],v["launches"],v["gbs"]) for k,v in j["kernels"].items()})
PY
