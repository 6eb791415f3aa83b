# round 2: ncu --set full of the non-Krylov kernels and the resident PCG (reports summarised on the box; only text comes back)
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none -f"
cap() {  # name, regex, skip, count, bench args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU -k regex:"$rx" --launch-skip $skip --launch-count $cnt -o /tmp/$name python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/r02r_$name.log 2>&1
  python scripts/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/r02r_${name}_ncu.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02r_${name}_raw.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep
}
cap explicit_kernels 'k_ebe|k_step1|k_step3|k_gt_rhs|k_spmv_multi' 35 7 --workload ne27000_explicit
cap pcg_smem 'k_pcg_smem' 6 1 --workload ne27000_explicit
cap implicit_kernels 'k_ebe|k_gq_rhs' 4 3 --workload ne27000_implicit
cap matrixfree_kernels 'k_ebe' 30 5 --workload ne27000_explicit --path ebe
CFD_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -a "cfdb200" 
du -sh gpurun_out
