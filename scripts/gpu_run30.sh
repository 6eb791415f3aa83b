# round 2: ncu --set full of the resident pressure-CG kernel (one whole solve on the shipped mesh); only the text summary comes back
set -x
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -f -k regex:'k_pcg_res' --launch-skip 6 --launch-count 1 -o /tmp/pcg_res python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload ne27000_explicit > gpurun_out/r02zb_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/pcg_res.ncu-rep > gpurun_out/r02_ne27000_pcg_res_ncu.txt 2>&1
ncu -i /tmp/pcg_res.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for k in ['smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_op_shared_ld.sum','launch__shared_mem_per_block_dynamic']:
    if k in h: print('%-90s %-12s %s' % (k, u[h.index(k)], v[h.index(k)]))
" >> gpurun_out/r02_ne27000_pcg_res_ncu.txt
cat gpurun_out/r02_ne27000_pcg_res_ncu.txt
