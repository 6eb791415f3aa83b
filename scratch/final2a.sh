set -x
python -m pytest tests -m gpu -q > gpurun_out/r01_gpu_suite.log 2>&1; tail -3 gpurun_out/r01_gpu_suite.log
python __graft_entry__.py smoke > gpurun_out/r01_smoke.log 2>&1; tail -3 gpurun_out/r01_smoke.log
python scripts/ref_parity_report.py > gpurun_out/r01_ref_parity_report.txt 2>&1
for wl in ne27000_explicit ne27000_implicit; do python bench.py --workload $wl --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r01_bench_$wl.json; python scratch/pj.py < gpurun_out/r01_bench_$wl.json; done
for wl in cavity64_q2_implicit channel_48x32x32_implicit cavity128_q1_implicit; do python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r01_bench_nocpu_$wl.json; python scratch/pj.py < gpurun_out/r01_bench_nocpu_$wl.json; done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_bicgstab3 -s 4 -c 1 -o gpurun_out/prof_r01_bicg_ne27000 -f python bench.py --workload ne27000_implicit --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bicg.log 2>&1; tail -2 gpurun_out/ncu_bicg.log
