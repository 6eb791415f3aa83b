python -m pytest tests -m gpu -q 2>&1 | tail -4
for wl in cavity256_q1_implicit cavity128_q1_implicit cavity128_q1_explicit cavity64_q2_explicit cavity64_q2_implicit ne27000_explicit ne27000_implicit; do python bench.py --workload $wl --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r01_bench_$wl.json; python scratch/pj.py < gpurun_out/r01_bench_$wl.json; done
