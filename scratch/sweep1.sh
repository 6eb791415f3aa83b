python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in ne27000_implicit cavity64_q2_implicit cavity128_q1_implicit ne27000_explicit; do python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done
