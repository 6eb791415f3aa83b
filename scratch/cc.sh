python -m pytest tests -m gpu -q 2>&1 | tail -2
for wl in ne27000_implicit cavity64_q2_implicit; do python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done
