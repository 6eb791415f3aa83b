python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in ne27000_explicit ne27000_implicit cavity128_q1_implicit cavity64_q2_implicit; do python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done
for T in 1 2 4; do echo VT=$T; for wl in ne27000_explicit ne27000_implicit; do CFD_V_T=$T python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done; done
