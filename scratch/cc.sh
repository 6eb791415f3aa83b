python -m pytest tests -m gpu -q 2>&1 | tail -3
for occ in 3 2; do for nc in 0 1; do echo OCC=$occ NOCOMP=$nc; if [ $nc = 1 ]; then export CFD_NO_COLCOMP=1; else unset CFD_NO_COLCOMP; fi; CFD_PCG_OCC=$occ python bench.py --workload cavity128_q1_implicit --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done; done
unset CFD_NO_COLCOMP
for wl in cavity256_q1_implicit cavity64_q2_explicit ne27000_explicit cavity128_q1_explicit; do python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done
