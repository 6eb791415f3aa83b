python -m pytest tests -m gpu -q 2>&1 | tail -3
for nd in 0 1; do echo NODYN=$nd; if [ $nd = 1 ]; then export CFD_NO_DYN=1; else unset CFD_NO_DYN; fi; for wl in cavity128_q1_implicit cavity256_q1_implicit; do python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done; done
