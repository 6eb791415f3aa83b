for T in 1 2 4 8; do echo ZT=$T; CFD_Z_T=$T python bench.py --workload cavity64_q2_explicit --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done
for T in 1 2 4; do echo VT=$T; CFD_V_T=$T python bench.py --workload cavity64_q2_implicit --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python scratch/pj.py; done
python bench.py --workload channel_48x32x32_implicit --steps 3 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_channel.json | python scratch/pj.py
