set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --workload ne27000_implicit --no-cpu-baseline > gpurun_out/r02k_bench_ne27000i.json 2> gpurun_out/r02k_bench_ne27000i.err
timeout 600 python bench.py --steps 3 --warmup 2 --workload channel_48x32x32_implicit --no-cpu-baseline > gpurun_out/r02k_bench_channel.json 2> gpurun_out/r02k_bench_channel.err
timeout 600 python bench.py --steps 5 --warmup 2 --workload cavity64_q2_implicit --no-cpu-baseline > gpurun_out/r02k_bench64_q2i.json 2> gpurun_out/r02k_bench64_q2i.err
CFD_VERBOSE=1 BENCH_VERBOSE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench256_verbose.json 2> gpurun_out/r02k_bench256_verbose.err
grep -a "cfdb200\|bench rank" gpurun_out/r02k_bench256_verbose.err
