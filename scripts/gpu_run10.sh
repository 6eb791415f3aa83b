set -x
mkdir -p gpurun_out
for p in csr ebe; do
timeout 600 python bench.py --steps 3 --warmup 3 --workload cavity128_q1_explicit --path $p --no-cpu-baseline > gpurun_out/r02j_bench128_q1_explicit_$p.json 2> gpurun_out/r02j_bench128_q1_explicit_$p.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload cavity64_q2_explicit --path $p --no-cpu-baseline > gpurun_out/r02j_bench64_q2_explicit_$p.json 2> gpurun_out/r02j_bench64_q2_explicit_$p.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --path $p --no-cpu-baseline > gpurun_out/r02j_bench_ne27000_explicit_$p.json 2> gpurun_out/r02j_bench_ne27000_explicit_$p.err
done
CFD_VERBOSE=1 BENCH_VERBOSE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench256_verbose.json 2> gpurun_out/r02j_bench256_verbose.err
grep -a "cfdb200\|bench rank" gpurun_out/r02j_bench256_verbose.err
