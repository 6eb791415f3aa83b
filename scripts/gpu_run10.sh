set -x
mkdir -p gpurun_out
free -g > gpurun_out/r02j_mem.txt; nproc >> gpurun_out/r02j_mem.txt
timeout 300 python scripts/operator_bench.py 128 8 > gpurun_out/r02j_operators_q1_128.jsonl 2> gpurun_out/r02j_operators_q1_128.err
timeout 300 python scripts/operator_bench.py 64 27 > gpurun_out/r02j_operators_q2_64.jsonl 2> gpurun_out/r02j_operators_q2_64.err
cat gpurun_out/r02j_operators_*.jsonl
for p in csr ebe; do
timeout 600 python bench.py --steps 3 --warmup 3 --workload cavity128_q1_explicit --path $p --no-cpu-baseline > gpurun_out/r02j_bench128_q1_explicit_$p.json 2> gpurun_out/r02j_bench128_q1_explicit_$p.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload cavity64_q2_explicit --path $p --no-cpu-baseline > gpurun_out/r02j_bench64_q2_explicit_$p.json 2> gpurun_out/r02j_bench64_q2_explicit_$p.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload ne27000_explicit --path $p --no-cpu-baseline > gpurun_out/r02j_bench_ne27000_explicit_$p.json 2> gpurun_out/r02j_bench_ne27000_explicit_$p.err
done
