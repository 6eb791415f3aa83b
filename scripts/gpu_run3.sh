set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c_gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "two_gpu or time_loop_parity or momentum or both_cg or spmv_and_pcg" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
CG_VARIANTS=1,2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py > gpurun_out/r02c_mgpu2.log 2>&1
tail -30 gpurun_out/r02c_mgpu2.log
for w in ne27000_explicit ne27000_implicit q1_128; do timeout 300 python scripts/krylov_phases.py pcg $w >> gpurun_out/r02c_pcg_phases.txt 2>&1; done
cat gpurun_out/r02c_pcg_phases.txt
for v in 1 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$v bench.py --gpus 2 --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit > gpurun_out/r02c_bench128_n2_v$v.json 2> gpurun_out/r02c_bench128_n2_v$v.err
done
for v in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --cg-variant $v --workload ne27000_explicit --no-cpu-baseline > gpurun_out/r02c_bench_ne27000e_v$v.json 2> gpurun_out/r02c_bench_ne27000e_v$v.err
timeout 600 python bench.py --steps 5 --warmup 2 --cg-variant $v --workload cavity128_q1_implicit --no-cpu-baseline > gpurun_out/r02c_bench128_n1_v$v.json 2> gpurun_out/r02c_bench128_n1_v$v.err
done
