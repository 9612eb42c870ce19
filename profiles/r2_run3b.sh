#!/bin/bash
# Round 2, GPU call 3b (2 GPUs): the sharded path on two real devices (tests/test_gpu_multi.py picks [0, 1] up by itself),
# then the torchrun bench at N = 2: ONE catalogue of 2 x 10^4 objects split by the library's LPT partition.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus_n2.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2c_pytest_n2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest_n2.log
tail -3 gpurun_out/r2c_pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1200 gpurun_out/r2c_bench_n2.json
timeout 600 python profiles/multi_single_process.py 2 > gpurun_out/r2c_multi_single_process_n2.json 2> gpurun_out/r2c_multi_single_process_n2.err; echo "single-process multi rc=$?"
cat gpurun_out/r2c_multi_single_process_n2.json
