#!/bin/bash
# Round 2, GPU call 5 (8 GPUs of one box): torchrun bench at N = 8 and N = 4 (ONE catalogue of N x 10^4 objects split by
# the library's LPT partition), and one host process driving all 8 GPUs through cpg_multi.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_gpus_n8.txt
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2f_bench_n$n.json 2> gpurun_out/r2f_bench_n$n.err; echo "bench n$n rc=$?"
done
timeout 900 python profiles/multi_single_process.py 8 > gpurun_out/r2f_multi_single_process_n8.json 2> gpurun_out/r2f_multi_single_process_n8.err; echo "single-process multi rc=$?"
cat gpurun_out/r2f_multi_single_process_n8.json
