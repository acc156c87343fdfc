#!/usr/bin/env bash
# round-2 GPU call 2 (2 GPUs): multi-GPU correctness of every N>1 path + the N=2 bench line
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/test_multigpu.py \
    > gpurun_out/r02_multigpu_2.log 2>&1; echo "multigpu rc=$?"
tail -5 gpurun_out/r02_multigpu_2.log | cut -c1-3000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 \
    > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r02_bench_n2.err; cut -c1-2500 gpurun_out/r02_bench_n2.json
