#!/usr/bin/env bash
# round-2 GPU call 6 (8 GPUs): every N>1 path on 8 ranks, then the N=8 bench line (ONE video on 2 CFG branches x 4 sequence-parallel ranks)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/test_multigpu.py \
    > gpurun_out/r02_multigpu_8.log 2>&1; echo "multigpu rc=$?"
grep -a "world" gpurun_out/r02_multigpu_8.log | tail -1 | cut -c1-2000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 3 --warmup 3 \
    > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8 rc=$?"
tail -c 1200 gpurun_out/r02_bench_n8.err; grep -a '^{"metric' gpurun_out/r02_bench_n8.json | cut -c1-3500
