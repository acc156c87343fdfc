#!/usr/bin/env bash
# round-2 GPU call 5 (4 GPUs): every N>1 path on 4 ranks, then the N=4 bench line (ONE video on 2 CFG branches x 2 sequence-parallel ranks)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tools/test_multigpu.py \
    > gpurun_out/r02_multigpu_4.log 2>&1; echo "multigpu rc=$?"
grep -a "world" gpurun_out/r02_multigpu_4.log | tail -1 | cut -c1-2000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 3 --warmup 3 \
    > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench4 rc=$?"
tail -c 1200 gpurun_out/r02_bench_n4.err; cut -c1-3500 gpurun_out/r02_bench_n4.json
