#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/debug_sp.py \
    > gpurun_out/r02_debug_sp.log 2>&1; echo "debug rc=$?"
grep -E "^\[rank|Error|error" gpurun_out/r02_debug_sp.log | head -30 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/test_multigpu.py \
    > gpurun_out/r02_multigpu_2.log 2>&1; echo "multigpu rc=$?"
tail -3 gpurun_out/r02_multigpu_2.log | cut -c1-3000
