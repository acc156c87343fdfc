#!/usr/bin/env bash
# round-2 GPU call 11 (1 GPU): the GPU suite with the two tests added after call 9 (native pipeline call, CUDA-graph capture)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest5.log
tail -15 gpurun_out/r02_pytest5.log | cut -c1-600
