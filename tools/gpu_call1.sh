#!/usr/bin/env bash
# round-2 GPU call 1: full GPU test suite (all failures listed), bench line, kernel micro-benchmarks
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -30 gpurun_out/r02_pytest1.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench1.err
cat gpurun_out/r02_bench1.json | cut -c1-3000
timeout 300 python tools/bench_kernels.py attn > gpurun_out/r02_attn_base.log 2>&1
for v in 0x210c 0x211c 0x290c; do EA_ATTN_VARIANT=$v EA_ATTN_NO_COMPARE=1 timeout 200 python tools/bench_kernels.py attn >> gpurun_out/r02_attn_3x64.log 2>&1; done
timeout 300 python tools/bench_kernels.py gemm > gpurun_out/r02_gemm_base.log 2>&1
tail -20 gpurun_out/r02_attn_base.log gpurun_out/r02_attn_3x64.log gpurun_out/r02_gemm_base.log
timeout 1300 python tools/try_conv_halo.py > gpurun_out/r02_conv_halo_bringup.log 2>&1
grep -E "PASS|bench|rc|timeout" gpurun_out/r02_conv_halo_bringup.log | tail -40
