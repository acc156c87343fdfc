#!/usr/bin/env bash
# round-2 GPU call 1: full GPU test suite, new bench line, kernel micro-benchmarks, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench1.err
python tools/bench_kernels.py attn > gpurun_out/r02_attn_base.log 2>&1
for v in 0x210c 0x211c 0x290c; do EA_ATTN_VARIANT=$v EA_ATTN_NO_COMPARE=1 python tools/bench_kernels.py attn >> gpurun_out/r02_attn_3x64.log 2>&1; done
python tools/bench_kernels.py gemm > gpurun_out/r02_gemm_base.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-vae --no-secondary --no-cpu-baseline > gpurun_out/r02_launches_step.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_tc -s 3 -c 2 -o gpurun_out/r02_gemm2 \
    python tools/bench_kernels.py gemm > /dev/null 2>&1
EA_VAE_SHAPES=48x48 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 29 -c 5 -o gpurun_out/r02_conv3d \
    python tools/bench_kernels.py vae > /dev/null 2>&1
EA_VAE_SHAPES=48x48 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_vae_tile.csv \
    python tools/bench_kernels.py vae > /dev/null 2>&1
ls -la gpurun_out | tail -15
