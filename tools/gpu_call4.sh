#!/usr/bin/env bash
# round-2 GPU call 4 (1 GPU): full GPU suite with the new kernels, VAE micro-bench (halo conv + faster GroupNorm), attention
# polynomial fractions on the 3 x 64 layout, ncu launch lists and --set full captures of the three tensor-core kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -8 gpurun_out/r02_pytest2.log
timeout 400 python tools/bench_kernels.py vae > gpurun_out/r02_vae_v2.log 2>&1; cat gpurun_out/r02_vae_v2.log | cut -c1-300
EA_GN_LEGACY_APPLY=1 EA_VAE_SHAPES=90x160 timeout 300 python tools/bench_kernels.py vae > gpurun_out/r02_vae_v2_legacy_gn.log 2>&1; cat gpurun_out/r02_vae_v2_legacy_gn.log | cut -c1-300
for v in 0x210c 0x214c 0x217c 0x290c 0x294c 0x297c; do EA_ATTN_VARIANT=$v EA_ATTN_NO_COMPARE=1 timeout 200 python tools/bench_kernels.py attn >> gpurun_out/r02_attn_poly.log 2>&1; done
cat gpurun_out/r02_attn_poly.log
# ncu: launch lists (durations only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-vae --no-secondary --no-cpu-baseline > gpurun_out/r02_launches_step.out 2>&1
EA_VAE_SHAPES=90x160 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_vae_720p.csv \
    python tools/bench_kernels.py vae > /dev/null 2>&1
# ncu --set full: attention (bench shape), CTA-pair GEMM, halo convolution
EA_ATTN_SHAPE=2,48,47056,256 EA_ATTN_NO_COMPARE=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn6_kernel -s 2 -c 1 -o gpurun_out/r02_attn_3x64 \
    python tools/bench_kernels.py attn > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_tc -s 3 -c 2 -o gpurun_out/r02_gemm2 \
    python tools/bench_kernels.py gemm > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo -s 40 -c 3 -o gpurun_out/r02_conv3d_halo \
    python tools/try_conv_halo.py one 0x0 bench > /dev/null 2>&1
ls -la gpurun_out | tail -20
