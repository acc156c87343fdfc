#!/usr/bin/env bash
# round-2 GPU call 10 (1 GPU): launch list of the bench step at HEAD (repo kernels only), ncu single-metric pass
mkdir -p gpurun_out
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:attn6|gemm|ln_modulate|skinny|patchify|cfg_euler|rmsnorm|timestep_emb|l1_sums|ew_' -c 4000 --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-vae --no-secondary --no-cpu-baseline > gpurun_out/r02_launches_step.out 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02_launches_step.csv; tail -c 300 gpurun_out/r02_launches_step.out
