#!/usr/bin/env bash
# round-2 GPU call 8 (1 GPU): full GPU suite, smoke, the bench line as the driver runs it, reference arm, 4 x 32 attention layout,
# step launch list (library kernels only)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -6 gpurun_out/r02_pytest3.log | cut -c1-300
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench2.err; cut -c1-1500 gpurun_out/r02_bench2.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench2_reference.json 2>&1; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02_bench2_reference.json
for v in 0x217c 0x410c 0x490c; do EA_ATTN_VARIANT=$v EA_ATTN_NO_COMPARE=1 timeout 200 python tools/bench_kernels.py attn >> gpurun_out/r02_attn_4x32.log 2>&1; done
cat gpurun_out/r02_attn_4x32.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:attn6|gemm|ln_modulate|skinny|patchify|cfg_euler|rmsnorm|timestep_emb|l1_sums|ew_' -c 4000 --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-vae --no-secondary --no-cpu-baseline > gpurun_out/r02_launches_step.out 2>&1
ls -la gpurun_out | tail -8
