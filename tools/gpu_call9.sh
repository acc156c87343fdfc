#!/usr/bin/env bash
# round-2 GPU call 9 (1 GPU, after the 4 x 32 attention layout was withdrawn): full GPU suite under the per-test watchdog, smoke,
# the bench line as the driver runs it, the reference arm.  Every step has its own timeout; the whole call is capped by gpurun.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest4.log
tail -8 gpurun_out/r02_pytest4.log | cut -c1-400
timeout 150 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
timeout 420 python bench.py > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02_bench3.err; cut -c1-1200 gpurun_out/r02_bench3.json
timeout 150 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench3_reference.json 2>&1; echo "ref rc=$?"; cut -c1-500 gpurun_out/r02_bench3_reference.json
