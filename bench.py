#!/usr/bin/env python
"""bench.py — EasyAnimateV5.1 sampling hot path on B200 (BASELINE.json metric: denoising-steps/sec @49f·720p bf16).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches one rank per GPU for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle port), rank 0 only

A "step" is ONE scheduler step of the denoise loop WITH classifier-free guidance: two MMDiT forwards (batch of 2), the
CFG combine and the flow-matching Euler update (pipeline_easyanimate.py:1069-1111).  Workload at N=1: BASELINE
configs[1] read as SURVEY.md §8(d) resolves it — "7B" synthetic MMDiT (d=3072, 48 heads, 28 layers, 6.87 B params),
49 frames @720x1280 => latent 13x90x160 => 46 800 video + 256 text tokens, synthetic text embeds, random-init weights.
Prints ONE JSON line (contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {
    # name: (latent F,h,w, layers, heads)
    "R720_7B": dict(F=13, h=90, w=160, layers=28, heads=48),
    "R720_12B": dict(F=13, h=90, w=160, layers=48, heads=48),
    "R512_7B": dict(F=13, h=64, w=64, layers=28, heads=48),
    "tiny": dict(F=3, h=16, w=24, layers=2, heads=4),
}
S_TEXT, E_TEXT, GUIDANCE = 256, 3584, 6.0
_CPU_THREADS = None


def dit_flops_per_forward(F, h, w, layers, heads, c_in=16, s_t=S_TEXT, e_text=E_TEXT):
    """BASELINE.md §2 closed form (MAC = 2 FLOP; softmax/LN/GELU excluded)."""
    d = heads * 64
    s_v = F * (h // 2) * (w // 2)
    s = s_v + s_t
    return layers * (24 * s * d * d + 4 * s * s * d) + 2 * s_v * 4 * c_in * d + 2 * s_t * e_text * d + 2 * s_v * d * 64


def model_cfg(p, text_dim=E_TEXT):
    return dict(num_attention_heads=p["heads"], attention_head_dim=64, in_channels=16, out_channels=16, patch_size=2,
                num_layers=p["layers"], time_embed_dim=512, add_norm_text_encoder=True, text_embed_dim=text_dim,
                text_embed_dim_t5=None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(seconds_budget: float = 15.0, threads: int | None = None):
    """Reference CPU path = the oracle restatement (diffusers is not installable offline): one
    EasyAnimateDiTBlock-deep model (d=3072, 48 heads) on 1 024 video + 256 text tokens, CFG batch 2, bf16, all host
    cores.  Returns (flop/s, description, cores, per-rep seconds)."""
    import torch
    from oracle import dit
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    p = dict(F=1, h=64, w=64, layers=1, heads=48)
    if threads is None:
        # "all the host threads it can use": torch's bf16 CPU kernels stop scaling (and regress badly) well before 128
        # threads on this small sample, so probe a few team sizes once and keep the fastest.
        global _CPU_THREADS
        if _CPU_THREADS is None:
            best = (float("inf"), avail)
            probe = dit.OracleTransformer3D(**model_cfg(dict(p, heads=8))).to(torch.bfloat16)
            x = torch.randn(2, 16, 1, 32, 32).to(torch.bfloat16)
            e = torch.randn(2, 64, E_TEXT).to(torch.bfloat16)
            rp = dit.rope_for_video(256, 256, 1)
            tt = torch.tensor([500.0, 500.0]).to(torch.bfloat16)
            for n in sorted({avail, max(1, avail // 2), max(1, avail // 4), min(avail, 32), min(avail, 16), min(avail, 8)}):
                torch.set_num_threads(n)
                with torch.no_grad():
                    probe(x, tt, encoder_hidden_states=e, image_rotary_emb=rp)
                    t0 = time.perf_counter()
                    probe(x, tt, encoder_hidden_states=e, image_rotary_emb=rp)
                    dt = time.perf_counter() - t0
                if dt < best[0]:
                    best = (dt, n)
            _CPU_THREADS = best[1]
        threads = _CPU_THREADS
    torch.set_num_threads(threads)
    m = dit.OracleTransformer3D(**model_cfg(p)).to(torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(2, 16, p["F"], p["h"], p["w"], generator=g).to(torch.bfloat16)
    enc = torch.randn(2, S_TEXT, E_TEXT, generator=g).to(torch.bfloat16)
    rope = dit.rope_for_video(p["h"] * 8, p["w"] * 8, p["F"])
    t = torch.tensor([500.0, 500.0]).to(torch.bfloat16)
    fl = 2 * dit_flops_per_forward(**p)
    times = []
    with torch.no_grad():
        m(lat, t, encoder_hidden_states=enc, image_rotary_emb=rope)  # warm-up
        t_end = time.perf_counter() + seconds_budget
        while time.perf_counter() < t_end or len(times) < 2:
            t0 = time.perf_counter()
            m(lat, t, encoder_hidden_states=enc, image_rotary_emb=rope)
            times.append(time.perf_counter() - t0)
    per = statistics.median(times)
    desc = (f"oracle/dit.py (torch-cpu bf16, {threads} of {avail} available threads — fastest team size probed): 1 MMDiT block d=3072/48 heads, CFG batch 2, 1024 video + "
            f"256 text tokens, {len(times)} reps, median {per:.3f} s/rep; steps/s extrapolated by FLOPs to the workload")
    return fl / per, desc, threads, per


def vae_decode_line(preset, dev, reps=3):
    """Secondary metric of the path: AutoencoderKLMagvit.decode of the workload's latent, untiled
    whole-sequence decode on one GPU, output MPix/s (device-resident latents, CUDA events, after one warm-up)."""
    import torch
    from easyanimate_b200.autoencoder_magvit import AutoencoderKLMagvit
    bf16 = torch.bfloat16
    with torch.device(dev):
        vae = AutoencoderKLMagvit(latent_channels=16, cache_mag_vae=True, spatial_group_norm=True,
                                  mid_block_attention_type="spatial", mini_batch_decoder=1, scaling_factor=0.7125).to(bf16)
    with torch.no_grad():
        for n, p in vae.named_parameters():
            if p.dim() >= 2:
                p.normal_(0, (1.0 / p[0].numel()) ** 0.5)
            elif "norm" in n and n.endswith("weight"):
                p.normal_(1.0, 0.05)
            else:
                p.normal_(0, 0.05)
    vae.use_tiling = False
    z = torch.randn((1, 16, preset["F"], preset["h"], preset["w"]), device=dev).to(bf16)
    out = vae.decode(z).sample
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vae.decode(z)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    mpix = out.shape[2] * out.shape[3] * out.shape[4] / 1e6
    return {"value": mpix / ms * 1e3, "unit": "MPix/s", "ms": ms, "frames": int(out.shape[2]),
            "resolution": [int(out.shape[3]), int(out.shape[4])], "mode": "untiled whole-sequence decode, random-init weights",
            "finite": bool(torch.isfinite(out).all())}


def run_reference_arm(args, preset):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flops_step = 2 * dit_flops_per_forward(**preset)
    per_step = []
    rate, desc, cores, per = cpu_oracle_rate(seconds_budget=2.0)
    for _ in range(args.warmup):
        cpu_oracle_rate(seconds_budget=0.0)
    for _ in range(args.steps):
        r, _, _, _ = cpu_oracle_rate(seconds_budget=0.0)
        per_step.append(flops_step / r)
    ms = statistics.mean(per_step) * 1e3
    value = 1e3 / ms
    line = {"impl": "reference", "metric": "denoising-steps/sec @49f·720p bf16 (CFG step = 2 MMDiT forwards)", "value": value,
            "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.preset, **preset},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default=os.environ.get("EA_BENCH_PRESET", "R720_7B"), choices=sorted(PRESETS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vae", action="store_true")
    args = ap.parse_args()
    preset = PRESETS[args.preset]
    if args.impl == "reference":
        return run_reference_arm(args, preset)

    import torch
    from easyanimate_b200 import _lib, ops
    from easyanimate_b200.pipeline import EasyAnimateSampler, rope_table
    from easyanimate_b200.transformer3d import EasyAnimateTransformer3DModel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    assert args.warmup >= 3 or args.preset == "tiny", "timing rules: at least 3 warm-up steps"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg_group = sp_group = None
    single_video = os.environ.get("EA_BENCH_TOPOLOGY", "") == "single_video"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        assert world % 2 == 0, "N>1 runs CFG-parallel pairs: N must be even"
        if single_video and world >= 4:
            # opt-in (EA_BENCH_TOPOLOGY=single_video): ONE video on all N GPUs = 2 CFG branches x N/2 sequence-parallel
            # ranks (Ulysses, easyanimate_b200/sequence_parallel.py).  Host logic verified on gloo
            # (tests/test_dist_sp_cpu.py); not yet validated on GPUs, hence not the default.
            P = world // 2
            for b in range(2):  # every rank must create every group
                grp = dist.new_group(list(range(b * P, (b + 1) * P)))
                if rank // P == b:
                    sp_group = grp
            for r in range(P):
                grp = dist.new_group([r, r + P])
                if rank % P == r:
                    cfg_group = grp
        else:
            for g0 in range(0, world, 2):  # every rank must create every group
                grp = dist.new_group([g0, g0 + 1])
                if rank in (g0, g0 + 1):
                    cfg_group = grp

    bf16 = torch.bfloat16
    F, h, w = preset["F"], preset["h"], preset["w"]
    torch.manual_seed(1234)  # identical weights on every rank
    with torch.device(dev):
        model = EasyAnimateTransformer3DModel(**model_cfg(preset)).to(bf16)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if prm.dim() == 1 and name.endswith("weight"):
                prm.normal_(1.0, 0.02)
            else:
                prm.normal_(0.0, 0.02)
    n_params = sum(p.numel() for p in model.parameters())
    if sp_group is not None:
        model.set_sequence_parallel_group(sp_group)
    sampler = EasyAnimateSampler(model, guidance_scale=GUIDANCE, cfg_group=cfg_group)
    total_steps = args.warmup + 2 * args.steps + 4
    sampler.set_timesteps(max(total_steps, 30), device="cpu")
    rope = rope_table(h * 8, w * 8, F, device=dev)
    video_seed = 100 + (0 if sp_group is not None else (rank // 2 if world > 1 else 0))  # one video per CFG pair
    g = torch.Generator(device=dev).manual_seed(video_seed)
    latents = torch.randn((1, 16, F, h, w), device=dev, generator=g).to(bf16)
    embeds = (torch.randn((2, S_TEXT, E_TEXT), device=dev, generator=g) * 10).to(bf16)  # cat(negative, positive)

    def sync_all():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    step_i = 0
    for _ in range(args.warmup):
        latents = sampler.step(latents, step_i, embeds, rope); step_i += 1

    # ---- timed region 1: device-resident inputs
    attn_events = []
    ops.ATTN_TIMING = attn_events
    clocks = ClockSampler(local_rank)
    sync_all()
    clocks.start()
    launches0 = _lib.ea_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        latents = sampler.step(latents, step_i, embeds, rope); step_i += 1
    e1.record()
    sync_all()
    launches = _lib.ea_launch_count() - launches0
    clock_info = clocks.stop()
    ops.ATTN_TIMING = None
    ms_total = e0.elapsed_time(e1)
    attn_ms = [a.elapsed_time(b) for a, b in attn_events]

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D inputs + D2H result every step)
    lat_host = torch.empty(latents.shape, dtype=bf16, pin_memory=True).copy_(latents)
    emb_host = torch.empty(embeds.shape, dtype=bf16, pin_memory=True).copy_(embeds)
    out_host = torch.empty(latents.shape, dtype=bf16, pin_memory=True)
    sampler.step_from_host(lat_host, step_i, emb_host, rope, out_host, device=dev); step_i += 1  # warm
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        sampler.step_from_host(lat_host, step_i, emb_host, rope, out_host, device=dev); step_i += 1
        lat_host, out_host = out_host, lat_host
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)

    t_dev = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t_dev[0]), float(t_dev[1])
    n_videos = 1 if sp_group is not None else max(1, world // 2)
    ms_per_step = ms_total / args.steps
    value = n_videos * args.steps / (ms_total / 1e3)
    e2e_value = n_videos * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        S = F * (h // 2) * (w // 2) + S_TEXT
        B_attn = 2 if cfg_group is None else 1
        heads_attn = preset["heads"] // (world // 2) if sp_group is not None else preset["heads"]  # Ulysses: heads sharded
        attn_flops = 4.0 * B_attn * heads_attn * S * S * 64
        attn_avg_ms = statistics.mean(attn_ms) if attn_ms else None
        achieved = attn_flops / (attn_avg_ms * 1e-3) / 1e12 if attn_avg_ms else None
        flops_step = 2 * dit_flops_per_forward(**preset)
        act_bytes = B_attn * S * preset["heads"] * 64 * 2  # one [B, S, d] bf16 activation
        line = {
            "metric": "denoising-steps/sec @49f·720p bf16 (CFG step = 2 MMDiT forwards)", "value": value, "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if sp_group is not None else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": args.preset, "model": f"MMDiT d={preset['heads'] * 64} heads={preset['heads']} layers={preset['layers']}",
                       "params": n_params, "latent": [1, 16, F, h, w], "video": f"{4 * (F - 1) + 1}f {h * 8}x{w * 8}",
                       "tokens": S, "text_tokens": S_TEXT, "guidance_scale": GUIDANCE, "scheduler": "flow-match Euler shift=1",
                       "parallelism": "single GPU (CFG batch 2)" if world == 1 else
                       (f"1 video on {world} GPUs: 2 CFG branches x {world // 2} sequence-parallel ranks (Ulysses)" if sp_group is not None else
                        f"{n_videos} video(s) x CFG-parallel pair (1 all_gather of {latents.numel() * 2 / 1e6:.2f} MB per step)"),
                       "l2": ("inputs_exceed_L2" if act_bytes > 126e6 else "inputs_fit_L2 (not a timing configuration)") +
                             f" (one activation tensor of a forward is {act_bytes / 1e6:.0f} MB, weights {n_params * 2 / 1e9:.1f} GB)",
                       "tflop_per_step": flops_step / 1e12},
            "model_tflops": flops_step * value / 1e12, "forwards_per_s": 2.0 * value,  # a CFG step is two MMDiT forwards
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": lat_host.numel() * 2 + emb_host.numel() * 2,
                    "d2h_bytes_per_step": out_host.numel() * 2},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": {"kernel": "a6::attn6_kernel (joint text+video attention, hd=64, ea_attn_args.variant 0x%x)" % ops.ATTN_VARIANT, "bound": "tensor",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                         "peak_source": peak_src + " (of measured)",
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the ncu --set full capture of
                         # this kernel on this shape (profiles/r01_ncu_attn_v6_step_summary.txt: 2.3135 GB at B=2, i.e.
                         # exactly Q+K+V+O once); not re-measured live
                         "traffic": (1.746926e9 + 0.566597e9) * B_attn / 2 if args.preset == "R720_7B" else None,
                         "traffic_algorithmic": 4.0 * B_attn * heads_attn * S * 64 * 2,
                         "launches_timed": len(attn_ms), "avg_ms": attn_avg_ms, "flops_per_launch": attn_flops,
                         "share_of_step": (sum(attn_ms) / ms_total) if attn_ms else None},
        }
        if world == 1 and not args.no_vae:
            line["vae_decode"] = vae_decode_line(preset, dev)
        if world == 1 and not args.no_cpu_baseline:
            rate, desc, cores, _ = cpu_oracle_rate(seconds_budget=12.0)
            line["cpu_baseline"] = {"value": rate / flops_step, "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
