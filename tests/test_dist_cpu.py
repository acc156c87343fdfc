"""world_size-2 gloo test of the CFG-parallel host logic (rank -> branch assignment, all_gather layout, identical
latents on both ranks, equality with the single-process batch-of-2 evaluation).  The transformer and the
CFG+Euler kernel are replaced by CPU stand-ins: this exercises the N>1 plumbing, not the kernels."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

bf16 = torch.bfloat16


class FakeTransformer(torch.nn.Module):
    """Deterministic stand-in with the transformer call signature; depends on latents, timestep and the embeds row."""

    def forward(self, hidden_states, timestep, encoder_hidden_states=None, image_rotary_emb=None, inpaint_latents=None,
                return_dict=True):
        bias = encoder_hidden_states.float().mean(dim=(1, 2)).view(-1, 1, 1, 1, 1)
        out = hidden_states.float() * 0.5 + bias + timestep.float().view(-1, 1, 1, 1, 1) * 1e-3
        return (out.to(hidden_states.dtype),)


def euler_cpu(noise_pred, latents, guidance_scale, sigma, sigma_next, use_cfg=True):
    """torch restatement of ea_cfg_euler_step (pipeline_easyanimate.py:1102-1111) for the CPU test."""
    if use_cfg:
        u, c = noise_pred.chunk(2)
        noise_pred = u + guidance_scale * (c - u)
    dt = torch.tensor(sigma_next, dtype=torch.float32) - torch.tensor(sigma, dtype=torch.float32)
    return (latents.float() + dt * noise_pred).to(noise_pred.dtype)


def _inputs():
    g = torch.Generator().manual_seed(0)
    return torch.randn(1, 4, 2, 4, 4, generator=g).to(bf16), torch.randn(2, 3, 8, generator=g).to(bf16)


def _single():
    from easyanimate_b200.pipeline import EasyAnimateSampler
    lat, emb = _inputs()
    s = EasyAnimateSampler(FakeTransformer(), guidance_scale=6.0, euler_fn=euler_cpu)
    s.set_timesteps(4, device="cpu")
    for i in range(4):
        lat = s.step(lat, i, emb, None)
    return lat


def _worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    dist.init_process_group("gloo", rank=rank, world_size=2)
    from easyanimate_b200.pipeline import EasyAnimateSampler
    grp = dist.new_group([0, 1])
    lat, emb = _inputs()
    s = EasyAnimateSampler(FakeTransformer(), guidance_scale=6.0, cfg_group=grp, euler_fn=euler_cpu)
    s.set_timesteps(4, device="cpu")
    for i in range(4):
        lat = s.step(lat, i, emb, None)
    q.put((rank, lat.float()))
    dist.barrier()
    dist.destroy_process_group()


def test_cfg_parallel_pair_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _single().float()
    assert torch.equal(got[0], got[1]), "both ranks of a CFG pair must hold identical latents"
    assert torch.equal(got[0], ref), "CFG-parallel must reproduce the batch-of-2 evaluation bit for bit"


def test_scheduler_mirror_matches_oracle():
    from oracle import dit
    from easyanimate_b200.scheduler import FlowMatchEulerDiscreteScheduler
    for shift, dyn in ((1.0, False), (3.0, False), (1.0, True)):
        ours = FlowMatchEulerDiscreteScheduler(shift=shift, use_dynamic_shifting=dyn)
        ref = dit.FlowMatchEulerScheduler(shift=shift, use_dynamic_shifting=dyn)
        for n in (25, 30, 50):
            ours.set_timesteps(n, device="cpu", mu=1.0)
            ref.set_timesteps(n, mu=1.0)
            assert torch.equal(ours.sigmas, ref.sigmas) and torch.equal(ours.timesteps, ref.timesteps)
    with pytest.raises(ValueError):
        FlowMatchEulerDiscreteScheduler(use_dynamic_shifting=True).set_timesteps(10)


def test_rope_table_mirror_matches_oracle():
    from oracle import dit
    from easyanimate_b200.pipeline import rope_table
    for (hh, ww, f) in ((720, 1280, 13), (512, 512, 13), (64, 64, 1), (384, 672, 7)):
        a, b = rope_table(hh, ww, f), dit.rope_for_video(hh, ww, f)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
