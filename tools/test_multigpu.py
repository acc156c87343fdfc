"""2-GPU checks (run under torchrun on a B200 box): CFG-parallel sampler == single-GPU batch-of-2 step; tile-parallel
VAE tiled_decode == single-GPU tiled_decode; with EA_TEST_SP=1 also Ulysses sequence-parallel forward vs the single-GPU
forward (reported, not yet part of `ok`: it has not run on GPUs).  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/test_multigpu.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from easyanimate_b200 import AutoencoderKLMagvit, EasyAnimateSampler, EasyAnimateTransformer3DModel, rope_table

bf16 = torch.bfloat16


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    grp = dist.new_group(list(range(world)))
    torch.manual_seed(0)
    # ---- DiT: CFG-parallel pair vs batch of 2
    cfg = dict(num_attention_heads=4, attention_head_dim=64, in_channels=16, out_channels=16, patch_size=2, num_layers=2,
               time_embed_dim=128, add_norm_text_encoder=True, text_embed_dim=256, text_embed_dim_t5=None)
    with torch.device(dev):
        model = EasyAnimateTransformer3DModel(**cfg).to(bf16)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.normal_(1.0 if (p.dim() == 1 and n.endswith("weight")) else 0.0, 0.05)
    g = torch.Generator(device=dev).manual_seed(5)
    lat = torch.randn((1, 16, 3, 16, 24), device=dev, generator=g).to(bf16)
    emb = (torch.randn((2, 40, 256), device=dev, generator=g) * 3).to(bf16)
    rope = rope_table(128, 192, 3, device=dev)
    single = EasyAnimateSampler(model, guidance_scale=6.0)
    pair = EasyAnimateSampler(model, guidance_scale=6.0, cfg_group=grp)
    single.set_timesteps(4, device="cpu"); pair.set_timesteps(4, device="cpu")
    a, b = lat.clone(), lat.clone()
    for i in range(4):
        a = single.step(a, i, emb, rope)
        b = pair.step(b, i, emb, rope)
    torch.cuda.synchronize()
    d1 = (a.float() - b.float()).abs().max().item()
    # ---- DiT: Ulysses sequence parallelism over the same ranks vs the single-GPU forward (first GPU run: round 2)
    d3 = None
    if os.environ.get("EA_TEST_SP", "0") == "1":
        t = torch.tensor([937.0], device=dev).to(bf16)
        one = model(lat, t, encoder_hidden_states=emb[:1], image_rotary_emb=rope, return_dict=False)[0]
        model.set_sequence_parallel_group(grp)
        sp = model(lat, t, encoder_hidden_states=emb[:1], image_rotary_emb=rope, return_dict=False)[0]
        model.set_sequence_parallel_group(None)
        torch.cuda.synchronize()
        d3 = (one.float() - sp.float()).abs().max().item()
    # ---- VAE: tile-parallel tiled decode vs single-GPU tiled decode
    with torch.device(dev):
        vae = AutoencoderKLMagvit(latent_channels=16, cache_mag_vae=True, spatial_group_norm=True,
                                  mid_block_attention_type="spatial", block_out_channels=[64, 64, 128, 128],
                                  use_tiling=True, tile_sample_min_size=64).to(bf16)
    with torch.no_grad():
        for n, p in vae.named_parameters():
            if p.dim() >= 2:
                p.normal_(0, (1.0 / p[0].numel()) ** 0.5)
            elif "norm" in n and n.endswith("weight"):
                p.normal_(1.0, 0.05)
            else:
                p.normal_(0, 0.05)
    z = torch.randn((1, 16, 2, 14, 20), device=dev, generator=g).to(bf16)
    ref = vae.decode(z).sample
    vae.set_tile_parallel_group(grp)
    par = vae.decode(z).sample
    torch.cuda.synchronize()
    d2 = (ref.float() - par.float()).abs().max().item()
    ok = torch.tensor([float(d1 == 0.0 and d2 == 0.0 and bool(torch.isfinite(par).all()))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print({"cfg_parallel_max_abs_diff": d1, "tile_parallel_max_abs_diff": d2, "sequence_parallel_max_abs_diff": d3,
               "shape": tuple(par.shape), "ok": bool(ok.item())})
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()
