"""world_size-2 gloo check of the VAE tile-parallel plumbing (partition of the reference's tiled_decode tiles over
ranks, padding, all-gather layout, reassembly order) with CPU stand-ins for the decoder pass and the copy kernel."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

bf16 = torch.bfloat16


def _fake_decode(self, zt):  # [C,T,h,w] -> [1,3,T',8h,8w]: deterministic function of the tile content
    C, T, h, w = zt.shape
    base = zt.float().mean(dim=0)[:1].repeat_interleave(8, dim=1).repeat_interleave(8, dim=2)  # [1,8h,8w]
    return (base[None, None].expand(1, 3, 4 * (T - 1) + 1, 8 * h, 8 * w) + 1.0).to(bf16).contiguous()


def _copy_cpu(src, dst, rows, cols, r0, c0):
    dst[..., r0:r0 + rows, c0:c0 + cols] = src[..., :rows, :cols]


def _tile_worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    dist.init_process_group("gloo", rank=rank, world_size=2)
    from easyanimate_b200 import vae_ops
    from easyanimate_b200.autoencoder_magvit import AutoencoderKLMagvit

    AutoencoderKLMagvit._decode_one = _fake_decode
    vae_ops.copy2d = _copy_cpu
    m = AutoencoderKLMagvit(latent_channels=16, cache_mag_vae=True, spatial_group_norm=True, mid_block_attention_type="spatial",
                            block_out_channels=[64, 64, 128, 128], use_tiling=True, tile_sample_min_size=64)
    z = torch.randn(16, 2, 14, 20, generator=torch.Generator().manual_seed(3)).to(bf16)
    tl, ov = m.tile_latent_min_size, int(m.tile_latent_min_size * 0.75)
    coords = [[(i, j) for j in range(0, 20, ov)] for i in range(0, 14, ov)]
    rows, corner = m._decode_tiles_parallel(z, coords, tl, dist.new_group([0, 1]))
    ok = True
    for row_c, row_t in zip(coords, rows):
        for (i, j), t in zip(row_c, row_t):
            ok &= torch.equal(t, _fake_decode(m, z[:, :, i:i + tl, j:j + tl]))
    ok &= torch.equal(corner, _fake_decode(m, z[:, :, -tl:, -tl:]))
    q.put((rank, bool(ok), len(rows), len(rows[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_vae_tile_parallel_gather_reassembles_every_tile():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_tile_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res) and res[0][2:] == (3, 4), res


# ---- the real module (all decoder layers, blending, corner pass) with torch stand-ins for the kernels (tests/cpu_ops.py) ----
def _install_cpu_vae_ops():
    from easyanimate_b200 import ops, vae_ops
    from tests import cpu_ops
    for name in ("conv3d_causal", "prepare_latents", "groupnorm", "upsample2x", "spatial_attention", "tile_blend", "copy2d",
                 "corner_blend"):
        setattr(vae_ops, name, getattr(cpu_ops, name))
    ops.gemm = cpu_ops.gemm


def _real_decode(group):
    from oracle import vae
    from easyanimate_b200.autoencoder_magvit import AutoencoderKLMagvit
    boc = [64, 64, 128, 128]
    ob = vae.init_weights_(vae.OracleAutoencoderKLMagvit(block_out_channels=boc, use_tiling=True, tile_sample_min_size=64), 21)
    m = AutoencoderKLMagvit(latent_channels=16, cache_mag_vae=True, spatial_group_norm=True, mid_block_attention_type="spatial",
                            block_out_channels=boc, use_tiling=True, tile_sample_min_size=64).to(bf16)
    m.load_state_dict({k: v.to(bf16) for k, v in ob.state_dict().items()}, strict=False)
    m.set_tile_parallel_group(group)
    z = torch.randn(16, 2, 14, 13, generator=torch.Generator().manual_seed(6)).to(bf16)
    with torch.no_grad():
        return m._tiled_decode_one(z)


def _real_worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    dist.init_process_group("gloo", rank=rank, world_size=2)
    _install_cpu_vae_ops()
    out = _real_decode(dist.new_group([0, 1]))
    q.put((rank, out.float()))
    dist.barrier()
    dist.destroy_process_group()


def test_vae_tile_parallel_decode_equals_single_process_tiled_decode():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 90)
    procs = [ctx.Process(target=_real_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import importlib
    from easyanimate_b200 import ops, vae_ops
    _install_cpu_vae_ops()
    try:
        single = _real_decode(None).float()
    finally:
        importlib.reload(ops)
        importlib.reload(vae_ops)
    assert single.shape == (1, 3, 5, 112, 104)
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][1], single)
