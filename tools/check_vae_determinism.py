"""Run-to-run determinism of the VAE decode kernels on one GPU (tile decode twice, tiled decode twice, padded-buffer
round trip as used by the tile-parallel path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyanimate_b200 import AutoencoderKLMagvit, vae_ops
bf16 = torch.bfloat16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with torch.device(dev):
    vae = AutoencoderKLMagvit(latent_channels=16, cache_mag_vae=True, spatial_group_norm=True, mid_block_attention_type="spatial",
                              block_out_channels=[64, 64, 128, 128], use_tiling=True, tile_sample_min_size=64).to(bf16)
with torch.no_grad():
    for n, p in vae.named_parameters():
        if p.dim() >= 2: p.normal_(0, (1.0 / p[0].numel()) ** 0.5)
        elif "norm" in n and n.endswith("weight"): p.normal_(1.0, 0.05)
        else: p.normal_(0, 0.05)
g = torch.Generator(device=dev).manual_seed(5)
z = torch.randn((1, 16, 2, 14, 20), device=dev, generator=g).to(bf16)
a = vae.decode(z).sample; b = vae.decode(z).sample
print("tiled decode twice: max abs diff", (a.float() - b.float()).abs().max().item())
tl = vae.tile_latent_min_size
for (i, j) in [(0, 0), (6, 6), (12, 18), (12, 0), (0, 18)]:
    t = z[0][:, :, i:i + tl, j:j + tl].contiguous()
    o1 = vae._decode_one(t); o2 = vae._decode_one(t)
    # decode again after other work of a different shape in between
    _ = vae._decode_one(z[0][:, :, 0:tl, 0:tl].contiguous())
    o3 = vae._decode_one(t)
    print("tile", (i, j), tuple(t.shape), "d12", (o1.float() - o2.float()).abs().max().item(), "d13", (o1.float() - o3.float()).abs().max().item())
    buf = torch.zeros((1, 3, o1.shape[2], 8 * tl, 8 * tl), device=dev, dtype=bf16)
    vae_ops.copy2d(o1, buf, o1.shape[3], o1.shape[4], 0, 0)
    back = buf[:, :, :, :o1.shape[3], :o1.shape[4]]
    print("   padded round trip equal:", torch.equal(back, o1))
