"""diffusers.models.embeddings: sinusoidal timestep projection, the 2-layer timestep MLP and the real-valued rotary
application (embeddings.py get_timestep_embedding / Timesteps / TimestepEmbedding / apply_rotary_emb, 0.30)."""
import math

import torch
from torch import nn

from ._placeholder import placeholder


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1, scale=1,
                           max_period=10000):
    assert len(timesteps.shape) == 1, "Timesteps should be a 1d-array"
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(start=0, end=half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift, scale=1):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos = num_channels, flip_sin_to_cos
        self.downscale_freq_shift, self.scale = downscale_freq_shift, scale

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift, scale=self.scale)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None, cond_proj_dim=None,
                 sample_proj_bias=True):
        super().__init__()
        assert act_fn == "silu" and post_act_fn is None and cond_proj_dim is None
        self.linear_1 = nn.Linear(in_channels, time_embed_dim, sample_proj_bias)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim, sample_proj_bias)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


def apply_rotary_emb(x, freqs_cis, use_real=True, use_real_unbind_dim=-1):
    assert use_real and use_real_unbind_dim == -1
    cos, sin = freqs_cis  # [S, D]
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)  # [B, H, S, D//2]
    x_rotated = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rotated.float() * sin).to(x.dtype)


PatchEmbed = placeholder("PatchEmbed")
PixArtAlphaTextProjection = placeholder("PixArtAlphaTextProjection")
SinusoidalPositionalEmbedding = placeholder("SinusoidalPositionalEmbedding")
CombinedTimestepLabelEmbeddings = placeholder("CombinedTimestepLabelEmbeddings")
PixArtAlphaCombinedTimestepSizeEmbeddings = placeholder("PixArtAlphaCombinedTimestepSizeEmbeddings")


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    """diffusers 0.30/0.31 embeddings.get_1d_sincos_pos_embed_from_grid (MAE): [M] positions -> [M, embed_dim] (sin | cos)."""
    import numpy as np
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    import numpy as np
    emb_h = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0])
    emb_w = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])
    return np.concatenate([emb_h, emb_w], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False, extra_tokens=0, interpolation_scale=1.0, base_size=16):
    """diffusers 0.30/0.31 embeddings.get_2d_sincos_pos_embed restated from its published algorithm (the reference calls it
    with (inner_dim, (post_patch_height, post_patch_width)), transformer3d.py:1424): numpy float64 [H*W, embed_dim]."""
    import numpy as np
    if isinstance(grid_size, int):
        grid_size = (grid_size, grid_size)
    grid_h = np.arange(grid_size[0], dtype=np.float32) / (grid_size[0] / base_size) / interpolation_scale
    grid_w = np.arange(grid_size[1], dtype=np.float32) / (grid_size[1] / base_size) / interpolation_scale
    grid = np.meshgrid(grid_w, grid_h)  # w goes first
    grid = np.stack(grid, axis=0).reshape([2, 1, grid_size[1], grid_size[0]])
    pos_embed = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    if cls_token and extra_tokens > 0:
        pos_embed = np.concatenate([np.zeros([extra_tokens, embed_dim]), pos_embed], axis=0)
    return pos_embed


def get_3d_sincos_pos_embed(*a, **k):
    raise NotImplementedError("diffusers shim: get_3d_sincos_pos_embed is not on the EasyAnimateV5.1 path")


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, temporal_size, theta=10000, use_real=True):
    """diffusers 0.30/0.31 embeddings.get_3d_rotary_pos_embed (called at pipeline_easyanimate.py:1006-1009): ONE restatement,
    kept in oracle/dit.py (third-party, unpinned - tests/test_diffusers_pin.py compares it with a real diffusers if present)."""
    from oracle.dit import get_3d_rotary_pos_embed as _impl

    assert use_real
    return _impl(embed_dim, crops_coords, grid_size, temporal_size, theta)


def get_2d_rotary_pos_embed(*a, **k):
    raise NotImplementedError("diffusers shim: get_2d_rotary_pos_embed is not on the EasyAnimateV5.1 path (3d_rope)")
