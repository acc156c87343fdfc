"""ORACLE helper (test infrastructure only): import the reference's own VAE Decoder from /root/reference.

Works only where /root/reference exists (the authoring container); used to validate oracle/vae.py and to mint
tests/golden/*.  Never imported on the GPU box or by the product path."""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = "/root/reference"
UP_BLOCK_TYPES = ("SpatialUpBlock3D", "SpatialTemporalUpBlock3D", "SpatialTemporalUpBlock3D", "SpatialTemporalUpBlock3D")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "easyanimate", "vae"))


def reference_decoder(cache_mag_vae: bool = True, mid_block_use_attention: bool = True, mini_batch_decoder: int = 1,
                      block_out_channels=(128, 256, 512, 512), latent_channels: int = 16):
    """Reference `Decoder` configured like v5.1 (config/easyanimate_video_v5.1_magvit_qwen.yaml:9-19 +
    vae/configs/autoencoder/autoencoder_kl_32x32x4_mag_v2.yaml:4-13)."""
    if not available():
        raise RuntimeError("/root/reference is not present here")
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_refshim")
    try:
        import diffusers  # noqa: F401  (a real install wins over the shim)
    except ImportError:
        if shim not in sys.path:
            sys.path.insert(0, shim)
    vae_pkg = os.path.join(REFERENCE_ROOT, "easyanimate", "vae")
    if vae_pkg not in sys.path:
        sys.path.insert(0, vae_pkg)  # `ldm` becomes importable without importing easyanimate/__init__ (needs diffusers)
    from ldm.models.omnigen_enc_dec import Decoder

    return Decoder(in_channels=latent_channels, out_channels=3, up_block_types=UP_BLOCK_TYPES,
                   block_out_channels=list(block_out_channels), mid_block_type="MidBlock3D",
                   mid_block_use_attention=mid_block_use_attention, mid_block_attention_type="spatial",
                   mid_block_num_attention_heads=1, layers_per_block=2, norm_num_groups=32, act_fn="silu",
                   num_attention_heads=1, slice_mag_vae=False, slice_compression_vae=False,
                   cache_compression_vae=False, cache_mag_vae=cache_mag_vae, spatial_group_norm=True,
                   mini_batch_decoder=mini_batch_decoder)


def reference_autoencoder(block_out_channels=(128, 256, 512, 512), mid_block_use_attention=True, use_tiling=False,
                          tile_sample_min_size=384, latent_channels=16):
    """The reference's own `AutoencoderKLMagvit` (easyanimate/models/autoencoder_magvit.py:59, executed unmodified) in the
    v5.1 configuration (config/easyanimate_video_v5.1_magvit_qwen.yaml:9-19): decode / tiled_decode wrapper around the
    Decoder above, post_quant_conv included.  Synthetic parent packages keep the package __init__ files (text encoders,
    pytorch_lightning, ...) from being executed."""
    import importlib
    import types

    if not available():
        raise RuntimeError("/root/reference is not present here")
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_refshim")
    try:
        import diffusers  # noqa: F401
    except ImportError:
        if shim not in sys.path:
            sys.path.insert(0, shim)
    for name, rel in (("easyanimate", "easyanimate"), ("easyanimate.models", "easyanimate/models"),
                      ("easyanimate.vae", "easyanimate/vae"), ("easyanimate.vae.ldm", "easyanimate/vae/ldm"),
                      ("easyanimate.vae.ldm.models", "easyanimate/vae/ldm/models")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            sys.modules[name] = pkg
    mod = importlib.import_module("easyanimate.models.autoencoder_magvit")
    down = ("SpatialDownBlock3D", "SpatialTemporalDownBlock3D", "SpatialTemporalDownBlock3D", "SpatialTemporalDownBlock3D")
    vae = mod.AutoencoderKLMagvit(
        in_channels=3, out_channels=3, down_block_types=down, up_block_types=UP_BLOCK_TYPES,
        block_out_channels=list(block_out_channels), latent_channels=latent_channels, layers_per_block=2,
        mid_block_type="MidBlock3D", mid_block_use_attention=mid_block_use_attention, mid_block_attention_type="spatial",
        mid_block_num_attention_heads=1, norm_num_groups=32, act_fn="silu", scaling_factor=0.7125,
        slice_mag_vae=False, slice_compression_vae=False, cache_compression_vae=False, cache_mag_vae=True,
        spatial_group_norm=True, mini_batch_encoder=4, mini_batch_decoder=1, use_tiling=use_tiling,
        tile_sample_min_size=tile_sample_min_size)
    return vae
