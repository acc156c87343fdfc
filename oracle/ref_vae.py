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
