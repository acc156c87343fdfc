"""Mint golden vectors from the REFERENCE ITSELF (run in the authoring container, where /root/reference exists).

VAE: the reference's own `Decoder` (easyanimate/vae/ldm/models/omnigen_enc_dec.py) is imported through
oracle/ref_vae.py and run in its v5.1 execution mode (cache_mag_vae=True, one latent frame per chunk,
padding_flag 3/4) in fp32 on CPU.  Weights come from oracle.vae.init_weights_(seed) so the fixtures stay small:
only inputs and outputs are stored.

DiT: the reference transformer cannot be imported here (diffusers missing), so DiT fixtures are produced by the
oracle restatement and are marked `source=oracle` (parity unpinned, see oracle/dit.py).

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import dit, ref_vae, vae  # noqa: E402

VAE_CASES = {
    # name: (block_out_channels, mid_attention, z shape, seed)
    "vae_small_attn": ((64, 64, 128, 128), True, (1, 16, 3, 8, 8), 11),
    "vae_small_noattn_ragged": ((64, 64, 128, 128), False, (1, 16, 2, 6, 10), 12),
    "vae_full_arch": ((128, 256, 512, 512), True, (1, 16, 2, 8, 8), 13),
}

DIT_CFG = dict(num_attention_heads=2, attention_head_dim=64, in_channels=16, out_channels=16, patch_size=2, num_layers=2,
               time_embed_dim=64, add_norm_text_encoder=True, text_embed_dim=128, text_embed_dim_t5=None)


def make_vae():
    for name, (boc, attn, zshape, seed) in VAE_CASES.items():
        ref = ref_vae.reference_decoder(cache_mag_vae=True, mid_block_use_attention=attn, block_out_channels=boc)
        mine = vae.init_weights_(vae.OracleDecoder(block_out_channels=boc, mid_block_use_attention=attn), seed)
        ref.load_state_dict(mine.state_dict(), strict=True)
        z = torch.randn(zshape, generator=torch.Generator().manual_seed(seed + 100))
        with torch.no_grad():
            out = ref(z)
        save_file({"z": z, "out": out.contiguous()}, os.path.join(HERE, f"{name}.safetensors"),
                  metadata={"source": "reference Decoder (cache_mag_vae=True, fp32, CPU)", "seed": str(seed),
                            "block_out_channels": str(boc), "mid_attention": str(attn)})
        print(name, tuple(out.shape), float(out.abs().max()))


def make_dit():
    m = dit.init_weights_(dit.OracleTransformer3D(**DIT_CFG), 1234)
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(1, 16, 3, 8, 12, generator=g)
    pe, ne = torch.randn(1, 9, 128, generator=g), torch.randn(1, 9, 128, generator=g)
    rope = dit.rope_for_video(64, 96, 3)
    out = dit.denoise_loop(m, lat, pe, ne, rope, num_steps=3, guidance_scale=6.0)
    sched = dit.FlowMatchEulerScheduler(shift=3.0)
    sched.set_timesteps(25, mu=1.0)
    cos720, sin720 = dit.rope_for_video(720, 1280, 2)
    save_file({"latents": lat, "prompt_embeds": pe, "negative_prompt_embeds": ne, "out_3steps_cfg6": out,
               "rope_cos": rope[0], "rope_sin": rope[1], "sigmas_shift3_25": sched.sigmas,
               "timesteps_shift3_25": sched.timesteps, "rope720_cos_head": cos720[:200].contiguous(),
               "rope720_sin_tail": sin720[-200:].contiguous()},
              os.path.join(HERE, "dit_tiny.safetensors"), metadata={"source": "oracle restatement (parity unpinned)"})
    print("dit_tiny", tuple(out.shape), float(out.std()))


if __name__ == "__main__":
    if not ref_vae.available():
        raise SystemExit("/root/reference not present: golden VAE vectors can only be minted in the authoring container")
    make_vae()
    make_dit()
