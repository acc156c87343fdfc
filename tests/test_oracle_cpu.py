"""CPU suite: the oracle against the golden vectors minted from the reference, against the reference itself when
/root/reference is present, and internal consistency of the restated diffusers pieces."""
import ast
import os

import pytest
import torch
from safetensors import safe_open
from safetensors.torch import load_file

from oracle import dit, ref_vae, vae

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _meta(path):
    with safe_open(path, framework="pt") as f:
        return f.metadata()


@pytest.mark.parametrize("name", ["vae_small_attn", "vae_small_noattn_ragged", "vae_full_arch"])
def test_vae_oracle_matches_reference_golden(name):
    path = os.path.join(GOLD, f"{name}.safetensors")
    t, meta = load_file(path), _meta(path)
    boc = ast.literal_eval(meta["block_out_channels"])
    dec = vae.init_weights_(vae.OracleDecoder(block_out_channels=boc, mid_block_use_attention=meta["mid_attention"] == "True"),
                            int(meta["seed"]))
    with torch.no_grad():
        out = dec(t["z"])
    assert out.shape == t["out"].shape
    torch.testing.assert_close(out, t["out"], rtol=1e-4, atol=1e-4)  # whole-sequence vs chunked fp32 round-off


@pytest.mark.skipif(not ref_vae.available(), reason="/root/reference only exists in the authoring container")
@pytest.mark.parametrize("cache", [True, False])
def test_vae_oracle_matches_live_reference(cache):
    boc = (64, 64, 128, 128)
    ref = ref_vae.reference_decoder(cache_mag_vae=cache, block_out_channels=boc)
    mine = vae.init_weights_(vae.OracleDecoder(block_out_channels=boc), 3)
    ref.load_state_dict(mine.state_dict(), strict=True)
    z = torch.randn(1, 16, 3, 6, 8, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        torch.testing.assert_close(mine(z), ref(z), rtol=1e-4, atol=1e-4)


def test_vae_tiled_decode_oracle_shapes_and_seams():
    m = vae.init_weights_(vae.OracleAutoencoderKLMagvit(block_out_channels=(64, 64, 128, 128), use_tiling=True,
                                                        tile_sample_min_size=64, mid_block_use_attention=False), 9)
    z = torch.randn(1, 16, 2, 12, 20, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        tiled = m.decode(z)[0]
        m.use_tiling = False
        full = m.decode(z)[0]
    assert tiled.shape == full.shape == (1, 3, 5, 96, 160)
    assert torch.isfinite(tiled).all()


def test_dit_golden_regression():
    t = load_file(os.path.join(GOLD, "dit_tiny.safetensors"))
    from tests.golden.make_golden import DIT_CFG
    m = dit.init_weights_(dit.OracleTransformer3D(**DIT_CFG), 1234)
    rope = dit.rope_for_video(64, 96, 3)
    torch.testing.assert_close(rope[0], t["rope_cos"])
    torch.testing.assert_close(rope[1], t["rope_sin"])
    out = dit.denoise_loop(m, t["latents"], t["prompt_embeds"], t["negative_prompt_embeds"], rope, 3, 6.0)
    torch.testing.assert_close(out, t["out_3steps_cfg6"], rtol=1e-4, atol=1e-4)
    s = dit.FlowMatchEulerScheduler(shift=3.0)
    s.set_timesteps(25, mu=1.0)
    torch.testing.assert_close(s.sigmas, t["sigmas_shift3_25"])
    c, sn = dit.rope_for_video(720, 1280, 2)
    torch.testing.assert_close(c[:200], t["rope720_cos_head"])
    torch.testing.assert_close(sn[-200:], t["rope720_sin_tail"])


def test_rope_properties():
    cos, sin = dit.rope_for_video(720, 1280, 13)
    assert cos.shape == (13 * 45 * 80, 64) and cos.dtype == torch.float32
    torch.testing.assert_close(cos * cos + sin * sin, torch.ones_like(cos))
    # pairs are repeat-interleaved, and the temporal band (first 16 dims) is constant inside a frame
    assert torch.equal(cos[:, 0::2], cos[:, 1::2])
    assert torch.equal(cos[:45 * 80, :16], cos[0:1, :16].expand(45 * 80, 16))
    # rotating by the table is norm-preserving
    x = torch.randn(1, 2, 13 * 45 * 80, 64)
    y = dit.apply_rotary_emb(x, (cos, sin))
    torch.testing.assert_close(y.norm(dim=-1), x.norm(dim=-1), rtol=1e-4, atol=1e-4)


def test_scheduler_properties():
    for shift, dyn in ((1.0, False), (3.0, False), (1.0, True)):
        s = dit.FlowMatchEulerScheduler(shift=shift, use_dynamic_shifting=dyn)
        for n in (25, 30, 50):
            s.set_timesteps(n, mu=1.0)
            assert s.sigmas.shape == (n + 1,) and s.sigmas[-1] == 0
            assert torch.all(s.sigmas[:-1] > s.sigmas[1:])
            torch.testing.assert_close(s.timesteps, s.sigmas[:-1] * 1000)
    # Euler steps telescope: with a constant velocity v the loop integrates x0 + (0 - sigma_0) * v
    s = dit.FlowMatchEulerScheduler()
    s.set_timesteps(10)
    x, v = torch.zeros(4), torch.ones(4)
    for _ in range(10):
        x = s.step(v, x)
    torch.testing.assert_close(x, -s.sigmas[0] * torch.ones(4))


def test_cfg_linearity():
    """CFG with guidance g is linear in (u, c): g=1 returns c, g=0 returns u."""
    u, c = torch.randn(5), torch.randn(5)
    assert torch.allclose(u + 1.0 * (c - u), c)
    assert torch.allclose(u + 0.0 * (c - u), u)
