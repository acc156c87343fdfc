"""Every `-m gpu` module test starts with a part that needs no GPU: read the fixture, build the oracle pair and the product
module, load the state dict and assert which keys may be missing.  Round 1's driver run went red inside exactly such a
prelude (an assertion that had not followed a new module), on the GPU box, where the author could no longer see it.  This
file runs all of those preludes on the authoring box, so that class of failure shows up in the CPU suite."""
import pytest
import torch

bf16 = torch.bfloat16


@pytest.mark.parametrize("name", ["vae_small_attn", "vae_small_noattn_ragged", "vae_full_arch"])
def test_prelude_vae_decode_golden(name):
    from tests import test_vae_gpu as T
    assert name in T.GOLDEN_DECODE_CASES
    t, o32, ob, ours = T.prelude_decode_golden(name)
    assert "z" in t and "out" in t
    assert all(p.dtype == bf16 for p in ours.parameters())


def test_prelude_vae_tiled():
    from tests import test_vae_gpu as T
    T.prelude_tiled_oracle()
    t, ob, ours = T.prelude_tiled_reference_golden()
    assert "z" in t and "out_tiled" in t


def test_prelude_vae_encode_golden():
    from tests import test_zz_vae_encode_gpu as T
    t, ob, ours = T.prelude_encode_golden()
    assert {"x", "moments", "moments_tiled32"} <= set(t)


@pytest.mark.parametrize("cfg_name", ["CFG_TINY", "CFG_BLOCK"])
def test_prelude_dit_build(cfg_name):
    from tests import test_dit_module_gpu as T
    o32, ob, ours = T._build(getattr(T, cfg_name), device="cpu")
    assert set(ours.state_dict()) == set(ob.state_dict())


@pytest.mark.parametrize("name", ["dit_ref_t2v", "dit_ref_i2v_inpaint", "dit_ref_3heads_3layers", "dit_ref_control_ref_clip"])
def test_prelude_dit_reference_golden(name):
    from tests import test_dit_module_gpu as T
    t, cfg, shape, (o32, ob, ours) = T.prelude_reference_golden(name, device="cpu")
    assert {"latents", "timestep", "encoder_hidden_states", "out"} <= set(t)


def test_prelude_sampler_and_pipeline_tests():
    from tests import test_pipeline_gpu as T
    T.prelude_sampler(device="cpu")
    T.prelude_decode_latents(device="cpu")
    T.prelude_from_pretrained_2d(device="cpu")
    T.prelude_reference_pipeline_golden(device="cpu")


def test_gpu_test_modules_import_without_a_gpu():
    """Collection-time errors in a GPU test file are only visible on the GPU box otherwise."""
    import importlib
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    for f in sorted(os.listdir(here)):
        if f.endswith("_gpu.py"):
            importlib.import_module("tests." + f[:-3])
