"""easyanimate_b200.EasyAnimateTransformer3DModel against the oracle restatement of the reference (oracle/dit.py)."""
import pytest
import torch

from tests.parity import three_way

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _build(cfg, seed=1234, device="cuda"):
    from oracle import dit
    from easyanimate_b200.transformer3d import EasyAnimateTransformer3DModel
    o32 = dit.init_weights_(dit.OracleTransformer3D(**cfg), seed).to(torch.float32)  # (.to: the float64 ref_pos_embedding buffer)
    ob = dit.OracleTransformer3D(**cfg).to(bf16)
    ob.load_state_dict({k: v.to(bf16) for k, v in o32.state_dict().items()})
    o32.load_state_dict({k: v.float() for k, v in ob.state_dict().items()})  # truth uses the bf16-rounded weights
    ours = EasyAnimateTransformer3DModel(**cfg).to(bf16)
    missing, unexpected = ours.load_state_dict(ob.state_dict(), strict=True)
    assert not missing and not unexpected
    return o32, ob, ours.to(device)


def _inputs(B, C, F, H, W, S_t, E, seed=0):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(B, C, F, H, W, generator=g)
    enc = torch.randn(B, S_t, E, generator=g) * 3.0
    t = torch.tensor([937.0, 421.0][:B])
    return lat, enc, t


CFG_TINY = dict(num_attention_heads=2, attention_head_dim=64, in_channels=16, out_channels=16, patch_size=2, num_layers=2,
                time_embed_dim=64, add_norm_text_encoder=True, text_embed_dim=128, text_embed_dim_t5=None)
# BASELINE.json configs[0]: single DiT block, d=3072/48 heads, 1 frame 16x16 latent, 256 text tokens of width 3584
CFG_BLOCK = dict(num_attention_heads=48, attention_head_dim=64, in_channels=16, out_channels=16, patch_size=2, num_layers=1,
                 time_embed_dim=512, add_norm_text_encoder=True, text_embed_dim=3584, text_embed_dim_t5=None)


@pytest.mark.parametrize("name,cfg,shape", [
    ("tiny_2blocks", CFG_TINY, (2, 16, 3, 8, 12, 40)),
    ("tiny_ragged", CFG_TINY, (1, 16, 2, 6, 10, 7)),
    ("config1_single_block", CFG_BLOCK, (2, 16, 1, 16, 16, 256)),
])
def test_transformer_forward_matches_oracle(name, cfg, shape):
    from oracle import dit
    B, C, F, H, W, S_t = shape
    o32, ob, ours = _build(cfg)
    lat, enc, t = _inputs(B, C, F, H, W, S_t, cfg["text_embed_dim"])
    rope = dit.rope_for_video(H * 8, W * 8, F)
    with torch.no_grad():
        # the pipeline rounds the timestep to bf16 before the transformer call (pipeline_easyanimate.py:1079-1081)
        tb = t.to(bf16)
        truth = o32(lat.to(bf16).float(), tb.float(), encoder_hidden_states=enc.to(bf16).float(), image_rotary_emb=rope)[0]
        ref = ob(lat.to(bf16), tb, encoder_hidden_states=enc.to(bf16), image_rotary_emb=rope)[0]
        got = ours(lat.to(bf16).cuda(), tb.cuda(), encoder_hidden_states=enc.to(bf16).cuda(),
                   image_rotary_emb=(rope[0].cuda(), rope[1].cuda()), return_dict=False)[0]
    assert got.shape == (B, C, F, H, W) and got.dtype == bf16
    three_way(got, ref, truth, name=name)


def prelude_reference_golden(name, device="cuda"):
    """Fixture + oracle pair + product module of the reference-golden test (no GPU needed for device='cpu')."""
    import ast
    import os
    from safetensors import safe_open
    from safetensors.torch import load_file
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.safetensors")
    t = load_file(path)
    with safe_open(path, framework="pt") as f:
        meta = f.metadata()
    cfg = ast.literal_eval(meta["config"])
    shape = ast.literal_eval(meta["shape"])
    return t, cfg, shape, _build(cfg, seed=int(meta["seed"]), device=device)


@pytest.mark.parametrize("name", ["dit_ref_t2v", "dit_ref_i2v_inpaint", "dit_ref_3heads_3layers"])
def test_transformer_forward_matches_reference_golden(name):
    """Against outputs of the REFERENCE's own EasyAnimateTransformer3DModel (fp32, minted by tests/golden/make_golden.py
    in the authoring container): our bf16 kernels are as close to it as the bf16 oracle is."""
    from oracle import dit
    t, cfg, (B, F, H, W, St), (o32, ob, ours) = prelude_reference_golden(name)
    rope = dit.rope_for_video(H * 8, W * 8, F)
    inp = t.get("inpaint_latents")
    with torch.no_grad():
        tb = t["timestep"].to(bf16)
        truth32 = o32(t["latents"].to(bf16).float(), tb.float(), encoder_hidden_states=t["encoder_hidden_states"].to(bf16).float(),
                      image_rotary_emb=rope, inpaint_latents=None if inp is None else inp.to(bf16).float())[0]
        ref = ob(t["latents"].to(bf16), tb, encoder_hidden_states=t["encoder_hidden_states"].to(bf16), image_rotary_emb=rope,
                 inpaint_latents=None if inp is None else inp.to(bf16))[0]
        got = ours(t["latents"].to(bf16).cuda(), tb.cuda(), encoder_hidden_states=t["encoder_hidden_states"].to(bf16).cuda(),
                   image_rotary_emb=(rope[0].cuda(), rope[1].cuda()), inpaint_latents=None if inp is None else inp.to(bf16).cuda(),
                   return_dict=False)[0]
    # the fixture is the reference on UNROUNDED fp32 weights and inputs; the bf16 rounding of weights/inputs is common to
    # `ref` and `got`, so both are compared against the fixture itself
    three_way(got, ref, t["out"], name=name + "_vs_reference_fixture")
    three_way(got, ref, truth32, name=name)


def test_control_ref_and_clip_tokens_match_reference_golden():
    """v5.1 Control with a reference image (+ CLIP tokens): ref_proj patch tokens + resized 2-D sin-cos table in place of the
    text tokens (transformer3d.py:1420-1429,1538-1561), against outputs of the REFERENCE module (fixture minted by
    tests/golden/make_golden.py) and the oracle."""
    from oracle import dit
    t, cfg, (B, F, H, W, St), (o32, ob, ours) = prelude_reference_golden("dit_ref_control_ref_clip")
    rope = dit.rope_for_video(H * 8, W * 8, F)
    tb = t["timestep"].to(bf16)
    lat, enc, refl, clip = (t[k].to(bf16) for k in ("latents", "encoder_hidden_states", "ref_latents", "clip_encoder_hidden_states"))
    for key, clip_in in (("out", clip), ("out_ref_only", None)):
        with torch.no_grad():
            truth32 = o32(lat.float(), tb.float(), encoder_hidden_states=enc.float(), image_rotary_emb=rope, ref_latents=refl.float(),
                          clip_encoder_hidden_states=None if clip_in is None else clip_in.float())[0]
            ref = ob(lat, tb, encoder_hidden_states=enc, image_rotary_emb=rope, ref_latents=refl, clip_encoder_hidden_states=clip_in)[0]
            got = ours(lat.cuda(), tb.cuda(), encoder_hidden_states=enc.cuda(), image_rotary_emb=(rope[0].cuda(), rope[1].cuda()),
                       ref_latents=refl.cuda(), clip_encoder_hidden_states=None if clip_in is None else clip_in.cuda(),
                       return_dict=False)[0]
        three_way(got, ref, t[key], name=f"control_ref_clip[{key}]_vs_reference_fixture")
        three_way(got, ref, truth32, name=f"control_ref_clip[{key}]")


def test_transformer_forward_is_deterministic():
    """No atomics on the forward path: two runs agree bit for bit (what CFG-parallel == batch-of-2 relies on)."""
    from oracle import dit
    _, _, ours = _build(CFG_TINY)
    lat, enc, t = _inputs(2, 16, 3, 8, 12, 40, CFG_TINY["text_embed_dim"])
    rope = dit.rope_for_video(8 * 8, 12 * 8, 3)
    args = (lat.to(bf16).cuda(), t.to(bf16).cuda())
    kw = dict(encoder_hidden_states=enc.to(bf16).cuda(), image_rotary_emb=(rope[0].cuda(), rope[1].cuda()), return_dict=False)
    with torch.no_grad():
        a = ours(*args, **kw)[0]
        b = ours(*args, **kw)[0]
    assert torch.equal(a, b)


def test_transformer_i2v_inpaint_channels():
    """predict_i2v path: inpaint_latents (1 mask + 16 masked-video channels) concatenated on channels -> in_channels 33."""
    from oracle import dit
    cfg = dict(CFG_TINY, in_channels=33)
    o32, ob, ours = _build(cfg)
    B, F, H, W, S_t = 2, 2, 8, 8, 16
    lat, enc, t = _inputs(B, 16, F, H, W, S_t, cfg["text_embed_dim"])
    inp = torch.randn(B, 17, F, H, W, generator=torch.Generator().manual_seed(5))
    rope = dit.rope_for_video(H * 8, W * 8, F)
    tb = t.to(bf16)
    with torch.no_grad():
        truth = o32(lat.to(bf16).float(), tb.float(), encoder_hidden_states=enc.to(bf16).float(), image_rotary_emb=rope,
                    inpaint_latents=inp.to(bf16).float())[0]
        ref = ob(lat.to(bf16), tb, encoder_hidden_states=enc.to(bf16), image_rotary_emb=rope, inpaint_latents=inp.to(bf16))[0]
        got = ours(lat.to(bf16).cuda(), tb.cuda(), encoder_hidden_states=enc.to(bf16).cuda(),
                   image_rotary_emb=(rope[0].cuda(), rope[1].cuda()), inpaint_latents=inp.to(bf16).cuda())
    assert got.sample.shape == (B, 16, F, H, W)
    three_way(got.sample, ref, truth, name="i2v_inpaint")


def test_config_surface_and_state_dict_keys():
    from oracle import dit
    from easyanimate_b200.transformer3d import EasyAnimateTransformer3DModel
    m = EasyAnimateTransformer3DModel(**CFG_TINY)
    assert m.config.in_channels == 16 and m.config.patch_size == 2 and m.config.attention_head_dim == 64
    assert m.config.get("time_position_encoding_type", "2d_rope") == "3d_rope"
    assert m.config.get("not_there", 7) == 7 and m.config.enable_text_attention_mask is True
    assert set(m.state_dict().keys()) == set(dit.OracleTransformer3D(**CFG_TINY).state_dict().keys())


def test_teacache_matches_oracle_decisions_and_outputs():
    """TeaCache (transformer3d.py:90-121,1563-1636): same skip decisions and outputs as the oracle over a 6-step loop."""
    from oracle import dit
    o32, ob, ours = _build(CFG_TINY)
    coeffs = [1.07862322, -4.19362456, 3.06725828, 0.33161686, 0.02374758]  # get_teacache_coefficients("v5.1-7b")
    steps, thresh = 6, 0.08
    ob.teacache = dit.OracleTeaCache(coeffs, steps, thresh)
    ours.enable_teacache(steps, thresh, coefficients=coeffs)
    B, C, F, H, W, S_t = 2, 16, 2, 8, 8, 16
    lat, enc, _ = _inputs(B, C, F, H, W, S_t, CFG_TINY["text_embed_dim"])
    rope = dit.rope_for_video(H * 8, W * 8, F)
    sched = dit.FlowMatchEulerScheduler()
    sched.set_timesteps(steps)
    x_ref, x_our = lat.to(bf16), lat.to(bf16).cuda()
    with torch.no_grad():
        for t in sched.timesteps:
            tb = torch.tensor([t, t]).to(bf16)
            r = ob(x_ref, tb, encoder_hidden_states=enc.to(bf16), image_rotary_emb=rope)[0]
            g = ours(x_our, tb.cuda(), encoder_hidden_states=enc.to(bf16).cuda(),
                     image_rotary_emb=(rope[0].cuda(), rope[1].cuda()), return_dict=False)[0]
            torch.testing.assert_close(g.float().cpu(), r.float(), rtol=0.05, atol=0.05)
            x_ref, x_our = (x_ref - 0.1 * r).to(bf16), (x_our - 0.1 * g).to(bf16)
    assert ours.teacache.skipped == ob.teacache.skipped and ours.teacache.skipped >= 1, (ours.teacache.skipped, ob.teacache.skipped)
    assert ours.teacache.cnt == 0  # reset after num_steps forwards, like the reference


def test_fp8_weight_storage_mode_matches_reference_arithmetic():
    """Row f4: parameters stored as float8_e4m3fn (predict_t2v.py:37,106), expanded block by block with ea_dequant_e4m3:
    (1) the expansion is exact (every e4m3 value is a bf16 value); (2) the forward equals the bf16 forward of a model whose
    weights are bf16(e4m3(w)) BIT FOR BIT (same kernels, same operands); (3) it is as close to the fp32 oracle with those
    weights as the bf16 oracle is."""
    from oracle import dit
    from easyanimate_b200 import ops
    from easyanimate_b200.transformer3d import EasyAnimateTransformer3DModel
    w8 = torch.arange(256, dtype=torch.uint8, device="cuda").repeat(9).view(torch.float8_e4m3fn)
    got = ops.dequant_e4m3(w8, torch.empty(w8.numel(), device="cuda", dtype=bf16))
    want = w8.to(bf16)
    assert torch.equal(torch.nan_to_num(got.float(), nan=7.0), torch.nan_to_num(want.float(), nan=7.0))
    o32, ob, _ = _build(CFG_TINY, device="cpu")
    q = {k: v.to(torch.float8_e4m3fn).to(bf16) for k, v in ob.state_dict().items()}
    ob.load_state_dict(q)
    o32.load_state_dict({k: v.float() for k, v in q.items()})
    plain = EasyAnimateTransformer3DModel(**CFG_TINY).to(bf16)
    plain.load_state_dict(q)
    plain = plain.cuda()
    stored = EasyAnimateTransformer3DModel(**CFG_TINY).to(bf16)
    stored.load_state_dict(q)
    stored = stored.to(torch.float8_e4m3fn).cuda()
    assert stored.dtype == torch.float8_e4m3fn
    B, C, F, H, W, S_t = 2, 16, 3, 8, 12, 40
    lat, enc, t = _inputs(B, C, F, H, W, S_t, CFG_TINY["text_embed_dim"])
    rope = dit.rope_for_video(H * 8, W * 8, F)
    tb = t.to(bf16)
    kw = dict(encoder_hidden_states=enc.to(bf16).cuda(), image_rotary_emb=(rope[0].cuda(), rope[1].cuda()), return_dict=False)
    with torch.no_grad():
        truth = o32(lat.to(bf16).float(), tb.float(), encoder_hidden_states=enc.to(bf16).float(), image_rotary_emb=rope)[0]
        ref = ob(lat.to(bf16), tb, encoder_hidden_states=enc.to(bf16), image_rotary_emb=rope)[0]
        a = plain(lat.to(bf16).cuda(), tb.cuda(), **kw)[0]
        b = stored(lat.to(bf16).cuda(), tb.cuda(), **kw)[0]
    assert torch.equal(a, b)
    three_way(b, ref, truth, name="fp8_weight_storage")


def test_control_latents_channel_concat():
    """v5.1 Control: hidden | inpaint | control channel concat (transformer3d.py:1523-1526) feeding the patch-embed GEMM."""
    from oracle import dit
    cfg = dict(CFG_TINY, in_channels=16 + 17 + 16)
    o32, ob, ours = _build(cfg)
    B, F, H, W, S_t = 2, 2, 8, 8, 16
    lat, enc, t = _inputs(B, 16, F, H, W, S_t, cfg["text_embed_dim"])
    g = torch.Generator().manual_seed(5)
    inp, ctl = torch.randn(B, 17, F, H, W, generator=g).to(bf16), torch.randn(B, 16, F, H, W, generator=g).to(bf16)
    rope = dit.rope_for_video(H * 8, W * 8, F)
    tb = t.to(bf16)
    cat = torch.cat([inp, ctl], 1)
    with torch.no_grad():
        truth = o32(lat.to(bf16).float(), tb.float(), encoder_hidden_states=enc.to(bf16).float(), image_rotary_emb=rope,
                    inpaint_latents=cat.float())[0]
        ref = ob(lat.to(bf16), tb, encoder_hidden_states=enc.to(bf16), image_rotary_emb=rope, inpaint_latents=cat)[0]
        got = ours(lat.to(bf16).cuda(), tb.cuda(), encoder_hidden_states=enc.to(bf16).cuda(),
                   image_rotary_emb=(rope[0].cuda(), rope[1].cuda()), inpaint_latents=inp.cuda(), control_latents=ctl.cuda(),
                   return_dict=False)[0]
    three_way(got, ref, truth, name="control_latents")


def test_transformer_forward_is_cuda_graph_capturable():
    """include/ea_b200.h promises enqueue-only entry points (no allocation, no synchronisation, no global state): the whole
    MMDiT forward - every kernel launch of the blocks plus torch's allocations from the graph's private pool - is captured ONCE
    and replayed on new contents of the static input buffers, and equals the eager forward bit for bit.  (What a deployment uses
    for the launch-bound small configurations, BASELINE configs[0]: 14 launches per block at a few microseconds each.)"""
    from oracle import dit
    _, _, ours = _build(CFG_TINY)
    lat, enc, t = _inputs(2, 16, 3, 8, 12, 40, 128)
    rope = tuple(r.cuda() for r in dit.rope_for_video(64, 96, 3))
    s_lat, s_enc, s_t = lat.to(bf16).cuda(), enc.to(bf16).cuda(), t.to(bf16).cuda()

    def fwd():
        return ours(s_lat, s_t, encoder_hidden_states=s_enc, image_rotary_emb=rope, return_dict=False)[0]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):  # warm-up off the capture: per-device launch attributes, packed / fused weight caches
            fwd()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(graph):
        out = fwd()
    lat2, enc2, _ = _inputs(2, 16, 3, 8, 12, 40, 128, seed=5)
    s_lat.copy_(lat2.to(bf16))
    s_enc.copy_(enc2.to(bf16))
    s_t.copy_(torch.tensor([611.0, 87.0]).to(bf16))
    graph.replay()
    torch.cuda.synchronize()
    got = out.clone()
    with torch.no_grad():
        want = fwd()
    assert torch.isfinite(got).all() and torch.equal(got, want)
