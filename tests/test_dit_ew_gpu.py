"""HBM-bound DiT kernels against PyTorch references executed op by op in bf16 like the reference modules."""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(bf16)


@pytest.mark.parametrize("d", [3072, 1920, 512, 64])
def test_layernorm_modulate(d):
    from easyanimate_b200 import ops
    B, S = 2, 77
    x = _rand((B * S, d), 2.0, 1)
    w, b = 1 + _rand((d,), 0.1, 2), _rand((d,), 0.1, 3)
    mod = _rand((B, 6 * d), 0.5, 4)
    shift, scale = mod[:, 0:d], mod[:, d:2 * d]
    out = ops.layernorm_modulate(x, w, b, 1e-5, shift=shift, scale=scale, rows_per_batch=S)
    n = torch.nn.functional.layer_norm(x.float(), (d,), w.float(), b.float(), 1e-5).to(bf16).view(B, S, d)
    ref = n * (1 + scale)[:, None, :] + shift[:, None, :]
    torch.testing.assert_close(out.view(B, S, d).float(), ref.float(), rtol=2 ** -7, atol=2e-2)
    # plain LN and double LN (norm_final -> norm_out)
    out2 = ops.layernorm_modulate(x, w, b, 1e-5)
    torch.testing.assert_close(out2.float(), n.view(B * S, d).float(), rtol=2 ** -7, atol=1e-2)
    w2, b2 = 1 + _rand((d,), 0.1, 5), _rand((d,), 0.1, 6)
    out3 = ops.layernorm_modulate(x, w2, b2, 1e-5, shift=shift, scale=scale, rows_per_batch=S, pre=(w, b, 1e-5))
    n2 = torch.nn.functional.layer_norm(n, (d,), w2, b2, 1e-5)
    ref3 = n2 * (1 + scale)[:, None, :] + shift[:, None, :]
    torch.testing.assert_close(out3.view(B, S, d).float(), ref3.float(), rtol=2 ** -6, atol=4e-2)


def test_rmsnorm():
    from easyanimate_b200 import ops
    x = _rand((512, 3584), 10.0, 1)
    w = 1 + _rand((3584,), 0.1, 2)
    out = ops.rmsnorm(x, w, 1e-6)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(bf16)
    torch.testing.assert_close(out.float(), ref.float(), rtol=2 ** -7, atol=1e-2)


def test_timestep_embedding_and_skinny_linear():
    import math
    from easyanimate_b200 import ops
    t = torch.tensor([999.0, 500.0, 3.0], device="cuda").to(bf16)
    dim = 3072
    out = ops.timestep_embedding(t, dim)
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device="cuda") / half
    emb = t.float()[:, None] * torch.exp(exponent)[None]
    ref = torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1).to(bf16)
    # arguments reach ~1e3 rad: fp32 sin/cos implementations differ by a few 1e-4
    torch.testing.assert_close(out.float(), ref.float(), rtol=0, atol=1e-2)
    w, b = _rand((512, dim), 0.02, 1), _rand((512,), 0.1, 2)
    y = ops.skinny_linear(out, w, b)
    ref_y = (out.float() @ w.float().t() + b.float()).to(bf16)
    torch.testing.assert_close(y.float(), ref_y.float(), rtol=2 ** -7, atol=1e-2)
    w2, b2 = _rand((18432, 512), 0.05, 3), _rand((18432,), 0.1, 4)
    y2 = ops.skinny_linear(y, w2, b2, act_in=1)
    ref_y2 = (torch.nn.functional.silu(y).float() @ w2.float().t() + b2.float()).to(bf16)
    torch.testing.assert_close(y2.float(), ref_y2.float(), rtol=2 ** -7, atol=1e-2)


@pytest.mark.parametrize("C2", [0, 17])
def test_patchify_unpatchify(C2):
    from easyanimate_b200 import ops
    B, C, F, H, W = 2, 16, 3, 12, 20
    x = _rand((B, C, F, H, W), 1.0, 1)
    x2 = _rand((B, C2, F, H, W), 1.0, 2) if C2 else None
    a = ops.patchify(x, x2)
    xx = torch.cat([x, x2], 1) if C2 else x
    CC = C + C2
    ref = xx.view(B, CC, F, H // 2, 2, W // 2, 2).permute(0, 2, 3, 5, 1, 4, 6).reshape(B * F * (H // 2) * (W // 2), CC * 4)
    assert torch.equal(a[:, :CC * 4], ref)
    assert torch.count_nonzero(a[:, CC * 4:]) == 0
    y = _rand((B * F * (H // 2) * (W // 2), 64), 1.0, 3)
    out = ops.unpatchify(y, B, 16, F, H, W)
    ref_o = y.reshape(B, F, H // 2, W // 2, 16, 2, 2).permute(0, 4, 1, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    assert torch.equal(out, ref_o)


def test_cfg_euler_step():
    from easyanimate_b200 import ops
    lat = _rand((1, 16, 3, 8, 10), 1.0, 1)
    pred = _rand((2, 16, 3, 8, 10), 1.0, 2)
    sig, sig_next = torch.tensor(0.9371), torch.tensor(0.9012)
    out = ops.cfg_euler_step(pred, lat, 6.0, sig.item(), sig_next.item())
    u, c = pred.chunk(2)
    v = u + 6.0 * (c - u)
    ref = (lat.float() + (sig_next - sig).cuda() * v).to(bf16)
    assert torch.equal(out, ref)
    out1 = ops.cfg_euler_step(pred[:1], lat, 1.0, sig.item(), sig_next.item(), use_cfg=False)
    ref1 = (lat.float() + (sig_next - sig).cuda() * pred[:1]).to(bf16)
    assert torch.equal(out1, ref1)


def test_ipc_export_names_the_allocation_that_holds_a_buffer():
    """ea_ipc_export (sequence-parallel peer buffers): a sub-allocation of torch's caching allocator is exported as the
    cudaIpcMemHandle_t of its cudaMalloc'ed segment + its byte offset; two buffers carved from one segment share the handle."""
    import ctypes as C
    from easyanimate_b200 import _lib as L
    a = torch.empty(1 << 16, device="cuda", dtype=torch.uint8)   # both come from the same 2 MB small-block segment
    b = torch.empty(1 << 16, device="cuda", dtype=torch.uint8)
    out = []
    for t in (a, b):
        h, off = (C.c_char * 64)(), L.i64(0)
        L.check(L.ea_ipc_export(t.data_ptr(), h, C.byref(off)), "ea_ipc_export")
        out.append((bytes(h.raw), off.value))
        assert off.value >= 0 and any(h.raw)
    if out[0][0] == out[1][0]:  # same segment: the offsets differ by the distance of the buffers
        assert out[1][1] - out[0][1] == b.data_ptr() - a.data_ptr()
    assert L.ea_ipc_export(None, (C.c_char * 64)(), C.byref(L.i64(0))) != 0
