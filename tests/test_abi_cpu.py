"""The C-ABI shared library loads without a GPU and exports every symbol include/ea_b200.h declares; host-side
mirrors of the reference's module surface behave like the reference (no compute calls here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ea_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ea_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "easyanimate_b200", "libea_b200.so"))
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ea_b200.h but not exported"
    lib.ea_abi_version.restype = ctypes.c_int
    from easyanimate_b200 import _lib as L
    assert lib.ea_abi_version() == L.ABI_VERSION == 4
    lib.ea_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.ea_last_error(), bytes)


def test_argument_validation_without_gpu():
    from easyanimate_b200 import _lib as L
    args = L.GemmArgs()  # all-null
    assert L.ea_gemm(ctypes.byref(args), None) == -1
    assert b"null pointer" in L.ea_last_error()
    a = L.AttnArgs(q=1, k=1, v=1, B=1, H=1, S=8, head_dim=128)
    assert L.ea_attn_fwd(ctypes.byref(a), None) == -1 and b"head_dim" in L.ea_last_error()


def test_ops_refuse_cpu_tensors():
    from easyanimate_b200 import ops, _lib as L
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(L.EaError, match="no CPU path"):
        ops.gemm(x, x)


def test_transformer_config_surface_and_keys():
    from oracle import dit
    from easyanimate_b200.transformer3d import EasyAnimateTransformer3DModel
    cfg = dict(num_attention_heads=2, attention_head_dim=64, in_channels=16, out_channels=16, patch_size=2, num_layers=2,
               time_embed_dim=64, add_norm_text_encoder=True, text_embed_dim=128, text_embed_dim_t5=None)
    m = EasyAnimateTransformer3DModel(**cfg)
    assert m.config.in_channels == 16 and m.config.patch_size == 2 and m.config.attention_head_dim == 64
    assert m.config.get("time_position_encoding_type", "2d_rope") == "3d_rope"
    assert m.config.get("not_there", 7) == 7 and m.config.enable_text_attention_mask is True
    assert m.resize_inpaint_mask_directly is False and m.enable_clip_in_inpaint is True and m.teacache is None
    m.enable_teacache(25, 0.08)
    assert m.teacache.num_steps == 25 and abs(m.teacache.rescale_func(0.1) - float(__import__('numpy').poly1d(m.teacache.coefficients)(0.1))) < 1e-9
    assert set(m.state_dict().keys()) == set(dit.OracleTransformer3D(**cfg).state_dict().keys())
    assert m.to(torch.bfloat16).dtype == torch.bfloat16
    with pytest.raises(ValueError):
        EasyAnimateTransformer3DModel(**dict(cfg, attention_head_dim=128))


def test_vae_config_surface_and_keys():
    from oracle import vae
    from easyanimate_b200.autoencoder_magvit import AutoencoderKLMagvit
    kw = dict(latent_channels=16, cache_mag_vae=True, spatial_group_norm=True, mid_block_attention_type="spatial",
              mini_batch_encoder=4, mini_batch_decoder=1, scaling_factor=0.7125, block_out_channels=[64, 64, 128, 128],
              up_block_types='("SpatialUpBlock3D", "SpatialTemporalUpBlock3D", "SpatialTemporalUpBlock3D", "SpatialTemporalUpBlock3D",)')
    m = AutoencoderKLMagvit(**kw)
    assert m.config.scaling_factor == 0.7125 and m.config.latent_channels == 16
    assert m.config.block_out_channels == [64, 64, 128, 128]
    assert m.quant_conv.weight.ndim == 5 and m.cache_mag_vae and m.mini_batch_decoder == 1 and m.mini_batch_encoder == 4
    assert m.tile_latent_min_size == 48
    # every key of the reference's AutoencoderKLMagvit (244 for this architecture; the oracle is pinned to it key by key)
    ref = set(vae.OracleAutoencoderKLMagvit(block_out_channels=(64, 64, 128, 128), with_encoder=True).state_dict())
    assert set(m.state_dict()) == ref and len(ref) == 244
    with pytest.raises(NotImplementedError):
        AutoencoderKLMagvit(**dict(kw, cache_mag_vae=False))
    from easyanimate_b200 import _lib as L
    with pytest.raises(L.EaError, match="no CPU path"):
        m.to(torch.bfloat16).encode(torch.zeros(1, 3, 5, 16, 16))
    with pytest.raises(L.EaError, match="no CPU path"):
        m.decode(torch.zeros(1, 16, 2, 4, 4))


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every argument struct of include/ea_b200.h, compiled by gcc, has the size and the field offsets of its ctypes
    mirror in easyanimate_b200/_lib.py (same field names, same order): the binding cannot drift from the ABI silently."""
    import subprocess
    from easyanimate_b200 import _lib as L
    pairs = {"ea_gemm_args": L.GemmArgs, "ea_qkv_args": L.QkvArgs, "ea_skinny_linear_args": L.SkinnyArgs,
             "ea_ln_args": L.LnArgs, "ea_rmsnorm_args": L.RmsArgs, "ea_attn_args": L.AttnArgs, "ea_conv3d_args": L.ConvArgs,
             "ea_qkv_peers": L.QkvPeers, "ea_attn_peers": L.AttnPeers}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ea_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    # and the header declares no struct the binding does not mirror
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ea_b200.h")).read(), flags=re.S)
    assert sorted(re.findall(r"^\}\s*(ea_[a-z0-9_]+);", hdr, flags=re.M)) == sorted(pairs)


def test_import_fails_loudly_without_the_built_library(tmp_path):
    """No CPU / PyTorch fallback: a copy of the package WITHOUT libea_b200.so cannot even be imported."""
    import shutil
    import subprocess
    import sys
    shutil.copytree(os.path.join(ROOT, "easyanimate_b200"), str(tmp_path / "easyanimate_b200"),
                    ignore=shutil.ignore_patterns("*.so", "_build", "__pycache__", "csrc"))
    r = subprocess.run([sys.executable, "-c", "import easyanimate_b200"], cwd=str(tmp_path), capture_output=True, text=True,
                       env={**os.environ, "PYTHONPATH": str(tmp_path)})
    assert r.returncode != 0 and "ImportError" in r.stderr and "no CPU/PyTorch fallback" in r.stderr


def test_product_never_reaches_the_oracle_or_the_cpu_stand_ins():
    """oracle/ and tests/cpu_ops.py are test infrastructure: no module of the product package imports them, and bench.py touches the
    oracle only in its CPU-baseline / torch-baseline / reference legs (never inside the timed region of the product arm)."""
    import ast
    import glob
    for path in glob.glob(os.path.join(ROOT, "easyanimate_b200", "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                assert not (n == "oracle" or n.startswith("oracle.") or n.startswith("tests") or "cpu_ops" in n), (path, n)
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    allowed = {"cpu_oracle_rate", "reference_arm", "torch_gpu_baseline", "cpu_baseline"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, (ast.Import, ast.ImportFrom)) and any(
            (a.name if isinstance(n, ast.Import) else (n.module or "")).split(".")[0] == "oracle" for a in n.names) for n in ast.walk(fn))
        if uses:
            assert fn.name in allowed or "baseline" in fn.name or "reference" in fn.name or "oracle" in fn.name, fn.name
