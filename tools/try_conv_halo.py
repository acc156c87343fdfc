"""Bring-up / A-B of the halo-tile convolution (csrc/conv3d_halo.cu) against the tap-per-box kernel (csrc/conv3d_tc.cu, validated)
and an fp32 F.conv3d reference.  Every variant runs in its own process under a timeout (a wrong descriptor can trap the context).
usage: python tools/try_conv_halo.py            (driver: all variants)
       python tools/try_conv_halo.py one <variant-hex> [bench]"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHECK_SHAPES = [(1, 8, 16, 64, 64), (3, 10, 20, 64, 128), (2, 9, 7, 128, 256), (4, 16, 32, 256, 128), (2, 24, 40, 512, 512),
                (5, 128, 160, 64, 128), (5, 122, 150, 64, 128), (3, 90, 160, 128, 256), (3, 208, 256, 128, 256),
                (3, 400, 250, 64, 128), (2, 33, 19, 64, 32)]
BENCH_SHAPES = [(13, 90, 160, 512, 512), (25, 360, 640, 256, 256), (13, 720, 1280, 128, 128), (25, 180, 320, 512, 512)]


def one(variant: int, bench: bool):
    import torch
    import torch.nn.functional as F
    os.environ["EA_CONV_VARIANT"] = hex(variant)
    from easyanimate_b200 import vae_ops
    from tools.bench_kernels import timeit
    bf16 = torch.bfloat16
    assert vae_ops.CONV_VARIANT == variant

    def rnd(shape, scale, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        return (torch.randn(shape, device="cuda", generator=g) * scale).to(bf16)

    worst = 0.0
    for (T, H, W, Cin, Cout) in CHECK_SHAPES:
        x, w, b = rnd((T, H, W, Cin), 1.0, 1), rnd((Cout, Cin, 3, 3, 3), (27 * Cin) ** -0.5, 2), rnd((Cout,), 0.1, 3)
        wp = vae_ops.pack_conv_weight(w, cout_pad=max(Cout, 32))
        out = vae_ops.conv3d_causal(x, wp, b, Cout)
        xx = F.pad(x.permute(3, 0, 1, 2)[None].float(), (0, 0, 0, 0, 2, 0), mode="replicate")
        ref = F.conv3d(xx, w.float(), b.float(), padding=(0, 1, 1))[0].permute(1, 2, 3, 0)
        err = (out.float() - ref).abs().max().item()
        res = rnd((T, H, W, Cout), 1.0, 4)
        out2 = vae_ops.conv3d_causal(x, wp, b, Cout, residual=res)
        err2 = (out2.float() - (ref.to(bf16) + res).float()).abs().max().item()
        ok3 = True
        if T > 1:
            out3 = vae_ops.conv3d_causal(x, wp, b, Cout, dup_frames=True)
            idx = [0] + [i for t in range(1, T) for i in (t, t)]
            ok3 = bool(torch.equal(out3, out[idx]))
        worst = max(worst, err, err2)
        print(json.dumps({"variant": hex(variant), "shape": [T, H, W, Cin, Cout], "max_err": round(err, 4), "max_err_res": round(err2, 4),
                          "dup_ok": ok3}), flush=True)
    print(json.dumps({"variant": hex(variant), "worst": worst, "PASS": worst < 0.05}), flush=True)
    if bench and worst < 0.05:
        for (T, H, W, Cin, Cout) in BENCH_SHAPES:
            x, w, b = rnd((T, H, W, Cin), 1.0, 1), rnd((Cout, Cin, 3, 3, 3), (27 * Cin) ** -0.5, 2), rnd((Cout,), 0.1, 3)
            wp = vae_ops.pack_conv_weight(w)
            del w
            med, best = timeit(lambda: vae_ops.conv3d_causal(x, wp, b, Cout), reps=5, warmup=2, flush=False)
            fl = 2.0 * T * H * W * Cout * 27 * Cin
            print(json.dumps({"variant": hex(variant), "bench": [T, H, W, Cin, Cout], "ms": round(med, 3), "tflops": round(fl / med / 1e9, 1),
                              "best_tflops": round(fl / best / 1e9, 1)}), flush=True)
            del x, wp
            torch.cuda.empty_cache()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(int(sys.argv[2], 0), len(sys.argv) > 3)
    else:
        for v in (0x4, 0x0, 0x10):  # tap-per-box kernel, halo kernel (default pitch), halo kernel (the other pitch)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", hex(v), "bench"], timeout=240, capture_output=True, text=True)
                print(r.stdout[-6000:], flush=True)
                if r.returncode != 0:
                    print(json.dumps({"variant": hex(v), "rc": r.returncode, "stderr": r.stderr[-800:]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"variant": hex(v), "timeout": True}), flush=True)
