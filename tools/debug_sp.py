"""Staged check of the peer-store sequence-parallel path (2+ GPUs, torchrun): IPC mapping, QKV peer stores, attention peer stores.
CUDA_LAUNCH_BLOCKING=1 is set so that a faulting kernel is reported where it is launched."""
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

bf16 = torch.bfloat16


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    grp = dist.new_group(list(range(world)))
    from easyanimate_b200 import ops, _lib as L
    from easyanimate_b200.sequence_parallel import PeerExchange

    def say(*a):
        print(f"[rank {rank}]", *a, flush=True)

    B, H, S_t, S_loc = 1, 8, 40, 192
    px = PeerExchange(grp, B, H, S_t, S_loc, dev)
    say("exchange built; q ptrs", [hex(px.qkv_video.q[r] or 0) for r in range(world)], "own", hex(px.q.data_ptr()))
    # (a) plain copy through the mapped pointers: every rank fills ITS out_video with its rank id, then reads the peers'
    px.out_video.fill_(float(rank + 1))
    torch.cuda.synchronize(); dist.barrier()
    # a kernel of THIS device storing into the peer's buffer: copy2d of 8 elements
    from easyanimate_b200 import _lib
    src = torch.full((1, 1, 1, 1, 8), float(10 + rank), device=dev, dtype=bf16)
    for r in range(world):
        if r != rank:
            _lib.check(_lib.ea_copy2d(src.data_ptr(), 8, 8, px.attn.out_video[r], 8, 8, 1, 1, 8, torch.cuda.current_stream().cuda_stream), "copy2d")
    torch.cuda.synchronize(); dist.barrier()
    say("after peers stored into my out_video:", px.out_video.view(-1)[:2].float().cpu().tolist())
    dist.barrier()
    # (b) QKV projection with peer stores
    d = H * 64
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn((B * S_loc, d), device=dev, generator=g).to(bf16)
    w = (torch.randn((3 * d, d), device=dev, generator=g) * d ** -0.5).to(bf16)
    b = torch.zeros((3 * d,), device=dev, dtype=bf16)
    ln = (torch.ones(64, device=dev, dtype=bf16), torch.zeros(64, device=dev, dtype=bf16))
    cos = torch.ones((S_loc, 64), device=dev); sin = torch.zeros((S_loc, 64), device=dev)
    ops.qkv_gemm_ln_rope(a, w, b, ln, ln, (cos, sin), px.q, px.k, px.v, rows_per_batch=S_loc, seq_offset=S_t + rank * S_loc,
                         peers=px.qkv_video)
    torch.cuda.synchronize(); say("qkv video peers ok")
    at = torch.randn((B * S_t, d), device=dev, generator=g).to(bf16)
    ops.qkv_gemm_ln_rope(at, w, b, ln, ln, None, px.q, px.k, px.v, rows_per_batch=S_t, seq_offset=0, peers=px.qkv_text)
    torch.cuda.synchronize(); say("qkv text peers ok")
    px.barrier(); torch.cuda.synchronize(); say("barrier ok")
    for variant in (0x10c, 0x210c):
        ops.attention(px.q, px.k, px.v, S_t, peers=px.attn, variant=variant)
        torch.cuda.synchronize(); say("attention peers ok variant", hex(variant))
        px.barrier()
    torch.cuda.synchronize()
    say("finite", bool(torch.isfinite(px.out_video.float()).all()), bool(torch.isfinite(px.out_text.float()).all()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
