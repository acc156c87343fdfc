"""Ulysses sequence parallelism for ONE video across the ranks of a group.  The reference has no multi-GPU inference code at
this commit (no xfuser / ulysses / ring); joint text+video attention (processor.py:287-289) makes a
per-block exchange unavoidable once one video spans more than the two CFG branches, and this module
is this framework's own design for it: every rank owns a contiguous slice of the video tokens for the
per-token work (AdaLN, projections, feed-forward) and, inside attention, all tokens of a slice of the HEADS.  The text
tokens are few (256) and replicated.  Two exchanges per block, both plain NCCL collectives here:

    q/k/v  [B, H, S_t + S_loc, 64]  --all_to_all-->  [B, H/P, S_t + S_v, 64]      (before attention)
    out    [B, S_v, (H/P)*64]       --all_to_all-->  [B, S_loc, H*64]  (+ all_gather of the text rows)

STATUS: the exchange logic is covered by a world_size-2 gloo test against single-process attention
(tests/test_dist_sp_cpu.py, which also runs the whole forward, a TeaCache sequence and the 2 CFG branches x 2 ranks sampler
topology with CPU stand-ins for the kernels); the model integration (`EasyAnimateTransformer3DModel.set_sequence_parallel_group`) has NOT run
on GPUs yet - tools/test_multigpu.py checks it against the single-GPU forward and is the first thing to run next round.
The B200-native form of these exchanges (P2P stores from the QKV-GEMM / attention epilogues into the peers' buffers) is
DESIGN.md section 8, item 3; this NCCL version is the baseline it will be measured against.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def group_size(group) -> int:
    return dist.get_world_size(group)


def all_reduce_floats(values, group) -> list:
    """Sum a short list of Python floats over `group` (TeaCache's rel-L1 pieces); fp64 on the wire."""
    t = torch.tensor(list(values), dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    return [float(v) for v in t.cpu()]


class UlyssesAttention:
    def __init__(self, group, attention_fn: Optional[Callable] = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._attention = attention_fn  # (q, k, v, S_t) -> (out_text [B,S_t,h*64], out_video [B,S-S_t,h*64]); default ops.attention

    # ---- token sharding of the per-token streams ----------------------------------------------------------------
    def local_range(self, S_v: int) -> Tuple[int, int]:
        if S_v % self.world:
            raise ValueError(f"sequence parallelism needs the video token count ({S_v}) to divide by the group size ({self.world})")
        n = S_v // self.world
        return self.rank * n, (self.rank + 1) * n

    def shard_tokens(self, x: torch.Tensor, B: int, S_v: int) -> torch.Tensor:
        """x [B*S_v, c] (token-major per batch element) -> this rank's [B*S_loc, c]."""
        s0, s1 = self.local_range(S_v)
        return x.view(B, S_v, -1)[:, s0:s1].reshape(B * (s1 - s0), -1).contiguous()

    def gather_tokens(self, x_loc: torch.Tensor, B: int, S_loc: int) -> torch.Tensor:
        """[B*S_loc, c] of every rank -> [B*S_v, c] on every rank (rank order = token order)."""
        c = x_loc.shape[1]
        out = torch.empty((self.world * B * S_loc, c), device=x_loc.device, dtype=x_loc.dtype)
        dist.all_gather_into_tensor(out, x_loc.contiguous(), group=self.group)
        return out.view(self.world, B, S_loc, c).permute(1, 0, 2, 3).reshape(B * self.world * S_loc, c).contiguous()

    def all_reduce_sums(self, values) -> list:
        return all_reduce_floats(values, self.group)

    # ---- attention with the two exchanges -----------------------------------------------------------------------
    def attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, S_t: int):
        """q, k, v [B, H, S_t + S_loc, 64]: ALL heads, the replicated text rows first, then this rank's video tokens.
        Returns (out_text [B, S_t, H*64] - identical on every rank, out_video [B, S_loc, H*64])."""
        P = self.world
        B, H, S_in, hd = q.shape
        if H % P:
            raise ValueError(f"sequence parallelism needs the head count ({H}) to divide by the group size ({P})")
        Hl, S_loc = H // P, S_in - S_t
        S_v = P * S_loc
        h0 = self.rank * Hl
        # (1) video rows: heads scattered, tokens gathered.  send[p] = my tokens, the heads rank p will own
        send = torch.stack([q[:, :, S_t:], k[:, :, S_t:], v[:, :, S_t:]])               # [3, B, H, S_loc, hd]
        send = send.view(3, B, P, Hl, S_loc, hd).permute(2, 0, 1, 3, 4, 5).contiguous()  # [P, 3, B, Hl, S_loc, hd]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)                             # recv[p] = rank p's tokens, my heads
        vid = recv.permute(1, 2, 3, 0, 4, 5).reshape(3, B, Hl, S_v, hd)                  # tokens back in rank (= token) order
        # (2) text rows are replicated: take this rank's heads
        qkv = torch.empty((3, B, Hl, S_t + S_v, hd), device=q.device, dtype=q.dtype)
        qkv[0, :, :, :S_t], qkv[1, :, :, :S_t], qkv[2, :, :, :S_t] = (t[:, h0:h0 + Hl, :S_t] for t in (q, k, v))
        qkv[:, :, :, S_t:] = vid
        # (3) local attention over all S tokens for Hl heads
        fn = self._attention
        if fn is None:
            from . import ops
            fn = ops.attention
        o_t, o_v = fn(qkv[0], qkv[1], qkv[2], S_t)                                       # [B,S_t,Hl*hd], [B,S_v,Hl*hd]
        # (4) video rows back: tokens scattered, heads gathered
        send2 = o_v.view(B, P, S_loc, Hl * hd).permute(1, 0, 2, 3).contiguous()          # [P, B, S_loc, Hl*hd]
        recv2 = torch.empty_like(send2)
        dist.all_to_all_single(recv2, send2, group=self.group)                           # recv2[p] = my tokens, rank p's heads
        out_video = recv2.permute(1, 2, 0, 3).reshape(B, S_loc, H * hd).contiguous()
        # (5) text rows: every rank needs all heads (the text stream is replicated)
        gath = torch.empty((P,) + tuple(o_t.shape), device=o_t.device, dtype=o_t.dtype)
        dist.all_gather_into_tensor(gath.view(P * B, S_t, Hl * hd), o_t.contiguous(), group=self.group)
        out_text = gath.permute(1, 2, 0, 3).reshape(B, S_t, H * hd).contiguous()
        return out_text, out_video
