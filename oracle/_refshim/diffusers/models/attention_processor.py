"""diffusers.models.attention_processor.Attention, the branch the reference configures (attention.py:1055-1073:
query_dim, dim_head, heads, qk_norm="layer_norm", eps=1e-6, bias=True, processor=...): a container of to_q/to_k/to_v,
per-head LayerNorms and to_out = [Linear, Dropout]; forward hands itself to the processor, passing only the keyword
arguments the processor's __call__ accepts (diffusers 0.30 attention_processor.py Attention.forward)."""
import inspect

import torch
from torch import nn

from ._placeholder import placeholder


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 qk_norm=None, eps=1e-5, out_bias=True, processor=None, **unused):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.query_dim = query_dim
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        if qk_norm is None:
            self.norm_q = self.norm_k = None
        elif qk_norm == "layer_norm":
            self.norm_q = nn.LayerNorm(dim_head, eps=eps, elementwise_affine=True)
            self.norm_k = nn.LayerNorm(dim_head, eps=eps, elementwise_affine=True)
        else:
            raise NotImplementedError(f"diffusers shim: qk_norm={qk_norm}")
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = processor

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        accepted = set(inspect.signature(self.processor.__call__).parameters.keys())
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in accepted}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


AttentionProcessor = placeholder("AttentionProcessor")
AttnProcessor2_0 = placeholder("AttnProcessor2_0")
HunyuanAttnProcessor2_0 = placeholder("HunyuanAttnProcessor2_0")
AttnAddedKVProcessor = placeholder("AttnAddedKVProcessor")
AttnProcessor = placeholder("AttnProcessor")
ADDED_KV_ATTENTION_PROCESSORS = ()
CROSS_ATTENTION_PROCESSORS = ()
