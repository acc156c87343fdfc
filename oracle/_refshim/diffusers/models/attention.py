"""diffusers.models.attention.FeedForward with activation_fn="gelu-approximate" (activations.py GELU: Linear then
F.gelu(approximate="tanh")), mult 4, dropout layers kept for the state_dict indices (net.0.proj, net.2)."""
import torch.nn.functional as F
from torch import nn

from ._placeholder import placeholder
from .attention_processor import Attention  # noqa: F401  (re-exported like diffusers does)


class GELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False,
                 inner_dim=None, bias=True):
        super().__init__()
        inner_dim = int(dim * mult) if inner_dim is None else inner_dim
        dim_out = dim if dim_out is None else dim_out
        if activation_fn == "gelu-approximate":
            act = GELU(dim, inner_dim, approximate="tanh", bias=bias)
        elif activation_fn == "gelu":
            act = GELU(dim, inner_dim, bias=bias)
        else:
            raise NotImplementedError(f"diffusers shim: activation_fn={activation_fn}")
        layers = [act, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out, bias=bias)]
        if final_dropout:
            layers.append(nn.Dropout(dropout))
        self.net = nn.ModuleList(layers)

    def forward(self, hidden_states, *args, **kwargs):
        for m in self.net:
            hidden_states = m(hidden_states)
        return hidden_states


BasicTransformerBlock = placeholder("BasicTransformerBlock")
