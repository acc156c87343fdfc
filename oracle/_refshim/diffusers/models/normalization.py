"""diffusers.models.normalization.AdaLayerNorm as the reference's norm_out uses it (transformer3d.py:1471-1477:
embedding_dim=time_embed_dim, output_dim=2*inner_dim, chunk_dim=1, temb given): shift comes FIRST in the chunk."""
from torch import nn

from ._placeholder import placeholder


class AdaLayerNorm(nn.Module):
    def __init__(self, embedding_dim, num_embeddings=None, output_dim=None, norm_elementwise_affine=False, norm_eps=1e-5,
                 chunk_dim=0):
        super().__init__()
        assert num_embeddings is None
        self.chunk_dim = chunk_dim
        output_dim = output_dim or embedding_dim * 2
        self.emb = None
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, output_dim)
        self.norm = nn.LayerNorm(output_dim // 2, norm_eps, norm_elementwise_affine)

    def forward(self, x, timestep=None, temb=None):
        temb = self.linear(self.silu(temb))
        if self.chunk_dim == 1:
            shift, scale = temb.chunk(2, dim=1)
            shift, scale = shift[:, None, :], scale[:, None, :]
        else:
            scale, shift = temb.chunk(2, dim=0)
        return self.norm(x) * (1 + scale) + shift


AdaLayerNormContinuous = placeholder("AdaLayerNormContinuous")
AdaLayerNormZero = placeholder("AdaLayerNormZero")
CogVideoXLayerNormZero = placeholder("CogVideoXLayerNormZero")
