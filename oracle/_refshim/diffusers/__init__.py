"""Minimal stand-in for `diffusers` (pinned by the reference's requirements.txt to >=0.30.1,<=0.31.0), used ONLY to
import the reference's own modules (easyanimate/vae/ldm/models/omnigen_enc_dec.py, easyanimate/models/transformer3d.py
and what they pull in) in a container where diffusers cannot be installed.  Test infrastructure (oracle pinning and
golden-vector minting, oracle/ref_vae.py, oracle/ref_dit.py); never on the product path, never on the GPU box.

What the EasyAnimateV5.1 path actually executes from diffusers is restated here from its published algorithm
(Attention container, FeedForward/GELU, AdaLayerNorm, Timesteps/TimestepEmbedding, apply_rotary_emb, ConfigMixin /
ModelMixin plumbing); every other imported name is a placeholder that raises if it is ever instantiated."""
__version__ = "0.31.0"


class AutoencoderKL:  # imported by easyanimate/models/autoencoder_magvit.py:41, never used on the decode path
    def __init__(self, *a, **k):
        raise NotImplementedError("diffusers shim: AutoencoderKL is a placeholder")

from .pipelines.pipeline_utils import DiffusionPipeline  # noqa: E402  (pipeline_easyanimate.py:22)
from .models._placeholder import placeholder as _placeholder  # noqa: E402

ImagePipelineOutput = _placeholder("ImagePipelineOutput")  # pipeline_easyanimate_control.py:24
