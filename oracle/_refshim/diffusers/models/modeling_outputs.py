from dataclasses import dataclass

import torch


@dataclass
class Transformer2DModelOutput:
    sample: "torch.Tensor"

    def __getitem__(self, i):
        return (self.sample,)[i]


@dataclass
class AutoencoderKLOutput:
    latent_dist: object

    def __getitem__(self, i):
        return (self.latent_dist,)[i]
