"""diffusers.models.autoencoders.vae: the two small value classes the reference's AutoencoderKLMagvit returns."""
from dataclasses import dataclass

import torch


@dataclass
class DecoderOutput:
    sample: "torch.Tensor"
    commit_loss: "torch.Tensor" = None

    def __getitem__(self, i):
        return (self.sample,)[i]


class DiagonalGaussianDistribution:
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean
