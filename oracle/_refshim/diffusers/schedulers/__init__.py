"""diffusers.schedulers: FlowMatchEulerDiscreteScheduler (0.30/0.31) restated from its published algorithm with the call
surface the reference pipeline uses (retrieve_timesteps -> set_timesteps(n, device=, mu=), .timesteps, .order,
step(model_output, t, sample, generator=, return_dict=False)[0]; no scale_model_input / init_noise_sigma attributes, which
the pipeline probes with hasattr).  Third-party, UNPINNED (no diffusers install to compare with): tests/test_diffusers_pin.py
compares it with the real class wherever `diffusers` is importable."""
import math

import numpy as np
import torch

from ..configuration_utils import FrozenDict
from ..models._placeholder import placeholder

DDIMScheduler = placeholder("DDIMScheduler")
DPMSolverMultistepScheduler = placeholder("DPMSolverMultistepScheduler")


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shift=1.0, use_dynamic_shifting=False, base_shift=0.5, max_shift=1.15,
                 base_image_seq_len=256, max_image_seq_len=4096):
        self.config = FrozenDict(num_train_timesteps=num_train_timesteps, shift=shift, use_dynamic_shifting=use_dynamic_shifting,
                                 base_shift=base_shift, max_shift=max_shift, base_image_seq_len=base_image_seq_len,
                                 max_image_seq_len=max_image_seq_len)
        timesteps = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        timesteps = torch.from_numpy(timesteps).to(dtype=torch.float32)
        sigmas = timesteps / num_train_timesteps
        if not use_dynamic_shifting:
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self._step_index = None
        self._begin_index = None
        self.sigmas = sigmas.to("cpu")
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()

    @property
    def step_index(self):
        return self._step_index

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    def time_shift(self, mu, sigma, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError(" you have a pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        if sigmas is None:
            self.num_inference_steps = num_inference_steps
            timesteps = np.linspace(self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min), num_inference_steps)
            sigmas = timesteps / self.config.num_train_timesteps
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self.config.shift * sigmas / (1 + (self.config.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        timesteps = sigmas * self.config.num_train_timesteps
        self.timesteps = timesteps.to(device=device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._step_index = None
        self._begin_index = None

    @property
    def begin_index(self):
        return self._begin_index

    def scale_noise(self, sample, timestep, noise=None):
        """Forward process of flow matching at `timestep`: sigma * noise + (1 - sigma) * sample, in the SAMPLE's dtype (the sigma
        table is cast to it first) - called by the inpaint pipeline for strength < 1 (pipeline_easyanimate_inpaint.py:896)."""
        sigmas = self.sigmas.to(device=sample.device, dtype=sample.dtype)
        schedule_timesteps = self.timesteps.to(sample.device)
        timestep = timestep.to(sample.device)
        if self.begin_index is None:
            step_indices = [self.index_for_timestep(t, schedule_timesteps) for t in timestep]
        elif self.step_index is not None:
            step_indices = [self.step_index] * timestep.shape[0]
        else:
            step_indices = [self.begin_index] * timestep.shape[0]
        sigma = sigmas[step_indices].flatten()
        while len(sigma.shape) < len(sample.shape):
            sigma = sigma.unsqueeze(-1)
        return sigma * noise + (1.0 - sigma) * sample

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        if schedule_timesteps is None:
            schedule_timesteps = self.timesteps
        indices = (schedule_timesteps == timestep).nonzero()
        pos = 1 if len(indices) > 1 else 0
        return indices[pos].item()

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            if isinstance(timestep, torch.Tensor):
                timestep = timestep.to(self.timesteps.device)
            self._step_index = self.index_for_timestep(timestep)
        else:
            self._step_index = self._begin_index

    def step(self, model_output, timestep, sample, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, generator=None,
             return_dict=True):
        if self._step_index is None:
            self._init_step_index(timestep)
        sample = sample.to(torch.float32)  # upcast to avoid precision issues when computing prev_sample
        sigma = self.sigmas[self._step_index]
        sigma_next = self.sigmas[self._step_index + 1]
        prev_sample = sample + (sigma_next - sigma) * model_output
        prev_sample = prev_sample.to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev_sample,)
        return FrozenDict(prev_sample=prev_sample)
