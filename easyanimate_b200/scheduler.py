"""Flow-matching Euler scheduler with the call surface the reference pipelines use on diffusers'
``FlowMatchEulerDiscreteScheduler`` (pinned 0.30.1-0.31.0 by /root/reference/requirements.txt:25; call sites
easyanimate/pipeline/pipeline_easyanimate.py:972 ``retrieve_timesteps(..., mu=1)`` -> ``set_timesteps`` and :1111
``scheduler.step(noise_pred, t, latents, return_dict=False)[0]``).

Host logic (the sigma schedule) is plain Python/torch on CPU like diffusers'; the per-step update
``x <- bf16(float(x) + bf16(bf16(sigma_next - sigma) * v))`` runs in the ``ea_cfg_euler_step`` kernel.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import ops
from .config import FrozenConfig


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = False,
                 base_shift: float = 0.5, max_shift: float = 1.15, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 4096):
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, shift=shift,
                                   use_dynamic_shifting=use_dynamic_shifting, base_shift=base_shift, max_shift=max_shift,
                                   base_image_seq_len=base_image_seq_len, max_image_seq_len=max_image_seq_len)
        timesteps = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        sigmas = torch.from_numpy(timesteps).to(dtype=torch.float32) / num_train_timesteps
        if not use_dynamic_shifting:
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self.sigmas = sigmas
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self._step_index: Optional[int] = None
        self.num_inference_steps: Optional[int] = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **overrides):
        """`FlowMatchEulerDiscreteScheduler.from_pretrained(model_name, subfolder="scheduler")` (predict_t2v.py:225-231): the
        released `scheduler/scheduler_config.json` (shift, use_dynamic_shifting, ...); unknown keys are ignored like diffusers'
        `from_config` ignores them."""
        import inspect
        import json
        import os
        path = pretrained_model_name_or_path if subfolder is None else os.path.join(pretrained_model_name_or_path, subfolder)
        with open(os.path.join(path, "scheduler_config.json")) as f:
            cfg = json.load(f)
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
        return cls(**{**{k: v for k, v in cfg.items() if k in accepted}, **{k: v for k, v in overrides.items() if k in accepted}})

    @property
    def step_index(self):
        return self._step_index

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    @staticmethod
    def time_shift(mu: float, sigma: float, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, sigmas=None, mu: Optional[float] = None):
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError(" you have a pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        if sigmas is None:
            self.num_inference_steps = num_inference_steps
            timesteps = np.linspace(self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min), num_inference_steps)
            sigmas = timesteps / self.config.num_train_timesteps
        else:
            sigmas = np.asarray(sigmas, dtype=np.float64)
            self.num_inference_steps = len(sigmas)
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self.config.shift * sigmas / (1 + (self.config.shift - 1) * sigmas)
        sigmas = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32)
        # timesteps follow the caller's device (the pipeline iterates over them); the sigma table stays on the host so
        # that step() never synchronises the stream to read it.
        self.timesteps = (sigmas * self.config.num_train_timesteps).to(device=device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self._sigmas_host = [float(s) for s in self.sigmas]  # exact fp32 values
        self._step_index = None

    def _init_step_index(self, timestep):
        t = float(timestep)
        ts = [float(x) for x in self.timesteps.cpu()]
        idx = [i for i, v in enumerate(ts) if v == t]
        pos = 1 if len(idx) > 1 else 0
        self._step_index = idx[pos] if idx else 0

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, **unused):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma, sigma_next = self._sigmas_host[self._step_index], self._sigmas_host[self._step_index + 1]
        prev = ops.cfg_euler_step(model_output.to(torch.bfloat16), sample.to(torch.bfloat16), 1.0, sigma, sigma_next,
                                  use_cfg=False)
        self._step_index += 1
        return (prev,) if not return_dict else FrozenConfig(prev_sample=prev)

    def scale_noise(self, sample: torch.Tensor, timestep, noise: torch.Tensor, index: Optional[int] = None) -> torch.Tensor:
        """diffusers' `scale_noise` (the start of a strength < 1 image-to-video / video-to-video call,
        pipeline_easyanimate_inpaint.py:896): sigma * noise + (1 - sigma) * sample with sigma rounded to the sample's dtype and
        tensor ops in that dtype.  One-off preparation before the loop (torch ops)."""
        if index is None:
            ts = [float(x) for x in self.timesteps.cpu()]
            t0 = float(torch.as_tensor(timestep).reshape(-1)[0])
            hits = [i for i, v in enumerate(ts) if v == t0]
            index = hits[1 if len(hits) > 1 else 0]
        sigma = self.sigmas[index].to(device=sample.device, dtype=sample.dtype)
        return sigma * noise + (1.0 - sigma) * sample

    def sigma_pair(self, i: int):
        return self._sigmas_host[i], self._sigmas_host[i + 1]
