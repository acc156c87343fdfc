"""diffusers.pipelines.pipeline_utils.DiffusionPipeline: what EasyAnimatePipeline.__init__ / __call__ use of it
(pipeline_easyanimate.py:221-240,895-1148): register_modules, _execution_device, progress_bar, maybe_free_model_hooks.
No hub / offload / device-map machinery."""
import contextlib

import torch


class _Bar:
    def update(self, n=1):
        pass


class DiffusionPipeline:
    def __init__(self):
        self._modules_registered = []

    def register_modules(self, **kwargs):
        for name, module in kwargs.items():
            setattr(self, name, module)
            self._modules_registered.append(name)

    @property
    def _execution_device(self):
        for name in self._modules_registered:
            m = getattr(self, name)
            if isinstance(m, torch.nn.Module):
                for p in m.parameters():
                    return p.device
        return torch.device("cpu")

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    def maybe_free_model_hooks(self):
        pass


def is_accelerate_available():
    return False


def is_accelerate_version(op, ver):
    return False
