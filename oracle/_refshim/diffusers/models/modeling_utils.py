import torch


class ModelMixin(torch.nn.Module):
    _supports_gradient_checkpointing = False
    _keys_to_ignore_on_load_unexpected = None

    def __init__(self):
        super().__init__()
        self.gradient_checkpointing = False

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def __getattr__(self, name):
        """diffusers ModelMixin.__getattr__: a registered config entry is readable as an attribute (deprecated there, but the
        reference relies on it: pipeline_easyanimate_inpaint.py:1296 reads `transformer.enable_clip_in_inpaint`, which the
        V5.1 class never assigns)."""
        if "_internal_dict" in self.__dict__ and name in self.__dict__["_internal_dict"] and name not in self.__dict__:
            return self.__dict__["_internal_dict"][name]
        return super().__getattr__(name)
