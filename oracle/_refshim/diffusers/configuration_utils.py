"""diffusers.configuration_utils: just enough of ConfigMixin / register_to_config for `self.config.<name>`."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        self._internal_dict = FrozenDict(kwargs)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        return cls(**{k: v for k, v in {**dict(config), **kwargs}.items() if k in sig})


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        init(self, *args, **kwargs)
        self.register_to_config(**cfg)

    return inner_init
