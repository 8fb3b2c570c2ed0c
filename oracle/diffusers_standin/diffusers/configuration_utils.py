"""ConfigMixin / register_to_config: records constructor kwargs on `self.config`."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kwargs)
        self._internal_dict = FrozenDict(cfg)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        args = {k: v for k, v in dict(config).items() if k in sig}
        args.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**args)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self" and not k.startswith("_")}
        if "kwargs" in cfg:
            cfg.update(cfg.pop("kwargs"))
        ConfigMixin.register_to_config(self, **cfg)
        init(self, *args, **kwargs)

    return inner
