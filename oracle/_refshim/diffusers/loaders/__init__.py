class FromOriginalModelMixin:  # single-file checkpoint loading: not used by the oracle helpers
    pass


FromOriginalVAEMixin = FromOriginalModelMixin
