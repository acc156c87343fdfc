def placeholder(name):
    class _P:  # imported by the reference but never instantiated on the EasyAnimateV5.1 path
        def __init__(self, *a, **k):
            raise NotImplementedError(f"diffusers shim: {name} is a placeholder (not on the EasyAnimateV5.1 path)")

    _P.__name__ = name
    return _P
