"""No-op `gin` stand-in: the reference only uses the decorators at import time."""


def configurable(fn_or_name=None, **_kw):
    if callable(fn_or_name):
        return fn_or_name

    def deco(fn):
        return fn

    return deco


def constants_from_enum(cls=None, **_kw):
    if cls is not None:
        return cls
    return lambda c: c


def parse_config_files_and_bindings(*a, **k):
    raise NotImplementedError("gin stand-in")


def clear_config(*a, **k):
    pass
