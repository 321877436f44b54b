"""Import shim for hydra (absent from this image).  Test infrastructure only.

``main`` composes the config from ``config_path``/``config_name`` when the task function is called
without arguments and passes straight through when called with a cfg (plots.py:157 does that).
"""
import functools
import sys

from mcluminescence_b200.config import compose as _compose


def main(version_base=None, config_path=None, config_name=None):  # noqa: ARG001
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            if args or kwargs:
                return fn(*args, **kwargs)
            cfg = _compose(config_name, sys.argv[1:], config_dir=config_path)
            return fn(cfg)
        return wrapper
    return deco


def compose(config_name=None, overrides=None, **_):
    return _compose(config_name, overrides)


class _Ctx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


initialize = _Ctx
initialize_config_dir = _Ctx
