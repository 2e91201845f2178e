"""A reader for exactly the gin-config subset the reference uses (src/scrabble_gan.gin, main.py:16-18,25,38,43,56):
`scope.param = <python literal>` bindings, `@name` references to registered configurables, `#` comments.
gin-config itself is not installable in this image; binding names are kept identical so the reference's .gin file
parses unchanged."""
from __future__ import annotations

import ast
import functools
import inspect
from typing import Any, Callable, Dict

_REGISTRY: Dict[str, Callable] = {}
_BINDINGS: Dict[str, Dict[str, Any]] = {}


class _Ref:
    def __init__(self, name: str):
        self.name = name

    def resolve(self):
        if self.name not in _REGISTRY:
            raise KeyError("gin reference @{} is not a registered configurable".format(self.name))
        return _REGISTRY[self.name]


def _wrap(fn: Callable, name: str) -> Callable:
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        bound = sig.bind_partial(*args, **kwargs)
        for k, v in _BINDINGS.get(name, {}).items():
            if k in sig.parameters and k not in bound.arguments:
                kwargs[k] = v.resolve() if isinstance(v, _Ref) else v
        return fn(*args, **kwargs)
    wrapper.__wrapped_configurable__ = fn
    return wrapper


def configurable(name_or_fn=None):
    """@gin.configurable or @gin.configurable('scope')."""
    if callable(name_or_fn):
        fn = name_or_fn
        w = _wrap(fn, fn.__name__)
        _REGISTRY[fn.__name__] = w
        return w

    def deco(fn):
        name = name_or_fn or fn.__name__
        w = _wrap(fn, name)
        _REGISTRY[name] = w
        return w
    return deco


def external_configurable(fn: Callable, name: str = None) -> Callable:
    name = name or fn.__name__
    _REGISTRY[name] = fn
    return fn


def parse_config(text: str) -> None:
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if "=" not in line:
            raise ValueError("cannot parse gin line: {!r}".format(raw))
        lhs, rhs = (s.strip() for s in line.split("=", 1))
        scope, _, param = lhs.rpartition(".")
        if not scope:
            raise ValueError("gin binding needs scope.param: {!r}".format(raw))
        value: Any = _Ref(rhs[1:]) if rhs.startswith("@") else ast.literal_eval(rhs)
        _BINDINGS.setdefault(scope, {})[param] = value


def parse_config_file(path: str) -> None:
    with open(path) as f:
        parse_config(f.read())


def clear_config() -> None:
    _BINDINGS.clear()


def query_parameter(key: str):
    scope, _, param = key.rpartition(".")
    v = _BINDINGS[scope][param]
    return v.resolve() if isinstance(v, _Ref) else v
