"""Config surface: the reference's Hydra ``conf/config_fp.yaml`` groups without Hydra.

The reference composes ``conf/config_fp.yaml`` (defaults list: ``_self_``,
``exp_type_fp: TL12``, ``physics_fp: basicTL12``; reference ``conf/config_fp.yaml:1-4``)
through ``hydra.main`` and hands a ``DictConfig`` to ``simulate`` / ``optimizer.main``
(reference ``src/class/simulate.py:11``, ``src/class/optimizer.py:89``).  hydra-core and
omegaconf are not part of this image, so this module provides the small slice of that
surface the hot path touches:

* :class:`DictConfig` / :class:`ListConfig` -- attribute + item access, ``.get``,
  mapping protocol (``Physics(**cfg.physics_fp)``), attribute assignment, shallow ``copy``;
* :func:`load_yaml` -- YAML 1.1 loader with OmegaConf's float resolver (``1e12`` is a
  float, not a string);
* :func:`compose` -- defaults list + ``group=choice`` + dotted ``a.b=v`` + ``+new=v`` overrides;
* :func:`initialize_runs` -- list-valued fields are sweep axes, zip-cycled to the longest
  (reference ``src/class/engine.py:13-38``; ``ValueError`` on non-divisible lengths).

If real omegaconf objects are handed in they work too: everything downstream only uses
the mapping protocol.
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, Iterable, Iterator, List, Mapping, Optional, Sequence

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conf")
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class ListConfig(list):
    """List node.  ``initialize_runs`` treats exactly these as sweep axes."""

    def copy(self) -> "ListConfig":  # keep the node type through copies
        return ListConfig(self)


class DictConfig(dict):
    """Dict node with attribute access (the part of omegaconf.DictConfig the path uses)."""

    def __init__(self, data: Optional[Mapping[str, Any]] = None):
        super().__init__()
        if data:
            for k, v in data.items():
                self[k] = v

    # -- node wrapping ---------------------------------------------------------
    @staticmethod
    def _wrap(v: Any) -> Any:
        if isinstance(v, (DictConfig, ListConfig)):
            return v
        if isinstance(v, Mapping):
            return DictConfig(v)
        if isinstance(v, (list, tuple)):
            return ListConfig(DictConfig._wrap(x) for x in v)
        return v

    def __setitem__(self, k: str, v: Any) -> None:
        super().__setitem__(k, DictConfig._wrap(v))

    # -- attribute protocol ----------------------------------------------------
    def __getattr__(self, k: str) -> Any:
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            return self[k]
        except KeyError:
            # omegaconf raises ConfigAttributeError (an AttributeError), which is what makes
            # ``getattr(cfg, "T_start", 0)`` (reference simulate.py:53) fall back to 0.
            raise AttributeError(f"Missing key {k}") from None

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = v

    def __delattr__(self, k: str) -> None:
        try:
            del self[k]
        except KeyError:
            raise AttributeError(k) from None

    def copy(self) -> "DictConfig":
        """Shallow, like omegaconf's ``DictConfig.copy()`` (nested nodes are shared)."""
        out = DictConfig()
        for k, v in self.items():
            dict.__setitem__(out, k, v)
        return out

    def deepcopy(self) -> "DictConfig":
        return DictConfig(to_container(self))


def to_container(node: Any) -> Any:
    """Plain python containers (OmegaConf.to_container)."""
    if isinstance(node, Mapping):
        return {k: to_container(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [to_container(v) for v in node]
    return node


class OmegaConf:
    """Name-compatible helpers (``OmegaConf.create`` is used at reference engine.py:36-37)."""

    @staticmethod
    def create(obj: Any = None) -> Any:
        if obj is None:
            return DictConfig()
        return DictConfig._wrap(obj)

    @staticmethod
    def to_container(node: Any, resolve: bool = True) -> Any:  # noqa: ARG004
        return to_container(node)

    @staticmethod
    def load(path: str) -> Any:
        return DictConfig._wrap(load_yaml(path))


# ---------------------------------------------------------------------------------------
# YAML with OmegaConf's float rule: exponent forms without a dot are floats.
# ---------------------------------------------------------------------------------------
class _Loader(yaml.SafeLoader):
    pass


_FLOAT_RE = re.compile(
    r"""^(?:
     [-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
    |\.[0-9_]+(?:[eE][-+][0-9]+)?
    |[-+]?\.(?:inf|Inf|INF)
    |\.(?:nan|NaN|NAN))$""",
    re.X,
)
_Loader.add_implicit_resolver("tag:yaml.org,2002:float", _FLOAT_RE, list("-+0123456789."))


def load_yaml(path: str) -> Any:
    with open(path, "r", encoding="utf-8") as fh:
        return yaml.load(fh, Loader=_Loader)


def parse_value(text: str) -> Any:
    """Parse an override value the way Hydra does for simple cases (YAML scalars and lists)."""
    return yaml.load(text, Loader=_Loader)


# ---------------------------------------------------------------------------------------
# Composition
# ---------------------------------------------------------------------------------------
def _resolve_group_file(config_dir: str, group: str, choice: str) -> str:
    path = os.path.join(config_dir, group, f"{choice}.yaml")
    if not os.path.isfile(path):
        avail = sorted(
            os.path.splitext(f)[0]
            for f in os.listdir(os.path.join(config_dir, group))
            if f.endswith(".yaml")
        ) if os.path.isdir(os.path.join(config_dir, group)) else []
        raise FileNotFoundError(
            f"Could not find '{group}/{choice}'. Available options in '{group}': {avail}"
        )
    return path


def compose(
    config_name: str = "config_fp",
    overrides: Optional[Sequence[str]] = None,
    config_dir: Optional[str] = None,
) -> DictConfig:
    """Compose ``<config_dir>/<config_name>.yaml`` like ``hydra.compose``.

    Supported override grammar (every form that appears in the reference's
    ``outputs/*/.hydra/overrides.yaml`` snapshots):

    ``exp_type_fp=TLlab``            swap a defaults-list group
    ``physics_fp.D=0.092``           set an existing (possibly nested) key
    ``exp_type_fp.T_rate=[20]``      values are YAML, lists allowed
    ``+tag=foo`` / ``++tag=foo``     add (or force) a key
    ``tag=foo``                      also accepted for top-level flags read with ``cfg.get``
    ``~key``                         delete a key
    """
    config_dir = config_dir or CONFIG_DIR
    overrides = list(overrides or [])
    root = load_yaml(os.path.join(config_dir, f"{config_name}.yaml")) or {}
    defaults = root.pop("defaults", [])
    root.pop("hydra", None)  # job.chdir etc. -- process control, not config content

    groups: Dict[str, str] = {}
    order: List[str] = []
    for item in defaults:
        if item == "_self_":
            order.append("_self_")
        elif isinstance(item, Mapping):
            for g, c in item.items():
                groups[g] = c
                order.append(g)
        else:  # bare config name
            order.append(str(item))

    value_overrides: List[tuple] = []
    for ov in overrides:
        if ov.startswith("~"):
            value_overrides.append(("del", ov[1:].split("=")[0], None))
            continue
        if "=" not in ov:
            raise ValueError(f"Cannot parse override '{ov}'")
        key, val = ov.split("=", 1)
        mode = "set"
        while key.startswith("+"):
            mode = "add"
            key = key[1:]
        if mode == "set" and key in groups and "." not in key:
            groups[key] = val
            continue
        value_overrides.append((mode, key, parse_value(val)))

    cfg = DictConfig()
    for name in order:
        if name == "_self_":
            for k, v in root.items():
                cfg[k] = v
        elif name in groups:
            cfg[name] = load_yaml(_resolve_group_file(config_dir, name, groups[name])) or {}
        else:
            extra = load_yaml(os.path.join(config_dir, f"{name}.yaml")) or {}
            for k, v in extra.items():
                cfg[k] = v

    for mode, key, val in value_overrides:
        parts = key.split(".")
        node = cfg
        for p in parts[:-1]:
            if p not in node:
                if mode == "add":
                    node[p] = {}
                else:
                    raise KeyError(f"Could not override '{key}': key '{p}' not in config")
            node = node[p]
        leaf = parts[-1]
        if mode == "del":
            node.pop(leaf, None)
        else:
            # Hydra proper refuses ``a.b=v`` for a missing b without '+'; the reference's own
            # docs call top-level flags without it ("task=train", optimizer.py:8), so accept.
            node[leaf] = val
    return cfg


# ---------------------------------------------------------------------------------------
# Sweep expansion
# ---------------------------------------------------------------------------------------
def _is_list_node(v: Any) -> bool:
    if isinstance(v, ListConfig):
        return True
    return type(v).__name__ == "ListConfig"  # a genuine omegaconf node


def initialize_runs(cfg: Mapping[str, Any]) -> Dict[int, DictConfig]:
    """Expand list-valued fields of ``exp_type_fp`` / ``physics_fp`` into run configs.

    Mirrors reference ``src/class/engine.py:13-38``: the longest list sets the number of
    runs; shorter lists are cycled and must divide it (``ValueError`` otherwise).  Despite the
    reference's docstring this is a zip, not a cartesian product.
    """
    subs = {"exp_type_fp": cfg["exp_type_fp"], "physics_fp": cfg["physics_fp"]}
    lengths = [len(v) for sub in subs.values() for v in sub.values() if _is_list_node(v)]
    max_length = max(lengths) if lengths else 1

    runs: Dict[int, DictConfig] = {}
    for i in range(max_length):
        run = DictConfig()
        for sub_name, sub in subs.items():
            node = DictConfig()
            for k, v in sub.items():
                if _is_list_node(v):
                    lst = list(v)
                    if max_length % len(lst) != 0:
                        raise ValueError(
                            f"max_length {max_length} is not divisible by "
                            f"len({sub_name}.{k})={len(lst)}"
                        )
                    node[k] = lst[i % len(lst)]
                else:
                    node[k] = v
            run[sub_name] = node
        runs[i] = run
    return runs


PHYSICS_FIELDS = (
    "alpha", "b", "s", "E_cb", "E_loc_1", "E_loc_2", "D0", "Retrap",
    "name", "D", "rho_trap", "k_b",
)
PHYSICS_REQUIRED = PHYSICS_FIELDS[:8]
K_B_DEFAULT = 8.617333262145e-5  # reference engine.py:60


def physics_record(phys: Mapping[str, Any]) -> Dict[str, Any]:
    """Validate a ``physics_fp`` node the way ``Physics(**cfg.physics_fp)`` does.

    The reference builds a frozen dataclass from the node (``tl_trap_lab.py:30``), so unknown
    keys (e.g. ``E`` in ``BG_basic.yaml``) and missing required ones raise ``TypeError``.
    """
    unknown = [k for k in phys.keys() if k not in PHYSICS_FIELDS]
    if unknown:
        raise TypeError(f"Physics.__init__() got an unexpected keyword argument '{unknown[0]}'")
    missing = [k for k in PHYSICS_REQUIRED if k not in phys]
    if missing:
        raise TypeError(
            f"Physics.__init__() missing {len(missing)} required positional argument(s): "
            + ", ".join(repr(m) for m in missing)
        )
    rec = {k: phys[k] for k in phys.keys()}
    rec.setdefault("name", None)
    rec.setdefault("D", None)
    rec.setdefault("rho_trap", None)
    rec.setdefault("k_b", K_B_DEFAULT)
    return rec
