"""Configuration for the two applications without Hydra.

The reference's CLIs are `@hydra.main(config_path="../config", config_name=...)` functions
(reference application/create_map.py:7-17, application/index_map.py:18-23) over the YAML tree in the reference's
`config/` (map_creation_cfg.yaml, map_indexing_cfg.yaml, map_config/vlmaps.yaml, params/default.yaml, ...).  Hydra and
OmegaConf are not dependencies of this package; `compose` implements the subset of their behaviour those files
use, so that the very same `config/` directory drives `avlmaps_b200.application`:

* a `defaults:` list whose entries are `{group: option}` (-> `<group>/<option>.yaml`, merged under key `group`),
  plain file names, and `_self_` (where the primary file's own keys merge; last if absent, like Hydra >= 1.1);
* `${a.b.c}` interpolation, absolute from the root, resolved on access (a whole-value reference keeps the type
  of its target, one embedded in a longer string is formatted into it);
* command-line overrides: `key.path=value` (value parsed as YAML), `+key.path=value` (add), `group=option`
  (swap an entry of the defaults list), `~key.path` (delete).

`Config` gives attribute AND key access, which the reference's classes mix freely (map.py:23-24,60-66).
"""
from __future__ import annotations

import copy
import re
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional

import yaml

_INTERP = re.compile(r"\$\{([^${}]+)\}")


class ConfigError(ValueError):
    pass


class Config:
    """Read-mostly view of a nested dict with `cfg.a.b` / `cfg["a"]["b"]` access and lazy `${...}` resolution."""

    __slots__ = ("_data", "_root", "_path")

    def __init__(self, data: Dict[str, Any], root: Optional["Config"] = None, path: str = ""):
        object.__setattr__(self, "_data", data)
        object.__setattr__(self, "_root", root if root is not None else self)
        object.__setattr__(self, "_path", path)

    # -- access
    def _wrap(self, where: str, value, depth: int = 0):
        if isinstance(value, dict):
            return Config(value, self._root, where)
        if isinstance(value, list):
            return [self._wrap(f"{where}[{i}]", v, depth) for i, v in enumerate(value)]
        if isinstance(value, str) and "${" in value:
            return self._root._resolve(value, where, depth)
        return value

    def _resolve(self, text: str, where: str, depth: int):
        """(root only) substitute every ${a.b.c} in `text`."""
        if depth > 32:
            raise ConfigError(f"interpolation cycle at {where!r}")
        whole = _INTERP.fullmatch(text)
        if whole:
            return self._lookup(whole.group(1).strip(), where, depth + 1)
        return _INTERP.sub(lambda m: str(self._lookup(m.group(1).strip(), where, depth + 1)), text)

    def _lookup(self, dotted: str, where: str, depth: int):
        if ":" in dotted:
            raise ConfigError(f"resolver ${{{dotted}}} at {where!r} is not supported (plain key paths only)")
        node: Any = self._data
        for part in dotted.split("."):
            if not isinstance(node, dict) or part not in node:
                raise ConfigError(f"interpolation ${{{dotted}}} at {where!r}: key {part!r} not found")
            node = node[part]
        return self._wrap(dotted, node, depth)

    def __getitem__(self, key):
        if key not in self._data:
            raise KeyError(f"{self._path + '.' if self._path else ''}{key}")
        return self._wrap(f"{self._path}.{key}" if self._path else str(key), self._data[key])

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(str(e)) from None

    def __setattr__(self, key, value):
        self._data[key] = value

    __setitem__ = __setattr__

    def get(self, key, default=None):
        return self[key] if key in self._data else default

    def __contains__(self, key):
        return key in self._data

    def __iter__(self):
        return iter(self._data)

    def __len__(self):
        return len(self._data)

    def keys(self):
        return self._data.keys()

    def items(self):
        return [(k, self[k]) for k in self._data]

    def values(self):
        return [self[k] for k in self._data]

    def to_dict(self) -> Dict[str, Any]:
        """Plain nested dict with every interpolation resolved."""
        def plain(v):
            if isinstance(v, Config):
                return {k: plain(v[k]) for k in v}
            if isinstance(v, list):
                return [plain(x) for x in v]
            return v
        return plain(self)

    def __repr__(self):
        return f"Config({self.to_dict()!r})"

    def __eq__(self, other):
        if isinstance(other, Config):
            other = other.to_dict()
        return self.to_dict() == other


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
    """Recursive dict merge, `src` wins; lists and scalars are replaced (OmegaConf.merge semantics)."""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _load_yaml(path: Path) -> Dict[str, Any]:
    if not path.exists():
        raise ConfigError(f"config file {path} does not exist")
    with open(path) as f:
        data = yaml.safe_load(f)
    if data is None:
        return {}
    if not isinstance(data, dict):
        raise ConfigError(f"{path}: top level must be a mapping")
    return data


def _find(config_dir: Path, rel: str) -> Path:
    p = config_dir / rel
    if p.suffix not in (".yaml", ".yml"):
        p = p.with_name(p.name + ".yaml")
    return p


def _set_path(data: Dict[str, Any], dotted: str, value, must_exist: bool, where: str):
    parts = dotted.split(".")
    node = data
    for part in parts[:-1]:
        if part not in node or not isinstance(node[part], dict):
            if must_exist:
                raise ConfigError(f"override {where!r}: key {dotted!r} is not in the config (use +{dotted}=... to add it)")
            node[part] = {}
        node = node[part]
    if must_exist and parts[-1] not in node:
        raise ConfigError(f"override {where!r}: key {dotted!r} is not in the config (use +{dotted}=... to add it)")
    node[parts[-1]] = value


def _del_path(data: Dict[str, Any], dotted: str, where: str):
    parts = dotted.split(".")
    node = data
    for part in parts[:-1]:
        if not isinstance(node, dict) or part not in node:
            raise ConfigError(f"override {where!r}: key {dotted!r} is not in the config")
        node = node[part]
    if parts[-1] not in node:
        raise ConfigError(f"override {where!r}: key {dotted!r} is not in the config")
    del node[parts[-1]]


def compose(config_dir, config_name: str, overrides: Iterable[str] = ()) -> Config:
    """The config Hydra would hand to `main(config)` for `config_dir / config_name` and the given CLI overrides."""
    config_dir = Path(config_dir)
    primary = _load_yaml(_find(config_dir, config_name))
    defaults: List[Any] = primary.pop("defaults", []) or []
    if not isinstance(defaults, list):
        raise ConfigError(f"{config_name}: `defaults` must be a list")
    groups = [next(iter(d)) for d in defaults if isinstance(d, dict) and len(d) == 1]

    value_overrides = []
    for ov in overrides:
        ov = ov.strip()
        if not ov:
            continue
        if ov.startswith("~"):
            value_overrides.append(("del", ov[1:], None, ov))
            continue
        if "=" not in ov:
            raise ConfigError(f"override {ov!r}: expected key=value")
        key, raw = ov.split("=", 1)
        add = key.startswith("+")
        key = key.lstrip("+")
        if not add and "." not in key and key in groups:      # swap a config-group option
            defaults = [({key: raw} if isinstance(d, dict) and next(iter(d)) == key else d) for d in defaults]
            continue
        try:
            value = yaml.safe_load(raw) if raw != "" else ""
        except yaml.YAMLError as e:
            raise ConfigError(f"override {ov!r}: value is not valid YAML ({e})") from None
        value_overrides.append(("add" if add else "set", key, value, ov))

    out: Dict[str, Any] = {}
    self_done = False
    for entry in defaults:
        if entry == "_self_":
            _merge(out, primary)
            self_done = True
        elif isinstance(entry, dict) and len(entry) == 1:
            group, option = next(iter(entry.items()))
            if option is None:
                continue
            sub = _load_yaml(_find(config_dir, f"{group}/{option}"))
            if "defaults" in sub:
                raise ConfigError(f"{group}/{option}.yaml: nested defaults lists are not supported")
            _merge(out.setdefault(group, {}), sub)
        elif isinstance(entry, str):
            _merge(out, _load_yaml(_find(config_dir, entry)))
        else:
            raise ConfigError(f"{config_name}: unsupported defaults entry {entry!r}")
    if not self_done:
        _merge(out, primary)

    for kind, key, value, ov in value_overrides:
        if kind == "del":
            _del_path(out, key, ov)
        else:
            _set_path(out, key, value, must_exist=(kind == "set"), where=ov)
    cfg = Config(out)
    cfg.to_dict()      # resolve everything once: a dangling ${...} fails here, not deep inside a build
    return cfg
