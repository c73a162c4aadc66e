"""Template-YAML + group-JSON config loading (the reference's input contract).

The reference merges an OmegaConf YAML template with every entry of a JSON list
(composite.py:94, inverse.py:143) and relies on ``${key}`` interpolation
(configs/group_composite/template.yaml:13,18,21,23).  OmegaConf is not a dependency here;
this is the small subset those files use: deep merge + ``${a.b}`` interpolation + attribute access.
"""
from __future__ import annotations

import json
import re
from typing import Any, Dict, Iterator, List

import yaml

_INTERP = re.compile(r"\$\{([^}]+)\}")


class Config(dict):
    """dict with attribute access (read/write), nested."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x: Any) -> Any:
    if isinstance(x, dict):
        return Config({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def deep_merge(base: Dict, over: Dict) -> Dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = deep_merge(out[k], v)
        else:
            out[k] = v
    return out


def _lookup(root: Dict, dotted: str) -> Any:
    cur: Any = root
    for part in dotted.split("."):
        cur = cur[part]
    return cur


def resolve(root: Dict) -> Dict:
    """Resolve ${a.b} references (strings only), iterating until stable."""

    def res(v: Any, depth: int = 0) -> Any:
        if isinstance(v, str):
            if depth > 16:
                raise ValueError(f"interpolation cycle in {v!r}")
            m = _INTERP.fullmatch(v)
            if m:  # whole-value reference keeps the referenced type
                return res(_lookup(root, m.group(1)), depth + 1)
            return _INTERP.sub(lambda mm: str(res(_lookup(root, mm.group(1)), depth + 1)), v)
        if isinstance(v, dict):
            return {k: res(x, depth) for k, x in v.items()}
        if isinstance(v, list):
            return [res(x, depth) for x in v]
        return v

    return res(root)


def load_template(path: str) -> Dict:
    with open(path) as f:
        return yaml.safe_load(f) or {}


def load_group(path: str) -> List[Dict]:
    with open(path) as f:
        return json.load(f)


def iter_configs(template_path: str, group_json_path: str, only_active: bool = True) -> Iterator[Config]:
    """Yield one merged, interpolated config per (active) entry of the JSON list (composite.py:87-94)."""
    template = load_template(template_path)
    for entry in load_group(group_json_path):
        if only_active and not entry.get("active", True):
            continue
        yield _wrap(resolve(deep_merge(template, entry)))


def merge(template: Dict, entry: Dict) -> Config:
    return _wrap(resolve(deep_merge(template, entry)))
