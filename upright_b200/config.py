"""ROS-free configuration loading for upright-style YAML trees.

Restates the behaviour of the reference's config layer
(`upright_core/src/upright_core/parsing.py:30-106` for `load_config`, the
number/array DSL and diagonal-matrix dicts; `:109-115` for package-relative
paths) without rospkg/xacrodoc: package names are resolved by searching a list
of root directories for `<root>/<package>/`.
"""
from __future__ import annotations

import copy
import os
from pathlib import Path

import numpy as np
import yaml

_REPO_ROOT = Path(__file__).resolve().parent.parent

#: Directories searched (in order) for `<package>/<path>` includes.  The
#: reference tree is only present in the build container; `configs/packages`
#: holds this repo's own configuration packages.
DEFAULT_PACKAGE_ROOTS = [
    str(_REPO_ROOT / "configs" / "packages"),
    "/root/reference",
]


def package_roots():
    env = os.environ.get("UPRIGHT_PACKAGE_ROOTS")
    roots = env.split(os.pathsep) if env else []
    return roots + DEFAULT_PACKAGE_ROOTS


def resolve_package_path(spec, roots=None) -> Path:
    """Resolve `{package: ..., path: ...}` to a file path (parsing.py:109-115)."""
    roots = package_roots() if roots is None else roots
    for root in roots:
        candidate = Path(root) / spec["package"] / spec["path"]
        if candidate.exists():
            return candidate
    raise FileNotFoundError(
        f"cannot resolve package path {spec['package']}/{spec['path']} in {roots}"
    )


def merge_dicts(base: dict, override: dict) -> dict:
    """Recursive dict merge; `override` wins, nested dicts are merged."""
    if not isinstance(base, dict) or not isinstance(override, dict):
        raise TypeError("merge_dicts needs two dicts")
    for key, val in override.items():
        if isinstance(val, dict) and isinstance(base.get(key), dict):
            base[key] = merge_dicts(base[key], val)
        else:
            base[key] = val
    return base


def load_config(path, roots=None, _depth=0, max_depth=5) -> dict:
    """Load a YAML file honouring `include: [{package, path, key?}]` lists.

    Included files are merged first (in order), optionally nested under `key`,
    and the including file's own content overrides them (parsing.py:30-60).
    """
    if _depth > max_depth:
        raise RecursionError(f"include depth {max_depth} exceeded at {path}")
    with open(path) as f:
        doc = yaml.safe_load(f) or {}
    merged: dict = {}
    for inc in doc.pop("include", []):
        sub = load_config(resolve_package_path(inc, roots), roots, _depth + 1, max_depth)
        if "key" in inc:
            sub = {inc["key"]: sub}
        merged = merge_dicts(merged, sub)
    return merge_dicts(merged, doc)


def parse_number(x, dtype=float):
    """`"0.5pi"` -> 0.5*pi; anything else through `dtype` (parsing.py:63-71)."""
    if isinstance(x, str) and x.endswith("pi"):
        return dtype(x[:-2]) * np.pi
    return dtype(x)


def _array_element(x):
    try:
        return [float(x)]
    except (ValueError, TypeError):
        pass
    if isinstance(x, str):
        if x.endswith("pi"):
            return [float(x[:-2]) * np.pi]
        if "rep" in x:
            val, count = x.split("rep")
            return [float(val)] * int(count)
    raise ValueError(f"could not convert {x!r} to array element")


def parse_array(seq) -> np.ndarray:
    """1-D array with the `"<v>rep<n>"` and `"<v>pi"` shorthands (parsing.py:74-91)."""
    out = []
    for x in seq:
        out.extend(_array_element(x))
    return np.array(out, dtype=float)


def parse_diag_matrix_dict(d) -> np.ndarray:
    """`{scale, diag}` -> scale*diag(diag) (parsing.py:94-106)."""
    return parse_number(d["scale"]) * np.diag(parse_array(d["diag"]))


def parse_support_offset(d) -> np.ndarray:
    """x/y plus optional polar (r, θ) offset of an object on its parent
    (parsing.py:132-151)."""
    x = d.get("x", 0)
    y = d.get("y", 0)
    has_r, has_t = "r" in d, "θ" in d
    if has_r != has_t:
        raise ValueError("support offset needs both r and θ")
    if has_r:
        th = parse_number(d["θ"])
        x = x + d["r"] * np.cos(th)
        y = y + d["r"] * np.sin(th)
    return np.array([x, y], dtype=float)


def deep_copy(cfg):
    return copy.deepcopy(cfg)
