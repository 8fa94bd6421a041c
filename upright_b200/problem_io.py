"""(De)serialise `ub_problem_desc_t` to JSON so that tests, smoke() and bench.py
run on the GPU box, where the reference tree (`/root/reference`) is absent.
The fixtures under `upright_b200/data/` are generated from the reference's own
YAML files by `tools/gen_fixtures.py` (committed beside them)."""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

import numpy as np

from . import bindings as B

DATA_DIR = Path(__file__).resolve().parent / "data"


def _to_py(obj):
    if isinstance(obj, C.Structure):
        return {name: _to_py(getattr(obj, name)) for name, _ in obj._fields_}
    if isinstance(obj, C.Array):
        return [_to_py(v) for v in obj]
    return obj


def _from_py(target, value):
    """Fill ctypes struct/array `target` in place from nested python data."""
    if isinstance(target, C.Structure):
        for name, ftype in target._fields_:
            if name not in value:
                continue
            cur = getattr(target, name)
            if isinstance(cur, (C.Structure, C.Array)):
                _from_py(cur, value[name])
            else:
                setattr(target, name, value[name])
    else:  # array
        for i, v in enumerate(value):
            if isinstance(target[i], (C.Structure, C.Array)):
                _from_py(target[i], v)
            else:
                target[i] = v


def desc_to_dict(desc: B.ProblemDesc) -> dict:
    return _to_py(desc)


def desc_from_dict(d: dict) -> B.ProblemDesc:
    desc = B.ProblemDesc()
    _from_py(desc, d)
    return desc


def save_fixture(name: str, desc: B.ProblemDesc, meta: dict, directory=DATA_DIR):
    directory.mkdir(parents=True, exist_ok=True)
    with open(directory / f"{name}.json", "w") as f:
        json.dump({"meta": meta, "desc": desc_to_dict(desc)}, f, indent=1)


def load_fixture(name: str, directory=DATA_DIR):
    """-> (desc, meta).  meta: x0 (home state), r_ee0, source, body_names, ..."""
    with open(Path(directory) / f"{name}.json") as f:
        doc = json.load(f)
    meta = doc["meta"]
    for k in ("x0", "r_ee0", "waypoint"):
        if k in meta:
            meta[k] = np.array(meta[k], dtype=float)
    return desc_from_dict(doc["desc"]), meta


FIXTURES = {
    "cfg1_ur10_demo": "upright_cmd/config/demos/ur10_demo.yaml",
    "cfg2_thing_demo": "upright_cmd/config/demos/thing_demo.yaml",
    "cfg3_thing_box_arch": "upright_b200_cfg/config/thing_box_arch.yaml",
    "cfg4_thing_obstacles2": "upright_b200_cfg/config/thing_obstacles2.yaml",
    "cfg5_thing_robust8": "upright_b200_cfg/config/thing_robust8.yaml",
}


def fp32_conditioning_estimate(desc, body_params=None):
    """Rough ratio between the largest and smallest curvature the force block of a stage matrix sees when the
    object-dynamics rows are SOFT: the rows are divided by the body mass (contact_constraints.h:84-100) and by
    sqrt(6 nb) (balancing_constraints.cpp:144-151) and penalised with Z directly, against a force weight of
    dt * force_weight (floored by the input regularisation).  Above ~1e7 (1 / fp32 epsilon) the fp32 factorisation
    of the kernels cannot be trusted: use the fp64 kernels or UB_RESCUE_F64.  1.0 when the
    rows are hard (normalised proximal weights) or balancing is off."""
    import numpy as np
    if not (desc.balancing_enabled and desc.nb > 0 and desc.slacks.enabled and desc.slacks.poly_ineq):
        return 1.0
    if body_params is None:
        masses = np.array([desc.body_params[b][0] for b in range(desc.nb)])
    else:
        masses = np.asarray(body_params, dtype=float).reshape(-1, desc.nb, 10)[:, :, 0].ravel()
    Z = desc.slacks.upper_L2_penalty if desc.slacks.upper_L2_penalty > 0 else 100.0
    floor = max(desc.dt * desc.force_weight, desc.reg_input, 1e-300)
    return float(Z / (6.0 * desc.nb * masses.min() ** 2) / floor)
