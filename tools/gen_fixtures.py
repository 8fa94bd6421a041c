#!/usr/bin/env python3
"""Generate the problem fixtures under upright_b200/data/ from the YAML
configuration trees (the reference's own files, read in place from
/root/reference, plus this repo's configs/packages).  Run in the build
container; the JSON outputs are committed so the GPU box needs no reference."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import config, problem_io, settings  # noqa: E402


def main():
    for name, rel in problem_io.FIXTURES.items():
        pkg, path = rel.split("/", 1)
        cfg_path = config.resolve_package_path({"package": pkg, "path": path})
        cfg = config.load_config(cfg_path)
        ctrl = cfg["controller"]
        s = settings.ControllerSettings(ctrl)
        desc = s.to_desc()
        q0 = s.initial_state[: desc.nq]
        r0, C0 = s.chain.tool_pose(q0)
        wp = ctrl["waypoints"][0]["position"] if "waypoints" in ctrl else [0, 0, 0]
        meta = {
            "source": rel,
            "x0": s.initial_state.tolist(),
            "r_ee0": r0.tolist(),
            "waypoint": list(map(float, wp)),
            "body_names": s.body_names(),
            "contact_names": [[c.object1_name, c.object2_name] for c in s.balancing_settings.contacts],
            "frictionless": bool(desc.nf == 1),
            "plane_spans": [np.asarray(c.span).tolist() for c in s.balancing_settings.contacts],
            # merged controller dictionary (what ControllerSettings(config) consumes), so the
            # reference-facing surface can be driven on the GPU box without the YAML tree
            "controller_config": cfg["controller"],
        }
        problem_io.save_fixture(name, desc, meta)
        print(f"{name}: nq={desc.nq} nx={desc.nx} nu={desc.nu} nb={desc.nb} nc={desc.nc} nf={desc.nf} "
              f"N={desc.N} eq={desc.n_eq} fric={desc.n_fric} obs={desc.n_obs} slacks={desc.slacks.enabled}")


if __name__ == "__main__":
    main()
