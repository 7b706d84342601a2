"""pydrake bridge: the probe BASELINE.md 3 asks for and the DOF-order derivation of SURVEY.md E.1.

pydrake has not been importable in the build container or on the GPU boxes so far; `derive_v_index` is therefore covered
by tests/test_pydrake_hooks.py against a stand-in plant exposing the same three pydrake calls the reference itself uses
(`GetJointByName`, `Joint.velocity_start`, `MakeActuationMatrix`; reference controllers/basic_controller.py:113,311-313).
"""
from __future__ import annotations

import importlib

import numpy as np

# internal joint order of wbc_model (legs LF RF LH RH x abduction / hip / knee) -> joint names in the two reference URDFs
LEG_JOINTS = {
    "mini_cheetah": [n for leg in ("fl", "fr", "hl", "hr")
                     for n in (f"torso_to_abduct_{leg}_j", f"abduct_{leg}_to_thigh_{leg}_j", f"thigh_{leg}_to_knee_{leg}_j")],
    "anymal_b": [f"{leg}_{j}" for leg in ("LF", "RF", "LH", "RH") for j in ("HAA", "HFE", "KFE")],
}


def probe() -> dict:
    """{"importable", "version", "error"} for `import pydrake` (run at harness start)."""
    try:
        m = importlib.import_module("pydrake")
        importlib.import_module("pydrake.all")
        return {"importable": True, "version": getattr(m, "__version__", None), "error": None}
    except Exception as e:  # noqa: BLE001
        return {"importable": False, "version": None, "error": f"{type(e).__name__}: {e}"}


def is_drake_plant(obj) -> bool:
    return all(hasattr(obj, a) for a in ("GetJointByName", "MakeActuationMatrix", "num_velocities"))


def derive_v_index(plant, robot="mini_cheetah"):
    """DOF order of a live MultibodyPlant -> (v_index[12], act_index[12]) of wbc_model. Drake's velocity numbering is version
    dependent (breadth-first in the reference's 2021-era Drake, depth-first since 2023), and the reference notes
    that the actuator <-> velocity map is not the identity (basic_controller.py:311-313): read both from the plant."""
    names = LEG_JOINTS[robot]
    v_index = np.array([int(plant.GetJointByName(n).velocity_start()) for n in names], dtype=np.int32)
    B = np.asarray(plant.MakeActuationMatrix(), dtype=float)
    if B.shape != (18, 12) or sorted(v_index.tolist()) != list(range(6, 18)):
        raise ValueError(f"unexpected dof layout: v_index {v_index.tolist()}, B {B.shape}")
    act_index = np.array([int(np.argmax(np.abs(B[v_index[k]]))) for k in range(12)], dtype=np.int32)
    if sorted(act_index.tolist()) != list(range(12)) or not np.allclose(np.abs(B).sum(axis=0), 1.0):
        raise ValueError(f"actuation matrix is not a selection: act_index {act_index.tolist()}")
    return v_index, act_index


def robot_of_plant(plant) -> str:
    """mini_cheetah or anymal_b, from the joint names the plant knows."""
    for robot, names in LEG_JOINTS.items():
        try:
            plant.GetJointByName(names[0])
            return robot
        except Exception:  # noqa: BLE001
            continue
    raise ValueError("plant is neither mini_cheetah nor anymal_b (no known leg joint names)")
