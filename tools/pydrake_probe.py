#!/usr/bin/env python
"""Probe for pydrake and the hooks that use it when it is there.

Neither this container nor the GPU boxes have had pydrake so far, so everything below the probe is exercised only by
`tests/test_pydrake_hooks.py` against a stand-in plant object; it is written against the pydrake API the reference
itself calls (reference controllers/basic_controller.py:29-70,110-113,180-195).

  probe()               -> {"importable": bool, "version": str | None, "error": str | None}
  derive_v_index(plant) -> (v_index[12], act_index[12]) in the internal joint order k = 3 leg + j of wbc_model, read from
                           plant.GetJointByName(name).velocity_start() and the actuator list (Drake's dof
                           numbering is version dependent - never assume it)
  drake_dynamics(plant, context, q, v, foot_frames) -> dict with the keys of tests/golden/*.npz (M, Cv, tau_g, J_feet,
                           Jdv_feet, p_feet) computed by MultibodyPlant itself; tools/make_golden.py --drake stores them as
                           drake_* next to the oracle's and tests/test_oracle_dynamics.py compares the two when present.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from quadruped_drake_b200.drake_bridge import LEG_JOINTS, derive_v_index, probe  # noqa: E402,F401


def drake_dynamics(plant, context, q, v, foot_frames, world_frame=None):
    """The pydrake calls of BasicController.CalcDynamics / CalcFramePositionQuantities (basic_controller.py:101-115,173-196)."""
    from pydrake.all import JacobianWrtVariable  # noqa: PLC0415
    W = world_frame if world_frame is not None else plant.world_frame()
    n = len(q)
    out = dict(M=np.zeros((n, 18, 18)), Cv=np.zeros((n, 18)), tau_g=np.zeros((n, 18)), J_feet=np.zeros((n, 4, 3, 18)),
               Jdv_feet=np.zeros((n, 4, 3)), p_feet=np.zeros((n, 4, 3)))
    for i in range(n):
        plant.SetPositions(context, q[i])
        plant.SetVelocities(context, v[i])
        out["M"][i] = plant.CalcMassMatrixViaInverseDynamics(context)
        out["Cv"][i] = plant.CalcBiasTerm(context)
        out["tau_g"][i] = -plant.CalcGravityGeneralizedForces(context)
        for k, f in enumerate(foot_frames):
            out["p_feet"][i, k] = plant.CalcPointsPositions(context, f, np.zeros(3), W).ravel()
            out["J_feet"][i, k] = plant.CalcJacobianTranslationalVelocity(context, JacobianWrtVariable.kV, f, np.zeros(3), W, W)
            out["Jdv_feet"][i, k] = plant.CalcBiasTranslationalAcceleration(context, JacobianWrtVariable.kV, f, np.zeros(3), W, W).ravel()
    return out


if __name__ == "__main__":
    json.dump(probe(), sys.stdout)
    print()
