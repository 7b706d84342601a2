"""Deterministic synthetic inputs for the benchmark and the parity tests.

State: base rpy ~ U(-0.4, 0.4)^3 (amplitude of BasicTrunkPlanner.OrientationTest, reference
planners/simple.py:93), base xy ~ U(-1, 1), z ~ nominal +- 0.05, joints = nominal + U(-0.3, 0.3),
base omega / v ~ U(-0.5, 0.5), joint rates ~ U(-2, 2).
Trajectory: nominal = actual task-space quantity + N(0, 0.02) (positions), N(0, 0.1) (velocities),
N(0, 1) (accelerations); rpy likewise. The actual foot positions/velocities come from a forward-
kinematics callable the caller supplies (the CUDA `wbc_dynamics` entry in the benchmark, the oracle
in CPU tests) - this module does no robot arithmetic of its own.
Contact masks per config: "stand" all four; "trot" = flying-trot stride (towr quadruped_gait_generator.cc
:224-241) LF+RH / flight / RF+LH / flight with weights .4/.1/.4/.1; "walk" = overlap-walk stride
(:182-204) 3- and 2-stance patterns; "mixed" = uniform over all 16 patterns.
"""
from __future__ import annotations

import numpy as np

from .model import NQ, NV, NTRAJ, RobotModel

PATTERNS = {
    "stand": ([[1, 1, 1, 1]], [1.0]),
    # LF RF LH RH; flying trot: bP_ (LF+RH), II_ (flight), Pb_ (RF+LH), II_
    "trot": ([[1, 0, 0, 1], [0, 0, 0, 0], [0, 1, 1, 0], [1, 1, 1, 1]], [0.4, 0.15, 0.4, 0.05]),
    # overlap walk: three-leg and diagonal/lateral two-leg supports
    "walk": ([[1, 1, 1, 0], [1, 1, 0, 1], [1, 0, 1, 1], [0, 1, 1, 1], [1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1]],
             [0.18, 0.18, 0.18, 0.18, 0.1, 0.1, 0.08]),
    "mixed": ([[(i >> 3) & 1, (i >> 2) & 1, (i >> 1) & 1, i & 1] for i in range(16)], [1 / 16.0] * 16),
}


def rpy_to_quat(rpy):
    r, p, y = rpy[..., 0] / 2, rpy[..., 1] / 2, rpy[..., 2] / 2
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], axis=-1)


def random_states(model: RobotModel, n: int, seed: int):
    rng = np.random.default_rng(seed)
    qn = model.nominal_q()
    q = np.tile(qn, (n, 1))
    rpy = rng.uniform(-0.4, 0.4, (n, 3))
    q[:, 0:4] = rpy_to_quat(rpy)
    q[:, 4:6] = rng.uniform(-1.0, 1.0, (n, 2))
    q[:, 6] = qn[6] + rng.uniform(-0.05, 0.05, n)
    q[:, 7:] += rng.uniform(-0.3, 0.3, (n, 12))
    v = np.empty((n, NV))
    v[:, 0:6] = rng.uniform(-0.5, 0.5, (n, 6))
    v[:, 6:] = rng.uniform(-2.0, 2.0, (n, 12))
    return q, v, rpy, rng


def generate(model: RobotModel, n: int, seed: int, pattern="stand", fk=None):
    """-> q[n,19], v[n,18], traj[n,54], contact[n,4] (uint8).
    fk(q, v) -> (p_feet[n,4,3], pd_feet[n,4,3]) in world coordinates."""
    q, v, rpy, rng = random_states(model, n, seed)
    if fk is None:
        raise ValueError("generate() needs a forward-kinematics callable (wbc_dynamics or the test oracle)")
    p_feet, pd_feet = fk(q, v)
    traj = np.zeros((n, NTRAJ))
    # rpy rates from world angular velocity: rpyd = N^-1 omega
    cp, sp, cy, sy = np.cos(rpy[:, 1]), np.sin(rpy[:, 1]), np.cos(rpy[:, 2]), np.sin(rpy[:, 2])
    w = v[:, 0:3]
    rd = (cy * w[:, 0] + sy * w[:, 1]) / cp
    pdot = -sy * w[:, 0] + cy * w[:, 1]
    yd = w[:, 2] + sp * rd
    rpyd = np.stack([rd, pdot, yd], axis=1)
    traj[:, 0:3] = q[:, 4:7] + rng.normal(0, 0.02, (n, 3))
    traj[:, 3:6] = v[:, 3:6] + rng.normal(0, 0.1, (n, 3))
    traj[:, 6:9] = rng.normal(0, 1.0, (n, 3))
    traj[:, 9:12] = rpy + rng.normal(0, 0.02, (n, 3))
    traj[:, 12:15] = rpyd + rng.normal(0, 0.1, (n, 3))
    traj[:, 15:18] = rng.normal(0, 1.0, (n, 3))
    traj[:, 18:30] = (p_feet + rng.normal(0, 0.02, (n, 4, 3))).reshape(n, 12)
    traj[:, 30:42] = (pd_feet + rng.normal(0, 0.1, (n, 4, 3))).reshape(n, 12)
    traj[:, 42:54] = rng.normal(0, 1.0, (n, 12))
    pats, wts = PATTERNS[pattern]
    idx = rng.choice(len(pats), size=n, p=np.asarray(wts) / np.sum(wts))
    contact = np.asarray(pats, dtype=np.uint8)[idx]
    return q, v, traj, contact


def nominal_state(model: RobotModel, n: int):
    """n copies of the nominal standing configuration at rest -> q[n,19], v[n,18]."""
    return np.tile(model.nominal_q(), (n, 1)), np.zeros((n, NV))
