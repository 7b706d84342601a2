"""Trunk planners on top of the device trajectory sampler.

Mirrors the reference's planner interface (output "trunk_trajectory": the dict of planners/simple.py:45-85):

  BasicTrunkPlanner   planners/simple.py   constant standing reference (+ the OrientationTest / RaiseFoot / EdgeTest motions)
  TowrTrunkPlanner    planners/towr.py     stand for 1 s, then the NEAREST stored 1 kHz sample of a TOWR spline solution

The reference obtains the spline solution by running IPOPT in a subprocess (towr/trunk_mpc.cpp) - out of scope. Here a
plan is a set of cubic Hermite node splines + phase durations in TOWR's own layout (`GaitPlan`); `make_gait_plan` fills
it with a synthetic heuristic solution from TOWR's gait tables, or it can be loaded with the nodes of a real solve.
Sampling - towr/trunk_mpc.cpp:19-68 publish_trunk_state + planners/towr.py:92-148 - runs on the GPU for N
(plan, time) pairs per launch through `wbc_sample_trajectory` (csrc/wbc_traj.cuh); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import np_ptr
from .controller import traj_to_dict
from .model import NTRAJ

FEET = ("LF", "RF", "LH", "RH")           # towr endeffector order (endeffector_mappings.h:44) = wbc.h foot order

# Contact-state codes of towr/src/quadruped_gait_generator.cc:38-73: first letter hind legs, second front legs;
# I none, P left, b right, B both.
_SIDE = {"I": (), "P": ("L",), "b": ("R",), "B": ("L", "R")}


def _contacts(code):
    on = {s + "H" for s in _SIDE[code[0]]} | {s + "F" for s in _SIDE[code[1]]}
    return tuple(f in on for f in FEET)


def _stride(times, codes):
    codes = codes.split()
    assert len(times) == len(codes)
    return list(times), [_contacts(c) for c in codes]


def _without_transition(stride):          # GaitGenerator::RemoveTransition, gait_generator.cc:130-143
    t, c = list(stride[0]), list(stride[1])
    t[-2] += t[-1]
    return t[:-1], c[:-1]


_WALK2 = _stride([0.25, 0.13, 0.25, 0.13] * 2, "bB bb Bb Pb PB PP BP bP")
_GALLOP = _stride([0.2, 0.3, 0.2, 0.2] * 2, "Bb BI BP bP bB IB PB Pb")
STRIDES = {                               # quadruped_gait_generator.cc:114-375
    "Stand": _stride([0.3], "BB"), "Flight": _stride([0.3], "Bb"),
    "Walk1": _stride([0.3, 0.2] * 4, "bB BB Bb BB PB BB BP BB"),
    "Walk2": _WALK2, "Walk2E": _without_transition(_WALK2),
    "Run1": _stride([0.3, 0.2] * 2, "bP BB Pb BB"),
    "Run2": _stride([0.4, 0.1] * 2, "bP II Pb II"), "Run2E": _stride([0.4], "bP"),
    "Run3": _stride([0.3, 0.1] * 2, "PP II bb II"), "Run3E": _stride([0.3], "PP"),
    "Hop1": _stride([0.3, 0.1] * 2, "BI II IB II"), "Hop1E": _stride([0.3], "BI"),
    "Hop2": _stride([0.3, 0.4, 0.3], "BB II BB"),
    "Hop3": _GALLOP, "Hop3E": _without_transition(_GALLOP),
    "Hop5": _stride([0.1, 0.2, 0.1] * 2, "Bb BB IP Bb BB IP"),
}
COMBOS = {name: ["Stand"] + [g] * 3 + [g + "E", "Stand"]          # SetCombo, quadruped_gait_generator.cc:76-88
          for name, g in (("walk", "Walk2"), ("trot", "Run2"), ("pace", "Run3"), ("bound", "Hop1"), ("gallop", "Hop3"))}
COMBO_IDS = ["walk", "trot", "pace", "bound", "gallop"]           # C0..C4, argv[1] of towr/trunk_mpc.cpp:82-98


def gait_phase_durations(combo, t_total):
    """Per foot: (phase durations scaled to t_total, in contact at start) - gait_generator.cc:55-111."""
    combo = COMBO_IDS[combo] if isinstance(combo, int) else combo
    times, contacts = [], []
    for g in COMBOS[combo]:
        times += STRIDES[g][0]
        contacts += STRIDES[g][1]
    out = []
    for ee in range(4):
        acc, phases = 0.0, []
        for ph in range(len(contacts) - 1):
            acc += times[ph]
            if contacts[ph][ee] != contacts[ph + 1][ee]:
                phases.append(acc)
                acc = 0.0
        phases.append(acc + times[-1])
        total = 0.0
        for x in phases:
            total += x
        out.append(([(x / total) * t_total for x in phases], bool(contacts[0][ee])))
    return out


def base_poly_durations(t_total, dt=0.1):                         # parameters.cc:83-98
    out, left = [], t_total
    while left > 1e-10:
        out.append(dt if left > dt else left)
        left -= dt
    return out


def split_phases(phase_dur, first_constant, n_changing):          # nodes_variables_phase_based.cc:36-83
    durs, const = [], first_constant
    for d in phase_dur:
        n = 1 if const else n_changing
        durs += [d / n] * n
        const = not const
    return durs


class SplineTable:
    """durations[n_poly] + nodes[n_poly + 1, 6] (position, velocity) of one 3-D cubic Hermite spline."""

    def __init__(self, durations, nodes):
        self.durations = np.ascontiguousarray(durations, dtype=np.float64)
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64).reshape(len(self.durations) + 1, 6)


class GaitPlan:
    """TOWR's SplineHolder as plain arrays (+ the planner settings of planners/towr.py)."""

    def __init__(self, base_linear, base_angular, ee_motion, ee_force, phase_durations, contact_at_start,
                 sample_dt=0.0, wait_time=0.0, standing=None, total_duration=None):
        self.base_linear, self.base_angular, self.ee_motion, self.ee_force = base_linear, base_angular, list(ee_motion), list(ee_force)
        self.phase_durations = [np.ascontiguousarray(p, dtype=np.float64) for p in phase_durations]
        self.contact_at_start = [bool(c) for c in contact_at_start]
        self.sample_dt, self.wait_time = float(sample_dt), float(wait_time)
        self.standing = np.zeros(NTRAJ) if standing is None else np.ascontiguousarray(standing, dtype=np.float64)
        # the duration asked of the solver (trunk_mpc.cpp:126); the spline total differs from it by rounding
        self.total_duration = round(self.total_time(), 9) if total_duration is None else float(total_duration)
        self.grid = self._grid()

    def total_time(self):
        s = 0.0
        for d in self.base_linear.durations:
            s += d
        return s

    def _grid(self):
        """Timestamps trunk_mpc publishes (trunk_mpc.cpp:168-174): accumulated t += dt while t < T, then T itself."""
        if self.sample_dt <= 0.0:
            return np.zeros(0)
        T = self.total_duration
        ts, t = [], 0.0
        while t < T:
            ts.append(t)
            t = t + self.sample_dt
        ts.append(T)
        return np.array(ts)


SIMPLE_STANDING = {                       # planners/simple.py:45-52,74 (mini cheetah literals; anymal ones are commented there)
    "mini_cheetah": (np.array([[0.175, 0.11, 0.0], [0.175, -0.11, 0.0], [-0.2, 0.11, 0.0], [-0.2, -0.11, 0.0]]), 0.3),
    "anymal_b": (np.array([[0.34, 0.19, 0.0], [0.34, -0.19, 0.0], [-0.34, 0.19, 0.0], [-0.34, -0.19, 0.0]]), 0.5),
}
NOMINAL_STANCE = {"mini_cheetah": (0.2, 0.11, -0.30, 9.0), "anymal_b": (0.34, 0.19, -0.42, 29.5)}   # towr models/examples/*.h


def simple_standing(robot="mini_cheetah"):
    """SimpleStanding (planners/simple.py:39-85) as (traj[54], contact[4])."""
    feet, height = SIMPLE_STANDING[robot]
    traj = np.zeros(NTRAJ)
    traj[2] = height
    traj[18:30] = feet.ravel()
    return traj, np.ones(4, np.uint8)


def make_gait_plan(robot="mini_cheetah", combo="walk", total_duration=5.0, goal=(1.5, 0.0), swing_height=0.05, yaw_goal=0.0,
                   sample_dt=0.0, wait_time=0.0, base_height=None):
    """Synthetic stand-in for the IPOPT solution in TOWR's variable layout: base nodes every 0.1 s on a straight line at
    constant velocity (SetByLinearInterpolation, nodes_variables.cc:127-149); per foot one constant polynomial per stance
    phase and two per swing phase with a lifted apex node (vertical velocity 0, nodes_variables_phase_based.cc:208-216);
    three force polynomials per stance phase carrying a quarter of the weight (nlp_formulation.cc:150-170)."""
    x, y, z, mass = NOMINAL_STANCE[robot]
    stance = np.array([[x, y, z], [x, -y, z], [-x, y, z], [-x, -y, z]])
    T = float(total_duration)
    bh = -z if base_height is None else float(base_height)
    p0, p1 = np.array([0.0, 0.0, bh]), np.array([goal[0], goal[1], bh])
    dp = p1 - p0
    bd = base_poly_durations(T)
    nb = len(bd) + 1
    lin, ang = np.zeros((nb, 6)), np.zeros((nb, 6))
    frac = np.arange(nb) / float(nb - 1)
    lin[:, :3] = p0 + frac[:, None] * dp
    lin[:, 3:] = dp / T
    ang[:, 2] = frac * yaw_goal
    ang[:, 5] = yaw_goal / T
    motions, forces, pds, c0s = [], [], [], []
    for ee, (pd, c0) in enumerate(gait_phase_durations(combo, T)):
        starts = np.concatenate([[0.0], np.cumsum(pd)])
        is_stance = [c0 if ph % 2 == 0 else (not c0) for ph in range(len(pd))]
        holds = {}
        for ph, st in enumerate(is_stance):
            if st:
                tm = 0.0 if ph == 0 else 0.5 * (starts[ph] + starts[ph + 1])
                b = p0 + min(max(tm / T, 0.0), 1.0) * dp
                holds[ph] = np.array([b[0] + stance[ee, 0], b[1] + stance[ee, 1], 0.0])
        md = split_phases(pd, c0, 2)
        nodes = np.zeros((len(md) + 1, 6))
        k = 0
        for ph, st in enumerate(is_stance):
            if st:
                nodes[k, :3] = nodes[k + 1, :3] = holds[ph]
                k += 1
            else:
                a = holds.get(ph - 1, holds.get(ph + 1))
                b = holds.get(ph + 1, holds.get(ph - 1))
                nodes[k, :3], nodes[k + 2, :3] = a, b
                nodes[k + 1, :3] = 0.5 * (a + b)
                nodes[k + 1, 2] = swing_height
                nodes[k + 1, 3:5] = (b - a)[:2] / pd[ph]
                k += 2
        motions.append(SplineTable(md, nodes))
        fd = split_phases(pd, not c0, 3)
        fn = np.zeros((len(fd) + 1, 6))
        k = 0
        for ph, st in enumerate(is_stance):
            if st:
                for j in range(4):
                    edge = (j == 0 and ph > 0) or (j == 3 and ph < len(pd) - 1)
                    fn[k + j, 2] = 0.0 if edge else mass * 9.81 / 4.0
                k += 3
            else:
                k += 1
        forces.append(SplineTable(fd, fn))
        pds.append(pd)
        c0s.append(c0)
    return GaitPlan(SplineTable(bd, lin), SplineTable(bd, ang), motions, forces, pds, c0s, sample_dt=sample_dt, wait_time=wait_time,
                    standing=simple_standing(robot)[0], total_duration=T)


def make_motion_plan(robot="mini_cheetah", motion="standing", total_duration=6.0, base_height=None, phase=0.0):
    """The reference's manual test motions (planners/simple.py:87-115) as a plan for the device sampler: the analytic base
    reference is sampled into Hermite nodes every 0.1 s (position and velocity exact at the nodes, O(h^4) in between).
    motion: "standing" (SimpleStanding), "orientation" (OrientationTest), "edge" (EdgeTest: constant trunk target offset
    (-0.1, 0.63, 0), friction rows active), "raise_foot" (RaiseFoot: the reference steps the RF target up by 0.1 m at t = 1 s;
    here it is lifted smoothly over the first half of its swing phase), "heave" (not in the reference: p_body z += 0.1 sin t,
    a vertical exercise used by the rollout tests).
    `phase` shifts the time argument of the sinusoids (different robots of a batch at different phases)."""
    feet, height = SIMPLE_STANDING[robot]
    bh = height if base_height is None else float(base_height)
    T = float(total_duration)
    bd = base_poly_durations(T)
    tn = np.concatenate([[0.0], np.cumsum(bd)]) + phase
    nb = len(tn)
    lin, ang = np.zeros((nb, 6)), np.zeros((nb, 6))
    lin[:, 2] = bh
    if motion == "orientation":           # rpy = [0, 0.4 sin t, 0.4 cos t]
        ang[:, 1], ang[:, 2], ang[:, 4], ang[:, 5] = 0.4 * np.sin(tn), 0.4 * np.cos(tn), 0.4 * np.cos(tn), -0.4 * np.sin(tn)
    elif motion == "edge":                # p_body += [-0.1, 0.63, 0]  (planners/simple.py:109-115)
        lin[:, 0], lin[:, 1] = -0.1, 0.63
    elif motion == "heave":               # p_body z += 0.1 sin t
        lin[:, 2] += 0.1 * np.sin(tn)
        lin[:, 5] = 0.1 * np.cos(tn)
    elif motion == "raise_foot":          # p_body += [-0.1, 0.05, 0]
        lin[:, 0], lin[:, 1] = -0.1, 0.05
    elif motion != "standing":
        raise KeyError(motion)
    motions, forces, pds, c0s = [], [], [], []
    mass = NOMINAL_STANCE[robot][3]
    for ee in range(4):
        if motion == "raise_foot" and ee == 1:       # RF: stance for 1 s, then in the air at +0.1 m
            pd = [1.0, T - 1.0]
            md = split_phases(pd, True, 2)
            nodes = np.zeros((4, 6))
            nodes[:, :3] = feet[ee]
            nodes[2:, 2] = 0.1
            fd = split_phases(pd, False, 3)
            fn = np.zeros((5, 6))
            fn[:3, 2] = mass * 9.81 / 4.0
        else:
            pd = [T]
            md, fd = [T], split_phases(pd, False, 3)
            nodes = np.zeros((2, 6))
            nodes[:, :3] = feet[ee]
            fn = np.zeros((4, 6))
            fn[:, 2] = mass * 9.81 / 4.0
        motions.append(SplineTable(md, nodes))
        forces.append(SplineTable(fd, fn))
        pds.append(pd)
        c0s.append(True)
    return GaitPlan(SplineTable(bd, lin), SplineTable(bd, ang), motions, forces, pds, c0s, standing=simple_standing(robot)[0], total_duration=T)


# ------------------------------------------------------------------------------ C ABI mirrors
class _SplineDesc(C.Structure):
    _fields_ = [("n_poly", C.c_int32), ("durations", C.c_void_p), ("nodes", C.c_void_p)]


class _PlanDesc(C.Structure):
    _fields_ = [("base_linear", _SplineDesc), ("base_angular", _SplineDesc), ("ee_motion", _SplineDesc * 4), ("ee_force", _SplineDesc * 4),
                ("n_phase", C.c_int32 * 4), ("phase_durations", C.c_void_p * 4), ("contact_at_start", C.c_uint8 * 4),
                ("n_grid", C.c_int32), ("grid_timestamps", C.c_void_p), ("wait_time", C.c_double), ("standing", C.c_double * NTRAJ)]


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class TrajectorySampler:
    """Device-resident set of plans + the sampling entry (wbc_plan_create / wbc_sample_trajectory)."""

    def __init__(self, ctl, plans):
        self.ctl, self.lib = ctl, ctl.lib
        self.plans = list(plans) if isinstance(plans, (list, tuple)) else [plans]
        descs = (_PlanDesc * len(self.plans))()
        for d, p in zip(descs, self.plans):
            def sd(dst, s):
                dst.n_poly, dst.durations, dst.nodes = len(s.durations), s.durations.ctypes.data, s.nodes.ctypes.data
            sd(d.base_linear, p.base_linear)
            sd(d.base_angular, p.base_angular)
            for k in range(4):
                sd(d.ee_motion[k], p.ee_motion[k])
                sd(d.ee_force[k], p.ee_force[k])
                d.n_phase[k] = len(p.phase_durations[k])
                d.phase_durations[k] = p.phase_durations[k].ctypes.data
                d.contact_at_start[k] = 1 if p.contact_at_start[k] else 0
            d.n_grid = len(p.grid)
            d.grid_timestamps = p.grid.ctypes.data if len(p.grid) else None
            d.wait_time = p.wait_time
            for i in range(NTRAJ):
                d.standing[i] = p.standing[i]
        self._p = C.c_void_p()
        fn = self.lib.wbc_plan_create
        fn.argtypes = [C.c_void_p, C.c_int32, C.POINTER(_PlanDesc), C.POINTER(C.c_void_p)]
        fn.restype = C.c_int
        self.ctl._check(fn(self.ctl._h, len(self.plans), descs, C.byref(self._p)), "wbc_plan_create")

    def close(self):
        if getattr(self, "_p", None):
            self.lib.wbc_plan_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sample(self, t, plan_index=None, forces=False):
        """t[N] (+ plan_index[N]) -> dict(traj[N,54], contact[N,4], f[N,12] or None, t_eval[N], status[N])."""
        if _is_torch(t):
            import torch
            n, dev = t.shape[0], t.device
            o = dict(traj=torch.empty((n, NTRAJ), dtype=torch.float64, device=dev), contact=torch.empty((n, 4), dtype=torch.uint8, device=dev),
                     f=torch.empty((n, 12), dtype=torch.float64, device=dev) if forces else None,
                     t_eval=torch.empty(n, dtype=torch.float64, device=dev), status=torch.empty(n, dtype=torch.int32, device=dev))
            p = lambda x: None if x is None else C.c_void_p(x.data_ptr())  # noqa: E731
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self.ctl._check(self.lib.wbc_sample_trajectory(self.ctl._h, self._p, n, p(plan_index), p(t), p(o["traj"]), p(o["contact"]), p(o["f"]),
                                                           p(o["t_eval"]), p(o["status"]), stream), "wbc_sample_trajectory")
            return o
        t = np.ascontiguousarray(t, dtype=np.float64).ravel()
        n = len(t)
        pi = None if plan_index is None else np.ascontiguousarray(plan_index, dtype=np.int32).reshape(n)
        o = dict(traj=np.empty((n, NTRAJ)), contact=np.empty((n, 4), np.uint8), f=np.empty((n, 12)) if forces else None,
                 t_eval=np.empty(n), status=np.empty(n, np.int32))
        opt = lambda a: None if a is None else np_ptr(a)  # noqa: E731
        self.ctl._check(self.lib.wbc_sample_trajectory_host(self.ctl._h, self._p, n, opt(pi), np_ptr(t), np_ptr(o["traj"]), np_ptr(o["contact"]),
                                                            opt(o["f"]), np_ptr(o["t_eval"]), np_ptr(o["status"])), "wbc_sample_trajectory_host")
        return o


# ------------------------------------------------------------------------------ planner mirrors
class BasicTrunkPlanner:
    """Drop-in for reference planners/simple.py:BasicTrunkPlanner: output "trunk_trajectory" = the standing dict."""

    def __init__(self, frame_ids=None, robot="mini_cheetah"):
        self.frame_ids, self.robot = frame_ids, robot
        self.output_dict = {}

    def SimpleStanding(self):
        traj, contact = simple_standing(self.robot)
        self.output_dict = traj_to_dict(traj, contact)

    def OrientationTest(self, t):          # planners/simple.py:87-95
        self.SimpleStanding()
        self.output_dict["rpy_body"] = np.array([0.0, 0.4 * np.sin(t), 0.4 * np.cos(t)])
        self.output_dict["rpyd_body"] = np.array([0.0, 0.4 * np.cos(t), -0.4 * np.sin(t)])
        self.output_dict["rpydd_body"] = np.array([0.0, -0.4 * np.sin(t), -0.4 * np.cos(t)])

    def RaiseFoot(self, t):                # planners/simple.py:97-107
        self.SimpleStanding()
        self.output_dict["p_body"] += np.array([-0.1, 0.05, 0.0])
        if t > 1:
            self.output_dict["contact_states"] = [True, False, True, True]
            self.output_dict["p_rf"] += np.array([0.0, 0.0, 0.1])

    def EdgeTest(self, t=None):            # planners/simple.py:109-115: trunk moved to the edge of feasibility (friction rows active)
        self.SimpleStanding()
        self.output_dict["p_body"] += np.array([-0.1, 0.63, 0.0])

    def SetTrunkOutputs(self, t):          # planners/simple.py:117-124
        self.SimpleStanding()
        return self.output_dict


class TowrTrunkPlanner:
    """Drop-in for reference planners/towr.py:TowrTrunkPlanner with the spline solution held on the device.
    `SetTrunkOutputs(t)` returns the reference dict for one time; `sample(t[N])` is the batched device entry."""

    def __init__(self, ctl, plan=None, robot="mini_cheetah", gait="walk", distance=(1.5, 0.0), wait_time=1.0, total_duration=5.0):
        # planners/towr.py:60 runs `trunk_mpc walk 0 1.5 0.0`; towr/trunk_mpc.cpp:126,168: 5 s sampled at 1 kHz
        self.plan = plan or make_gait_plan(robot, gait, total_duration, distance, sample_dt=1e-3, wait_time=wait_time)
        self.sampler = TrajectorySampler(ctl, self.plan)
        self.wait_time = self.plan.wait_time
        self.output_dict = {}
        self.u2_max = self.ComputeMaxControlInputs()

    def ComputeMaxControlInputs(self):
        """planners/towr.py:70-90: max over the stored samples of |[foot accelerations; base rpydd; base pdd]|_2."""
        g = self.plan.grid
        if g is None or len(g) == 0:
            return 0.0
        o = self.sampler.sample(np.asarray(g, float) + self.wait_time)
        tr = o["traj"]
        u2 = np.concatenate([tr[:, 42:54], tr[:, 15:18], tr[:, 6:9]], axis=1)
        return float(np.linalg.norm(u2, axis=1).max())

    def sample(self, t, forces=False):
        return self.sampler.sample(t, forces=forces)

    def SetTrunkOutputs(self, t):
        o = self.sampler.sample(np.array([float(t)]), forces=True)
        # u2_max is only attached once the motion has started (planners/towr.py:148; SimpleStanding leaves it at 0)
        self.output_dict = traj_to_dict(o["traj"][0], o["contact"][0], f_plan=o["f"][0], u2_max=self.u2_max if t >= self.wait_time else 0.0)
        return self.output_dict
