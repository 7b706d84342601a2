"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the reference's trunk-trajectory *sampling* path: what turns a TOWR spline
solution into the 54-double / 4-bool controller input at time t. The NLP solve itself (IPOPT) is out of scope.

  gait tables          towr/src/quadruped_gait_generator.cc:38-375   (contact sequences and phase times)
  phase durations      towr/src/gait_generator.cc:55-120,130-143
  polynomial durations towr/src/nodes_variables_phase_based.cc:36-83, towr/src/parameters.cc:40-45,83-98
  cubic Hermite        towr/src/polynomial.cc:49-63 (GetPoint), :98-104 (UpdateCoeff)
  spline lookup        towr/src/spline.cc:49-90
  contact flag         towr/src/phase_durations.cc:120-124
  sample layout        towr/trunk_mpc.cpp:19-68, :168-174 (1 kHz, accumulated timestamps, final `finished` sample)
  planner lookup       planners/towr.py:92-148 (stand for 1 s, then NEAREST stored sample, no interpolation),
                       planners/simple.py:39-85 (SimpleStanding)

Pinned against the reference's own sources compiled here (oracle/_ref/libtowr_ref.so, see oracle/ref_build/) by
tests/test_oracle_trajectory.py, and frozen in tests/golden/towr_spline.npz (tools/make_golden_traj.py).
The synthetic node generator `make_gait_plan` is the builder's (it stands in for the IPOPT solution); only its use
of the tables above follows the reference (initialisation style of towr/src/nlp_formulation.cc:94-170).
"""
from __future__ import annotations

import numpy as np

LF, RF, LH, RH = 0, 1, 2, 3           # towr/include/towr/models/endeffector_mappings.h:44


def _cs(*feet):
    c = [False] * 4
    for f in feet:
        c[f] = True
    return tuple(c)


# contact states, quadruped_gait_generator.cc:38-73 (capital = left, small = right; first letter hind, second front)
II = _cs()
PI, bI, IP, Ib = _cs(LH), _cs(RH), _cs(LF), _cs(RF)
Pb, bP, BI, IB, PP, bb = _cs(LH, RF), _cs(RH, LF), _cs(LH, RH), _cs(LF, RF), _cs(LH, LF), _cs(RH, RF)
Bb, BP, bB, PB = _cs(LH, RH, RF), _cs(LH, RH, LF), _cs(RH, LF, RF), _cs(LH, LF, RF)
BB = _cs(LF, RF, LH, RH)


def _remove_transition(g):               # gait_generator.cc:130-143
    t, c = list(g[0]), list(g[1])
    last = t.pop()
    t[-1] += last
    c.pop()
    return t, c


def _strides():                          # quadruped_gait_generator.cc:114-375
    s = {}
    s["Stand"] = ([0.3], [BB])
    s["Flight"] = ([0.3], [Bb])
    s["Hop2"] = ([0.3, 0.4, 0.3], [BB, II, BB])
    s["Walk1"] = ([0.3, 0.2] * 4, [bB, BB, Bb, BB, PB, BB, BP, BB])
    s["Walk2"] = ([0.25, 0.13, 0.25, 0.13, 0.25, 0.13, 0.25, 0.13], [bB, bb, Bb, Pb, PB, PP, BP, bP])
    s["Walk2E"] = _remove_transition(s["Walk2"])
    s["Run1"] = ([0.3, 0.2, 0.3, 0.2], [bP, BB, Pb, BB])
    s["Run2"] = ([0.4, 0.1, 0.4, 0.1], [bP, II, Pb, II])
    s["Run2E"] = ([0.4], [bP])
    s["Run3"] = ([0.3, 0.1, 0.3, 0.1], [PP, II, bb, II])
    s["Run3E"] = ([0.3], [PP])
    s["Hop1"] = ([0.3, 0.1, 0.3, 0.1], [BI, II, IB, II])
    s["Hop1E"] = ([0.3], [BI])
    A, B, C = 0.3, 0.2, 0.2
    s["Hop3"] = ([B, A, B, C, B, A, B, C], [Bb, BI, BP, bP, bB, IB, PB, Pb])
    s["Hop3E"] = _remove_transition(s["Hop3"])
    A, B, C = 0.1, 0.2, 0.1
    s["Hop5"] = ([A, B, C, A, B, C], [Bb, BB, IP, Bb, BB, IP])
    return s


STRIDES = _strides()
COMBOS = {                                # QuadrupedGaitGenerator::SetCombo, quadruped_gait_generator.cc:76-88
    0: ["Stand", "Walk2", "Walk2", "Walk2", "Walk2E", "Stand"],     # overlap-walk (trunk_mpc "walk")
    1: ["Stand", "Run2", "Run2", "Run2", "Run2E", "Stand"],         # flying trot  (trunk_mpc "trot")
    2: ["Stand", "Run3", "Run3", "Run3", "Run3E", "Stand"],         # pace
    3: ["Stand", "Hop1", "Hop1", "Hop1", "Hop1E", "Stand"],         # bound
    4: ["Stand", "Hop3", "Hop3", "Hop3", "Hop3E", "Stand"],         # gallop
}
COMBO_NAMES = {"walk": 0, "trot": 1, "pace": 2, "bound": 3, "gallop": 4}      # towr/trunk_mpc.cpp:82-98


def gait_sequence(combo):
    """SetGaits (gait_generator.cc:113-128): concatenated (times, contact states) of the combo's strides."""
    times, contacts = [], []
    for g in COMBOS[combo]:
        t, c = STRIDES[g]
        times += list(t)
        contacts += list(c)
    return times, contacts


def foot_phase_durations(combo):
    """GaitGenerator::GetPhaseDurations() (gait_generator.cc:77-105): per foot, the durations of its alternating
    contact / swing phases (unnormalised)."""
    times, contacts = gait_sequence(combo)
    acc = [0.0] * 4
    out = [[] for _ in range(4)]
    for ph in range(len(contacts) - 1):
        for ee in range(4):
            acc[ee] += times[ph]
            if contacts[ph][ee] != contacts[ph + 1][ee]:
                out[ee].append(acc[ee])
                acc[ee] = 0.0
    for ee in range(4):
        out[ee].append(acc[ee] + times[-1])
    return out


def phase_durations(combo, t_total, ee):
    """GetPhaseDurations(t_total, ee) (gait_generator.cc:55-75): normalised by the foot's total, scaled to t_total."""
    v = foot_phase_durations(combo)[ee]
    total = 0.0
    for x in v:
        total += x                        # std::accumulate
    return [(x / total) * t_total for x in v]


def contact_at_start(combo, ee):
    return bool(gait_sequence(combo)[1][0][ee])


def base_poly_durations(t_total, dt=0.1):
    """Parameters::GetBasePolyDurations (parameters.cc:83-98)."""
    out, t_left = [], t_total
    while t_left > 1e-10:
        out.append(dt if t_left > dt else t_left)
        t_left -= dt
    return out


def phase_to_poly_durations(phase_dur, first_phase_constant, n_polys_in_changing_phase):
    """BuildPolyInfos + ConvertPhaseToPolyDurations (nodes_variables_phase_based.cc:36-83): a constant phase is one
    polynomial, a changing phase is split evenly into n. Returns (poly durations, phase index of each poly)."""
    durs, phase_of = [], []
    const = first_phase_constant
    for i, d in enumerate(phase_dur):
        n = 1 if const else n_polys_in_changing_phase
        for _ in range(n):
            durs.append(d / n)
            phase_of.append(i)
        const = not const
    return durs, phase_of


# ------------------------------------------------------------------------------ splines
def segment_id(t_global, durations):
    """Spline::GetSegmentID (spline.cc:49-66): at junctions returns the previous polynomial."""
    eps = 1e-10
    t = 0.0
    for i, d in enumerate(durations):
        t += d
        if t >= t_global - eps:
            return i
    raise AssertionError("t beyond the spline")            # reference: assert(false)


def local_time(t_global, durations):
    """Spline::GetLocalTime (spline.cc:68-79): sequential subtraction of the previous durations."""
    i = segment_id(t_global, durations)
    tl = t_global
    for k in range(i):
        tl -= durations[k]
    return i, tl


def hermite_coeff(p0, v0, p1, v1, T):
    """CubicHermitePolynomial::UpdateCoeff (polynomial.cc:98-104)."""
    p0, v0, p1, v1 = (np.asarray(x, float) for x in (p0, v0, p1, v1))
    a = p0
    b = v0
    c = -(3 * (p0 - p1) + T * (2 * v0 + v1)) / T ** 2
    d = (2 * (p0 - p1) + T * (v0 + v1)) / T ** 3
    return a, b, c, d


def poly_point(coeff, t):
    """Polynomial::GetPoint (polynomial.cc:49-63): sum over the coefficients of d^k/dt^k t^c."""
    a, b, c, d = coeff
    p = a + t * b + t ** 2 * c + t ** 3 * d
    v = b + 2 * t * c + 3 * t ** 2 * d
    acc = 2 * c + 6 * t * d
    return p, v, acc


class Spline:
    """3-D cubic Hermite spline over nodes[(n_poly + 1), 6] = (p, v) (towr NodeSpline / Spline)."""

    def __init__(self, durations, nodes):
        self.durations = [float(x) for x in durations]
        self.nodes = np.asarray(nodes, float).reshape(len(self.durations) + 1, 6)

    def point(self, t):
        i, tl = local_time(t, self.durations)
        n0, n1 = self.nodes[i], self.nodes[i + 1]
        return poly_point(hermite_coeff(n0[:3], n0[3:], n1[:3], n1[3:], self.durations[i]), tl)

    def total_time(self):
        s = 0.0
        for d in self.durations:
            s += d
        return s


def is_contact_phase(t, phase_dur, in_contact_at_start):
    """PhaseDurations::IsContactPhase (phase_durations.cc:120-124)."""
    ph = segment_id(t, phase_dur)
    return in_contact_at_start if ph % 2 == 0 else (not in_contact_at_start)


class Plan:
    """The SplineHolder of a solved (here: synthetic) trunk trajectory (towr/include/towr/variables/spline_holder.h):
    base_linear, base_angular, ee_motion[4], ee_force[4] splines + per-foot phase durations."""

    def __init__(self, base_linear, base_angular, ee_motion, ee_force, phase_dur, contact_start):
        self.base_linear, self.base_angular, self.ee_motion, self.ee_force = base_linear, base_angular, ee_motion, ee_force
        self.phase_dur, self.contact_start = phase_dur, [bool(c) for c in contact_start]

    def total_time(self):
        return self.base_linear.total_time()

    def sample(self, t):
        """publish_trunk_state (towr/trunk_mpc.cpp:19-68): (traj[54] in the wbc.h order, contact[4], f[12])."""
        traj = np.zeros(54)
        p, v, a = self.base_linear.point(t)
        traj[0:3], traj[3:6], traj[6:9] = p, v, a
        p, v, a = self.base_angular.point(t)
        traj[9:12], traj[12:15], traj[15:18] = p, v, a
        f = np.zeros(12)
        contact = np.zeros(4, np.uint8)
        for ee in range(4):
            p, v, a = self.ee_motion[ee].point(t)
            traj[18 + 3 * ee:21 + 3 * ee], traj[30 + 3 * ee:33 + 3 * ee], traj[42 + 3 * ee:45 + 3 * ee] = p, v, a
            contact[ee] = is_contact_phase(t, self.phase_dur[ee], self.contact_start[ee])
            f[3 * ee:3 * ee + 3] = self.ee_force[ee].point(t)[0]
        return traj, contact, f


def publish_timestamps(total_duration, dt=1e-3):
    """The sample times trunk_mpc sends (trunk_mpc.cpp:168-174): `for (t = 0; t < T; t = t + dt)` with the
    accumulated floating-point t, then one final sample at exactly T carrying finished = true."""
    ts, t = [], 0.0
    while t < total_duration:
        ts.append(t)
        t = t + dt
    ts.append(total_duration)
    return np.array(ts)


SIMPLE_STANDING = {                       # planners/simple.py:39-85 (mini cheetah literals)
    "p_feet": np.array([[0.175, 0.11, 0.0], [0.175, -0.11, 0.0], [-0.2, 0.11, 0.0], [-0.2, -0.11, 0.0]]),
    "p_body": np.array([0.0, 0.0, 0.3]),
}


def simple_standing_traj():
    traj = np.zeros(54)
    traj[0:3] = SIMPLE_STANDING["p_body"]
    traj[18:30] = SIMPLE_STANDING["p_feet"].ravel()
    return traj, np.ones(4, np.uint8)


def towr_planner_output(plan, timestamps, t, wait_time=1.0):
    """TowrTrunkPlanner.SetTrunkOutputs (planners/towr.py:92-148): SimpleStanding for t < wait_time, afterwards the
    stored sample whose timestamp is closest to t - wait_time (np.abs(ts - t).argmin(): first minimum wins)."""
    if t < wait_time:
        return simple_standing_traj() + (np.zeros(12),)
    idx = int(np.abs(np.asarray(timestamps) - (t - wait_time)).argmin())
    return plan.sample(float(timestamps[idx]))


# ------------------------------------------------------------------------------ synthetic plan (builder's generator)
NOMINAL_STANCE = {                        # towr/include/towr/models/examples/{mini_cheetah,anymal}_model.h
    "mini_cheetah": (0.2, 0.11, -0.30, 9.0),
    "anymal_b": (0.34, 0.19, -0.42, 29.5),
}


def make_gait_plan(robot="mini_cheetah", combo=0, total_duration=5.0, goal=(1.5, 0.0), swing_height=0.05, yaw_goal=0.0, base_height=None):
    """Synthetic stand-in for the IPOPT solution: same variable layout as TOWR (base nodes every 0.1 s, one constant
    polynomial per stance phase, two per swing phase with a lifted mid node whose vertical velocity is zero, three force
    polynomials per stance phase), node values from a simple heuristic: base on a straight line with constant velocity
    (SetByLinearInterpolation, nodes_variables.cc:127-149), footholds under the base at the middle of each stance phase,
    weight shared equally by the stance feet (nlp_formulation.cc:150-170)."""
    x, y, z, mass = NOMINAL_STANCE[robot]
    stance = np.array([[x, y, z], [x, -y, z], [-x, y, z], [-x, -y, z]])
    T = float(total_duration)
    bh = -z if base_height is None else float(base_height)
    p_init, p_goal = np.array([0.0, 0.0, bh]), np.array([goal[0], goal[1], bh])
    bd = base_poly_durations(T)
    nb = len(bd) + 1
    dp = p_goal - p_init
    lin = np.zeros((nb, 6))
    ang = np.zeros((nb, 6))
    for i in range(nb):
        lin[i, :3] = p_init + i / float(nb - 1) * dp
        lin[i, 3:] = dp / T
        ang[i, 2] = i / float(nb - 1) * yaw_goal
        ang[i, 5] = yaw_goal / T
    base_linear, base_angular = Spline(bd, lin), Spline(bd, ang)

    def base_xy(t):
        return p_init + min(max(t / T, 0.0), 1.0) * dp

    ee_motion, ee_force, pds, cstart = [], [], [], []
    for ee in range(4):
        pd = phase_durations(combo, T, ee)
        c0 = contact_at_start(combo, ee)
        starts = np.concatenate([[0.0], np.cumsum(pd)])
        # footholds: nominal stance under the base at the middle of the stance phase (first one at the start)
        holds = {}
        for ph in range(len(pd)):
            stance_ph = c0 if ph % 2 == 0 else (not c0)
            if stance_ph:
                tm = 0.0 if ph == 0 else 0.5 * (starts[ph] + starts[ph + 1])
                b = base_xy(tm)
                holds[ph] = np.array([b[0] + stance[ee, 0], b[1] + stance[ee, 1], 0.0])
        md, mphase = phase_to_poly_durations(pd, c0, 2)
        nodes = np.zeros((len(md) + 1, 6))
        k = 0
        for ph in range(len(pd)):
            stance_ph = c0 if ph % 2 == 0 else (not c0)
            prev_hold = holds.get(ph - 1, holds.get(ph + 1))
            next_hold = holds.get(ph + 1, holds.get(ph - 1))
            if stance_ph:
                nodes[k, :3] = holds[ph]; nodes[k + 1, :3] = holds[ph]
                k += 1
            else:
                if prev_hold is None:                       # swing with no neighbouring stance (cannot happen in the combos)
                    prev_hold = next_hold = np.array([stance[ee, 0], stance[ee, 1], 0.0])
                nodes[k, :3] = prev_hold
                mid = 0.5 * (prev_hold + next_hold)
                mid[2] = swing_height
                nodes[k + 1, :3] = mid
                nodes[k + 1, 3:5] = (next_hold - prev_hold)[:2] / pd[ph]   # horizontal velocity, vz = 0 at the apex
                nodes[k + 2, :3] = next_hold
                k += 2
        ee_motion.append(Spline(md, nodes))
        fd, fphase = phase_to_poly_durations(pd, not c0, 3)
        fn = np.zeros((len(fd) + 1, 6))
        k = 0
        for ph in range(len(pd)):
            stance_ph = c0 if ph % 2 == 0 else (not c0)
            if stance_ph:
                for j in range(4):
                    edge = (j == 0 and ph > 0) or (j == 3 and ph < len(pd) - 1)
                    fn[k + j, 2] = 0.0 if edge else mass * 9.81 / 4.0       # zero at touch-down / lift-off
                k += 3
            else:
                k += 1
        ee_force.append(Spline(fd, fn))
        pds.append(pd)
        cstart.append(c0)
    return Plan(base_linear, base_angular, ee_motion, ee_force, pds, cstart)
