#!/usr/bin/env python
"""Generate tests/golden/lcm_wire.npz from the REFERENCE's own generated LCM codecs.

The generated Python codecs (lcm_types/trunklcm/trunk_state_t.py, lcm_types/cheetahlcm/robot_state_control_lcmt.py)
need only `struct`, so they import in this container even though the LCM runtime is absent. This script runs them on
seeded field values (plus special values: +-0, denormals, +-inf, extremes, false/true booleans) and freezes the wire
bytes and the decoded fields. The GPU box has no /root/reference: tests read only the committed .npz.

Usage: python tools/make_golden_lcm.py          (needs /root/reference)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference/lcm_types")
from trunklcm.trunk_state_t import trunk_state_t  # noqa: E402
from cheetahlcm.robot_state_control_lcmt import robot_state_control_lcmt  # noqa: E402

TRUNK_FIELDS = ["base_p", "base_pd", "base_pdd", "base_rpy", "base_rpyd", "base_rpydd", "lf_p", "rf_p", "lh_p", "rh_p",
                "lf_pd", "rf_pd", "lh_pd", "rh_pd", "lf_pdd", "rf_pdd", "lh_pdd", "rh_pdd"]          # = traj[54] order
FORCE_FIELDS = ["lf_f", "rf_f", "lh_f", "rh_f"]
CONTACT_FIELDS = ["lf_contact", "rf_contact", "lh_contact", "rh_contact"]
SPECIAL64 = [0.0, -0.0, 5e-324, -2.2250738585072014e-308, 1.7976931348623157e308, float("inf"), float("-inf"), 1.0, -1.0,
             np.pi, 1e-300, 123456789.123456789]
SPECIAL32 = [0.0, -0.0, 1e-45, -1.1754944e-38, 3.4028234e38, float("inf"), float("-inf"), 1.0, -1.0, 0.1, 1e-7,
             16777217.0, 3.0000001, 2.5000000001, 1e-50, -1e-50]   # incl. round-to-nearest-even and underflow cases


def main():
    rng = np.random.default_rng(20260117 + 100)
    n = 96
    ts = rng.uniform(0, 5, n)
    fin = rng.integers(0, 2, n).astype(np.uint8)
    traj = rng.normal(0, 1, (n, 54)) * np.exp(rng.uniform(-20, 20, (n, 54)))
    f = rng.normal(0, 50, (n, 12))
    contact = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    sp = np.array(SPECIAL64)
    traj[0, :len(sp)] = sp
    f[1, :len(sp)] = sp
    ts[2] = -0.0
    msgs = np.zeros((n, 549), np.uint8)
    for i in range(n):
        m = trunk_state_t()
        m.timestamp, m.finished = float(ts[i]), bool(fin[i])
        for k, name in enumerate(TRUNK_FIELDS):
            setattr(m, name, [float(x) for x in traj[i, 3 * k:3 * k + 3]])
        for k, name in enumerate(FORCE_FIELDS):
            setattr(m, name, [float(x) for x in f[i, 3 * k:3 * k + 3]])
        for k, name in enumerate(CONTACT_FIELDS):
            setattr(m, name, bool(contact[i, k]))
        b = m.encode()
        assert len(b) == 549
        msgs[i] = np.frombuffer(b, np.uint8)
        d = trunk_state_t.decode(b)                    # the reference decoder round-trips its own bytes
        assert d.timestamp == ts[i] or (np.isnan(d.timestamp) and np.isnan(ts[i]))
        back = np.concatenate([getattr(d, name) for name in TRUNK_FIELDS])
        assert np.array_equal(back.view(np.uint64), traj[i].view(np.uint64))
        assert [d.lf_contact, d.rf_contact, d.lh_contact, d.rh_contact] == [bool(x) for x in contact[i]]

    # robot_state_control_lcmt: float64 inputs (what the controller holds) -> float32 wire values
    nr = 96
    q = rng.normal(0, 1, (nr, 19)); v = rng.normal(0, 3, (nr, 18)); tau = rng.normal(0, 20, (nr, 12))
    s32 = np.array(SPECIAL32)
    q[0, :len(s32)] = s32
    tau[1, :12] = s32[:12]
    v[2, :len(s32)] = s32
    rmsgs = np.zeros((nr, 204), np.uint8)
    dq, dv, dtau = np.zeros((nr, 19)), np.zeros((nr, 18)), np.zeros((nr, 12))
    for i in range(nr):
        m = robot_state_control_lcmt()
        m.q, m.v, m.tau = [float(x) for x in q[i]], [float(x) for x in v[i]], [float(x) for x in tau[i]]
        b = m.encode()
        assert len(b) == 204
        rmsgs[i] = np.frombuffer(b, np.uint8)
        d = robot_state_control_lcmt.decode(b)
        dq[i], dv[i], dtau[i] = d.q, d.v, d.tau
    # the controller's own outgoing message: fresh message, torques only (basic_controller.py:309-314)
    tmsgs = np.zeros((nr, 204), np.uint8)
    for i in range(nr):
        m = robot_state_control_lcmt()
        m.tau = [float(x) for x in tau[i]]
        tmsgs[i] = np.frombuffer(m.encode(), np.uint8)
    # what the reference does on a too-large double: struct.pack('>f') raises OverflowError
    m = robot_state_control_lcmt(); m.tau = [1e39] + [0.0] * 11
    try:
        m.encode(); overflow_raises = False
    except OverflowError:
        overflow_raises = True
    # and on a corrupted fingerprint: ValueError("Decode error")
    bad = bytearray(tmsgs[0].tobytes()); bad[3] ^= 0x10
    try:
        robot_state_control_lcmt.decode(bytes(bad)); bad_raises = False
    except ValueError:
        bad_raises = True
    out = ROOT / "tests" / "golden" / "lcm_wire.npz"
    np.savez_compressed(out, trunk_msgs=msgs, trunk_timestamp=ts, trunk_finished=fin, trunk_traj=traj, trunk_contact=contact,
                        trunk_f=f, robot_msgs=rmsgs, robot_q_in=q, robot_v_in=v, robot_tau_in=tau, robot_q=dq, robot_v=dv,
                        robot_tau=dtau, robot_tau_only_msgs=tmsgs, overflow_raises=overflow_raises, bad_fingerprint_raises=bad_raises,
                        trunk_fingerprint=np.frombuffer(trunk_state_t._get_packed_fingerprint(), np.uint8),
                        robot_fingerprint=np.frombuffer(robot_state_control_lcmt._get_packed_fingerprint(), np.uint8))
    print("wrote", out, "overflow_raises", overflow_raises, "bad_fingerprint_raises", bad_raises)


if __name__ == "__main__":
    main()
