"""oracle/lcm_codec.py against the golden wire vectors produced by the reference's generated LCM codecs
(tools/make_golden_lcm.py): bit-exact in both directions."""
from pathlib import Path

import numpy as np

from oracle import lcm_codec as lc

G = np.load(Path(__file__).parent / "golden" / "lcm_wire.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def test_fingerprints_follow_from_the_schema():
    # lcm-gen struct hash restated from the .lcm member lists reproduces the constants in the generated files
    assert lc.struct_hash(lc.TRUNK_STATE_MEMBERS) == 0xbd03c56c9649d0b6          # trunk_state_t.py:125
    assert lc.struct_hash(lc.ROBOT_STATE_MEMBERS) == 0xbe14089c923ad667          # robot_state_control_lcmt.py:56
    assert bytes(G["trunk_fingerprint"]) == lc.TRUNK_STATE_FINGERPRINT
    assert bytes(G["robot_fingerprint"]) == lc.ROBOT_STATE_FINGERPRINT


def test_trunk_state_decode_matches_reference_fields():
    d = lc.decode_trunk_state(G["trunk_msgs"])
    assert (d["status"] == 0).all()
    assert np.array_equal(bits(d["timestamp"]), bits(G["trunk_timestamp"]))
    assert np.array_equal(d["finished"], G["trunk_finished"])
    assert np.array_equal(bits(d["traj"]), bits(G["trunk_traj"]))
    assert np.array_equal(d["contact"], G["trunk_contact"])
    assert np.array_equal(bits(d["f"]), bits(G["trunk_f"]))


def test_trunk_state_encode_matches_reference_bytes():
    m = lc.encode_trunk_state(G["trunk_timestamp"], G["trunk_finished"], G["trunk_traj"], G["trunk_contact"], G["trunk_f"])
    assert np.array_equal(m, G["trunk_msgs"])


def test_robot_state_roundtrip_matches_reference():
    m = lc.encode_robot_state(G["robot_q_in"], G["robot_v_in"], G["robot_tau_in"])
    assert np.array_equal(m, G["robot_msgs"])
    d = lc.decode_robot_state(G["robot_msgs"])
    for k in ("q", "v", "tau"):
        assert np.array_equal(bits(d[k]), bits(G["robot_" + k]))
    assert np.array_equal(lc.encode_robot_state(None, None, G["robot_tau_in"]), G["robot_tau_only_msgs"])


def test_bad_fingerprint_is_flagged():
    assert bool(G["bad_fingerprint_raises"])                  # the reference decoder raises ValueError
    msgs = G["trunk_msgs"].copy()
    msgs[5, 2] ^= 0x40
    d = lc.decode_trunk_state(msgs)
    assert d["status"][5] == 1 and (np.delete(d["status"], 5) == 0).all()
    assert (d["traj"][5] == 0).all()
    r = G["robot_msgs"].copy()
    r[7, 7] ^= 1
    assert lc.decode_robot_state(r)["status"][7] == 1
