"""GPU parity of the LCM wire codecs (wbc_lcm_*): bit-exact against the golden vectors produced by the reference's
generated codecs (tests/golden/lcm_wire.npz) and against oracle/lcm_codec.py on large seeded batches."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).parent / "golden" / "lcm_wire.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def codec(built):
    from quadruped_drake_b200.controller import BatchedController
    from quadruped_drake_b200.wire import WireCodec
    ctl = BatchedController("mini_cheetah", device=0)
    yield WireCodec(ctl)
    ctl.close()


def test_trunk_state_decode_golden(codec):
    d = codec.decode_trunk_state(G["trunk_msgs"])
    assert (d["status"] == 0).all()
    assert np.array_equal(bits(d["timestamp"]), bits(G["trunk_timestamp"]))
    assert np.array_equal(d["finished"], G["trunk_finished"])
    assert np.array_equal(bits(d["traj"]), bits(G["trunk_traj"]))
    assert np.array_equal(d["contact"], G["trunk_contact"])
    assert np.array_equal(bits(d["f"]), bits(G["trunk_f"]))


def test_trunk_state_encode_golden(codec):
    m = codec.encode_trunk_state(G["trunk_timestamp"], G["trunk_finished"], G["trunk_traj"], G["trunk_contact"], G["trunk_f"])
    assert np.array_equal(m, G["trunk_msgs"])


def test_robot_state_golden(codec):
    m, st = codec.encode_robot_state(G["robot_q_in"], G["robot_v_in"], G["robot_tau_in"])
    assert (st == 0).all() and np.array_equal(m, G["robot_msgs"])
    d = codec.decode_robot_state(G["robot_msgs"])
    assert (d["status"] == 0).all()
    for k in ("q", "v", "tau"):
        assert np.array_equal(bits(d[k]), bits(G["robot_" + k]))
    m, st = codec.encode_robot_state(None, None, G["robot_tau_in"])
    assert np.array_equal(m, G["robot_tau_only_msgs"])


def test_error_statuses(codec):
    from quadruped_drake_b200.wire import WIRE_BADFINGERPRINT, WIRE_OVERFLOW
    msgs = G["trunk_msgs"].copy()
    msgs[5, 2] ^= 0x40
    msgs[40, 7] ^= 0x01
    d = codec.decode_trunk_state(msgs)
    assert d["status"][5] == WIRE_BADFINGERPRINT and d["status"][40] == WIRE_BADFINGERPRINT and (np.delete(d["status"], [5, 40]) == 0).all()
    assert (d["traj"][5] == 0).all() and (d["contact"][40] == 0).all() and d["timestamp"][5] == 0
    r = G["robot_msgs"].copy()
    r[7, 0] ^= 0x80
    dr = codec.decode_robot_state(r)
    assert dr["status"][7] == WIRE_BADFINGERPRINT and (dr["q"][7] == 0).all() and (np.delete(dr["status"], 7) == 0).all()
    assert bool(G["overflow_raises"])            # struct.pack('>f', 1e39) raises in the reference codec
    tau = np.zeros((3, 12)); tau[1, 4] = 1e39; tau[2, 0] = np.inf      # inf itself packs fine in the reference
    m, st = codec.encode_robot_state(None, None, tau)
    assert list(st) == [0, WIRE_OVERFLOW, 0]


@pytest.mark.parametrize("n", [1, 31, 32, 33, 63, 64, 65, 1000, 65536 + 7])
def test_ragged_sizes_match_oracle(codec, n):
    from oracle import lcm_codec as lc
    rng = np.random.default_rng(n)
    ts, fin = rng.uniform(0, 5, n), rng.integers(0, 2, n).astype(np.uint8)
    traj, f = rng.normal(0, 3, (n, 54)), rng.normal(0, 40, (n, 12))
    contact = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    ref = lc.encode_trunk_state(ts, fin, traj, contact, f)
    m = codec.encode_trunk_state(ts, fin, traj, contact, f)
    assert np.array_equal(m, ref)
    d = codec.decode_trunk_state(ref)
    assert np.array_equal(bits(d["traj"]), bits(traj)) and np.array_equal(d["contact"], contact) and np.array_equal(bits(d["f"]), bits(f))
    assert np.array_equal(bits(d["timestamp"]), bits(ts)) and np.array_equal(d["finished"], fin)
    q, v, tau = rng.normal(0, 1, (n, 19)), rng.normal(0, 3, (n, 18)), rng.normal(0, 20, (n, 12))
    rref = lc.encode_robot_state(q, v, tau)
    rm, st = codec.encode_robot_state(q, v, tau)
    assert np.array_equal(rm, rref) and (st == 0).all()
    dr, do = codec.decode_robot_state(rref), lc.decode_robot_state(rref)
    for k in ("q", "v", "tau"):
        assert np.array_equal(bits(dr[k]), bits(do[k]))


def test_device_path_roundtrip_and_velocity_order(codec):
    """torch device tensors (no host copies); encode(decode(x)) == x; tau sent in velocity order like (S.T @ u)[-12:]."""
    import torch
    from oracle import lcm_codec as lc
    n = 4096
    rng = np.random.default_rng(3)
    traj, f = rng.normal(0, 3, (n, 54)), rng.normal(0, 40, (n, 12))
    contact = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    ref = lc.encode_trunk_state(np.arange(n) * 1e-3, np.zeros(n, np.uint8), traj, contact, f)
    dm = torch.from_numpy(ref).cuda()
    d = codec.decode_trunk_state(dm)
    back = codec.encode_trunk_state(d["timestamp"], d["finished"], d["traj"], d["contact"], d["f"])
    torch.cuda.synchronize()
    assert torch.equal(back, dm)
    assert np.array_equal(bits(d["traj"].cpu().numpy()), bits(traj))
    # torque ordering: message slot j (velocity index 6 + j) carries the actuator that drives that joint
    model = codec.ctl.model
    tau = rng.normal(0, 10, (n, 12))
    msgs, st = codec.encode_robot_state(None, None, torch.from_numpy(tau).cuda(), tau_in_actuator_order=True)
    B = np.zeros((18, 12))
    ms = model.as_struct()
    for k in range(12):
        B[ms.v_index[k], ms.act_index[k]] = 1.0       # MakeActuationMatrix: one 1 per actuator column at its joint's velocity row
    expect = lc.encode_robot_state(None, None, (tau @ B.T)[:, 6:])
    assert np.array_equal(msgs.cpu().numpy(), expect)


def test_lcm_bridge_of_the_leafsystem_mirror(built):
    """use_lcm=True branch of BasicController.DoSetControlTorques (basic_controller.py:291-317): state from the latest
    robot_current_state message, torques published as robot_control_input, zeros to the Drake port."""
    from oracle import lcm_codec as lc
    from quadruped_drake_b200.controller import IDController, BatchedController
    from quadruped_drake_b200.synth import generate
    ctl = IDController("mini_cheetah", 5e-3, use_lcm=True)
    q, v, traj, contact = generate(ctl.batched.model, 1, 5, "stand", ctl.batched.fk)
    state_msg = lc.encode_robot_state(q, v, np.zeros((1, 12)))[0].tobytes()
    ctl.lcm_callback("robot_current_state", state_msg)
    q32, v32 = q.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64)
    assert np.array_equal(ctl.q, q32[0]) and np.array_equal(ctl.v, v32[0])
    from quadruped_drake_b200.controller import traj_to_dict
    ctx = ctl.CreateDefaultContext()
    ctx.FixValue(0, np.zeros(37))                 # the Drake state port is ignored in LCM mode
    ctx.FixValue(1, traj_to_dict(traj[0], contact[0]))
    u = ctl.EvalOutput(ctx, 0)
    assert (u == 0).all()                          # basic_controller.py:317
    chan, data = ctl.published[-1]
    assert chan == "robot_control_input"
    ref = BatchedController("mini_cheetah").step("id", q32, v32, traj, contact)
    B = np.zeros((18, 12)); ms = ctl.batched.model.as_struct()
    for k in range(12):
        B[ms.v_index[k], ms.act_index[k]] = 1.0
    expect = lc.encode_robot_state(None, None, (ref.tau @ B.T)[:, 6:])[0].tobytes()
    assert data == expect
    with pytest.raises(ValueError):
        ctl.lcm_callback("robot_current_state", b"\0" * 204)
