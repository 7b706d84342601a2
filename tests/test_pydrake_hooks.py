"""pydrake probe and the DOF-order derivation against a stand-in plant that exposes the three pydrake calls the
reference itself uses (GetJointByName, Joint.velocity_start, MakeActuationMatrix). pydrake itself has never been importable
in this image; when it is, the same functions run on the real MultibodyPlant."""
import numpy as np
import pytest

from quadruped_drake_b200 import drake_bridge, load_robot


class _Joint:
    def __init__(self, start):
        self._s = start

    def velocity_start(self):
        return self._s


class FakePlant:
    """MultibodyPlant stand-in numbering the leg joints in a given order (floating base first, 6 velocities)."""

    def __init__(self, robot, order, actuator_order=None):
        self.names = drake_bridge.LEG_JOINTS[robot]
        self.pos = {self.names[k]: 6 + i for i, k in enumerate(order)}
        act = actuator_order or list(range(12))               # actuator a drives internal joint act[a]
        self.B = np.zeros((18, 12))
        for a, k in enumerate(act):
            self.B[self.pos[self.names[k]], a] = 1.0

    def GetJointByName(self, name):
        if name not in self.pos:
            raise RuntimeError("no joint " + name)
        return _Joint(self.pos[name])

    def MakeActuationMatrix(self):
        return self.B

    def num_velocities(self):
        return 18


def test_probe_reports_the_import_result():
    p = drake_bridge.probe()
    assert set(p) == {"importable", "version", "error"} and isinstance(p["importable"], bool)
    assert p["importable"] or "pydrake" in p["error"]


@pytest.mark.parametrize("robot", ["mini_cheetah", "anymal_b"])
def test_dof_order_is_read_from_the_plant(robot):
    depth = list(range(12))
    breadth = [3 * l + j for j in range(3) for l in range(4)]         # 2021-era Drake: all abductions, all hips, all knees
    for order, preset in ((depth, "depth_first"), (breadth, "breadth_first")):
        plant = FakePlant(robot, order)
        assert drake_bridge.is_drake_plant(plant) and drake_bridge.robot_of_plant(plant) == robot
        v_index, act_index = drake_bridge.derive_v_index(plant, robot)
        m = load_robot(robot, dof_order=preset)
        assert np.array_equal(v_index, m.v_index) and np.array_equal(act_index, m.act_index)
    # an arbitrary numbering and a non-identity actuator map (the reference warns about it, basic_controller.py:311-313)
    rng = np.random.default_rng(0)
    order, act = rng.permutation(12).tolist(), rng.permutation(12).tolist()
    plant = FakePlant(robot, order, act)
    v_index, act_index = drake_bridge.derive_v_index(plant, robot)
    for pos, k in enumerate(order):
        assert v_index[k] == 6 + pos
    for a, k in enumerate(act):
        assert act_index[k] == a
    assert np.array_equal(plant.B, _actuation(v_index, act_index))


def _actuation(v_index, act_index):
    B = np.zeros((18, 12))
    for k in range(12):
        B[v_index[k], act_index[k]] = 1.0
    return B


def test_bad_layout_is_rejected():
    plant = FakePlant("mini_cheetah", list(range(12)))
    plant.pos[drake_bridge.LEG_JOINTS["mini_cheetah"][0]] = 7       # two joints on the same velocity index
    with pytest.raises(ValueError):
        drake_bridge.derive_v_index(plant, "mini_cheetah")
