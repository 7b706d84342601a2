"""URDF -> neutral robot description (plain dict / JSON).

The reference hands the URDF to Drake's Parser (reference simulate.py:31-40);
here we only need the kinematic tree and the inertial data, so a small
xml.etree reader is enough.  The output is a *neutral* description: every
link and joint of the file, nothing merged, nothing reordered.  Both the
product flattener (model.py) and the test oracle (oracle/) start from it.

Semantics restated from Drake's URDF parser:
  * <inertia> is about the link CoM, expressed in the inertial frame
    (<inertial><origin rpy>), which we rotate into the link frame;
  * `continuous` is an unlimited revolute joint;
  * fixed joints weld; links without <inertial> are massless frames;
  * actuator order is the order of the <transmission> elements.
"""
from __future__ import annotations

import json
import math
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np


def _floats(text, n, default):
    if text is None:
        return list(default)
    vals = [float(t) for t in text.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return vals


def rpy_to_matrix(rpy):
    """R = Rz(yaw) Ry(pitch) Rx(roll) (URDF and Drake RollPitchYaw convention)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


def parse_urdf(path) -> dict:
    """Read a URDF file into the neutral description used across this repo."""
    root = ET.parse(str(path)).getroot()
    links = []
    for el in root.findall("link"):
        ine = el.find("inertial")
        mass, com, inertia = 0.0, [0.0, 0.0, 0.0], [0.0] * 6
        if ine is not None:
            mass = float(ine.find("mass").get("value"))
            org = ine.find("origin")
            xyz = _floats(org.get("xyz") if org is not None else None, 3, (0, 0, 0))
            rpy = _floats(org.get("rpy") if org is not None else None, 3, (0, 0, 0))
            it = ine.find("inertia")
            ixx, iyy, izz = (float(it.get(k)) for k in ("ixx", "iyy", "izz"))
            ixy, ixz, iyz = (float(it.get(k)) for k in ("ixy", "ixz", "iyz"))
            I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
            R = rpy_to_matrix(rpy)
            I = R @ I @ R.T  # inertial frame -> link frame axes
            com = xyz
            inertia = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
        links.append({"name": el.get("name"), "mass": mass, "com": [float(c) for c in com],
                      "inertia_com": [float(v) for v in inertia]})
    joints = []
    for el in root.findall("joint"):
        org = el.find("origin")
        xyz = _floats(org.get("xyz") if org is not None else None, 3, (0, 0, 0))
        rpy = _floats(org.get("rpy") if org is not None else None, 3, (0, 0, 0))
        ax = el.find("axis")
        axis = _floats(ax.get("xyz") if ax is not None else None, 3, (1, 0, 0))
        lim = el.find("limit")
        effort = float(lim.get("effort")) if lim is not None and lim.get("effort") else float("inf")
        jtype = el.get("type")
        if jtype == "continuous":
            jtype = "revolute"
        if jtype not in ("revolute", "fixed"):
            raise ValueError(f"unsupported joint type {jtype!r} ({el.get('name')})")
        joints.append({"name": el.get("name"), "type": jtype,
                       "parent": el.find("parent").get("link"), "child": el.find("child").get("link"),
                       "xyz": xyz, "rpy": rpy, "axis": axis, "effort": effort})
    actuators = []
    for el in root.findall("transmission"):
        j = el.find("joint")
        if j is not None:
            actuators.append(j.get("name"))
    return {"name": root.get("name"), "links": links, "joints": joints, "actuated_joints": actuators}


def save_description(desc: dict, path) -> None:
    Path(path).write_text(json.dumps(desc, indent=1) + "\n")


def load_description(path) -> dict:
    return json.loads(Path(path).read_text())
