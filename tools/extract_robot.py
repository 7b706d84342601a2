#!/usr/bin/env python
"""Extract the neutral robot descriptions shipped under quadruped_drake_b200/robots/.

Run in the build container (the reference tree is not present on the GPU box):
    python tools/extract_robot.py [/root/reference]
Reads the two URDFs the reference loads (reference simulate.py:31 and the ANYmal
alternative, models/anymal_b_simple_description/urdf/anymal_drake.urdf) and writes
plain-number JSON (kinematic tree + inertial data + actuator order). No reference
code is copied; these are the robots' physical parameters.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from quadruped_drake_b200.urdf import parse_urdf, save_description  # noqa: E402

ROBOTS = {
    "mini_cheetah": ("models/mini_cheetah/mini_cheetah_mesh.urdf", "body"),
    "anymal_b": ("models/anymal_b_simple_description/urdf/anymal_drake.urdf", "base"),
}
FEET = ["LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"]  # reference basic_controller.py:67-70


def main():
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    out = Path(__file__).resolve().parents[1] / "quadruped_drake_b200" / "robots"
    for name, (rel, base) in ROBOTS.items():
        d = parse_urdf(ref / rel)
        d["name"] = name
        d["source"] = rel
        d["base_link"] = base          # reference basic_controller.py:65 ("body" / "base")
        d["foot_frames"] = FEET
        save_description(d, out / f"{name}.json")
        print(name, len(d["links"]), "links", len(d["joints"]), "joints",
              "mass", sum(l["mass"] for l in d["links"]))


if __name__ == "__main__":
    main()
