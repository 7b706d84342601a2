#!/usr/bin/env python
"""Generate tests/golden/towr_spline.npz from the REFERENCE's own TOWR sources (oracle/_ref/libtowr_ref.so, built by
oracle/ref_build/Makefile from /root/reference/towr/src/*.cc): gait phase durations of the five combos, spline
segment ids and Spline::GetPoint values of seeded random Hermite splines, IsContactPhase flags.
Usage: make -C oracle/ref_build && python tools/make_golden_traj.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import towr_ref as ref  # noqa: E402


def main():
    assert ref.available(), "build oracle/_ref first: make -C oracle/ref_build"
    data = {}
    cs = np.zeros((5, 4), np.uint8)
    contact_ts = np.concatenate([np.linspace(0, 5, 201), [0.3 / 4.81 * 5.0, 2.5000000001]])
    flags = np.zeros((5, 4, len(contact_ts)), np.uint8)
    for combo in range(5):
        for ee in range(4):
            pd = ref.phase_durations(combo, 5.0, ee)
            data[f"phase_c{combo}_e{ee}"] = pd
            cs[combo, ee] = ref.contact_at_start(combo, ee)
            flags[combo, ee] = [ref.is_contact_phase(t, pd, cs[combo, ee]) for t in contact_ts]
    data.update(contact_start=cs, contact_ts=contact_ts, contact_flags=flags)
    rng = np.random.default_rng(20260117 + 200)
    n_splines = 12
    for k in range(n_splines):
        n = int(rng.integers(1, 40))
        d = rng.uniform(0.02, 0.4, n)
        nodes = rng.normal(0, 1.5, (n + 1, 6))
        T = d.sum()
        ts = np.concatenate([rng.uniform(0, T * 0.999, 48), np.cumsum(d)[:-1][:8], [0.0]])
        data[f"dur_{k}"], data[f"nodes_{k}"], data[f"ts_{k}"] = d, nodes, ts
        data[f"pts_{k}"] = ref.spline_points(d, nodes, ts)
        data[f"seg_{k}"] = np.array([ref.segment_id(t, d) for t in ts])
    data["n_splines"] = n_splines
    out = ROOT / "tests" / "golden" / "towr_spline.npz"
    np.savez_compressed(out, **data)
    print("wrote", out)


if __name__ == "__main__":
    main()
