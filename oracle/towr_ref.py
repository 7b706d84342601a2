"""ORACLE SUPPORT: ctypes access to oracle/_ref/libtowr_ref.so - the reference's own TOWR spline / gait sources
compiled by oracle/ref_build/Makefile. Test infrastructure only (pins oracle/trajectory.py, regenerates golden vectors)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB = Path(__file__).resolve().parent / "_ref" / "libtowr_ref.so"
_lib = None


def available() -> bool:
    return LIB.exists()


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(str(LIB))
        dp = C.POINTER(C.c_double)
        l.towr_ref_phase_durations.argtypes = [C.c_int, C.c_double, C.c_int, dp, C.c_int]
        l.towr_ref_contact_at_start.argtypes = [C.c_int, C.c_int]
        l.towr_ref_segment_id.argtypes = [C.c_double, dp, C.c_int]
        l.towr_ref_is_contact_phase.argtypes = [C.c_double, dp, C.c_int, C.c_int]
        l.towr_ref_spline_point.argtypes = [C.c_int, dp, dp, C.c_int, dp, dp]
        l.towr_ref_spline_point.restype = None
        _lib = l
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def phase_durations(combo, t_total, ee):
    out = np.zeros(64)
    n = lib().towr_ref_phase_durations(combo, t_total, ee, _dp(out), 64)
    return out[:n].copy()


def contact_at_start(combo, ee):
    return bool(lib().towr_ref_contact_at_start(combo, ee))


def segment_id(t, durations):
    d = np.ascontiguousarray(durations, float)
    return int(lib().towr_ref_segment_id(float(t), _dp(d), len(d)))


def is_contact_phase(t, durations, contact_start):
    d = np.ascontiguousarray(durations, float)
    return bool(lib().towr_ref_is_contact_phase(float(t), _dp(d), len(d), int(bool(contact_start))))


def spline_points(durations, nodes, ts):
    """-> [len(ts), 9] = p, v, a of Spline::GetPoint."""
    d = np.ascontiguousarray(durations, float)
    nd = np.ascontiguousarray(nodes, float).reshape(len(d) + 1, 6)
    t = np.ascontiguousarray(ts, float)
    out = np.zeros((len(t), 9))
    lib().towr_ref_spline_point(len(d), _dp(d), _dp(nd), len(t), _dp(t), _dp(out))
    return out
