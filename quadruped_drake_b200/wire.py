"""LCM wire codecs through the C ABI: batched, on the device.

    trunk_state_t             reference lcm_types/trunk_state_t.lcm, lcm_types/trunklcm/trunk_state_t.py
    robot_state_control_lcmt  reference lcm_types/robot_state_control_lcmt.lcm, lcm_types/cheetahlcm/...

`WireCodec(ctl)` works on uint8[N, 549] / uint8[N, 204] message arrays. NumPy arguments go through the `_host` entry
points (copies inside the library); torch CUDA tensors go straight to the device entry points on torch's current
stream. Per-message problems come back in `status` (1 = fingerprint mismatch, 2 = float32 overflow) where the
reference's generated codecs raise ValueError / OverflowError. There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import np_ptr
from .model import NQ, NTRAJ, NU, NV

TRUNK_STATE_BYTES, ROBOT_STATE_BYTES = 549, 204
WIRE_BADFINGERPRINT, WIRE_OVERFLOW = 1, 2


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _tp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class WireCodec:
    def __init__(self, ctl):
        self.ctl, self.lib, self._h = ctl, ctl.lib, ctl._h

    def _stream(self, t):
        import torch
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    # ------------------------------------------------------------------ trunk_state_t
    def decode_trunk_state(self, msgs):
        """uint8[N,549] -> dict(timestamp, finished, traj[N,54], contact[N,4], f[N,12], status[N])."""
        if _is_torch(msgs):
            import torch
            n, dev = msgs.shape[0], msgs.device
            assert msgs.dtype == torch.uint8 and msgs.is_contiguous() and msgs.shape[1] == TRUNK_STATE_BYTES
            o = dict(timestamp=torch.empty(n, dtype=torch.float64, device=dev), finished=torch.empty(n, dtype=torch.uint8, device=dev),
                     traj=torch.empty((n, NTRAJ), dtype=torch.float64, device=dev), contact=torch.empty((n, 4), dtype=torch.uint8, device=dev),
                     f=torch.empty((n, 12), dtype=torch.float64, device=dev), status=torch.empty(n, dtype=torch.int32, device=dev))
            self.ctl._check(self.lib.wbc_lcm_decode_trunk_state(self._h, n, _tp(msgs), _tp(o["timestamp"]), _tp(o["finished"]), _tp(o["traj"]),
                                                                _tp(o["contact"]), _tp(o["f"]), _tp(o["status"]), self._stream(msgs)),
                            "wbc_lcm_decode_trunk_state")
            return o
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8).reshape(-1, TRUNK_STATE_BYTES)
        n = len(msgs)
        o = dict(timestamp=np.empty(n), finished=np.empty(n, np.uint8), traj=np.empty((n, NTRAJ)), contact=np.empty((n, 4), np.uint8),
                 f=np.empty((n, 12)), status=np.empty(n, np.int32))
        self.ctl._check(self.lib.wbc_lcm_decode_trunk_state_host(self._h, n, np_ptr(msgs), np_ptr(o["timestamp"]), np_ptr(o["finished"]),
                                                                 np_ptr(o["traj"]), np_ptr(o["contact"]), np_ptr(o["f"]), np_ptr(o["status"])),
                        "wbc_lcm_decode_trunk_state_host")
        return o

    def encode_trunk_state(self, timestamp, finished, traj, contact, f=None):
        """-> uint8[N,549] (what towr/trunk_mpc.cpp:19-68 publishes per sample)."""
        if _is_torch(traj):
            import torch
            n, dev = traj.shape[0], traj.device
            msgs = torch.empty((n, TRUNK_STATE_BYTES), dtype=torch.uint8, device=dev)
            self.ctl._check(self.lib.wbc_lcm_encode_trunk_state(self._h, n, _tp(timestamp), _tp(finished), _tp(traj), _tp(contact), _tp(f),
                                                                _tp(msgs), self._stream(traj)), "wbc_lcm_encode_trunk_state")
            return msgs
        traj = np.ascontiguousarray(traj, dtype=np.float64).reshape(-1, NTRAJ)
        n = len(traj)
        ts = None if timestamp is None else np.ascontiguousarray(timestamp, dtype=np.float64).reshape(n)
        fin = None if finished is None else np.ascontiguousarray(np.asarray(finished) != 0, dtype=np.uint8).reshape(n)
        contact = np.ascontiguousarray(np.asarray(contact) != 0, dtype=np.uint8).reshape(n, 4)
        fp = None if f is None else np.ascontiguousarray(f, dtype=np.float64).reshape(n, 12)
        msgs = np.empty((n, TRUNK_STATE_BYTES), np.uint8)
        opt = lambda a: None if a is None else np_ptr(a)  # noqa: E731
        self.ctl._check(self.lib.wbc_lcm_encode_trunk_state_host(self._h, n, opt(ts), opt(fin), np_ptr(traj), np_ptr(contact), opt(fp),
                                                                 np_ptr(msgs)), "wbc_lcm_encode_trunk_state_host")
        return msgs

    # ------------------------------------------------------------------ robot_state_control_lcmt
    def decode_robot_state(self, msgs):
        """uint8[N,204] -> dict(q[N,19], v[N,18], tau[N,12], status[N]) (basic_controller.py:79-87)."""
        if _is_torch(msgs):
            import torch
            n, dev = msgs.shape[0], msgs.device
            assert msgs.dtype == torch.uint8 and msgs.is_contiguous() and msgs.shape[1] == ROBOT_STATE_BYTES
            o = dict(q=torch.empty((n, NQ), dtype=torch.float64, device=dev), v=torch.empty((n, NV), dtype=torch.float64, device=dev),
                     tau=torch.empty((n, NU), dtype=torch.float64, device=dev), status=torch.empty(n, dtype=torch.int32, device=dev))
            self.ctl._check(self.lib.wbc_lcm_decode_robot_state(self._h, n, _tp(msgs), _tp(o["q"]), _tp(o["v"]), _tp(o["tau"]), _tp(o["status"]),
                                                                self._stream(msgs)), "wbc_lcm_decode_robot_state")
            return o
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8).reshape(-1, ROBOT_STATE_BYTES)
        n = len(msgs)
        o = dict(q=np.empty((n, NQ)), v=np.empty((n, NV)), tau=np.empty((n, NU)), status=np.empty(n, np.int32))
        self.ctl._check(self.lib.wbc_lcm_decode_robot_state_host(self._h, n, np_ptr(msgs), np_ptr(o["q"]), np_ptr(o["v"]), np_ptr(o["tau"]),
                                                                 np_ptr(o["status"])), "wbc_lcm_decode_robot_state_host")
        return o

    def encode_robot_state(self, q, v, tau, tau_in_actuator_order=False):
        """-> (uint8[N,204], status[N]). q / v None: zeros (the controller's outgoing message, basic_controller.py:309-314)."""
        flag = 1 if tau_in_actuator_order else 0
        if _is_torch(tau):
            import torch
            n, dev = tau.shape[0], tau.device
            msgs = torch.empty((n, ROBOT_STATE_BYTES), dtype=torch.uint8, device=dev)
            st = torch.empty(n, dtype=torch.int32, device=dev)
            self.ctl._check(self.lib.wbc_lcm_encode_robot_state(self._h, n, _tp(q), _tp(v), _tp(tau), flag, _tp(msgs), _tp(st),
                                                                self._stream(tau)), "wbc_lcm_encode_robot_state")
            return msgs, st
        tau = np.ascontiguousarray(tau, dtype=np.float64).reshape(-1, NU)
        n = len(tau)
        q = None if q is None else np.ascontiguousarray(q, dtype=np.float64).reshape(n, NQ)
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float64).reshape(n, NV)
        msgs, st = np.empty((n, ROBOT_STATE_BYTES), np.uint8), np.empty(n, np.int32)
        opt = lambda a: None if a is None else np_ptr(a)  # noqa: E731
        self.ctl._check(self.lib.wbc_lcm_encode_robot_state_host(self._h, n, opt(q), opt(v), np_ptr(tau), flag, np_ptr(msgs), np_ptr(st)),
                        "wbc_lcm_encode_robot_state_host")
        return msgs, st
