// wbc_rollout.cuh — state integration for the closed-loop batched rollout.
//
// The caller of the control step in the reference is Drake's Simulator stepping a discrete MultibodyPlant(time_step=dt)
// (simulate.py:38,160-182): each step pulls the controller torques and advances (q, v) by dt with a symplectic Euler
// scheme and a compliant ground-contact solve. Drake's contact solver is not restatable here, so the rollout uses the
// contact-consistent acceleration the QP itself returns (vd satisfies M vd + Cv + tau_g = B tau + J_c' f and the no-slip
// rows of AddContactConstraint for the planned stance feet): "planned contacts hold". Integration is Drake's
// semi-implicit Euler form: v+ = v + dt vd, q+ = q + dt N(q) v+ with N the quaternion-rate map of a floating base whose
// angular velocity is expressed in the world frame, followed by a renormalisation of the quaternion.
#pragma once
#include <stdint.h>
#include "wbc.h"

namespace wbcroll {

// One thread per instance: 37 doubles in / out, 18 in. HBM bound (and tiny next to the control step).
__global__ void __launch_bounds__(256) integrate_kernel(long long n, double dt, double* __restrict__ q, double* __restrict__ v,
                                                        const double* __restrict__ vd, double* __restrict__ t,
                                                        const int* __restrict__ status, int* __restrict__ status_or,
                                                        const double* __restrict__ metrics, double* __restrict__ err_max,
                                                        double* __restrict__ metrics_log, const int* __restrict__ step_counter) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* qi = q + i * WBC_NQ;
  double* vi = v + i * WBC_NV;
  const double* ai = vd + i * WBC_NV;
  // an instance whose QP failed (the reference asserts there, inverse_dynamics_controller.py:224) or whose accelerations are
  // not finite is frozen: its state stays where it was and the failure stays visible in status_or
  int st = status ? status[i] : 0;
  double vn[WBC_NV];
  bool finite = true;
#pragma unroll
  for (int k = 0; k < WBC_NV; ++k) { vn[k] = vi[k] + dt * ai[k]; finite = finite && (fabs(vn[k]) < 1e6); }
  if (!finite) st |= WBC_ST_DIVERGED;
  if (status_or) status_or[i] |= st;
  if (err_max && metrics) { const double e = metrics[i * WBC_NMETRIC + 1]; if (e > err_max[i]) err_max[i] = e; }
  if (metrics_log && metrics) {
    double* dst = metrics_log + ((long long)(*step_counter) * n + i) * WBC_NMETRIC;
#pragma unroll
    for (int k = 0; k < WBC_NMETRIC; ++k) dst[k] = metrics[i * WBC_NMETRIC + k];
  }
  if (st != 0) {      // any status bit (gimbal lock included): the controller returned zero torques / accelerations for this robot
    if (t) t[i] += dt;
    return;
  }
#pragma unroll
  for (int k = 0; k < WBC_NV; ++k) vi[k] = vn[k];
  // quaternion rate for world-frame angular velocity: qdot = 1/2 [0, w] (x) q
  const double qw = qi[0], qx = qi[1], qy = qi[2], qz = qi[3];
  const double wx = vn[0], wy = vn[1], wz = vn[2];
  double nw = qw + 0.5 * dt * (-wx * qx - wy * qy - wz * qz);
  double nx = qx + 0.5 * dt * (wx * qw + wy * qz - wz * qy);
  double ny = qy + 0.5 * dt * (wy * qw + wz * qx - wx * qz);
  double nz = qz + 0.5 * dt * (wz * qw + wx * qy - wy * qx);
  const double inv = 1.0 / sqrt(nw * nw + nx * nx + ny * ny + nz * nz);
  qi[0] = nw * inv; qi[1] = nx * inv; qi[2] = ny * inv; qi[3] = nz * inv;
#pragma unroll
  for (int k = 0; k < 3; ++k) qi[4 + k] += dt * vn[3 + k];
#pragma unroll
  for (int k = 0; k < WBC_NU; ++k) qi[7 + k] += dt * vn[6 + k];
  if (t) t[i] += dt;
}

// Step counter of a rollout (device resident so that a captured CUDA graph can be replayed unchanged).
__global__ void bump_counter_kernel(int* c) { *c += 1; }

}  // namespace wbcroll
