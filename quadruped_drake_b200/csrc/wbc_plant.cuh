// wbc_plant.cuh — one time step of the simulated robot on flat ground ("semi-implicit integrator + flat-ground
// contact"), the other half of the closed loop of simulate.py:36-57,160-182.
//
// The reference steps a discrete MultibodyPlant(time_step = dt) with a ground half-space (static = dynamic friction 1.0,
// simulate.py:44-46) through Drake's implicit contact solver, which is third party and not restatable. This file is a
// time-stepping scheme of the same family, defined here and pinned by invariants (tests/test_gpu_rollout.py) and by the numpy
// restatement in oracle/rollout.py - NOT by Drake:
//
//   v_free = v + dt M^-1 (B tau - C v - tau_g)                          forward dynamics from the APPLIED torques
//   u      = J_c v+ = J_c v_free + A p,   A = J_c M^-1 J_c'              foot velocities as a function of the contact impulses p
//   per foot:  0 <= p_n  _|_  u_n + bias_n >= 0                          no penetration (velocity level, Signorini)
//              |p_x|, |p_y| <= mu p_n, tangential velocity driven to 0   Coulomb friction, pyramid as in the controller
//   v+ = v_free + M^-1 J_c' p,   q+ = q + dt N(q) v+                      semi-implicit (symplectic) Euler
//
// bias_n = phi / dt for an open gap phi > 0 (the foot may close the gap in this step but not pass it) and erp * phi / dt for a
// penetration (Baumgarte push-out). The complementarity problem is solved by `iters` sweeps of projected Gauss-Seidel over
// the 12 rows in the fixed order foot LF RF LH RH x (normal, x, y), from p = 0. Contact state comes from the geometry here;
// the controller keeps using the PLANNED contact flags.
//
// One warp per robot, like the control-step kernels; the dynamics block is the same `dynamics_phase` the controllers use, and
// M^-1 is applied through M's block-arrow structure (four closed-form 3x3 leg-block inverses + a 6x6 Cholesky of the Schur
// complement), 13 right-hand sides at once with one lane each.
#pragma once
#include "wbc_device.cuh"

namespace wbcplant {
using namespace wbc;

struct alignas(16) PlantSmem {
  double X[18][13];        // M^-1 [J_c' | B tau - h]: column c < 12 contact row c, column 12 the free acceleration
  double A[12][13];        // Delassus matrix J_c M^-1 J_c' (row per lane, odd stride)
  double Sb[6][6];         // Schur complement of the leg blocks -> its Cholesky factor
  double Dinv[4][6];       // inverses of the 3x3 leg blocks (symmetric storage)
  double p[12];            // contact impulses (N s), row 3 k + i = foot k, world axis i
  double tauk[12];         // applied torque of internal joint k
};

struct PlantArgs {
  double* q; double* v; const double* tau; double* t; const int32_t* ctrl_status; int32_t* status_or; double* f_contact;
  const double* metrics; double* err_max; double* metrics_log; const int* step_counter;     // rollout bookkeeping (optional)
  long long n; double dt, mu, erp; int iters;
};

// Freeze mask: a robot whose controller reported any failure keeps its state (the reference asserts there) and stays flagged.
constexpr int PLANT_FREEZE = WBC_ST_MAXITER | WBC_ST_INFEASIBLE | WBC_ST_RANKDEF | WBC_ST_GIMBAL | WBC_ST_NOTPD | WBC_ST_BADQUAT |
                             WBC_ST_UNSUPPORTED | WBC_ST_DIVERGED;

WBC_DEV void plant_step_instance(WarpSmem& s, PlantSmem& ps, const wbc_model& md, const PlantArgs& a, long long inst, int lane) {
  int status = a.ctrl_status ? a.ctrl_status[inst] : 0;
  for (int i = lane; i < WBC_NQ; i += 32) s.q[i] = a.q[inst * WBC_NQ + i];
  for (int i = lane; i < WBC_NV; i += 32) s.v[i] = a.v[inst * WBC_NV + i];
  if (lane < 12) ps.tauk[lane] = a.tau[inst * WBC_NU + md.act_index[lane]];
  __syncwarp();
  dynamics_phase<DYN_STEP>(s, md, lane, status, nullptr);
  // ---- M^-1 through the block-arrow structure
  if (lane < 4) {
    const double* d = s.Mleg[lane];
    const double aa = d[0], b = d[1], c = d[2], e = d[3], f = d[4], g = d[5];     // [[a b c],[b e f],[c f g]]
    const double c00 = e * g - f * f, c01 = c * f - b * g, c02 = b * f - c * e;
    const double det = aa * c00 + b * c01 + c * c02, id = 1.0 / det;
    ps.Dinv[lane][0] = c00 * id; ps.Dinv[lane][1] = c01 * id; ps.Dinv[lane][2] = c02 * id;
    ps.Dinv[lane][3] = (aa * g - c * c) * id; ps.Dinv[lane][4] = (b * c - aa * f) * id; ps.Dinv[lane][5] = (aa * e - b * b) * id;
  }
  __syncwarp();
  for (int e = lane; e < 36; e += 32) {
    const int i = e / 6, j = e % 6;
    double acc = s.Mb[j][i];
    for (int k = 0; k < 4; ++k)
      for (int x = 0; x < 3; ++x)
        for (int y = 0; y < 3; ++y) acc = fma(-s.Mb[6 + 3 * k + x][i] * sym3get(ps.Dinv[k], x, y), s.Mb[6 + 3 * k + y][j], acc);
    ps.Sb[i][j] = acc;
  }
  __syncwarp();
  for (int j = 0; j < 6; ++j) {                     // Cholesky of Sb (lower), lanes = rows
    const double dj = ps.Sb[j][j];
    if (!(dj > 1e-300)) status |= WBC_ST_NOTPD;
    const double inv = 1.0 / sqrt(dj > 1e-300 ? dj : 1.0);
    __syncwarp();
    if (lane < 6 && lane >= j) ps.Sb[lane][j] *= inv;
    __syncwarp();
    if (lane < 6 && lane > j)
      for (int k = j + 1; k <= lane; ++k) ps.Sb[lane][k] = fma(-ps.Sb[lane][j], ps.Sb[k][j], ps.Sb[lane][k]);
    __syncwarp();
  }
  // ---- column c = lane < 13 of X = M^-1 [J_c' | B tau - h]
  if (lane < 13) {
    double rb[6], rl[4][3];
    if (lane < 12) {
      const int k = lane / 3, i = lane % 3;
      const V3 rh = ld3(s.rho[k]);
      // J_k = [-skew(rho) | 1 | L_k]  ->  J_k' e_i = [skew(rho) e_i ; e_i ; L_k[i][:]]
#pragma unroll
      for (int c = 0; c < 3; ++c) { rb[c] = -skew_ent(rh, i, c); rb[3 + c] = (c == i) ? 1.0 : 0.0; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int x = 0; x < 3; ++x) rl[kk][x] = (kk == k) ? s.L[k][i][x] : 0.0;
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) rb[c] = -s.hb[c];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int x = 0; x < 3; ++x) rl[kk][x] = ps.tauk[3 * kk + x] - s.hj[3 * kk + x];
    }
    double tl[4][3];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int x = 0; x < 3; ++x)
        tl[kk][x] = sym3get(ps.Dinv[kk], x, 0) * rl[kk][0] + sym3get(ps.Dinv[kk], x, 1) * rl[kk][1] + sym3get(ps.Dinv[kk], x, 2) * rl[kk][2];
    double xb[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double acc = rb[i];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int x = 0; x < 3; ++x) acc = fma(-s.Mb[6 + 3 * kk + x][i], tl[kk][x], acc);
      xb[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {                   // L y = cb
      double acc = xb[i];
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k < i) acc = fma(-ps.Sb[i][k], xb[k], acc);
      xb[i] = acc / ps.Sb[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {                  // L' x = y
      double acc = xb[i];
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k > i) acc = fma(-ps.Sb[k][i], xb[k], acc);
      xb[i] = acc / ps.Sb[i][i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) ps.X[i][lane] = xb[i];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      double c3[3];
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        double acc = rl[kk][x];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fma(-s.Mb[6 + 3 * kk + x][i], xb[i], acc);
        c3[x] = acc;
      }
#pragma unroll
      for (int x = 0; x < 3; ++x)
        ps.X[6 + 3 * kk + x][lane] = sym3get(ps.Dinv[kk], x, 0) * c3[0] + sym3get(ps.Dinv[kk], x, 1) * c3[1] + sym3get(ps.Dinv[kk], x, 2) * c3[2];
    }
  }
  __syncwarp();
  // ---- contact rows: lane r < 12 owns row r = 3 k + i of J_c: A[r][:] = J_r X[:, :12], u_free = J_r (v + dt X[:, 12])
  const int rk = lane < 12 ? lane / 3 : 0, ri = lane < 12 ? lane % 3 : 0;
  double u = 0.0, ainv = 0.0, bias = 0.0;
  if (lane < 12) {
    const V3 rh = ld3(s.rho[rk]);
    auto jrow = [&](int c) {
      double val = ps.X[3 + ri][c];
#pragma unroll
      for (int x = 0; x < 3; ++x) val = fma(-skew_ent(rh, ri, x), ps.X[x][c], val);
#pragma unroll
      for (int x = 0; x < 3; ++x) val = fma(s.L[rk][ri][x], ps.X[6 + 3 * rk + x][c], val);
      return val;
    };
    for (int c = 0; c < 12; ++c) ps.A[lane][c] = jrow(c);
    u = fma(a.dt, jrow(12), s.vf[rk][ri]);
    ainv = 1.0 / ps.A[lane][lane];
    if (ri == 2) {
      const double phi = s.q[6] + s.rho[rk][2];      // foot height over the ground plane z = 0
      bias = (phi > 0.0 ? phi : a.erp * phi) / a.dt;
    }
  }
  __syncwarp();
  // ---- projected Gauss-Seidel from p = 0: feet in order, normal row first, then the two tangential rows
  double p = 0.0;
  for (int it = 0; it < a.iters; ++it) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      double pn;
      {
        const int r = 3 * k + 2;
        const double pnew = fmax(0.0, p - (u + bias) * ainv);
        const double dl = shfl(pnew - p, r);
        pn = shfl(pnew, r);
        if (lane == r) p = pnew;
        if (lane < 12) u = fma(ps.A[lane][r], dl, u);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = 3 * k + i;
        const double lim = a.mu * pn;
        const double pnew = fmin(fmax(p - u * ainv, -lim), lim);
        const double dl = shfl(pnew - p, r);
        if (lane == r) p = pnew;
        if (lane < 12) u = fma(ps.A[lane][r], dl, u);
      }
    }
  }
  if (lane < 12) ps.p[lane] = p;
  __syncwarp();
  // ---- v+ = v + dt X[:, 12] + X[:, :12] p (rows in internal joint order), then q+ = q + dt N(q) v+
  double vn = 0.0;
  int vi = 0;
  if (lane < 18) {
    vi = lane < 6 ? lane : md.v_index[lane - 6];
    double acc = fma(a.dt, ps.X[lane][12], s.v[vi]);
#pragma unroll
    for (int c = 0; c < 12; ++c) acc = fma(ps.X[lane][c], ps.p[c], acc);
    vn = acc;
  }
  const bool finite = __all_sync(WBC_FULL, lane >= 18 || fabs(vn) < 1e6);
  if (!finite) status |= WBC_ST_DIVERGED;
  if (lane == 0 && a.status_or) a.status_or[inst] |= status;
  if (a.metrics) {
    if (lane == 0 && a.err_max) { const double e = a.metrics[inst * WBC_NMETRIC + 1]; if (e > a.err_max[inst]) a.err_max[inst] = e; }
    if (a.metrics_log && lane < WBC_NMETRIC)
      a.metrics_log[((long long)(*a.step_counter) * a.n + inst) * WBC_NMETRIC + lane] = a.metrics[inst * WBC_NMETRIC + lane];
  }
  if (a.f_contact && lane < 12) a.f_contact[inst * 12 + lane] = (status & PLANT_FREEZE) ? 0.0 : p / a.dt;
  if (lane == 0 && a.t) a.t[inst] += a.dt;
  if (status & PLANT_FREEZE) return;
  const double wx = shfl(vn, 0), wy = shfl(vn, 1), wz = shfl(vn, 2);
  if (lane < 18) a.v[inst * WBC_NV + vi] = vn;
  if (lane == 0) {
    const double qw = s.q[0], qx = s.q[1], qy = s.q[2], qz = s.q[3];
    const double nw = qw + 0.5 * a.dt * (-wx * qx - wy * qy - wz * qz);
    const double nx = qx + 0.5 * a.dt * (wx * qw + wy * qz - wz * qy);
    const double ny = qy + 0.5 * a.dt * (wy * qw + wz * qx - wx * qz);
    const double nz = qz + 0.5 * a.dt * (wz * qw + wx * qy - wy * qx);
    const double inv = 1.0 / sqrt(nw * nw + nx * nx + ny * ny + nz * nz);
    double* qo = a.q + inst * WBC_NQ;
    qo[0] = nw * inv; qo[1] = nx * inv; qo[2] = ny * inv; qo[3] = nz * inv;
  }
  if (lane >= 3 && lane < 18) a.q[inst * WBC_NQ + 1 + vi] = fma(a.dt, vn, s.q[1 + vi]);   // base position (q 4..6) and joints
  __syncwarp();
}

}  // namespace wbcplant
