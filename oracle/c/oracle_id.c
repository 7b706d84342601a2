/* ORACLE (test infrastructure / CPU baseline only - never linked into the product library).
 *
 * Plain-C restatement of the reference's per-step path for the ID controller
 * (reference controllers/inverse_dynamics_controller.py:103-234 on top of
 * controllers/basic_controller.py:101-115,173-196), i.e. the same algorithm class the reference runs
 * through pydrake: mass matrix by nv inverse-dynamics passes (CalcMassMatrixViaInverseDynamics), bias and
 * gravity by two more passes, foot Jacobians / Jdot*v, then the FULL-SIZE QP the reference hands to OSQP
 * (variables [vd(18); tau(12); f(3 nc)], 18+3nc equalities, 4nc friction rows) solved by a dense Mehrotra
 * predictor-corrector interior-point method. pydrake / OSQP are un-vendored, unpinned dependencies that
 * cannot be installed here: **parity unpinned** (see oracle/dynamics.py). This twin is validated against
 * the numpy oracle in tests/test_oracle_cport.py (dynamics to round-off, vd to 1e-5; its IPM has no
 * active-set polish, so tau / f carry ~1e-2 of error along the tie-break directions - it is the TIMED
 * baseline, the numpy oracle is the parity reference) and is what bench.py times as `cpu_baseline` (kind
 * "port") and as the `--impl reference` arm, one instance stream per host thread.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "wbc.h"

#define NV 18
#define NU 12
#define NMAX 42 /* 30 + 3*4 */
#define MEMAX 30
#define MIMAX 16
#define KMAX (NMAX + MEMAX)

typedef struct { double x, y, z; } v3;
static v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 scl(double s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static v3 crs(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static double dt3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 mv(const double R[9], v3 a) { /* row-major 3x3 */
  return V(R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z, R[6] * a.x + R[7] * a.y + R[8] * a.z);
}
static void mm(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

/* world-frame kinematics of the 13 bodies for state (q, v) and generalized acceleration vd (Drake coordinates) */
typedef struct {
  double R[13][9]; v3 p[13], w[13], al[13], vo[13], ao[13], axis[13];
} kin_t;

static void kinematics(const wbc_model* md, const double* q, const double* v, const double* vd, kin_t* k) {
  double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
  double n = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  qw /= n; qx /= n; qy /= n; qz /= n;
  double* R = k->R[0];
  R[0] = 1 - 2 * (qy * qy + qz * qz); R[1] = 2 * (qx * qy - qw * qz); R[2] = 2 * (qx * qz + qw * qy);
  R[3] = 2 * (qx * qy + qw * qz); R[4] = 1 - 2 * (qx * qx + qz * qz); R[5] = 2 * (qy * qz - qw * qx);
  R[6] = 2 * (qx * qz - qw * qy); R[7] = 2 * (qy * qz + qw * qx); R[8] = 1 - 2 * (qx * qx + qy * qy);
  k->p[0] = V(q[4], q[5], q[6]);
  k->w[0] = V(v[0], v[1], v[2]); k->vo[0] = V(v[3], v[4], v[5]);
  k->al[0] = V(vd[0], vd[1], vd[2]); k->ao[0] = V(vd[3], vd[4], vd[5]);
  for (int j = 0; j < 12; ++j) {
    const int b = j + 1, pa = (j % 3 == 0) ? 0 : j; /* parent body */
    const int vi = md->v_index[j];
    const double th = q[vi + 1], thd = v[vi], thdd = vd[vi];
    v3 r = mv(k->R[pa], V(md->joint_xyz[j][0], md->joint_xyz[j][1], md->joint_xyz[j][2]));
    v3 a = V(md->joint_axis[j][0], md->joint_axis[j][1], md->joint_axis[j][2]);
    v3 aw = mv(k->R[pa], a);
    double s = sin(th), c = cos(th), oc = 1 - c;
    double Rot[9] = {c + oc * a.x * a.x, -s * a.z + oc * a.x * a.y, s * a.y + oc * a.x * a.z,
                     s * a.z + oc * a.y * a.x, c + oc * a.y * a.y, -s * a.x + oc * a.y * a.z,
                     -s * a.y + oc * a.z * a.x, s * a.x + oc * a.z * a.y, c + oc * a.z * a.z};
    mm(k->R[pa], Rot, k->R[b]);
    k->p[b] = add(k->p[pa], r);
    k->vo[b] = add(k->vo[pa], crs(k->w[pa], r));
    k->ao[b] = add(add(k->ao[pa], crs(k->al[pa], r)), crs(k->w[pa], crs(k->w[pa], r)));
    k->w[b] = add(k->w[pa], scl(thd, aw));
    k->al[b] = add(add(k->al[pa], scl(thdd, aw)), scl(thd, crs(k->w[pa], aw)));
    k->axis[b] = aw;
  }
}

/* generalized force of a wrench (F at point pt, N) on body b: tau += J_v' F + J_w' N */
static void project(const wbc_model* md, const kin_t* k, int b, v3 pt, v3 F, v3 N, double* tau) {
  v3 m0 = add(crs(sub(pt, k->p[0]), F), N); /* -skew(pt-p0)' F = (pt-p0) x F */
  tau[0] += m0.x; tau[1] += m0.y; tau[2] += m0.z; tau[3] += F.x; tau[4] += F.y; tau[5] += F.z;
  while (b > 0) {
    const int j = b - 1;
    tau[md->v_index[j]] += dt3(crs(k->axis[b], sub(pt, k->p[b])), F) + dt3(k->axis[b], N);
    b = (j % 3 == 0) ? 0 : j;
  }
}

/* tau = M vd + C v + tau_g (controller sign); gravity optional */
static void inverse_dynamics(const wbc_model* md, const double* q, const double* v, const double* vd, int gravity, double* tau) {
  kin_t k;
  kinematics(md, q, v, vd, &k);
  memset(tau, 0, NV * sizeof(double));
  v3 g = gravity ? V(md->gravity[0], md->gravity[1], md->gravity[2]) : V(0, 0, 0);
  for (int b = 0; b < 13; ++b) {
    const double m = md->mass[b];
    if (m == 0.0) continue;
    v3 rc = mv(k.R[b], V(md->com[b][0], md->com[b][1], md->com[b][2]));
    v3 w = k.w[b], al = k.al[b];
    v3 ac = add(add(k.ao[b], crs(al, rc)), crs(w, crs(w, rc)));
    const double* ic = md->inertia_com[b];
    double I[9] = {ic[0], ic[3], ic[4], ic[3], ic[1], ic[5], ic[4], ic[5], ic[2]}, T[9], Rt[9], Iw[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt[3 * i + j] = k.R[b][3 * j + i];
    mm(k.R[b], I, T); mm(T, Rt, Iw);
    v3 F = scl(m, sub(ac, g));
    v3 N = add(mv(Iw, al), crs(w, mv(Iw, w)));
    project(md, &k, b, add(k.p[b], rc), F, N, tau);
  }
}

static void foot_quantities(const wbc_model* md, const double* q, const double* v, int leg, double p[3], double J[3][NV], double Jdv[3]) {
  static const double zero[NV] = {0};
  kin_t k;
  kinematics(md, q, v, zero, &k);
  const int b = 3 * leg + 3;
  v3 off = mv(k.R[b], V(md->foot_xyz[leg][0], md->foot_xyz[leg][1], md->foot_xyz[leg][2]));
  v3 pt = add(k.p[b], off);
  v3 a = add(add(k.ao[b], crs(k.al[b], off)), crs(k.w[b], crs(k.w[b], off)));
  p[0] = pt.x; p[1] = pt.y; p[2] = pt.z; Jdv[0] = a.x; Jdv[1] = a.y; Jdv[2] = a.z;
  memset(J, 0, 3 * NV * sizeof(double));
  v3 d = sub(pt, k.p[0]);
  J[0][1] = d.z; J[0][2] = -d.y; J[1][0] = -d.z; J[1][2] = d.x; J[2][0] = d.y; J[2][1] = -d.x; /* -skew(d) */
  J[0][3] = J[1][4] = J[2][5] = 1.0;
  for (int bb = b; bb > 3 * leg; --bb) {
    v3 c = crs(k.axis[bb], sub(pt, k.p[bb]));
    const int vi = md->v_index[bb - 1];
    J[0][vi] = c.x; J[1][vi] = c.y; J[2][vi] = c.z;
  }
}

/* dense LU with partial pivoting, in place; returns 0 on success */
static int lu_factor(double* A, int n, int lda, int* piv) {
  for (int k = 0; k < n; ++k) {
    int p = k; double mx = fabs(A[k * lda + k]);
    for (int i = k + 1; i < n; ++i) if (fabs(A[i * lda + k]) > mx) { mx = fabs(A[i * lda + k]); p = i; }
    if (mx < 1e-300) return 1;
    piv[k] = p;
    if (p != k) for (int j = 0; j < n; ++j) { double t = A[k * lda + j]; A[k * lda + j] = A[p * lda + j]; A[p * lda + j] = t; }
    const double inv = 1.0 / A[k * lda + k];
    for (int i = k + 1; i < n; ++i) {
      const double f = A[i * lda + k] * inv;
      A[i * lda + k] = f;
      if (f != 0.0) for (int j = k + 1; j < n; ++j) A[i * lda + j] -= f * A[k * lda + j];
    }
  }
  return 0;
}
static void lu_solve(const double* A, int n, int lda, const int* piv, double* b) {
  for (int k = 0; k < n; ++k) if (piv[k] != k) { double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }   /* P b */
  for (int k = 0; k < n; ++k) for (int i = k + 1; i < n; ++i) b[i] -= A[i * lda + k] * b[k];           /* L y = P b */
  for (int k = n - 1; k >= 0; --k) { for (int j = k + 1; j < n; ++j) b[k] -= A[k * lda + j] * b[j]; b[k] /= A[k * lda + k]; }
}

/* min 1/2 x'Px + q'x  s.t. Ax = b, Gx <= h   (dense Mehrotra predictor-corrector) */
static int qp_ipm(int n, int me, int mi, const double* P, const double* qv, const double* A, const double* b, const double* G,
                  const double* h, double* x, int* iters_out) {
  double nu[MEMAX], s[MIMAX], lam[MIMAX], K[KMAX * KMAX], rhs[KMAX], rd[NMAX], re[MEMAX], ri[MIMAX], rc[MIMAX];
  double dsa[MIMAX], dla[MIMAX], best[NMAX];
  int piv[KMAX];
  const int nk = n + me;
  memset(x, 0, n * sizeof(double)); memset(nu, 0, sizeof(nu));
  for (int i = 0; i < mi; ++i) { s[i] = h[i] > 1.0 ? h[i] : 1.0; lam[i] = 1.0; }
  double best_score = 1e300, prev_mu = 1e300;
  int it, stall = 0, slow = 0;
  memcpy(best, x, n * sizeof(double));
  for (it = 1; it <= 100; ++it) {
    double score = 0.0, mu = 0.0;
    for (int i = 0; i < n; ++i) {
      double r = qv[i];
      for (int j = 0; j < n; ++j) r += P[i * n + j] * x[j];
      for (int j = 0; j < me; ++j) r += A[j * n + i] * nu[j];
      for (int j = 0; j < mi; ++j) r += G[j * n + i] * lam[j];
      rd[i] = r; if (fabs(r) > score) score = fabs(r);
    }
    for (int i = 0; i < me; ++i) { double r = -b[i]; for (int j = 0; j < n; ++j) r += A[i * n + j] * x[j]; re[i] = r; if (fabs(r) > score) score = fabs(r); }
    for (int i = 0; i < mi; ++i) { double r = s[i] - h[i]; for (int j = 0; j < n; ++j) r += G[i * n + j] * x[j]; ri[i] = r; if (fabs(r) > score) score = fabs(r); mu += lam[i] * s[i]; }
    if (mi) mu /= mi;
    /* jamming guard: when the complementarity gap stops shrinking for two iterations (a badly centred pair caps the step length), the next
       iterations are pure centring steps (sigma = 0.7, no second-order term) until it moves again */
    slow = (mi && it > 12 && mu > 0.9 * prev_mu) ? slow + 1 : 0;
    if (slow >= 2 || it > 40) stall = 1;      /* sticky: rare (a handful of instances per 4096) */
    prev_mu = mu;
    if (mu > score) score = mu;
    if (!(score == score)) break;
    if (score < best_score) { best_score = score; memcpy(best, x, n * sizeof(double)); }
    if (score < 1e-11) break;
    /* K = [[P + G' D G, A'], [A, 0]] */
    memset(K, 0, sizeof(double) * nk * nk);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) K[i * nk + j] = P[i * n + j];
    for (int c = 0; c < mi; ++c) {
      const double d = lam[c] / s[c];
      for (int i = 0; i < n; ++i) { const double gi = G[c * n + i]; if (gi == 0.0) continue; for (int j = 0; j < n; ++j) K[i * nk + j] += d * gi * G[c * n + j]; }
    }
    for (int i = 0; i < me; ++i) for (int j = 0; j < n; ++j) { K[(n + i) * nk + j] = A[i * n + j]; K[j * nk + n + i] = A[i * n + j]; }
    if (lu_factor(K, nk, nk, piv)) break;
    double alpha = 1.0, sigma = 0.0;
    for (int pass = 0; pass < (mi ? 2 : 1); ++pass) {
      for (int i = 0; i < mi; ++i) rc[i] = (pass == 0) ? lam[i] * s[i] : lam[i] * s[i] + (stall ? 0.0 : dsa[i] * dla[i]) - sigma * mu;
      for (int i = 0; i < n; ++i) {
        double r = -rd[i];
        for (int c = 0; c < mi; ++c) r += G[c * n + i] * (rc[c] - lam[c] * ri[c]) / s[c];
        rhs[i] = r;
      }
      for (int i = 0; i < me; ++i) rhs[n + i] = -re[i];
      lu_solve(K, nk, nk, piv, rhs);
      alpha = 1.0;
      for (int c = 0; c < mi; ++c) {
        double gdx = 0.0;
        for (int j = 0; j < n; ++j) gdx += G[c * n + j] * rhs[j];
        dsa[c] = -ri[c] - gdx;
        dla[c] = -(rc[c] + lam[c] * dsa[c]) / s[c];
        if (dsa[c] < 0 && -s[c] / dsa[c] < alpha) alpha = -s[c] / dsa[c];
        if (dla[c] < 0 && -lam[c] / dla[c] < alpha) alpha = -lam[c] / dla[c];
      }
      if (pass == 0 && mi) {
        double mua = 0.0;
        for (int c = 0; c < mi; ++c) mua += (lam[c] + alpha * dla[c]) * (s[c] + alpha * dsa[c]);
        mua /= mi;
        sigma = mu > 0 ? (mua / mu) * (mua / mu) * (mua / mu) : 0.0;
        if (stall && sigma < 0.7) sigma = 0.7;
      }
    }
    if (mi) { alpha *= 0.99; if (alpha > 1.0) alpha = 1.0; }
    for (int i = 0; i < n; ++i) x[i] += alpha * rhs[i];
    for (int i = 0; i < me; ++i) nu[i] += alpha * rhs[n + i];
    for (int c = 0; c < mi; ++c) { s[c] += alpha * dsa[c]; lam[c] += alpha * dla[c]; }
  }
  memcpy(x, best, n * sizeof(double));
  if (iters_out) *iters_out = it;
  return best_score < 1e-6 ? 0 : 1;
}

static void rpy_from_R(const double R[9], double rpy[3]) {
  rpy[0] = atan2(R[7], R[8]); rpy[1] = atan2(-R[6], hypot(R[0], R[3])); rpy[2] = atan2(R[3], R[0]);
}

/* one IDController.ControlLaw; returns 0 if the QP converged */
int oracle_id_step(const wbc_model* md, const wbc_params* pr, const double* q, const double* v, const double* traj,
                   const uint8_t* contact, double* tau, double* vd_out, double* f_out) {
  static const double zero[NV] = {0};
  double M[NV][NV], Cv[NV], tg[NV], e[NV], col[NV];
  /* CalcMassMatrixViaInverseDynamics: nv passes */
  for (int j = 0; j < NV; ++j) {
    memset(e, 0, sizeof(e)); e[j] = 1.0;
    inverse_dynamics(md, q, zero, e, 0, col);
    for (int i = 0; i < NV; ++i) M[i][j] = col[i];
  }
  inverse_dynamics(md, q, v, zero, 0, Cv);      /* CalcBiasTerm */
  inverse_dynamics(md, q, zero, zero, 1, tg);   /* -CalcGravityGeneralizedForces */
  double pf[4][3], J[4][3][NV], Jdv[4][3];
  for (int k = 0; k < 4; ++k) foot_quantities(md, q, v, k, pf[k], J[k], Jdv[k]);
  /* body pose */
  kin_t kin; kinematics(md, q, v, zero, &kin);
  double rpy[3]; rpy_from_R(kin.R[0], rpy);
  const double cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
  const double N[3][3] = {{cy * cp, -sy, 0}, {sy * cp, cy, 0}, {-sp, 0, 1}};
  double rpyd[3];
  rpyd[0] = (cy * v[0] + sy * v[1]) / cp; rpyd[1] = -sy * v[0] + cy * v[1]; rpyd[2] = v[2] + sp * rpyd[0];
  double ades[6], rdd[3];
  for (int i = 0; i < 3; ++i) {
    rdd[i] = traj[15 + i] - pr->id_kp_body_rpy * (rpy[i] - traj[9 + i]) - pr->id_kd_body_rpy * (rpyd[i] - traj[12 + i]);
    ades[3 + i] = traj[6 + i] - pr->id_kp_body_p * (q[4 + i] - traj[i]) - pr->id_kd_body_p * (v[3 + i] - traj[3 + i]);
  }
  for (int i = 0; i < 3; ++i) ades[i] = N[i][0] * rdd[0] + N[i][1] * rdd[1] + N[i][2] * rdd[2];
  int cont[4], nc = 0;
  for (int k = 0; k < 4; ++k) if (contact[k]) cont[nc++] = k;
  const int n = 30 + 3 * nc, me = 18 + 3 * nc, mi = 4 * nc;
  double P[NMAX * NMAX], qv[NMAX], A[MEMAX * NMAX], b[MEMAX], G[MIMAX * NMAX], h[MIMAX], x[NMAX];
  memset(P, 0, sizeof(P)); memset(qv, 0, sizeof(qv)); memset(A, 0, sizeof(A)); memset(b, 0, sizeof(b));
  memset(G, 0, sizeof(G)); memset(h, 0, sizeof(h));
  for (int i = 0; i < 6; ++i) { P[i * n + i] += pr->id_w_body; qv[i] += pr->id_w_body * (0.0 - ades[i]); }   /* J_body = [I6 0], Jdv_body = 0 */
  for (int k = 0; k < 4; ++k) {
    if (contact[k]) continue;
    for (int r = 0; r < 3; ++r) {
      const double pv = J[k][r][0] * v[0];
      (void)pv;
      double pd = 0.0;
      for (int j = 0; j < NV; ++j) pd += J[k][r][j] * v[j];
      const double as = traj[42 + 3 * k + r] - pr->id_kp_foot * (pf[k][r] - traj[18 + 3 * k + r]) - pr->id_kd_foot * (pd - traj[30 + 3 * k + r]);
      for (int i = 0; i < NV; ++i) {
        qv[i] += pr->id_w_foot * J[k][r][i] * (Jdv[k][r] - as);
        for (int j = 0; j < NV; ++j) P[i * n + j] += pr->id_w_foot * J[k][r][i] * J[k][r][j];
      }
    }
  }
  for (int i = 0; i < NV; ++i) P[i * n + i] += pr->reg_vd;
  for (int i = 0; i < NU; ++i) P[(18 + i) * n + 18 + i] += pr->reg_tau;
  for (int i = 0; i < 3 * nc; ++i) P[(30 + i) * n + 30 + i] += pr->reg_f;
  for (int i = 0; i < NV; ++i) { for (int j = 0; j < NV; ++j) A[i * n + j] = M[i][j]; b[i] = -Cv[i] - tg[i]; }
  for (int k = 0; k < NU; ++k) A[md->v_index[k] * n + 18 + md->act_index[k]] = -1.0;                              /* -B */
  for (int s = 0; s < nc; ++s) {
    const int k = cont[s];
    for (int r = 0; r < 3; ++r) {
      double pd = 0.0;
      for (int j = 0; j < NV; ++j) { A[j * n + 30 + 3 * s + r] = -J[k][r][j]; A[(18 + 3 * s + r) * n + j] = J[k][r][j]; pd += J[k][r][j] * v[j]; }
      b[18 + 3 * s + r] = -pr->contact_damping * pd - Jdv[k][r];
    }
    const double mu = pr->mu;
    const double Ai[4][3] = {{1, 0, -mu}, {-1, 0, -mu}, {0, 1, -mu}, {0, -1, -mu}};
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 3; ++c) G[(4 * s + r) * n + 30 + 3 * s + c] = Ai[r][c];
  }
  int iters = 0;
  const int rc = qp_ipm(n, me, mi, P, qv, A, b, G, h, x, &iters);
  for (int i = 0; i < NU; ++i) tau[i] = x[18 + i];
  if (vd_out) memcpy(vd_out, x, NV * sizeof(double));
  if (f_out) { memset(f_out, 0, 12 * sizeof(double)); for (int s = 0; s < nc; ++s) for (int r = 0; r < 3; ++r) f_out[3 * cont[s] + r] = x[30 + 3 * s + r]; }
  return rc;
}

typedef struct {
  const wbc_model* md; const wbc_params* pr; const double *q, *v, *traj; const uint8_t* contact;
  double *tau, *vd, *f; int32_t* status; int64_t lo, hi;
} job_t;
static void* worker(void* p) {
  job_t* j = (job_t*)p;
  for (int64_t i = j->lo; i < j->hi; ++i)
    j->status[i] = oracle_id_step(j->md, j->pr, j->q + 19 * i, j->v + 18 * i, j->traj + 54 * i, j->contact + 4 * i, j->tau + 12 * i,
                                  j->vd ? j->vd + 18 * i : NULL, j->f ? j->f + 12 * i : NULL);
  return NULL;
}

/* n instances split evenly over `threads` host threads */
int oracle_id_batch(const wbc_model* md, const wbc_params* pr, int64_t n, const double* q, const double* v, const double* traj,
                    const uint8_t* contact, double* tau, double* vd, double* f, int32_t* status, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256]; job_t jobs[256];
  for (int t = 0; t < threads; ++t) {
    job_t j = {md, pr, q, v, traj, contact, tau, vd, f, status, n * t / threads, n * (t + 1) / threads};
    jobs[t] = j;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* CalcDynamics + foot queries for one state (parity tests of the port itself) */
void oracle_dynamics(const wbc_model* md, const double* q, const double* v, double* M, double* Cv, double* tg, double* Jfeet,
                     double* Jdv, double* pfeet) {
  static const double zero[NV] = {0};
  double e[NV], col[NV];
  for (int j = 0; j < NV; ++j) {
    memset(e, 0, sizeof(e)); e[j] = 1.0;
    inverse_dynamics(md, q, zero, e, 0, col);
    for (int i = 0; i < NV; ++i) M[i * NV + j] = col[i];
  }
  inverse_dynamics(md, q, v, zero, 0, Cv);
  inverse_dynamics(md, q, zero, zero, 1, tg);
  for (int k = 0; k < 4; ++k) foot_quantities(md, q, v, k, pfeet + 3 * k, (double(*)[NV])(Jfeet + 54 * k), Jdv + 3 * k);
}
