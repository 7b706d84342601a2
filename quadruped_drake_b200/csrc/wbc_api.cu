// wbc_api.cu — C ABI (include/wbc.h) and kernel launches of the batched whole-body controller.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

// named barrier `id` over `threads` threads of the CTA (a multiple of 32; skipped when at most one warp takes part)
#ifndef WBC_PC_MEET
#define WBC_PC_MEET 1
#endif
#if WBC_PC_MEET
#define WBC_CTA_MEET(id, threads)                                                                    \
  do {                                                                                               \
    const int nt_ = (threads);                                                                       \
    if (nt_ > 32) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nt_) : "memory");                   \
  } while (0)
#endif
#include "wbc_device.cuh"
#include "wbc_wire.cuh"
#include "wbc_traj.cuh"
#include "wbc_rollout.cuh"
#include "wbc_plant.cuh"
#include <vector>

namespace {

// Warps (= instances) per CTA of the reduce / dynamics kernels. One warp per CTA puts the per-warp block at a compile-time
// shared-memory address, which frees the registers that held its base (spill 536 -> 44 B; 4-warp CTAs measured 3-8 % slower
// on device-resident buffers). The 4-warp shape is kept for page-locked HOST buffers (zero-copy path of wbc_step_host): there
// the CTA stages all four input arrays of its 4 consecutive instances with one bulk copy each - few large requests, which is
// what the host link wants (e2e 25.3 vs 22.3 M steps/s at 4096).
constexpr int WARPS = 1;         // device-resident buffers
#ifndef WBC_WARPS_HOST
#define WBC_WARPS_HOST 4
#endif
constexpr int WARPS_HOST = WBC_WARPS_HOST;    // host-mapped buffers (ID / CLF)
#ifndef WBC_WARPS_PC
#define WBC_WARPS_PC 6
#endif
constexpr int WARPS_PC = WBC_WARPS_PC;        // PC / MPTC reduce kernel
#ifndef WBC_PC_MIN_WARPS
#define WBC_PC_MIN_WARPS 12      // resident warps per SM the PC reduce kernel is compiled for (168 registers)
#endif
#ifndef WBC_MIN_WARPS
#define WBC_MIN_WARPS 16         // resident warps per SM the reduce kernels are compiled for (128 registers)
#endif

struct alignas(16) DevConst { wbc_model md; wbc_params pr; wbc::Derived dv; };

// CTA-level input staging of the reduce kernels: the rows of the CTA's W consecutive instances are contiguous in the caller's
// arrays, so an array whose W rows are a multiple of 16 bytes arrives with ONE bulk copy instead of 8-byte loads per lane
// (v: 144 B and traj: 432 B rows always; q: 152 B and contact: 4 B rows only for W = 4).
template <int W>
struct alignas(16) InStage {
  alignas(16) double q[W * WBC_NQ];
  alignas(16) double v[W * WBC_NV];
  alignas(16) double traj[W * WBC_NTRAJ];
  alignas(16) uint8_t contact[W * 4];
  alignas(16) unsigned long long mbar;
  static constexpr bool BULK_Q = (W * WBC_NQ * 8) % 16 == 0, BULK_C = (W * 4) % 16 == 0;
};
static_assert((WBC_NV * 8) % 16 == 0 && (WBC_NTRAJ * 8) % 16 == 0, "v / traj rows must be multiples of 16 bytes");

template <int W>
struct SmemLayoutT {
  DevConst dc;
  InStage<W> in;
  wbc::WarpSmem w[W];
};
template <int W>
struct SmemLayoutPCT {     // PC controller / Coriolis entry: extra operational-space workspace per warp
  DevConst dc;
  InStage<W> in;
  wbc::WarpSmem w[W];
  wbc::PcSmem pc[W];
};
using SmemLayout = SmemLayoutT<WARPS>;
using SmemLayoutPC = SmemLayoutPCT<WARPS>;

template <typename L>
__device__ __forceinline__ const DevConst& stage_consts(L* sm, const DevConst* g) {
  // model + gains into shared memory once per CTA (lane-divergent table lookups stay on chip), 16 bytes per thread and trip
  static_assert(sizeof(DevConst) % 16 == 0, "DevConst is staged with 16-byte vectors");
  const int nvec = sizeof(DevConst) / 16;
  const uint4* src = reinterpret_cast<const uint4*>(g);
  uint4* dst = reinterpret_cast<uint4*>(&sm->dc);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return sm->dc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Split path (ID / CLF): the step is launched as two kernels with their own register / occupancy budgets.
//   wbc_reduce_kernel  phases 0-4 (dynamics, equality elimination, reduced rows): register hungry (128 regs, 16 warps / SM);
//                      leaves [Y | cw | ct] + 16 carry words per instance in a hand-over record (4224 B, device scratch).
//   wbc_solve_kernel   phases 5-7 (reduced Hessian, Cholesky, Goldfarb-Idnani, outputs): 72 registers and 7.8 KB of shared
//                      memory per warp -> 28 warps / SM, so the dependent-instruction latency of the active-set iterations is
//                      hidden by 1.75x more resident instances than the fused kernel could hold.
// The 4 KB block moves shared -> global and global -> shared with one bulk asynchronous copy each (TMA 1-D), issued by one
// lane; the solve warp waits on an mbarrier.
#ifndef WBC_SOLVE_WARPS
#define WBC_SOLVE_WARPS 1      // one warp per CTA: the shared-memory block sits at a compile-time address, so no per-warp base
#endif                         // (thread index -> warp -> offset) has to be kept live or re-derived under the 72-register budget
#ifndef WBC_SOLVE_CTAS
#define WBC_SOLVE_CTAS (21 / WBC_SOLVE_WARPS)
#endif
constexpr int SOLVE_WARPS = WBC_SOLVE_WARPS;
template <bool VD> struct SmemLayoutSolveT { wbc::SolveSmemT<VD> w[SOLVE_WARPS]; };
static_assert(sizeof(wbc::SolveSmem) % 16 == 0 && sizeof(wbc::SolveSmemVd) % 16 == 0, "SolveSmem must keep 16-byte alignment per warp");
static_assert(offsetof(wbc::WarpSmem, Y) % 16 == 0 && sizeof(wbc::WarpSmem) % 16 == 0 && sizeof(DevConst) % 16 == 0, "bulk copy alignment");
static_assert(offsetof(wbc::WarpSmem, cw) == offsetof(wbc::WarpSmem, Y) + sizeof(double) * wbc::YROWS * wbc::YS &&
              offsetof(wbc::WarpSmem, ct) == offsetof(wbc::WarpSmem, cw) + sizeof(double) * wbc::YROWS, "[Y | cw | ct] must be contiguous");
static_assert(offsetof(wbc::SolveSmem, cw) == sizeof(double) * wbc::YROWS * wbc::YS &&
              offsetof(wbc::SolveSmem, ct) == offsetof(wbc::SolveSmem, cw) + sizeof(double) * wbc::YROWS, "[Y | cw | ct] must be contiguous");
constexpr unsigned REC_Y_BYTES = wbc::REC_Y * sizeof(double);

// Programmatic dependent launch: the solve kernel is launched with the programmatic-stream-serialization attribute, so its
// CTAs become resident while the last reduce CTAs are still running (they have all started by then) and wait here; this
// hides the launch latency and CTA ramp of the second kernel. Both instructions are no-ops in an ordinary launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// Lane index the compiler cannot re-derive from the special register: under the register caps of the step kernels it
// otherwise rematerialises every lane-dependent shared-memory address with an S2R (a slow special-register read) + IMAD
// pair at each use instead of keeping one register live (WBC_OPAQUE_LANE=0 restores the plain form for A/B runs).
#ifndef WBC_OPAQUE_LANE
#define WBC_OPAQUE_LANE 1
#endif
__device__ __forceinline__ int lane_index() {
  int lane = (int)(threadIdx.x & 31);
#if WBC_OPAQUE_LANE
  asm volatile("" : "+r"(lane));
#endif
  return lane;
}
// shared -> global bulk copy of `bytes` (multiple of 16), issued by the calling thread; returns once the source may be reused
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the warp -> visible to the copy engine
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// global -> shared bulk copy completing on an mbarrier (armed here by the issuing thread)
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned b = smem_addr(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sdst)),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void bulk_load_on(void* sdst, const void* gsrc, unsigned bytes, unsigned bar_smem) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sdst)),
               "l"(gsrc), "r"(bytes), "r"(bar_smem)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = smem_addr(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(b), "r"(parity)
      : "memory");
}

// Hand-over record of one instance: [Y | cw | ct] by one bulk copy + the carry words (called by one lane).
__device__ __forceinline__ void store_record(double* r, const wbc::StepCarry& c, const wbc::WarpSmem& s) {
  double2* m = reinterpret_cast<double2*>(r + wbc::REC_Y);
  m[0] = make_double2((double)c.status, (double)c.cmask);
  m[1] = make_double2((double)c.nf, (double)c.nextra);
  m[2] = make_double2(c.ok ? 1.0 : 0.0, c.pc_ok ? 1.0 : 0.0);
  m[3] = make_double2(c.extra_bound, c.err);
  m[4] = make_double2(c.Vl, c.PFl);
  m[5] = make_double2(c.csum, c.Vpc);
  if (c.ok) bulk_store(r, &s.Y[0][0], REC_Y_BYTES);
}

// Issues the CTA's four input bulk copies (thread 0) - call before stage_consts so that they fly during the constant staging -
// and, after the CTA barrier inside stage_consts, waits for them. Returns false when this CTA must use the per-lane path
// (tail CTA with fewer than WARPS instances, or buffers that are not 16-byte aligned: `bulk_ok` from the host).
template <int W>
__device__ __forceinline__ bool stage_inputs_issue(InStage<W>& in, const wbc::StepArgs& a, bool bulk_ok) {
  constexpr bool BQ = InStage<W>::BULK_Q, BC = InStage<W>::BULK_C;
  const long long first = (long long)blockIdx.x * W;
  const bool full = bulk_ok && first + W <= a.n;
  if (full && threadIdx.x == 0) {
    const unsigned b = smem_addr(&in.mbar);
    const unsigned bytes = W * (WBC_NV + WBC_NTRAJ) * 8 + (BQ ? W * WBC_NQ * 8 : 0) + (BC ? W * 4 : 0);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (BQ) bulk_load_on(in.q, a.q + first * WBC_NQ, W * WBC_NQ * 8, b);
    bulk_load_on(in.v, a.v + first * WBC_NV, W * WBC_NV * 8, b);
    bulk_load_on(in.traj, a.traj + first * WBC_NTRAJ, W * WBC_NTRAJ * 8, b);
    if (BC) bulk_load_on(in.contact, a.contact + first * 4, W * 4, b);
  }
  return full;
}
template <int W>
__device__ __forceinline__ wbc::StagedInputs staged_rows(const InStage<W>& in, int warp) {
  return wbc::StagedInputs{InStage<W>::BULK_Q ? in.q + warp * WBC_NQ : nullptr, in.v + warp * WBC_NV, in.traj + warp * WBC_NTRAJ,
                           InStage<W>::BULK_C ? in.contact + warp * 4 : nullptr};
}

template <int KIND, int W>
__global__ void __launch_bounds__(W * 32, WBC_MIN_WARPS / W) wbc_reduce_kernel(const DevConst* __restrict__ gdc, wbc::StepArgs a,
                                                                            double* __restrict__ rec, double* __restrict__ vdmap, int bulk_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayoutT<W>* sm = reinterpret_cast<SmemLayoutT<W>*>(smem_raw);
  pdl_launch_dependents();
  const bool staged = stage_inputs_issue(sm->in, a, bulk_ok != 0);
  const DevConst& dc = stage_consts(sm, gdc);   // (reading the tables through L1 instead of staging them measured neutral)
  const int warp = W == 1 ? 0 : (int)(threadIdx.x >> 5), lane = lane_index();
  const long long inst = (long long)blockIdx.x * W + warp;
  if (inst >= a.n) return;
  wbc::WarpSmem& s = sm->w[warp];
  wbc::StepCarry c;
  const wbc::StagedInputs in = staged_rows(sm->in, warp);
  if (staged) mbar_wait(&sm->in.mbar, 0);
  wbc::reduce_instance<KIND>(s, dc.md, dc.pr, dc.dv, a, inst, lane, c, nullptr, vdmap ? vdmap + inst * wbc::VDMAP_DOUBLES : nullptr,
                             staged ? &in : nullptr);
  __syncwarp();
  if (lane == 0) store_record(rec + inst * wbc::REC_DOUBLES, c, s);
}

// VD: the accelerations are requested (rollout, debug outputs) - the Cholesky factor is kept for the recovery of w, which
// costs 1.3 KB more shared memory per warp (24 instead of 28 resident warps).
template <int KIND, bool VD>
__global__ void __launch_bounds__(SOLVE_WARPS * 32, VD ? (24 / SOLVE_WARPS) : WBC_SOLVE_CTAS) wbc_solve_kernel(
    const DevConst* __restrict__ gdc, wbc::StepArgs a, const double* __restrict__ rec, const double* __restrict__ vdmap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayoutSolveT<VD>* sm = reinterpret_cast<SmemLayoutSolveT<VD>*>(smem_raw);
  const int warp = SOLVE_WARPS == 1 ? 0 : (int)(threadIdx.x >> 5), lane = lane_index();
  const long long inst = (long long)blockIdx.x * SOLVE_WARPS + warp;
  if (inst >= a.n) return;
  wbc::SolveSmemT<VD>& s = sm->w[warp];
  const double* r = rec + inst * wbc::REC_DOUBLES;
  const double2* m = reinterpret_cast<const double2*>(r + wbc::REC_Y);
  pdl_wait();                      // the records of the reduce kernel are complete and visible from here on
  const double2 m2 = m[2];
  const bool ok = m2.x != 0.0;
  if (ok && lane == 0) bulk_load(&s.Y[0][0], r, REC_Y_BYTES, &s.mbar);
  const double2 m0 = m[0], m1 = m[1], m3 = m[3], m4 = m[4], m5 = m[5];
  wbc::StepCarry c;
  c.status = (int)m0.x; c.cmask = (unsigned)m0.y; c.nf = (int)m1.x; c.nextra = (int)m1.y;
  c.ok = ok; c.pc_ok = m2.y != 0.0; c.extra_bound = m3.x; c.err = m3.y; c.Vl = m4.x; c.PFl = m4.y; c.csum = m5.x; c.Vpc = m5.y;
  __syncwarp();                    // the barrier is initialised before any lane polls it
  if (ok) mbar_wait(&s.mbar, 0);
  wbc::solve_instance<KIND, wbc::SolveSmemT<VD>>(s, gdc->md, gdc->pr, a, inst, lane, c, vdmap ? vdmap + inst * wbc::VDMAP_DOUBLES : nullptr, r);
}

// PC / MPTC reduce half: same hand-over, extra operational-space workspace per warp (168 registers, 12 warps / SM).
template <int W>
__global__ void __launch_bounds__(W * 32, WBC_PC_MIN_WARPS / W) wbc_reduce_pc_kernel(const DevConst* __restrict__ gdc, wbc::StepArgs a,
                                                                    double* __restrict__ rec, double* __restrict__ vdmap, int bulk_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayoutPCT<W>* sm = reinterpret_cast<SmemLayoutPCT<W>*>(smem_raw);
  pdl_launch_dependents();
  const bool staged = stage_inputs_issue(sm->in, a, bulk_ok != 0);
  const DevConst& dc = stage_consts(sm, gdc);
  const int warp = W == 1 ? 0 : (int)(threadIdx.x >> 5), lane = lane_index();
  const long long inst = (long long)blockIdx.x * W + warp;
  if (inst >= a.n) return;
  wbc::WarpSmem& s = sm->w[warp];
  wbc::StepCarry c;
  const wbc::StagedInputs in = staged_rows(sm->in, warp);
  if (staged) mbar_wait(&sm->in.mbar, 0);
  {
    // meeting points of the CTA's warps ahead of the dynamics passes (wbc_device.cuh: WBC_CTA_MEET): every warp with an instance
    // takes the state pass, every warp whose instance has a stance foot takes the two polarisation passes
    const long long first = (long long)blockIdx.x * W;
    const int n_act = (int)((a.n - first) < W ? (a.n - first) : W);
    int n_pc = 0;
    for (int w = 0; w < n_act; ++w) {
      const uint8_t* cp = (staged && in.contact) ? sm->in.contact + w * 4 : a.contact + (first + w) * 4;
      n_pc += (cp[0] | cp[1] | cp[2] | cp[3]) ? 1 : 0;
    }
    sm->pc[warp].sync_all = 32 * n_act;
    sm->pc[warp].sync_pc = 32 * n_pc;
  }
  wbc::reduce_instance<WBC_CTRL_PC>(s, dc.md, dc.pr, dc.dv, a, inst, lane, c, &sm->pc[warp], vdmap ? vdmap + inst * wbc::VDMAP_DOUBLES : nullptr,
                                    staged ? &in : nullptr);
  __syncwarp();
  if (lane == 0) store_record(rec + inst * wbc::REC_DOUBLES, c, s);
}

__global__ void __launch_bounds__(WARPS * 32, 12 / WARPS) wbc_coriolis_kernel(const DevConst* __restrict__ gdc, const double* q,
                                                                     const double* v, double* Cout, double* Jdout, long long n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayoutPC* sm = reinterpret_cast<SmemLayoutPC*>(smem_raw);
  const DevConst& dc = stage_consts(sm, gdc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long inst = (long long)blockIdx.x * WARPS + warp;
  if (inst < n) wbc::coriolis_instance(sm->w[warp], sm->pc[warp], dc.md, q, v, Cout, Jdout, inst, lane);
}

__global__ void __launch_bounds__(WARPS * 32) wbc_dynamics_kernel(const DevConst* __restrict__ gdc, const double* q,
                                                                   const double* v, wbc::DynOut o, long long n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayout* sm = reinterpret_cast<SmemLayout*>(smem_raw);
  const DevConst& dc = stage_consts(sm, gdc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long inst = (long long)blockIdx.x * WARPS + warp;
  if (inst < n) wbc::dynamics_instance(sm->w[warp], dc.md, q, v, o, inst, lane);
}

// One time step of the simulated robot on the ground (wbc_plant.cuh), one single-warp CTA per robot.
struct SmemLayoutPlant { DevConst dc; wbc::WarpSmem w; wbcplant::PlantSmem p; };
#ifndef WBC_PLANT_CTAS
#define WBC_PLANT_CTAS 16     // 128 registers: 22.2 M robot-steps/s at 4096 robots against 19.2 M at 8 CTAs (202 registers)
#endif
__global__ void __launch_bounds__(32, WBC_PLANT_CTAS) wbc_plant_kernel(const DevConst* __restrict__ gdc, wbcplant::PlantArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayoutPlant* sm = reinterpret_cast<SmemLayoutPlant*>(smem_raw);
  const DevConst& dc = stage_consts(sm, gdc);
  const int lane = lane_index();
  const long long inst = blockIdx.x;
  if (inst < a.n) wbcplant::plant_step_instance(sm->w, sm->p, dc.md, a, inst, lane);
}

__global__ void wbc_pd_kernel(const DevConst* __restrict__ gdc, const double* q, const double* v, double* tau, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * WBC_NU) wbc::pd_element(gdc->md, gdc->pr, q, v, tau, t / WBC_NU, (int)(t % WBC_NU));
}

// Register-resident DFMA loop: 8 independent chains per thread, 2 flop per FMA.
__global__ void fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

constexpr int WBC_NSLOT = 4;
struct wbc_handle {
  int device = 0;
  DevConst* d_const = nullptr;
  wbc_params params;
  std::string err;
  int64_t launches = 0;
  int sm_count = 148;
  int* d_tau_map = nullptr;            // message slot (velocity order) -> actuator index, for wbc_lcm_encode_robot_state
  cudaStream_t stream = nullptr;       // used by the host entry points
  cudaStream_t stream2 = nullptr;      // second lane of the chunked wbc_step_host pipeline
  // device staging for wbc_step_host
  int64_t cap = 0;
  double *d_q = nullptr, *d_v = nullptr, *d_traj = nullptr, *d_tau = nullptr, *d_metrics = nullptr, *d_vd = nullptr,
         *d_f = nullptr, *d_info = nullptr, *d_lam = nullptr;
  uint8_t* d_contact = nullptr;
  int32_t* d_status = nullptr;
  // scratch of wbc_rollout (sized for ro_cap instances)
  int64_t ro_cap = 0;
  double *ro_traj = nullptr, *ro_vd = nullptr, *ro_metrics = nullptr, *ro_tau = nullptr, *ro_t = nullptr;
  uint8_t* ro_contact = nullptr;
  int32_t* ro_status = nullptr;
  int* ro_counter = nullptr;
  // hand-over records of the split step (reduce -> solve), one slot per internal stream lane
  cudaEvent_t prof_ev[3] = {nullptr, nullptr, nullptr};   // wbc_profile_step: before reduce / between / after solve
  bool prof_on = false;
  bool side_by_side = false;                              // set while a step is issued as several concurrent chains
  bool host_mapped = false;                               // set around the zero-copy launches of wbc_step_host
  // per-slot ordering of the scratch users: the stream of the last step that used the slot and an event to chain a new one
  cudaStream_t last_stream[WBC_NSLOT] = {};
  bool slot_used[WBC_NSLOT] = {};
  cudaEvent_t order_ev = nullptr, join_ev = nullptr;
  double* d_rec[WBC_NSLOT] = {};
  double* d_vdmap[WBC_NSLOT] = {};
  int64_t rec_cap[WBC_NSLOT] = {}, vdmap_cap[WBC_NSLOT] = {};
  cudaStream_t xstream[WBC_NSLOT] = {};                  // lanes 2.. of a step issued as more than two chunks (created on demand)
};

#define WBC_CUDA(h, call)                                                                       \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      return WBC_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

static int fail_arg(wbc_handle* h, const char* msg) {
  if (h) h->err = msg;
  return WBC_ERR_ARG;
}

extern "C" int wbc_default_params(wbc_params* p) {
  if (!p) return WBC_ERR_ARG;
  memset(p, 0, sizeof(*p));
  p->id_kp_body_p = 500; p->id_kd_body_p = 50; p->id_kp_body_rpy = 500; p->id_kd_body_rpy = 50;
  p->id_kp_foot = 100; p->id_kd_foot = 20; p->id_w_body = 10; p->id_w_foot = 1;
  p->clf_q_body_p = 5000; p->clf_q_body_pd = 200; p->clf_q_body_rpy = 5000; p->clf_q_body_rpyd = 200;
  p->clf_q_foot_p = 200; p->clf_q_foot_pd = 20; p->clf_r = 1; p->clf_w_delta = 1000;
  p->pc_kp_body_p = 100; p->pc_kd_body_p = 10; p->pc_kp_body_rpy = 100; p->pc_kd_body_rpy = 10;
  p->pc_kp_foot = 200; p->pc_kd_foot = 20; p->pc_w_body = 10; p->pc_w_foot = 1;
  p->mu = 0.7; p->contact_damping = 100; p->reg_f = 1e-6; p->reg_tau = 0; p->reg_vd = 0;
  p->torque_limits = 0; p->max_iter = 200;
  p->pd_kp = 30; p->pd_kd = 1.5; p->pd_clip = 150;
  for (int l = 0; l < 4; ++l) { p->pd_q_nom[3 * l] = 0.0; p->pd_q_nom[3 * l + 1] = -0.8; p->pd_q_nom[3 * l + 2] = 1.6; }   // basic_controller.py:335-340
  return WBC_OK;
}

static int set_smem_attr(wbc_handle* h) {
  const int bytes = (int)sizeof(SmemLayout);
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_reduce_kernel<WBC_CTRL_ID, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_reduce_kernel<WBC_CTRL_CLF, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_reduce_kernel<WBC_CTRL_ID, WARPS_HOST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutT<WARPS_HOST>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_reduce_kernel<WBC_CTRL_CLF, WARPS_HOST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutT<WARPS_HOST>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_reduce_pc_kernel<WARPS_PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutPCT<WARPS_PC>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_solve_kernel<WBC_CTRL_ID, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutSolveT<false>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_solve_kernel<WBC_CTRL_CLF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutSolveT<false>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_solve_kernel<WBC_CTRL_PC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutSolveT<false>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_solve_kernel<WBC_CTRL_ID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutSolveT<true>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_solve_kernel<WBC_CTRL_CLF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutSolveT<true>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_solve_kernel<WBC_CTRL_PC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutSolveT<true>)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_coriolis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutPC)));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_dynamics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  WBC_CUDA(h, cudaFuncSetAttribute(wbc_plant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayoutPlant)));
  return WBC_OK;
}

extern "C" int wbc_create(const wbc_model* model, const wbc_params* params, int device, wbc_handle** out) {
  if (!model || !out) return WBC_ERR_ARG;
  *out = nullptr;
  wbc_handle* h = new (std::nothrow) wbc_handle();
  if (!h) return WBC_ERR_NOMEM;
  h->device = device;
  if (params) h->params = *params; else wbc_default_params(&h->params);
  int rc = WBC_OK;
  auto bail = [&](int code) { *out = h; return code; };  // hand the handle back so wbc_last_error works
  if (h->params.reg_vd != 0.0) { h->err = "reg_vd is not supported yet (must be 0)"; return bail(WBC_ERR_ARG); }
  if (!(h->params.reg_f > 0.0)) { h->err = "reg_f must be > 0 (tie-break that makes the QP strictly convex)"; return bail(WBC_ERR_ARG); }
  for (int k = 0; k < WBC_NU; ++k) {
    if (model->v_index[k] < 6 || model->v_index[k] >= WBC_NV || model->act_index[k] < 0 || model->act_index[k] >= WBC_NU) {
      h->err = "wbc_model: v_index/act_index out of range"; return bail(WBC_ERR_ARG);
    }
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return bail(WBC_ERR_CUDA); }
  DevConst hc;
  hc.md = *model; hc.pr = h->params;
  wbc::derive_constants(h->params, hc.dv);
  e = cudaMalloc(&h->d_const, sizeof(DevConst));
  if (e != cudaSuccess) { h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return bail(WBC_ERR_CUDA); }
  e = cudaMemcpy(h->d_const, &hc, sizeof(DevConst), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { h->err = std::string("cudaMemcpy: ") + cudaGetErrorString(e); return bail(WBC_ERR_CUDA); }
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
  // lanes 2.. of a chunked step and the fork / join events: created here, not on first use, so that a step never calls a
  // resource-creating API while some other stream of the process is being captured
  for (int l = 2; l < WBC_NSLOT && e == cudaSuccess; ++l) e = cudaStreamCreateWithFlags(&h->xstream[l], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming);
  if (e != cudaSuccess) { h->err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); return bail(WBC_ERR_CUDA); }
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    int tmap[WBC_NU];
    for (int k = 0; k < WBC_NU; ++k) tmap[model->v_index[k] - 6] = model->act_index[k];
    e = cudaMalloc(&h->d_tau_map, sizeof(tmap));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tau_map, tmap, sizeof(tmap), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { h->err = std::string("tau map upload: ") + cudaGetErrorString(e); return bail(WBC_ERR_CUDA); }
  }
  rc = set_smem_attr(h);
  *out = h;
  return rc;
}

static void free_staging(wbc_handle* h) {
  cudaFree(h->d_q); cudaFree(h->d_v); cudaFree(h->d_traj); cudaFree(h->d_tau); cudaFree(h->d_metrics);
  cudaFree(h->d_vd); cudaFree(h->d_f); cudaFree(h->d_info); cudaFree(h->d_lam); cudaFree(h->d_contact); cudaFree(h->d_status);
  h->d_q = h->d_v = h->d_traj = h->d_tau = h->d_metrics = h->d_vd = h->d_f = h->d_info = h->d_lam = nullptr;
  h->d_contact = nullptr; h->d_status = nullptr; h->cap = 0;
}

extern "C" int wbc_destroy(wbc_handle* h) {
  if (!h) return WBC_OK;
  cudaSetDevice(h->device);
  free_staging(h);
  if (h->d_const) cudaFree(h->d_const);
  if (h->d_tau_map) cudaFree(h->d_tau_map);
  cudaFree(h->ro_traj); cudaFree(h->ro_vd); cudaFree(h->ro_metrics); cudaFree(h->ro_tau); cudaFree(h->ro_t);
  cudaFree(h->ro_contact); cudaFree(h->ro_status); cudaFree(h->ro_counter);
  for (int i = 0; i < WBC_NSLOT; ++i) { cudaFree(h->d_rec[i]); cudaFree(h->d_vdmap[i]); if (h->xstream[i]) cudaStreamDestroy(h->xstream[i]); }
  for (int i = 0; i < 3; ++i) if (h->prof_ev[i]) cudaEventDestroy(h->prof_ev[i]);
  if (h->order_ev) cudaEventDestroy(h->order_ev);
  if (h->join_ev) cudaEventDestroy(h->join_ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  delete h;
  return WBC_OK;
}

extern "C" const char* wbc_last_error(const wbc_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t wbc_launch_count(const wbc_handle* h) { return h ? h->launches : 0; }

extern "C" int wbc_dynamics(wbc_handle* h, int64_t n, const double* q, const double* v, double* M, double* Cv,
                            double* taug, double* Jfeet, double* Jdv, double* pfeet, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!q || !v))) return fail_arg(h, "wbc_dynamics: null input");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  wbc::DynOut o{M, Cv, taug, Jfeet, Jdv, pfeet};
  const unsigned grid = (unsigned)((n + WARPS - 1) / WARPS);
  wbc_dynamics_kernel<<<grid, WARPS * 32, sizeof(SmemLayout), (cudaStream_t)stream>>>(h->d_const, q, v, o, n);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_coriolis(wbc_handle* h, int64_t n, const double* q, const double* v, double* Cm, double* Jd, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!q || !v))) return fail_arg(h, "wbc_coriolis: null input");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((n + WARPS - 1) / WARPS);
  wbc_coriolis_kernel<<<grid, WARPS * 32, sizeof(SmemLayoutPC), (cudaStream_t)stream>>>(h->d_const, q, v, Cm, Jd, n);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

namespace {
// Scratch device buffers of the host-buffer debug / codec entries: freed when the scope ends (early error returns included).
struct DevScratch {
  void* p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // at most 8 buffers per scratch
  int used = 0;
  cudaError_t err = cudaSuccess;
  template <typename T> T* in(const T* host, size_t count, cudaStream_t st) {      // NULL stays NULL
    if (!host || err != cudaSuccess) return nullptr;
    T* d = out<T>(host, count);
    if (d) err = cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, st);
    return d;
  }
  template <typename T> T* out(const T* host, size_t count) {
    if (!host || err != cudaSuccess) return nullptr;
    void* d = nullptr;
    err = cudaMalloc(&d, count * sizeof(T) > 0 ? count * sizeof(T) : 1);
    if (err != cudaSuccess) return nullptr;
    p[used++] = d;
    return static_cast<T*>(d);
  }
  template <typename T> void back(T* host, const T* dev, size_t count, cudaStream_t st) {
    if (host && dev && err == cudaSuccess) err = cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st);
  }
  ~DevScratch() { for (int i = 0; i < used; ++i) cudaFree(p[i]); }
};
}  // namespace

#define WBC_SCRATCH_CHECK(h, s)                                                                  \
  do {                                                                                           \
    if ((s).err != cudaSuccess) { (h)->err = std::string("host-buffer entry: ") + cudaGetErrorString((s).err); return WBC_ERR_CUDA; } \
  } while (0)

extern "C" int wbc_coriolis_host(wbc_handle* h, int64_t n, const double* q, const double* v, double* Cm, double* Jd) {
  if (!h) return WBC_ERR_ARG;
  if (n <= 0) return n == 0 ? WBC_OK : fail_arg(h, "wbc_coriolis_host: n < 0");
  if (!q || !v) return fail_arg(h, "wbc_coriolis_host: null input");
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  const double* dq = s.in(q, N * WBC_NQ, st); const double* dv = s.in(v, N * WBC_NV, st);
  double* dC = s.out(Cm, N * 324); double* dJ = s.out(Jd, N * 216);
  WBC_SCRATCH_CHECK(h, s);
  int rc = wbc_coriolis(h, n, dq, dv, dC, dJ, st);
  if (rc) return rc;
  s.back(Cm, dC, N * 324, st); s.back(Jd, dJ, N * 216, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

extern "C" int wbc_step_pd(wbc_handle* h, int64_t n, const double* q, const double* v, double* tau, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!q || !v || !tau))) return fail_arg(h, "wbc_step_pd: null buffer");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  const long long total = (long long)n * WBC_NU;
  wbc_pd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d_const, q, v, tau, n);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

// Instances per reduce / solve launch pair of the split path: bounds the hand-over scratch (4224 B per instance) at 1.1 GB
// while each kernel still runs for milliseconds, so the drain at the kernel boundaries stays below ~1 % of the step.
constexpr int64_t SPLIT_CHUNK = 262144;

// Grows the hand-over scratch of `slot` to `n` instances (never inside a stream capture: callers that capture call this first).
static int ensure_split_scratch(wbc_handle* h, int slot, int64_t n, bool with_vd) {
  const int64_t m = n < SPLIT_CHUNK ? n : SPLIT_CHUNK;
  if (m > h->rec_cap[slot]) {
    cudaFree(h->d_rec[slot]); h->d_rec[slot] = nullptr; h->rec_cap[slot] = 0;
    WBC_CUDA(h, cudaMalloc(&h->d_rec[slot], m * wbc::REC_DOUBLES * sizeof(double)));
    h->rec_cap[slot] = m;
  }
  if (with_vd && m > h->vdmap_cap[slot]) {
    cudaFree(h->d_vdmap[slot]); h->d_vdmap[slot] = nullptr; h->vdmap_cap[slot] = 0;
    WBC_CUDA(h, cudaMalloc(&h->d_vdmap[slot], m * wbc::VDMAP_DOUBLES * sizeof(double)));
    h->vdmap_cap[slot] = m;
  }
  return WBC_OK;
}

static wbc::StepArgs offset_args(const wbc_io* io, int64_t o, int64_t m, int kind) {
  return wbc::StepArgs{io->q + o * WBC_NQ, io->v + o * WBC_NV, io->traj + o * WBC_NTRAJ, io->contact + o * 4, io->tau + o * WBC_NU,
                       io->metrics + o * WBC_NMETRIC, io->status + o, io->vd ? io->vd + o * WBC_NV : nullptr,
                       io->f ? io->f + o * 12 : nullptr, io->qp_info ? io->qp_info + o * 4 : nullptr,
                       io->lam ? io->lam + o * WBC_NLAM : nullptr, (long long)m, kind};
}

static int pdl_mode() {   // WBC_PDL=0: ordinary launch of the solve kernel (A/B comparisons)
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("WBC_PDL"); mode = e ? atoi(e) : 1; }
  return mode;
}
static bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static int bulk_in_mode() {   // WBC_BULK_IN=0: per-lane input loads everywhere (A/B comparisons)
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("WBC_BULK_IN"); mode = e ? atoi(e) : 1; }
  return mode;
}

// The hand-over scratch of a slot belongs to one step at a time. A step on another stream than the slot's previous user is
// ordered behind everything submitted to that stream so far (event dependency; never inside a stream capture, where the
// caller owns the ordering).
static int order_slot(wbc_handle* h, int slot, cudaStream_t st) {
  if (h->slot_used[slot] && h->last_stream[slot] != st) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      if (!h->order_ev) WBC_CUDA(h, cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
      if (cudaEventRecord(h->order_ev, h->last_stream[slot]) == cudaSuccess) WBC_CUDA(h, cudaStreamWaitEvent(st, h->order_ev, 0));
      else cudaGetLastError();      // the previous stream no longer exists: its work has completed
    }
  }
  h->slot_used[slot] = true;
  h->last_stream[slot] = st;
  return WBC_OK;
}

template <int KIND>
static int launch_split(wbc_handle* h, int kind, int64_t n, const wbc_io* io, cudaStream_t st, int slot) {
  int rc = ensure_split_scratch(h, slot, n, io->vd != nullptr);
  if (rc) return rc;
  rc = order_slot(h, slot, st);
  if (rc) return rc;
  double* vdmap = io->vd ? h->d_vdmap[slot] : nullptr;
  for (int64_t o = 0; o < n; o += SPLIT_CHUNK) {
    const int64_t m = (n - o) < SPLIT_CHUNK ? (n - o) : SPLIT_CHUNK;
    const wbc::StepArgs a = offset_args(io, o, m, kind);
    // bulk input staging needs 16-byte aligned rows at the CTA boundaries (chunk offsets are multiples of 4)
    const bool al_vt = aligned16p(a.v) && aligned16p(a.traj), al_qc = aligned16p(a.q) && aligned16p(a.contact);
    if (h->prof_on) cudaEventRecord(h->prof_ev[0], st);
    if (h->host_mapped || KIND == WBC_CTRL_PC) {   // zero-copy call of wbc_step_host: 4-warp CTAs, every array staged per CTA
                                                   // (PC / MPTC too: 20.0 vs 19.7 M steps/s with single-warp CTAs)
      constexpr int W = (KIND == WBC_CTRL_PC) ? WARPS_PC : WARPS_HOST;
      const unsigned grid = (unsigned)((m + W - 1) / W);
      const int bulk_ok = bulk_in_mode() && al_vt && al_qc;
      if constexpr (KIND == WBC_CTRL_PC) wbc_reduce_pc_kernel<W><<<grid, W * 32, sizeof(SmemLayoutPCT<W>), st>>>(h->d_const, a, h->d_rec[slot], vdmap, bulk_ok);
      else wbc_reduce_kernel<KIND, W><<<grid, W * 32, sizeof(SmemLayoutT<W>), st>>>(h->d_const, a, h->d_rec[slot], vdmap, bulk_ok);
    } else {
      constexpr int W = WARPS;
      const unsigned grid = (unsigned)((m + W - 1) / W);
      const int bulk_ok = bulk_in_mode() && al_vt;
      if constexpr (KIND != WBC_CTRL_PC) wbc_reduce_kernel<KIND, W><<<grid, W * 32, sizeof(SmemLayoutT<W>), st>>>(h->d_const, a, h->d_rec[slot], vdmap, bulk_ok);
    }
    if (h->prof_on) cudaEventRecord(h->prof_ev[1], st);
    const unsigned sgrid = (unsigned)((m + SOLVE_WARPS - 1) / SOLVE_WARPS);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    // programmatic dependent launch of the solve kernel (its CTAs become resident while the last reduce CTAs run): +1.5 % at 4096
    // instances, neutral above; never inside a stream capture ...
    static const long long pdl_min = getenv("WBC_PDL_MIN") ? atoll(getenv("WBC_PDL_MIN")) : 4096;
    // ... and chains that run beside other chains of the same step (chunked issue): there the early-resident solve CTAs take the
    // slots the other chunk's reduce CTAs need (-22 % at 2 x 2048). A lone chain gains at every size (+3-4 % at 64 - 1024).
    const bool pdl = pdl_mode() && (m >= pdl_min || !h->side_by_side);
    if (pdl) cudaStreamIsCapturing(st, &cap);
    const bool with_vd = io->vd != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sgrid); cfg.blockDim = dim3(SOLVE_WARPS * 32); cfg.stream = st;
    cfg.dynamicSmemBytes = with_vd ? sizeof(SmemLayoutSolveT<true>) : sizeof(SmemLayoutSolveT<false>);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl && cap == cudaStreamCaptureStatusNone) ? 1 : 0;
    const double* crec = h->d_rec[slot];
    const double* cvd = vdmap;
    if (with_vd) WBC_CUDA(h, cudaLaunchKernelEx(&cfg, wbc_solve_kernel<KIND, true>, (const DevConst*)h->d_const, a, crec, cvd));
    else WBC_CUDA(h, cudaLaunchKernelEx(&cfg, wbc_solve_kernel<KIND, false>, (const DevConst*)h->d_const, a, crec, cvd));
    if (h->prof_on) cudaEventRecord(h->prof_ev[2], st);
    h->launches += 2;
  }
  return WBC_OK;
}

// One control step = reduce kernel (per controller kind) + solve kernel, stream ordered; `slot` selects the hand-over scratch
// (the host pipelines run two streams side by side).
static int step_launch(wbc_handle* h, int kind, int64_t n, const wbc_io* io, cudaStream_t st, int slot) {
  int rc = WBC_OK;
  switch (kind) {
    case WBC_CTRL_ID: rc = launch_split<WBC_CTRL_ID>(h, kind, n, io, st, slot); break;
    case WBC_CTRL_CLF: rc = launch_split<WBC_CTRL_CLF>(h, kind, n, io, st, slot); break;
    case WBC_CTRL_PC:
    case WBC_CTRL_MPTC: rc = launch_split<WBC_CTRL_PC>(h, kind, n, io, st, slot); break;   // MPTC: a.kind drops the passivity row
    default: return fail_arg(h, "wbc_step: unknown controller kind");
  }
  if (rc) return rc;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

static int two_stream_mode() {   // WBC_TWO_STREAMS=0: one reduce -> solve chain per step (A/B comparisons); k >= 2: k chunks;
  static int mode = -1;          // unset (1): chunks of about 2048 instances, 2 to 4 of them
  if (mode < 0) { const char* e = getenv("WBC_TWO_STREAMS"); mode = e ? atoi(e) : 1; }
  return mode;
}
static int64_t two_stream_max() {
  static int64_t v = -1;
  if (v < 0) { const char* e = getenv("WBC_TWO_STREAMS_MAX"); v = e ? atoll(e) : 65536; }
  return v;
}

extern "C" int wbc_step(wbc_handle* h, int kind, int64_t n, const wbc_io* io, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (!io || n < 0) return fail_arg(h, "wbc_step: bad arguments");
  if (n == 0) return WBC_OK;
  if (kind == WBC_CTRL_PD) return wbc_step_pd(h, n, io->q, io->v, io->tau, stream);
  if (!io->q || !io->v || !io->traj || !io->contact || !io->tau || !io->metrics || !io->status)
    return fail_arg(h, "wbc_step: q, v, traj, contact, tau, metrics and status are required");
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (two_stream_mode() && !h->prof_on && n >= 4096 && n <= two_stream_max()) {
    // Small batches go through as chunks of about 2048 instances, independent reduce -> solve chains on as many streams (the
    // caller's and internal ones, joined by events): the solve kernel of one chunk overlaps the reduce kernel and the ramp-down
    // of the others. Measured against one chain: +5.5 % at 4096 instances (2 chunks; 3: +4.5 %, 4: +2.5 %), +7.7 % at 8192
    // (4 chunks; 2: +4.1 %), +2.6 % at 16384, +1 % at 65536, -2 % at 2048 (profiles/README.md). Never inside a stream capture.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      if (!h->order_ev) WBC_CUDA(h, cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
      if (!h->join_ev) WBC_CUDA(h, cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
      int chunks = two_stream_mode() < 2 ? (n >= 7168 ? 4 : n >= 5120 ? 3 : 2) : two_stream_mode();
      if (chunks > WBC_NSLOT) chunks = WBC_NSLOT;
      const int64_t per = (((n + chunks - 1) / chunks) + 3) & ~(int64_t)3;
      WBC_CUDA(h, cudaEventRecord(h->order_ev, st));
      for (int c = 1; c < chunks; ++c) {
        if (c >= 2 && !h->xstream[c]) WBC_CUDA(h, cudaStreamCreateWithFlags(&h->xstream[c], cudaStreamNonBlocking));
        WBC_CUDA(h, cudaStreamWaitEvent(c == 1 ? h->stream2 : h->xstream[c], h->order_ev, 0));
      }
      for (int c = 0; c < chunks; ++c) {
        const int64_t o = c * per, m = (o + per <= n) ? per : n - o;
        if (m <= 0) break;
        const wbc_io cio{io->q + o * WBC_NQ, io->v + o * WBC_NV, io->traj + o * WBC_NTRAJ, io->contact + o * 4, io->tau + o * WBC_NU,
                         io->metrics + o * WBC_NMETRIC, io->status + o, io->vd ? io->vd + o * WBC_NV : nullptr,
                         io->f ? io->f + o * 12 : nullptr, io->qp_info ? io->qp_info + o * 4 : nullptr, io->lam ? io->lam + o * WBC_NLAM : nullptr};
        cudaStream_t cs = c == 0 ? st : c == 1 ? h->stream2 : h->xstream[c];
        h->side_by_side = true;
        const int rc = step_launch(h, kind, m, &cio, cs, c);
        h->side_by_side = false;
        if (rc) return rc;
        if (c) {
          WBC_CUDA(h, cudaEventRecord(h->join_ev, cs));
          WBC_CUDA(h, cudaStreamWaitEvent(st, h->join_ev, 0));
        }
      }
      return WBC_OK;
    }
  }
  return step_launch(h, kind, n, io, st, 0);
}

#define WBC_STEP_WRAPPER(name, kind)                                                                               \
  extern "C" int name(wbc_handle* h, int64_t n, const double* q, const double* v, const double* traj,              \
                      const uint8_t* contact, double* tau, double* metrics, int32_t* status, void* stream) {       \
    wbc_io io{q, v, traj, contact, tau, metrics, status, nullptr, nullptr, nullptr};                               \
    return wbc_step(h, kind, n, &io, stream);                                                                      \
  }
WBC_STEP_WRAPPER(wbc_step_id, WBC_CTRL_ID)
WBC_STEP_WRAPPER(wbc_step_clf, WBC_CTRL_CLF)
WBC_STEP_WRAPPER(wbc_step_pc, WBC_CTRL_PC)
WBC_STEP_WRAPPER(wbc_step_mptc, WBC_CTRL_MPTC)

static int ensure_staging(wbc_handle* h, int64_t n) {
  if (n <= h->cap) return WBC_OK;
  free_staging(h);
  WBC_CUDA(h, cudaMalloc(&h->d_q, n * WBC_NQ * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_v, n * WBC_NV * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_traj, n * WBC_NTRAJ * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_contact, n * 4));
  WBC_CUDA(h, cudaMalloc(&h->d_tau, n * WBC_NU * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_metrics, n * WBC_NMETRIC * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_status, n * sizeof(int32_t)));
  WBC_CUDA(h, cudaMalloc(&h->d_vd, n * WBC_NV * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_f, n * 12 * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_info, n * 4 * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->d_lam, n * WBC_NLAM * sizeof(double)));
  h->cap = n;
  return WBC_OK;
}

// Enqueues one host-buffer step on the handle's two internal streams without waiting for it; `wait` must follow before the
// outputs are read or the next step is enqueued on this handle (wbc_step_host = enqueue + wait; wbc_multi_step_host enqueues on
// every device first and waits afterwards, so the devices run side by side from one host thread).
static int step_host_wait(wbc_handle* h) {
  WBC_CUDA(h, cudaSetDevice(h->device));
  WBC_CUDA(h, cudaStreamSynchronize(h->stream));
  WBC_CUDA(h, cudaStreamSynchronize(h->stream2));
  return WBC_OK;
}

static int step_host_enqueue(wbc_handle* h, int kind, int64_t n, const wbc_io* io) {
  if (!h) return WBC_ERR_ARG;
  if (!io || n < 0) return fail_arg(h, "wbc_step_host: bad arguments");
  if (n == 0) return WBC_OK;
  const bool pd = kind == WBC_CTRL_PD;
  if (!io->q || !io->v || !io->tau || (!pd && (!io->traj || !io->contact || !io->metrics || !io->status)))
    return fail_arg(h, "wbc_step_host: q, v, traj, contact, tau, metrics and status are required");
  WBC_CUDA(h, cudaSetDevice(h->device));
  int rc = WBC_OK;
  // Zero-copy path: when every buffer is page-locked host memory (wbc_host_alloc / cudaHostRegister) the kernel reads
  // its 732 B of inputs and writes its 132 B of outputs per instance straight over the host link - one launch, no
  // staging copies, the transfers of one warp overlap the arithmetic of the others.
  {
    int mode = -1;   // auto; WBC_HOST_ZEROCOPY=0 / 1 forces the staged / zero-copy path (experiments)
    if (const char* env = getenv("WBC_HOST_ZEROCOPY")) mode = atoi(env);
    const void* ptrs[11] = {io->q, io->v, io->traj, io->contact, io->tau, io->metrics, io->status, io->vd, io->f, io->qp_info, io->lam};
    void* dev[11] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // Measured on B200 (profiles/README.md): reading the inputs over the host link from the SMs (zero-copy, ~32 GB/s) wins up
    // to ~48 k instances per call (no staging copies, no extra API calls); above that the inputs go through the copy engine
    // (~55 GB/s) in a chunked two-stream pipeline. Page-locked outputs are written straight to host memory in both modes.
    if (mode < 0) mode = 1;
    bool pinned = mode != 0;
    for (int i = 0; i < 11 && pinned; ++i) {
      if (!ptrs[i]) continue;
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, ptrs[i]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) { pinned = false; cudaGetLastError(); }
      else dev[i] = at.devicePointer;
    }
    if (pinned) {
      const wbc_io hio{(const double*)dev[0], (const double*)dev[1], (const double*)dev[2], (const uint8_t*)dev[3], (double*)dev[4],
                       (double*)dev[5], (int32_t*)dev[6], (double*)dev[7], (double*)dev[8], (double*)dev[9], (double*)dev[10]};
      // The batch goes through in chunks alternating between the two internal streams (each with its own hand-over scratch); the
      // outputs are written straight to host memory by the solve kernel in every mode.
      //  * below 24576 instances the reduce kernel reads its inputs over the host link (zero-copy), two halves from 4096 on:
      //    the solve kernel of one half overlaps the link-bound reduce kernel of the other (e2e +4.6 % at 4096, +3.8 % at
      //    16384; three or more chunks lose);
      //  * above, the inputs go through the copy engine into device staging (WBC_ZC_STAGE bit 0: traj, bit 1: q, v, contact) and
      //    the single-warp reduce CTAs read resident inputs: e2e +5 % at 32768, +16 % at 65536, +10 % at 2^17 - 2^20.
      static const long long stage_min = getenv("WBC_ZC_MAX") ? atoll(getenv("WBC_ZC_MAX")) : 24576;
      static const int stage_env = getenv("WBC_ZC_STAGE") ? atoi(getenv("WBC_ZC_STAGE")) : -1;
      // PC / MPTC are compute bound at a third of the rate: the zero-copy reads hide under the reduce kernel up to 131072
      // instances (e2e 22.3 M steps/s against 20.0 M staged at 65536)
      const bool pc_kind = kind == WBC_CTRL_PC || kind == WBC_CTRL_MPTC;
      const int stage_in = stage_env >= 0 ? stage_env : (n >= (pc_kind ? 131072 : stage_min) ? 3 : 0);
      int zc = n >= 4096 ? 2 : 1;
      if (stage_in) {   // four chunks up to 65536 instances, then chunks of n / 8 clamped to [16384, 32768] instances
        int64_t per_c = n / 8; per_c = per_c < 16384 ? 16384 : (per_c > 32768 ? 32768 : per_c);
        if (n < 98304) per_c = ((n + 3) / 4 + 3) & ~(int64_t)3;
        zc = (int)((n + per_c - 1) / per_c);
      }
      if (const char* env = getenv("WBC_ZC_CHUNKS")) { const int v = atoi(env); if (v >= 1 && v <= 64) zc = v; }
      if ((int64_t)zc > n) zc = (int)n;
      const int64_t per = ((n + zc - 1) / zc + 3) & ~(int64_t)3;
      cudaStream_t lanes[2] = {h->stream, h->stream2};
      int used = 0;
      if (stage_in && !pd) { rc = ensure_staging(h, n); if (rc) return rc; }
      for (int c = 0; c < zc; ++c) {
        const int64_t o = c * per, m = (o + per <= n) ? per : n - o;
        if (m <= 0) break;
        const bool staged = stage_in && !pd;
        wbc_io dio = hio;
        if (staged) {
          cudaStream_t cs = lanes[c & 1];
          if (stage_in & 2) {
            WBC_CUDA(h, cudaMemcpyAsync(h->d_q + o * WBC_NQ, io->q + o * WBC_NQ, m * WBC_NQ * sizeof(double), cudaMemcpyHostToDevice, cs));
            WBC_CUDA(h, cudaMemcpyAsync(h->d_v + o * WBC_NV, io->v + o * WBC_NV, m * WBC_NV * sizeof(double), cudaMemcpyHostToDevice, cs));
            WBC_CUDA(h, cudaMemcpyAsync(h->d_contact + o * 4, io->contact + o * 4, m * 4, cudaMemcpyHostToDevice, cs));
            dio.q = h->d_q; dio.v = h->d_v; dio.contact = h->d_contact;
          }
          if (stage_in & 1) {
            WBC_CUDA(h, cudaMemcpyAsync(h->d_traj + o * WBC_NTRAJ, io->traj + o * WBC_NTRAJ, m * WBC_NTRAJ * sizeof(double), cudaMemcpyHostToDevice, cs));
            dio.traj = h->d_traj;
          }
        }
        const wbc_io cio{dio.q + o * WBC_NQ, dio.v + o * WBC_NV, dio.traj ? dio.traj + o * WBC_NTRAJ : nullptr,
                         dio.contact ? dio.contact + o * 4 : nullptr, dio.tau + o * WBC_NU,
                         dio.metrics ? dio.metrics + o * WBC_NMETRIC : nullptr, dio.status ? dio.status + o : nullptr,
                         dio.vd ? dio.vd + o * WBC_NV : nullptr, dio.f ? dio.f + o * 12 : nullptr, dio.qp_info ? dio.qp_info + o * 4 : nullptr,
                         dio.lam ? dio.lam + o * WBC_NLAM : nullptr};
        h->host_mapped = !staged || (stage_in & 3) != 3;   // the reduce kernel reads host memory (4-warp CTAs, everything staged per CTA)
        h->side_by_side = zc > 1;
        rc = pd ? wbc_step_pd(h, m, cio.q, cio.v, cio.tau, lanes[c & 1]) : step_launch(h, kind, m, &cio, lanes[c & 1], c & 1);
        h->host_mapped = false; h->side_by_side = false;
        if (rc) return rc;
        used |= 1 << (c & 1);
      }
      (void)used;
      return WBC_OK;
    }
  }
  rc = ensure_staging(h, n);
  if (rc) return rc;
  // Chunked two-stream pipeline: the upload of chunk c + 1 overlaps the kernel of chunk c, the download of chunk c the
  // kernel of chunk c + 1 (separate copy engines); small batches go through in one piece.
  // chunk = n / 8 clamped to [8192, 65536] instances: short pipeline fill, copies long enough for the copy engines
  int n_chunks = n >= 2048 ? 2 : 1;
  if (n >= 32768) {
    int64_t per_c = n / 8; per_c = per_c < 8192 ? 8192 : (per_c > 65536 ? 65536 : per_c);
    const int64_t c = (n + per_c - 1) / per_c;
    n_chunks = (int)(c > 64 ? 64 : c);
  }
  if (const char* env = getenv("WBC_HOST_CHUNKS")) { const int v = atoi(env); if (v >= 1 && v <= 64) n_chunks = v; }
  if ((int64_t)n_chunks > n) n_chunks = (int)n;
  const int64_t per = (n + n_chunks - 1) / n_chunks;
  cudaStream_t lanes[2] = {h->stream, h->stream2};
  for (int c = 0; c < n_chunks; ++c) {
    const int64_t o = c * per, m = (o + per <= n) ? per : n - o;
    if (m <= 0) break;
    cudaStream_t st = lanes[c & 1];
    WBC_CUDA(h, cudaMemcpyAsync(h->d_q + o * WBC_NQ, io->q + o * WBC_NQ, m * WBC_NQ * sizeof(double), cudaMemcpyHostToDevice, st));
    WBC_CUDA(h, cudaMemcpyAsync(h->d_v + o * WBC_NV, io->v + o * WBC_NV, m * WBC_NV * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!pd) {
      WBC_CUDA(h, cudaMemcpyAsync(h->d_traj + o * WBC_NTRAJ, io->traj + o * WBC_NTRAJ, m * WBC_NTRAJ * sizeof(double), cudaMemcpyHostToDevice, st));
      WBC_CUDA(h, cudaMemcpyAsync(h->d_contact + o * 4, io->contact + o * 4, m * 4, cudaMemcpyHostToDevice, st));
    }
    wbc_io dio{h->d_q + o * WBC_NQ, h->d_v + o * WBC_NV, h->d_traj + o * WBC_NTRAJ, h->d_contact + o * 4, h->d_tau + o * WBC_NU,
               h->d_metrics + o * WBC_NMETRIC, h->d_status + o,
               io->vd ? h->d_vd + o * WBC_NV : nullptr, io->f ? h->d_f + o * 12 : nullptr, io->qp_info ? h->d_info + o * 4 : nullptr,
               io->lam ? h->d_lam + o * WBC_NLAM : nullptr};
    h->side_by_side = n_chunks > 1;
    rc = pd ? wbc_step_pd(h, m, dio.q, dio.v, dio.tau, st) : step_launch(h, kind, m, &dio, st, c & 1);
    h->side_by_side = false;
    if (rc) return rc;
    WBC_CUDA(h, cudaMemcpyAsync(io->tau + o * WBC_NU, h->d_tau + o * WBC_NU, m * WBC_NU * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!pd) {
      WBC_CUDA(h, cudaMemcpyAsync(io->metrics + o * WBC_NMETRIC, h->d_metrics + o * WBC_NMETRIC, m * WBC_NMETRIC * sizeof(double), cudaMemcpyDeviceToHost, st));
      WBC_CUDA(h, cudaMemcpyAsync(io->status + o, h->d_status + o, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      if (io->vd) WBC_CUDA(h, cudaMemcpyAsync(io->vd + o * WBC_NV, h->d_vd + o * WBC_NV, m * WBC_NV * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (io->f) WBC_CUDA(h, cudaMemcpyAsync(io->f + o * 12, h->d_f + o * 12, m * 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (io->qp_info) WBC_CUDA(h, cudaMemcpyAsync(io->qp_info + o * 4, h->d_info + o * 4, m * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (io->lam) WBC_CUDA(h, cudaMemcpyAsync(io->lam + o * WBC_NLAM, h->d_lam + o * WBC_NLAM, m * WBC_NLAM * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
  }
  return WBC_OK;
}

extern "C" int wbc_step_host(wbc_handle* h, int kind, int64_t n, const wbc_io* io) {
  const int rc = step_host_enqueue(h, kind, n, io);
  if (rc || !h || n <= 0) return rc;
  return step_host_wait(h);
}

// ------------------------------------------------------------------------------ all GPUs of the box from one call
// north star: "instances shard trivially across the 8 GPUs of one box, with no NCCL beyond an optional host gather". One
// handle per device; a call splits [0, n) into contiguous equal shards, enqueues each shard on its device and
// waits for all of them: the outputs land in the caller's (page-locked or pageable) host arrays, which IS the host gather.
struct wbc_multi {
  std::vector<wbc_handle*> h;
  std::string err;
};

extern "C" int wbc_multi_destroy(wbc_multi* m) {
  if (!m) return WBC_OK;
  for (wbc_handle* h : m->h) wbc_destroy(h);
  delete m;
  return WBC_OK;
}

extern "C" int wbc_multi_create(const wbc_model* model, const wbc_params* params, int n_devices, const int* devices, wbc_multi** out) {
  if (!model || !out || n_devices <= 0) return WBC_ERR_ARG;
  *out = nullptr;
  wbc_multi* m = new (std::nothrow) wbc_multi();
  if (!m) return WBC_ERR_NOMEM;
  for (int i = 0; i < n_devices; ++i) {
    wbc_handle* h = nullptr;
    const int rc = wbc_create(model, params, devices ? devices[i] : i, &h);
    if (h) m->h.push_back(h);
    if (rc) { m->err = h ? h->err : "wbc_create failed"; *out = m; return rc; }
  }
  *out = m;
  return WBC_OK;
}

extern "C" const char* wbc_multi_last_error(const wbc_multi* m) { return m ? m->err.c_str() : "null handle"; }
extern "C" int wbc_multi_device_count(const wbc_multi* m) { return m ? (int)m->h.size() : 0; }
extern "C" int64_t wbc_multi_launch_count(const wbc_multi* m) {
  int64_t s = 0;
  if (m) for (const wbc_handle* h : m->h) s += h->launches;
  return s;
}

// Shard r of n over w devices: contiguous, sizes differ by at most one (the same rule as quadruped_drake_b200/sharding.py).
static void shard_of(int64_t n, int r, int w, int64_t* lo, int64_t* hi) {
  const int64_t base = n / w, rem = n % w;
  *lo = r * base + (r < rem ? r : rem);
  *hi = *lo + base + (r < rem ? 1 : 0);
}

extern "C" int wbc_multi_step_host(wbc_multi* m, int kind, int64_t n, const wbc_io* io) {
  if (!m) return WBC_ERR_ARG;
  if (!io || n < 0) { m->err = "wbc_multi_step_host: bad arguments"; return WBC_ERR_ARG; }
  const int w = (int)m->h.size();
  int rc = WBC_OK;
  int enq = 0;
  for (int r = 0; r < w && rc == WBC_OK; ++r) {
    int64_t lo, hi;
    shard_of(n, r, w, &lo, &hi);
    if (hi <= lo) { ++enq; continue; }
    const wbc_io sio{io->q + lo * WBC_NQ, io->v + lo * WBC_NV, io->traj ? io->traj + lo * WBC_NTRAJ : nullptr,
                     io->contact ? io->contact + lo * 4 : nullptr, io->tau + lo * WBC_NU,
                     io->metrics ? io->metrics + lo * WBC_NMETRIC : nullptr, io->status ? io->status + lo : nullptr,
                     io->vd ? io->vd + lo * WBC_NV : nullptr, io->f ? io->f + lo * 12 : nullptr,
                     io->qp_info ? io->qp_info + lo * 4 : nullptr, io->lam ? io->lam + lo * WBC_NLAM : nullptr};
    rc = step_host_enqueue(m->h[r], kind, hi - lo, &sio);
    if (rc) m->err = m->h[r]->err;
    ++enq;
  }
  for (int r = 0; r < enq && r < w; ++r) {                 // always drain what was enqueued, even after an error
    const int rw = step_host_wait(m->h[r]);
    if (rw && rc == WBC_OK) { rc = rw; m->err = m->h[r]->err; }
  }
  return rc;
}

extern "C" int wbc_dynamics_host(wbc_handle* h, int64_t n, const double* q, const double* v, double* M, double* Cv,
                                 double* taug, double* Jfeet, double* Jdv, double* pfeet) {
  if (!h) return WBC_ERR_ARG;
  if (n <= 0) return n == 0 ? WBC_OK : fail_arg(h, "wbc_dynamics_host: n < 0");
  if (!q || !v) return fail_arg(h, "wbc_dynamics_host: null input");
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  const double* dq = s.in(q, N * WBC_NQ, st); const double* dv = s.in(v, N * WBC_NV, st);
  double* dM = s.out(M, N * 324); double* dCv = s.out(Cv, N * 18); double* dtg = s.out(taug, N * 18);
  double* dJ = s.out(Jfeet, N * 216); double* dJdv = s.out(Jdv, N * 12); double* dp = s.out(pfeet, N * 12);
  WBC_SCRATCH_CHECK(h, s);
  int rc = wbc_dynamics(h, n, dq, dv, dM, dCv, dtg, dJ, dJdv, dp, st);
  if (rc) return rc;
  s.back(M, dM, N * 324, st); s.back(Cv, dCv, N * 18, st); s.back(taug, dtg, N * 18, st);
  s.back(Jfeet, dJ, N * 216, st); s.back(Jdv, dJdv, N * 12, st); s.back(pfeet, dp, N * 12, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

extern "C" int wbc_time_step(wbc_handle* h, int kind, int64_t n, const wbc_io* io, int reps, void* stream,
                             double* ms_per_launch) {
  if (!h) return WBC_ERR_ARG;
  if (!ms_per_launch || reps <= 0) return fail_arg(h, "wbc_time_step: bad arguments");
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  WBC_CUDA(h, cudaEventCreate(&e0));
  WBC_CUDA(h, cudaEventCreate(&e1));
  WBC_CUDA(h, cudaEventRecord(e0, st));
  for (int r = 0; r < reps; ++r) {
    int rc = wbc_step(h, kind, n, io, stream);
    if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
  }
  WBC_CUDA(h, cudaEventRecord(e1, st));
  WBC_CUDA(h, cudaEventSynchronize(e1));
  float ms = 0.f;
  WBC_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_launch = (double)ms / reps;
  return WBC_OK;
}

// Per-kernel share of one control step: CUDA events before the reduce kernel, between the two kernels and after the solve
// kernel, averaged over `reps` steps (n <= 262144 so that the step is one launch pair). Measurement aid for bench.py.
extern "C" int wbc_profile_step(wbc_handle* h, int kind, int64_t n, const wbc_io* io, int reps, void* stream, double* ms_reduce,
                                double* ms_solve) {
  if (!h) return WBC_ERR_ARG;
  if (!ms_reduce || !ms_solve || reps <= 0 || n <= 0 || n > SPLIT_CHUNK || kind == WBC_CTRL_PD) return fail_arg(h, "wbc_profile_step: bad arguments");
  WBC_CUDA(h, cudaSetDevice(h->device));
  for (int i = 0; i < 3; ++i) if (!h->prof_ev[i]) WBC_CUDA(h, cudaEventCreate(&h->prof_ev[i]));
  double r = 0.0, s = 0.0;
  for (int k = 0; k < reps; ++k) {
    h->prof_on = true;
    const int rc = wbc_step(h, kind, n, io, stream);
    h->prof_on = false;
    if (rc) return rc;
    WBC_CUDA(h, cudaEventSynchronize(h->prof_ev[2]));
    float a = 0.f, b = 0.f;
    WBC_CUDA(h, cudaEventElapsedTime(&a, h->prof_ev[0], h->prof_ev[1]));
    WBC_CUDA(h, cudaEventElapsedTime(&b, h->prof_ev[1], h->prof_ev[2]));
    r += a; s += b;
  }
  *ms_reduce = r / reps; *ms_solve = s / reps;
  return WBC_OK;
}

// ------------------------------------------------------------------------------ LCM wire codecs (wbc_wire.cuh)
static unsigned wire_grid(const wbc_handle* h, int64_t n, int tile) {
  const int64_t tiles = (n + tile - 1) / tile, cap = (int64_t)h->sm_count * 8;     // grid-stride over tiles
  return (unsigned)(tiles < cap ? tiles : cap);
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int wbc_lcm_decode_trunk_state(wbc_handle* h, int64_t n, const uint8_t* msgs, double* timestamp, uint8_t* finished,
                                          double* traj, uint8_t* contact, double* f_plan, int32_t* status, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !traj || !contact))) return fail_arg(h, "wbc_lcm_decode_trunk_state: msgs, traj and contact are required");
  if (!aligned16(msgs)) return fail_arg(h, "wbc_lcm_decode_trunk_state: msgs must be 16-byte aligned");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  wbcwire::decode_trunk_kernel<<<wire_grid(h, n, wbcwire::TILE), wbcwire::THREADS, 0, (cudaStream_t)stream>>>(
      msgs, n, timestamp, finished, traj, contact, f_plan, status);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_lcm_encode_trunk_state(wbc_handle* h, int64_t n, const double* timestamp, const uint8_t* finished,
                                          const double* traj, const uint8_t* contact, const double* f_plan, uint8_t* msgs,
                                          void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !traj || !contact))) return fail_arg(h, "wbc_lcm_encode_trunk_state: msgs, traj and contact are required");
  if (!aligned16(msgs)) return fail_arg(h, "wbc_lcm_encode_trunk_state: msgs must be 16-byte aligned");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  wbcwire::encode_trunk_kernel<<<wire_grid(h, n, wbcwire::TILE), wbcwire::THREADS, 0, (cudaStream_t)stream>>>(
      msgs, n, timestamp, finished, traj, contact, f_plan);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_lcm_decode_robot_state(wbc_handle* h, int64_t n, const uint8_t* msgs, double* q, double* v, double* tau,
                                          int32_t* status, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !q || !v))) return fail_arg(h, "wbc_lcm_decode_robot_state: msgs, q and v are required");
  if (!aligned16(msgs)) return fail_arg(h, "wbc_lcm_decode_robot_state: msgs must be 16-byte aligned");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  wbcwire::decode_robot_kernel<<<wire_grid(h, n, wbcwire::RTILE), wbcwire::THREADS, 0, (cudaStream_t)stream>>>(msgs, n, q, v, tau, status);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_lcm_encode_robot_state(wbc_handle* h, int64_t n, const double* q, const double* v, const double* tau,
                                          int tau_in_actuator_order, uint8_t* msgs, int32_t* status, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !tau))) return fail_arg(h, "wbc_lcm_encode_robot_state: msgs and tau are required");
  if (!aligned16(msgs)) return fail_arg(h, "wbc_lcm_encode_robot_state: msgs must be 16-byte aligned");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  wbcwire::encode_robot_kernel<<<wire_grid(h, n, wbcwire::RTILE), wbcwire::THREADS, 0, (cudaStream_t)stream>>>(
      msgs, n, q, v, tau, tau_in_actuator_order ? h->d_tau_map : nullptr, status);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_lcm_decode_trunk_state_host(wbc_handle* h, int64_t n, const uint8_t* msgs, double* timestamp, uint8_t* finished,
                                               double* traj, uint8_t* contact, double* f_plan, int32_t* status) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !traj || !contact))) return fail_arg(h, "wbc_lcm_decode_trunk_state_host: msgs, traj and contact are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  uint8_t* dm = s.in(msgs, N * WBC_LCM_TRUNK_STATE_BYTES, st);
  double* dts = s.out(timestamp, N); uint8_t* dfin = s.out(finished, N);
  double* dtr = s.out(traj, N * WBC_NTRAJ); uint8_t* dc = s.out(contact, N * 4);
  double* dfp = s.out(f_plan, N * 12); int32_t* dst = s.out(status, N);
  WBC_SCRATCH_CHECK(h, s);
  int rc = wbc_lcm_decode_trunk_state(h, n, dm, dts, dfin, dtr, dc, dfp, dst, st);
  if (rc) return rc;
  s.back(timestamp, dts, N, st); s.back(finished, dfin, N, st); s.back(traj, dtr, N * WBC_NTRAJ, st);
  s.back(contact, dc, N * 4, st); s.back(f_plan, dfp, N * 12, st); s.back(status, dst, N, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

extern "C" int wbc_lcm_encode_trunk_state_host(wbc_handle* h, int64_t n, const double* timestamp, const uint8_t* finished,
                                               const double* traj, const uint8_t* contact, const double* f_plan, uint8_t* msgs) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !traj || !contact))) return fail_arg(h, "wbc_lcm_encode_trunk_state_host: msgs, traj and contact are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  const double* dts = s.in(timestamp, N, st); const uint8_t* dfin = s.in(finished, N, st);
  const double* dtr = s.in(traj, N * WBC_NTRAJ, st); const uint8_t* dc = s.in(contact, N * 4, st);
  const double* dfp = s.in(f_plan, N * 12, st);
  uint8_t* dm = s.out(msgs, N * WBC_LCM_TRUNK_STATE_BYTES);
  WBC_SCRATCH_CHECK(h, s);
  int rc = wbc_lcm_encode_trunk_state(h, n, dts, dfin, dtr, dc, dfp, dm, st);
  if (rc) return rc;
  s.back(msgs, dm, N * WBC_LCM_TRUNK_STATE_BYTES, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

extern "C" int wbc_lcm_decode_robot_state_host(wbc_handle* h, int64_t n, const uint8_t* msgs, double* q, double* v, double* tau,
                                               int32_t* status) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !q || !v))) return fail_arg(h, "wbc_lcm_decode_robot_state_host: msgs, q and v are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  uint8_t* dm = s.in(msgs, N * WBC_LCM_ROBOT_STATE_BYTES, st);
  double* dq = s.out(q, N * WBC_NQ); double* dv = s.out(v, N * WBC_NV); double* dt = s.out(tau, N * WBC_NU);
  int32_t* dst = s.out(status, N);
  WBC_SCRATCH_CHECK(h, s);
  int rc = wbc_lcm_decode_robot_state(h, n, dm, dq, dv, dt, dst, st);
  if (rc) return rc;
  s.back(q, dq, N * WBC_NQ, st); s.back(v, dv, N * WBC_NV, st); s.back(tau, dt, N * WBC_NU, st); s.back(status, dst, N, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

extern "C" int wbc_lcm_encode_robot_state_host(wbc_handle* h, int64_t n, const double* q, const double* v, const double* tau,
                                               int tau_in_actuator_order, uint8_t* msgs, int32_t* status) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!msgs || !tau))) return fail_arg(h, "wbc_lcm_encode_robot_state_host: msgs and tau are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  const double* dq = s.in(q, N * WBC_NQ, st); const double* dv = s.in(v, N * WBC_NV, st); const double* dt = s.in(tau, N * WBC_NU, st);
  uint8_t* dm = s.out(msgs, N * WBC_LCM_ROBOT_STATE_BYTES); int32_t* dst = s.out(status, N);
  WBC_SCRATCH_CHECK(h, s);
  int rc = wbc_lcm_encode_robot_state(h, n, dq, dv, dt, tau_in_actuator_order, dm, dst, st);
  if (rc) return rc;
  s.back(msgs, dm, N * WBC_LCM_ROBOT_STATE_BYTES, st); s.back(status, dst, N, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

// ------------------------------------------------------------------------------ trajectory sampler (wbc_traj.cuh)
struct wbc_plan {
  wbc_handle* h = nullptr;
  wbctraj::PlanTables t{};
  std::vector<void*> bufs;
};

extern "C" int wbc_plan_destroy(wbc_plan* p) {
  if (!p) return WBC_OK;
  if (p->h) cudaSetDevice(p->h->device);
  for (void* b : p->bufs) cudaFree(b);
  delete p;
  return WBC_OK;
}

extern "C" int wbc_plan_create(wbc_handle* h, int32_t n_plans, const wbc_plan_desc* plans, wbc_plan** out) {
  if (!h) return WBC_ERR_ARG;
  if (!out || !plans || n_plans <= 0) return fail_arg(h, "wbc_plan_create: bad arguments");
  *out = nullptr;
  using wbctraj::NSPLINE;
  // ---- flatten on the host: offsets, running sums of the durations (in the reference's summation order), nodes
  std::vector<int> poly_off, phase_off, grid_off, node_of_poly;
  std::vector<unsigned char> cstart;
  std::vector<double> tend, dur, nodes, phase_tend, grid_ts, wait, standing;
  grid_off.push_back(0);
  for (int p = 0; p < n_plans; ++p) {
    const wbc_plan_desc& d = plans[p];
    const wbc_spline_desc* sp[NSPLINE] = {&d.base_linear, &d.base_angular, &d.ee_motion[0], &d.ee_motion[1], &d.ee_motion[2],
                                          &d.ee_motion[3], &d.ee_force[0], &d.ee_force[1], &d.ee_force[2], &d.ee_force[3]};
    for (int s = 0; s < NSPLINE; ++s) {
      if (sp[s]->n_poly <= 0 || !sp[s]->durations || !sp[s]->nodes) return fail_arg(h, "wbc_plan_create: empty spline");
      poly_off.push_back((int)tend.size());
      double acc = 0.0;
      const int node0 = (int)(nodes.size() / 6);
      for (int i = 0; i < sp[s]->n_poly; ++i) {
        if (!(sp[s]->durations[i] > 0.0)) return fail_arg(h, "wbc_plan_create: polynomial duration must be > 0");
        acc += sp[s]->durations[i];
        tend.push_back(acc);
        dur.push_back(sp[s]->durations[i]);
        node_of_poly.push_back(node0 + i);
      }
      nodes.insert(nodes.end(), sp[s]->nodes, sp[s]->nodes + (size_t)(sp[s]->n_poly + 1) * 6);
    }
    poly_off.push_back((int)tend.size());
    for (int k = 0; k < WBC_NLEG; ++k) {
      if (d.n_phase[k] <= 0 || !d.phase_durations[k]) return fail_arg(h, "wbc_plan_create: empty phase list");
      phase_off.push_back((int)phase_tend.size());
      double acc = 0.0;
      for (int i = 0; i < d.n_phase[k]; ++i) { acc += d.phase_durations[k][i]; phase_tend.push_back(acc); }
      cstart.push_back(d.contact_at_start[k] ? 1 : 0);
    }
    phase_off.push_back((int)phase_tend.size());
    if (d.n_grid < 0 || (d.n_grid > 0 && !d.grid_timestamps)) return fail_arg(h, "wbc_plan_create: bad sample grid");
    for (int i = 0; i < d.n_grid; ++i) {
      if (i > 0 && !(d.grid_timestamps[i] > d.grid_timestamps[i - 1])) return fail_arg(h, "wbc_plan_create: grid timestamps must increase");
      grid_ts.push_back(d.grid_timestamps[i]);
    }
    grid_off.push_back((int)grid_ts.size());
    wait.push_back(d.wait_time);
    standing.insert(standing.end(), d.standing, d.standing + WBC_NTRAJ);
  }
  if (grid_ts.empty()) grid_ts.push_back(0.0);
  WBC_CUDA(h, cudaSetDevice(h->device));
  wbc_plan* pl = new (std::nothrow) wbc_plan();
  if (!pl) return WBC_ERR_NOMEM;
  pl->h = h;
  cudaError_t err = cudaSuccess;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* dptr = nullptr;
    if (err != cudaSuccess) return nullptr;
    err = cudaMalloc(&dptr, bytes ? bytes : 1);
    if (err != cudaSuccess) return nullptr;
    pl->bufs.push_back(dptr);
    if (src) err = cudaMemcpy(dptr, src, bytes, cudaMemcpyHostToDevice);
    return dptr;
  };
  const int n_poly_total = (int)tend.size();
  pl->t.n_plans = n_plans;
  pl->t.poly_off = (const int*)up(poly_off.data(), poly_off.size() * sizeof(int));
  pl->t.phase_off = (const int*)up(phase_off.data(), phase_off.size() * sizeof(int));
  pl->t.grid_off = (const int*)up(grid_off.data(), grid_off.size() * sizeof(int));
  pl->t.contact_start = (const unsigned char*)up(cstart.data(), cstart.size());
  pl->t.tend = (const double*)up(tend.data(), tend.size() * sizeof(double));
  pl->t.phase_tend = (const double*)up(phase_tend.data(), phase_tend.size() * sizeof(double));
  pl->t.grid_ts = (const double*)up(grid_ts.data(), grid_ts.size() * sizeof(double));
  pl->t.wait_time = (const double*)up(wait.data(), wait.size() * sizeof(double));
  pl->t.standing = (const double*)up(standing.data(), standing.size() * sizeof(double));
  double* d_coef = (double*)up(nullptr, (size_t)n_poly_total * 12 * sizeof(double));
  pl->t.coef = d_coef;
  const int* d_node_of = (const int*)up(node_of_poly.data(), node_of_poly.size() * sizeof(int));
  const double* d_nodes = (const double*)up(nodes.data(), nodes.size() * sizeof(double));
  const double* d_dur = (const double*)up(dur.data(), dur.size() * sizeof(double));
  if (err == cudaSuccess) {
    wbctraj::hermite_coeff_kernel<<<(n_poly_total * 3 + 255) / 256, 256, 0, h->stream>>>(n_poly_total, d_node_of, d_nodes, d_dur, d_coef);
    h->launches++;
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
  }
  if (err != cudaSuccess) {
    h->err = std::string("wbc_plan_create: ") + cudaGetErrorString(err);
    wbc_plan_destroy(pl);
    return WBC_ERR_CUDA;
  }
  *out = pl;
  return WBC_OK;
}

extern "C" int wbc_sample_trajectory(wbc_handle* h, const wbc_plan* plan, int64_t n, const int32_t* plan_index, const double* t,
                                     double* traj, uint8_t* contact, double* f_plan, double* t_eval, int32_t* status, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (!plan || n < 0 || (n > 0 && (!t || !traj || !contact))) return fail_arg(h, "wbc_sample_trajectory: plan, t, traj and contact are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  if (reinterpret_cast<uintptr_t>(contact) & 3u) return fail_arg(h, "wbc_sample_trajectory: contact must be 4-byte aligned");
  // instances per warp: 32 (lane = instance) is the throughput shape; small batches are latency bound on the 14 dependent binary
  // searches per instance, so they spread them over 4 or 16 lanes per instance
  static const int ipw_env = getenv("WBC_SAMPLE_IPW") ? atoi(getenv("WBC_SAMPLE_IPW")) : 0;
  const int ipw = ipw_env ? ipw_env : (n >= 32768 ? 32 : n >= 8192 ? 8 : 2);   // rollout of 4096 robots: 45.2 / 48.7 / 49.7 M robot-steps/s with 32 / 8 / 2
  const long long chunks = (n + ipw - 1) / ipw, blocks = (chunks + wbctraj::SAMPLE_WARPS - 1) / wbctraj::SAMPLE_WARPS, cap = (long long)h->sm_count * 16;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  cudaStream_t sst = (cudaStream_t)stream;
  if (ipw == 32) wbctraj::sample_kernel<32><<<grid, wbctraj::SAMPLE_WARPS * 32, 0, sst>>>(plan->t, n, plan_index, t, traj, contact, f_plan, t_eval, status);
  else if (ipw == 8) wbctraj::sample_kernel<8><<<grid, wbctraj::SAMPLE_WARPS * 32, 0, sst>>>(plan->t, n, plan_index, t, traj, contact, f_plan, t_eval, status);
  else wbctraj::sample_kernel<2><<<grid, wbctraj::SAMPLE_WARPS * 32, 0, sst>>>(plan->t, n, plan_index, t, traj, contact, f_plan, t_eval, status);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_sample_trajectory_host(wbc_handle* h, const wbc_plan* plan, int64_t n, const int32_t* plan_index, const double* t,
                                          double* traj, uint8_t* contact, double* f_plan, double* t_eval, int32_t* status) {
  if (!h) return WBC_ERR_ARG;
  if (!plan || n < 0 || (n > 0 && (!t || !traj || !contact))) return fail_arg(h, "wbc_sample_trajectory_host: plan, t, traj and contact are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  const int32_t* dpi = s.in(plan_index, N, st); const double* dt = s.in(t, N, st);
  double* dtr = s.out(traj, N * WBC_NTRAJ); uint8_t* dc = s.out(contact, N * 4); double* dfp = s.out(f_plan, N * 12);
  double* dte = s.out(t_eval, N); int32_t* dst = s.out(status, N);
  WBC_SCRATCH_CHECK(h, s);
  int rc = wbc_sample_trajectory(h, plan, n, dpi, dt, dtr, dc, dfp, dte, dst, st);
  if (rc) return rc;
  s.back(traj, dtr, N * WBC_NTRAJ, st); s.back(contact, dc, N * 4, st); s.back(f_plan, dfp, N * 12, st);
  s.back(t_eval, dte, N, st); s.back(status, dst, N, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

// Control step whose trunk targets come from a device-resident plan: the host sends q, v and the plan time only (300 B + 8 B
// per instance instead of 732 B - the 54-double trajectory row is 59 % of the step's input bytes), wbc_sample_trajectory fills
// traj / contact in device scratch and the step kernels read them from there. Page-locked buffers are read / written by the
// kernels directly (zero-copy), pageable ones are staged.
static int ensure_rollout_scratch(wbc_handle* h, int64_t n);
extern "C" int wbc_step_plan_host(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, const double* q, const double* v,
                                  const double* t, const int32_t* plan_index, double* tau, double* metrics, int32_t* status) {
  if (!h) return WBC_ERR_ARG;
  if (!plan || n < 0 || kind == WBC_CTRL_PD) return fail_arg(h, "wbc_step_plan_host: plan and a QP controller kind are required");
  if (n > 0 && (!q || !v || !t || !tau || !metrics || !status)) return fail_arg(h, "wbc_step_plan_host: q, v, t, tau, metrics and status are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_rollout_scratch(h, n);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  const void* ptrs[7] = {q, v, t, plan_index, tau, metrics, status};
  void* dev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool pinned = true;
  for (int i = 0; i < 7 && pinned; ++i) {
    if (!ptrs[i]) continue;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptrs[i]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) { pinned = false; cudaGetLastError(); }
    else dev[i] = at.devicePointer;
  }
  if (pinned) {
    rc = wbc_sample_trajectory(h, plan, n, (const int32_t*)dev[3], (const double*)dev[2], h->ro_traj, h->ro_contact, nullptr, nullptr, nullptr, st);
    if (rc) return rc;
    // the step itself as in the zero-copy mode of wbc_step_host: two halves on the two internal streams. q and v are 304 B per
    // instance, which the reduce CTAs pull over the link without slowing down at any batch size (57.3 M steps/s at 65536,
    // 60.8 M at 2^20; copy-engine staging of q and v: 51.5 / 60.6 M)
    const int zc = n >= 4096 ? 2 : 1;
    const int64_t per = ((n + zc - 1) / zc + 3) & ~(int64_t)3;
    cudaStream_t lanes[2] = {st, h->stream2};
    if (zc > 1) {
      if (!h->order_ev) WBC_CUDA(h, cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
      WBC_CUDA(h, cudaEventRecord(h->order_ev, st));                 // the sampled rows
      WBC_CUDA(h, cudaStreamWaitEvent(h->stream2, h->order_ev, 0));
    }
    for (int c = 0; c < zc; ++c) {
      const int64_t o = c * per, m = (o + per <= n) ? per : n - o;
      if (m <= 0) break;
      const wbc_io io{(const double*)dev[0] + o * WBC_NQ, (const double*)dev[1] + o * WBC_NV, h->ro_traj + o * WBC_NTRAJ, h->ro_contact + o * 4,
                      (double*)dev[4] + o * WBC_NU, (double*)dev[5] + o * WBC_NMETRIC, (int32_t*)dev[6] + o};
      h->host_mapped = true; h->side_by_side = zc > 1;
      rc = step_launch(h, kind, m, &io, lanes[c & 1], c & 1);
      h->host_mapped = false; h->side_by_side = false;
      if (rc) return rc;
    }
    return step_host_wait(h);
  }
  rc = ensure_staging(h, n);
  if (rc) return rc;
  DevScratch s;
  const int32_t* dpi = s.in(plan_index, (size_t)n, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaMemcpyAsync(h->d_q, q, n * WBC_NQ * sizeof(double), cudaMemcpyHostToDevice, st));
  WBC_CUDA(h, cudaMemcpyAsync(h->d_v, v, n * WBC_NV * sizeof(double), cudaMemcpyHostToDevice, st));
  WBC_CUDA(h, cudaMemcpyAsync(h->ro_t, t, n * sizeof(double), cudaMemcpyHostToDevice, st));
  rc = wbc_sample_trajectory(h, plan, n, dpi, h->ro_t, h->ro_traj, h->ro_contact, nullptr, nullptr, nullptr, st);
  if (rc) return rc;
  const wbc_io io{h->d_q, h->d_v, h->ro_traj, h->ro_contact, h->d_tau, h->d_metrics, h->d_status};
  rc = step_launch(h, kind, n, &io, st, 0);
  if (rc) return rc;
  WBC_CUDA(h, cudaMemcpyAsync(tau, h->d_tau, n * WBC_NU * sizeof(double), cudaMemcpyDeviceToHost, st));
  WBC_CUDA(h, cudaMemcpyAsync(metrics, h->d_metrics, n * WBC_NMETRIC * sizeof(double), cudaMemcpyDeviceToHost, st));
  WBC_CUDA(h, cudaMemcpyAsync(status, h->d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

// ------------------------------------------------------------------------------ closed-loop rollout (wbc_rollout.cuh)
extern "C" int wbc_integrate(wbc_handle* h, int64_t n, double dt, double* q, double* v, const double* vd, double* t, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || (n > 0 && (!q || !v || !vd))) return fail_arg(h, "wbc_integrate: q, v and vd are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  wbcroll::integrate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, dt, q, v, vd, t, nullptr, nullptr, nullptr,
                                                                                          nullptr, nullptr, nullptr);
  h->launches++;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

static int ensure_rollout_scratch(wbc_handle* h, int64_t n) {
  if (!h->ro_counter) WBC_CUDA(h, cudaMalloc(&h->ro_counter, sizeof(int)));
  if (n <= h->ro_cap) return WBC_OK;
  cudaFree(h->ro_traj); cudaFree(h->ro_vd); cudaFree(h->ro_metrics); cudaFree(h->ro_tau); cudaFree(h->ro_t);
  cudaFree(h->ro_contact); cudaFree(h->ro_status);
  h->ro_traj = h->ro_vd = h->ro_metrics = h->ro_tau = h->ro_t = nullptr; h->ro_contact = nullptr; h->ro_status = nullptr; h->ro_cap = 0;
  WBC_CUDA(h, cudaMalloc(&h->ro_traj, n * WBC_NTRAJ * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->ro_vd, n * WBC_NV * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->ro_metrics, n * WBC_NMETRIC * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->ro_tau, n * WBC_NU * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->ro_t, n * sizeof(double)));
  WBC_CUDA(h, cudaMalloc(&h->ro_contact, n * 4));
  WBC_CUDA(h, cudaMalloc(&h->ro_status, n * sizeof(int32_t)));
  h->ro_cap = n;
  return WBC_OK;
}

extern "C" int wbc_default_plant_opts(wbc_plant_opts* o) {
  if (!o) return WBC_ERR_ARG;
  o->mu = 1.0; o->erp = 0.2; o->iters = 30; o->reserved = 0;      // simulate.py:44-46: static = dynamic friction 1.0
  return WBC_OK;
}

static int plant_launch(wbc_handle* h, int64_t n, double dt, const wbc_plant_opts* opts, double* q, double* v, const double* tau,
                        double* t, const int32_t* ctrl_status, int32_t* status_or, double* f_contact, const double* metrics,
                        double* err_max, double* metrics_log, const int* counter, cudaStream_t st) {
  wbc_plant_opts o;
  if (opts) o = *opts; else wbc_default_plant_opts(&o);
  if (!(o.mu >= 0.0) || !(o.erp >= 0.0 && o.erp <= 1.0) || o.iters < 1 || o.iters > 1000) return fail_arg(h, "wbc_plant_step: bad options");
  const wbcplant::PlantArgs a{q, v, tau, t, ctrl_status, status_or, f_contact, metrics, err_max, metrics_log, counter, (long long)n, dt,
                              o.mu, o.erp, o.iters};
  wbc_plant_kernel<<<(unsigned)n, 32, sizeof(SmemLayoutPlant), st>>>(h->d_const, a);
  h->launches++;
  return WBC_OK;
}

extern "C" int wbc_plant_step(wbc_handle* h, int64_t n, double dt, const wbc_plant_opts* opts, double* q, double* v, const double* tau,
                              double* t, const int32_t* ctrl_status, int32_t* status_or, double* f_contact, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || !(dt > 0.0) || (n > 0 && (!q || !v || !tau))) return fail_arg(h, "wbc_plant_step: q, v, tau and dt > 0 are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  const int rc = plant_launch(h, n, dt, opts, q, v, tau, t, ctrl_status, status_or, f_contact, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
  if (rc) return rc;
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_plant_step_host(wbc_handle* h, int64_t n, double dt, const wbc_plant_opts* opts, double* q, double* v, const double* tau,
                                   double* t, const int32_t* ctrl_status, int32_t* status_or, double* f_contact) {
  if (!h) return WBC_ERR_ARG;
  if (n < 0 || !(dt > 0.0) || (n > 0 && (!q || !v || !tau))) return fail_arg(h, "wbc_plant_step_host: q, v, tau and dt > 0 are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  double* dq = s.in(q, N * WBC_NQ, st); double* dv = s.in(v, N * WBC_NV, st); const double* dtau = s.in(tau, N * WBC_NU, st);
  double* dt_ = s.in(t, N, st); const int32_t* dcs = s.in(ctrl_status, N, st); int32_t* dso = s.in(status_or, N, st);
  double* df = s.out(f_contact, N * 12);
  WBC_SCRATCH_CHECK(h, s);
  const int rc = wbc_plant_step(h, n, dt, opts, dq, dv, dtau, dt_, dcs, dso, df, st);
  if (rc) return rc;
  s.back(q, dq, N * WBC_NQ, st); s.back(v, dv, N * WBC_NV, st); s.back(t, dt_, N, st); s.back(status_or, dso, N, st);
  s.back(f_contact, df, N * 12, st);
  WBC_SCRATCH_CHECK(h, s);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

extern "C" int wbc_rollout_ex(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                              const wbc_rollout_io* io, const wbc_rollout_opts* opts, void* stream) {
  if (!h) return WBC_ERR_ARG;
  if (!plan || !io || n < 0 || n_steps < 0 || !(dt > 0.0)) return fail_arg(h, "wbc_rollout: bad arguments");
  const bool plant = opts && opts->plant != 0;
  const int use_graph = opts ? opts->use_graph : 1;
  if (kind == WBC_CTRL_PD && !plant) return fail_arg(h, "wbc_rollout: the PD law returns no accelerations to integrate (use the ground plant)");
  if (n > 0 && (!io->q || !io->v || !io->t)) return fail_arg(h, "wbc_rollout: q, v and t are required");
  if (n == 0 || n_steps == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_rollout_scratch(h, n);
  if (rc) return rc;
  rc = ensure_split_scratch(h, 0, n, !plant);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  double* tau = io->tau ? io->tau : h->ro_tau;
  double* metrics = io->metrics ? io->metrics : h->ro_metrics;
  WBC_CUDA(h, cudaMemsetAsync(h->ro_counter, 0, sizeof(int), st));
  if (io->status_or) WBC_CUDA(h, cudaMemsetAsync(io->status_or, 0, n * sizeof(int32_t), st));
  if (io->err_max) WBC_CUDA(h, cudaMemsetAsync(io->err_max, 0, n * sizeof(double), st));
  if (kind == WBC_CTRL_PD) {                       // the PD law has no QP: no status / metrics of its own
    WBC_CUDA(h, cudaMemsetAsync(h->ro_status, 0, n * sizeof(int32_t), st));
    WBC_CUDA(h, cudaMemsetAsync(metrics, 0, n * WBC_NMETRIC * sizeof(double), st));
  }
  // with the ground plant the controller's accelerations are not needed: the step runs without the vd outputs
  wbc_io sio{io->q, io->v, h->ro_traj, h->ro_contact, tau, metrics, h->ro_status, plant ? nullptr : h->ro_vd, nullptr, nullptr};
  auto one_step = [&]() -> int {
    int r = wbc_sample_trajectory(h, plan, n, io->plan_index, io->t, h->ro_traj, h->ro_contact, nullptr, nullptr, nullptr, st);
    if (r) return r;
    r = wbc_step(h, kind, n, &sio, st);
    if (r) return r;
    if (plant) {
      r = plant_launch(h, n, dt, &opts->plant_opts, io->q, io->v, tau, io->t, h->ro_status, io->status_or, opts->f_contact, metrics,
                       io->err_max, io->metrics_log, h->ro_counter, st);
      if (r) return r;
    } else {
      wbcroll::integrate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, dt, io->q, io->v, h->ro_vd, io->t, h->ro_status, io->status_or,
                                                                           metrics, io->err_max, io->metrics_log, h->ro_counter);
      h->launches++;
    }
    if (io->metrics_log) { wbcroll::bump_counter_kernel<<<1, 1, 0, st>>>(h->ro_counter); h->launches++; }
    return WBC_OK;
  };
  if (use_graph && n_steps > 1 && st != nullptr && st != cudaStreamLegacy) {   // the legacy default stream cannot be captured
    // capture one control step (3-4 kernels) once, replay it n_steps times: one graph launch per step instead of 3-4 kernel launches
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const int64_t launches0 = h->launches;
    WBC_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    rc = one_step();
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    const int64_t per_step = h->launches - launches0;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { h->err = std::string("wbc_rollout: graph capture: ") + cudaGetErrorString(e); return WBC_ERR_CUDA; }
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); h->err = std::string("wbc_rollout: graph instantiate: ") + cudaGetErrorString(e); return WBC_ERR_CUDA; }
    h->launches = launches0;
    for (int k = 0; k < n_steps && e == cudaSuccess; ++k) { e = cudaGraphLaunch(exec, st); h->launches += per_step; }
    // the exec graph must outlive its launches: wait for the stream before destroying it
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { h->err = std::string("wbc_rollout: graph launch: ") + cudaGetErrorString(e); return WBC_ERR_CUDA; }
  } else {
    for (int k = 0; k < n_steps; ++k) {
      rc = one_step();
      if (rc) return rc;
    }
  }
  WBC_CUDA(h, cudaGetLastError());
  return WBC_OK;
}

extern "C" int wbc_rollout(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                           const wbc_rollout_io* io, int use_graph, void* stream) {
  wbc_rollout_opts o{};
  o.use_graph = use_graph; o.plant = 0; o.f_contact = nullptr;
  wbc_default_plant_opts(&o.plant_opts);
  return wbc_rollout_ex(h, kind, plan, n, n_steps, dt, io, &o, stream);
}

extern "C" int wbc_rollout_ex_host(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                                   const wbc_rollout_io* io, const wbc_rollout_opts* opts) {
  if (!h) return WBC_ERR_ARG;
  if (!plan || !io || n < 0 || n_steps < 0) return fail_arg(h, "wbc_rollout_host: bad arguments");
  if (n > 0 && (!io->q || !io->v || !io->t)) return fail_arg(h, "wbc_rollout_host: q, v and t are required");
  if (n == 0) return WBC_OK;
  WBC_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevScratch s;
  const size_t N = (size_t)n;
  wbc_rollout_io d{};
  d.q = s.in(io->q, N * WBC_NQ, st); d.v = s.in(io->v, N * WBC_NV, st); d.t = s.in(io->t, N, st);
  d.plan_index = s.in(io->plan_index, N, st);
  DevScratch s2;     // DevScratch holds 8 buffers
  d.tau = s2.out(io->tau, N * WBC_NU); d.metrics = s2.out(io->metrics, N * WBC_NMETRIC);
  d.status_or = s2.out(io->status_or, N); d.err_max = s2.out(io->err_max, N);
  d.metrics_log = s2.out(io->metrics_log, N * WBC_NMETRIC * (size_t)n_steps);
  wbc_rollout_opts o{};
  if (opts) o = *opts; else { o.use_graph = 1; wbc_default_plant_opts(&o.plant_opts); }
  double* host_f = o.f_contact;
  o.f_contact = s2.out(host_f, N * 12);
  WBC_SCRATCH_CHECK(h, s); WBC_SCRATCH_CHECK(h, s2);
  int rc = wbc_rollout_ex(h, kind, plan, n, n_steps, dt, &d, &o, st);
  if (rc) return rc;
  s.back(io->q, d.q, N * WBC_NQ, st); s.back(io->v, d.v, N * WBC_NV, st); s.back(io->t, d.t, N, st);
  s2.back(io->tau, d.tau, N * WBC_NU, st); s2.back(io->metrics, d.metrics, N * WBC_NMETRIC, st);
  s2.back(io->status_or, d.status_or, N, st); s2.back(io->err_max, d.err_max, N, st);
  s2.back(io->metrics_log, d.metrics_log, N * WBC_NMETRIC * (size_t)n_steps, st);
  s2.back(host_f, o.f_contact, N * 12, st);
  WBC_SCRATCH_CHECK(h, s); WBC_SCRATCH_CHECK(h, s2);
  WBC_CUDA(h, cudaStreamSynchronize(st));
  return WBC_OK;
}

extern "C" int wbc_rollout_host(wbc_handle* h, int kind, const wbc_plan* plan, int64_t n, int32_t n_steps, double dt,
                                const wbc_rollout_io* io, int use_graph) {
  wbc_rollout_opts o{};
  o.use_graph = use_graph; o.plant = 0; o.f_contact = nullptr;
  wbc_default_plant_opts(&o.plant_opts);
  return wbc_rollout_ex_host(h, kind, plan, n, n_steps, dt, io, &o);
}

extern "C" void* wbc_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void wbc_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int wbc_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return WBC_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return WBC_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return WBC_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return WBC_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fp64_peak_kernel<<<blocks, threads>>>(out, 1024);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return WBC_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return WBC_OK;
}
