// Host harness around csrc/wbc_device.cuh (see cuda_emu.h). Exports two C functions used by
// tests/test_emulator.py. TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"
#include "../../quadruped_drake_b200/csrc/wbc_device.cuh"
#include <vector>

thread_local int emu_lane;
thread_local EmuWarp* emu_warp;

namespace {
struct Job {
  EmuWarp* warp; int lane; wbc::WarpSmem* sm; wbc::PcSmem* pcs; double* Cout; double* Jdout; const wbc_model* md; const wbc_params* pr; wbc::Derived dv;
  wbc::StepArgs args; wbc::DynOut dyn; const double* q; const double* v; long long n; int mode;
};
void* lane_main(void* p) {
  Job* j = (Job*)p;
  emu_lane = j->lane;
  emu_warp = j->warp;
  for (long long i = 0; i < j->n; ++i) {
    if (j->mode == 0) {
      if (j->args.kind == WBC_CTRL_ID) wbc::step_instance<WBC_CTRL_ID>(*j->sm, *j->md, *j->pr, j->dv, j->args, i, j->lane);
      if (j->args.kind == WBC_CTRL_PC || j->args.kind == WBC_CTRL_MPTC) wbc::step_instance<WBC_CTRL_PC>(*j->sm, *j->md, *j->pr, j->dv, j->args, i, j->lane, j->pcs);
      if (j->args.kind == WBC_CTRL_CLF) wbc::step_instance<WBC_CTRL_CLF>(*j->sm, *j->md, *j->pr, j->dv, j->args, i, j->lane);
    } else if (j->mode == 1) {
      wbc::dynamics_instance(*j->sm, *j->md, j->q, j->v, j->dyn, i, j->lane);
    } else {
      wbc::coriolis_instance(*j->sm, *j->pcs, *j->md, j->q, j->v, j->Cout, j->Jdout, i, j->lane);
    }
  }
  return nullptr;
}
int run(Job proto) {
  EmuWarp warp;
  pthread_barrier_init(&warp.bar, nullptr, 32);
  wbc::WarpSmem* sm = new wbc::WarpSmem();
  memset(sm, 0, sizeof(*sm));
  wbc::PcSmem* pcs = new wbc::PcSmem();
  memset(pcs, 0, sizeof(*pcs));
  std::vector<Job> jobs(32, proto);
  std::vector<pthread_t> th(32);
  for (int l = 0; l < 32; ++l) {
    jobs[l].warp = &warp; jobs[l].lane = l; jobs[l].sm = sm; jobs[l].pcs = pcs;
    pthread_create(&th[l], nullptr, lane_main, &jobs[l]);
  }
  for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
  pthread_barrier_destroy(&warp.bar);
  delete sm;
  delete pcs;
  return 0;
}
}  // namespace

extern "C" int emu_step(const wbc_model* md, const wbc_params* pr, int kind, long long n, const wbc_io* io) {
  Job j{};
  wbc::derive_constants(*pr, j.dv);
  j.md = md; j.pr = pr; j.n = n; j.mode = 0;
  j.args.q = io->q; j.args.v = io->v; j.args.traj = io->traj; j.args.contact = io->contact;
  j.args.tau = io->tau; j.args.metrics = io->metrics; j.args.status = io->status;
  j.args.vd = io->vd; j.args.f = io->f; j.args.qp_info = io->qp_info; j.args.n = n; j.args.kind = kind;
  return run(j);
}
extern "C" int emu_dynamics(const wbc_model* md, long long n, const double* q, const double* v, double* M, double* Cv,
                            double* taug, double* Jfeet, double* Jdv, double* pfeet) {
  Job j{};
  j.md = md; j.n = n; j.mode = 1; j.q = q; j.v = v;
  j.dyn.M = M; j.dyn.Cv = Cv; j.dyn.taug = taug; j.dyn.Jfeet = Jfeet; j.dyn.Jdv = Jdv; j.dyn.pfeet = pfeet;
  return run(j);
}
extern "C" int emu_coriolis(const wbc_model* md, long long n, const double* q, const double* v, double* Cm, double* Jd) {
  Job j{};
  j.md = md; j.n = n; j.mode = 2; j.q = q; j.v = v; j.Cout = Cm; j.Jdout = Jd;
  return run(j);
}
extern "C" int emu_smem_bytes() { return (int)sizeof(wbc::WarpSmem); }
