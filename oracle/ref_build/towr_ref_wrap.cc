// ORACLE SUPPORT (test infrastructure): thin extern "C" wrapper around the REFERENCE's own TOWR sources, compiled where
// they lie under /root/reference by oracle/ref_build/Makefile into oracle/_ref/libtowr_ref.so. Used only to pin
// oracle/trajectory.py (and to regenerate tests/golden/towr_spline.npz); never linked or loaded by the product path.
//
// Reference sources compiled: towr/src/{state,polynomial,spline}.cc (against the Eigen stand-in next to this file) and
// towr/src/{gait_generator,quadruped_gait_generator,monoped_gait_generator,biped_gait_generator}.cc (plain STL).
// phase_durations.cc needs ifopt (absent): IsContactPhase (phase_durations.cc:120-124) is restated below on top of the
// reference's Spline::GetSegmentID.
#include <vector>

#include <towr/initialization/gait_generator.h>
#include <towr/variables/spline.h>

namespace {
struct RefSpline : public towr::Spline {
  RefSpline(const VecTimes& d, int dim) : towr::Spline(d, dim) {}
  void SetNodes(int poly, const towr::Node& n0, const towr::Node& n1) { cubic_polys_.at(poly).SetNodes(n0, n1); }
  void Update() { UpdatePolynomialCoeff(); }
};
}  // namespace

extern "C" {

// GaitGenerator::SetCombo + GetPhaseDurations(t_total, ee) (towr/trunk_mpc.cpp:131-137). Returns the phase count.
int towr_ref_phase_durations(int combo, double t_total, int ee, double* out, int cap) {
  auto gen = towr::GaitGenerator::MakeGaitGenerator(4);
  gen->SetCombo(static_cast<towr::GaitGenerator::Combos>(combo));
  const std::vector<double> d = gen->GetPhaseDurations(t_total, ee);
  for (int i = 0; i < static_cast<int>(d.size()) && i < cap; ++i) out[i] = d[i];
  return static_cast<int>(d.size());
}

int towr_ref_contact_at_start(int combo, int ee) {
  auto gen = towr::GaitGenerator::MakeGaitGenerator(4);
  gen->SetCombo(static_cast<towr::GaitGenerator::Combos>(combo));
  return gen->IsInContactAtStart(ee) ? 1 : 0;
}

int towr_ref_segment_id(double t, const double* durations, int n) {
  return towr::Spline::GetSegmentID(t, std::vector<double>(durations, durations + n));
}

// PhaseDurations::IsContactPhase (phase_durations.cc:120-124)
int towr_ref_is_contact_phase(double t, const double* durations, int n, int contact_at_start) {
  const int phase = towr::Spline::GetSegmentID(t, std::vector<double>(durations, durations + n));
  return (phase % 2 == 0) ? (contact_at_start != 0) : (contact_at_start == 0);
}

// Spline::GetPoint(t) of a 3-D cubic Hermite spline with nodes[(n_poly + 1)][6] = (p, v): out[9] = p, v, a.
void towr_ref_spline_point(int n_poly, const double* durations, const double* nodes, int n_t, const double* t, double* out) {
  RefSpline s(std::vector<double>(durations, durations + n_poly), 3);
  for (int i = 0; i < n_poly; ++i) {
    towr::Node n0(3), n1(3);
    for (int d = 0; d < 3; ++d) {
      n0.at(towr::kPos)(d) = nodes[6 * i + d];       n0.at(towr::kVel)(d) = nodes[6 * i + 3 + d];
      n1.at(towr::kPos)(d) = nodes[6 * (i + 1) + d]; n1.at(towr::kVel)(d) = nodes[6 * (i + 1) + 3 + d];
    }
    s.SetNodes(i, n0, n1);
  }
  s.Update();
  for (int k = 0; k < n_t; ++k) {
    const towr::State st = s.GetPoint(t[k]);
    for (int d = 0; d < 3; ++d) {
      out[9 * k + d] = st.p()(d); out[9 * k + 3 + d] = st.v()(d); out[9 * k + 6 + d] = st.a()(d);
    }
  }
}

}  // extern "C"
