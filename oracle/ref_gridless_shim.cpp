// ref_gridless_shim.cpp -- C entry points around the REFERENCE's own stand-alone movers
// (srcEarth/gridless/GridlessParticleMovers.cpp), compiled from the sources where they lie under $(REF) by
// oracle/Makefile into oracle/_ref/libref_gridless.so.  TEST INFRASTRUCTURE: used to check the physics of the
// restated Relativistic::Boris (gyration, |p| conservation) against code of the reference itself.  The gridless
// BorisStep is a drift-kick-drift splitting (:781-802), not the kick-drift of pic_mover_relativistic_boris.cpp,
// so the momentum rotation is comparable bit-for-bit in intent but positions differ at O(dt^2).
#include <cmath>

#include "GridlessParticleMovers.h"

namespace {
class UniformB : public IGridlessFieldEvaluator {
 public:
  V3 B;
  void GetB_T(const V3 &, V3 &B_T) const override { B_T = B; }
};
class DipoleB : public IGridlessFieldEvaluator {
 public:
  double B0, R0;  // equatorial surface field [T], body radius [m]; moment along -z
  void GetB_T(const V3 &x, V3 &B_T) const override {
    const double r2 = x.x * x.x + x.y * x.y + x.z * x.z, r = std::sqrt(r2), r5 = r2 * r2 * r;
    const double k = -B0 * R0 * R0 * R0;  // m_z (up to mu0/4pi)
    B_T.x = k * 3.0 * x.z * x.x / r5;
    B_T.y = k * 3.0 * x.z * x.y / r5;
    B_T.z = k * (3.0 * x.z * x.z - r2) / r5;
  }
};
}  // namespace

extern "C" {
// n Boris steps in a uniform field; x[3] (m), p[3] (kg m/s) are updated in place
void ref_boris_uniform(double *x, double *p, double q_C, double m_kg, double dt, const double *B_T, int n) {
  UniformB f;
  f.B = {B_T[0], B_T[1], B_T[2]};
  V3 xx{x[0], x[1], x[2]}, pp{p[0], p[1], p[2]};
  for (int i = 0; i < n; i++) BorisStep(xx, pp, q_C, m_kg, dt, f);
  x[0] = xx.x, x[1] = xx.y, x[2] = xx.z, p[0] = pp.x, p[1] = pp.y, p[2] = pp.z;
}
void ref_boris_dipole(double *x, double *p, double q_C, double m_kg, double dt, double B0, double R0, int n) {
  DipoleB f;
  f.B0 = B0, f.R0 = R0;
  V3 xx{x[0], x[1], x[2]}, pp{p[0], p[1], p[2]};
  for (int i = 0; i < n; i++) BorisStep(xx, pp, q_C, m_kg, dt, f);
  x[0] = xx.x, x[1] = xx.y, x[2] = xx.z, p[0] = pp.x, p[1] = pp.y, p[2] = pp.z;
}
}
