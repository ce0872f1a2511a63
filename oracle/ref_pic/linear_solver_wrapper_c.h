// Stand-in for SWMF share/Library/src/linear_solver_wrapper_c.h (un-vendored), OURS: the GMRES of the field solve is not on
// the particle path; calling it in this build is an error.
#pragma once
#include <cstdlib>
template <class... A>
inline void linear_solver_wrapper(A...) { abort(); }
// the matrix-free operator the solver calls back (set by the reference before every Solve)
extern void (*linear_solver_matvec_c)(double *VecIn, double *VecOut, int n);
