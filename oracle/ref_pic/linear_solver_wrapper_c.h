// Stand-in for SWMF share/Library/src/linear_solver_wrapper_c.h (un-vendored: Config.pl clones SWMFsoftware/share at HEAD). OURS,
// test infrastructure.  The reference calls
//     linear_solver_wrapper("GMRES", &Tol, &nMaxIter, &nVar, &nDim, &nI, &nJ, &nK, &nBlock, &iComm, Rhs_I, Sol_I, &PrecondParam, NULL, &lTest)
// (srcInterface/LinearSystemCornerNode.h:3282) with its matrix-free operator in linear_solver_matvec_c.  The SWMF routine is a
// restarted GMRES without preconditioner that stops on the relative residual; gmres_single.cpp implements that published
// algorithm (Saad & Schultz 1986, modified Gram-Schmidt + Givens rotations) for one rank.
#pragma once
extern void (*linear_solver_matvec_c)(double *VecIn, double *VecOut, int n);
void linear_solver_wrapper(const char *method, double *Tol, int *nMaxIter, int *nVar, int *nDim, int *nI, int *nJ, int *nK, int *nBlock, int *iComm,
                           double *Rhs_I, double *Sol_I, double *PrecondParam, double *precond_matrix, int *lTest);
// what the last call did (read by the shim)
extern int ref_gmres_last_iterations;
extern double ref_gmres_last_relative_residual;
