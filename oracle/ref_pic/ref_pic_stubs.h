// Additions to the single-process mpi.h stand-in (oracle/ref_mesh/mpi.h) that the PIC translation units need. OURS (test
// infrastructure): the reference is run as ONE rank, so every collective is a local copy and no point-to-point message exists.
#pragma once
#include <climits>
#include "mpi.h"
#ifndef MPI_UNSIGNED_SHORT
#define MPI_UNSIGNED_SHORT 15
#define MPI_LONG_DOUBLE 16
#endif
typedef int MPI_Fint;
typedef int MPI_Info;
#define MPI_BOTTOM ((void *)0)
#define MPI_COMM_SELF 1
#define MPI_INFO_NULL 0
#define MPI_COMM_TYPE_SHARED 1
#define MPI_THREAD_FUNNELED 1
#ifdef __cplusplus
extern "C" {
#endif
MPI_Fint MPI_Comm_c2f(MPI_Comm);
MPI_Comm MPI_Comm_f2c(MPI_Fint);
int MPI_Init_thread(int *, char ***, int, int *);
int MPI_Allgatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm);
int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *);
int MPI_Type_indexed(int, const int *, const int *, MPI_Datatype, MPI_Datatype *);
MPI_Aint MPI_Aint_diff(MPI_Aint, MPI_Aint);
int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm *);
#ifdef __cplusplus
}
#endif

/* the gyrokinetic variant (REF_PIC_VARIANT=gk) routes MoveParticles through the shim so that a test can choose the guiding-centre
   mover of the reference at run time (ref_pic_shim.cpp); node is a cTreeNodeAMR<PIC::Mesh::cDataBlockAMR>* */
#ifdef __cplusplus
extern "C" int ref_pic_user_mover(long int ptr, double dt, void *node);
#endif
