// One-rank implementation of the MPI calls the reference makes (declared in oracle/ref_mesh/mpi.h and ref_pic_stubs.h). OURS,
// test infrastructure: collectives copy the send buffer, a message to another rank is an error (there is none), requests
// complete at once.  Only for running the reference's own PIC code as a single-process checker.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include "ref_pic_stubs.h"

static size_t tsize(MPI_Datatype t) {
  switch (t) {
    case MPI_BYTE: case MPI_CHAR: case MPI_UNSIGNED_CHAR: case MPI_C_BOOL: return 1;
    case MPI_SHORT: case MPI_UNSIGNED_SHORT: return 2;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
    case MPI_LONG_DOUBLE: case MPI_LONG_INT: return 16;
    default: return 8;
  }
}
static void bad(const char *what) {
  fprintf(stderr, "mpi_single: %s called in the one-rank build\n", what);
  abort();
}
static int copy(const void *s, void *r, int n, MPI_Datatype t) {
  if (s != MPI_IN_PLACE && s != r && n > 0) memcpy(r, s, (size_t)n * tsize(t));
  return 0;
}
extern "C" {
int MPI_Init(int *, char ***) { return 0; }
int MPI_Init_thread(int *, char ***, int, int *p) { if (p) *p = MPI_THREAD_FUNNELED; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm, int c) { fprintf(stderr, "MPI_Abort(%d)\n", c); abort(); }
int MPI_Barrier(MPI_Comm) { return 0; }
int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *o) { *o = c; return 0; }
int MPI_Comm_free(MPI_Comm *) { return 0; }
int MPI_Comm_split(MPI_Comm c, int, int, MPI_Comm *o) { *o = c; return 0; }
int MPI_Comm_split_type(MPI_Comm c, int, int, MPI_Info, MPI_Comm *o) { *o = c; return 0; }
MPI_Fint MPI_Comm_c2f(MPI_Comm c) { return c; }
MPI_Comm MPI_Comm_f2c(MPI_Fint c) { return c; }
int MPI_Initialized(int *f) { *f = 1; return 0; }
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { bad("MPI_Send"); return 1; }
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { bad("MPI_Recv"); return 1; }
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { bad("MPI_Isend"); return 1; }
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { bad("MPI_Irecv"); return 1; }
int MPI_Sendrecv(const void *, int, MPI_Datatype, int, int, void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { bad("MPI_Sendrecv"); return 1; }
int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return 0; }
int MPI_Waitany(int, MPI_Request *, int *i, MPI_Status *) { *i = MPI_UNDEFINED; return 0; }
int MPI_Test(MPI_Request *, int *f, MPI_Status *) { *f = 1; return 0; }
int MPI_Testany(int, MPI_Request *, int *i, int *f, MPI_Status *) { *i = MPI_UNDEFINED, *f = 1; return 0; }
int MPI_Testall(int, MPI_Request *, int *f, MPI_Status *) { *f = 1; return 0; }
int MPI_Request_free(MPI_Request *) { return 0; }
int MPI_Cancel(MPI_Request *) { return 0; }
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { return copy(s, r, n, t); }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { return copy(s, r, n, t); }
int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm) { return copy(s, r, n, t); }
int MPI_Gatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *displ, MPI_Datatype, int, MPI_Comm) {
  return copy(s, (char *)r + (displ ? (size_t)displ[0] * tsize(t) : 0), n, t);
}
int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm) { return copy(s, r, n, t); }
int MPI_Allgatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *displ, MPI_Datatype, MPI_Comm) {
  return copy(s, (char *)r + (displ ? (size_t)displ[0] * tsize(t) : 0), n, t);
}
int MPI_Scatter(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm) { return copy(s, r, n, t); }
int MPI_Alltoall(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm) { return copy(s, r, n, t); }
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *c) { *c = 0; return 0; }
int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { bad("MPI_Probe"); return 1; }
int MPI_Iprobe(int, int, MPI_Comm, int *f, MPI_Status *) { *f = 0; return 0; }
int MPI_Type_create_hindexed_block(int, int, const MPI_Aint *, MPI_Datatype, MPI_Datatype *o) { *o = MPI_BYTE; return 0; }
int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *o) { *o = MPI_BYTE; return 0; }
int MPI_Type_indexed(int, const int *, const int *, MPI_Datatype, MPI_Datatype *o) { *o = MPI_BYTE; return 0; }
int MPI_Type_contiguous(int, MPI_Datatype, MPI_Datatype *o) { *o = MPI_BYTE; return 0; }
int MPI_Type_commit(MPI_Datatype *) { return 0; }
int MPI_Type_free(MPI_Datatype *) { return 0; }
int MPI_Type_size(MPI_Datatype t, int *s) { *s = (int)tsize(t); return 0; }
int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)p; return 0; }
MPI_Aint MPI_Aint_diff(MPI_Aint a, MPI_Aint b) { return a - b; }
int MPI_Get_processor_name(char *n, int *l) { strcpy(n, "single"); *l = 6; return 0; }
double MPI_Wtime(void) {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
}
