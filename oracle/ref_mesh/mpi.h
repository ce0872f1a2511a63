#ifndef STUB_MPI_H
#define STUB_MPI_H
// minimal serial stand-in for <mpi.h>: enough for the reference's mesh headers to compile in a single process
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op; typedef int MPI_Request; typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL -1
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_ANY_SOURCE -1
#define MPI_ANY_TAG -1
#define MPI_REQUEST_NULL -1
#define MPI_UNDEFINED -32766
#define MPI_IN_PLACE ((void*)1)
#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_LONG 4
#define MPI_DOUBLE 5
#define MPI_UNSIGNED_LONG 6
#define MPI_UNSIGNED 7
#define MPI_FLOAT 8
#define MPI_UNSIGNED_CHAR 9
#define MPI_LONG_LONG 10
#define MPI_SHORT 11
#define MPI_UNSIGNED_LONG_LONG 12
#define MPI_LONG_INT 13
#define MPI_C_BOOL 14
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_LAND 5
#define MPI_BOR 6
#define MPI_MAX_PROCESSOR_NAME 256
#ifdef __cplusplus
extern "C" {
#endif
int MPI_Init(int*, char***); int MPI_Finalize(void); int MPI_Abort(MPI_Comm,int); int MPI_Barrier(MPI_Comm);
int MPI_Comm_rank(MPI_Comm,int*); int MPI_Comm_size(MPI_Comm,int*); int MPI_Comm_dup(MPI_Comm, MPI_Comm*); int MPI_Comm_free(MPI_Comm*);
int MPI_Send(const void*,int,MPI_Datatype,int,int,MPI_Comm); int MPI_Recv(void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Status*);
int MPI_Isend(const void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Request*); int MPI_Irecv(void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Request*);
int MPI_Wait(MPI_Request*,MPI_Status*); int MPI_Waitall(int,MPI_Request*,MPI_Status*); int MPI_Waitany(int,MPI_Request*,int*,MPI_Status*);
int MPI_Test(MPI_Request*,int*,MPI_Status*); int MPI_Testany(int,MPI_Request*,int*,int*,MPI_Status*); int MPI_Testall(int,MPI_Request*,int*,MPI_Status*);
int MPI_Bcast(void*,int,MPI_Datatype,int,MPI_Comm); int MPI_Reduce(const void*,void*,int,MPI_Datatype,MPI_Op,int,MPI_Comm);
int MPI_Allreduce(const void*,void*,int,MPI_Datatype,MPI_Op,MPI_Comm); int MPI_Gather(const void*,int,MPI_Datatype,void*,int,MPI_Datatype,int,MPI_Comm);
int MPI_Gatherv(const void*,int,MPI_Datatype,void*,const int*,const int*,MPI_Datatype,int,MPI_Comm);
int MPI_Allgather(const void*,int,MPI_Datatype,void*,int,MPI_Datatype,MPI_Comm); int MPI_Scatter(const void*,int,MPI_Datatype,void*,int,MPI_Datatype,int,MPI_Comm);
int MPI_Sendrecv(const void*,int,MPI_Datatype,int,int,void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Status*);
int MPI_Get_count(const MPI_Status*,MPI_Datatype,int*); int MPI_Probe(int,int,MPI_Comm,MPI_Status*); int MPI_Iprobe(int,int,MPI_Comm,int*,MPI_Status*);
int MPI_Type_create_hindexed_block(int,int,const MPI_Aint*,MPI_Datatype,MPI_Datatype*); int MPI_Type_commit(MPI_Datatype*); int MPI_Type_free(MPI_Datatype*);
int MPI_Type_contiguous(int,MPI_Datatype,MPI_Datatype*); int MPI_Type_size(MPI_Datatype,int*);
int MPI_Get_address(const void*,MPI_Aint*); int MPI_Get_processor_name(char*,int*); double MPI_Wtime(void);
int MPI_Request_free(MPI_Request*); int MPI_Cancel(MPI_Request*); int MPI_Initialized(int*);
int MPI_Alltoall(const void*,int,MPI_Datatype,void*,int,MPI_Datatype,MPI_Comm);
int MPI_Comm_split(MPI_Comm,int,int,MPI_Comm*);
#ifdef __cplusplus
}
#endif
#endif
