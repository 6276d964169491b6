/* mini-MPI: the subset of MPI the reference (TRC-HPC/LFM_Public) calls, so that its sources compile and
 * run UNCHANGED in an image without an MPI installation (SURVEY.md section 8(c) lists the symbols).
 *
 * TEST INFRASTRUCTURE ONLY -- used to build oracle/_ref (the reference's own CPU solver).  The product
 * (lfm_public_b200/) never includes this header.
 *
 * Ranks are forked from the first process inside MPI_Init when LFM_MPI_NP=<n> is set (no mpirun);
 * point-to-point runs over AF_UNIX socket pairs with a progress engine, collectives over a shared
 * anonymous mapping.  One-sided communication is served for general active target synchronisation
 * (MPI_Win_post / start / complete / wait + MPI_Get / MPI_Rget: the reference's haloCommType 3/4), emulated
 * with messages (mpi_shim.cpp); passive target locks and neighbourhood collectives are link-only stubs
 * that abort.
 */
#ifndef LFM_MINI_MPI_H
#define LFM_MINI_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Win;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef struct MPI_Status {
	int MPI_SOURCE, MPI_TAG, MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_INFO_NULL 0
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MODE_NOCHECK 1
#define MPI_MODE_NOPUT 2
#define MPI_LOCK_SHARED 1

/* predefined datatypes: handle = id; sizes in mpi_shim.cpp */
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_UNSIGNED 3
#define MPI_UNSIGNED_SHORT 4
#define MPI_FLOAT 5
#define MPI_DOUBLE 6
#define MPI_LOGICAL 7
#define MPI_BYTE 8

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

int MPI_Init(int* argc, char*** argv);
int MPI_Initialized(int* flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Get_processor_name(char* name, int* len);
int MPI_Pcontrol(const int level, ...);

int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Gather(const void* sbuf, int scount, MPI_Datatype st, void* rbuf, int rcount, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype st, void* rbuf, int rcount, MPI_Datatype rt, MPI_Comm comm);

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Send_init(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Recv_init(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Start(MPI_Request* req);
int MPI_Startall(int n, MPI_Request* reqs);
int MPI_Wait(MPI_Request* req, MPI_Status* st);
int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* st);
int MPI_Testall(int n, MPI_Request* reqs, int* flag, MPI_Status* st);

int MPI_Type_create_struct(int n, const int* blocklens, const MPI_Aint* disps, const MPI_Datatype* types, MPI_Datatype* newtype);
int MPI_Type_create_resized(MPI_Datatype old, MPI_Aint lb, MPI_Aint extent, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* t);

/* link-only (abort when called) */
int MPI_Win_create(void* base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win* win);
int MPI_Win_free(MPI_Win* win);
int MPI_Win_post(MPI_Group g, int assert_, MPI_Win win);
int MPI_Win_start(MPI_Group g, int assert_, MPI_Win win);
int MPI_Win_complete(MPI_Win win);
int MPI_Win_wait(MPI_Win win);
int MPI_Win_lock(int type, int rank, int assert_, MPI_Win win);
int MPI_Win_lock_all(int assert_, MPI_Win win);
int MPI_Get(void* o, int oc, MPI_Datatype ot, int rank, MPI_Aint disp, int tc, MPI_Datatype tt, MPI_Win win);
int MPI_Rget(void* o, int oc, MPI_Datatype ot, int rank, MPI_Aint disp, int tc, MPI_Datatype tt, MPI_Win win, MPI_Request* req);
int MPI_Comm_group(MPI_Comm comm, MPI_Group* g);
int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out);
int MPI_Group_free(MPI_Group* g);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* out);
int MPI_Dist_graph_create_adjacent(MPI_Comm comm, int indeg, const int* src, const int* sw, int outdeg, const int* dst,
                                   const int* dw, MPI_Info info, int reorder, MPI_Comm* out);
int MPI_Dist_graph_neighbors(MPI_Comm comm, int maxin, int* src, int* sw, int maxout, int* dst, int* dw);

#ifdef __cplusplus
}
#endif
#endif
