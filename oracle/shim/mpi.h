/* TEST INFRASTRUCTURE — single-rank stand-in for <mpi.h>, used only to compile the
 * UNMODIFIED reference translation units (main.cpp / cuda.cu) into oracle/_ref/.
 * The image has no MPI.  Semantics are those of a communicator of size 1
 * (SURVEY.md Appendix A lists every call site and what it needs):
 *   - Allreduce/Iallreduce with MPI_IN_PLACE: identity; with distinct buffers: copy.
 *   - Allgather/Iallgather/Alltoall/Reduce: copy send -> recv.
 *   - Exscan: leave recv untouched.   - Test: flag = 1.
 *   - point-to-point: never reached with data at size 1 (empty neighbour sets /
 *     MPI_PROC_NULL): no-ops.
 *   - MPI_File_*: plain stdio.
 * One extension: a translation unit that defines CUP2D_REF_HOOK_TU before including
 * this header gets `cup2d_ref_hook(op, buf, count)` called from its MPI_Allreduce
 * (the harness uses the unique MPI_MAX reduction at main.cpp:6592 = "start of a time
 * step, after umax" to seed / dump fields without touching the reference source).
 */
#ifndef CUP2D_ORACLE_MPI_SHIM_H
#define CUP2D_ORACLE_MPI_SHIM_H
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Op;
typedef int MPI_Info;
typedef struct { int count_bytes; } MPI_Status;
typedef FILE *MPI_File;
typedef long long MPI_Offset;

#define MPI_COMM_WORLD 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_PROC_NULL (-2)
#define MPI_INFO_NULL 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 2
#define MPI_SUCCESS 0

/* datatypes: the value is the size in bytes */
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_LONG_LONG 108
#define MPI_LONG 208
#define MPI_UINT8_T 1
#define MPI_BYTE 301

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

static inline size_t cup2d_mpi_sizeof(MPI_Datatype t) {
  switch (t) {
  case MPI_DOUBLE: return 8;
  case MPI_INT: return 4;
  case MPI_LONG_LONG: return 8;
  case MPI_LONG: return sizeof(long);
  case MPI_UINT8_T: return 1;
  case MPI_BYTE: return 1;
  }
  fprintf(stderr, "mpi shim: unknown datatype %d\n", t);
  abort();
}

#ifdef CUP2D_REF_HOOK_TU
void cup2d_ref_hook(int op, void *buf, int count);
#endif

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op,
                                MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r)
    memcpy(r, s, n * cup2d_mpi_sizeof(t));
#ifdef CUP2D_REF_HOOK_TU
  cup2d_ref_hook(op, r, n);
#else
  (void)op;
#endif
  return 0;
}
static inline int MPI_Iallreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op,
                                 MPI_Comm, MPI_Request *q) {
  if (s != MPI_IN_PLACE && s != r)
    memcpy(r, s, n * cup2d_mpi_sizeof(t));
  if (q) *q = 0;
  return 0;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int,
                             MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r)
    memcpy(r, s, n * cup2d_mpi_sizeof(t));
  return 0;
}
static inline int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int,
                                MPI_Datatype, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r)
    memcpy(r, s, n * cup2d_mpi_sizeof(t));
  return 0;
}
static inline int MPI_Iallgather(const void *s, int n, MPI_Datatype t, void *r, int,
                                 MPI_Datatype, MPI_Comm, MPI_Request *q) {
  if (s != MPI_IN_PLACE && s != r)
    memcpy(r, s, n * cup2d_mpi_sizeof(t));
  if (q) *q = 0;
  return 0;
}
static inline int MPI_Alltoall(const void *s, int n, MPI_Datatype t, void *r, int,
                               MPI_Datatype, MPI_Comm) {
  if (s != MPI_IN_PLACE && s != r)
    memcpy(r, s, n * cup2d_mpi_sizeof(t));
  return 0;
}
static inline int MPI_Exscan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm) {
  return 0;
}
static inline int MPI_Isend(const void *, int n, MPI_Datatype, int dest, int, MPI_Comm,
                            MPI_Request *q) {
  if (dest != MPI_PROC_NULL && n > 0) {
    fprintf(stderr, "mpi shim: MPI_Isend with data reached at size 1\n");
    abort();
  }
  if (q) *q = 0;
  return 0;
}
static inline int MPI_Irecv(void *, int n, MPI_Datatype, int src, int, MPI_Comm,
                            MPI_Request *q) {
  if (src != MPI_PROC_NULL && n > 0) {
    fprintf(stderr, "mpi shim: MPI_Irecv with data reached at size 1\n");
    abort();
  }
  if (q) *q = 0;
  return 0;
}
static inline int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
static inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return 0; }
static inline int MPI_Test(MPI_Request *, int *flag, MPI_Status *) { *flag = 1; return 0; }
static inline int MPI_Probe(int, int, MPI_Comm, MPI_Status *st) {
  if (st) st->count_bytes = 0;
  return 0;
}
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *c) { *c = 0; return 0; }
static inline int MPI_File_open(MPI_Comm, const char *path, int, MPI_Info, MPI_File *f) {
  *f = fopen(path, "wb");
  return *f ? 0 : 1;
}
static inline int MPI_File_write_at_all(MPI_File f, MPI_Offset off, const void *buf, int n,
                                        MPI_Datatype t, MPI_Status *) {
  fseek(f, (long)off, SEEK_SET);
  fwrite(buf, cup2d_mpi_sizeof(t), n, f);
  return 0;
}
static inline int MPI_File_close(MPI_File *f) {
  fclose(*f);
  *f = NULL;
  return 0;
}
#endif
