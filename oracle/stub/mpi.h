/* Single-process stand-in for <mpi.h>, TEST INFRASTRUCTURE ONLY.
 *
 * It exists so that the UNMODIFIED reference sources src/dense.cpp, src/tet.cpp
 * and src/volume.cpp (under /root/reference) compile and run in one process
 * without an MPI installation (none exists in this image).  Only the symbols
 * those three files (and the headers they pull in) touch are provided:
 *   - rank/size/barrier/wtime:           one rank, rank 0
 *   - Allreduce / Reduce:                copy sendbuf -> recvbuf
 *   - MPI-IO with a C-order subarray view: implemented on top of pwrite so the
 *     reference's own WriteGrid (dense.cpp:751-870) produces a real dense.raw
 *   - Isend / Recv / Waitall:            abort (never reached with one rank,
 *     see dense.cpp:974-987: rank == root_rank always)
 * Nothing in the product (tess2_b200/) includes this file.
 */
#ifndef TESSB200_ORACLE_MPI_STUB_H
#define TESSB200_ORACLE_MPI_STUB_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <fcntl.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Info;
typedef int MPI_Request;
typedef long long MPI_Offset;

typedef struct { int count; } MPI_Status;

/* datatypes: basic ones are small ints holding their size class; derived
 * (subarray) types are indices >= 16 into a small table */
typedef int MPI_Datatype;
typedef int MPI_Op;

struct mpistub_file { int fd; MPI_Datatype view; };
typedef struct mpistub_file *MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_SOURCE (-1)

#define MPI_BYTE 1
#define MPI_INT 2
#define MPI_FLOAT 3
#define MPI_DOUBLE 4
#define MPI_UNSIGNED_CHAR 5
#define MPI_LONG_LONG 6

#define MPI_MIN 1
#define MPI_MAX 2
#define MPI_SUM 3

#define MPI_ORDER_C 0
#define MPI_MODE_WRONLY 1
#define MPI_MODE_CREATE 2
#define MPI_MODE_RDONLY 4

#define MPISTUB_MAX_TYPES 64
struct mpistub_subarray { int used, ndims, sizes[3], subsizes[3], starts[3], elem; };

static inline struct mpistub_subarray *mpistub_types(void)
{
  static struct mpistub_subarray t[MPISTUB_MAX_TYPES];
  return t;
}

static inline size_t mpistub_sizeof(MPI_Datatype t)
{
  switch (t) {
  case MPI_BYTE: case MPI_UNSIGNED_CHAR: return 1;
  case MPI_INT: case MPI_FLOAT: return 4;
  case MPI_DOUBLE: case MPI_LONG_LONG: return 8;
  default: fprintf(stderr, "mpi stub: unknown datatype %d\n", t); abort();
  }
}

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  (void)op; (void)c;
  memcpy(r, s, (size_t)n * mpistub_sizeof(t));
  return 0;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
  (void)op; (void)c; (void)root;
  memcpy(r, s, (size_t)n * mpistub_sizeof(t));
  return 0;
}
static inline int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{
  (void)nr; (void)tr; (void)c;
  memcpy(r, s, (size_t)ns * mpistub_sizeof(ts));
  return 0;
}
static inline int MPI_Allgatherv(const void *s, int ns, MPI_Datatype ts, void *r, const int *counts, const int *displs, MPI_Datatype tr, MPI_Comm c)
{
  (void)counts; (void)c;
  memcpy((char *)r + (size_t)displs[0] * mpistub_sizeof(tr), s, (size_t)ns * mpistub_sizeof(ts));
  return 0;
}
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "mpi stub: MPI_Abort(%d)\n", code); abort(); }
static inline int MPI_Error_string(int code, char *s, int *len)
{
  *len = snprintf(s, MPI_MAX_ERROR_STRING, "mpi stub error %d", code);
  return 0;
}
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *rq)
{
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; (void)rq;
  fprintf(stderr, "mpi stub: MPI_Isend is unreachable with one rank\n"); abort();
}
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st)
{
  (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)st;
  fprintf(stderr, "mpi stub: MPI_Recv is unreachable with one rank\n"); abort();
}
static inline int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *st) { (void)n; (void)rq; (void)st; return 0; }
static inline int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) { (void)t; *count = st->count; return 0; }

/* ---- derived datatypes (C-order subarrays of a basic type) ---- */
static inline int MPI_Type_create_subarray(int ndims, const int *sizes, const int *subsizes, const int *starts,
                                           int order, MPI_Datatype old, MPI_Datatype *newtype)
{
  struct mpistub_subarray *t = mpistub_types();
  (void)order;
  for (int i = 0; i < MPISTUB_MAX_TYPES; i++)
    if (!t[i].used) {
      t[i].used = 1; t[i].ndims = ndims; t[i].elem = (int)mpistub_sizeof(old);
      for (int d = 0; d < 3; d++) { t[i].sizes[d] = 1; t[i].subsizes[d] = 1; t[i].starts[d] = 0; }
      /* right-align so that dimension 2 is always the fastest one */
      for (int d = 0; d < ndims; d++) {
        t[i].sizes[3 - ndims + d] = sizes[d];
        t[i].subsizes[3 - ndims + d] = subsizes[d];
        t[i].starts[3 - ndims + d] = starts[d];
      }
      *newtype = 16 + i;
      return 0;
    }
  fprintf(stderr, "mpi stub: out of datatype slots\n"); abort();
}
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype *t) { if (*t >= 16) mpistub_types()[*t - 16].used = 0; *t = 0; return 0; }

/* ---- MPI-IO subset ---- */
static inline int MPI_File_open(MPI_Comm c, const char *name, int mode, MPI_Info info, MPI_File *fh)
{
  (void)c; (void)info;
  int flags = (mode & MPI_MODE_RDONLY) ? O_RDONLY : O_WRONLY;
  if (mode & MPI_MODE_CREATE) flags |= O_CREAT;
  int fd = open(name, flags, 0644);
  if (fd < 0) return 1;
  *fh = (MPI_File)malloc(sizeof(struct mpistub_file));
  (*fh)->fd = fd; (*fh)->view = MPI_BYTE;
  return 0;
}
static inline int MPI_File_set_size(MPI_File fh, MPI_Offset sz) { return ftruncate(fh->fd, sz) ? 1 : 0; }
static inline int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype,
                                    const char *rep, MPI_Info info)
{
  (void)disp; (void)etype; (void)rep; (void)info;
  fh->view = filetype;
  return 0;
}
static inline int MPI_File_write_all(MPI_File fh, const void *buf, int count, MPI_Datatype t, MPI_Status *st)
{
  size_t es = mpistub_sizeof(t);
  st->count = 0;
  if (count == 0) return 0;
  if (fh->view < 16) {
    if (pwrite(fh->fd, buf, es * (size_t)count, 0) != (ssize_t)(es * (size_t)count)) return 1;
    st->count = count;
    return 0;
  }
  const struct mpistub_subarray *v = &mpistub_types()[fh->view - 16];
  const char *src = (const char *)buf;
  int written = 0;
  for (int i = 0; i < v->subsizes[0] && written < count; i++)
    for (int j = 0; j < v->subsizes[1] && written < count; j++) {
      int n = v->subsizes[2];
      if (written + n > count) n = count - written;
      MPI_Offset off = (((MPI_Offset)(v->starts[0] + i) * v->sizes[1] + (v->starts[1] + j)) * v->sizes[2]
                        + v->starts[2]) * (MPI_Offset)es;
      if (pwrite(fh->fd, src + (size_t)written * es, es * (size_t)n, off) != (ssize_t)(es * (size_t)n)) return 1;
      written += n;
    }
  st->count = written;
  return 0;
}
static inline int MPI_File_close(MPI_File *fh) { close((*fh)->fd); free(*fh); *fh = 0; return 0; }

#ifdef __cplusplus
}
#endif
#endif
