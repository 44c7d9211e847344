// Single-rank MPI stand-in for the oracle build (TEST INFRASTRUCTURE ONLY).
//
// The reference (arashb/tbslas) needs <mpi.h>; this container has no MPI.  The
// oracle only ever runs the reference's np == 1 path, so every collective
// degenerates to a local copy.  Results of the hot path are partition
// invariant (each point is evaluated by the one leaf that contains it), so the
// np == 1 run also pins the multi-GPU results.
#ifndef TBSLAS_ORACLE_SHIM_MPI_H_
#define TBSLAS_ORACLE_SHIM_MPI_H_

#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;  // element size in bytes
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUCCESS 0

inline int MPI_Comm_rank(MPI_Comm, int *rank) {
  *rank = 0;
  return MPI_SUCCESS;
}
inline int MPI_Comm_size(MPI_Comm, int *size) {
  *size = 1;
  return MPI_SUCCESS;
}
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype stype,
                         void *rbuf, int, MPI_Datatype, MPI_Comm) {
  std::memcpy(rbuf, sbuf, (size_t)scount * (size_t)stype);
  return MPI_SUCCESS;
}
inline int MPI_Allreduce(const void *sbuf, void *rbuf, int count,
                         MPI_Datatype type, MPI_Op, MPI_Comm) {
  std::memcpy(rbuf, sbuf, (size_t)count * (size_t)type);
  return MPI_SUCCESS;
}

#endif  // TBSLAS_ORACLE_SHIM_MPI_H_
