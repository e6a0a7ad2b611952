/*
 * oracle/mpi_stub/mpi.h -- TEST INFRASTRUCTURE.  Single-rank stand-in for <mpi.h> so that the
 * reference's metric headers (CBench/metrics/*.hpp) compile in this image, which has no MPI.
 * Every collective is the identity on one rank: Allreduce copies send -> recv.
 * Only the handful of symbols those headers touch are provided.
 */
#ifndef ZFB_ORACLE_MPI_STUB_H
#define ZFB_ORACLE_MPI_STUB_H
#include <cstring>
#include <cstddef>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0

/* datatype handles encode the element size in bytes */
#define MPI_DOUBLE 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_FLOAT 4
#define MPI_INT 4
#define MPI_LONG 8
#define MPI_INT8_T 1
#define MPI_INT16_T 2
#define MPI_INT32_T 4
#define MPI_INT64_T 8
#define MPI_UINT8_T 1
#define MPI_UINT16_T 2
#define MPI_UINT32_T 4
#define MPI_UINT64_T 8

#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3

static inline int MPI_Comm_size(MPI_Comm, int *size) { *size = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op, MPI_Comm)
{
  std::memcpy(recv, send, (size_t)count * (size_t)type);
  return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op, int, MPI_Comm)
{
  std::memcpy(recv, send, (size_t)count * (size_t)type);
  return MPI_SUCCESS;
}
#endif
