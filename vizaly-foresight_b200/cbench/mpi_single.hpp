// mpi_single.hpp -- one-rank stand-in for the few MPI calls the metric plugins make, used only when the
// plugins are built outside an MPI environment (this repository's driver and tests; the image has no MPI).
// With one rank every reduction is the identity.
#ifndef ZFB_MPI_SINGLE_HPP
#define ZFB_MPI_SINGLE_HPP
#include <cstddef>
#include <cstring>

typedef int MPI_Comm;
typedef size_t MPI_Datatype;  // element size in bytes
typedef int MPI_Op;
enum { MPI_COMM_WORLD = 0 };
enum { MPI_MAX = 1, MPI_MIN = 2, MPI_SUM = 3 };
static const MPI_Datatype MPI_DOUBLE = sizeof(double);
static const MPI_Datatype MPI_UNSIGNED_LONG_LONG = sizeof(unsigned long long);

inline int MPI_Comm_size(MPI_Comm, int* n) { *n = 1; return 0; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
inline int MPI_Barrier(MPI_Comm) { return 0; }
inline int MPI_Allreduce(const void* in, void* out, int count, MPI_Datatype elem, MPI_Op, MPI_Comm)
{
  std::memcpy(out, in, (size_t)count * elem);
  return 0;
}
#endif
