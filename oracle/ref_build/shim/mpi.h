/* Build shim (test infrastructure): MPI is absent; only type names are needed by headers. */
#ifndef PF_SHIM_MPI_H
#define PF_SHIM_MPI_H
typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Win;
typedef int MPI_Datatype;
struct MPI_Status { int MPI_SOURCE, MPI_TAG; };
#define MPI_COMM_WORLD 0
#endif
