// stand-in for Communicate/Communicator.h (oracle/ref_shim): rank/size are plain settable ints so
// FieldLayout::findNeighbors can be evaluated for every rank inside one process.
#pragma once
using MPI_Comm = int;
constexpr MPI_Comm MPI_COMM_WORLD = 0;
namespace refshim { inline int g_rank = 0; inline int g_size = 1; }
namespace ippl { namespace mpi {
    class Communicator {
    public:
        Communicator(MPI_Comm = MPI_COMM_WORLD) {}
        int rank() const { return refshim::g_rank; }
        int size() const { return refshim::g_size; }
    };
}}  // namespace ippl::mpi
