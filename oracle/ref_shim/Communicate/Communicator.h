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
        // one process plays the whole communicator and holds the global data: a sum over ranks is the identity
        template <typename T, class Op> void allreduce(const T* in, T* out, int n, Op) { for (int i = 0; i < n; ++i) out[i] = in[i]; }
    };
}}  // namespace ippl::mpi
