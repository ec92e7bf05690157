// stand-in for Communicate/Communicator.h (oracle/ref_shim): rank/size are plain settable ints so
// FieldLayout::findNeighbors can be evaluated for every rank inside one process.
#pragma once
#include <cstddef>
#include <memory>
using MPI_Comm = int;
using MPI_Request = int;
#define MPI_STATUSES_IGNORE nullptr
inline int MPI_Waitall(int, MPI_Request*, void*) { return 0; }
constexpr MPI_Comm MPI_COMM_WORLD = 0;
namespace refshim { inline int g_rank = 0; inline int g_size = 1; }
namespace ippl { namespace mpi {
    namespace tag { constexpr int HALO = 20000; }
    class Communicator {
    public:
        Communicator(MPI_Comm = MPI_COMM_WORLD) {}
        int rank() const { return refshim::g_rank; }
        int size() const { return refshim::g_size; }
        // names HaloCells.hpp refers to outside dependent contexts (its exchange code is parsed, never run, here)
        template <class MemorySpace> using buffer_type = std::shared_ptr<int>;
        template <class MemorySpace, class T> buffer_type<MemorySpace> getBuffer(std::size_t) { return {}; }
        template <class Buffer, class Archive> void isend(int, int, Buffer&, Archive&, int&, std::size_t) {}
        template <class Buffer, class Archive> void recv(int, int, Buffer&, Archive&, std::size_t, std::size_t) {}
        void freeAllBuffers() {}
        double getDefaultOverallocation() const { return 1.0; }
        // one process plays the whole communicator and holds the global data: a sum over ranks is the identity
        template <typename T, class Op> void allreduce(const T* in, T* out, int n, Op) { for (int i = 0; i < n; ++i) out[i] = in[i]; }
    };
}}  // namespace ippl::mpi
