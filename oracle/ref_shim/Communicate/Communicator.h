// stand-in for Communicate/Communicator.h (oracle/ref_shim): rank/size are plain settable ints so
// FieldLayout::findNeighbors can be evaluated for every rank inside one process, and -- for HaloCells::exchangeBoundaries
// -- an in-process mailbox: every rank is a thread (g_rank is thread-local), isend copies the packed halo buffer into the
// mailbox of (source, destination, tag), recv blocks until that message is there.  Message order per (source,
// destination, tag) is FIFO, as MPI guarantees.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>
using MPI_Comm = int;
using MPI_Request = int;
#define MPI_STATUSES_IGNORE nullptr
inline int MPI_Waitall(int, MPI_Request*, void*) { return 0; }
constexpr MPI_Comm MPI_COMM_WORLD = 0;
namespace refshim {
    inline thread_local int g_rank = 0;
    inline int g_size              = 1;
    struct Mailbox {
        std::mutex m;
        std::condition_variable cv;
        std::map<std::tuple<int, int, int>, std::deque<std::vector<unsigned char>>> q;
    };
    inline Mailbox& mailbox() {
        static Mailbox b;
        return b;
    }
    struct ArchiveStub {
        void resetWritePos() {}
        void resetReadPos() {}
    };
}  // namespace refshim
namespace ippl { namespace mpi {
    namespace tag { constexpr int HALO = 20000; }
    class Communicator {
    public:
        Communicator(MPI_Comm = MPI_COMM_WORLD) {}
        int rank() const { return refshim::g_rank; }
        int size() const { return refshim::g_size; }
        template <class MemorySpace> using buffer_type = std::shared_ptr<refshim::ArchiveStub>;
        template <class MemorySpace, class T> buffer_type<MemorySpace> getBuffer(std::size_t) {
            return std::make_shared<refshim::ArchiveStub>();
        }
        // Communicator::isend / recv of a FieldBufferData (Communicate/Communicator.h:176-206): the first n elements of
        // fd.buffer travel
        template <class Buffer, class Archive> void isend(int dest, int tag, Buffer& fd, Archive&, MPI_Request&, std::size_t n) {
            using T = typename decltype(fd.buffer)::value_type;
            std::vector<unsigned char> bytes(n * sizeof(T));
            if (n) std::memcpy(bytes.data(), fd.buffer.data(), bytes.size());
            auto& mb = refshim::mailbox();
            {
                std::lock_guard<std::mutex> lk(mb.m);
                mb.q[{rank(), dest, tag}].push_back(std::move(bytes));
            }
            mb.cv.notify_all();
        }
        template <class Buffer, class Archive> void recv(int src, int tag, Buffer& fd, Archive&, std::size_t, std::size_t n) {
            using T = typename decltype(fd.buffer)::value_type;
            auto& mb = refshim::mailbox();
            std::vector<unsigned char> bytes;
            {
                std::unique_lock<std::mutex> lk(mb.m);
                auto key = std::make_tuple(src, rank(), tag);
                mb.cv.wait(lk, [&] { return !mb.q[key].empty(); });
                bytes = std::move(mb.q[key].front());
                mb.q[key].pop_front();
            }
            if (fd.buffer.size() < n) Kokkos::realloc(fd.buffer, n);
            if (n) std::memcpy(fd.buffer.data(), bytes.data(), n * sizeof(T));
        }
        void freeAllBuffers() {}
        double getDefaultOverallocation() const { return 1.0; }
        // one process plays the whole communicator and holds the global data: a sum over ranks is the identity
        template <typename T, class Op> void allreduce(const T* in, T* out, int n, Op) { for (int i = 0; i < n; ++i) out[i] = in[i]; }
    };
}}  // namespace ippl::mpi
