// Ippl.h -- host-side C++ mirror of IPPL's API for the particle-mesh hot path, on top of the C-ABI
// (include/ippl_b200.h).  Same names, argument meaning and error behaviour as the reference for the path
// SURVEY.md section 8 scopes (paths below are relative to the reference tree):
//
//   ippl::Vector                    src/Types/Vector.h
//   ippl::Index / NDIndex           src/Index/Index.h, src/Index/NDIndex.h
//   ippl::FieldLayout               src/FieldLayout/FieldLayout.h           (boxes from ipplb_layout_*)
//   ippl::UniformCartesian          src/Meshes/UniformCartesian.h
//   ippl::Field (BareField)         src/Field/BareField.h / .hpp            (= scalar, sum, accumulateHalo, fillHalo)
//   ippl::ParticleAttrib            src/Particle/ParticleAttrib.h / .hpp    (scatter, gather, = expression)
//   ippl::ParticleSpatialLayout     src/Particle/ParticleSpatialLayout.h    (update: BC + migration)
//   ippl::ParticleBase              src/Particle/ParticleBase.h             (create, addAttribute, update)
//   ippl::scatter / ippl::gather    src/Particle/ParticleAttrib.hpp:304-358
//   ippl::FFTPeriodicPoissonSolver  src/PoissonSolvers/FFTPeriodicPoissonSolver.h (non-owned stage, cuFFT)
//   IpplTimings, Inform, IpplException, ippl::Comm, ippl::initialize / finalize
//
// Differences, all behind the same API: particle vectors are stored SoA on the device (the reference stores
// AoS Kokkos::View<Vector<T,3>*>); host access goes through getHostMirror() + ippl::deep_copy like the
// reference's create_mirror_view / deep_copy; there is no Kokkos, so driver-side KOKKOS_LAMBDA kernels are out
// of this header's scope.  Header-only, C++17, links libippl_b200.so + libcudart.  No CPU fallback: every
// operation ends in an ipplb_* call and throws IpplException when that fails.
#ifndef IPPL_B200_FACADE_H
#define IPPL_B200_FACADE_H

#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <variant>
#include <vector>

#include <unistd.h>

#include "../ippl_b200.h"

// functions that driver-side device lambdas may call (include/ippl/KokkosShim.cuh, nvcc): host + device under nvcc,
// plain host functions for the host compiler
#ifdef __CUDACC__
#define IPPL_HD __host__ __device__
#else
#define IPPL_HD
#endif

// ---- errors (src/Utility/IpplException.h) ---------------------------------------------------------------------
class IpplException : public std::runtime_error {
public:
    IpplException(const std::string& meth, const std::string& descr)
        : std::runtime_error(meth + ": " + descr), meth_(meth), descr_(descr) {}
    const std::string& where() const { return meth_; }
    const std::string& what_str() const { return descr_; }

private:
    std::string meth_, descr_;
};

// ---- Inform (src/Utility/Inform.h): message stream of one rank (default: rank 0) to stdout or to a file -----------------
constexpr int INFORM_ALL_NODES = -1;
class Inform {
public:
    enum WriteMode { OVERWRITE, APPEND };
    explicit Inform(const char* name = nullptr, int printNode = 0) : name_(name ? name : ""), node_(printNode) {}
    // Inform(name, file, mode, node): the drivers' CSV dumps (e.g. LandauDampingManager.h:377)
    Inform(const char* name, const char* fname, WriteMode mode, int printNode = 0) : name_(name ? name : ""), node_(printNode) {
        if (node_ == INFORM_ALL_NODES || rank_ref() == node_)
            file_ = std::make_unique<std::ofstream>(fname, mode == APPEND ? std::ios::app : std::ios::trunc);
        to_file_ = true;
    }
    template <typename T>
    Inform& operator<<(const T& v) {
        buf_ << v;
        return *this;
    }
    Inform& operator<<(Inform& (*f)(Inform&)) { return f(*this); }
    std::streamsize precision(std::streamsize p) { return buf_.precision(p); }
    std::ios::fmtflags setf(std::ios::fmtflags f, std::ios::fmtflags mask) { return buf_.setf(f, mask); }
    // endl: the message goes out (Inform.cpp:42-45, outputMessage); flush(): only the destination stream is flushed
    // (Inform.h:113) -- `csvout << endl; csvout.flush();` in the drivers' dumps writes ONE line
    Inform& outputMessage() {
        const bool mine = node_ == INFORM_ALL_NODES || rank_ref() == node_;
        if (to_file_) {
            if (file_) *file_ << buf_.str() << std::endl;
        } else if (level_on() && mine) {
            std::cout << (name_.empty() ? "" : name_ + "> ") << buf_.str() << std::endl;
        }
        buf_.str("");
        return *this;
    }
    void flush() {
        if (to_file_) {
            if (file_) file_->flush();
        } else {
            std::cout.flush();
        }
    }
    static bool& level_on() {
        static bool on = true;
        return on;
    }
    static int& rank_ref() {  // set by ippl::initialize
        static int r = 0;
        return r;
    }

private:
    std::string name_;
    int node_ = 0;
    bool to_file_ = false;
    std::unique_ptr<std::ofstream> file_;
    std::ostringstream buf_;
};
inline Inform& endl(Inform& m) { return m.outputMessage(); }

namespace ippl {

namespace detail {
    using size_type = std::size_t;
}

// ---- runtime: one context per process/GPU (src/Ippl.cpp, src/Communicate/Communicator.h) -------------------------
namespace b200 {
    inline ipplb_ctx*& ctx_ref() {
        static ipplb_ctx* c = nullptr;
        return c;
    }
    inline ipplb_ctx* ctx() {
        if (!ctx_ref()) throw IpplException("ippl::b200::ctx", "ippl::initialize has not been called");
        return ctx_ref();
    }
    inline std::string& id_file() {
        static std::string f;
        return f;
    }
    inline void check(int rc, const char* where) {
        if (rc != IPPLB_OK) throw IpplException(where, ipplb_last_error());
    }
    inline void cuda_check(cudaError_t e, const char* where) {
        if (e != cudaSuccess) throw IpplException(where, cudaGetErrorString(e));
    }
    template <typename T>
    T* device_alloc(std::size_t n) {
        T* p = nullptr;
        cuda_check(cudaMalloc(&p, sizeof(T) * std::max<std::size_t>(n, 1)), "cudaMalloc");
        return p;
    }
    // IPPL_B200_FUSE=1 (or fusion_enabled() = true before the particles are created): the leapfrog expression sequence of an
    // unchanged driver is recognised lazily and executed by the fused single-pass step (detail::FusionEngine below).
    inline bool& fusion_enabled() {
        static bool on = [] {
            const char* e = std::getenv("IPPL_B200_FUSE");
            return e && std::atoi(e) != 0;
        }();
        return on;
    }
    struct FusionStats {
        long fused_steps = 0, materialised = 0;
    };
    inline FusionStats& fusion_stats() {
        static FusionStats st;
        return st;
    }
}  // namespace b200

class Communicator {
public:
    int rank() const { return rank_; }
    int size() const { return size_; }
    int getCommunicator() const { return 0; }   // the MPI_Comm stand-in of include/ippl/compat (one process per GPU, NCCL underneath)
    // Communicator::barrier (src/Communicate/Communicator.h): drains this rank's stream, then a rank barrier (one small
    // all-reduce over NCCL, host-synchronous like MPI_Barrier)
    void barrier() {
        b200::check(ipplb_sync(b200::ctx()), "Comm::barrier");
        if (size_ > 1) {
            long one = 1;
            b200::check(ipplb_allreduce_sum_i64(b200::ctx(), &one), "Comm::barrier");
        }
    }
    [[noreturn]] void abort() {
        std::cerr << "ippl::Comm->abort()" << std::endl;
        std::abort();
    }
    double getDefaultOverallocation() const { return overalloc_; }
    void setDefaultOverallocation(double f) { overalloc_ = f; }
    // reduce / allreduce of one value over ranks (single rank: identity; multi rank: NCCL through the C-ABI)
    // std::plus -> sum, std::greater -> max (the two operations the alpine drivers reduce with)
    template <typename T, typename Op>
    void reduce(const T& in, T& out, int /*count*/, Op, int /*root*/ = 0) {
        out = in;
        if (size_ > 1) {
            if constexpr (std::is_same_v<Op, std::greater<T>>) allreduce_max(out);
            else allreduce_sum(out);
        }
    }
    template <typename T, typename Op>
    void allreduce(T& inout, int /*count*/, Op) {
        if (size_ > 1) {
            if constexpr (std::is_same_v<Op, std::greater<T>>) allreduce_max(inout);
            else allreduce_sum(inout);
        }
    }
    void set(int rank, int size) {
        rank_ = rank;
        size_ = size;
    }

private:
    void allreduce_sum(double& v) { b200::check(ipplb_allreduce_sum_f64(b200::ctx(), &v), "Comm::allreduce"); }
    void allreduce_max(double& v) { b200::check(ipplb_allreduce_max_f64(b200::ctx(), &v), "Comm::allreduce"); }
    void allreduce_max(std::size_t&) { throw IpplException("Comm::allreduce", "max over ranks is wired for double only"); }
    void allreduce_sum(std::size_t& v) {
        long t = (long)v;
        b200::check(ipplb_allreduce_sum_i64(b200::ctx(), &t), "Comm::allreduce");
        v = (std::size_t)t;
    }
    int rank_ = 0, size_ = 1;
    double overalloc_ = 1.0;
};
inline std::unique_ptr<Communicator>& comm_holder() {
    static std::unique_ptr<Communicator> c;
    return c;
}
inline Communicator* Comm = nullptr;

inline void initialize(int& argc, char**& argv) {
    int device = 0;
    if (const char* lr = std::getenv("LOCAL_RANK")) device = std::atoi(lr);
    b200::cuda_check(cudaSetDevice(device), "ippl::initialize");
    // the CUDA default stream: the facade's blocking cudaMemcpy calls and the kernels stay ordered
    b200::check(ipplb_ctx_create(&b200::ctx_ref(), device, nullptr, 0), "ippl::initialize");
    comm_holder() = std::make_unique<Communicator>();
    Comm          = comm_holder().get();
    // one process per GPU, launched like `torchrun --no-python --nproc-per-node N <driver> ...` (RANK / WORLD_SIZE /
    // LOCAL_RANK in the environment).  The NCCL id is published by rank 0 through a file named after the launcher's pid
    // (the reference uses MPI_Init + MPI_COMM_WORLD here, src/Ippl.cpp:24-33; MPI is not part of this build).
    int rank = 0, size = 1;
    if (const char* e = std::getenv("RANK")) rank = std::atoi(e);
    if (const char* e = std::getenv("WORLD_SIZE")) size = std::atoi(e);
    if (size > 1) {
        std::string path = "/tmp/ipplb_nccl_id_" + std::to_string((long)getppid());
        if (const char* e = std::getenv("IPPLB_NCCL_ID_FILE")) path = e;
        char id[IPPLB_NCCL_ID_BYTES];
        if (rank == 0) {
            b200::check(ipplb_nccl_unique_id(id), "ippl::initialize");
            const std::string tmp = path + ".tmp";
            FILE* f               = std::fopen(tmp.c_str(), "wb");
            if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) throw IpplException("ippl::initialize", "cannot publish the NCCL id");
            std::fclose(f);
            std::rename(tmp.c_str(), path.c_str());
        } else {
            bool got = false;
            for (int tries = 0; tries < 6000 && !got; ++tries) {  // up to 60 s
                if (FILE* f = std::fopen(path.c_str(), "rb")) {
                    got = std::fread(id, 1, sizeof(id), f) == sizeof(id);
                    std::fclose(f);
                }
                if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
            if (!got) throw IpplException("ippl::initialize", "timed out waiting for the NCCL id of rank 0");
        }
        b200::check(ipplb_comm_init(b200::ctx_ref(), rank, size, id), "ippl::initialize");
        Comm->set(rank, size);
        b200::id_file() = rank == 0 ? path : "";
    }
    Inform::rank_ref() = rank;
    for (int i = 1; i < argc; ++i) {  // global flags of src/Ippl.cpp:35-92 that matter here
        const std::string a = argv[i];
        if (a == "--overallocate" && i + 1 < argc) Comm->setDefaultOverallocation(std::atof(argv[i + 1]));
        if (a == "--info" && i + 1 < argc) Inform::level_on() = std::atoi(argv[i + 1]) > 0;
    }
}
inline void finalize() {
    if (b200::fusion_enabled() && Inform::rank_ref() == 0)
        std::cout << "ippl_b200 fusion: " << b200::fusion_stats().fused_steps << " fused steps, " << b200::fusion_stats().materialised
                  << " materialisations" << std::endl;
    if (b200::ctx_ref()) ipplb_ctx_destroy(b200::ctx_ref());
    b200::ctx_ref() = nullptr;
    if (!b200::id_file().empty()) std::remove(b200::id_file().c_str());
}
inline void fence() { b200::check(ipplb_sync(b200::ctx()), "ippl::fence"); }

// ---- Vector (src/Types/Vector.h) -----------------------------------------------------------------------------------
template <typename T, unsigned Dim>
class Vector {
public:
    IPPL_HD Vector() {
        for (unsigned i = 0; i < Dim; ++i) data_[i] = T();
    }
    IPPL_HD Vector(const T& v) {
        for (unsigned i = 0; i < Dim; ++i) data_[i] = v;
    }
    IPPL_HD Vector(std::initializer_list<T> l) {
        unsigned i = 0;
        for (const T* p = l.begin(); p != l.end() && i < Dim; ++p, ++i) data_[i] = *p;
        for (; i < Dim; ++i) data_[i] = T();
    }
    IPPL_HD T& operator[](unsigned i) { return data_[i]; }
    IPPL_HD const T& operator[](unsigned i) const { return data_[i]; }
    IPPL_HD T& operator()(unsigned i) { return data_[i]; }   // element access with call syntax (src/Types/Vector.h)
    IPPL_HD const T& operator()(unsigned i) const { return data_[i]; }
    IPPL_HD T* begin() { return data_; }
    IPPL_HD T* end() { return data_ + Dim; }
    IPPL_HD const T* begin() const { return data_; }
    IPPL_HD const T* end() const { return data_ + Dim; }
    // converting copy (Vector<long> -> Vector<double> etc.)
    template <typename U>
    IPPL_HD Vector(const Vector<U, Dim>& o) {
        for (unsigned i = 0; i < Dim; ++i) data_[i] = static_cast<T>(o[i]);
    }
#define IPPL_VEC_OP(op)                                                           \
    template <typename U>                                                         \
    IPPL_HD Vector& operator op##=(const Vector<U, Dim>& o) {                     \
        for (unsigned i = 0; i < Dim; ++i) data_[i] op## = o[i];                  \
        return *this;                                                             \
    }                                                                             \
    IPPL_HD Vector& operator op##=(const T& s) {                                  \
        for (unsigned i = 0; i < Dim; ++i) data_[i] op## = s;                     \
        return *this;                                                             \
    }
    IPPL_VEC_OP(+)
    IPPL_VEC_OP(-)
    IPPL_VEC_OP(*)
    IPPL_VEC_OP(/)
#undef IPPL_VEC_OP
    friend std::ostream& operator<<(std::ostream& os, const Vector& v) {
        os << "( ";
        for (unsigned i = 0; i < Dim; ++i) os << v[i] << (i + 1 < Dim ? " , " : " )");
        return os;
    }

private:
    T data_[Dim];
};

// element-wise arithmetic with the usual promotions (Vector<long> + 0.5 is a Vector<double>, Vector<double> / Vector<int> too):
// the drivers mix index vectors and coordinates (LandauDampingManager.h:90-92, :193-194)
#define IPPL_VEC_BIN(op)                                                                                          \
    template <typename T, typename U, unsigned Dim>                                                               \
    IPPL_HD Vector<std::common_type_t<T, U>, Dim> operator op(const Vector<T, Dim>& a, const Vector<U, Dim>& b) {  \
        Vector<std::common_type_t<T, U>, Dim> r;                                                                   \
        for (unsigned i = 0; i < Dim; ++i) r[i] = a[i] op b[i];                                                    \
        return r;                                                                                                  \
    }                                                                                                              \
    template <typename T, typename U, unsigned Dim, std::enable_if_t<std::is_arithmetic<U>::value, int> = 0>       \
    IPPL_HD Vector<std::common_type_t<T, U>, Dim> operator op(const Vector<T, Dim>& a, const U& s) {               \
        Vector<std::common_type_t<T, U>, Dim> r;                                                                   \
        for (unsigned i = 0; i < Dim; ++i) r[i] = a[i] op s;                                                       \
        return r;                                                                                                  \
    }                                                                                                              \
    template <typename T, typename U, unsigned Dim, std::enable_if_t<std::is_arithmetic<U>::value, int> = 0>       \
    IPPL_HD Vector<std::common_type_t<T, U>, Dim> operator op(const U& s, const Vector<T, Dim>& a) {               \
        Vector<std::common_type_t<T, U>, Dim> r;                                                                   \
        for (unsigned i = 0; i < Dim; ++i) r[i] = s op a[i];                                                       \
        return r;                                                                                                  \
    }
IPPL_VEC_BIN(+)
IPPL_VEC_BIN(-)
IPPL_VEC_BIN(*)
IPPL_VEC_BIN(/)
#undef IPPL_VEC_BIN

// ---- views handed to driver-side device lambdas (Kokkos::View stand-ins; see include/ippl/KokkosShim.cuh) -----------------------
namespace detail {
    // view(i)[d] on SoA component arrays: what ParticleAttrib<Vector<T, 3>>::getView() returns in the reference
    // (Kokkos::View<Vector<T, 3>*>, AoS) seen through a proxy
    struct SoARef3 {
        double *x, *y, *z;
        IPPL_HD double& operator[](unsigned d) const { return d == 0 ? *x : (d == 1 ? *y : *z); }
        IPPL_HD operator Vector<double, 3>() const {
            Vector<double, 3> v;
            v[0] = *x; v[1] = *y; v[2] = *z;
            return v;
        }
        IPPL_HD const SoARef3& operator=(const Vector<double, 3>& v) const {
            *x = v[0]; *y = v[1]; *z = v[2];
            return *this;
        }
    };
    template <int NC>
    struct AttribView;
    template <>
    struct AttribView<3> {
        double* c[3];
        std::size_t n;
        IPPL_HD SoARef3 operator()(std::size_t i) const { return SoARef3{c[0] + i, c[1] + i, c[2] + i}; }
        IPPL_HD std::size_t extent(int) const { return n; }
        IPPL_HD std::size_t size() const { return n; }
    };
    template <>
    struct AttribView<1> {
        double* c[1];
        std::size_t n;
        IPPL_HD double& operator()(std::size_t i) const { return c[0][i]; }
        IPPL_HD std::size_t extent(int) const { return n; }
        IPPL_HD std::size_t size() const { return n; }
    };
    // ghosted field view, x fastest: view(i, j, k) -> double& (scalar field) or a Vector-like reference (AoS-3 field)
    struct AoSRef3 {
        double* p;
        IPPL_HD double& operator[](unsigned d) const { return p[d]; }
        IPPL_HD operator Vector<double, 3>() const {
            Vector<double, 3> v;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
            return v;
        }
        IPPL_HD const AoSRef3& operator=(const Vector<double, 3>& v) const {
            p[0] = v[0]; p[1] = v[1]; p[2] = v[2];
            return *this;
        }
    };
    template <int NC>
    struct FieldView;
    template <>
    struct FieldView<1> {
        double* d;
        int e[3];  // ghosted extents
        IPPL_HD double& operator()(long i, long j, long k) const { return d[i + (long)e[0] * (j + (long)e[1] * k)]; }
        IPPL_HD int extent(int a) const { return e[a]; }
    };
    template <>
    struct FieldView<3> {
        double* d;
        int e[3];
        IPPL_HD AoSRef3 operator()(long i, long j, long k) const { return AoSRef3{d + 3 * (i + (long)e[0] * (j + (long)e[1] * k))}; }
        IPPL_HD int extent(int a) const { return e[a]; }
    };
}  // namespace detail

// ---- Index / NDIndex (src/Index) -------------------------------------------------------------------------------------
class Index {
public:
    IPPL_HD Index() : first_(0), length_(0) {}
    IPPL_HD explicit Index(int n) : first_(0), length_(n) {}
    IPPL_HD Index(int first, int last) : first_(first), length_(last - first + 1) {}
    IPPL_HD int first() const { return first_; }
    IPPL_HD int last() const { return first_ + length_ - 1; }
    IPPL_HD int length() const { return length_; }

private:
    int first_, length_;
};
template <unsigned Dim>
class NDIndex {
public:
    IPPL_HD NDIndex() {}
    template <typename... Idx>
    IPPL_HD NDIndex(const Idx&... i) : idx_{i...} {}
    IPPL_HD Index& operator[](unsigned d) { return idx_[d]; }
    IPPL_HD const Index& operator[](unsigned d) const { return idx_[d]; }
    IPPL_HD std::size_t size() const {
        std::size_t s = 1;
        for (unsigned d = 0; d < Dim; ++d) s *= idx_[d].length();
        return s;
    }
    // first / last / length of every dimension as index vectors (src/Index/NDIndex.hpp)
    IPPL_HD Vector<int, Dim> first() const {
        Vector<int, Dim> v;
        for (unsigned d = 0; d < Dim; ++d) v[d] = idx_[d].first();
        return v;
    }
    IPPL_HD Vector<int, Dim> last() const {
        Vector<int, Dim> v;
        for (unsigned d = 0; d < Dim; ++d) v[d] = idx_[d].last();
        return v;
    }
    IPPL_HD Vector<int, Dim> length() const {
        Vector<int, Dim> v;
        for (unsigned d = 0; d < Dim; ++d) v[d] = idx_[d].length();
        return v;
    }

private:
    Index idx_[Dim];
};

template <unsigned Dim>
std::ostream& operator<<(std::ostream& os, const NDIndex<Dim>& n) {
    os << "{";
    for (unsigned d = 0; d < Dim; ++d) os << "[" << n[d].first() << ":" << n[d].last() << ":1]" << (d + 1 < Dim ? "," : "");
    return os << "}";
}

// ippl::RangePolicy<Dim> (src/Utility/ParallelDispatch.h): the index box of a field kernel; the device side of
// ippl::parallel_for / parallel_reduce over it is in include/ippl/KokkosShim.cuh
template <unsigned Dim>
struct RangePolicy {
    static_assert(Dim == 3, "the B200 path is three-dimensional");
    using index_type       = long;
    using index_array_type = Vector<long, Dim>;
    struct policy_type {
        long lo[3], hi[3];   // [lo, hi) in ghosted local indices
        long count() const { return (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]); }
    };
};

enum BC { PERIODIC, REFLECTIVE, SINK, NO };

// ---- FieldLayout (src/FieldLayout/FieldLayout.h) ----------------------------------------------------------------------
template <unsigned Dim>
class FieldLayout {
    static_assert(Dim == 3, "the B200 path is three-dimensional");

public:
    FieldLayout() = default;
    // FieldLayout(comm, domain, isParallel, isAllPeriodic): FieldLayout.hpp:76-134 (comm is implicit: ippl::Comm)
    template <typename CommT>
    FieldLayout(const CommT&, const NDIndex<Dim>& domain, std::array<bool, Dim> decomp, bool isAllPeriodic = false) {
        initialize(domain, decomp, isAllPeriodic);
    }
    ~FieldLayout() {
        if (h_) ipplb_layout_destroy(h_);
    }
    FieldLayout(const FieldLayout&)            = delete;
    FieldLayout& operator=(const FieldLayout&) = delete;
    void initialize(const NDIndex<Dim>& domain, std::array<bool, Dim> decomp, bool isAllPeriodic, int nghost = 1) {
        domain_   = domain;
        periodic_ = isAllPeriodic;
        int ng[3], par[3];
        for (unsigned d = 0; d < Dim; ++d) {
            ng[d]  = domain[d].length();
            par[d] = decomp[d];
        }
        b200::check(ipplb_layout_create(&h_, ng, par, Comm->size(), isAllPeriodic, nghost), "FieldLayout::initialize");
        std::vector<int> boxes(6 * Comm->size());
        b200::check(ipplb_layout_boxes(h_, boxes.data()), "FieldLayout::initialize");
        const int* b = &boxes[6 * Comm->rank()];
        local_       = NDIndex<Dim>(Index(b[0], b[3]), Index(b[1], b[4]), Index(b[2], b[5]));
    }
    // FieldLayout::updateLayout(domains), FieldLayout.hpp:136-160: the rank boxes of an ORB repartition
    void updateLayout(const std::vector<NDIndex<Dim>>& domains) {
        std::vector<int> boxes(6 * domains.size());
        for (std::size_t r = 0; r < domains.size(); ++r)
            for (unsigned d = 0; d < Dim; ++d) {
                boxes[6 * r + d]     = domains[r][d].first();
                boxes[6 * r + 3 + d] = domains[r][d].last();
            }
        b200::check(ipplb_layout_set_boxes(h_, boxes.data()), "FieldLayout::updateLayout");
        local_ = domains[Comm->rank()];
    }
    const NDIndex<Dim>& getDomain() const { return domain_; }
    const NDIndex<Dim>& getLocalNDIndex() const { return local_; }
    bool isAllPeriodic() const { return periodic_; }
    ipplb_layout* handle() const { return h_; }

private:
    ipplb_layout* h_ = nullptr;
    NDIndex<Dim> domain_, local_;
    bool periodic_ = false;
};

// ---- UniformCartesian (src/Meshes/UniformCartesian.h) ------------------------------------------------------------------
struct Cell {};
template <typename T, unsigned Dim>
class UniformCartesian {
public:
    using DefaultCentering = Cell;
    using vector_type      = Vector<T, Dim>;
    UniformCartesian() = default;
    UniformCartesian(const NDIndex<Dim>& domain, const vector_type& hx, const vector_type& origin) {
        initialize(domain, hx, origin);
    }
    void initialize(const NDIndex<Dim>& domain, const vector_type& hx, const vector_type& origin) {
        domain_ = domain;
        hx_     = hx;
        origin_ = origin;
    }
    const vector_type& getMeshSpacing() const { return hx_; }
    const vector_type& getOrigin() const { return origin_; }
    T getCellVolume() const { return std::accumulate(hx_.begin(), hx_.end(), T(1), std::multiplies<T>()); }
    const NDIndex<Dim>& getGridsize() const { return domain_; }

private:
    NDIndex<Dim> domain_;
    vector_type hx_, origin_;
};

namespace detail {
    template <typename T>
    struct ncomp_of {
        static constexpr int value = 1;
    };
    template <typename T, unsigned D>
    struct ncomp_of<Vector<T, D>> {
        static constexpr int value = (int)D;
    };
}  // namespace detail

namespace detail {
    // field / scalar and field - scalar: the two expressions AlpineManager::getDensity assigns (AlpineManager.h:225-245)
    template <class F>
    struct FieldAffine {
        const F* f;
        double divisor, shift;
    };
}  // namespace detail

// ---- Field / BareField (src/Field/BareField.h, BareField.hpp) --------------------------------------------------------------
template <typename T, unsigned Dim, class Mesh = UniformCartesian<double, Dim>, class Centering = Cell>
class Field {
    static_assert(Dim == 3, "the B200 path is three-dimensional");

public:
    using Mesh_t   = Mesh;
    using Layout_t = FieldLayout<Dim>;
    static constexpr int ncomp = detail::ncomp_of<T>::value;
    Field() = default;
    Field(Mesh& m, Layout_t& l, int nghost = 1) { initialize(m, l, nghost); }
    ~Field() {
        if (data_) cudaFree(data_);
    }
    Field(const Field&)            = delete;
    Field& operator=(const Field&) = delete;
    void initialize(Mesh& m, Layout_t& l, int nghost = 1) {
        mesh_p_   = &m;
        layout_p_ = &l;
        nghost_   = nghost;
        double o[3], h[3];
        for (int d = 0; d < 3; ++d) {
            o[d] = m.getOrigin()[d];
            h[d] = m.getMeshSpacing()[d];
        }
        b200::check(ipplb_layout_mesh(l.handle(), Comm->rank(), o, h, &mesh_), "Field::initialize");
        cells_ = 1;
        for (int d = 0; d < 3; ++d) cells_ *= mesh_.nl[d] + 2 * nghost;
        if (data_) cudaFree(data_);
        data_ = b200::device_alloc<double>(cells_ * ncomp);
        *this = 0.0;
    }
    // BareField::updateLayout, BareField.hpp:124-129: storage follows the new local box (contents are not kept)
    void updateLayout(Layout_t& l, int nghost = 1) { initialize(*mesh_p_, l, nghost); }
    // field = field of the same layout (ORB keeps a copy of rho: OrthogonalRecursiveBisection.hpp:10-12)
    void copyFrom(const Field& o) {
        if (o.cells_ != cells_) throw IpplException("Field::copyFrom", "layouts differ");
        b200::cuda_check(cudaMemcpy(data_, o.data_, sizeof(double) * cells_ * ncomp, cudaMemcpyDeviceToDevice), "Field::copyFrom");
    }
    // field = scalar (BareField.hpp:182-185)
    Field& operator=(double v) {
        b200::check(ipplb_field_fill(b200::ctx(), data_, (long)(cells_ * ncomp), v), "Field::operator=");
        return *this;
    }
    // field = expression object that knows how to evaluate itself into a field (include/ippl/compat: dot(E, E))
    template <class Expr, class = decltype(std::declval<const Expr&>().assign_to(std::declval<Field&>()))>
    Field& operator=(const Expr& e) {
        e.assign_to(*this);
        return *this;
    }
    // field = field / c  and  field = field - s on the interior (expression assignment, BareField.hpp:187-205)
    Field& operator=(const detail::FieldAffine<Field>& e) {
        if (e.f != this) throw IpplException("Field::operator=", "only f = f / c and f = f - s are supported by the facade");
        b200::check(ipplb_field_density(b200::ctx(), &mesh_, data_, e.divisor, e.shift), "Field::operator=");
        return *this;
    }
    friend detail::FieldAffine<Field> operator/(const Field& f, double c) { return {&f, c, 0.0}; }
    friend detail::FieldAffine<Field> operator-(const Field& f, double s) { return {&f, 1.0, s}; }
    // rho.sum(): interior cells, reduced over ranks (BareField.hpp:224-240)
    double sum() const {
        double s = 0;
        b200::check(ipplb_field_sum(b200::ctx(), &mesh_, data_, &s), "Field::sum");
        double g = s;
        Comm->allreduce(g, 1, std::plus<double>());
        return g;
    }
    // BareField::accumulateHalo / fillHalo (BareField.hpp:152-172): exchange, then in-rank periodic wrap
    void accumulateHalo() { halo(1); }
    void fillHalo() { halo(0); }
    int getNghost() const { return nghost_; }
    // field boundary conditions: the drivers set an all-periodic set for the potential of the CG / FEM solvers; the
    // periodic FFT path has the periodicity in the FieldLayout, so the object is only accepted
    template <class B>
    void setFieldBC(const B&) {}
    int getFieldBC() const { return 0; }
    // getFieldRangePolicy(nghost): the field's index box without its ghost layers (BareField.hpp)
    typename RangePolicy<Dim>::policy_type getFieldRangePolicy(int nghost = -1) const {
        typename RangePolicy<Dim>::policy_type p;
        const int g = nghost < 0 ? nghost_ : nghost;
        for (int d = 0; d < 3; ++d) {
            p.lo[d] = g;
            p.hi[d] = mesh_.nl[d] + 2 * nghost_ - g;
        }
        return p;
    }
    // getView(): ghosted local array for driver-side kernels, view(i, j, k) with ghosted local indices (BareField.h)
    using view_type = detail::FieldView<ncomp>;
    view_type getView() const {
        view_type v;
        v.d = data_;
        for (int d = 0; d < 3; ++d) v.e[d] = mesh_.nl[d] + 2 * nghost_;
        return v;
    }
    Mesh& get_mesh() const { return *mesh_p_; }
    Layout_t& getLayout() const { return *layout_p_; }
    const ipplb_mesh& b200_mesh() const { return mesh_; }
    double* data() const { return data_; }
    std::size_t cells() const { return cells_; }
    // host copy of the ghosted storage (x fastest), ncomp values per cell -- Kokkos::create_mirror_view_and_copy
    std::vector<double> getHostMirror() const {
        std::vector<double> h(cells_ * ncomp);
        fence();
        b200::cuda_check(cudaMemcpy(h.data(), data_, sizeof(double) * h.size(), cudaMemcpyDeviceToHost), "Field::getHostMirror");
        return h;
    }

private:
    void halo(int accumulate) {
        if (Comm->size() > 1) {
            b200::check(ipplb_halo_exchange(b200::ctx(), data_, ncomp, accumulate), "Field::halo");
        } else {
            int mask = 0;
            for (int d = 0; d < 3; ++d)
                if (mesh_.nl[d] == mesh_.ng[d] && layout_p_->isAllPeriodic()) mask |= 1 << d;
            b200::check(accumulate ? ipplb_halo_accumulate_periodic(b200::ctx(), &mesh_, data_, ncomp, mask)
                                   : ipplb_halo_fill_periodic(b200::ctx(), &mesh_, data_, ncomp, mask),
                        "Field::halo");
        }
    }
    Mesh* mesh_p_       = nullptr;
    Layout_t* layout_p_ = nullptr;
    ipplb_mesh mesh_{};
    int nghost_        = 1;
    std::size_t cells_ = 0;
    double* data_      = nullptr;
};

// ---- ParticleAttrib (src/Particle/ParticleAttrib.h / .hpp) ----------------------------------------------------------------------
namespace detail {
    class ParticleAttribBase {
    public:
        virtual ~ParticleAttribBase()                 = default;
        virtual void create(std::size_t n)            = 0;
        virtual void setCount(std::size_t n)          = 0;
        // what the multi-rank exchange needs to know about an attribute without its type
        virtual int components() const                = 0;
        virtual double* component_ptr(int c) const    = 0;
        virtual std::size_t capacity() const          = 0;
        virtual void reserve_storage(std::size_t n)   = 0;
        // a scalar attribute whose every element holds one known value (attrib = scalar): what the fused step needs of q
        virtual bool uniform(double* /*value*/) const { return false; }
        virtual void restore_uniform(double /*value*/) {}
        void set_name(const std::string& n) { name_ = n; }
        const std::string& get_name() const { return name_; }
        void set_engine(class FusionEngine* e) { engine_ = e; }
        class FusionEngine* engine() const { return engine_; }

    protected:
        std::string name_;
        class FusionEngine* engine_ = nullptr;   // the container's lazy-fusion engine (set by ParticleBase)
    };

    // Lazy fusion of the alpine leapfrog sequence onto ipplb_bins_step, for drivers that are NOT changed.  An unchanged
    // driver spells one step as separate attribute expressions and calls (LandauDampingManager.h:265-320):
    //     gather(E_p, E, R);  P = P - 0.5 dt E_p;  |  P = P - 0.5 dt E_p;  R = R + dt P;  pc->update();  scatter(q, rho, R);
    // (the bar is the step boundary).  With fusion on, none of the first five does any work: each is RECORDED.  When the
    // scatter arrives and the record is exactly [gather, kick (, kick), drift, periodic BC] with matching coefficients and a
    // uniform charge, ONE fused kernel does all of it on the bucketed particle store (where the particles then stay).  Any
    // other access to the container's attributes in between -- getView(), a driver lambda, a dump that reads P, create(),
    // an expression that does not fit -- MATERIALISES first: the particles go back to the contiguous attribute arrays
    // (ipplb_bins_compact) and the recorded operations run one by one through the ordinary kernels, in order, so the
    // driver always sees what the unfused sequence would have produced (per particle bit for bit; rho to summation
    // order).  One thing cannot be reproduced: E_p AFTER a fused step (the gather it belongs to was consumed inside the
    // kernel); reading it then throws.  Several ranks: the step also writes the leavers out and ipplb_bins_migrate exchanges them
    // (what the restated drivers' --fused path does, demos/Alpine.h).
    class FusionEngine {
    public:
        enum Kind { GATHER, KICK, DRIFT, BCS };
        struct Op {
            Kind kind;
            double coef;
            std::function<void()> eager;
        };
        ParticleAttribBase* R = nullptr;   // positions
        const std::size_t* local_num = nullptr;
        std::function<void(std::size_t)> set_local_num;   // several ranks: the container's count follows the migration
        bool busy = false;                 // inside materialise(): every hook is a plain call

        bool active() const { return b200::fusion_enabled() && !busy && R != nullptr && Comm; }
        bool pending() const { return !chain_.empty(); }
        ~FusionEngine() { release(); }

        // -- recording (each returns false when the operation does not continue the pattern) ---------------------------------
        void record_gather(ParticleAttribBase* target, const double* field, std::function<void()> fill_halo, std::function<void()> eager) {
            if (pending()) materialise();   // an older, unconsumed record goes first
            chain_.push_back(Op{GATHER, 0.0, std::move(eager)});
            e_attr_ = target;
            e_field_ = field;
            e_fill_halo_ = std::move(fill_halo);
        }
        bool record_axpy(ParticleAttribBase* y, const ParticleAttribBase* x, double a, std::function<void()> eager) {
            if (chain_.empty() || chain_.front().kind != GATHER) return false;
            const Kind last = chain_.back().kind;
            if (x == e_attr_ && y != R && (last == GATHER || (last == KICK && y == vel_ && kicks() < 2))) {
                vel_ = y;
                chain_.push_back(Op{KICK, a, std::move(eager)});
                return true;
            }
            if (y == R && x == vel_ && last == KICK) {
                chain_.push_back(Op{DRIFT, a, std::move(eager)});
                return true;
            }
            return false;
        }
        bool record_bc(std::function<void()> eager) {
            if (chain_.empty() || chain_.back().kind != DRIFT) return false;
            chain_.push_back(Op{BCS, 0.0, std::move(eager)});
            return true;
        }
        // -- the fused step: true when the record was the whole pattern and has been executed (rho's halo is NOT accumulated) ------
        // `region` (min[3], max[3] of this rank, RegionLayout) is needed on several ranks only.  Collective there: every rank
        // runs the same driver code, so every rank arrives here with the same record; the one rank-local reason not to fuse
        // (the store has to be rebuilt for a new mesh / more particles while the particles are inside it) is agreed on first.
        bool fused_scatter(double q, double* rho, const ipplb_mesh& mesh, const double* region) {
            if (chain_.size() < 4 || chain_.back().kind != BCS) return false;
            const int nk = kicks();
            const double c = chain_[1].coef, dt = chain_[chain_.size() - 2].coef;
            if (nk < 1 || (nk == 2 && chain_[2].coef != c) || c != -(0.5 * dt)) return false;   // not the leapfrog coefficients
            ipplb_ctx* ctx   = b200::ctx();
            const long n     = (long)*local_num;
            const bool multi = Comm->size() > 1;
            const bool stale = bins_ && (n + n / 8 > cap_ || std::memcmp(&mesh, &bins_mesh_, sizeof(mesh)) != 0);
            double veto      = stale && in_bins_ ? 1.0 : 0.0;   // cannot re-bucket from the store: the caller materialises
            Comm->allreduce(veto, 1, std::greater<double>());
            if (veto != 0.0) return false;
            if (stale) release();
            if (!bins_) {
                cap_ = (multi ? 2 * n : n + n / 4) + 65536;   // migration head-room on several ranks
                b200::check(ipplb_bins_create(ctx, &mesh, cap_, &bins_), "fusion: bins_create");
                bins_mesh_ = mesh;
                for (auto& p : spare_) p = b200::device_alloc<double>((std::size_t)cap_);
                cur_ = ipplb_particles{spare_[0], spare_[1], spare_[2], spare_[3], spare_[4], spare_[5], nullptr, q, 0, cap_};
                nxt_ = ipplb_particles{spare_[6], spare_[7], spare_[8], spare_[9], spare_[10], spare_[11], nullptr, q, 0, cap_};
                if (multi) {   // leavers of a step, as records for ipplb_bins_migrate
                    exit_cap_ = (int)std::max<long>(n / 4, 1 << 16);
                    exit_buf_ = b200::device_alloc<double>(6 * (std::size_t)exit_cap_);
                }
            }
            if (!in_bins_) {   // bucket the contiguous attribute arrays once; the particles then live in the store
                ipplb_particles in{R->component_ptr(0),    R->component_ptr(1),    R->component_ptr(2), vel_->component_ptr(0),
                                   vel_->component_ptr(1), vel_->component_ptr(2), nullptr,             q,
                                   n,                      (long)std::min(R->capacity(), vel_->capacity())};
                b200::check(ipplb_bins_build(ctx, bins_, &in, &cur_), "fusion: bins_build");
                in_bins_ = true;
            }
            cur_.q_scalar = nxt_.q_scalar = q;
            ipplb_push push{};
            push.kind     = IPPLB_PUSH_LEAPFROG;
            push.dt       = dt;
            push.do_kick2 = nk == 2;
            push.do_kick1 = push.do_drift = push.do_bc = 1;
            e_fill_halo_();
            b200::check(ipplb_bins_step(ctx, bins_, &push, &cur_, &nxt_, e_field_, rho, exit_buf_, exit_cap_, multi ? region : nullptr,
                                        multi ? region + 3 : nullptr),
                        "fusion: bins_step");
            std::swap(cur_, nxt_);
            if (multi) {   // ParticleSpatialLayout::update for the bucketed store: exchange, append, deposit the arrivals
                b200::check(ipplb_bins_migrate(ctx, bins_, &cur_, exit_buf_, exit_cap_, rho, nullptr, nullptr), "fusion: bins_migrate");
                busy = true;   // the attribute arrays only follow the count (they are stale while the particles are in the store)
                struct Unbusy {
                    bool& b;
                    ~Unbusy() { b = false; }
                } guard{busy};
                set_local_num((std::size_t)cur_.n);
            }
            chain_.clear();
            e_consumed_ = true;
            ++b200::fusion_stats().fused_steps;
            return true;
        }
        // -- back to what the unfused sequence would hold: contiguous arrays valid, pending operations executed in order --------
        void materialise() {
            if (busy || (chain_.empty() && !in_bins_)) return;
            busy = true;
            struct Unbusy {
                bool& b;
                ~Unbusy() { b = false; }
            } guard{busy};
            if (in_bins_) {
                ipplb_particles out{R->component_ptr(0),    R->component_ptr(1),    R->component_ptr(2), vel_->component_ptr(0),
                                    vel_->component_ptr(1), vel_->component_ptr(2), nullptr,             cur_.q_scalar,
                                    0,                      (long)std::min(R->capacity(), vel_->capacity())};
                b200::check(ipplb_bins_compact(b200::ctx(), bins_, &cur_, &out), "fusion: bins_compact");
                in_bins_ = false;
                if (out.n != (long)*local_num)
                    throw IpplException("ippl_b200 fusion", "the bucketed store holds " + std::to_string(out.n) + " particles, the container "
                                                                + std::to_string(*local_num) + " (ipplb_bins_status flags tell why)");
            }
            std::vector<Op> ops;
            ops.swap(chain_);
            if (!ops.empty() && ops.front().kind == GATHER) e_consumed_ = false;   // the gather is about to be computed for real
            for (auto& op : ops) op.eager();
            ++b200::fusion_stats().materialised;
        }
        // called by every public accessor of an attribute of the container
        void sync(const ParticleAttribBase* who) {
            if (busy) return;
            materialise();
            if (who == e_attr_ && e_consumed_)
                throw IpplException("ippl_b200 fusion",
                                    "the field at the particles was consumed by the fused step and never stored; run with IPPL_B200_FUSE=0 "
                                    "if the driver reads it between scatter and the next gather");
        }

    private:
        int kicks() const {
            int k = 0;
            for (auto& op : chain_) k += op.kind == KICK;
            return k;
        }
        void release() {
            if (bins_) ipplb_bins_destroy(bins_);
            bins_ = nullptr;
            for (auto& p : spare_) {
                if (p) cudaFree(p);
                p = nullptr;
            }
            if (exit_buf_) cudaFree(exit_buf_);
            exit_buf_ = nullptr;
            exit_cap_ = 0;
            in_bins_  = false;
        }
        std::vector<Op> chain_;
        ParticleAttribBase *e_attr_ = nullptr, *vel_ = nullptr;   // gather target, velocity attribute
        const double* e_field_ = nullptr;
        std::function<void()> e_fill_halo_;
        bool e_consumed_ = false;
        ipplb_bins* bins_ = nullptr;
        ipplb_mesh bins_mesh_{};
        ipplb_particles cur_{}, nxt_{};
        double* spare_[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        long cap_     = 0;
        bool in_bins_ = false;
        double* exit_buf_ = nullptr;   // several ranks
        int exit_cap_     = 0;
    };
    // a * attrib  and  attrib +/- a * attrib : the expressions the alpine pushes are made of
    template <class A>
    struct Scaled {
        double a;
        const A* x;
    };
    template <class A>
    struct Axpy {
        const A* y;
        double a;
        const A* x;
    };
    // detail::hash_type (src/Types/ViewTypes.h): an index remap in device memory (the reference: Kokkos::View<int*>);
    // here a non-owning (pointer, length) pair
    struct hash_type {
        const int* d = nullptr;
        std::size_t n = 0;
        hash_type() = default;
        hash_type(const int* device_ptr, std::size_t count) : d(device_ptr), n(count) {}
        const int* data() const { return d; }
        std::size_t extent(int) const { return n; }
    };
}  // namespace detail

// Kokkos::RangePolicy<exec_space>(begin, end) as the alpine drivers build it for scatter (AlpineManager.h:178-181)
struct RangePolicy1D {
    long b = 0, e = 0;
    RangePolicy1D() = default;
    RangePolicy1D(long begin_, long end_) : b(begin_), e(end_) {}
    long begin() const { return b; }
    long end() const { return e; }
};

template <typename T>
class ParticleAttrib : public detail::ParticleAttribBase {
public:
    static constexpr int ncomp = detail::ncomp_of<T>::value;
    using value_type           = T;
    using hash_type            = detail::hash_type;
    ParticleAttrib()           = default;
    ~ParticleAttrib() override {
        for (auto p : d_)
            if (p) cudaFree(p);
    }
    // create(n): grows with the communicator's over-allocation factor truncated to int (ParticleAttrib.hpp:37-49)
    void create(std::size_t n) override {
        touch();
        const std::size_t need = count_ + n;
        if (need > capacity_) {
            const int over         = std::max(1, (int)Comm->getDefaultOverallocation());
            const std::size_t ncap = need * over;
            for (int c = 0; c < ncomp; ++c) {
                double* p = b200::device_alloc<double>(ncap);
                if (d_[c]) {
                    b200::cuda_check(cudaMemcpy(p, d_[c], sizeof(double) * count_, cudaMemcpyDeviceToDevice), "ParticleAttrib::create");
                    cudaFree(d_[c]);
                }
                d_[c] = p;
            }
            capacity_ = ncap;
        }
        count_ = need;
    }
    void setCount(std::size_t n) override {
        // from the fusion engine (the count follows a migration inside the bucketed store): a uniform attribute (q = Q / N)
        // stays uniform, its new slots are filled.  From anywhere else the storage is being rewritten by the caller.
        const bool from_engine = engine_ && engine_->busy;
        if (!from_engine) touch();
        reserve(n);
        if constexpr (ncomp == 1) {
            if (from_engine && uniform_valid_ && n > count_)
                b200::check(ipplb_field_fill(b200::ctx(), d_[0] + count_, (long)(n - count_), uniform_value_), "ParticleAttrib::setCount");
        }
        count_ = n;
    }
    // capacity for at least n particles, contents kept
    void reserve(std::size_t n) {
        if (n <= capacity_) return;
        for (int c = 0; c < ncomp; ++c) {
            double* p = b200::device_alloc<double>(n);
            if (d_[c]) {
                b200::cuda_check(cudaMemcpy(p, d_[c], sizeof(double) * count_, cudaMemcpyDeviceToDevice), "ParticleAttrib::reserve");
                cudaFree(d_[c]);
            }
            d_[c] = p;
        }
        capacity_ = n;
    }
    std::size_t size() const { return capacity_; }
    int components() const override { return ncomp; }
    double* component_ptr(int c) const override { return d_[c]; }
    std::size_t capacity() const override { return capacity_; }
    void reserve_storage(std::size_t n) override { reserve(n); }
    bool uniform(double* value) const override {
        if (ncomp != 1 || !uniform_valid_) return false;
        *value = uniform_value_;
        return true;
    }
    void restore_uniform(double value) override {
        if (ncomp == 1) {
            uniform_valid_ = true;
            uniform_value_ = value;
        }
    }
    // getView(): the reference returns a Kokkos::View<T*> for driver-side kernels; here a view(i)[d] / view(i) proxy over
    // the SoA component arrays, usable inside device lambdas (include/ippl/KokkosShim.cuh)
    using view_type = detail::AttribView<ncomp>;
    IPPL_HD view_type getView() const {
#ifndef __CUDA_ARCH__
        touch();   // a view hands out the storage: pending lazy operations run first, the charge is no longer known to be uniform
#endif
        view_type v;
        for (int c = 0; c < ncomp; ++c) v.c[c] = d_[c];
        v.n = count_;
        return v;
    }
    // attrib(i): element access inside kernels (src/Particle/ParticleAttrib.h)
    IPPL_HD auto operator()(std::size_t i) const -> decltype(std::declval<view_type>()(i)) { return getView()(i); }
    std::size_t getParticleCount() const { return count_; }
    double* component(int c) const {
        touch();
        return d_[c];
    }
    // attrib = scalar (ParticleAttrib.hpp:105-116)
    ParticleAttrib& operator=(const T& v) {
        sync();
        for (int c = 0; c < ncomp; ++c)
            b200::check(ipplb_field_fill(b200::ctx(), d_[c], (long)count_, comp(v, c)), "ParticleAttrib::operator=");
        if constexpr (ncomp == 1) {
            uniform_valid_ = true;
            uniform_value_ = comp(v, 0);
        }
        return *this;
    }
    // lazy fusion (detail::FusionEngine): make this attribute's storage say what the unfused sequence would have produced
    void sync() const {
        if (engine_) engine_->sync(this);
    }
    // attrib = expression (ParticleAttrib.hpp:118-130) for y = y +/- a * x: one axpy per component, same rounding
    // as the reference's per-particle y - (a * x)
    ParticleAttrib& operator=(const detail::Axpy<ParticleAttrib>& e) {
        if (e.y != this) throw IpplException("ParticleAttrib::operator=", "only y = y + a * x is supported by the facade");
        if (engine_ && engine_->active() && ncomp == 3) {   // a kick or the drift of a recorded leapfrog step?
            ParticleAttrib* self = this;
            const detail::Axpy<ParticleAttrib> copy = e;
            if (engine_->record_axpy(this, e.x, e.a, [self, copy] { *self = copy; })) return *this;
        }
        sync();
        e.x->sync();
        for (int c = 0; c < ncomp; ++c)
            b200::check(ipplb_axpy(b200::ctx(), (long)count_, e.a, e.x->d_[c], d_[c]), "ParticleAttrib::operator=");
        return *this;
    }
    // sum over local particles (then ranks): ParticleAttrib.hpp:511-532
    double sum(int c = 0) const {
        sync();
        // reuse the interior-sum kernel on a 1-D "mesh" of count_ cells without ghosts
        ipplb_mesh m{};
        m.ng[0] = m.nl[0] = (int)count_;
        m.ng[1] = m.nl[1] = m.ng[2] = m.nl[2] = 1;
        m.nghost                                = 0;
        m.h[0] = m.h[1] = m.h[2] = 1.0;
        double s = 0;
        b200::check(ipplb_field_sum(b200::ctx(), &m, d_[c], &s), "ParticleAttrib::sum");
        Comm->allreduce(s, 1, std::plus<double>());
        return s;
    }
    // host mirror (AoS of T like the reference's view(i)); deep_copy moves it to / from the device SoA
    using HostMirror = std::vector<T>;
    HostMirror getHostMirror() const { return HostMirror(count_); }
    void copyFromHost(const HostMirror& h) {
        touch();
        std::vector<double> tmp(count_);
        for (int c = 0; c < ncomp; ++c) {
            for (std::size_t i = 0; i < count_; ++i) tmp[i] = comp(h[i], c);
            b200::cuda_check(cudaMemcpy(d_[c], tmp.data(), sizeof(double) * count_, cudaMemcpyHostToDevice), "deep_copy");
        }
    }
    void copyToHost(HostMirror& h) const {
        sync();
        h.resize(count_);
        std::vector<double> tmp(count_);
        fence();
        for (int c = 0; c < ncomp; ++c) {
            b200::cuda_check(cudaMemcpy(tmp.data(), d_[c], sizeof(double) * count_, cudaMemcpyDeviceToHost), "deep_copy");
            for (std::size_t i = 0; i < count_; ++i) comp(h[i], c) = tmp[i];
        }
    }
    // scatter(f, pp, policy, hash) (ParticleAttrib.hpp:132-191): deposits the particles policy.begin() .. policy.end(),
    // taken through the index remap hash_array when one is given
    template <typename Field, typename PT>
    void scatter(Field& f, const ParticleAttrib<Vector<PT, 3>>& pp, const RangePolicy1D& policy,
                 const hash_type& hash_array = {}) const {
        static_assert(ncomp == 1, "scatter deposits a scalar attribute");
        sync();
        b200::check(ipplb_scatter_cic(b200::ctx(), &f.b200_mesh(), policy.begin(), policy.end(), pp.component(0),
                                      pp.component(1), pp.component(2), d_[0], 0.0, hash_array.extent(0) ? hash_array.data() : nullptr,
                                      f.data()),
                    "ParticleAttrib::scatter");
        f.accumulateHalo();
    }
    // scatter / gather members (ParticleAttrib.hpp:132-246)
    template <typename Field, typename PT>
    void scatter(Field& f, const ParticleAttrib<Vector<PT, 3>>& pp) const {
        static_assert(ncomp == 1, "scatter deposits a scalar attribute");
        if (engine_ && engine_->active() && engine_->pending() && uniform_valid_
            && static_cast<const detail::ParticleAttribBase*>(&pp) == engine_->R && Field::ncomp == 1) {
            double region[6] = {0, 0, 0, 0, 0, 0};
            if (Comm->size() > 1) {   // this rank's physical region (RegionLayout), for the ownership test inside the kernel
                double o[3], h[3];
                for (int d = 0; d < 3; ++d) {
                    o[d] = f.get_mesh().getOrigin()[d];
                    h[d] = f.get_mesh().getMeshSpacing()[d];
                }
                std::vector<double> all(6 * (std::size_t)Comm->size());
                b200::check(ipplb_layout_regions(f.getLayout().handle(), o, h, all.data()), "ParticleAttrib::scatter (regions)");
                std::copy_n(all.begin() + 6 * Comm->rank(), 6, region);
            }
            if (engine_->fused_scatter(uniform_value_, f.data(), f.b200_mesh(), region)) {
                f.accumulateHalo();   // [gather, kick(s), drift, update] + this scatter ran as ONE fused step
                return;
            }
        }
        sync();
        b200::check(ipplb_scatter_cic(b200::ctx(), &f.b200_mesh(), 0, (long)pp.getParticleCount(), pp.component(0),
                                      pp.component(1), pp.component(2), d_[0], 0.0, nullptr, f.data()),
                    "ParticleAttrib::scatter");
        f.accumulateHalo();
    }
    template <typename Field, typename PT>
    void gather(Field& f, const ParticleAttrib<Vector<PT, 3>>& pp, bool addToAttribute = false) {
        if (engine_ && engine_->active() && ncomp == 3 && Field::ncomp == 3 && !addToAttribute
            && static_cast<const detail::ParticleAttribBase*>(&pp) == engine_->R) {
            ParticleAttrib* self = this;
            Field* fp            = &f;
            const auto* ppp      = &pp;
            engine_->record_gather(this, f.data(), [fp] { fp->fillHalo(); }, [self, fp, ppp] { self->gather(*fp, *ppp, false); });
            return;
        }
        sync();
        f.fillHalo();
        double* out[3] = {d_[0], ncomp > 1 ? d_[1] : nullptr, ncomp > 2 ? d_[2] : nullptr};
        b200::check(ipplb_gather_cic(b200::ctx(), &f.b200_mesh(), (long)pp.getParticleCount(), pp.component(0),
                                     pp.component(1), pp.component(2), f.data(), Field::ncomp, out, addToAttribute),
                    "ParticleAttrib::gather");
    }

private:
    static double comp(const double& v, int) { return v; }
    static double& comp(double& v, int) { return v; }
    template <typename U, unsigned D>
    static double comp(const Vector<U, D>& v, int c) { return v[c]; }
    template <typename U, unsigned D>
    static double& comp(Vector<U, D>& v, int c) { return v[c]; }
    void touch() const {   // storage handed out / rewritten
        sync();
        uniform_valid_ = false;
    }
    std::array<double*, 3> d_{nullptr, nullptr, nullptr};
    std::size_t count_ = 0, capacity_ = 0;
    mutable bool uniform_valid_ = false;   // every element holds uniform_value_ (set by attrib = scalar): what the fused step needs of q
    double uniform_value_       = 0.0;
};

template <typename T>
detail::Scaled<ParticleAttrib<T>> operator*(double a, const ParticleAttrib<T>& x) { return {a, &x}; }
template <class A>
detail::Scaled<A> operator*(double a, const detail::Scaled<A>& s) { return {a * s.a, s.x}; }
template <typename T>
detail::Axpy<ParticleAttrib<T>> operator+(const ParticleAttrib<T>& y, const detail::Scaled<ParticleAttrib<T>>& s) { return {&y, s.a, s.x}; }
template <typename T>
detail::Axpy<ParticleAttrib<T>> operator-(const ParticleAttrib<T>& y, const detail::Scaled<ParticleAttrib<T>>& s) { return {&y, -s.a, s.x}; }

template <typename T>
void deep_copy(ParticleAttrib<T>& dst, const typename ParticleAttrib<T>::HostMirror& src) { dst.copyFromHost(src); }
template <typename T>
void deep_copy(typename ParticleAttrib<T>::HostMirror& dst, const ParticleAttrib<T>& src) { src.copyToHost(dst); }

// free functions (ParticleAttrib.hpp:304-358)
template <typename Attrib1, typename Field, typename Attrib2>
void scatter(const Attrib1& attrib, Field& f, const Attrib2& pp) { attrib.scatter(f, pp); }
template <typename Attrib1, typename Field, typename Attrib2>
void gather(Attrib1& attrib, Field& f, const Attrib2& pp, bool addToAttribute = false) { attrib.gather(f, pp, addToAttribute); }
// ... and the (policy, hash) overload (ParticleAttrib.hpp:332-334)
template <typename Attrib1, typename Field, typename Attrib2>
void scatter(const Attrib1& attrib, Field& f, const Attrib2& pp, const RangePolicy1D& iteration_policy,
             const typename Attrib1::hash_type& hash_array = {}) {
    attrib.scatter(f, pp, iteration_policy, hash_array);
}

// ---- ParticleSpatialLayout (src/Particle/ParticleSpatialLayout.h / .hpp) -----------------------------------------------------------
template <typename T, unsigned Dim, class Mesh = UniformCartesian<T, Dim>>
class ParticleSpatialLayout {
    static_assert(Dim == 3, "the B200 path is three-dimensional");

public:
    using vector_type            = Vector<T, Dim>;
    using particle_position_type = ParticleAttrib<vector_type>;
    ParticleSpatialLayout(FieldLayout<Dim>& fl, Mesh& mesh, bool /*fem*/ = false) : fl_(&fl), mesh_(&mesh) {
        if (Comm->size() > 1) {
            double o[3], h[3];
            for (int d = 0; d < 3; ++d) {
                o[d] = mesh.getOrigin()[d];
                h[d] = mesh.getMeshSpacing()[d];
            }
            b200::check(ipplb_ctx_set_layout(b200::ctx(), fl.handle(), o, h), "ParticleSpatialLayout");
        }
    }
    // ParticleSpatialLayout::updateLayout(fl, mesh), ParticleSpatialLayout.hpp:100-113: new regions for the ownership test
    void updateLayout(FieldLayout<Dim>& fl, Mesh& mesh) {
        fl_   = &fl;
        mesh_ = &mesh;
        if (Comm->size() > 1) {
            double o[3], h[3];
            for (int d = 0; d < 3; ++d) {
                o[d] = mesh.getOrigin()[d];
                h[d] = mesh.getMeshSpacing()[d];
            }
            b200::check(ipplb_ctx_set_layout(b200::ctx(), fl.handle(), o, h), "ParticleSpatialLayout::updateLayout");
        }
    }
    void setParticleBC(BC bc) { bc_ = bc; }
    // update(): applyBC (ParticleLayout.hpp:34-74, PeriodicBC ParticleBC.h:73-76), early return on one rank
    // (ParticleSpatialLayout.hpp:128), else ownership + exchange + compaction (ipplb_update)
    template <class PC>
    void update(PC& pc) {
        if (auto* eng = pc.R.engine(); eng && eng->active() && bc_ == PERIODIC) {   // the BC of a recorded leapfrog step?
            ParticleSpatialLayout* self = this;
            PC* pcp                     = &pc;
            if (eng->record_bc([self, pcp] { self->update(*pcp); })) return;
        }
        auto& R = pc.R;
        long n  = (long)pc.getLocalNum();
        double lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            lo[d] = 0 * mesh_->getMeshSpacing()[d] + mesh_->getOrigin()[d];
            hi[d] = fl_->getDomain()[d].length() * mesh_->getMeshSpacing()[d] + mesh_->getOrigin()[d];
        }
        if (bc_ == PERIODIC)
            b200::check(ipplb_apply_periodic_bc(b200::ctx(), n, R.component(0), R.component(1), R.component(2), lo, hi, 7),
                        "ParticleSpatialLayout::update");
        if (Comm->size() < 2) return;
        pc.migrate();
    }

private:
    FieldLayout<Dim>* fl_;
    Mesh* mesh_;
    BC bc_ = NO;
};

// ---- ParticleBase (src/Particle/ParticleBase.h / .hpp) -------------------------------------------------------------------------------
template <class PLayout>
class ParticleBase {
public:
    using particle_position_type = typename PLayout::particle_position_type;
    using vector_type            = typename PLayout::vector_type;
    particle_position_type R;
    ParticleBase() {
        attributes_.push_back(&R);
        engine_.R             = &R;
        engine_.local_num     = &localNum_;
        engine_.set_local_num = [this](std::size_t n) { setLocalNum(n); };
        R.set_engine(&engine_);
    }
    explicit ParticleBase(PLayout& l) : ParticleBase() { initialize(l); }
    virtual ~ParticleBase() = default;
    void initialize(PLayout& l) { layout_ = &l; }
    void addAttribute(detail::ParticleAttribBase& a) {
        attributes_.push_back(&a);
        a.set_engine(&engine_);
    }
    void setParticleBC(BC bc) { layout_->setParticleBC(bc); }
    void create(std::size_t nLocal) {
        for (auto* a : attributes_) a->create(nLocal);
        localNum_ += nLocal;
    }
    std::size_t getLocalNum() const { return localNum_; }
    std::size_t getTotalNum() const {
        std::size_t t = localNum_;
        Comm->allreduce(t, 1, std::plus<std::size_t>());
        return t;
    }
    void setLocalNum(std::size_t n) {
        localNum_ = n;
        for (auto* a : attributes_) a->setCount(n);
    }
    PLayout& getLayout() { return *layout_; }
    void update() { layout_->update(*this); }
    // Multi-rank exchange behind update() (ParticleSpatialLayout.hpp:150-314, ParticleBase.hpp:175-393).  The C-ABI moves a
    // fixed bundle: the positions R, one vector attribute (the first one registered: the alpine containers' velocity P) and
    // one scalar attribute (the first one registered: the charge q).  Further attributes (the alpine containers' E) are
    // resized only: the drivers recompute them before they read them (gather after every update).  Two collective halves:
    // the plan exchanges the counts, the attributes grow to hold the arrivals (times the over-allocation factor), the
    // commit moves the particles.
    virtual void migrate() {
        detail::ParticleAttribBase *vec = nullptr, *sca = nullptr;
        for (auto* a : attributes_) {
            if (a == &R) continue;
            if (!vec && a->components() == 3) vec = a;
            if (!sca && a->components() == 1) sca = a;
        }
        auto bundle = [&]() {
            ipplb_particles b{};
            b.x = R.component(0); b.y = R.component(1); b.z = R.component(2);
            if (vec) { b.px = vec->component_ptr(0); b.py = vec->component_ptr(1); b.pz = vec->component_ptr(2); }
            if (sca) b.q = sca->component_ptr(0);
            b.n        = (long)localNum_;
            b.capacity = (long)R.size();
            if (vec) b.capacity = std::min<long>(b.capacity, (long)vec->capacity());
            if (sca) b.capacity = std::min<long>(b.capacity, (long)sca->capacity());
            return b;
        };
        // lazy fusion wants to know whether the charge is still ONE value after the exchange: it is when every rank's was
        // uniform with the same value (two small reductions, only with fusion on)
        double uq = 0.0;
        bool uniform_everywhere = false;
        if (b200::fusion_enabled() && sca) {
            const bool mine = sca->uniform(&uq);
            double hi = mine ? uq : std::numeric_limits<double>::max(), neg_lo = mine ? -uq : std::numeric_limits<double>::max();
            Comm->allreduce(hi, 1, std::greater<double>());
            Comm->allreduce(neg_lo, 1, std::greater<double>());
            uniform_everywhere = mine && hi == uq && -neg_lo == uq;
        }
        ipplb_particles b = bundle();
        long n_after      = b.n;
        const int rc      = ipplb_update_plan(b200::ctx(), &b, &n_after, nullptr, nullptr);
        if (rc != IPPLB_OK && rc != IPPLB_ERR_CAPACITY) b200::check(rc, "ParticleBase::update (plan)");
        if (n_after > b.capacity) {
            const std::size_t want = (std::size_t)n_after * (std::size_t)std::max(1, (int)Comm->getDefaultOverallocation());
            for (auto* a : attributes_) a->reserve_storage(want);
            b = bundle();
        }
        b200::check(ipplb_update_commit(b200::ctx(), &b), "ParticleBase::update (commit)");
        setLocalNum((std::size_t)b.n);
        if (uniform_everywhere) sca->restore_uniform(uq);
    }

protected:
    PLayout* layout_ = nullptr;
    std::vector<detail::ParticleAttribBase*> attributes_;
    std::size_t localNum_ = 0;
    detail::FusionEngine engine_;   // lazy fusion of the leapfrog sequence (off unless IPPL_B200_FUSE=1)
};

// ---- RegionLayout (src/Region/RegionLayout.h): physical region of every rank ---------------------------------------------------------
namespace detail {
    template <typename T, unsigned Dim, class Mesh = UniformCartesian<T, Dim>>
    class RegionLayout {
    public:
        RegionLayout() = default;
        RegionLayout(const FieldLayout<Dim>& fl, const Mesh& mesh, bool /*fem*/ = false) {
            double o[3], h[3];
            for (int d = 0; d < 3; ++d) {
                o[d] = mesh.getOrigin()[d];
                h[d] = mesh.getMeshSpacing()[d];
            }
            regions_.resize(6 * (std::size_t)Comm->size());
            b200::check(ipplb_layout_regions(fl.handle(), o, h, regions_.data()), "RegionLayout");
        }
        const std::vector<double>& regions() const { return regions_; }  // [rank][min 3, max 3]

    private:
        std::vector<double> regions_;
    };
}  // namespace detail

// ---- Random (src/Random/Distribution.h, NormalDistribution.h, InverseTransformSampling.h, Randn.h) on the device sampler ------------------
// (include/ippl/compat/Random/*.h provide the reference's functor-shaped classes instead when the reference's own drivers are compiled)
#ifndef IPPL_B200_REFERENCE_SHAPED_RANDOM
namespace random {
    // Distribution<T, Dim, 2 * Dim, Functions>: the reference takes host/device functors (CDF, PDF, Estimate); the
    // facade names the three families the alpine managers use (one per dimension) and the device evaluates them
    enum Kind { UNIFORM = IPPLB_DIST_UNIFORM, COSINE = IPPLB_DIST_COSINE, NORMAL = IPPLB_DIST_NORMAL };
    template <typename T, unsigned Dim>
    class Distribution {
        static_assert(Dim == 3, "the B200 path is three-dimensional");

    public:
        Distribution(const std::array<Kind, Dim>& kind, const T* par_p) {
            for (unsigned d = 0; d < Dim; ++d) {
                d_.kind[d]        = kind[d];
                d_.par[2 * d]     = par_p[2 * d];
                d_.par[2 * d + 1] = par_p[2 * d + 1];
            }
        }
        const ipplb_dist& handle() const { return d_; }
        // rho(cell) = getFullPdf((global index + 0.5) * hr + origin) on the interior: the weights of the first
        // repartition (LandauDampingManager.h:188-199)
        template <class FieldT>
        void fillFullPdf(FieldT& rho) const {
            b200::check(ipplb_field_fill_pdf(b200::ctx(), &rho.b200_mesh(), &d_, rho.data()), "Distribution::getFullPdf");
        }

    private:
        ipplb_dist d_{};
    };
    template <typename T, unsigned Dim>
    class NormalDistribution : public Distribution<T, Dim> {
    public:
        explicit NormalDistribution(const T* par_p) : Distribution<T, Dim>({NORMAL, NORMAL, NORMAL}, par_p) {}
    };

    // InverseTransformSampling<T, Dim, DeviceType, Distribution> (InverseTransformSampling.h:30-256).  generate()
    // takes a seed instead of a Kokkos::Random_XorShift64_Pool: the stream is counter based (Philox4x32-10).
    template <typename T, unsigned Dim, class DeviceType, class Dist>
    class InverseTransformSampling {
    public:
        using size_type = detail::size_type;
        template <class RegionLayout>
        InverseTransformSampling(Dist& dist, Vector<T, Dim>& rmax, Vector<T, Dim>& rmin, const RegionLayout& rlayout,
                                 size_type ntotal)
            : dist_(dist) {
            const int nr = Comm->size();
            std::vector<long> nloc(nr);
            std::vector<double> ub(6 * (std::size_t)nr);
            double lo[3], hi[3];
            for (int d = 0; d < 3; ++d) {
                lo[d] = rmin[d];
                hi[d] = rmax[d];
            }
            b200::check(ipplb_sample_counts(&dist.handle(), lo, hi, rlayout.regions().data(), nr, (long)ntotal, nloc.data(),
                                            ub.data()),
                        "InverseTransformSampling");
            nlocal_ = (size_type)nloc[Comm->rank()];
            for (int d = 0; d < 3; ++d) {
                umin_[d] = ub[6 * Comm->rank() + d];
                umax_[d] = ub[6 * Comm->rank() + 3 + d];
            }
        }
        size_type getLocalSamplesNum() const { return nlocal_; }
        void setLocalSamplesNum(size_type n) { nlocal_ = n; }
        // generate(view, rand_pool64), :235-244
        void generate(ParticleAttrib<Vector<T, Dim>>& R, std::uint64_t seed) { generate(R, 0, nlocal_, seed); }
        void generate(ParticleAttrib<Vector<T, Dim>>& R, size_type begin, size_type end, std::uint64_t seed) {
            b200::check(ipplb_sample_positions(b200::ctx(), &dist_.handle(), umin_, umax_, seed, (long)begin, (long)(end - begin),
                                               R.component(0) + begin, R.component(1) + begin, R.component(2) + begin),
                        "InverseTransformSampling::generate");
        }

    private:
        Dist dist_;
        size_type nlocal_ = 0;
        double umin_[3], umax_[3];
    };

    // Kokkos::parallel_for(RangePolicy(begin, end), randn<T, Dim>(P, pool, mu, sd)), Randn.h:30-94
    template <typename T, unsigned Dim>
    void randn(ParticleAttrib<Vector<T, Dim>>& P, std::uint64_t seed, const T* mu, const T* sd, std::size_t begin, std::size_t end) {
        b200::check(ipplb_sample_normal(b200::ctx(), mu, sd, seed, (long)begin, (long)(end - begin), P.component(0) + begin,
                                        P.component(1) + begin, P.component(2) + begin),
                    "random::randn");
    }
}  // namespace random
#endif  // IPPL_B200_REFERENCE_SHAPED_RANDOM

// ---- OrthogonalRecursiveBisection (src/Decomposition/OrthogonalRecursiveBisection.h / .hpp) ---------------------------------------------
template <class FieldT, class Tp = double>
class OrthogonalRecursiveBisection {
    static constexpr unsigned Dim = 3;
    using mesh_type               = typename FieldT::Mesh_t;

public:
    FieldT bf_m;  // the weights: a copy of rho on the first repartition, scatterR(R) afterwards

    // initialize(fl, mesh, rho), .hpp:8-12
    void initialize(FieldLayout<Dim>& fl, mesh_type& mesh, const FieldT& rho) {
        bf_m.initialize(mesh, fl);
        bf_m.copyFrom(rho);
    }
    // binaryRepartition(R, fl, isFirstRepartition), .hpp:14-105: plane sums on the device, reduced over ranks; the cuts
    // (findCutAxis / findMedian / cutDomain) on the host; false when a box would get an axis of length 1
    template <typename Attrib>
    bool binaryRepartition(const Attrib& R, FieldLayout<Dim>& fl, const bool& isFirstRepartition) {
        if (!isFirstRepartition) scatterR(R);
        const int nr = Comm->size();
        std::vector<int> boxes(6 * (std::size_t)nr);
        int ok = 0;
        b200::check(ipplb_orb_repartition(b200::ctx(), &bf_m.b200_mesh(), nr, bf_m.data(), boxes.data(), &ok),
                    "OrthogonalRecursiveBisection::binaryRepartition");
        if (!ok) return false;
        std::vector<NDIndex<Dim>> domains(nr);
        for (int r = 0; r < nr; ++r)
            domains[r] = NDIndex<Dim>(Index(boxes[6 * r], boxes[6 * r + 3]), Index(boxes[6 * r + 1], boxes[6 * r + 4]),
                                      Index(boxes[6 * r + 2], boxes[6 * r + 5]));
        fl.updateLayout(domains);
        bf_m.updateLayout(fl);
        return true;
    }
    // scatterR, .hpp:234-300: CIC deposit of weight 1 per particle, then accumulateHalo
    template <typename Attrib>
    void scatterR(const Attrib& r) {
        bf_m = 0.0;
        b200::check(ipplb_scatter_cic(b200::ctx(), &bf_m.b200_mesh(), 0, (long)r.getParticleCount(), r.component(0), r.component(1),
                                      r.component(2), nullptr, 1.0, nullptr, bf_m.data()),
                    "OrthogonalRecursiveBisection::scatterR");
        bf_m.accumulateHalo();
    }
};

// ---- FFTPeriodicPoissonSolver (src/PoissonSolvers/FFTPeriodicPoissonSolver.h; non-owned stage, cuFFT) ------------------------------------
template <class FieldLHS, class FieldRHS>
class FFTPeriodicPoissonSolver {
public:
    enum OutputType { SOL = 1, GRAD = 2, SOL_AND_GRAD = 3 };
    FFTPeriodicPoissonSolver() = default;
    FFTPeriodicPoissonSolver(FieldLHS& lhs, FieldRHS& rhs) {
        setLhs(lhs);
        setRhs(rhs);
    }
    ~FFTPeriodicPoissonSolver() {
        if (h_) ipplb_poisson_destroy(h_);
    }
    // heFFTe options / output type: the cuFFT stage always returns the gradient (GRAD), the rest does not apply
    template <class PL>
    void mergeParameters(const PL&) {}
    void setRhs(FieldRHS& rhs) {
        rhs_ = &rhs;
        if (h_) ipplb_poisson_destroy(h_);
        h_ = nullptr;
        if (Comm->size() > 1) {  // replicated solve over NCCL (ipplb_poisson_create_dist)
            double o[3], h[3];
            for (int d = 0; d < 3; ++d) {
                o[d] = rhs.get_mesh().getOrigin()[d];
                h[d] = rhs.get_mesh().getMeshSpacing()[d];
            }
            b200::check(ipplb_poisson_create_dist(b200::ctx(), rhs.getLayout().handle(), o, h, &h_), "FFTPeriodicPoissonSolver::setRhs");
        } else {
            b200::check(ipplb_poisson_create(b200::ctx(), &rhs.b200_mesh(), &h_), "FFTPeriodicPoissonSolver::setRhs");
        }
    }
    void setLhs(FieldLHS& lhs) { lhs_ = &lhs; }
    // solve(): GRAD output -- E written to lhs interior, rho clobbered like the reference (:53-169)
    void solve() { b200::check(ipplb_poisson_solve(h_, rhs_->data(), lhs_->data()), "FFTPeriodicPoissonSolver::solve"); }

private:
    ipplb_poisson* h_ = nullptr;
    FieldLHS* lhs_    = nullptr;
    FieldRHS* rhs_    = nullptr;
};

}  // namespace ippl

// ---- IpplTimings (src/Utility/IpplTimings.h): named wall timers with a stream fence on start/stop -----------------------------------------------
namespace ippl {
// ippl::ParameterList (src/Utility/ParameterList.h:29-160): named solver / FFT parameters of mixed type, nested lists
// allowed.  Same interface (add / get / get with default / contains / merge / update / operator<<).
class ParameterList {
public:
    using variant_t = std::variant<double, float, bool, std::string, unsigned int, int, std::shared_ptr<ParameterList>>;
    template <typename T>
    void add(const std::string& key, const T& value) {
        if (params_.count(key)) throw IpplException("ParameterList::add()", "Parameter '" + key + "' already exists.");
        params_[key] = wrap(value);
    }
    template <typename T>
    T get(const std::string& key) const {
        auto it = params_.find(key);
        if (it == params_.end()) throw IpplException("ParameterList::get()", "Parameter '" + key + "' not contained.");
        return unwrap<T>(it->second);
    }
    template <typename T>
    T get(const std::string& key, const T& defval) const {
        auto it = params_.find(key);
        return it == params_.end() ? defval : unwrap<T>(it->second);
    }
    bool contains(const std::string& key) const { return params_.count(key) != 0; }
    void merge(const ParameterList& p) noexcept {
        for (const auto& kv : p.params_) params_[kv.first] = kv.second;
    }
    void update(const ParameterList& p) noexcept {
        for (const auto& kv : p.params_)
            if (params_.count(kv.first)) params_[kv.first] = kv.second;
    }
    template <typename T>
    void update(const std::string& key, const T& value) {
        if (!params_.count(key)) throw IpplException("ParameterList::update()", "Parameter '" + key + "' does not exist.");
        params_[key] = wrap(value);
    }
    friend std::ostream& operator<<(std::ostream& os, const ParameterList& p) {
        p.print(os, 0);
        return os;
    }

private:
    template <typename T>
    static variant_t wrap(const T& v) {
        if constexpr (std::is_same_v<T, ParameterList>) return std::make_shared<ParameterList>(v);
        else if constexpr (std::is_enum_v<T>) return static_cast<int>(v);
        else if constexpr (std::is_convertible_v<T, std::string> && !std::is_arithmetic_v<T>) return std::string(v);
        else return v;
    }
    template <typename T>
    static T unwrap(const variant_t& v) {
        if constexpr (std::is_same_v<T, ParameterList>) return *std::get<std::shared_ptr<ParameterList>>(v);
        else return std::get<T>(v);
    }
    void print(std::ostream& os, int indent) const {
        std::size_t k = 0;
        for (const auto& kv : params_) {
            os << std::string(indent, ' ') << std::left << std::setw(20) << kv.first << " ";
            std::visit([&](const auto& a) {
                using A = std::decay_t<decltype(a)>;
                if constexpr (std::is_same_v<A, std::shared_ptr<ParameterList>>) { os << "\n"; a->print(os, indent + 4); }
                else os << a;
            }, kv.second);
            if (++k != params_.size()) os << "\n";
        }
    }
    std::map<std::string, variant_t> params_;
};
}  // namespace ippl

// IpplTimings (src/Utility/IpplTimings.h / .cpp:226-330): wall-clock timers fenced on the context's stream (the reference
// fences Kokkos); print() reduces every timer over the ranks (max / average / min) and lists the measurement counts,
// print(file) writes the same block to a file (the drivers' "timing.dat").
class IpplTimings {
public:
    using TimerRef = int;
    static TimerRef getTimer(const char* name) {
        auto& s = state();
        auto it = s.index.find(name);
        if (it != s.index.end()) return it->second;
        s.names.push_back(name);
        s.total.push_back(0.0);
        s.count.push_back(0);
        s.start.push_back({});
        return s.index[name] = (int)s.names.size() - 1;
    }
    static void startTimer(TimerRef t) {
        if (ippl::b200::ctx_ref()) ipplb_sync(ippl::b200::ctx_ref());
        state().start[t] = std::chrono::steady_clock::now();
    }
    static void stopTimer(TimerRef t) {
        if (ippl::b200::ctx_ref()) ipplb_sync(ippl::b200::ctx_ref());
        state().total[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - state().start[t]).count();
        state().count[t] += 1;
    }
    static double seconds(TimerRef t) { return state().total[t]; }
    static void print(std::ostream& os = std::cout) {
        auto& s = state();
        if (s.names.empty()) return;
        const int nranks = ippl::Comm ? ippl::Comm->size() : 1;
        const int rank   = ippl::Comm ? ippl::Comm->rank() : 0;
        std::ostringstream msg;
        msg << "---------------------------------------------\n";
        msg << "     Timing results for " << nranks << " rank(s):\n";
        msg << "---------------------------------------------\n";
        for (std::size_t i = 0; i < s.names.size(); ++i) {
            double wmax = s.total[i], wneg = -s.total[i], wsum = s.total[i];
            if (nranks > 1) {   // collective: every rank walks the same timer list (IpplTimings.cpp:254-258)
                ippl::Comm->allreduce(wmax, 1, std::greater<double>());
                ippl::Comm->allreduce(wneg, 1, std::greater<double>());
                ippl::Comm->allreduce(wsum, 1, std::plus<double>());
            }
            const std::string nm = s.names[i].substr(0, std::min<std::size_t>(s.names[i].size(), 19));
            const std::string pad(20 - nm.size(), '.');
            if (i == 0) {
                msg << nm << pad << " Wall tot = " << std::setw(10) << wmax << "\n\n";
            } else {
                msg << nm << pad << " Wall max = " << std::setw(10) << wmax << "\n"
                    << std::string(20, ' ') << " Wall avg = " << std::setw(10) << wsum / nranks << "\n"
                    << std::string(20, ' ') << " Wall min = " << std::setw(10) << -wneg << "\n\n";
            }
        }
        msg << "---------------------------------------------\n     Measurement counts:\n---------------------------------------------\n";
        for (std::size_t i = 0; i < s.names.size(); ++i) {
            const std::string nm = s.names[i].substr(0, std::min<std::size_t>(s.names[i].size(), 19));
            msg << nm << std::string(20 - nm.size(), '.') << " Count = " << std::setw(10) << s.count[i] << "\n";
        }
        msg << "---------------------------------------------\n";
        if (rank == 0) os << msg.str() << std::flush;
    }
    // Timing::print(fn, problemSize) (IpplTimings.cpp:311-330): the same block into a file, written by rank 0
    static void print(const std::string& fn, const std::map<std::string, unsigned int>& problemSize = {}) {
        std::ostringstream body;
        print(body);
        if (ippl::Comm && ippl::Comm->rank() != 0) return;
        std::ofstream f(fn.c_str(), std::ios::out);
        if (!problemSize.empty()) {
            f << "Problem size:\n";
            for (auto& kv : problemSize) f << "    " << std::setw(10) << kv.first << ": " << kv.second << "\n";
            f << "\n";
        }
        f << body.str();
    }

private:
    struct State {
        std::map<std::string, int> index;
        std::vector<std::string> names;
        std::vector<double> total;
        std::vector<long> count;
        std::vector<std::chrono::steady_clock::time_point> start;
    };
    static State& state() {
        static State s;
        return s;
    }
};

#endif  // IPPL_B200_FACADE_H
