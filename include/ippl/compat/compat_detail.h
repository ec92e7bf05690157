// compat_detail.h -- framework names the reference's drivers touch beyond the facade's classes: MPI (two calls), field
// boundary-condition objects (periodic only), ViewType, the heFFTe communication enum.  Included by compat/Ippl.h.
#pragma once

// ---- MPI as the drivers use it (LoadBalancer.hpp:152-153, FieldContainer.hpp:19): one process per GPU, NCCL underneath ----------
using MPI_Comm     = int;
using MPI_Datatype = int;
using MPI_Op       = int;
constexpr MPI_Comm MPI_COMM_WORLD = 0;
constexpr MPI_Datatype MPI_INT = 1, MPI_DOUBLE = 2, MPI_UNSIGNED_LONG = 3;
constexpr MPI_Op MPI_SUM = 1;
// MPI_Allgather of `count` ints per rank: one sum all-reduce per slot (each rank contributes its own slots, zeros elsewhere)
inline int MPI_Allgather(const void* send, int count, MPI_Datatype st, void* recv, int, MPI_Datatype, MPI_Comm) {
    if (st != MPI_INT) throw IpplException("MPI_Allgather", "only MPI_INT is wired (LoadBalancer::balance)");
    const int nr = ippl::Comm->size(), me = ippl::Comm->rank();
    const int* s = static_cast<const int*>(send);
    int* r       = static_cast<int*>(recv);
    for (int k = 0; k < nr * count; ++k) {
        long v = (k / count == me) ? s[k % count] : 0;
        if (nr > 1) ippl::b200::check(ipplb_allreduce_sum_i64(ippl::b200::ctx(), &v), "MPI_Allgather");
        r[k] = (int)v;
    }
    return 0;
}
// MPI_Reduce(sum) of doubles to rank 0 (every rank receives the sum: a superset)
inline int MPI_Reduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
    if (t != MPI_DOUBLE) throw IpplException("MPI_Reduce", "only MPI_DOUBLE sums are wired");
    const double* s = static_cast<const double*>(send);
    double* r       = static_cast<double*>(recv);
    for (int k = 0; k < count; ++k) {
        double v = s[k];
        if (ippl::Comm->size() > 1) ippl::b200::check(ipplb_allreduce_sum_f64(ippl::b200::ctx(), &v), "MPI_Reduce");
        r[k] = v;
    }
    return 0;
}

// MPI_Allreduce(sum) of unsigned longs / ints / doubles
inline int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op, MPI_Comm) {
    for (int k = 0; k < count; ++k) {
        if (t == MPI_DOUBLE) {
            double v = static_cast<const double*>(send)[k];
            if (ippl::Comm->size() > 1) ippl::b200::check(ipplb_allreduce_sum_f64(ippl::b200::ctx(), &v), "MPI_Allreduce");
            static_cast<double*>(recv)[k] = v;
        } else {
            long v = t == MPI_INT ? (long)static_cast<const int*>(send)[k] : (long)static_cast<const unsigned long*>(send)[k];
            if (ippl::Comm->size() > 1) ippl::b200::check(ipplb_allreduce_sum_i64(ippl::b200::ctx(), &v), "MPI_Allreduce");
            if (t == MPI_INT) static_cast<int*>(recv)[k] = (int)v;
            else static_cast<unsigned long*>(recv)[k] = (unsigned long)v;
        }
    }
    return 0;
}

// MPI_Bcast of doubles from `root`: the root's value summed with zeros
inline int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm) {
    if (t != MPI_DOUBLE) throw IpplException("MPI_Bcast", "only MPI_DOUBLE is wired");
    for (int k = 0; k < count; ++k) {
        double v = ippl::Comm->rank() == root ? static_cast<double*>(buf)[k] : 0.0;
        if (ippl::Comm->size() > 1) ippl::b200::check(ipplb_allreduce_sum_f64(ippl::b200::ctx(), &v), "MPI_Bcast");
        static_cast<double*>(buf)[k] = v;
    }
    return 0;
}

namespace ippl {
// heFFTe's communication pattern names (src/FFT): the periodic solver here is cuFFT, the value is accepted and ignored
enum FFTComm { a2a = 0, a2av = 1, p2p = 2, p2p_pl = 3 };

namespace detail {
    // ViewType<T, 1>::view_type (src/Types/ViewTypes.h): what getView() of a particle attribute returns
    template <typename T, unsigned Rank>
    struct ViewType;
    template <typename T>
    struct ViewType<Vector<T, 3>, 1> {
        using view_type = AttribView<3>;
    };
    template <>
    struct ViewType<double, 1> {
        using view_type = AttribView<1>;
    };
}  // namespace detail

// field boundary conditions (src/Field/BcTypes.h): the drivers build an all-periodic set for the CG / FEM solvers' potential
template <class FieldT>
struct PeriodicFace {
    explicit PeriodicFace(unsigned face_) : face(face_) {}
    unsigned face;
};
template <class FieldT, unsigned Dim>
struct BConds {
    std::array<std::shared_ptr<PeriodicFace<FieldT>>, 2 * Dim> bc;
    std::shared_ptr<PeriodicFace<FieldT>>& operator[](unsigned i) { return bc[i]; }
};
}  // namespace ippl

// ---- field = dot(vector field, vector field) (PenningTrapManager.h:350: rho = dot(E, E) for the potential energy) -------------------
namespace ippl {
namespace detail {
    template <class VF>
    struct FieldDot {
        const VF* a;
        const VF* b;
        // cell by cell over the whole ghosted array (the reference's expression assignment covers the ghost layers too)
        template <class SF>
        void assign_to(SF& f) const {
            auto av = a->getView();
            auto bv = b->getView();
            auto fv = f.getView();
            using index_array_type = typename ippl::RangePolicy<3>::index_array_type;
            ippl::parallel_for(
                "field = dot(a, b)", ippl::getRangePolicy(fv, 0), KOKKOS_LAMBDA(const index_array_type& args) {
                    ippl::apply(fv, args) = ippl::detail::dot3(ippl::apply(av, args), ippl::apply(bv, args)).apply();
                });
        }
    };
}  // namespace detail
template <typename T, unsigned Dim, class M, class C>
detail::FieldDot<Field<Vector<T, Dim>, Dim, M, C>> dot(const Field<Vector<T, Dim>, Dim, M, C>& a, const Field<Vector<T, Dim>, Dim, M, C>& b) {
    return {&a, &b};
}
}  // namespace ippl

// ---- two-dimensional stand-ins: BumponTailInstabilityManager.h carries a phase-space dump (struct PhaseDump, :372-436) that
// the reference compiles out (`EnablePhaseDump = false`, :21) but that still has to parse and instantiate: 2-D layout,
// mesh, field and particle attribute types.  The B200 path is three-dimensional; these satisfy the types and throw if
// anything ever calls them.
namespace ippl {
namespace detail {
    [[noreturn]] inline void no_2d() { throw IpplException("PhaseDump", "two-dimensional fields are not part of the B200 path"); }
}
template <>
class FieldLayout<2> {
public:
    FieldLayout() = default;
    template <typename CommT>
    FieldLayout(const CommT&, const NDIndex<2>&, std::array<bool, 2>, bool = false) {}
};
template <class M, class C>
class Field<double, 2, M, C> {
public:
    struct view_type {
        double* d     = nullptr;
        std::size_t n = 0;
        double* data() const { return d; }
        std::size_t size() const { return n; }
    };
    void initialize(M&, FieldLayout<2>&) { detail::no_2d(); }
    NDIndex<2> getOwned() const { return NDIndex<2>(); }
    Field& operator=(double) { detail::no_2d(); }
    view_type& getView() { return view_; }
    void write(Inform&) { detail::no_2d(); }
    double max() { detail::no_2d(); }
    double min() { detail::no_2d(); }

private:
    view_type view_;
};
template <>
class ParticleAttrib<Vector<double, 2>> {
public:
    void realloc(std::size_t) { detail::no_2d(); }
    __host__ __device__ Vector<double, 2>& operator()(std::size_t i) const { return d_[i]; }

private:
    Vector<double, 2>* d_ = nullptr;
};
template <typename Attrib1, class M, class C>
void scatter(const Attrib1&, Field<double, 2, M, C>&, const ParticleAttrib<Vector<double, 2>>&) { detail::no_2d(); }
}  // namespace ippl
