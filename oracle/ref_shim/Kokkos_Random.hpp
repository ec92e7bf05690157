// stand-in for <Kokkos_Random.hpp> (oracle/ref_shim): a "pool" that replays a caller-supplied sequence of uniforms, so
// that the reference's samplers (fill_random, randn) can be driven with the SAME uniforms as the restatement.
#pragma once
#include <cstddef>
namespace refshim {
    struct ReplayGenerator {
        const double* u;
        std::size_t* cursor;
        double drand() { return u[(*cursor)++]; }
        double drand(double lo, double hi) { return lo + (hi - lo) * drand(); }
        double normal(double mean = 0.0, double sd = 1.0) { return mean + sd * drand(); }  // replays pre-made normals
    };
    struct ReplayPool {
        using generator_type = ReplayGenerator;
        const double* u;
        std::size_t* cursor;
        generator_type get_state() const { return generator_type{u, cursor}; }
        void free_state(const generator_type&) const {}
    };
}  // namespace refshim
namespace Kokkos {
    template <class Device = void> using Random_XorShift64_Pool = refshim::ReplayPool;
}
