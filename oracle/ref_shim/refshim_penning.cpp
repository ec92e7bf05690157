// refshim_penning.cpp -- TEST INFRASTRUCTURE.  The PenningTrap kicks as the reference writes them: the bodies of the
// "Kick1" / "Kick2" lambdas of demos/alpine/PenningTrapManager.h (:256-272, :313-333) are cut out of the reference file
// at build time (gen_snippets.py -> penning_kick{1,2}.inc in a temporary include directory) and compiled here,
// unchanged, inside a plain loop over the particles.  Views are AoS Vector<double,3>-like: view(j)[d].
#include <Kokkos_Core.hpp>

#include <cstddef>

namespace {
    struct AoS3 {
        double* p;
        double* operator()(std::size_t j) const { return p + 3 * j; }
    };
}  // namespace

extern "C" {

// which = 1: Kick1 (P += alpha (E + P x B), x then y with the updated Px), 2: Kick2 (the implicit half).  R, P, E: [n][3]
void refpenning_kick(int which, long n, const double* R, double* P, const double* E, const double* origin_,
                     const double* length_, double V0, double alpha, double Bext, double DrInv) {
    const double origin[3] = {origin_[0], origin_[1], origin_[2]};
    const double length[3] = {length_[0], length_[1], length_[2]};
    if (which == 1) {
        AoS3 Rview{const_cast<double*>(R)}, Pview{P}, Eview{const_cast<double*>(E)};
        for (std::size_t j = 0; j < (std::size_t)n; ++j) {
#include "penning_kick1.inc"
        }
    } else {
        AoS3 R2view{const_cast<double*>(R)}, P2view{P}, E2view{const_cast<double*>(E)};
        for (std::size_t j = 0; j < (std::size_t)n; ++j) {
#include "penning_kick2.inc"
        }
    }
    (void)DrInv;
}

}  // extern "C"
