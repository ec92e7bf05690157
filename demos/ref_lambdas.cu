// ref_lambdas.cu -- the reference drivers' OWN Kokkos lambdas on the B200 facade, unchanged.
//
// The alpine mini-apps carry device code that is not part of IPPL (see include/ippl/KokkosShim.cuh).  Here the bodies of
//   "Kick1" / "Kick2"  (demos/alpine/PenningTrapManager.h:256-272, 313-333)
//   "Ex stats"         (demos/alpine/LandauDampingManager.h:346-360 over the field, :401-406 over the particles)
//   "Particle Kinetic Energy" / "Vector E reduce"  (PenningTrapManager.h:354-360, :374-382)
//   "Ex inner product" / "Ex max norm"             (BumponTailInstabilityManager.h:460-468, :478-488)
// are cut out of the reference tree AT BUILD TIME (demos/gen_ref_lambdas.py -> a temporary include directory; the
// Makefile target `ref_lambdas` exists only where /root/reference does) and compiled by nvcc exactly as the drivers
// write them -- same view names, same captures, same Kokkos calls -- on include/ippl/KokkosShim.cuh, against
// ParticleAttrib / Field objects of the facade.  Every kernel is checked against the C-ABI kernel that replaces it in
// demos/*.cpp: ipplb_penning_kick (bit for bit: this file is compiled with --fmad=false like the oracle),
// ipplb_field_ex_stats, ipplb_field_energy_stats, ipplb_particles_kinetic, and a host sum.  Needs a CUDA device to run; there is no CPU fallback.
constexpr unsigned Dim = 3;
using T                = double;
const char* TestName   = "ref_lambdas";

#include "ippl/KokkosShim.cuh"

#include "Alpine.h"

#include <random>

static int check(const char* what, bool ok) {
    std::cout << "  " << what << (ok ? ": ok" : ": FAILED") << std::endl;
    return ok ? 0 : 1;
}

int main(int argc, char* argv[]) {
    try {
        ippl::initialize(argc, argv);
    } catch (const IpplException& ex) {
        std::cerr << TestName << ": cannot start: " << ex.what() << " (a CUDA device is required; there is no CPU fallback)" << std::endl;
        return 2;
    }
    int rc = 0;
    {
        const size_type n = 200000;
        Vector_t<int, Dim> nr(16);
        ippl::NDIndex<Dim> domain;
        for (unsigned d = 0; d < Dim; ++d) domain[d] = ippl::Index(nr[d]);
        // PenningTrapManager::pre_run constants (PenningTrapManager.h:56-74)
        Vector_t<double, Dim> rmin(0.0), rmax(20.0), origin_v(0.0);
        Vector_t<double, Dim> hr = rmax / 16.0;
        std::array<bool, Dim> decomp{true, true, true};
        auto fc = std::make_shared<FieldContainer<T, Dim>>(hr, rmin, rmax, decomp, domain, origin_v, true);
        fc->initializeFields();
        auto pc = std::make_shared<ParticleContainer<T, Dim>>(fc->getMesh(), fc->getFL());
        pc->create(n);
        std::mt19937_64 gen(7);
        std::uniform_real_distribution<double> ur(0.0, 20.0);
        std::normal_distribution<double> nd(0.0, 1.0);
        std::vector<ippl::Vector<double, 3>> hR(n), hP(n), hE(n);
        for (size_type i = 0; i < n; ++i)
            for (unsigned d = 0; d < Dim; ++d) {
                hR[i][d] = ur(gen);
                hP[i][d] = nd(gen);
                hE[i][d] = 0.3 * nd(gen);
            }
        pc->R.copyFromHost(hR);
        pc->P.copyFromHost(hP);
        pc->E.copyFromHost(hE);

        // ---- PenningTrapManager::LeapFrogStep: the two kicks, driver text -----------------------------------------------
        const double dt                 = 0.5 * 20.0 / 2048;
        double alpha                    = -0.5 * dt;
        double Bext                     = 5.0;
        double DrInv                    = 1.0 / (1 + (std::pow((alpha * Bext), 2)));
        Vector_t<double, Dim> length    = rmax - rmin;
        Vector_t<double, Dim> origin    = origin_v;
        double V0                       = 30 * length[2];
        // the C-ABI kernels on a copy of P
        ipplb_push push{};
        push.kind = IPPLB_PUSH_PENNING; push.dt = dt; push.do_kick1 = push.do_kick2 = push.do_drift = push.do_bc = 1;
        for (unsigned d = 0; d < Dim; ++d) { push.origin[d] = origin[d]; push.length[d] = length[d]; }
        push.V0 = V0; push.alpha = alpha; push.Bext = Bext; push.DrInv = DrInv;
        ippl::ParticleAttrib<ippl::Vector<double, 3>> Pc;
        for (int which = 1; which <= 2; ++which) {
            Pc.create(which == 1 ? n : 0);
            for (int c = 0; c < 3; ++c)
                cudaMemcpy(Pc.component(c), pc->P.component(c), sizeof(double) * n, cudaMemcpyDeviceToDevice);
            ippl::b200::check(ipplb_penning_kick(ippl::b200::ctx(), which, &push, (long)n, pc->R.component(0), pc->R.component(1),
                                                 pc->R.component(2), Pc.component(0), Pc.component(1), Pc.component(2),
                                                 pc->E.component(0), pc->E.component(1), pc->E.component(2)),
                              "ipplb_penning_kick");
            if (which == 1) {
                auto Rview = pc->R.getView();
                auto Pview = pc->P.getView();
                auto Eview = pc->E.getView();
                Kokkos::parallel_for(
                    "Kick1", pc->getLocalNum(), KOKKOS_LAMBDA(const size_t j) {
#include "penning_kick1.inc"
                    });
                Kokkos::fence();
            } else {
                auto R2view = pc->R.getView();
                auto P2view = pc->P.getView();
                auto E2view = pc->E.getView();
                Kokkos::parallel_for(
                    "Kick2", pc->getLocalNum(), KOKKOS_LAMBDA(const size_t j) {
#include "penning_kick2.inc"
                    });
                Kokkos::fence();
            }
            std::vector<ippl::Vector<double, 3>> a(n), b(n);
            pc->P.copyToHost(a);
            Pc.copyToHost(b);
            long differ = 0;
            double worst = 0.0;
            for (size_type i = 0; i < n; ++i)
                for (unsigned d = 0; d < Dim; ++d)
                    if (a[i][d] != b[i][d]) {
                        ++differ;
                        worst = std::max(worst, std::fabs(a[i][d] - b[i][d]) / std::max(std::fabs(b[i][d]), 1e-300));
                    }
            std::cout << "  Kick" << which << ": " << differ << " of " << 3 * n << " components differ from ipplb_penning_kick, worst relative "
                      << worst << (differ == 0 ? " (bit for bit)" : "") << std::endl;
            rc |= check(which == 1 ? "Kick1 lambda vs ipplb_penning_kick(1)" : "Kick2 lambda vs ipplb_penning_kick(2)", worst <= 4.5e-16);
        }

        // ---- LandauDampingManager::dumpLandau(Eview): "Ex stats" over the field -------------------------------------------
        {
            const ipplb_mesh& m = fc->getE().b200_mesh();
            const long cells    = (long)(m.nl[0] + 2) * (m.nl[1] + 2) * (m.nl[2] + 2);
            std::vector<double> hf(3 * cells);
            for (auto& v : hf) v = nd(gen);
            cudaMemcpy(fc->getE().data(), hf.data(), sizeof(double) * hf.size(), cudaMemcpyHostToDevice);
            double want[2];
            ippl::b200::check(ipplb_field_ex_stats(ippl::b200::ctx(), &m, fc->getE().data(), want), "ipplb_field_ex_stats");

            auto Eview        = fc->getE().getView();
            const int nghostE = fc->getE().getNghost();
            using index_array_type = typename ippl::RangePolicy<Dim>::index_array_type;
            double localEx2 = 0, localExNorm = 0;
            ippl::parallel_reduce(
                "Ex stats", ippl::getRangePolicy(Eview, nghostE),
                KOKKOS_LAMBDA(const index_array_type& args, double& E2, double& ENorm) {
#include "landau_ex_stats_field.inc"
                },
                Kokkos::Sum<double>(localEx2), Kokkos::Max<double>(localExNorm));
            rc |= check("\"Ex stats\" (field) lambda == ipplb_field_ex_stats",
                        std::fabs(localEx2 - want[0]) <= 1e-12 * want[0] && localExNorm == want[1]);
        }
        // ---- LandauDampingManager::dumpLandau(): "Ex stats" over the particles ----------------------------------------------
        {
            auto Eview               = pc->E.getView();
            size_type localParticles = pc->getLocalNum();
            using exec_space         = Kokkos::DefaultExecutionSpace;
            using policy_type        = Kokkos::RangePolicy<exec_space>;
            policy_type iteration_policy(0, localParticles);
            double localEx2 = 0;
            Kokkos::parallel_reduce(
                "Ex stats", iteration_policy,
                KOKKOS_LAMBDA(const size_t i, double& E2) {
#include "landau_ex_stats_particles.inc"
                },
                Kokkos::Sum<double>(localEx2));
            double want = 0;
            for (size_type i = 0; i < n; ++i) want += hE[i][0] * hE[i][0];
            rc |= check("\"Ex stats\" (particles) lambda == host sum", std::fabs(localEx2 - want) <= 1e-12 * want);
        }
        // ---- PenningTrapManager::dumpData: "Particle Kinetic Energy" and "Vector E reduce" ----------------------------------------
        {
            auto Pview       = pc->P.getView();
            double kinEnergy = 0.0;
            Kokkos::parallel_reduce(
                "Particle Kinetic Energy", pc->getLocalNum(),
                KOKKOS_LAMBDA(const int i, double& valL) {
#include "penning_kinetic.inc"
                },
                Kokkos::Sum<double>(kinEnergy));
            double want = 0.0;
            ippl::b200::check(ipplb_particles_kinetic(ippl::b200::ctx(), (long)n, pc->P.component(0), pc->P.component(1),
                                                      pc->P.component(2), &want), "ipplb_particles_kinetic");
            rc |= check("\"Particle Kinetic Energy\" lambda == ipplb_particles_kinetic", std::fabs(kinEnergy - want) <= 1e-12 * want);

            const ipplb_mesh& m = fc->getE().b200_mesh();
            double st[7];
            ippl::b200::check(ipplb_field_energy_stats(ippl::b200::ctx(), &m, fc->getE().data(), st), "ipplb_field_energy_stats");
            const int nghostE = fc->getE().getNghost();
            auto Eview        = fc->getE().getView();
            using index_array_type = typename ippl::RangePolicy<Dim>::index_array_type;
            bool ok = true;
            for (unsigned d = 0; d < Dim; ++d) {
                T temp = 0.0;
                ippl::parallel_reduce(
                    "Vector E reduce", ippl::getRangePolicy(Eview, nghostE),
                    KOKKOS_LAMBDA(const index_array_type& args, T& valL) {
#include "penning_vector_e.inc"
                    },
                    Kokkos::Sum<T>(temp));
                Kokkos::fence();
                ok &= std::fabs(temp - st[d]) <= 1e-12 * st[d];
            }
            rc |= check("\"Vector E reduce\" lambda == ipplb_field_energy_stats (sum E_d^2)", ok);

            // ---- BumponTailInstabilityManager::dumpBumponTailInstability: "Ex inner product" and "Ex max norm" ---------------
            double temp = 0.0;
            ippl::parallel_reduce(
                "Ex inner product", ippl::getRangePolicy(Eview, nghostE),
                KOKKOS_LAMBDA(const index_array_type& args, double& valL) {
#include "bumpontail_inner.inc"
                },
                Kokkos::Sum<double>(temp));
            double tempMax = 0.0;
            ippl::parallel_reduce(
                "Ex max norm", ippl::getRangePolicy(Eview, nghostE),
                KOKKOS_LAMBDA(const index_array_type& args, double& valL) {
#include "bumpontail_max.inc"
                },
                Kokkos::Max<double>(tempMax));
            rc |= check("\"Ex inner product\" / \"Ex max norm\" lambdas == ipplb_field_energy_stats (E_z)",
                        std::fabs(temp - st[2]) <= 1e-12 * st[2] && tempMax == st[5]);
        }
    }
    ippl::finalize();
    std::cout << TestName << (rc ? ": FAILED" : ": all reference lambdas agree with the C-ABI kernels") << std::endl;
    return rc;
}
