// fusion_check: the facade's lazy-fusion engine (include/ippl/Ippl.h, detail::FusionEngine) against the plain call sequence.
// The same leapfrog loop -- written as an unchanged driver writes it (LandauDampingManager.h:265-320: attribute expressions,
// update(), scatter, solve, gather) -- runs twice from the same initial condition: once with every call executed as it comes,
// once with fusion on, where [gather, kick, kick, drift, update, scatter] becomes one ipplb_bins_step.  In the fused run the
// loop also PEEKS at the particles at awkward moments (after the closing kick, between drift and update, right after a
// fused scatter), which forces the engine to bring the particles back from the bucketed store and replay what it had
// recorded.  Both runs must end with the same particles (compared as sorted coordinate lists) and the same field energy
// history.  Built by `make -C demos`; needs a CUDA device (no CPU fallback); the CPU suite runs it against oracle/mock.
constexpr unsigned Dim = 3;
using T = double;
const char* TestName = "fusion_check";
#include "Alpine.h"

#include <algorithm>
#include <random>

struct Result {
    std::vector<double> x, px;        // sorted
    std::vector<double> energy, psum;  // per step: sum E_x^2 on the mesh, sum of P_x where it was peeked
    long fused = 0, materialised = 0;
};

static Result run(bool fuse, int nsteps, int n, int grid, double kick_factor = 0.5, bool peek = true) {
    ippl::b200::fusion_enabled() = fuse;
    const long fused0 = ippl::b200::fusion_stats().fused_steps, mat0 = ippl::b200::fusion_stats().materialised;
    Result out;
    {
        ippl::NDIndex<3> domain;
        for (unsigned d = 0; d < 3; ++d) domain[d] = ippl::Index(grid);
        const double L = 4 * M_PI;
        Vector_t<double, 3> hr(L / grid), origin(0.0), rmin(0.0), rmax(L);
        std::array<bool, 3> decomp{true, true, true};
        FieldContainer<double, 3> fc(hr, rmin, rmax, decomp, domain, origin, true);
        fc.initializeFields();
        ParticleContainer<double, 3> pc(fc.getMesh(), fc.getFL());
        pc.create(n);
        const double Q = -L * L * L;
        pc.q           = Q / n;
        std::mt19937_64 gen(7);
        std::uniform_real_distribution<double> U(0.0, L);
        std::normal_distribution<double> G(0.0, 1.0);
        std::vector<ippl::Vector<double, 3>> hR(n), hP(n);
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) {
                hR[i][d] = U(gen) * (d == 0 ? (0.9 + 0.1 * std::cos(0.5 * i)) : 1.0);
                hP[i][d] = G(gen);
            }
        pc.R.copyFromHost(hR);
        pc.P.copyFromHost(hP);
        auto& rho = fc.getRho();
        auto& E   = fc.getE();
        ippl::FFTPeriodicPoissonSolver<VField_t<double, 3>, Field_t<3>> solver(E, rho);
        const double dt = 0.05, vol = hr[0] * hr[1] * hr[2];
        auto deposit = [&] {
            rho = 0.0;
            scatter(pc.q, rho, pc.R);
            rho = rho / vol;
            rho = rho - (Q / (L * L * L));
            solver.solve();
        };
        auto energy = [&] {
            double st[7];
            ippl::b200::check(ipplb_field_energy_stats(ippl::b200::ctx(), &E.b200_mesh(), E.data(), st), "energy");
            return st[0];
        };
        deposit();
        gather(pc.E, E, pc.R);
        out.energy.push_back(energy());
        for (int it = 0; it < nsteps; ++it) {
            pc.P = pc.P - kick_factor * dt * pc.E;
            pc.R = pc.R + dt * pc.P;
            if (peek && it == 3) out.psum.push_back(pc.R.sum(1));   // peek between drift and update
            pc.update();
            deposit();
            if (peek && it == 5) out.psum.push_back(pc.P.sum(2));   // peek right after a (fused) scatter: particles are in the store
            gather(pc.E, E, pc.R);
            pc.P = pc.P - kick_factor * dt * pc.E;
            if (peek && it % 4 == 1) out.psum.push_back(pc.P.sum(0));   // peek after the closing kick
            out.energy.push_back(energy());
        }
        std::vector<ippl::Vector<double, 3>> r, p;
        pc.R.copyToHost(r);
        pc.P.copyToHost(p);
        for (int i = 0; i < n; ++i) {
            out.x.push_back(r[i][0]);
            out.px.push_back(p[i][0]);
        }
        std::sort(out.x.begin(), out.x.end());
        std::sort(out.px.begin(), out.px.end());
    }
    out.fused        = ippl::b200::fusion_stats().fused_steps - fused0;
    out.materialised = ippl::b200::fusion_stats().materialised - mat0;
    return out;
}

static double worst(const std::vector<double>& a, const std::vector<double>& b, double floor_) {
    if (a.size() != b.size()) return 1e300;
    double w = 0;
    for (std::size_t i = 0; i < a.size(); ++i) w = std::max(w, std::fabs(a[i] - b[i]) / std::max(std::fabs(a[i]), floor_));
    return w;
}

int main(int argc, char* argv[]) {
    try {
        ippl::initialize(argc, argv);
    } catch (const IpplException& ex) {
        std::cerr << TestName << ": cannot start: " << ex.what() << " (a CUDA device is required; there is no CPU fallback)" << std::endl;
        return 2;
    }
    int rc = 0;
    try {
        const int nsteps = 12, n = 200000, grid = 16;
        const Result plain = run(false, nsteps, n, grid);
        const Result fused = run(true, nsteps, n, grid);
        const double ex = worst(plain.x, fused.x, 1e-3), ep = worst(plain.px, fused.px, 1e-3), ee = worst(plain.energy, fused.energy, 1e-300),
                     es = worst(plain.psum, fused.psum, 1.0);
        std::cout << "fusion_check: " << fused.fused << " fused steps, " << fused.materialised << " materialisations in the fused run ("
                  << plain.fused << " / " << plain.materialised << " in the plain run); worst relative difference: x " << ex << ", px " << ep
                  << ", field energy " << ee << ", peeked sums " << es << std::endl;
        // a peek after the closing kick (it = 1, 5, 9) computes that step's gather for real, so the NEXT step cannot fuse; the
        // peek between drift and update (it = 3) breaks that step's record: steps 2, 3, 6, 10 run unfused, the other 8 fused.
        // Materialisations: it = 1, 3, 5 (twice: out of the store after the scatter, then the replay after the kick), 9, and
        // the final copy to the host.
        rc |= plain.fused != 0 || plain.materialised != 0 || fused.fused != 8 || fused.materialised != 6;
        rc |= !(ex <= 1e-9 && ep <= 1e-9 && ee <= 1e-9 && es <= 1e-9);
        // a sequence that LOOKS like the leapfrog step but is not (kicks of 0.3 dt): recorded, refused at the scatter, replayed
        // through the ordinary kernels -- nothing fuses, nothing changes
        const Result odd_plain = run(false, 4, n, grid, 0.3, false), odd_fused = run(true, 4, n, grid, 0.3, false);
        const double ox = worst(odd_plain.x, odd_fused.x, 1e-3), oe = worst(odd_plain.energy, odd_fused.energy, 1e-300);
        std::cout << "fusion_check: non-leapfrog coefficients: " << odd_fused.fused << " fused steps, " << odd_fused.materialised
                  << " materialisations; worst relative difference: x " << ox << ", field energy " << oe << std::endl;
        rc |= odd_fused.fused != 0 || odd_fused.materialised != 5 || !(ox <= 1e-9 && oe <= 1e-9);
        // and the clean case: no peeks, every step fused, one materialisation (the final copy to the host)
        const Result clean = run(true, 6, n, grid, 0.5, false), clean_plain = run(false, 6, n, grid, 0.5, false);
        rc |= clean.fused != 6 || clean.materialised != 1 || !(worst(clean_plain.x, clean.x, 1e-3) <= 1e-9);
        std::cout << (rc ? "fusion_check: FAILED" : "fusion_check: ok") << std::endl;
    } catch (const IpplException& ex) {
        std::cerr << "fusion_check: IPPL exception: " << ex.what() << std::endl;
        rc = 1;
    }
    ippl::b200::fusion_enabled() = false;
    ippl::finalize();
    return rc;
}
