// LandauDamping on the B200 facade -- the reference's mini-app (demos/alpine/LandauDamping.cpp +
// LandauDampingManager.h + AlpineManager.h) restated against include/ippl/Ippl.h: same command line, same
// pre_run / LeapFrogStep / dump structure, same CSV (data/FieldLandau_<ranks>_manager.csv).
//
//   LandauDamping <nx> <ny> <nz> <Np> <Nt> FFT <lbthres> LeapFrog [--overallocate f] [--info n] [--fused]
//
// --fused runs the step through the single-pass fused kernel (ipplb_bins_step: gather + kick + kick + drift + BC
// + re-bucketing + scatter in one pass) instead of the reference-shaped sequence of attribute expressions;
// both paths produce the same energies (tests/test_facade.py).
// Particle initialisation is host-side (inverse-CDF sampling of 1 + alpha cos(k x) by Newton iterations and
// Box-Muller normals, seed 42 + 100 * rank like LandauDampingManager.h:216-240); the reference's
// Kokkos::Random_XorShift64_Pool stream is backend-dependent, so the numbers are statistically, not bitwise,
// comparable with a reference run -- which is why the reference checks its own CSV at tolerance 0.4.
constexpr unsigned Dim = 3;
using T                = double;
const char* TestName   = "LandauDamping";

#include "ippl/Ippl.h"

#include <filesystem>
#include <fstream>
#include <random>

template <unsigned D>
using Mesh_t = ippl::UniformCartesian<double, D>;
template <typename T_, unsigned D>
using PLayout_t = ippl::ParticleSpatialLayout<T_, D, Mesh_t<D>>;
template <unsigned D>
using FieldLayout_t = ippl::FieldLayout<D>;
template <typename T_, unsigned D>
using Vector_t = ippl::Vector<T_, D>;
template <unsigned D>
using Field_t = ippl::Field<double, D, Mesh_t<D>, typename Mesh_t<D>::DefaultCentering>;
template <typename T_, unsigned D>
using VField_t = ippl::Field<Vector_t<T_, D>, D, Mesh_t<D>, typename Mesh_t<D>::DefaultCentering>;
using size_type = ippl::detail::size_type;

// demos/alpine/ParticleContainer.hpp
template <typename T_, unsigned D = 3>
class ParticleContainer : public ippl::ParticleBase<PLayout_t<T_, D>> {
    using Base = ippl::ParticleBase<PLayout_t<T_, D>>;

public:
    ippl::ParticleAttrib<double> q;           // charge
    typename Base::particle_position_type P;  // particle velocity
    typename Base::particle_position_type E;  // electric field at particle position
    ParticleContainer(Mesh_t<D>& mesh, FieldLayout_t<D>& FL) : pl_m(FL, mesh) {
        this->initialize(pl_m);
        P.set_name("velocity");
        q.set_name("charge");
        E.set_name("electric_field");
        this->addAttribute(q);
        this->addAttribute(P);
        this->addAttribute(E);
        this->setParticleBC(ippl::BC::PERIODIC);
    }
    // multi-rank exchange of R, P, q (E is recomputed by the next gather)
    void migrate() override {
        ipplb_particles b{};
        b.x = this->R.component(0); b.y = this->R.component(1); b.z = this->R.component(2);
        b.px = P.component(0); b.py = P.component(1); b.pz = P.component(2);
        b.q = q.component(0);
        b.n = (long)this->getLocalNum();
        b.capacity = (long)this->R.size();
        ippl::b200::check(ipplb_update(ippl::b200::ctx(), &b, nullptr, nullptr), "ParticleContainer::migrate");
        this->setLocalNum((size_type)b.n);
    }

private:
    PLayout_t<T_, D> pl_m;
};

// demos/alpine/FieldContainer.hpp
template <typename T_, unsigned D = 3>
class FieldContainer {
public:
    FieldContainer(Vector_t<T_, D>& hr, Vector_t<T_, D>& rmin, Vector_t<T_, D>& rmax, std::array<bool, D> decomp,
                   ippl::NDIndex<D> domain, Vector_t<T_, D> origin, bool isAllPeriodic)
        : hr_m(hr), rmin_m(rmin), rmax_m(rmax), mesh_m(domain, hr, origin), fl_m(0, domain, decomp, isAllPeriodic) {}
    void initializeFields() {
        E_m.initialize(mesh_m, fl_m);
        rho_m.initialize(mesh_m, fl_m);
    }
    VField_t<T_, D>& getE() { return E_m; }
    Field_t<D>& getRho() { return rho_m; }
    Vector_t<double, D>& getHr() { return hr_m; }
    Mesh_t<D>& getMesh() { return mesh_m; }
    FieldLayout_t<D>& getFL() { return fl_m; }

private:
    Vector_t<double, D> hr_m, rmin_m, rmax_m;
    VField_t<T_, D> E_m;
    Field_t<D> rho_m;
    Mesh_t<D> mesh_m;
    FieldLayout_t<D> fl_m;
};

// demos/alpine/AlpineManager.h + LandauDampingManager.h
template <typename T_, unsigned D>
class LandauDampingManager {
    using ParticleContainer_t = ParticleContainer<T_, D>;
    using FieldContainer_t    = FieldContainer<T_, D>;
    using Solver_t            = ippl::FFTPeriodicPoissonSolver<VField_t<T_, D>, Field_t<D>>;

public:
    LandauDampingManager(size_type totalP, int nt, Vector_t<int, D>& nr, double lbt, std::string solver,
                         std::string stepMethod, bool fused)
        : totalP_m(totalP), nt_m(nt), nr_m(nr), lbt_m(lbt), solver_m(solver), stepMethod_m(stepMethod), fused_m(fused) {}
    ~LandauDampingManager() {
        if (bins_m) ipplb_bins_destroy(bins_m);
        for (auto* p : spare_m)
            if (p) cudaFree(p);
    }
    int getNt() const { return nt_m; }
    void setTime(double t) { time_m = t; }

    void pre_run() {
        Inform m("Pre Run");
        const double pi = std::acos(-1.0);
        if (solver_m != "FFT") throw IpplException("LandauDamping", "only the FFT solver is wired to the facade");
        for (unsigned i = 0; i < D; i++) domain_m[i] = ippl::Index(nr_m[i]);
        decomp_m.fill(true);
        kw_m    = 0.5;
        alpha_m = 0.05;
        rmin_m  = 0.0;
        rmax_m  = 2 * pi / kw_m;
        for (unsigned d = 0; d < D; ++d) hr_m[d] = rmax_m[d] / nr_m[d];
        // Q = -\int\int f dx dv
        Q_m      = std::accumulate(rmax_m.begin(), rmax_m.end(), -1., std::multiplies<double>());
        origin_m = rmin_m;
        dt_m     = std::min(.05, 0.5 * *std::min_element(hr_m.begin(), hr_m.end()));
        it_m     = 0;
        time_m   = 0.0;
        m << "Discretization:" << endl << "nt " << nt_m << " Np= " << totalP_m << " grid = " << nr_m << endl;
        fcontainer_m = std::make_shared<FieldContainer_t>(hr_m, rmin_m, rmax_m, decomp_m, domain_m, origin_m, true);
        pcontainer_m = std::make_shared<ParticleContainer_t>(fcontainer_m->getMesh(), fcontainer_m->getFL());
        fcontainer_m->initializeFields();
        fsolver_m = std::make_shared<Solver_t>(fcontainer_m->getE(), fcontainer_m->getRho());
        initializeParticles();
        fcontainer_m->getRho() = 0.0;
        fsolver_m->solve();  // warm-up solve on rho = 0 (LandauDampingManager.h:137-141)
        par2grid();
        static IpplTimings::TimerRef SolveTimer = IpplTimings::getTimer("solve");
        IpplTimings::startTimer(SolveTimer);
        fsolver_m->solve();
        IpplTimings::stopTimer(SolveTimer);
        grid2par();
        dump();
        m << "Done" << endl;
    }

    // inverse-transform sampling of f(x) = 1 + alpha cos(k x) per dimension on this rank's region, v ~ N(0,1)
    void initializeParticles() {
        Inform m("Initialize Particles");
        const auto& ldom       = fcontainer_m->getFL().getLocalNDIndex();
        const int rank         = ippl::Comm->rank();
        size_type nlocal       = totalP_m / ippl::Comm->size();
        if ((size_type)rank < totalP_m % ippl::Comm->size()) ++nlocal;
        pcontainer_m->create(nlocal);
        std::mt19937_64 eng(42 + 100 * rank);
        std::uniform_real_distribution<double> unif(0.0, 1.0);
        std::normal_distribution<double> normal(0.0, 1.0);
        auto Rh = pcontainer_m->R.getHostMirror();
        auto Ph = pcontainer_m->P.getHostMirror();
        auto cdf = [&](double x) { return x + (alpha_m / kw_m) * std::sin(kw_m * x); };
        for (size_type i = 0; i < nlocal; ++i) {
            for (unsigned d = 0; d < D; ++d) {
                const double lo = ldom[d].first() * hr_m[d] + origin_m[d];
                const double hi = (ldom[d].last() + 1) * hr_m[d] + origin_m[d];
                const double u  = cdf(lo) + unif(eng) * (cdf(hi) - cdf(lo));
                double x        = lo + (hi - lo) * 0.5;
                for (int k = 0; k < 20; ++k) {  // Newton, atol 1e-12
                    const double f = cdf(x) - u;
                    if (std::fabs(f) < 1e-12) break;
                    x -= f / (1.0 + alpha_m * std::cos(kw_m * x));
                }
                Rh[i][d] = std::min(std::max(x, std::nextafter(lo, hi)), hi);
                Ph[i][d] = normal(eng);
            }
        }
        ippl::deep_copy(pcontainer_m->R, Rh);
        ippl::deep_copy(pcontainer_m->P, Ph);
        pcontainer_m->q = Q_m / totalP_m;
        m << "particles created and initial conditions assigned " << endl;
    }

    void run(int nt) {
        for (int it = 0; it < nt; ++it) {
            advance();
            time_m += dt_m;
            it_m++;
            dump();
            Inform m("Post-step:");
            m << "Finished time step: " << it_m << " time: " << time_m << endl;
        }
    }
    void advance() {
        if (stepMethod_m != "LeapFrog") throw IpplException(TestName, "Step method is not set/recognized!");
        if (fused_m) FusedLeapFrogStep();
        else LeapFrogStep();
    }

    // LandauDampingManager.h:265-320, verbatim structure
    void LeapFrogStep() {
        static IpplTimings::TimerRef PTimer      = IpplTimings::getTimer("pushVelocity");
        static IpplTimings::TimerRef RTimer      = IpplTimings::getTimer("pushPosition");
        static IpplTimings::TimerRef updateTimer = IpplTimings::getTimer("update");
        static IpplTimings::TimerRef SolveTimer  = IpplTimings::getTimer("solve");
        double dt                               = dt_m;
        std::shared_ptr<ParticleContainer_t> pc = pcontainer_m;
        IpplTimings::startTimer(PTimer);
        pc->P = pc->P - 0.5 * dt * pc->E;
        IpplTimings::stopTimer(PTimer);
        IpplTimings::startTimer(RTimer);
        pc->R = pc->R + dt * pc->P;
        IpplTimings::stopTimer(RTimer);
        IpplTimings::startTimer(updateTimer);
        pc->update();
        IpplTimings::stopTimer(updateTimer);
        par2grid();
        IpplTimings::startTimer(SolveTimer);
        fsolver_m->solve();
        IpplTimings::stopTimer(SolveTimer);
        grid2par();
        IpplTimings::startTimer(PTimer);
        pc->P = pc->P - 0.5 * dt * pc->E;
        IpplTimings::stopTimer(PTimer);
    }

    // the same step through the fused single-pass kernel: [kick2 of the previous step] + kick1 + drift + BC +
    // re-bucketing + scatter in ONE pass; E at the particles is never materialised (single rank)
    void FusedLeapFrogStep() {
        static IpplTimings::TimerRef FTimer     = IpplTimings::getTimer("fusedStep");
        static IpplTimings::TimerRef SolveTimer = IpplTimings::getTimer("solve");
        if (ippl::Comm->size() > 1) throw IpplException(TestName, "the facade's fused step is single-rank in this round");
        auto* ctx = ippl::b200::ctx();
        auto& pc  = *pcontainer_m;
        auto& rho = fcontainer_m->getRho();
        auto& E   = fcontainer_m->getE();
        const long n = (long)pc.getLocalNum();
        if (!bins_m) {  // bucket the particles once; the fused step keeps them bucketed
            const long cap = n + n / 4 + 65536;
            ippl::b200::check(ipplb_bins_create(ctx, &rho.b200_mesh(), cap, &bins_m), "bins_create");
            for (int b = 0; b < 2; ++b)
                for (int a = 0; a < 6; ++a) {
                    spare_m[6 * b + a] = ippl::b200::device_alloc<double>(cap);
                }
            ipplb_particles in{pc.R.component(0), pc.R.component(1), pc.R.component(2), pc.P.component(0), pc.P.component(1),
                               pc.P.component(2), nullptr, Q_m / totalP_m, n, (long)pc.R.size()};
            cur_m = bundle(0, cap);
            nxt_m = bundle(1, cap);
            ippl::b200::check(ipplb_bins_build(ctx, bins_m, &in, &cur_m), "bins_build");
            pending_kick2_m = true;  // pre_run's gather result E(t0) enters through the fused gather instead
        }
        ipplb_push push{};
        push.kind = IPPLB_PUSH_LEAPFROG;
        push.dt   = dt_m;
        // first fused step: only kick1 is outstanding (pre_run did not kick); later steps fold the previous kick2 in
        push.do_kick2 = it_m > 0;
        push.do_kick1 = push.do_drift = push.do_bc = 1;
        IpplTimings::startTimer(FTimer);
        E.fillHalo();
        rho = 0.0;
        ippl::b200::check(ipplb_bins_step(ctx, bins_m, &push, &cur_m, &nxt_m, E.data(), rho.data(), nullptr, 0, nullptr, nullptr),
                          "bins_step");
        std::swap(cur_m, nxt_m);
        rho.accumulateHalo();
        IpplTimings::stopTimer(FTimer);
        finishScatter();
        IpplTimings::startTimer(SolveTimer);
        fsolver_m->solve();
        IpplTimings::stopTimer(SolveTimer);
    }

    void grid2par() { gather(pcontainer_m->E, fcontainer_m->getE(), pcontainer_m->R); }

    // AlpineManager::scatterCIC (AlpineManager.h:157-175)
    void par2grid() {
        fcontainer_m->getRho() = 0.0;
        ippl::ParticleAttrib<double>* q = &pcontainer_m->q;
        auto* R                         = &pcontainer_m->R;
        Field_t<D>* rho                 = &fcontainer_m->getRho();
        scatter(*q, *rho, *R);
        finishScatter();
    }
    void finishScatter() {
        Inform m("scatter ");
        Field_t<D>* rho = &fcontainer_m->getRho();
        double relError = std::fabs((Q_m - (*rho).sum()) / Q_m);
        m << relError << endl;
        // checkChargeConservation (AlpineManager.h:208-223)
        size_type TotalParticles = 0, localParticles = pcontainer_m->getLocalNum();
        ippl::Comm->reduce(localParticles, TotalParticles, 1, std::plus<size_type>());
        if (ippl::Comm->rank() == 0 && (TotalParticles != totalP_m || relError > 1e-10)) {
            m << "Total particles in the sim. " << totalP_m << " after update: " << TotalParticles << endl;
            m << "Rel. error in charge conservation: " << relError << endl;
            ippl::Comm->abort();
        }
        // getDensity (AlpineManager.h:225-245)
        double cellVolume = std::accumulate(hr_m.begin(), hr_m.end(), 1., std::multiplies<double>());
        (*rho)            = (*rho) / cellVolume;
        double size       = 1;
        for (unsigned d = 0; d < D; d++) size *= rmax_m[d] - rmin_m[d];
        *rho = *rho - (Q_m / size);
    }

    // dumpLandau (LandauDampingManager.h:339-386): Ex field energy and max norm over interior cells
    void dump() {
        auto& E = fcontainer_m->getE();
        double st[2];
        ippl::b200::check(ipplb_field_ex_stats(ippl::b200::ctx(), &E.b200_mesh(), E.data(), st), "dump");
        double globaltemp = 0.0, ExAmp = st[1];
        ippl::Comm->reduce(st[0], globaltemp, 1, std::plus<double>());
        double fieldEnergy = std::accumulate(hr_m.begin(), hr_m.end(), globaltemp, std::multiplies<double>());
        if (ippl::Comm->rank() == 0) {
            std::filesystem::create_directory("data");
            std::stringstream fname;
            fname << "data/FieldLandau_" << ippl::Comm->size() << "_manager.csv";
            std::ofstream csvout(fname.str(), std::fabs(time_m) < 1e-14 ? std::ios::trunc : std::ios::app);
            csvout.precision(16);
            csvout.setf(std::ios::scientific, std::ios::floatfield);
            if (std::fabs(time_m) < 1e-14) csvout << "time, Ex_field_energy, Ex_max_norm" << std::endl;
            csvout << time_m << " " << fieldEnergy << " " << ExAmp << std::endl;
        }
        ippl::Comm->barrier();
    }

private:
    ipplb_particles bundle(int b, long cap) {
        double** s = &spare_m[6 * b];
        return ipplb_particles{s[0], s[1], s[2], s[3], s[4], s[5], nullptr, Q_m / totalP_m, 0, cap};
    }
    size_type totalP_m;
    int nt_m;
    Vector_t<int, D> nr_m;
    double lbt_m;
    std::string solver_m, stepMethod_m;
    bool fused_m;
    double kw_m = 0, alpha_m = 0, Q_m = 0, dt_m = 0, time_m = 0;
    int it_m = 0;
    Vector_t<double, D> rmin_m, rmax_m, hr_m, origin_m;
    ippl::NDIndex<D> domain_m;
    std::array<bool, D> decomp_m;
    std::shared_ptr<FieldContainer_t> fcontainer_m;
    std::shared_ptr<ParticleContainer_t> pcontainer_m;
    std::shared_ptr<Solver_t> fsolver_m;
    ipplb_bins* bins_m = nullptr;
    ipplb_particles cur_m{}, nxt_m{};
    std::array<double*, 12> spare_m{};
    bool pending_kick2_m = false;
};

int main(int argc, char* argv[]) {
    ippl::initialize(argc, argv);
    int exit_code = 0;
    {
        try {
            Inform msg(TestName);
            static IpplTimings::TimerRef mainTimer = IpplTimings::getTimer("total");
            IpplTimings::startTimer(mainTimer);
            int arg = 1;
            Vector_t<int, Dim> nr;
            for (unsigned d = 0; d < Dim; d++) nr[d] = std::atoi(argv[arg++]);
            size_type totalP        = std::atoll(argv[arg++]);
            int nt                  = std::atoi(argv[arg++]);
            std::string solver      = argv[arg++];
            double lbt              = std::atof(argv[arg++]);
            std::string step_method = argv[arg++];
            bool fused              = false;
            for (int i = arg; i < argc; ++i) fused |= std::string(argv[i]) == "--fused";
            LandauDampingManager<T, Dim> manager(totalP, nt, nr, lbt, solver, step_method, fused);
            manager.pre_run();
            manager.setTime(0.0);
            msg << "Starting iterations ..." << endl;
            manager.run(manager.getNt());
            msg << "End." << endl;
            IpplTimings::stopTimer(mainTimer);
            IpplTimings::print();
        } catch (const IpplException& ex) {
            Inform err(TestName);
            err << "IPPL exception: " << ex.what() << endl;
            exit_code = 1;
        } catch (const std::exception& ex) {
            Inform err(TestName);
            err << "Unhandled std::exception: " << ex.what() << endl;
            exit_code = 1;
        }
    }
    ippl::finalize();
    return exit_code;
}
