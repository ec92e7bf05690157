// Alpine.h -- what the three mini-apps share, restated on the B200 facade (include/ippl/Ippl.h) from the reference's
// demos/alpine/{ParticleContainer.hpp, FieldContainer.hpp, AlpineManager.h}: the particle / field containers and the
// manager base with par2grid (scatterCIC + charge conservation check + getDensity), grid2par (gatherCIC), the run
// loop, and -- B200 specific -- the fused single-pass step (ipplb_bins_step) any of the apps can switch to with
// --fused.  Each driver defines `Dim`, `T` and `TestName` before including this header, like the reference's .cpp files.
#pragma once
#include <algorithm>
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
    // multi-rank exchange of R, P, q (E is recomputed by the next gather).  Two collective halves, like the reference,
    // which grows the attributes on receive (src/Particle/ParticleBase.hpp:300-393): the plan exchanges the counts, the
    // attributes are grown to hold the arrivals (times the over-allocation factor), the commit moves the particles.
    void migrate() override {
        auto bundle = [&]() {
            ipplb_particles b{};
            b.x = this->R.component(0); b.y = this->R.component(1); b.z = this->R.component(2);
            b.px = P.component(0); b.py = P.component(1); b.pz = P.component(2);
            b.q = q.component(0);
            b.n = (long)this->getLocalNum();
            b.capacity = (long)std::min(std::min(this->R.size(), P.size()), q.size());
            return b;
        };
        ipplb_particles b = bundle();
        long n_after = b.n;
        const int rc = ipplb_update_plan(ippl::b200::ctx(), &b, &n_after, nullptr, nullptr);
        if (rc != IPPLB_OK && rc != IPPLB_ERR_CAPACITY) ippl::b200::check(rc, "ParticleContainer::migrate (plan)");
        if (n_after > b.capacity) {
            const std::size_t want = (std::size_t)n_after * (std::size_t)std::max(1, (int)ippl::Comm->getDefaultOverallocation());
            this->R.reserve(want);
            P.reserve(want);
            q.reserve(want);
            b = bundle();
        }
        ippl::b200::check(ipplb_update_commit(ippl::b200::ctx(), &b), "ParticleContainer::migrate (commit)");
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


// demos/alpine/LoadBalancer.hpp
template <typename T_, unsigned D>
class LoadBalancer {
    using Solver_t = ippl::FFTPeriodicPoissonSolver<VField_t<T_, D>, Field_t<D>>;
    using ORB      = ippl::OrthogonalRecursiveBisection<Field_t<D>, T_>;

public:
    LoadBalancer(double lbs, std::shared_ptr<FieldContainer<T_, D>>& fc, std::shared_ptr<ParticleContainer<T_, D>>& pc,
                 std::shared_ptr<Solver_t>& fs)
        : loadbalancethreshold_m(lbs), rho_m(&fc->getRho()), E_m(&fc->getE()), pc_m(pc), fs_m(fs) {}
    double getLoadBalanceThreshold() const { return loadbalancethreshold_m; }
    void setLoadBalanceFreq(unsigned f) { loadbalancefreq_m = f; }

    // :54-88
    void updateLayout(FieldLayout_t<D>* fl, Mesh_t<D>* mesh, bool& isFirstRepartition) {
        static IpplTimings::TimerRef tupdateLayout = IpplTimings::getTimer("updateLayout");
        IpplTimings::startTimer(tupdateLayout);
        (*E_m).updateLayout(*fl);
        (*rho_m).updateLayout(*fl);
        pc_m->getLayout().updateLayout(*fl, *mesh);
        IpplTimings::stopTimer(tupdateLayout);
        static IpplTimings::TimerRef tupdatePLayout = IpplTimings::getTimer("updatePB");
        IpplTimings::startTimer(tupdatePLayout);
        if (!isFirstRepartition) pc_m->update();
        IpplTimings::stopTimer(tupdatePLayout);
    }
    void initializeORB(FieldLayout_t<D>* fl, Mesh_t<D>* mesh) { orb.initialize(*fl, *mesh, *rho_m); }
    // :94-123
    bool repartition(FieldLayout_t<D>* fl, Mesh_t<D>* mesh, bool& isFirstRepartition) {
        bool res = orb.binaryRepartition(pc_m->R, *fl, isFirstRepartition);
        if (res != true) {
            std::cout << "Could not repartition!" << std::endl;
            return false;
        }
        this->updateLayout(fl, mesh, isFirstRepartition);
        fs_m->setRhs(*rho_m);
        return true;
    }
    // :125-151 (the fixed-frequency branch is the reference's UniformPlasmaTest rule, selected here with --lb-every)
    bool balance(size_type totalP, const unsigned int nstep) {
        if (ippl::Comm->size() < 2 || loadbalancethreshold_m == 1.0) return false;
        if (loadbalancefreq_m) return (nstep % loadbalancefreq_m == 0);
        double equalPart = (double)totalP / ippl::Comm->size();
        double dev       = std::abs((double)pc_m->getLocalNum() - equalPart) / totalP;
        double local     = dev > loadbalancethreshold_m ? 1.0 : 0.0, any = 0.0;
        ippl::Comm->reduce(local, any, 1, std::greater<double>());  // MPI_Allgather + any-of in the reference
        return any > 0.0;
    }

private:
    double loadbalancethreshold_m;
    Field_t<D>* rho_m;
    VField_t<T_, D>* E_m;
    std::shared_ptr<ParticleContainer<T_, D>> pc_m;
    std::shared_ptr<Solver_t> fs_m;
    unsigned int loadbalancefreq_m = 0;
    ORB orb;
};

// demos/alpine/AlpineManager.h
template <typename T_, unsigned D>
class AlpineManager {
public:
    using ParticleContainer_t = ParticleContainer<T_, D>;
    using FieldContainer_t    = FieldContainer<T_, D>;
    using Solver_t            = ippl::FFTPeriodicPoissonSolver<VField_t<T_, D>, Field_t<D>>;

    AlpineManager(size_type totalP, int nt, Vector_t<int, D>& nr, double lbt, std::string solver, std::string stepMethod,
                  bool fused)
        : totalP_m(totalP), nt_m(nt), nr_m(nr), lbt_m(lbt), solver_m(solver), stepMethod_m(stepMethod), fused_m(fused) {}
    virtual ~AlpineManager() { releaseBins(); }
    void releaseBins() {
        if (bins_m) ipplb_bins_destroy(bins_m);
        bins_m = nullptr;
        for (auto*& p : spare_m) {
            if (p) cudaFree(p);
            p = nullptr;
        }
        if (exit_buf_m) cudaFree(exit_buf_m);
        exit_buf_m = nullptr;
    }
    void setLoadBalanceFreq(unsigned f) { lbfreq_m = f; }
    int repartitions() const { return repartitions_m; }
    int getNt() const { return nt_m; }
    void setTime(double t) { time_m = t; }
    virtual void pre_run() = 0;
    virtual void advance() = 0;
    virtual void dump()    = 0;

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

protected:
    // containers + solver + the two solves / scatter / gather / dump every pre_run ends with
    void setupContainers() {
        if (solver_m != "FFT") throw IpplException(TestName, "only the FFT solver is wired to the facade");
        for (unsigned i = 0; i < D; i++) domain_m[i] = ippl::Index(nr_m[i]);
        decomp_m.fill(true);
        fcontainer_m = std::make_shared<FieldContainer_t>(hr_m, rmin_m, rmax_m, decomp_m, domain_m, origin_m, true);
        pcontainer_m = std::make_shared<ParticleContainer_t>(fcontainer_m->getMesh(), fcontainer_m->getFL());
        fcontainer_m->initializeFields();
        fsolver_m = std::make_shared<Solver_t>(fcontainer_m->getE(), fcontainer_m->getRho());
        loadbalancer_m = std::make_shared<LoadBalancer<T_, D>>(lbt_m, fcontainer_m, pcontainer_m, fsolver_m);
        loadbalancer_m->setLoadBalanceFreq(lbfreq_m);
    }
    // the first repartition of initializeParticles (LandauDampingManager.h:179-205): ORB on the analytic density
    template <class DistT>
    void firstRepartition(const DistT& distR) {
        if ((lbt_m == 1.0) || (ippl::Comm->size() < 2)) return;
        Inform m("Initialize Particles");
        m << "Starting first repartition" << endl;
        static IpplTimings::TimerRef domainDecomposition = IpplTimings::getTimer("loadBalance");
        IpplTimings::startTimer(domainDecomposition);
        isFirstRepartition_m = true;
        distR.fillFullPdf(fcontainer_m->getRho());
        loadbalancer_m->initializeORB(&fcontainer_m->getFL(), &fcontainer_m->getMesh());
        loadbalancer_m->repartition(&fcontainer_m->getFL(), &fcontainer_m->getMesh(), isFirstRepartition_m);
        IpplTimings::stopTimer(domainDecomposition);
    }
    // the balance check of LeapFrogStep (LandauDampingManager.h:290-298), reference-shaped path
    void maybeRepartition() {
        static IpplTimings::TimerRef domainDecomposition = IpplTimings::getTimer("loadBalance");
        bool isFirstRepartition = false;
        if (loadbalancer_m->balance(totalP_m, it_m + 1)) {
            IpplTimings::startTimer(domainDecomposition);
            loadbalancer_m->repartition(&fcontainer_m->getFL(), &fcontainer_m->getMesh(), isFirstRepartition);
            IpplTimings::stopTimer(domainDecomposition);
            ++repartitions_m;
        }
    }
    void firstSolve() {
        fcontainer_m->getRho() = 0.0;
        fsolver_m->solve();  // warm-up solve on rho = 0 (LandauDampingManager.h:137-141)
        par2grid();
        static IpplTimings::TimerRef SolveTimer = IpplTimings::getTimer("solve");
        IpplTimings::startTimer(SolveTimer);
        fsolver_m->solve();
        IpplTimings::stopTimer(SolveTimer);
        grid2par();
        dump();
    }

    // One step through the fused single-pass kernel: [closing kick of the previous step] + opening kick + drift + BC +
    // re-bucketing + scatter in ONE pass over the particles; E at the particles is never materialised.  The push is
    // the app's (leapfrog or Penning), do_kick2 is cleared on the first step (pre_run did not kick).
    void fusedStep(ipplb_push push) {
        static IpplTimings::TimerRef FTimer     = IpplTimings::getTimer("fusedStep");
        static IpplTimings::TimerRef SolveTimer = IpplTimings::getTimer("solve");
        auto* ctx = ippl::b200::ctx();
        const bool multi = ippl::Comm->size() > 1;
        auto& pc  = *pcontainer_m;
        auto& rho = fcontainer_m->getRho();
        auto& E   = fcontainer_m->getE();
        const long n = (long)pc.getLocalNum();
        if (!bins_m) {  // bucket the particles once; the fused step keeps them bucketed
            const long cap = (multi ? 2 * n : n + n / 4) + 65536;  // migration head-room on several ranks
            ippl::b200::check(ipplb_bins_create(ctx, &rho.b200_mesh(), cap, &bins_m), "bins_create");
            for (int b = 0; b < 2; ++b)
                for (int a = 0; a < 6; ++a) spare_m[6 * b + a] = ippl::b200::device_alloc<double>(cap);
            ipplb_particles in{pc.R.component(0), pc.R.component(1), pc.R.component(2), pc.P.component(0), pc.P.component(1),
                               pc.P.component(2), nullptr, Q_m / totalP_m, n, (long)pc.R.size()};
            cur_m = bundle(0, cap);
            nxt_m = bundle(1, cap);
            ippl::b200::check(ipplb_bins_build(ctx, bins_m, &in, &cur_m), "bins_build");
            if (multi) {  // leavers of a step: records per destination rank; this rank's physical region
                exit_cap_m = (int)std::max<long>(n / 4, 1 << 16);
                exit_buf_m = ippl::b200::device_alloc<double>(6 * (std::size_t)exit_cap_m);
                ippl::detail::RegionLayout<double, D, Mesh_t<D>> rl(fcontainer_m->getFL(), fcontainer_m->getMesh());
                for (int d = 0; d < 3; ++d) {
                    region_m[d]     = rl.regions()[6 * ippl::Comm->rank() + d];
                    region_m[3 + d] = rl.regions()[6 * ippl::Comm->rank() + 3 + d];
                }
            }
        }
        push.do_kick2 = it_m > 0;
        push.do_kick1 = push.do_drift = push.do_bc = 1;
        IpplTimings::startTimer(FTimer);
        E.fillHalo();
        rho = 0.0;
        ippl::b200::check(ipplb_bins_step(ctx, bins_m, &push, &cur_m, &nxt_m, E.data(), rho.data(), exit_buf_m, exit_cap_m,
                                          multi ? region_m : nullptr, multi ? region_m + 3 : nullptr),
                          "bins_step");
        std::swap(cur_m, nxt_m);
        if (multi) {  // ParticleSpatialLayout::update for the bucketed store: exchange, append, deposit the arrivals
            ippl::b200::check(ipplb_bins_migrate(ctx, bins_m, &cur_m, exit_buf_m, exit_cap_m, rho.data(), nullptr, nullptr), "bins_migrate");
            pc.setLocalNum((size_type)cur_m.n);
        }
        if (loadbalancer_m->balance(totalP_m, it_m + 1)) {
            // LoadBalancer::repartition for the bucketed store: back to the contiguous attribute arrays, new layout,
            // pc->update() to the new owners, rho re-deposited on the new boxes; the next step re-buckets.  (The closing
            // kick of this step is still pending in P, exactly as it is inside the buckets.)
            IpplTimings::stopTimer(FTimer);
            static IpplTimings::TimerRef domainDecomposition = IpplTimings::getTimer("loadBalance");
            IpplTimings::startTimer(domainDecomposition);
            pc.setLocalNum((size_type)cur_m.n);  // grows R, P, E, q if the rank gained particles
            ipplb_particles out{pc.R.component(0), pc.R.component(1), pc.R.component(2), pc.P.component(0), pc.P.component(1),
                                pc.P.component(2), nullptr, Q_m / totalP_m, 0, (long)pc.R.size()};
            ippl::b200::check(ipplb_bins_compact(ctx, bins_m, &cur_m, &out), "bins_compact");
            pc.q = Q_m / totalP_m;
            releaseBins();
            bool isFirstRepartition = false;
            loadbalancer_m->repartition(&fcontainer_m->getFL(), &fcontainer_m->getMesh(), isFirstRepartition);
            ++repartitions_m;
            IpplTimings::stopTimer(domainDecomposition);
            par2grid();
        } else {
            rho.accumulateHalo();
            IpplTimings::stopTimer(FTimer);
            finishScatter();
        }
        IpplTimings::startTimer(SolveTimer);
        fsolver_m->solve();
        IpplTimings::stopTimer(SolveTimer);
    }
    // sum_i dot(P_i, P_i) wherever the particles currently live (attribute arrays or the bucketed store)
    double sumP2() {
        double s = 0.0;
        if (bins_m)
            ippl::b200::check(ipplb_bins_kinetic(ippl::b200::ctx(), bins_m, &cur_m, &s), "bins_kinetic");
        else
            ippl::b200::check(ipplb_particles_kinetic(ippl::b200::ctx(), (long)pcontainer_m->getLocalNum(), pcontainer_m->P.component(0),
                                                      pcontainer_m->P.component(1), pcontainer_m->P.component(2), &s),
                              "particles_kinetic");
        return s;
    }
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
    double Q_m = 0, dt_m = 0, time_m = 0;
    int it_m = 0;
    Vector_t<double, D> kw_m, rmin_m, rmax_m, hr_m, origin_m;
    ippl::NDIndex<D> domain_m;
    std::array<bool, D> decomp_m;
    std::shared_ptr<FieldContainer_t> fcontainer_m;
    std::shared_ptr<ParticleContainer_t> pcontainer_m;
    std::shared_ptr<Solver_t> fsolver_m;
    ipplb_bins* bins_m = nullptr;
    ipplb_particles cur_m{}, nxt_m{};
    std::array<double*, 12> spare_m{};
    double* exit_buf_m = nullptr;
    int exit_cap_m     = 0;
    std::shared_ptr<LoadBalancer<T_, D>> loadbalancer_m;
    bool isFirstRepartition_m = false;
    unsigned lbfreq_m         = 0;
    int repartitions_m        = 0;
    double region_m[6] = {0, 0, 0, 0, 0, 0};
};

// the reference's main() of the three drivers (demos/alpine/LandauDamping.cpp:38-95): same positional arguments
template <class Manager>
int alpine_main(int argc, char* argv[]) {
    try {
        ippl::initialize(argc, argv);
    } catch (const std::exception& ex) {
        // no CUDA device / driver: the library has no CPU fallback, and neither has the driver
        std::cerr << TestName << ": cannot start: " << ex.what() << " (a CUDA device is required; there is no CPU fallback)" << std::endl;
        return 2;
    }
    int exit_code = 0;
    {
        try {
            Inform msg(TestName);
            static IpplTimings::TimerRef mainTimer = IpplTimings::getTimer("total");
            IpplTimings::startTimer(mainTimer);
            if (argc < 9) throw IpplException(TestName, "usage: <nx> <ny> <nz> <Np> <Nt> FFT <lbthres> LeapFrog [--overallocate f] [--info n] [--fused]");
            int arg = 1;
            Vector_t<int, Dim> nr;
            for (unsigned d = 0; d < Dim; d++) nr[d] = std::atoi(argv[arg++]);
            size_type totalP        = std::atoll(argv[arg++]);
            int nt                  = std::atoi(argv[arg++]);
            std::string solver      = argv[arg++];
            double lbt              = std::atof(argv[arg++]);
            std::string step_method = argv[arg++];
            bool fused              = false;
            unsigned lbevery        = 0;
            for (int i = arg; i < argc; ++i) {
                fused |= std::string(argv[i]) == "--fused";
                if (std::string(argv[i]) == "--lb-every" && i + 1 < argc) lbevery = (unsigned)std::atoi(argv[i + 1]);
            }
            Manager manager(totalP, nt, nr, lbt, solver, step_method, fused);
            manager.setLoadBalanceFreq(lbevery);
            manager.pre_run();
            manager.setTime(0.0);
            msg << "Starting iterations ..." << endl;
            manager.run(manager.getNt());
            msg << "End." << endl;
            if (ippl::Comm->rank() == 0) std::cout << "ORB repartitions during the run: " << manager.repartitions() << std::endl;
            IpplTimings::stopTimer(mainTimer);
            IpplTimings::print();
            IpplTimings::print(std::string("timing.dat"));   // demos/alpine/LandauDamping.cpp:102-103
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
