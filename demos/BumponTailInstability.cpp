// BumponTailInstability (and TwoStreamInstability) on the B200 facade -- the reference's mini-app
// (demos/alpine/BumponTailInstability.cpp + BumponTailInstabilityManager.h) restated against include/ippl/Ippl.h:
// same command line, same pre_run / LeapFrogStep / dump structure, same CSV (data/FieldBumponTail_<ranks>_manager.csv).
//
//   BumponTailInstability <nx> <ny> <nz> <Np> <Nt> FFT <lbthres> LeapFrog [--two-stream] [--info n] [--fused]
//
// Positions: uniform in x, y and 1 + delta cos(k z) in z; velocities: bulk N(muBulk, sigma) for the first
// (1 - epsilon) of the particles, beam N(muBeam, sigma) for the rest, along z (BumponTailInstabilityManager.h:110-135,
// 196-303).  The reference selects the two-stream variant with TestName; here a flag does.  EnablePhaseDump is false
// in the reference (:20) and not built.
constexpr unsigned Dim = 3;
using T                = double;
const char* TestName   = "BumponTailInstability";

#include "Alpine.h"

static bool g_two_stream = false;

template <typename T_, unsigned D>
class BumponTailInstabilityManager : public AlpineManager<T_, D> {
    using Base = AlpineManager<T_, D>;

public:
    using Base::Base;

    void pre_run() override {
        Inform m("Pre Run");
        const double pi = std::acos(-1.0);
        if (!g_two_stream) {  // BumponTailInstabilityManager.h:109-116
            this->kw_m = 0.21;
            sigma_m    = 1.0 / std::sqrt(2.0);
            epsilon_m  = 0.1;
            muBulk_m   = 0.0;
            muBeam_m   = 4.0;
            delta_m    = 0.01;
        } else {  // :117-125
            this->kw_m = 0.5;
            sigma_m    = 0.1;
            epsilon_m  = 0.5;
            muBulk_m   = -pi / 2.0;
            muBeam_m   = pi / 2.0;
            delta_m    = 0.01;
        }
        this->rmin_m = 0.0;
        this->rmax_m = 2 * pi / this->kw_m;
        for (unsigned d = 0; d < D; ++d) this->hr_m[d] = this->rmax_m[d] / this->nr_m[d];
        this->Q_m      = std::accumulate(this->rmax_m.begin(), this->rmax_m.end(), -1., std::multiplies<double>());
        this->origin_m = this->rmin_m;
        this->dt_m     = std::min(.05, 0.5 * *std::min_element(this->hr_m.begin(), this->hr_m.end()));
        this->it_m     = 0;
        this->time_m   = 0.0;
        m << "Discretization:" << endl << "nt " << this->nt_m << " Np= " << this->totalP_m << " grid = " << this->nr_m << endl;
        this->setupContainers();
        initializeParticles();
        this->firstSolve();
        m << "Done" << endl;
    }

    // BumponTailInstabilityManager.h:196-303
    void initializeParticles() {
        Inform m("Initialize Particles");
        auto* mesh    = &this->fcontainer_m->getMesh();
        auto* FL      = &this->fcontainer_m->getFL();
        using DistR_t = ippl::random::Distribution<double, D>;
        double parR[2 * D];
        for (unsigned int i = 0; i < D; i++) {
            parR[i * 2]     = delta_m;
            parR[i * 2 + 1] = this->kw_m[i];
        }
        // CustomDistributionFunctions (:23-52): the perturbation acts along the last dimension only
        DistR_t distR({ippl::random::UNIFORM, ippl::random::UNIFORM, ippl::random::COSINE}, parR);
        this->firstRepartition(distR);
        static IpplTimings::TimerRef particleCreation = IpplTimings::getTimer("particlesCreation");
        IpplTimings::startTimer(particleCreation);
        ippl::detail::RegionLayout<double, D, Mesh_t<D>> rlayout(*FL, *mesh);
        size_type totalP              = this->totalP_m;
        int seed                      = 42;
        const std::uint64_t pool_seed = (std::uint64_t)(seed + 100 * ippl::Comm->rank());
        using samplingR_t             = ippl::random::InverseTransformSampling<double, D, void, DistR_t>;
        Vector_t<double, D> rmin      = this->rmin_m;
        Vector_t<double, D> rmax      = this->rmax_m;
        samplingR_t samplingR(distR, rmax, rmin, rlayout, totalP);
        size_type nlocal = samplingR.getLocalSamplesNum();

        double factorVelBulk = 1.0 - epsilon_m;
        double factorVelBeam = 1.0 - factorVelBulk;
        size_type nlocBulk   = (size_type)(factorVelBulk * nlocal);
        size_type nlocBeam   = (size_type)(factorVelBeam * nlocal);
        nlocal               = nlocBulk + nlocBeam;
        int rank             = ippl::Comm->rank();
        size_type nglobal    = nlocal;
        ippl::Comm->allreduce(nglobal, 1, std::plus<size_type>());
        int rest = (int)(totalP - nglobal);
        if (rank < rest) ++nlocal;

        this->pcontainer_m->create(nlocal);
        samplingR.setLocalSamplesNum(nlocal);
        samplingR.generate(this->pcontainer_m->R, pool_seed);

        double mu[D], sd[D];
        for (unsigned int i = 0; i < D; i++) {
            mu[i] = 0.0;
            sd[i] = sigma_m;
        }
        // sample first nlocBulk with muBulk as mean velocity, the remaining with muBeam
        mu[D - 1] = muBulk_m;
        ippl::random::randn<double, D>(this->pcontainer_m->P, pool_seed, mu, sd, 0, nlocBulk);
        mu[D - 1] = muBeam_m;
        ippl::random::randn<double, D>(this->pcontainer_m->P, pool_seed, mu, sd, nlocBulk, nlocal);
        ippl::fence();
        ippl::Comm->barrier();
        IpplTimings::stopTimer(particleCreation);
        this->pcontainer_m->q = this->Q_m / totalP;
        m << "particles created and initial conditions assigned " << endl;
    }

    void advance() override {
        if (this->stepMethod_m != "LeapFrog") throw IpplException(TestName, "Step method is not set/recognized!");
        if (this->fused_m) {
            ipplb_push push{};
            push.kind = IPPLB_PUSH_LEAPFROG;
            push.dt   = this->dt_m;
            this->fusedStep(push);
        } else {
            LeapFrogStep();
        }
    }

    // BumponTailInstabilityManager.h:315-369 (load balancing is a no-op on one rank)
    void LeapFrogStep() {
        static IpplTimings::TimerRef PTimer      = IpplTimings::getTimer("pushVelocity");
        static IpplTimings::TimerRef RTimer      = IpplTimings::getTimer("pushPosition");
        static IpplTimings::TimerRef updateTimer = IpplTimings::getTimer("update");
        static IpplTimings::TimerRef SolveTimer  = IpplTimings::getTimer("solve");
        double dt                                              = this->dt_m;
        std::shared_ptr<typename Base::ParticleContainer_t> pc = this->pcontainer_m;
        IpplTimings::startTimer(PTimer);
        pc->P = pc->P - 0.5 * dt * pc->E;
        IpplTimings::stopTimer(PTimer);
        IpplTimings::startTimer(RTimer);
        pc->R = pc->R + dt * pc->P;
        IpplTimings::stopTimer(RTimer);
        IpplTimings::startTimer(updateTimer);
        pc->update();
        IpplTimings::stopTimer(updateTimer);
        this->maybeRepartition();
        this->par2grid();
        IpplTimings::startTimer(SolveTimer);
        this->fsolver_m->solve();
        IpplTimings::stopTimer(SolveTimer);
        this->grid2par();
        IpplTimings::startTimer(PTimer);
        pc->P = pc->P - 0.5 * dt * pc->E;
        IpplTimings::stopTimer(PTimer);
    }

    // dumpBumponTailInstability (:448-500): Ez field energy and max norm over interior cells
    void dump() override {
        auto& E = this->fcontainer_m->getE();
        double st[7];
        ippl::b200::check(ipplb_field_energy_stats(ippl::b200::ctx(), &E.b200_mesh(), E.data(), st), "dump");
        double globaltemp = 0.0, EzAmp = 0.0;
        ippl::Comm->reduce(st[D - 1], globaltemp, 1, std::plus<double>());
        ippl::Comm->reduce(st[3 + (D - 1)], EzAmp, 1, std::greater<double>());
        double fieldEnergy = std::accumulate(this->hr_m.begin(), this->hr_m.end(), globaltemp, std::multiplies<double>());
        if (ippl::Comm->rank() == 0) {
            std::filesystem::create_directory("data");
            std::stringstream fname;
            fname << "data/FieldBumponTail_" << ippl::Comm->size() << "_manager.csv";
            std::ofstream csvout(fname.str(), std::fabs(this->time_m) < 1e-14 ? std::ios::trunc : std::ios::app);
            csvout.precision(16);
            csvout.setf(std::ios::scientific, std::ios::floatfield);
            if (std::fabs(this->time_m) < 1e-14) csvout << "time, Ez_field_energy, Ez_max_norm" << std::endl;
            csvout << this->time_m << " " << fieldEnergy << " " << EzAmp << std::endl;
        }
        ippl::Comm->barrier();
    }

private:
    double sigma_m = 0, epsilon_m = 0, muBulk_m = 0, muBeam_m = 0, delta_m = 0;
};

int main(int argc, char* argv[]) {
    for (int i = 1; i < argc; ++i) g_two_stream |= std::string(argv[i]) == "--two-stream";
    return alpine_main<BumponTailInstabilityManager<T, Dim>>(argc, argv);
}
