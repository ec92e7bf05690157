// Compile-time census of the facade's API surface beyond what the three drivers use: the (policy, hash) overload of
// ippl::scatter (src/Particle/ParticleAttrib.hpp:332-334), ippl::ParameterList (src/Utility/ParameterList.h), the two
// forms of IpplTimings::print.  Built by `make -C demos` (part of __graft_entry__.build()); like the drivers it needs
// a CUDA device to run and has no CPU fallback.
constexpr unsigned Dim = 3;
using T = double;
const char* TestName = "api_check";
#include "Alpine.h"

int main(int argc, char* argv[]) {
    try {
        ippl::initialize(argc, argv);
    } catch (const IpplException& ex) {
        std::cerr << TestName << ": cannot start: " << ex.what() << " (a CUDA device is required; there is no CPU fallback)" << std::endl;
        return 2;
    }
    int rc = 0;
    {
        ippl::ParameterList params, fft;
        fft.add("use_heffte_defaults", false);
        fft.add("r2c_direction", 0);
        params.add("output_type", 1);
        params.add("tolerance", 1e-10);
        params.add("solver", "FFT");
        params.add("fft", fft);
        params.update("tolerance", 1e-12);
        ippl::ParameterList over;
        over.add("tolerance", 1e-8);
        over.add("unknown", 3);
        params.update(over);   // only keys that exist are updated
        rc |= params.get<double>("tolerance") != 1e-8 || params.contains("unknown") || params.get<int>("missing", 7) != 7;
        rc |= params.get<ippl::ParameterList>("fft").get<int>("r2c_direction") != 0;
        std::cout << params << std::endl;

        const int n = 1000;
        ippl::Vector<int, 3> nr(8);
        ippl::NDIndex<3> domain;
        for (unsigned d = 0; d < 3; ++d) domain[d] = ippl::Index(nr[d]);
        Vector_t<double, 3> hr(1.0 / 8), origin(0.0), rmin(0.0), rmax(1.0);
        std::array<bool, 3> decomp{true, true, true};
        FieldContainer<double, 3> fc(hr, rmin, rmax, decomp, domain, origin, true);
        fc.initializeFields();
        ParticleContainer<double, 3> pc(fc.getMesh(), fc.getFL());
        pc.create(n);
        pc.q = 1.0 / n;
        std::vector<ippl::Vector<double, 3>> host(n);
        for (int i = 0; i < n; ++i) {
            host[i][0] = (i % 97) / 97.0;
            host[i][1] = (i % 89) / 89.0;
            host[i][2] = (i % 83) / 83.0;
        }
        pc.R.copyFromHost(host);
        // every second particle through a hash remap over the first half of the range == plain scatter of those particles
        std::vector<int> h(n / 2);
        for (int i = 0; i < n / 2; ++i) h[i] = 2 * i;
        int* d_hash = ippl::b200::device_alloc<int>(h.size());
        cudaMemcpy(d_hash, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice);
        fc.getRho() = 0.0;
        ippl::scatter(pc.q, fc.getRho(), pc.R, ippl::RangePolicy1D(0, n / 2), ippl::detail::hash_type(d_hash, h.size()));
        const double half = fc.getRho().sum();
        fc.getRho() = 0.0;
        ippl::scatter(pc.q, fc.getRho(), pc.R, ippl::RangePolicy1D(0, n));
        const double all = fc.getRho().sum();
        cudaFree(d_hash);
        rc |= std::fabs(half - 0.5) > 1e-12 || std::fabs(all - 1.0) > 1e-12;
        std::cout << "scatter(policy, hash): charge " << half << " of " << all << (rc ? "  FAILED" : "  ok") << std::endl;
        IpplTimings::print();
    }
    ippl::finalize();
    return rc;
}
