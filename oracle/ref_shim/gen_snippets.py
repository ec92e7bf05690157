"""TEST INFRASTRUCTURE.  Cuts small pieces of source out of the reference's alpine manager headers -- which cannot be
included, they pull in the whole framework -- so that the shims can compile the reference's OWN text in place:
  * the bodies of the Kokkos lambdas "Kick1" / "Kick2" of demos/alpine/PenningTrapManager.h
      -> penning_kick{1,2}.inc            (oracle/ref_shim/refshim_penning.cpp)
  * struct CustomDistributionFunctions of LandauDampingManager.h and BumponTailInstabilityManager.h
      -> landau_dist.inc, bumpontail_dist.inc   (oracle/ref_shim/refshim_random.cpp)
  * from src/Particle/ParticleSpatialLayout.hpp: the return statements of positionInRegion / positionInRegionInclusive
    and the body of the destRankOf lambda of locateParticlesPacked -> psl_in_region.inc, psl_in_region_inclusive.inc,
    psl_dest_rank.inc (refshim_locate.cpp)
  * the bodies of the lambdas "ParticleAttrib::scatter" and "ParticleAttrib::gather" of src/Particle/ParticleAttrib.hpp
    (the header needs the whole particle framework) -> attrib_scatter.inc, attrib_gather.inc (refshim_attrib.cpp)
  * the body of the k-space lambda "Gradient FFTPeriodicPoissonSolver" of src/PoissonSolvers/
    FFTPeriodicPoissonSolver.hpp (the solver header needs heFFTe) -> poisson_grad_lambda.inc (refshim_poisson.cpp)
The outputs go to the directory given on the command line -- the Makefile passes a temporary directory and removes it
after the compile: no reference text is left in the tree.
usage: python gen_snippets.py <reference root> <output dir>"""
import os
import sys


def lambda_body(lines, name):
    start = next(i for i, l in enumerate(lines) if f'"{name}"' in l and "KOKKOS_LAMBDA" in l)
    body = []
    for l in lines[start + 1:]:
        if l.strip() == "});":
            return body
        body.append(l)
    raise RuntimeError(f"end of lambda {name} not found")


def struct_text(lines, name):
    start = next(i for i, l in enumerate(lines) if l.startswith(f"struct {name} {{"))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] == "};")
    return lines[start:end + 1]


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    psl = open(os.path.join(ref, "src", "Particle", "ParticleSpatialLayout.hpp")).read().splitlines()
    for fn, inc in (("positionInRegionInclusive(", "psl_in_region_inclusive.inc"), ("positionInRegion(", "psl_in_region.inc")):
        # the out-of-class definition: "ParticleSpatialLayout<...>::<fn>" followed by its single return statement
        start = next(i for i, l in enumerate(psl) if l.strip().startswith("ParticleSpatialLayout<") and l.strip().endswith("::" + fn))
        ret = next(i for i in range(start, start + 6) if psl[i].strip().startswith("return "))
        assert psl[ret].strip().endswith(";"), psl[ret]
        with open(os.path.join(out, inc), "w") as f:
            f.write(psl[ret] + "\n")
    start = next(i for i, l in enumerate(psl) if "auto destRankOf = KOKKOS_LAMBDA" in l)
    body = []
    for l in psl[start + 1:]:
        if l.rstrip() == "        };":
            break
        body.append(l)
    assert 15 <= len(body) <= 35 and any("positionInRegionInclusive" in l for l in body), len(body)
    with open(os.path.join(out, "psl_dest_rank.inc"), "w") as f:
        f.write("\n".join(body) + "\n")
    pa = open(os.path.join(ref, "src", "Particle", "ParticleAttrib.hpp")).read().splitlines()
    for name, inc in (("ParticleAttrib::scatter", "attrib_scatter.inc"), ("ParticleAttrib::gather", "attrib_gather.inc")):
        start = next(i for i, l in enumerate(pa) if f'"{name}"' in l)
        while "KOKKOS_LAMBDA" not in pa[start]:
            start += 1
        body = []
        for l in pa[start + 1:]:
            if l.strip() == "});":
                break
            body.append(l)
        assert 8 <= len(body) <= 25 and any("invdx" in l for l in body), (name, len(body))
        with open(os.path.join(out, inc), "w") as f:
            f.write("\n".join(body) + "\n")
    hpp = open(os.path.join(ref, "src", "PoissonSolvers", "FFTPeriodicPoissonSolver.hpp")).read().splitlines()
    start = next(i for i, l in enumerate(hpp) if '"Gradient FFTPeriodicPoissonSolver"' in l)
    assert "KOKKOS_LAMBDA" in hpp[start + 1]
    body = []
    for l in hpp[start + 2:]:
        if l.strip() == "});":
            break
        body.append(l)
    assert 20 <= len(body) <= 45 and any("notMid" in l for l in body), len(body)
    with open(os.path.join(out, "poisson_grad_lambda.inc"), "w") as f:
        f.write("\n".join(body) + "\n")
    for app, inc in (("LandauDampingManager.h", "landau_dist.inc"), ("BumponTailInstabilityManager.h", "bumpontail_dist.inc")):
        src = open(os.path.join(ref, "demos", "alpine", app)).read().splitlines()
        text = struct_text(src, "CustomDistributionFunctions")
        assert 15 <= len(text) <= 45 and any("struct CDF" in l for l in text), (app, len(text))
        with open(os.path.join(out, inc), "w") as f:
            f.write("\n".join(text) + "\n")
    lines = open(os.path.join(ref, "demos", "alpine", "PenningTrapManager.h")).read().splitlines()
    os.makedirs(out, exist_ok=True)
    for k in (1, 2):
        body = lambda_body(lines, f"Kick{k}")
        assert 10 <= len(body) <= 30 and any("Bext" in l for l in body), (k, len(body))
        with open(os.path.join(out, f"penning_kick{k}.inc"), "w") as f:
            f.write("\n".join(body) + "\n")


if __name__ == "__main__":
    main()
