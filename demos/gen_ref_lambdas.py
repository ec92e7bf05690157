"""Cuts the Kokkos lambda BODIES of the reference's alpine drivers out of the reference tree, so that demos/ref_lambdas.cu
can compile the drivers' own device code unchanged on include/ippl/KokkosShim.cuh:
  * "Kick1" / "Kick2" of demos/alpine/PenningTrapManager.h (:256-272, :313-333)          -> penning_kick{1,2}.inc
  * "Ex stats" of demos/alpine/LandauDampingManager.h: over the field (:346-360)           -> landau_ex_stats_field.inc
                                                       over the particles (:401-406)       -> landau_ex_stats_particles.inc
  * "Particle Kinetic Energy" / "Vector E reduce" of PenningTrapManager.h (:354-360, :374-382) -> penning_kinetic.inc,
                                                                                               penning_vector_e.inc
  * "Ex inner product" / "Ex max norm" of BumponTailInstabilityManager.h (:460-468, :478-488) -> bumpontail_inner.inc,
                                                                                               bumpontail_max.inc
The outputs go to the directory given on the command line; the Makefile passes a temporary directory and removes it after
the compile, so no reference text stays in this tree (only the built binary, which is git-ignored).
usage: python gen_ref_lambdas.py <reference root> <output dir>"""
import os
import sys


def lambda_bodies(lines, name):
    """bodies of every `"name", ..., KOKKOS_LAMBDA(...) {` ... `}` in order of appearance (closing line: `}` + `,` or `);`)"""
    out = []
    i = 0
    while i < len(lines):
        if f'"{name}"' in lines[i]:
            j = i
            while "KOKKOS_LAMBDA" not in lines[j]:
                j += 1
            assert lines[j].rstrip().endswith("{"), lines[j]
            indent = len(lines[j]) - len(lines[j].lstrip())
            body = []
            k = j + 1
            while not (lines[k].strip() in ("});", "},") and len(lines[k]) - len(lines[k].lstrip()) <= indent):
                body.append(lines[k])
                k += 1
            out.append(body)
            i = k
        i += 1
    return out


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    pen = open(os.path.join(ref, "demos", "alpine", "PenningTrapManager.h")).read().splitlines()
    for k in (1, 2):
        (body,) = lambda_bodies(pen, f"Kick{k}")
        assert 10 <= len(body) <= 30 and any("Bext" in l for l in body), (k, len(body))
        open(os.path.join(out, f"penning_kick{k}.inc"), "w").write("\n".join(body) + "\n")
    lan = open(os.path.join(ref, "demos", "alpine", "LandauDampingManager.h")).read().splitlines()
    field, particles = lambda_bodies(lan, "Ex stats")
    assert any("ippl::apply(Eview, args)" in l for l in field) and any("ENorm" in l for l in field), field
    assert any("Eview(i)[0]" in l for l in particles), particles
    open(os.path.join(out, "landau_ex_stats_field.inc"), "w").write("\n".join(field) + "\n")
    open(os.path.join(out, "landau_ex_stats_particles.inc"), "w").write("\n".join(particles) + "\n")
    for name, inc, key in (("Particle Kinetic Energy", "penning_kinetic.inc", "dot(Pview(i), Pview(i))"),
                           ("Vector E reduce", "penning_vector_e.inc", "ippl::apply(Eview, args)[d]")):
        (body,) = lambda_bodies(pen, name)
        assert any(key in l for l in body), (name, body)
        open(os.path.join(out, inc), "w").write("\n".join(body) + "\n")
    bum = open(os.path.join(ref, "demos", "alpine", "BumponTailInstabilityManager.h")).read().splitlines()
    for name, inc in (("Ex inner product", "bumpontail_inner.inc"), ("Ex max norm", "bumpontail_max.inc")):
        (body,) = lambda_bodies(bum, name)
        assert any("ippl::apply(Eview, args)[Dim - 1]" in l for l in body), (name, body)
        open(os.path.join(out, inc), "w").write("\n".join(body) + "\n")


if __name__ == "__main__":
    main()
