"""ippl_b200 -- B200-native (sm_100a) particle-mesh hot path behind IPPL's API.

This Python package is the thin host-side mirror used by tests and bench.py: it loads the C-ABI
library (ippl_b200/libippl_b200.so, built from ippl_b200/csrc by `__graft_entry__.build()`), and
wraps device memory held in torch tensors.  The product is the C-ABI + the CUDA kernels; the C++
facade that keeps IPPL's ParticleAttrib / Field / FieldLayout / ParticleSpatialLayout surface is in
include/ippl/.  There is NO CPU fallback: without the library or without a CUDA device every
compute call raises.
"""
from .lib import (Bins, Context, Dist, IpplbError, Layout, Loop, Mesh, Orb, Particles, Poisson, Push, SlabPlan, lib,  # noqa: F401
                  lib_path, exported_symbols, leapfrog_push, lib_particles_array, nccl_unique_id, penning_push,
                  sample_counts)

__all__ = ["Bins", "Context", "Dist", "IpplbError", "Layout", "Loop", "Mesh", "Orb", "Particles", "Poisson", "Push", "SlabPlan", "lib",
           "lib_path", "exported_symbols", "leapfrog_push", "nccl_unique_id", "penning_push", "sample_counts"]
