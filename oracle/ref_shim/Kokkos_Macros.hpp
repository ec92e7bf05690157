// Minimal stand-in for <Kokkos_Macros.hpp> -- TEST INFRASTRUCTURE (oracle/ref_shim).
// Lets a handful of pure-arithmetic headers of the reference compile on the host with g++ so the
// oracle restatement can be checked against the reference's own code.  Not Kokkos; not shipped.
#pragma once
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_DEFAULTED_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_CLASS_LAMBDA [=, *this]
