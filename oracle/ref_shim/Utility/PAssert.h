// stand-in for Utility/PAssert.h (oracle/ref_shim): assertions throw.
#pragma once
#include <stdexcept>
#define PAssert(c) do { if (!(c)) throw std::runtime_error("PAssert failed: " #c); } while (0)
#define PAssert_EQ(a, b) PAssert((a) == (b))
#define PAssert_NE(a, b) PAssert((a) != (b))
#define PAssert_LT(a, b) PAssert((a) < (b))
#define PAssert_LE(a, b) PAssert((a) <= (b))
#define PAssert_GT(a, b) PAssert((a) > (b))
#define PAssert_GE(a, b) PAssert((a) >= (b))
#define PInsist(c, m) PAssert(c)
