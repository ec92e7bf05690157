// compat: ippl::OrthogonalRecursiveBisection lives in include/ippl/Ippl.h
#pragma once
#include "Ippl.h"
