// compat/Utility/IpplTimings.h -- IpplTimings lives in include/ippl/Ippl.h
#pragma once
#include "Ippl.h"
