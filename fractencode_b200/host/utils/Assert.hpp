// Drop-in forwarding header: same include path as the reference's utils/Assert.hpp; the implementation lives in frac_b200/core.hpp.
#pragma once
#include "frac_b200/core.hpp"
