// Drop-in forwarding header: same include path as the reference's encode/encode_parameters.h; the implementation lives in frac_b200/encode.hpp.
#pragma once
#include "frac_b200/encode.hpp"
