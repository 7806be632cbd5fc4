// Drop-in forwarding header: same include path as the reference's encode/EncodingEngine2.hpp; the implementation lives in frac_b200/encode.hpp.
#pragma once
#include "frac_b200/encode.hpp"
