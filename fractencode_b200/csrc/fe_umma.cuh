// fe_umma.cuh -- tile constants shared by the tcgen05 search kernels (fe_search_f16.cu, fe_search_i8.cu).
#pragma once
#include <algorithm>

#include "fe_internal.cuh"

constexpr int UM_NT = 128;         // f16 kind: domain columns per tile (UMMA N)
constexpr int UM_ROWS = 128;       // rows per tile = 32 range blocks x 4 rotations (UMMA M)
constexpr int UM_WGS = 2;          // compute warpgroups (each owns 2 TMEM accumulators)
constexpr int UM_HALF = UM_NT / 2; // f16 kind: columns a compute thread holds per register set
// i8 kind: one issuer thread per accumulator buffer (4): the issue loop of a tile is latency bound (barrier probes ~200 cycles,
// commit ~170) and four loops overlap; one producer warp in front of them.
constexpr int UM_ISSUERS_I8 = 2 * UM_WGS;
constexpr int UM_THREADS_I8 = 32 + 32 * UM_ISSUERS_I8 + 256 * UM_WGS;
constexpr int I8_NT = 64;          // i8 kind: domain columns per tile (two s32 accumulators, low/high byte plane, of 64 columns)
constexpr int I8_KC = 256;         // i8 kind: bytes of K per shared-memory stage
constexpr int I8_MAX_STAGES = 4;
