// fe_umma_epi.cuh -- row-argmin epilogue of the tcgen05 kind::f16 search kernels: one thread owns one row x UM_HALF
// accumulator columns of a tile (fe_search_f16.cu).
#pragma once
#include "fe_umma_dev.cuh"

namespace umma_dev {

// 128 accumulator values of one row -> running (bestV, bestp, bestcol) and first threshold hit.
// Fast path: FMNMX3 tree for the tile minimum; the per-column scan only runs for lanes whose tile
// minimum can improve their row (or cross the threshold) and works on the registers already loaded.
struct RowState {
    float bestV;
    uint32_t bestp;            // parity of sum(b^2) of the best column; 2 = not looked up yet
    uint32_t bestcol, hit;
    float vthr0, vthr1;
    const uint32_t* par_item;  // tile metadata of the item's first tile (4 words per half tile: parity lo, parity hi, valid, bucket)
};
__device__ __forceinline__ uint32_t row_parity(const RowState& st, uint32_t col) {
    return (st.par_item[(col >> 6) * 4 + ((col >> 5) & 1u)] >> (col & 31)) & 1u;
}

// Exhaustive scan of a half tile held in registers (rare: exact ties, threshold crossings).
// Parity words of a half tile: held in registers when the tile metadata is read anyway (META), else fetched from the
// metadata record only on the rare paths that need them.
struct ParitySrc {
    uint32_t reg[UM_HALF / 32];
    const uint32_t* ptr;
};
template <bool META>
__device__ __forceinline__ uint32_t par_word(const ParitySrc& p, int wd) { return META ? p.reg[wd] : __ldg(p.ptr + wd); }

template <bool META>
__device__ __forceinline__ void scan_half_full(const uint32_t (&v)[UM_HALF], RowState& st, bool need_best, bool need_hit, uint32_t colbase,
                                               const ParitySrc& par) {
    float cx = 3.0e38f;
    uint32_t cp = 1, ccol = FE_NONE32, chit = FE_NONE32;
#pragma unroll
    for (int wd = 0; wd < UM_HALF / 32; ++wd) {
        const uint32_t pw = par_word<META>(par, wd);
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int i = wd * 32 + b;
            const float x = __uint_as_float(v[i]);
            const uint32_t p = (pw >> b) & 1u;
            if (x < cx || (x == cx && p < cp)) { cx = x; cp = p; ccol = colbase + i; }
            if (chit == FE_NONE32 && x <= (p ? st.vthr1 : st.vthr0)) chit = colbase + i;
        }
    }
    if (need_best && (cx < st.bestV || (cx == st.bestV && cp < st.bestp))) { st.bestV = cx; st.bestp = cp; st.bestcol = ccol; }
    if (need_hit && chit != FE_NONE32) st.hit = chit;
}

// UM_HALF accumulator values of one row (columns colbase .. colbase+UM_HALF-1 of the work item).
template <bool META>
__device__ __forceinline__ void process_half(uint32_t (&v)[UM_HALF], RowState& st, bool row_ok, uint32_t colbase, uint32_t nvalid,
                                             const ParitySrc& par, bool want_min) {
    if (nvalid < UM_HALF) {
#pragma unroll
        for (int i = 0; i < UM_HALF; ++i)
            if ((uint32_t)i >= nvalid) v[i] = 0x7F61B1E6u; // 3.0e38f
    }
    float grp[UM_HALF / 8]; // minimum of every group of 8 columns; written level by level for ILP (8 independent chains)
    float ga[UM_HALF / 8], gb[UM_HALF / 8];
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; ++k) ga[k] = fmin3(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1]), __uint_as_float(v[8 * k + 2]));
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; ++k) gb[k] = fmin3(__uint_as_float(v[8 * k + 3]), __uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5]));
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; ++k) ga[k] = fmin3(ga[k], __uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7]));
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; ++k) grp[k] = fminf(ga[k], gb[k]);
    float t0 = 3.0e38f, t1 = 3.0e38f;
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; k += 4) {
        t0 = fmin3(t0, grp[k], grp[k + 1]);
        t1 = fmin3(t1, grp[k + 2], grp[k + 3]);
    }
    const float tmin = fminf(t0, t1);
    // A row that already crossed the threshold is finished for this work item: its range is decided by the first hit in
    // scan order, and no later column can come earlier (columns are visited in increasing order).
    const bool live = row_ok && st.hit == FE_NONE32;
    const bool improve = want_min && live && tmin < st.bestV;
    const bool tie = want_min && live && tmin == st.bestV;           // an equal V with even parity could win
    const bool need_hit = live && tmin <= st.vthr0;
    if (improve | tie | need_hit) {
        if (need_hit) {
            // first column with V <= vthr(parity): walk the groups of 8 in order, look inside the first that can hold one
            uint32_t chit = FE_NONE32;
#pragma unroll
            for (int k = 0; k < UM_HALF / 8; ++k) {
                if (chit == FE_NONE32 && grp[k] <= st.vthr0) {
                    const uint32_t pw = par_word<META>(par, (8 * k) >> 5) >> ((8 * k) & 31);
#pragma unroll
                    for (int e = 7; e >= 0; --e) {
                        const float x = __uint_as_float(v[8 * k + e]);
                        if (x <= (((pw >> e) & 1u) ? st.vthr1 : st.vthr0)) chit = (uint32_t)(8 * k + e);
                    }
                }
            }
            if (chit != FE_NONE32) { st.hit = colbase + chit; return; }
        }
        bool full = false;
        if (tie) {
            if (st.bestp == 2) st.bestp = row_parity(st, st.bestcol);
            full = st.bestp == 1;
        }
        if (improve && !full) {
            // common case: the minimum is held by exactly one column -> it is the best of these columns
            // whatever its parity (looked up lazily); locate it through its group of 8
            int gi = UM_HALF / 8 - 1, ng = 0;
#pragma unroll
            for (int k = UM_HALF / 8 - 1; k >= 0; --k) {
                const bool e = grp[k] == tmin;
                gi = e ? k : gi;
                ng += e ? 1 : 0;
            }
            int ei = 7, ne = 0;
#pragma unroll
            for (int k = 0; k < UM_HALF / 8; ++k) {
                if (k == gi) {
#pragma unroll
                    for (int e = 7; e >= 0; --e) {
                        const bool q = __uint_as_float(v[8 * k + e]) == tmin;
                        ei = q ? e : ei;
                        ne += q ? 1 : 0;
                    }
                }
            }
            if (ng == 1 && ne == 1) {
                st.bestV = tmin; st.bestp = 2; st.bestcol = colbase + (uint32_t)(gi * 8 + ei);
            } else {
                full = true;                       // several columns tie on V: parity decides
            }
        }
        if (full) {
            if (st.bestp == 2 && st.bestcol != FE_NONE32) st.bestp = row_parity(st, st.bestcol);
            scan_half_full<META>(v, st, true, false, colbase, par);
        }
    }
}

// Lower-bound prefilter (fe_lb.cu): every column of the half tile whose value is at or below the row's threshold is a
// CANDIDATE and is handed to `emit` (column inside the half tile); nothing else is tracked.
template <bool META, typename Emit>
__device__ __forceinline__ void candidates_half(uint32_t (&v)[UM_HALF], float vthr0, float vthr1, bool row_ok, uint32_t nvalid, const ParitySrc& par,
                                                Emit&& emit) {
    if (nvalid < UM_HALF) {
#pragma unroll
        for (int i = 0; i < UM_HALF; ++i)
            if ((uint32_t)i >= nvalid) v[i] = 0x7F61B1E6u; // 3.0e38f
    }
    float grp[UM_HALF / 8];
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; ++k) {
        const float a = fmin3(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1]), __uint_as_float(v[8 * k + 2]));
        const float b = fmin3(__uint_as_float(v[8 * k + 3]), __uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5]));
        grp[k] = fmin3(fminf(a, b), __uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7]));
    }
    float t0 = 3.0e38f, t1 = 3.0e38f;
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; k += 4) {
        t0 = fmin3(t0, grp[k], grp[k + 1]);
        t1 = fmin3(t1, grp[k + 2], grp[k + 3]);
    }
    if (!(row_ok && fminf(t0, t1) <= vthr0)) return;
#pragma unroll
    for (int k = 0; k < UM_HALF / 8; ++k) {
        if (grp[k] <= vthr0) {
            const uint32_t pw = par_word<META>(par, (8 * k) >> 5) >> ((8 * k) & 31);
#pragma unroll
            for (int e2 = 0; e2 < 8; ++e2)
                if (__uint_as_float(v[8 * k + e2]) <= (((pw >> e2) & 1u) ? vthr1 : vthr0)) emit((uint32_t)(8 * k + e2));
        }
    }
}

} // namespace umma_dev
