// fe_search_umma.cu -- tcgen05/TMEM contraction path (placeholder until the kernel lands).
#include "fe_internal.cuh"

int umma_level_supported(const LevelGeom&) { return 0; }
int launch_search_umma(fe_ctx* ctx, const LevelGeom&, const SearchArgs&, const void*, const void*) {
    return fe_fail(ctx, FE_ERR_UNSUPPORTED, "tcgen05 search path not built");
}
