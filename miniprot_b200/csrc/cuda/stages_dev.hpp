// stages_dev.hpp -- host entry points of the seeding/chaining and refinement stages (seed_chain.cu, refine.cu).
#pragma once
#include "ctx.hpp"

namespace mpb {
namespace cuda {

// S1 (map.c:155-195): per protein sketch -> index lookup -> anchor sort -> pre-chain -> main chain
void seed_chain_run(mpb_ctx_s *ctx, const mp_idx_t *mi, const mp_mapopt_t *opt, const Batch &b, const std::vector<int32_t> &aa_off, const char *d_aa,
                    ChainSet &out);
// S2 (map.c:41-97): per window 5-mer join with the protein + base-level chain, best chain kept
void refine_run(mpb_ctx_s *ctx, const mp_idx_t *mi, const mp_mapopt_t *opt, const Batch &b, const std::vector<int32_t> &aa_off, const char *d_aa,
                const std::vector<RefineJob> &jobs, RefineSet &out);

} // namespace cuda
} // namespace mpb
