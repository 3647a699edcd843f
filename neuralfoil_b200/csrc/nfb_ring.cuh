// The weight stream of the column-group kernels: one ring of TMA bulk copies per CTA that every warp consumes in the same
// order ("The weight ring").
//
// Slice j of the CTA's run lands in ring slot j % STAGES and is issued by warp j % WARPS when that warp starts slice
// j - LEAD.  The issuer waits until every warp has released slice j - STAGES, so a warp runs at most STAGES - LEAD slices
// ahead of the slowest one without blocking, and finds its weights as long as it is less than LEAD slices ahead of the
// issuers.
//
// Round 2: all indices are kept incrementally.  The first version recomputed j % STAGES, (j / STAGES) & 1 and
// j % slices_per_tile (a run-time divisor: I2F, MUFU.RCP, ...) for every slice and decided "is it my turn to issue" with a
// chain of four dependent predicate instructions; together with the BSSY / BSYNC pair around the single-lane issue and the
// WARPSYNC before the release that was 108 issue cycles per warp and slice on the path that does NOT issue -- a tenth of the
// 1,024 FMA-pipe cycles of a 512-wide slice, a fifth of a 64-wide one (SASS of round 1, profiles/r2_slice_loop_overhead.txt);
// switching the ring off altogether measured 4.5 % of the xxxlarge kernel's time.  Now the consuming path is: one countdown
// test, one try_wait, one predicated arrive, two counter updates.
#pragma once
#include "nfb_common.cuh"

namespace nfb {

template <int STAGES, int LEAD, int WARPS, int SLOT_FLOATS>
struct WeightRing {
  static_assert(LEAD >= 1 && LEAD < STAGES && LEAD <= WARPS, "bad ring geometry");
  float* ring;             // [STAGES][SLOT_FLOATS]
  const uint2* table;      // {float offset in the blob, bytes} of every slice of a tile
  const float* blob;
  uint32_t bar_full, bar_empty;
  uint32_t slices_per_tile, total_slices;
  uint32_t slot, phase;    // of the slice this warp consumes next: it % STAGES, (it / STAGES) & 1
  uint32_t turn;           // slices this warp still consumes before it issues again
  uint32_t j_next;         // the slice it issues then, with its slot, phase and position in the tile's table
  uint32_t iss_slot, iss_phase, iss_tpos;
  bool lane0;

  // Call with all threads of the CTA after `table` is written and before anything else; ends with a CTA barrier.
  __device__ __forceinline__ void init(float* ring_, const uint2* table_, const float* blob_, uint32_t bar_full_,
                                       uint32_t bar_empty_, uint32_t slices_per_tile_, uint32_t total_slices_) {
    ring = ring_; table = table_; blob = blob_; bar_full = bar_full_; bar_empty = bar_empty_;
    slices_per_tile = slices_per_tile_; total_slices = total_slices_;
    const uint32_t warp = threadIdx.x >> 5;
    lane0 = (threadIdx.x & 31) == 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_empty + 8 * s, WARPS);
      }
      fence_mbar_init();
    }
    __syncthreads();
    slot = 0; phase = 0;
    j_next = warp >= uint32_t(LEAD) ? warp : warp + WARPS;  // first slice >= LEAD whose turn is this warp's
    turn = j_next - LEAD;
    iss_slot = j_next % STAGES;
    iss_phase = (j_next / STAGES) & 1;
    iss_tpos = j_next % slices_per_tile;
    if (threadIdx.x == 0)
      for (uint32_t j = 0; j < uint32_t(LEAD) && j < total_slices; ++j) copy(j % slices_per_tile, j % STAGES);
  }

  __device__ __forceinline__ void copy(uint32_t tpos, uint32_t s) {  // one thread
    const uint2 e = table[tpos];
    mbar_arrive_expect_tx(bar_full + 8 * s, e.y);
    bulk_copy_g2s(smem_u32(ring + s * SLOT_FLOATS), blob + e.x, e.y, bar_full + 8 * s);
  }

  // Makes the next slice visible (and keeps the ring fed); returns its slot.
  __device__ __forceinline__ const float* wait() {
    if (turn == 0) {  // warp-uniform (a vote to prove it to the compiler costs more than the BSSY / BSYNC pair it saves)
      if (j_next < total_slices) {
        mbar_wait(bar_empty + 8 * iss_slot, iss_phase ^ 1);  // every warp has released slice j_next - STAGES
        if (lane0) copy(iss_tpos, iss_slot);
      }
      j_next += WARPS;
      constexpr uint32_t DS = WARPS % STAGES, DP = (WARPS / STAGES) & 1;
      iss_slot += DS;
      iss_phase ^= DP;
      if (iss_slot >= uint32_t(STAGES)) { iss_slot -= STAGES; iss_phase ^= 1; }
      iss_tpos += WARPS;
      while (iss_tpos >= slices_per_tile) iss_tpos -= slices_per_tile;
      turn = WARPS;
    }
    --turn;
    mbar_wait(bar_full + 8 * slot, phase);
    return ring + slot * SLOT_FLOATS;
  }

  // The warp is converged here and its loads from the slot have all returned (their consumers have issued), so lane 0 may
  // hand the slot back without a warp barrier.
  __device__ __forceinline__ void release() {
    if (lane0) mbar_arrive(bar_empty + 8 * slot);
    if (++slot == uint32_t(STAGES)) { slot = 0; phase ^= 1; }
  }
};

}  // namespace nfb
