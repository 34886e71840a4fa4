// fdk_color.cuh -- one-time, per cluster: where the 3x3 blocks of the balanced hex8 kernel (fdk_assemble_iso.cuh) sit in
// the shared-memory staging array, chosen so that NEITHER side of the exchange has shared-memory bank conflicts.
//
// Producer side: a half-warp is 4 incidences x 4 column parts and stores, for jj = 0, 1, the 16 blocks
// (it, part + 4 jj) -- call that a producer group G = (it / 4) * 2 + jj.  Consumer side: a half-warp is 16 consecutive
// CSR block slots, and at step t every lane loads the t-th contribution of its slot -- a consumer group (s / 16, t).
// A block is 9 doubles at position P: its first 8-byte bank is 9 P mod 16, a bijection of P mod 16.  So give group G
// the window of positions [16 G, 16 G + 16) and each of its blocks a distinct slot of the window (the stores of the
// group are then conflict-free whatever the permutation), and choose the permutations so that the blocks of a
// consumer group also carry distinct slots: a bipartite edge colouring (producer groups x consumer groups, degree
// <= 16, hence 16-colourable by Koenig's theorem).  Greedy colouring with a rotating first choice, then a few
// sweeps of pairwise swaps inside a window for the blocks still in conflict: ~7 % excess wavefronts instead of
// ~90 % with positions it * stride + j * 9 (simulated and measured).
//
// No reference counterpart: the reference's COO -> CSR conversion is a sparse matrix product (fedoo/core/
// _sparsematrix.py:256-302); this is how its cluster-local analogue is laid out in shared memory.
#pragma once
#include "fdk_assemble.cuh"

namespace fdk {

constexpr int COLOR_MAX_BLOCKS = 2048;  // 256 incidences x 8 column blocks
constexpr int COLOR_MAX_CG = 512;       // consumer groups: (slots / 16 + 1) x HEAVY_T

__global__ void __launch_bounds__(32) k_color_blocks(const fdk_plan p, unsigned char* __restrict__ blk_slot,
                                                     unsigned short* __restrict__ ent_pos) {
  constexpr int NNE = 8, TPI = 4, W = 16;
  __shared__ unsigned short s_cg[COLOR_MAX_BLOCKS];   // consumer group of every block (0xFFFF: unconstrained)
  __shared__ unsigned char s_slot[COLOR_MAX_BLOCKS];  // window slot of every block
  __shared__ unsigned char s_use[COLOR_MAX_CG][W];    // how many blocks of a consumer group sit in each slot
  const int c = blockIdx.x, lane = threadIdx.x;
  const ClusterHdr h = load_hdr(p.cl_hdr, c);
  const int n_inc = h.n_inc, n_slots = h.n_slots, n_blocks = n_inc * NNE;
  const unsigned* rec = p.slot_rec + h.slot0 + c;
  const unsigned short* ent = p.ent_src + h.ent0;
  for (int b = lane; b < n_blocks; b += 32) {
    s_cg[b] = 0xFFFF;
    s_slot[b] = 0;
  }
  for (int i = lane; i < COLOR_MAX_CG * W; i += 32) (&s_use[0][0])[i] = 0;
  __syncwarp();
  // consumer group of every block: (half-warp of its slot, rank among the slot's contributions); the other
  // contributions of a heavy slot are read one after the other by a single thread and stay unconstrained
  for (int s = lane; s < n_slots; s += 32) {
    const unsigned r0 = rec[s], r1 = rec[s + 1];
    const int e0 = r0 & 0xFFFF;
    const int cnt = (int)(r1 & 0xFFFF) - e0 - (((r0 ^ r1) >> 24) ? 1 : 0);
    // (a heavy slot is pre-reduced into its FIRST entry, which the gather then reads at step 0 like any other slot)
    const int n_con = cnt <= HEAVY_T ? cnt : 1;
    for (int t = 0; t < n_con; ++t) {
      const int cg = (s >> 4) * HEAVY_T + t;
      s_cg[ent[e0 + t]] = cg < COLOR_MAX_CG ? (unsigned short)cg : 0xFFFF;
    }
  }
  __syncwarp();
  if (lane == 0) {
    const int n_groups = ((n_inc + TPI - 1) / TPI) * 2;
    auto block_of = [&](int G, int l) {  // lane l of producer group G -> block id, or -1 past the last incidence
      const int it = (G >> 1) * TPI + (l >> 2);
      return it < n_inc ? it * NNE + iso_col(l & 3, G & 1) : -1;
    };
    for (int G = 0; G < n_groups; ++G) {  // greedy
      unsigned pmask = 0;
      for (int l = 0; l < W; ++l) {
        const int b = block_of(G, l);
        if (b < 0) continue;
        const int cg = s_cg[b];
        unsigned busy = pmask;
        if (cg != 0xFFFF)
          for (int k = 0; k < W; ++k) busy |= (s_use[cg][k] ? 1u : 0u) << k;
        unsigned avail = ~busy & 0xFFFFu;
        if (avail == 0) avail = ~pmask & 0xFFFFu;  // no slot free on both sides: keep the window a permutation
        const int r = (l * 5 + G * 3) & 15;       // rotating first choice: the slots fill evenly
        const unsigned rot = ((avail >> r) | (avail << (W - r))) & 0xFFFFu;
        const int sl = (__ffs(rot) - 1 + r) & 15;
        s_slot[b] = (unsigned char)sl;
        pmask |= 1u << sl;
        if (cg != 0xFFFF) ++s_use[cg][sl];
      }
    }
    for (int sweep = 0; sweep < 3; ++sweep) {  // repair: swap with another block of the window when both then fit
      int fixed = 0;
      for (int G = 0; G < n_groups; ++G) {
        for (int l = 0; l < W; ++l) {
          const int b = block_of(G, l);
          if (b < 0) continue;
          const int cg = s_cg[b];
          if (cg == 0xFFFF || s_use[cg][s_slot[b]] <= 1) continue;
          const int s1 = s_slot[b];
          bool done = false;
          for (int l2 = 0; l2 < W && !done; ++l2) {
            const int b2 = block_of(G, l2);
            if (b2 < 0 || b2 == b) continue;
            const int cg2 = s_cg[b2], s2 = s_slot[b2];
            if (cg2 == cg) continue;
            if (s_use[cg][s2] != 0) continue;
            if (cg2 != 0xFFFF && s_use[cg2][s1] != 0) continue;
            --s_use[cg][s1];
            ++s_use[cg][s2];
            if (cg2 != 0xFFFF) {
              --s_use[cg2][s2];
              ++s_use[cg2][s1];
            }
            s_slot[b] = (unsigned char)s2;
            s_slot[b2] = (unsigned char)s1;
            done = true;
            ++fixed;
          }
          if (!done) {  // a free slot of the window (groups of the last, partial incidence quadruple have holes)
            unsigned pm = 0;
            for (int l2 = 0; l2 < W; ++l2) {
              const int b2 = block_of(G, l2);
              if (b2 >= 0) pm |= 1u << s_slot[b2];
            }
            for (int k = 0; k < W && !done; ++k) {
              if (((pm >> k) & 1u) || s_use[cg][k] != 0) continue;
              --s_use[cg][s1];
              ++s_use[cg][k];
              s_slot[b] = (unsigned char)k;
              done = true;
              ++fixed;
            }
          }
        }
      }
      if (fixed == 0) break;
    }
  }
  __syncwarp();
  unsigned char* out_slot = blk_slot + (int64_t)h.inc0 * NNE;
  for (int b = lane; b < n_blocks; b += 32) out_slot[b] = s_slot[b];
  unsigned short* out_pos = ent_pos + h.ent0;
  const int n_ent = n_blocks + h.n_owned;
  for (int e = lane; e < n_ent; e += 32) {
    const int b = ent[e];  // gap entries hold 0: harmless, never read
    const int it = b / NNE, j = b - it * NNE;
    out_pos[e] = b < n_blocks ? (unsigned short)((((it >> 2) * 2 + iso_jj(j)) << 4) + s_slot[b]) : 0;
  }
}

}  // namespace fdk
