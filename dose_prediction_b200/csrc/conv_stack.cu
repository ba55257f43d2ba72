// Depth-stacked implicit-GEMM 3-D convolution for small C_out (16 / 32) on tcgen05 tensor cores.
//
// Why: with N = C_out = 16 a plain implicit GEMM issues one 128x16x16 MMA per tap and re-reads its 4 KB A
// operand from shared memory every time — ncu shows the tensor core's shared-memory read pipe at 61 % and
// the tensor pipe at 13 % (profiles/r1_conv_k7_plain.md).  Here the k taps along D are stacked into the
// N dimension instead: for one INPUT plane, one in-plane tap (kh,kw) and one 16-channel chunk, a single
// MMA with B = [W(kd=k-1) | ... | W(kd=0)]  (N = k*C_out = 112 / 224 for 7^3) updates the accumulators of
// the k OUTPUT planes that input plane contributes to.  The accumulators of consecutive output planes
// live in a ring of G = k+1 column groups per tile in TMEM (slot = plane counter mod G).  The weights are
// stored in G pre-rotated copies, each with one all-zero group, so that for input plane dz the copy
// rot = slot(dz - pad) has, at row block s, exactly the depth tap that ring slot s needs (and zeros for the
// one slot that is being drained): a single MMA with N = G*C_out covers the whole ring — no wrap-around
// splits, B is never rotated at run time (the producer just picks the copy).  A is read once per k taps.
//
// A CTA owns T adjacent 16(H) x 8(W) tiles (one wide TMA halo patch, B shared by the T tiles) and marches
// along a segment of L output planes; work items = (image, D segment, H tile, W tile group), persistent.
// Ring bookkeeping: every output plane gets a running counter pc; slot = pc % G, parity = (pc / G) & 1.
//   done_bar[slot]  MMA  -> epilogue   (tcgen05.commit after the last input plane contributing to it)
//   free_bar[slot]  epilogue -> MMA    (8 epilogue warps have drained it; next owner may start with acc=0)
// The first MMA that touches a new plane uses accumulate=0 on that plane's group only (the window is
// split into an old and a new part for that one tap), so TMEM never needs zeroing.
#include <cstdlib>

#include "common.cuh"
#include "dose_b200.h"

namespace dp {

struct StackParams {
  int N, D, H, W;
  int n_chunks, cb_total_in;
  int cout;                 // N0: MMA columns per ring slot
  int fold;                 // 1: output channel c = column c + column c + N0/2 (W_hi | W_lo operand split folded into N)
  int T, G, L;              // tiles per CTA, ring slots, segment length
  int tps, n_bst;           // taps per weight stage, weight stages per (plane, chunk)
  int PH, PWw;
  int a_stages, b_stages;
  uint32_t a_bytes, a_stride, b_bytes, b_stride, b_off;
  int tiles_h, wgroups, nseg, num_items, tiles_w;
  int balanced;             // 1: every CTA streams one contiguous share of the (column, plane) space (ItemIter), 0: fixed D segments
  long long total_planes;   // columns x D
  const __half* wpack;      // [G rotations][chunk][kh][kw][2][G*cout][8]
  const float* scale;
  const float* shift;
  int relu;
  float* out_f32;
  __half* out_hi;
  __half* out_lo;
  int cb_total_out, cb_out_off;
  double* stats;
  int* err_flag;
  int debug;                // timing experiments only (DP_STACK_DEBUG): 1 = skip weight copies, 2 = skip patch loads, 4 = skip epilogue work
  uint8_t chunk_cb[96];
};

constexpr int kStackThreads = 352;      // producer + MMA + 8 epilogue warps (2 per TMEM lane quarter) + second MMA issuer
constexpr int kStackEpiWarps = 8;
constexpr int kStackMmaWarpB = 10;      // issues the MMAs of the upper half of the CTA's tiles (single-thread issue is the
                                        // limiter: ncu shows ~100 cycles per MMA against a 64-cycle tensor floor)
constexpr int kMaxAStages = 4, kMaxBStages = 6, kMaxSlots = 32;

__device__ __forceinline__ void umma_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float colsum16s(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float keep = h4 ? v[j + 8] : v[j], send = h4 ? v[j] : v[j + 8];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float keep = h3 ? a[j + 4] : a[j], send = h3 ? a[j] : a[j + 4];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float keep = h2 ? b[j + 2] : b[j], send = h2 ? b[j] : b[j + 2];
    c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h1 ? c[1] : c[0], send = h1 ? c[0] : c[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// NT consecutive kw taps x TT tiles, fully unrolled with constant operand offsets (lets ptxas batch the
// descriptor moves): tap j of tile t uses A start + j + 8t, B start + j*tap_b16, accumulator tmem + t*tile_cols.
template <int NT, int TT>
__device__ __forceinline__ void issue_taps(uint32_t tmem, uint32_t tile_cols, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                           uint32_t b_hi, uint32_t tap_b16, uint32_t idesc) {
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int t = 0; t < TT; ++t) umma_split(tmem + t * tile_cols, a_lo + j + t * 8, a_hi, b_lo + j * tap_b16, b_hi, idesc, 1u);
  }
}

struct Item {
  int n, d0, d1, z0, z1, h0, w0, ntile;
};

template <int KS>
__device__ __forceinline__ Item decode_item(const StackParams& p, int item) {
  constexpr int pad = (KS - 1) / 2;
  Item it;
  int t = item;
  const int wg = t % p.wgroups; t /= p.wgroups;
  const int th = t % p.tiles_h; t /= p.tiles_h;
  const int sg = t % p.nseg;
  it.n = t / p.nseg;
  it.d0 = sg * p.L;
  it.d1 = min(p.D, it.d0 + p.L);
  it.z0 = max(0, it.d0 - pad);
  it.z1 = min(p.D - 1, it.d1 - 1 + pad);
  it.h0 = th * 16;
  it.w0 = wg * p.T * 8;
  it.ntile = min(p.T, p.tiles_w - wg * p.T);
  return it;
}

// Work distribution.  A column = (image, H tile, W tile group); its D planes are streamed in order.  Fixed D segments dealt
// round-robin leave the busiest SM with ceil(items / SMs) x (L + k - 1) planes (at 128^3, batch 8, k = 7: 7 x 38 = 266 input
// planes against 256 x 128 / 148 = 221 output planes per SM).  Balanced mode instead cuts the linearised (column, plane) space
// into one contiguous share per CTA; a share crosses at most two column boundaries, so it costs its planes plus <= 3 halos
// (221 + 3 x 6).  Cut points closer than kMinSeg planes to a column boundary snap onto it (no one-plane segments, whose halo
// would be all of their work).
constexpr int kMinSeg = 4;

__device__ __forceinline__ long long snap_cut(const StackParams& p, long long raw) {
  const long long d = raw % p.D;
  if (d < kMinSeg) return raw - d;
  if (p.D - d < kMinSeg) return raw + (p.D - d);
  return raw;
}

struct ItemIter {
  long long pos, end;
  int item, full_left, full_col;
  __device__ __forceinline__ explicit ItemIter(const StackParams& p) {
    item = blockIdx.x;
    // phase 1: whole columns dealt round-robin (CTAs b and b + 1 stream ADJACENT columns at the same depth, so their H / W
    // halos meet in L2: with contiguous shares alone the neighbours drift apart in depth and the 7^3 layers' DRAM traffic
    // rose from 1.07x to 1.75x of the algorithmic bytes); phase 2: the remaining columns as one contiguous share per CTA
    const long long cols = p.total_planes / p.D;
    full_left = static_cast<int>(cols / gridDim.x);
    full_col = blockIdx.x;
    const long long base = static_cast<long long>(full_left) * gridDim.x * p.D, rest = p.total_planes - base;
    pos = base + snap_cut(p, static_cast<long long>(blockIdx.x) * rest / gridDim.x);
    end = (blockIdx.x + 1 == gridDim.x) ? p.total_planes : base + snap_cut(p, static_cast<long long>(blockIdx.x + 1) * rest / gridDim.x);
  }
  template <int KS>
  __device__ __forceinline__ void column(const StackParams& p, int col, Item& it) {
    constexpr int pad = (KS - 1) / 2;
    int t = col;
    const int wg = t % p.wgroups; t /= p.wgroups;
    const int th = t % p.tiles_h;
    it.n = t / p.tiles_h;
    it.z0 = max(0, it.d0 - pad);
    it.z1 = min(p.D - 1, it.d1 - 1 + pad);
    it.h0 = th * 16;
    it.w0 = wg * p.T * 8;
    it.ntile = min(p.T, p.tiles_w - wg * p.T);
  }
  template <int KS>
  __device__ __forceinline__ bool next(const StackParams& p, Item& it) {
    if (!p.balanced) {
      if (item >= p.num_items) return false;
      it = decode_item<KS>(p, item);
      item += gridDim.x;
      return true;
    }
    if (full_left > 0) {
      it.d0 = 0;
      it.d1 = p.D;
      column<KS>(p, full_col, it);
      full_col += gridDim.x;
      --full_left;
      return true;
    }
    if (pos >= end) return false;
    it.d0 = static_cast<int>(pos % p.D);
    it.d1 = static_cast<int>(min(static_cast<long long>(p.D), it.d0 + (end - pos)));
    column<KS>(p, static_cast<int>(pos / p.D), it);
    pos += it.d1 - it.d0;
    return true;
  }
};

template <int KS, int TT>
__global__ void __launch_bounds__(kStackThreads, 1)
conv3d_stack_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ StackParams p) {
  constexpr int pad = (KS - 1) / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  __shared__ uint64_t done_bar[kMaxSlots], free_bar[kMaxSlots];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float stat_acc[kStackEpiWarps][32][2];
  __shared__ float s_scale[32], s_shift[32];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N0 = p.cout;
  const int NO = p.fold ? (N0 >> 1) : N0;     // output channels

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_in);
    constexpr uint32_t n_issuers = TT >= 2 ? 2 : 1;
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], n_issuers); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], n_issuers); }
    for (int i = 0; i <= KS; ++i) { mbar_init(&done_bar[i], n_issuers); mbar_init(&free_bar[i], kStackEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
  for (int i = threadIdx.x; i < kStackEpiWarps * 32 * 2; i += kStackThreads) (&stat_acc[0][0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < NO; i += kStackThreads) { s_scale[i] = p.scale[i]; s_shift[i] = p.shift[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  constexpr int G = KS + 1;                               // ring slots per tile (power of two for k = 3, 7)
  constexpr uint32_t gmask = G - 1;
  constexpr uint32_t gshift = (G == 8) ? 3 : 2;
  const uint32_t tap_b_bytes = 32u * G * N0;              // one tap: [2][G*N0][8] fp16
  const size_t chunk_w_halfs = static_cast<size_t>(KS) * KS * 16 * G * N0;
  const size_t rot_w_halfs = chunk_w_halfs * p.n_chunks;

  if (warp == 0) {
    // ===================================================================== producer
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t pc_base = 0;
    Item it;
    for (ItemIter iter(p); iter.template next<KS>(p, it);) {
      for (int dz = it.z0; dz <= it.z1; ++dz) {
        const uint32_t rot = (pc_base + static_cast<uint32_t>(dz - pad - it.d0)) & gmask;   // slot of (virtual) plane dz-pad
        for (int c = 0; c < p.n_chunks; ++c) {
          if (!mbar_wait_relaxed(&a_empty[ia], pa ^ 1, p.err_flag)) goto teardown;
          if (elect_one()) {
            if (p.debug & 2) {
              mbar_arrive(&a_full[ia]);
            } else {
              mbar_arrive_expect_tx(&a_full[ia], p.a_bytes);
              tma_load_5d(smem + static_cast<size_t>(ia) * p.a_stride, &tmap_in, &a_full[ia], 0, it.w0 - pad, it.h0 - pad, dz,
                          it.n * p.cb_total_in + p.chunk_cb[c]);
            }
          }
          __syncwarp();
          if (++ia == p.a_stages) { ia = 0; pa ^= 1; }
          const __half* wsrc = p.wpack + rot * rot_w_halfs + static_cast<size_t>(c) * chunk_w_halfs;
          for (int g = 0; g < p.n_bst; ++g) {
            const int t0 = g * p.tps;
            const int cnt = min(p.tps, KS * KS - t0);
            if (!mbar_wait_relaxed(&b_empty[ib], pb ^ 1, p.err_flag)) goto teardown;
            if (elect_one()) {
              const uint32_t bytes = static_cast<uint32_t>(cnt) * tap_b_bytes;
              if (p.debug & 1) {
                mbar_arrive(&b_full[ib]);
              } else {
                mbar_arrive_expect_tx(&b_full[ib], bytes);
                bulk_load_1d(smem + p.b_off + static_cast<size_t>(ib) * p.b_stride, wsrc + static_cast<size_t>(t0) * (tap_b_bytes / 2), bytes,
                             &b_full[ib]);
              }
            }
            __syncwarp();
            if (++ib == p.b_stages) { ib = 0; pb ^= 1; }
          }
        }
      }
      pc_base += static_cast<uint32_t>(it.d1 - it.d0);
    }
  } else if (warp == 1 || (warp == kStackMmaWarpB && TT >= 2)) {
    // ===================================================================== MMA issuers
    // One elected thread per issuing warp runs this role.  Issue cost is ~5.5 cycles per SASS instruction of THAT
    // thread (scripts/micro/mma_rate.cu), so the interior-plane path is fully unrolled per kernel row, and for
    // T >= 2 the CTA's tiles are split between two issuing warps (on different SM sub-partitions): both follow
    // the same barrier protocol, every hand-off barrier counts two arrivals.
    constexpr int TTs = TT >= 2 ? TT / 2 : TT;                 // tiles per issuer
    const int t_off = (warp == 1) ? 0 : TTs;
    if (elect_one()) {
    const uint32_t idesc0 = make_idesc_f16(128, 0);
    const uint32_t idesc_full = idesc0 | (static_cast<uint32_t>((G * N0) >> 3) << 17);
    const uint32_t a_lo_c = (static_cast<uint32_t>(p.PH * p.PWw) & 0x3FFFu) << 16;   // LBO: c8-block pitch
    const uint32_t a_hi = (static_cast<uint32_t>(p.PWw) & 0x3FFFu) | (1u << 14);     // SBO: patch row pitch
    const uint32_t b_lo_c = (static_cast<uint32_t>(G * N0) & 0x3FFFu) << 16;         // LBO: k-half pitch = rows*16 B
    const uint32_t b_hi = 8u | (1u << 14);                                           // SBO: 128 B
    const uint32_t tap_b16 = tap_b_bytes >> 4;
    const uint32_t tile_cols = static_cast<uint32_t>(G * N0);
    const uint32_t smem16 = smem_u32(smem) >> 4;
    const uint32_t a_stride16 = p.a_stride >> 4, b_stride16 = p.b_stride >> 4, b_off16 = p.b_off >> 4;
    const uint32_t tm0 = tmem_base + static_cast<uint32_t>(t_off) * tile_cols;       // this issuer's first tile
    const uint32_t a_toff = static_cast<uint32_t>(t_off) * 8;
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t pc_base = 0;                      // running plane counter of this item's plane d0
    Item it;
    for (ItemIter iter(p); iter.template next<KS>(p, it);) {
      const int my_tiles = max(0, min(TTs, it.ntile - t_off));
      const bool all_tiles = (my_tiles == TTs);
      for (int dz = it.z0; dz <= it.z1; ++dz) {
        // output planes fed by this input plane, oldest first
        const int p_lo = max(dz - pad, it.d0), p_hi = min(dz + pad, it.d1 - 1);
        const int nw = p_hi - p_lo + 1;
        const int n_new = (dz == it.z0) ? nw : ((dz + pad <= it.d1 - 1) ? 1 : 0);
        const int n_old = nw - n_new;
        const uint32_t pc_lo = pc_base + static_cast<uint32_t>(p_lo - it.d0);
        const int slot_lo = static_cast<int>(pc_lo & gmask);
        const bool fast = (nw == KS) && all_tiles;            // interior plane: one MMA per tile covers the ring
        // new planes: their ring slot must have been drained by the epilogue
        for (int i = n_old; i < nw; ++i) {
          const uint32_t pc = pc_lo + i;
          if (!mbar_wait(&free_bar[pc & gmask], ((pc >> gshift) & 1) ^ 1, p.err_flag)) goto teardown;
        }
        tc_fence_after();
        // MMA pieces = contiguous runs of ring slots with one accumulate flag (B row block == ring slot in the
        // rotated weight copy).  pc_*: regular taps of a clipped window; fp_*: first tap (old | new planes).
        // A window of ring slots wraps at most once, so every range is one or two runs; everything stays in
        // registers (dynamically indexed arrays would live in local memory, and this is the single-thread
        // critical path).
        auto run_of = [&](int a0, int a1, uint32_t& o0, uint32_t& i0, uint32_t& o1, uint32_t& i1) -> int {
          if (a1 <= a0) return 0;
          const int sl = (slot_lo + a0) & static_cast<int>(gmask);
          const int len = min(a1 - a0, G - sl);
          o0 = static_cast<uint32_t>(sl * N0);
          i0 = idesc0 | (static_cast<uint32_t>((len * N0) >> 3) << 17);
          if (a0 + len >= a1) return 1;
          o1 = 0u;                                                   // the wrapped remainder starts at ring slot 0
          i1 = idesc0 | (static_cast<uint32_t>(((a1 - a0 - len) * N0) >> 3) << 17);
          return 2;
        };
        uint32_t pc_o[2] = {0, 0}, pc_i[2] = {0, 0};
        const int npc = run_of(0, nw, pc_o[0], pc_i[0], pc_o[1], pc_i[1]);
        uint32_t fo_o[2] = {0, 0}, fo_i[2] = {0, 0}, fn_o[2] = {0, 0}, fn_i[2] = {0, 0};
        const int nfo = run_of(0, n_old, fo_o[0], fo_i[0], fo_o[1], fo_i[1]);       // old planes: accumulate
        const int nfn = run_of(n_old, nw, fn_o[0], fn_i[0], fn_o[1], fn_i[1]);      // new planes: overwrite
        bool first_tap = true;
        for (int c = 0; c < p.n_chunks; ++c) {
          if (!mbar_wait(&a_full[ia], pa, p.err_flag)) goto teardown;
          const uint32_t a_c = a_lo_c + smem16 + static_cast<uint32_t>(ia) * a_stride16;
          int kh = 0, kw = 0;
          uint32_t a_row = a_c;                                  // descriptor low word of (kh, kw = 0), tile 0
          for (int g = 0; g < p.n_bst; ++g) {
            const int cnt = min(p.tps, KS * KS - g * p.tps);
            if (!mbar_wait(&b_full[ib], pb, p.err_flag)) goto teardown;
            tc_fence_after();
            uint32_t b_lo = b_lo_c + smem16 + b_off16 + static_cast<uint32_t>(ib) * b_stride16;
            int tl = 0;
            if (first_tap) {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                if (q < nfo)
                  for (int t = 0; t < my_tiles; ++t)
                    umma_split(tm0 + t * tile_cols + fo_o[q], a_row + a_toff + kw + t * 8, a_hi, b_lo + fo_o[q], b_hi, fo_i[q], 1u);
                if (q < nfn)
                  for (int t = 0; t < my_tiles; ++t)
                    umma_split(tm0 + t * tile_cols + fn_o[q], a_row + a_toff + kw + t * 8, a_hi, b_lo + fn_o[q], b_hi, fn_i[q], 0u);
              }
              first_tap = false;
              b_lo += tap_b16;
              if (++kw == KS) { kw = 0; ++kh; a_row += p.PWw; }
              tl = 1;
            }
            if (fast && KS == 3 && tl == 0 && cnt == KS * KS) {
              // a whole 3x3 tap plane in one go (no per-row loop / switch on the issuing thread)
#pragma unroll
              for (int r3 = 0; r3 < 3; ++r3)
                issue_taps<3, TTs>(tm0, tile_cols, a_row + a_toff + r3 * p.PWw, a_hi, b_lo + r3 * 3 * tap_b16, b_hi, tap_b16, idesc_full);
              a_row += 3 * p.PWw;
              tl = cnt;
            } else if (fast) {
              while (tl < cnt) {
                const int seg = min(cnt - tl, KS - kw);           // taps left in this kernel row
                const uint32_t a_lo = a_row + a_toff + kw;
                switch (seg) {
                  case 7: issue_taps<7, TTs>(tm0, tile_cols, a_lo, a_hi, b_lo, b_hi, tap_b16, idesc_full); break;
                  case 6: issue_taps<6, TTs>(tm0, tile_cols, a_lo, a_hi, b_lo, b_hi, tap_b16, idesc_full); break;
                  case 5: issue_taps<5, TTs>(tm0, tile_cols, a_lo, a_hi, b_lo, b_hi, tap_b16, idesc_full); break;
                  case 4: issue_taps<4, TTs>(tm0, tile_cols, a_lo, a_hi, b_lo, b_hi, tap_b16, idesc_full); break;
                  case 3: issue_taps<3, TTs>(tm0, tile_cols, a_lo, a_hi, b_lo, b_hi, tap_b16, idesc_full); break;
                  case 2: issue_taps<2, TTs>(tm0, tile_cols, a_lo, a_hi, b_lo, b_hi, tap_b16, idesc_full); break;
                  default: issue_taps<1, TTs>(tm0, tile_cols, a_lo, a_hi, b_lo, b_hi, tap_b16, idesc_full); break;
                }
                tl += seg;
                b_lo += seg * tap_b16;
                kw += seg;
                if (kw == KS) { kw = 0; a_row += p.PWw; }
              }
            } else {
              for (; tl < cnt; ++tl) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                  if (q < npc)
                    for (int t = 0; t < my_tiles; ++t)
                      umma_split(tm0 + t * tile_cols + pc_o[q], a_row + a_toff + kw + t * 8, a_hi, b_lo + pc_o[q], b_hi, pc_i[q], 1u);
                b_lo += tap_b16;
                if (++kw == KS) { kw = 0; a_row += p.PWw; }
              }
            }
            umma_commit(&b_empty[ib]);
            if (++ib == p.b_stages) { ib = 0; pb ^= 1; }
          }
          umma_commit(&a_empty[ia]);
          if (++ia == p.a_stages) { ia = 0; pa ^= 1; }
        }
        // planes completed by this input plane
        if (dz < it.z1) {
          const int pdone = dz - pad;
          if (pdone >= it.d0) umma_commit(&done_bar[(pc_base + static_cast<uint32_t>(pdone - it.d0)) & gmask]);
        } else {
          for (int q = max(it.d0, it.z1 - pad); q < it.d1; ++q)
            umma_commit(&done_bar[(pc_base + static_cast<uint32_t>(q - it.d0)) & gmask]);
        }
      }
      pc_base += static_cast<uint32_t>(it.d1 - it.d0);
    }
    }
  } else if (warp >= 2 && warp < 2 + kStackEpiWarps) {
    // ===================================================================== epilogue (warps 2..9)
    // two warps per TMEM lane quarter; the pair splits the CTA's W tiles (even / odd)
    const int quarter = warp & 3;
    const int ew = warp - 2;
    const int tgrp = ew >> 2;
    const int row = quarter * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    const int mycol = (lane >> 1) & 15;
    const float relu_floor = p.relu ? 0.f : -INFINITY;      // one FMNMX per value instead of a predicated pair
    int cur_n = -1;
    uint32_t pc = 0;
    auto flush_stats = [&](int n) {
      if (p.stats == nullptr || n < 0) return;
      __syncwarp();
      for (int c = lane; c < NO; c += 32) {
        atomicAdd(&p.stats[(static_cast<size_t>(n) * NO + c) * 2 + 0], static_cast<double>(stat_acc[ew][c][0]));
        atomicAdd(&p.stats[(static_cast<size_t>(n) * NO + c) * 2 + 1], static_cast<double>(stat_acc[ew][c][1]));
        stat_acc[ew][c][0] = 0.f;
        stat_acc[ew][c][1] = 0.f;
      }
      __syncwarp();
    };
    float ts1[32], ts2[32];                  // this thread's sum / sum of squares per output channel (NO <= 32)
#pragma unroll
    for (int j = 0; j < 32; ++j) { ts1[j] = 0.f; ts2[j] = 0.f; }
    auto fold_item_stats = [&]() {
      if (p.stats == nullptr) return;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (cc * 16 >= NO) break;
        float a[16], b[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { a[j] = ts1[cc * 16 + j]; b[j] = ts2[cc * 16 + j]; ts1[cc * 16 + j] = 0.f; ts2[cc * 16 + j] = 0.f; }
        const float s1 = colsum16s(a, lane);
        const float s2 = colsum16s(b, lane);
        if ((lane & 1) == 0) {
          stat_acc[ew][cc * 16 + mycol][0] += s1;
          stat_acc[ew][cc * 16 + mycol][1] += s2;
        }
      }
    };
    const size_t plane = static_cast<size_t>(p.D) * p.H * p.W;
    Item it;
    for (ItemIter iter(p); iter.template next<KS>(p, it);) {
      if (it.n != cur_n) { flush_stats(cur_n); cur_n = it.n; }
      const int h = it.h0 + hl;
      for (int d = it.d0; d < it.d1; ++d, ++pc) {
        const int slot = static_cast<int>(pc & gmask);
        if (!mbar_wait_relaxed(&done_bar[slot], (pc >> gshift) & 1, p.err_flag)) goto teardown;
        tc_fence_after();
        for (int t = tgrp; t < it.ntile && !(p.debug & 4); t += 2) {     // debug 4: epilogue only hands the slot back
          const int w = it.w0 + t * 8 + wl;
          const bool valid = (h < p.H) && (w < p.W);
          const size_t vox = (static_cast<size_t>(d) * p.H + h) * p.W + w;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((t * G + slot) * N0);
          for (int c0 = 0; c0 < NO; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            if (p.fold) {
              uint32_t r2[16];
              tmem_ld16(taddr + NO + c0, r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            } else {
              tmem_ld_wait();
            }
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = fmaf(__uint_as_float(r[j]), s_scale[c0 + j], s_shift[c0 + j]);
              x = fmaxf(x, relu_floor);
              v[j] = x;
            }
            if (p.stats != nullptr && valid) {
              // per-thread partial sums over the item; the cross-lane reduction happens once per item (fold_item_stats)
              if (c0 == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { ts1[j] += v[j]; ts2[j] = fmaf(v[j], v[j], ts2[j]); }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { ts1[16 + j] += v[j]; ts2[16 + j] = fmaf(v[j], v[j], ts2[16 + j]); }
              }
            }
            if (valid) {
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                const size_t cb = static_cast<size_t>(it.n) * p.cb_total_out + p.cb_out_off + (c0 >> 3) + b;
                const size_t off = (cb * plane + vox) * 8;
                if (p.out_f32 != nullptr) {
                  float y8[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) y8[j] = v[b * 8 + j];
                  st_global_v8f(p.out_f32 + off, y8);
                }
                if (p.out_hi != nullptr) {
                  __align__(16) __half hi[8];
                  __align__(16) __half lo[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    hi[j] = __float2half_rn(v[b * 8 + j]);
                    lo[j] = __float2half_rn(v[b * 8 + j] - __half2float(hi[j]));
                  }
                  *reinterpret_cast<uint4*>(p.out_hi + off) = *reinterpret_cast<const uint4*>(hi);
                  if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(p.out_lo + off) = *reinterpret_cast<const uint4*>(lo);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&free_bar[slot]);
      }
      fold_item_stats();
    }
    flush_stats(cur_n);
  }

teardown:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ===================================================================================================================
// 3^3, C_out = 16, 3-term operand split folded into N with a SPLIT-HALF column layout (fold == 2).
//
// Why: the folded layout above keeps [W_hi | W_lo] per ring slot, so the x_lo chunks (which only need W_hi) still issue
// N = 128 MMAs against [W_hi | 0].  On a power-capped B200 the MMA stream's cost is its executed columns
// (scripts/micro/mma_mix.cu: N = 128 + 128 per tap 65.4 ms, N = 128 + 64 54.0 ms for the same MMA count), so here a
// tile's 128 TMEM columns are [x.W_hi: slot 0..3 x 16 | x_hi.W_lo: slot 0..3 x 16]: hi chunks issue N = 128 against
// [W_hi rows | W_lo rows], lo chunks N = 64 against the W_hi rows alone.
//
// The ring runs over VIRTUAL planes as well: an item streams input planes z0..z1 and every input plane dz feeds ring
// planes dz-1, dz, dz+1 whether or not they are real output planes of the item (d0 <= q < d1); the epilogue hands the
// slots of virtual planes straight back.  Every input plane is therefore an interior plane — one full-width MMA per
// tap and tile, no clipped windows.  A new plane's slot is not contiguous any more (16 hi + 16 lo columns), so the first
// tap of an input plane is: two N = 16 MMAs with accumulate = 0 on the new plane's columns, then one N = 128 MMA against
// a copy of that tap's weights whose new-plane rows are zero as well (wpack tail: one such block per rotation).  The
// first input plane of an item overwrites the whole ring (all four slots are waited free).
template <int TT>
__global__ void __launch_bounds__(kStackThreads, 1)
conv3d_stack3h_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ StackParams p) {
  constexpr int KS = 3, pad = 1, G = 4;
  constexpr uint32_t kTileCols = 128, kTapBytes = 32u * 128u, kTapB16 = kTapBytes >> 4;
  constexpr size_t kChunkHalfs = 9 * 2 * 128 * 8, kFirstHalfs = 2 * 128 * 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  __shared__ uint64_t done_bar[G], free_bar[G];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float stat_acc[kStackEpiWarps][16][2];
  __shared__ float s_scale[16], s_shift[16];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t n_issuers = TT >= 2 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_in);
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], n_issuers); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], n_issuers); }
    for (int i = 0; i < G; ++i) { mbar_init(&done_bar[i], n_issuers); mbar_init(&free_bar[i], kStackEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
  for (int i = threadIdx.x; i < kStackEpiWarps * 16 * 2; i += kStackThreads) (&stat_acc[0][0][0])[i] = 0.f;
  if (threadIdx.x < 16) { s_scale[threadIdx.x] = p.scale[threadIdx.x]; s_shift[threadIdx.x] = p.shift[threadIdx.x]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const size_t rot_w_halfs = kChunkHalfs * p.n_chunks;
  const int nh = p.n_chunks >> 1;                          // chunks [0, nh): x_hi blocks, [nh, n_chunks): x_lo blocks

  if (warp == 0) {
    // ===================================================================== producer
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t pc_base = 0;
    Item it;
    for (ItemIter iter(p); iter.template next<KS>(p, it);) {
      for (int dz = it.z0; dz <= it.z1; ++dz) {
        const uint32_t rot = (pc_base + static_cast<uint32_t>(dz - it.z0)) & 3u;      // slot of ring plane dz - 1
        for (int c = 0; c < p.n_chunks; ++c) {
          if (!mbar_wait_relaxed(&a_empty[ia], pa ^ 1, p.err_flag)) goto teardown;
          if (elect_one()) {
            if (p.debug & 2) {
              mbar_arrive(&a_full[ia]);
            } else {
              mbar_arrive_expect_tx(&a_full[ia], p.a_bytes);
              tma_load_5d(smem + static_cast<size_t>(ia) * p.a_stride, &tmap_in, &a_full[ia], 0, it.w0 - pad, it.h0 - pad, dz,
                          it.n * p.cb_total_in + p.chunk_cb[c]);
            }
          }
          __syncwarp();
          if (++ia == p.a_stages) { ia = 0; pa ^= 1; }
          if (!mbar_wait_relaxed(&b_empty[ib], pb ^ 1, p.err_flag)) goto teardown;
          if (elect_one()) {
            if (p.debug & 1) {
              mbar_arrive(&b_full[ib]);
            } else {
              uint8_t* sb = smem + p.b_off + static_cast<size_t>(ib) * p.b_stride;
              mbar_arrive_expect_tx(&b_full[ib], 9u * kTapBytes + (c == 0 ? kTapBytes : 0u));
              bulk_load_1d(sb, p.wpack + rot * rot_w_halfs + static_cast<size_t>(c) * kChunkHalfs, 9u * kTapBytes, &b_full[ib]);
              if (c == 0) bulk_load_1d(sb + 9u * kTapBytes, p.wpack + 4 * rot_w_halfs + rot * kFirstHalfs, kTapBytes, &b_full[ib]);
            }
          }
          __syncwarp();
          if (++ib == p.b_stages) { ib = 0; pb ^= 1; }
        }
      }
      pc_base += static_cast<uint32_t>(it.z1 - it.z0 + 3);
    }
  } else if (warp == 1 || (warp == kStackMmaWarpB && TT >= 2)) {
    // ===================================================================== MMA issuers (see conv3d_stack_kernel)
    constexpr int TTs = TT >= 2 ? TT / 2 : TT;
    const int t_off = (warp == 1) ? 0 : TTs;
    if (elect_one()) {
      const uint32_t idesc_hi = make_idesc_f16(128, 128), idesc_lo = make_idesc_f16(128, 64), idesc_16 = make_idesc_f16(128, 16);
      const uint32_t a_lo_c = (static_cast<uint32_t>(p.PH * p.PWw) & 0x3FFFu) << 16;
      const uint32_t a_hi = (static_cast<uint32_t>(p.PWw) & 0x3FFFu) | (1u << 14);
      const uint32_t b_lo_c = 128u << 16;                                 // LBO: k-half pitch = 128 rows x 16 B
      const uint32_t b_hi = 8u | (1u << 14);
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const uint32_t a_stride16 = p.a_stride >> 4, b_stride16 = p.b_stride >> 4, b_off16 = p.b_off >> 4;
      const uint32_t tm0 = tmem_base + static_cast<uint32_t>(t_off) * kTileCols;
      const uint32_t a_toff = static_cast<uint32_t>(t_off) * 8;
      int ia = 0, ib = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t pc_base = 0;
      Item it;
      for (ItemIter iter(p); iter.template next<KS>(p, it);) {
        const int my_tiles = max(0, min(TTs, it.ntile - t_off));
        const bool all_tiles = (my_tiles == TTs);
        for (int dz = it.z0; dz <= it.z1; ++dz) {
          const uint32_t pc_old = pc_base + static_cast<uint32_t>(dz - it.z0);     // ring plane dz - 1
          const uint32_t pc_new = pc_old + 2;                                      // ring plane dz + 1
          const bool item_start = (dz == it.z0);
          if (item_start) {
            for (uint32_t i = 0; i < 4; ++i) {
              const uint32_t pc = pc_base + i;
              if (!mbar_wait(&free_bar[pc & 3u], ((pc >> 2) & 1u) ^ 1u, p.err_flag)) goto teardown;
            }
          } else {
            if (!mbar_wait(&free_bar[pc_new & 3u], ((pc_new >> 2) & 1u) ^ 1u, p.err_flag)) goto teardown;
          }
          tc_fence_after();
          const uint32_t sn16 = (pc_new & 3u) * 16u;
          for (int c = 0; c < p.n_chunks; ++c) {
            if (!mbar_wait(&a_full[ia], pa, p.err_flag)) goto teardown;
            if (!mbar_wait(&b_full[ib], pb, p.err_flag)) goto teardown;
            tc_fence_after();
            const uint32_t a_c = a_lo_c + smem16 + static_cast<uint32_t>(ia) * a_stride16 + a_toff;
            const uint32_t b_lo = b_lo_c + smem16 + b_off16 + static_cast<uint32_t>(ib) * b_stride16;
            const uint32_t idesc = c < nh ? idesc_hi : idesc_lo;
            if (c == 0) {
              // first tap of this input plane
              for (int t = 0; t < my_tiles; ++t) {
                const uint32_t d = tm0 + t * kTileCols, a = a_c + t * 8;
                if (item_start) {
                  umma_split(d, a, a_hi, b_lo, b_hi, idesc_hi, 0u);
                } else {
                  umma_split(d + sn16, a, a_hi, b_lo + sn16, b_hi, idesc_16, 0u);
                  umma_split(d + 64u + sn16, a, a_hi, b_lo + 64u + sn16, b_hi, idesc_16, 0u);
                  umma_split(d, a, a_hi, b_lo + 9u * kTapB16, b_hi, idesc_hi, 1u);
                }
              }
              if (all_tiles) {
                issue_taps<2, TTs>(tm0, kTileCols, a_c + 1, a_hi, b_lo + kTapB16, b_hi, kTapB16, idesc_hi);
                issue_taps<3, TTs>(tm0, kTileCols, a_c + p.PWw, a_hi, b_lo + 3 * kTapB16, b_hi, kTapB16, idesc_hi);
                issue_taps<3, TTs>(tm0, kTileCols, a_c + 2 * p.PWw, a_hi, b_lo + 6 * kTapB16, b_hi, kTapB16, idesc_hi);
              } else {
                for (int tap = 1; tap < 9; ++tap)
                  for (int t = 0; t < my_tiles; ++t)
                    umma_split(tm0 + t * kTileCols, a_c + (tap / 3) * p.PWw + (tap % 3) + t * 8, a_hi, b_lo + tap * kTapB16, b_hi, idesc_hi, 1u);
              }
            } else if (all_tiles) {
#pragma unroll
              for (int r3 = 0; r3 < 3; ++r3)
                issue_taps<3, TTs>(tm0, kTileCols, a_c + r3 * p.PWw, a_hi, b_lo + r3 * 3 * kTapB16, b_hi, kTapB16, idesc);
            } else {
              for (int tap = 0; tap < 9; ++tap)
                for (int t = 0; t < my_tiles; ++t)
                  umma_split(tm0 + t * kTileCols, a_c + (tap / 3) * p.PWw + (tap % 3) + t * 8, a_hi, b_lo + tap * kTapB16, b_hi, idesc, 1u);
            }
            umma_commit(&b_empty[ib]);
            if (++ib == p.b_stages) { ib = 0; pb ^= 1; }
            umma_commit(&a_empty[ia]);
            if (++ia == p.a_stages) { ia = 0; pa ^= 1; }
          }
          // ring planes completed by this input plane: dz - 1, and at the item's last input plane the two younger ones too
          umma_commit(&done_bar[pc_old & 3u]);
          if (dz == it.z1) {
            umma_commit(&done_bar[(pc_old + 1) & 3u]);
            umma_commit(&done_bar[(pc_old + 2) & 3u]);
          }
        }
        pc_base += static_cast<uint32_t>(it.z1 - it.z0 + 3);
      }
    }
  } else if (warp >= 2 && warp < 2 + kStackEpiWarps) {
    // ===================================================================== epilogue (warps 2..9)
    const int quarter = warp & 3;
    const int ew = warp - 2;
    const int tgrp = ew >> 2;
    const int row = quarter * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    const int mycol = (lane >> 1) & 15;
    const float relu_floor = p.relu ? 0.f : -INFINITY;
    int cur_n = -1;
    uint32_t pc = 0;
    auto flush_stats = [&](int n) {
      if (p.stats == nullptr || n < 0) return;
      __syncwarp();
      if (lane < 16) {
        atomicAdd(&p.stats[(static_cast<size_t>(n) * 16 + lane) * 2 + 0], static_cast<double>(stat_acc[ew][lane][0]));
        atomicAdd(&p.stats[(static_cast<size_t>(n) * 16 + lane) * 2 + 1], static_cast<double>(stat_acc[ew][lane][1]));
        stat_acc[ew][lane][0] = 0.f;
        stat_acc[ew][lane][1] = 0.f;
      }
      __syncwarp();
    };
    float ts1[16], ts2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { ts1[j] = 0.f; ts2[j] = 0.f; }
    const size_t plane = static_cast<size_t>(p.D) * p.H * p.W;
    Item it;
    for (ItemIter iter(p); iter.template next<KS>(p, it);) {
      if (it.n != cur_n) { flush_stats(cur_n); cur_n = it.n; }
      const int h = it.h0 + hl;
      for (int q = it.z0 - 1; q <= it.z1 + 1; ++q, ++pc) {
        const uint32_t slot = pc & 3u;
        if (!mbar_wait_relaxed(&done_bar[slot], (pc >> 2) & 1u, p.err_flag)) goto teardown;
        tc_fence_after();
        if (q >= it.d0 && q < it.d1 && !(p.debug & 4)) {
          for (int t = tgrp; t < it.ntile; t += 2) {
            const int w = it.w0 + t * 8 + wl;
            const bool valid = (h < p.H) && (w < p.W);
            const size_t vox = (static_cast<size_t>(q) * p.H + h) * p.W + w;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(t) * kTileCols + slot * 16u;
            uint32_t r[16], r2[16];
            tmem_ld16(taddr, r);
            tmem_ld16(taddr + 64u, r2);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x = fmaf(__uint_as_float(r[j]) + __uint_as_float(r2[j]), s_scale[j], s_shift[j]);
              v[j] = fmaxf(x, relu_floor);
            }
            if (valid) {
              if (p.stats != nullptr) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { ts1[j] += v[j]; ts2[j] = fmaf(v[j], v[j], ts2[j]); }
              }
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                const size_t cb = static_cast<size_t>(it.n) * p.cb_total_out + p.cb_out_off + b;
                const size_t off = (cb * plane + vox) * 8;
                if (p.out_f32 != nullptr) {
                  float y8[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) y8[j] = v[b * 8 + j];
                  st_global_v8f(p.out_f32 + off, y8);
                }
                if (p.out_hi != nullptr) {
                  __align__(16) __half hi[8];
                  __align__(16) __half lo[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    hi[j] = __float2half_rn(v[b * 8 + j]);
                    lo[j] = __float2half_rn(v[b * 8 + j] - __half2float(hi[j]));
                  }
                  *reinterpret_cast<uint4*>(p.out_hi + off) = *reinterpret_cast<const uint4*>(hi);
                  if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(p.out_lo + off) = *reinterpret_cast<const uint4*>(lo);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&free_bar[slot]);
      }
      if (p.stats != nullptr) {          // cross-lane reduction of the item's per-thread partial sums
        float a[16], b[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { a[j] = ts1[j]; b[j] = ts2[j]; ts1[j] = 0.f; ts2[j] = 0.f; }
        const float s1 = colsum16s(a, lane);
        const float s2 = colsum16s(b, lane);
        if ((lane & 1) == 0) {
          stat_acc[ew][mycol][0] += s1;
          stat_acc[ew][mycol][1] += s2;
        }
      }
    }
    flush_stats(cur_n);
  }

teardown:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace dp

extern "C" int dp_conv3d_stack(const void* in_c8, int cb_total_in, const uint8_t* chunk_cb, int n_chunks,
                               const void* wpack_stack, int N, int D, int H, int W, int cout, int k,
                               const float* scale, const float* shift, int relu, float* out_f32, void* out_hi,
                               void* out_lo, int cb_total_out, int cb_out_off, double* stats, int* err_flag,
                               int seg_len, int tiles_per_cta, int fold, cudaStream_t stream) {
  using namespace dp;
  DP_REQUIRE(k == 3 || k == 7, "dp_conv3d_stack: kernel size %d unsupported (3 or 7)", k);
  DP_REQUIRE(cout == 16 || (cout == 32 && !fold), "dp_conv3d_stack: C_out=%d unsupported (16, or 32 without fold)", cout);
  DP_REQUIRE(fold >= 0 && fold <= 2 && (fold != 2 || (k == 3 && n_chunks % 2 == 0)), "dp_conv3d_stack: fold=%d needs k=3 and hi+lo chunks", fold);
  const bool split_half = (fold == 2);           // conv3d_stack3h_kernel: [W_hi | W_lo] halves of a tile's 128 columns
  cout = fold ? 2 * cout : cout;                 // MMA columns per ring slot
  DP_REQUIRE(n_chunks >= 1 && n_chunks <= 96, "dp_conv3d_stack: n_chunks=%d out of range", n_chunks);
  DP_REQUIRE(out_f32 != nullptr || out_hi != nullptr, "dp_conv3d_stack: no output tensor given");
  StackParams p{};
  p.N = N; p.D = D; p.H = H; p.W = W; p.n_chunks = n_chunks; p.cb_total_in = cb_total_in; p.cout = cout; p.fold = fold ? 1 : 0;
  for (int i = 0; i < n_chunks; ++i) p.chunk_cb[i] = chunk_cb[i];
  p.tiles_h = (H + 15) / 16;
  p.tiles_w = (W + 7) / 8;
  // tiles per CTA: share B between as many W tiles as TMEM allows while keeping a ring of >= k+1 slots
  const int G = k + 1;                           // ring slots per tile (8 or 4)
  int T = tiles_per_cta > 0 ? tiles_per_cta : 4;   // adjacent W tiles sharing one weight stream
  while (T > 1 && T * G * cout > 512) T >>= 1;
  while (T > 1 && T > p.tiles_w) T >>= 1;          // T in {1, 2, 4}
  if (T == 3) T = 2;
  p.T = T; p.G = G;
  DP_REQUIRE(T * G * cout <= 512, "dp_conv3d_stack: accumulator ring does not fit TMEM");
  p.wgroups = (p.tiles_w + T - 1) / T;
  // segment length: enough work items to balance the SMs, but long segments amortise the (k-1)-plane halo
  const int sms = sm_count();
  int L = seg_len > 0 ? seg_len : D;
  if (seg_len <= 0) {
    // pick the number of D segments minimising (halo overhead) x (round-robin imbalance over the SMs)
    const long long cols = static_cast<long long>(N) * p.tiles_h * p.wgroups;
    double best = 1e30;
    for (int nseg = 1; nseg <= D; ++nseg) {
      const int len = (D + nseg - 1) / nseg;
      if (len < 4 && nseg > 1) break;
      const int segs = (D + len - 1) / len;
      const long long items = cols * segs;
      const double rounds = static_cast<double>((items + sms - 1) / sms);
      const double cost = rounds * (len + k - 1);          // planes streamed by the busiest SM
      if (cost < best - 1e-9) { best = cost; L = len; }
    }
  }
  p.L = L;
  p.nseg = (D + L - 1) / L;
  p.num_items = N * p.nseg * p.tiles_h * p.wgroups;
  static const bool no_balance = [] { const char* e = getenv("DP_STACK_BALANCED"); return e && atoi(e) == 0; }();   // A/B switch
  p.total_planes = static_cast<long long>(N) * p.tiles_h * p.wgroups * D;
  // balanced shares pay off when a share is long against the snapping granularity; small layers (a few planes per SM)
  // keep the fixed segments (measured at 32^3, batch 8: balanced +25 %)
  p.balanced = (seg_len <= 0 && !no_balance && p.total_planes >= 4LL * kMinSeg * sms) ? 1 : 0;
  p.PH = 16 + k - 1;
  p.PWw = 8 * T + k - 1;
  p.a_bytes = 2u * p.PH * p.PWw * 16u;
  p.a_stride = ((p.a_bytes + 1023u) / 1024u) * 1024u;
  const uint32_t tap_b = 32u * G * cout;
  p.tps = (k == 3) ? 9 : static_cast<int>((30u * 1024u) / tap_b);     // ~30 KB weight stages
  if (p.tps < 1) p.tps = 1;
  if (p.tps > k * k) p.tps = k * k;
  p.n_bst = (k * k + p.tps - 1) / p.tps;
  p.b_bytes = static_cast<uint32_t>(p.tps + (split_half ? 1 : 0)) * tap_b;     // split-half: + the first-tap block
  p.b_stride = ((p.b_bytes + 1023u) / 1024u) * 1024u;
  p.a_stages = 3;
  p.b_off = p.a_stages * p.a_stride;
  int bs = static_cast<int>((205u * 1024u - p.b_off) / p.b_stride);
  if (bs > kMaxBStages) bs = kMaxBStages;
  DP_REQUIRE(bs >= 2, "dp_conv3d_stack: weight stage of %u bytes does not fit twice", p.b_stride);
  p.b_stages = bs;
  p.wpack = static_cast<const __half*>(wpack_stack);
  p.scale = scale; p.shift = shift; p.relu = relu;
  p.out_f32 = out_f32; p.out_hi = static_cast<__half*>(out_hi); p.out_lo = static_cast<__half*>(out_lo);
  p.cb_total_out = cb_total_out; p.cb_out_off = cb_out_off; p.stats = stats; p.err_flag = err_flag;
  { static const int dbg = [] { const char* e = getenv("DP_STACK_DEBUG"); return e ? atoi(e) : 0; }(); p.debug = dbg; }   // read once

  CUtensorMap tmap;
  const uint64_t dims[5] = {8, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(D),
                            static_cast<uint64_t>(N) * cb_total_in};
  const uint64_t strides[4] = {16, static_cast<uint64_t>(W) * 16, static_cast<uint64_t>(H) * W * 16,
                               static_cast<uint64_t>(D) * H * W * 16};
  const uint32_t box[5] = {8, static_cast<uint32_t>(p.PWw), static_cast<uint32_t>(p.PH), 1, 2};
  if (int rc = encode_tiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, in_c8, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
    return rc;

  const size_t smem = static_cast<size_t>(p.b_off) + static_cast<size_t>(p.b_stages) * p.b_stride + 1024;
  typedef void (*KernelFn)(const CUtensorMap, const StackParams);
  KernelFn fn = nullptr;
  if (split_half) fn = T == 4 ? conv3d_stack3h_kernel<4> : (T == 2 ? conv3d_stack3h_kernel<2> : conv3d_stack3h_kernel<1>);
  else if (k == 3) fn = T == 4 ? conv3d_stack_kernel<3, 4> : (T == 2 ? conv3d_stack_kernel<3, 2> : conv3d_stack_kernel<3, 1>);
  else fn = T == 4 ? conv3d_stack_kernel<7, 4> : (T == 2 ? conv3d_stack_kernel<7, 2> : conv3d_stack_kernel<7, 1>);
  if (first_use_on_device(KF_CONV_STACK)) {
    const int max_smem = 212 * 1024;
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack_kernel<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack_kernel<7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack_kernel<7, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack3h_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack3h_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DP_CHECK(cudaFuncSetAttribute(conv3d_stack3h_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  }
  int grid = sms < p.num_items ? sms : p.num_items;
  if (p.balanced) {          // one contiguous share of >= 2 planes per CTA
    const long long want = p.total_planes / 2 > 0 ? p.total_planes / 2 : 1;
    grid = static_cast<int>(want < sms ? want : sms);
  }
  fn<<<grid, kStackThreads, smem, stream>>>(tmap, p);
  DP_CHECK(cudaGetLastError());
  return 0;
}
