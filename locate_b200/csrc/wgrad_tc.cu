// tcgen05 weight-gradient GEMM: dW[n][m][tap] = sum_pixels G_tap[pixel][m] * D[pixel][n]
//   G = gathered operand (x for Conv, dy for ConvTranspose), D = dense operand, both bf16 channels-last,
//   mode-0 gather around D's pixel grid (gathered pixel = dense*stride - pad + tap).
// The reduction axis is the PIXEL axis, which is the slow axis of channels-last data, so both operands
// are MN-major for the tensor core: a TMA box {64 channels, 64 pixels} lands as 64 rows (pixels = K) of
// 128 swizzled bytes (64 channels = M or N), which is exactly the SWIZZLE_128B MN-major canonical layout
// (8-row K atoms 1024 B apart = SBO, 64-channel groups one box apart = LBO).  a_major = b_major = MN.
//
// Two main loops:
//   k_wgrad_tc    one tap per CTA (grid.y = taps): small maps (< 8 pixels wide) and 1x1 layers;
//   k_wgrad_halo  one GROUP of taps per CTA: all taps of one stride-parity view whose accumulators fit TMEM
//                 (taps x block_n <= 512 columns).  The gathered operand is loaded ONCE per pixel tile as a halo tile
//                 (8 + shift extent pixels wide), and every tap's A operand is that tile read through a descriptor
//                 whose start address is offset by the tap's shift -- the dense tile and the halo tile move L2 -> SM
//                 once per pixel tile instead of once per tap (the per-tap kernel was L2 -> SM bound at ~8-10 TB/s
//                 on the 9-, 16- and 25-tap layers of the 64^2 / 128^2 maps).
// Split-K over pixel tiles is DETERMINISTIC: every split writes its own partial [split][tap][n][m] with plain stores
// and k_wgrad_reduce adds the splits in index order while transposing to the master weight layout [n][m][tap]
// (coalesced on both sides through shared memory) -- no floating-point atomics, no zero-fill of the output.
// Warp roles as conv_tc.cu: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2-5 = epilogue.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kM = 128;                 // gathered-operand channels per tile (2 boxes of 64)
constexpr int kBK = 64;                 // pixels per stage
constexpr int kBox = 64;                // channels per TMA box (128 B)
constexpr int kBoxBytes = kBK * kBox * 2;   // 8 KB
constexpr int kMaxViews = 4;

struct WgMaps {
  CUtensorMap g[kMaxViews];             // gathered operand: parity views
  CUtensorMap d;                        // dense operand
};

struct WgParams {
  int batch, dst_w, dst_h;              // dense pixel grid
  int tile_w, tile_h, tile_b;           // pixel box, product = 64
  int tiles_w, tiles_h, tiles_total;    // pixel tiles
  int tiles_per_split;
  int g_c, d_c;                         // channel counts (M axis, N axis)
  int m_tiles, block_n, n_boxes, stages;
  int kw, stride, pad;
  int view_empty;
  float* part;                          // [split][taps][d_c][g_c]
  size_t split_stride;                  // taps * d_c * g_c
};

__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

__global__ void __launch_bounds__(192, 2) k_wgrad_tc(const __grid_constant__ WgMaps maps, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_acc;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = 2 * kBoxBytes;                       // 128 channels
  const int stage_bytes = a_bytes + p.n_boxes * kBoxBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int m_tile = blockIdx.x % p.m_tiles, n_tile = blockIdx.x / p.m_tiles;
  const int m0 = m_tile * kM, n0 = n_tile * p.block_n;
  const int tap = blockIdx.y;
  const int ty = tap / p.kw, tx = tap % p.kw;
  const int t_begin = blockIdx.z * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);
  const int iters = max(0, t_end - t_begin);
  const uint32_t tmem_cols = p.block_n <= 32 ? 32 : p.block_n <= 64 ? 64 : p.block_n <= 128 ? 128 : 256;

  // tap -> parity view + shift (mode 0): in = s*(out + a) + q
  const int ay = floordiv(ty - p.pad, p.stride), ax = floordiv(tx - p.pad, p.stride);
  const int qy = (ty - p.pad) - ay * p.stride, qx = (tx - p.pad) - ax * p.stride;
  const int view = qy * p.stride + qx;
  const bool empty_view = (p.view_empty >> view) & 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&maps.g[view]); tc::tma_prefetch_desc(&maps.d); }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // programmatic dependent launch: the successor may be scheduled only now that this CTA owns its TMEM columns (a
  // successor CTA allocating first, then waiting for this grid, would deadlock the SM's allocator); everything above ran
  // while the predecessor was still draining, nothing below may start before it has completed
  lb_pdl_trigger();
  lb_pdl_wait();

  if (warp == 0) {
    if (tc::elect_one()) {
      // pixel-tile coordinates and the ring position advance incrementally (no divisions in the producer loop)
      int tw = t_begin % p.tiles_w, th = (t_begin / p.tiles_w) % p.tiles_h, tb = t_begin / (p.tiles_w * p.tiles_h);
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        const int x0 = tw * p.tile_w, y0 = th * p.tile_h, b0 = tb * p.tile_b;
        tc::mbar_wait(&bar_empty[s], ph ^ 1u);
        uint8_t* sa = smem + s * stage_bytes;
        // a 64-channel box beyond the operand's channels would be pure zero fill: its accumulator rows / columns are
        // never stored, so it is not loaded at all (the MMA reads stale shared memory there, rows stay independent)
        const bool m_hi = m0 + kBox < p.g_c;
        int n_live = (p.d_c - n0 + kBox - 1) / kBox;
        if (n_live > p.n_boxes) n_live = p.n_boxes;
        tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)((1 + (m_hi ? 1 : 0) + n_live) * kBoxBytes));
        const int gb = empty_view ? p.batch : b0;
        tc::tma_load_4d(sa, &maps.g[view], &bar_full[s], m0, x0 + ax, y0 + ay, gb);
        if (m_hi) tc::tma_load_4d(sa + kBoxBytes, &maps.g[view], &bar_full[s], m0 + kBox, x0 + ax, y0 + ay, gb);
        for (int j = 0; j < n_live; ++j)
          tc::tma_load_4d(sa + a_bytes + j * kBoxBytes, &maps.d, &bar_full[s], n0 + j * kBox, x0, y0, b0);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
        if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tb; } }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_bf16(kM, p.block_n, 1, 1);
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        tc::mbar_wait(&bar_full[s], ph);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + s * stage_bytes);
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // 16 pixels = two 8-row K atoms = 2048 bytes; LBO = next 64-channel box, SBO = next K atom
          const uint64_t ad = tc::smem_desc_sw128(sa + k * 2048, kBoxBytes, 1024);
          const uint64_t bd = tc::smem_desc_sw128(sb + k * 2048, kBoxBytes, 1024);
          tc::umma_bf16(tmem_base, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        tc::umma_commit(&bar_empty[s]);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (iters > 0) tc::umma_commit(&bar_acc);
    }
  } else if (iters > 0) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    tc::mbar_wait(&bar_acc, 0);
    tc::tc_fence_after();
    float* base = p.part + blockIdx.z * p.split_stride + (size_t)tap * p.d_c * p.g_c + m;
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      float v[16];
      __syncwarp();
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (m < p.g_c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + c0 + i;
          if (n < p.d_c) base[(size_t)n * p.g_c] = v[i];       // this split's own partial: plain store
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- tap-group main loop over halo tiles --------------------------------------------------------------------------
constexpr int kMaxGroups = 40;
constexpr int kMaxGroupTaps = 16;

struct WhView { int ay0, ax0, ey, ex, box_tx; };     // smallest tap shift, shift extents, bytes one halo box transfers
struct WhGroup {
  short view, ntaps;
  unsigned char tap[kMaxGroupTaps], dy[kMaxGroupTaps], dx[kMaxGroupTaps];   // tap index, shift relative to (ay0, ax0)
};
struct WhParams {
  int batch, tile_h, tile_b, log_tile_h;
  int tiles_w, tiles_h, tiles_total, tiles_per_split;
  int g_c, d_c, m_tiles, block_n, n_boxes, m_boxes, stages;
  int g_box_bytes, stage_bytes;            // halo box slot (1024-aligned, largest view), one pipeline stage
  int view_empty, ngroups;
  uint32_t tmem_cols;
  float* part;
  size_t split_stride;
  WhView view[kMaxViews];
  WhGroup group[kMaxGroups];
};

__global__ void __launch_bounds__(192, 1) k_wgrad_halo(const __grid_constant__ WgMaps maps, const __grid_constant__ WhParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_acc;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x % p.m_tiles, n_tile = blockIdx.x / p.m_tiles;
  const int m0 = m_tile * kM, n0 = n_tile * p.block_n;
  const WhGroup& grp = p.group[blockIdx.y];
  const int view = grp.view, ntaps = grp.ntaps;
  const WhView vw = p.view[view];
  const int t_begin = blockIdx.z * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);
  const int iters = max(0, t_end - t_begin);
  const bool empty_view = (p.view_empty >> view) & 1;
  const int g_bytes = p.m_boxes * p.g_box_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&maps.g[view]); tc::tma_prefetch_desc(&maps.d); }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // programmatic dependent launch: the successor may be scheduled only now that this CTA owns its TMEM columns (a
  // successor CTA allocating first, then waiting for this grid, would deadlock the SM's allocator); everything above ran
  // while the predecessor was still draining, nothing below may start before it has completed
  lb_pdl_trigger();
  lb_pdl_wait();

  if (warp == 0) {
    if (tc::elect_one()) {
      int tw = t_begin % p.tiles_w, th = (t_begin / p.tiles_w) % p.tiles_h, tb = t_begin / (p.tiles_w * p.tiles_h);
      int s = 0; uint32_t ph = 0;
      const bool m_hi = m0 + kBox < p.g_c;
      int n_live = (p.d_c - n0 + kBox - 1) / kBox;
      if (n_live > p.n_boxes) n_live = p.n_boxes;
      const uint32_t tx_bytes = (uint32_t)((1 + (m_hi ? 1 : 0)) * vw.box_tx + n_live * kBoxBytes);
      for (int it = 0; it < iters; ++it) {
        const int x0 = tw * 8, y0 = th * p.tile_h, b0 = tb * p.tile_b;
        tc::mbar_wait(&bar_empty[s], ph ^ 1u);
        uint8_t* sa = smem + s * p.stage_bytes;
        tc::mbar_arrive_expect_tx(&bar_full[s], tx_bytes);
        const int gb = empty_view ? p.batch : b0;
        tc::tma_load_4d(sa, &maps.g[view], &bar_full[s], m0, x0 + vw.ax0, y0 + vw.ay0, gb);
        if (m_hi) tc::tma_load_4d(sa + p.g_box_bytes, &maps.g[view], &bar_full[s], m0 + kBox, x0 + vw.ax0, y0 + vw.ay0, gb);
        for (int j = 0; j < n_live; ++j)
          tc::tma_load_4d(sa + g_bytes + j * kBoxBytes, &maps.d, &bar_full[s], n0 + j * kBox, x0, y0, b0);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
        if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tb; } }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_bf16(kM, p.block_n, 1, 1);
      const int pitch = 8 + vw.ex;                       // halo pixels per tile row = rows of 128 bytes
      // descriptor halves: hi = SBO | version | swizzle; lo = start >> 4 | LBO << 16.  The halo start address is a
      // multiple of 128 bytes, not of the 1024-byte swizzle atom: the swizzle is a function of the absolute
      // shared-memory address (TMA writes it the same way), so the base_offset field stays 0.
      const uint32_t a_hi = (((uint32_t)(pitch * 128) >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
      const uint32_t b_hi = ((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
      const uint32_t a_lbo = (((uint32_t)p.g_box_bytes >> 4) & 0x3FFF) << 16;
      const uint32_t b_lbo = (((uint32_t)kBoxBytes >> 4) & 0x3FFF) << 16;
      // K step k = pixel rows 2k, 2k+1 of the 8 x tile_h x tile_b tile: halo row of its first pixel
      uint32_t krow[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = 2 * k, b = r >> p.log_tile_h, y = r & (p.tile_h - 1);
        krow[k] = (uint32_t)((b * (p.tile_h + vw.ey) + y) * pitch);
      }
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        tc::mbar_wait(&bar_full[s], ph);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + s * p.stage_bytes);
        const uint32_t sb = sa + (uint32_t)g_bytes;
#pragma unroll 1
        for (int i = 0; i < ntaps; ++i) {
          const uint32_t row0 = (uint32_t)grp.dy[i] * (uint32_t)pitch + grp.dx[i];
          const uint32_t acc = tmem_base + (uint32_t)(i * p.block_n);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a_lo = (((sa + (krow[k] + row0) * 128u) >> 4) & 0x3FFF) | a_lbo;
            const uint32_t b_lo = (((sb + (uint32_t)k * 2048u) >> 4) & 0x3FFF) | b_lbo;
            const uint64_t ad = ((uint64_t)a_hi << 32) | a_lo, bd = ((uint64_t)b_hi << 32) | b_lo;
            tc::umma_bf16(acc, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(&bar_empty[s]);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (iters > 0) tc::umma_commit(&bar_acc);
    }
  } else if (iters > 0) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    tc::mbar_wait(&bar_acc, 0);
    tc::tc_fence_after();
    for (int i = 0; i < ntaps; ++i) {
      float* base = p.part + blockIdx.z * p.split_stride + (size_t)grp.tap[i] * p.d_c * p.g_c + m;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        float v[16];
        __syncwarp();
        tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * p.block_n + c0), v);
        if (m < p.g_c) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c0 + j;
            if (n < p.d_c) base[(size_t)n * p.g_c] = v[j];
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---- ordered split reduction + transposition to the master layout (+ the spectral-norm dot product) ------------------
// part: [splits][taps][d_c][g_c] (m contiguous) -> dwn[(n * g_c + m) * taps + tap].  One WARP per tile (one n, 32 m): it
// walks the taps (coalesced 128-byte reads per split, summed in split order -- all taps' loads in flight together), turns
// the tile through its own slice of shared memory and stores 32 x taps consecutive floats.  With `w` (the master weight)
// the same pass accumulates dot = sum dwn * w, reduced over the grid in fixed order: the spectral-norm correction
// (lb_sn_weight_grad) then needs no pass of its own over dwn and w.
constexpr int kRedTaps = 32;
// Tile form (large layers): unit of work = one n, 32 m, every tap, per WARP, turned through shared memory.
//   TC > 0: the taps are walked in chunks of exactly TC (taps % TC == 0), everything unrolled at compile time: the W loads
//           and every tap's partial load of a chunk are in flight together (a W load inside the store loop serialised an
//           earlier version on the memory latency; a run-time-unrolled one needed 255 registers and 1600 instructions per tile);
//   TC == 0: run-time chunks of <= 32 taps, plain loops (odd tap counts).
template <int TC>
__global__ void __launch_bounds__(256, 2) k_wgrad_reduce_tile(const float* __restrict__ part, float* __restrict__ dwn, const float* __restrict__ w,
                                                           int splits, size_t split_stride, int taps, int d_c, int g_c, int m_chunks,
                                                           int units, double* __restrict__ dot_out, double* __restrict__ stat_work) {
  lb_pdl_enter();
  constexpr int kChunk = TC ? TC : kRedTaps;
  __shared__ float tile_s[8][32][kChunk + 1];
  __shared__ double scratch[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t tap_stride = (size_t)d_c * g_c;
  float (*tile)[kChunk + 1] = tile_s[warp];
  double dacc = 0.0;
  for (int u = blockIdx.x * 8 + warp; u < units; u += gridDim.x * 8) {
    float fpart = 0.0f;
    const int n = u / m_chunks, m0 = (u - n * m_chunks) * 32;
    const int mm = min(32, g_c - m0);
    for (int t0 = 0; t0 < taps; t0 += kChunk) {
      const size_t base = ((size_t)n * g_c + m0) * taps + t0;
      const float* src = part + ((size_t)t0 * d_c + n) * g_c + m0 + lane;
      if (TC) {
        const int cnt = mm * TC;
        float wv[kChunk], r[kChunk];
#pragma unroll
        for (int it = 0; it < kChunk; ++it) {
          const int j = lane + 32 * it, m = j / kChunk, t = j - m * kChunk;
          wv[it] = (w && j < cnt) ? __ldg(w + base + (size_t)m * taps + t) : 0.0f;
        }
#pragma unroll
        for (int t = 0; t < kChunk; ++t) r[t] = lane < mm ? __ldcs(src + t * tap_stride) : 0.0f;
        for (int sp = 1; sp < splits; ++sp) {            // splits in index order
          const float* ss = src + (size_t)sp * split_stride;
#pragma unroll
          for (int t = 0; t < kChunk; ++t) if (lane < mm) r[t] += __ldcs(ss + t * tap_stride);
        }
#pragma unroll
        for (int t = 0; t < kChunk; ++t) tile[lane][t] = r[t];
        __syncwarp();
#pragma unroll
        for (int it = 0; it < kChunk; ++it) {
          const int j = lane + 32 * it, m = j / kChunk, t = j - m * kChunk;
          if (j < cnt) {
            const float v = tile[m][t];
            dwn[base + (size_t)m * taps + t] = v;
            fpart = fmaf(v, wv[it], fpart);
          }
        }
        __syncwarp();
      } else {
        const int tc_ = min(kChunk, taps - t0);
        if (lane < mm) {
          for (int t = 0; t < tc_; ++t) {
            float acc = 0.0f;
            for (int sp = 0; sp < splits; ++sp) acc += __ldcs(src + t * tap_stride + (size_t)sp * split_stride);
            tile[lane][t] = acc;
          }
        }
        __syncwarp();
        for (int j = lane; j < mm * tc_; j += 32) {
          const int m = j / tc_, t = j - m * tc_;
          const float v = tile[m][t];
          const size_t idx = base + (size_t)m * taps + t;
          dwn[idx] = v;
          if (w) fpart = fmaf(v, __ldg(w + idx), fpart);
        }
        __syncwarp();
      }
    }
    dacc += (double)fpart;                              // short fp32 runs (<= taps floats per lane), fp64 across tiles
  }
  if (w) {
    dacc = lb_block_sum(dacc, scratch);
    lb_grid_sum2_ordered(dacc, 0.0, stat_work, dot_out, scratch);
  }
}

// Per-tap form (small layers with many splits): unit of work = (tile, tap) per CTA.  The 8 warps take every 8th split each
// (8 loads in flight per lane) and their sums are combined in warp order -- a fixed association, so still
// bit-reproducible; the 4-byte strided stores are a few KB in total.
__global__ void __launch_bounds__(256) k_wgrad_reduce_tap(const float* __restrict__ part, float* __restrict__ dwn, const float* __restrict__ w,
                                                          int splits, size_t split_stride, int taps, int d_c, int g_c, int m_chunks,
                                                          int units, double* __restrict__ dot_out, double* __restrict__ stat_work) {
  lb_pdl_enter();
  __shared__ float sums[8][32];
  __shared__ double scratch[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double dacc = 0.0;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int tl = u / taps, t = u - tl * taps;
    const int n = tl / m_chunks, m0 = (tl - n * m_chunks) * 32;
    const bool live = m0 + lane < g_c;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float* src = part + ((size_t)t * d_c + n) * g_c + m0 + lane;
      int sp = warp;
      for (; sp + 56 < splits; sp += 64) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += __ldcs(src + (size_t)(sp + 8 * k) * split_stride);
      }
      for (int k = 0; sp < splits; sp += 8, ++k) a[k] += __ldcs(src + (size_t)sp * split_stride);
    }
    sums[warp][lane] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    __syncthreads();
    if (warp == 0 && live) {
      const float v = ((sums[0][lane] + sums[1][lane]) + (sums[2][lane] + sums[3][lane])) +
                      ((sums[4][lane] + sums[5][lane]) + (sums[6][lane] + sums[7][lane]));
      const size_t idx = ((size_t)n * g_c + m0 + lane) * taps + t;
      dwn[idx] = v;
      if (w) dacc += (double)(v * __ldg(w + idx));
    }
    __syncthreads();
  }
  if (w) {
    dacc = lb_block_sum(dacc, scratch);
    lb_grid_sum2_ordered(dacc, 0.0, stat_work, dot_out, scratch);
  }
}

int wg_pow2_ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }
int wg_floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

bool wg_geom_ok(const lb_conv_geom* g) {
  if (!g || g->mode != 0) return false;
  if (g->stride != 1 && g->stride != 2) return false;
  if (g->ld_in % 8 || g->ld_out % 8) return false;         // TMA: 16-byte global strides (rows padded, channels arbitrary)
  if (g->in_c < 1 || g->out_c < 1 || g->kh * g->kw > 1024) return false;
  return true;
}

// Everything the launch and the workspace query have to agree on.
struct WgPlan {
  bool halo;
  int taps, splits, tiles_total, tiles_per_split;
  int tile_w, tile_h, tile_b, tiles_w, tiles_h;
  int m_tiles, n_tiles, block_n, n_boxes, stages, smem_bytes;
  // halo only
  int m_boxes, g_box_bytes, stage_bytes, ngroups, group_taps;
  uint32_t tmem_cols;
  WhView view[kMaxViews];
  int view_taps[kMaxViews];
};

int wg_halo_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LB_WGRAD_HALO");
    mode = e ? atoi(e) : 1;
  }
  return mode;
}

bool wg_plan(const lb_conv_geom* g, WgPlan& pl) {
  const int taps = g->kh * g->kw, vs = g->stride;
  pl.taps = taps;
  pl.m_tiles = (g->in_c + kM - 1) / kM;
  pl.halo = false;
  if (wg_halo_mode() && taps > 1 && taps <= 64 && g->out_w >= 8 && g->out_h >= 2) {
    // taps by parity view, shift bounding box per view
    for (int v = 0; v < kMaxViews; ++v) { pl.view_taps[v] = 0; pl.view[v] = WhView{1 << 20, 1 << 20, -(1 << 20), -(1 << 20), 0}; }
    for (int t = 0; t < taps; ++t) {
      const int ty = t / g->kw, tx = t % g->kw;
      const int ay = wg_floordiv(ty - g->pad, vs), ax = wg_floordiv(tx - g->pad, vs);
      const int v = ((ty - g->pad) - ay * vs) * vs + ((tx - g->pad) - ax * vs);
      WhView& w = pl.view[v];
      ++pl.view_taps[v];
      if (ay < w.ay0) w.ay0 = ay;
      if (ax < w.ax0) w.ax0 = ax;
      if (ay > w.ey) w.ey = ay;                               // max for now
      if (ax > w.ex) w.ex = ax;
    }
    pl.tile_w = 8;
    pl.tile_h = wg_pow2_ceil(g->out_h) < 8 ? wg_pow2_ceil(g->out_h) : 8;
    pl.tile_b = kBK / (8 * pl.tile_h);
    int max_view_taps = 0, box_max = 0;
    bool ok = true;
    for (int v = 0; v < vs * vs; ++v) {
      WhView& w = pl.view[v];
      if (!pl.view_taps[v]) { w = WhView{0, 0, 0, 0, 0}; continue; }
      w.ey -= w.ay0; w.ex -= w.ax0;
      if (w.ex > 8 || w.ey > 8) ok = false;
      w.box_tx = (8 + w.ex) * (pl.tile_h + w.ey) * pl.tile_b * 128;
      if (w.box_tx > box_max) box_max = w.box_tx;
      if (pl.view_taps[v] > max_view_taps) max_view_taps = pl.view_taps[v];
    }
    // channel tile: the padded MMA work per pixel tile is proportional to n_tiles x block_n whatever the taps per group,
    // and a 128 x 128 tile reads shared memory at the full 128 B/clk (A and B are both re-read per MMA), so wide tiles win:
    // the widest of {256, 192, 128} that minimises the padding; taps per group = what still fits TMEM (512 columns).
    int bn = (g->out_c + 15) / 16 * 16;
    if (bn > 256) {
      const int cand[3] = {256, 192, 128};
      int best = 0;
      long long best_cost = 0;
      for (int i = 0; i < 3; ++i) {
        const long long cost = (long long)((g->out_c + cand[i] - 1) / cand[i]) * cand[i];
        if (!best || cost < best_cost) { best = cand[i]; best_cost = cost; }
      }
      bn = best;
    }
    int gt = 512 / bn;
    if (gt > max_view_taps) gt = max_view_taps;
    if (gt > kMaxGroupTaps) gt = kMaxGroupTaps;
    int ngroups = 0;
    for (int v = 0; v < vs * vs; ++v) ngroups += (pl.view_taps[v] + gt - 1) / gt;
    pl.block_n = bn; pl.n_boxes = (bn + kBox - 1) / kBox;
    pl.n_tiles = (g->out_c + bn - 1) / bn;
    pl.group_taps = gt; pl.ngroups = ngroups;
    pl.m_boxes = g->in_c > kBox ? 2 : 1;
    pl.g_box_bytes = (box_max + 1023) / 1024 * 1024;
    pl.stage_bytes = pl.m_boxes * pl.g_box_bytes + pl.n_boxes * kBoxBytes;
    const int slack = 1024 + (pl.m_boxes == 1 ? pl.g_box_bytes : 0);   // a missing upper channel box is read as garbage rows
    pl.stages = (216 * 1024 - slack) / pl.stage_bytes;
    if (pl.stages > 6) pl.stages = 6;
    uint32_t cols = 32;
    while ((int)cols < gt * bn) cols <<= 1;
    pl.tmem_cols = cols;
    if (ok && ngroups <= kMaxGroups && pl.stages >= 2 && cols <= 512) {
      pl.halo = true;
      pl.smem_bytes = pl.stages * pl.stage_bytes + slack;
      pl.tiles_w = (g->out_w + 7) / 8;
      pl.tiles_h = (g->out_h + pl.tile_h - 1) / pl.tile_h;
      const int tiles_b = (g->batch + pl.tile_b - 1) / pl.tile_b;
      pl.tiles_total = pl.tiles_w * pl.tiles_h * tiles_b;
      const long long ctas = (long long)pl.m_tiles * pl.n_tiles * ngroups;
      long long splits = (LB_SMS + ctas - 1) / ctas;
      if (splits > pl.tiles_total) splits = pl.tiles_total;
      if (splits < 1) splits = 1;
      pl.tiles_per_split = (int)((pl.tiles_total + splits - 1) / splits);
      pl.splits = (pl.tiles_total + pl.tiles_per_split - 1) / pl.tiles_per_split;
      return true;
    }
  }
  pl.tile_w = wg_pow2_ceil(g->out_w) < kBK ? wg_pow2_ceil(g->out_w) : kBK;
  const int rest = kBK / pl.tile_w;
  pl.tile_h = wg_pow2_ceil(g->out_h) < rest ? wg_pow2_ceil(g->out_h) : rest;
  pl.tile_b = rest / pl.tile_h;
  pl.tiles_w = (g->out_w + pl.tile_w - 1) / pl.tile_w;
  pl.tiles_h = (g->out_h + pl.tile_h - 1) / pl.tile_h;
  const int tiles_b = (g->batch + pl.tile_b - 1) / pl.tile_b;
  pl.tiles_total = pl.tiles_w * pl.tiles_h * tiles_b;
  int bn = (g->out_c + 63) / 64 * 64;
  if (bn > 256) bn = 256;
  pl.block_n = bn; pl.n_boxes = bn / kBox;
  pl.n_tiles = (g->out_c + bn - 1) / bn;
  const int stage_bytes = (2 + pl.n_boxes) * kBoxBytes;
  pl.stages = (100 * 1024) / stage_bytes;                   // <= ~100 KB: two CTAs per SM
  if (pl.stages < 2) pl.stages = 2;
  if (pl.stages > 6) pl.stages = 6;
  pl.smem_bytes = pl.stages * stage_bytes + 1024;
  const long long ctas = (long long)pl.m_tiles * pl.n_tiles * taps;
  long long splits = (LB_SMS * 2 + ctas - 1) / ctas;
  if (splits > pl.tiles_total) splits = pl.tiles_total;
  if (splits < 1) splits = 1;
  pl.tiles_per_split = (int)((pl.tiles_total + splits - 1) / splits);
  pl.splits = (pl.tiles_total + pl.tiles_per_split - 1) / pl.tiles_per_split;
  return true;
}

}  // namespace

extern "C" int lb_wgrad_tc_supported(const lb_conv_geom* g) { return wg_geom_ok(g) ? 1 : 0; }

// floats of split-K workspace lb_wgrad_tc needs for this geometry (splits x the weight's element count)
extern "C" size_t lb_wgrad_tc_workspace_floats(const lb_conv_geom* g) {
  WgPlan pl;
  if (!wg_geom_ok(g) || !wg_plan(g, pl)) return 0;
  return (size_t)pl.splits * pl.taps * g->out_c * g->in_c;
}

// geom as lb_conv_wgrad: in_* = gathered operand, out_* = dense operand.  dwn: fp32, the master weight's layout
// [out_c][in_c][kh][kw] of this geometry's (dense, gathered) channel pair, OVERWRITTEN.
// w / dot_out / stat_work (all or none): dot_out[0] = sum dwn * w in fixed order (stat_work as lb_norm_stats).
extern "C" int lb_wgrad_tc(const void* gathered_bf16, const void* dense_bf16, float* dwn, const lb_conv_geom* g, float* work,
                           size_t work_floats, const float* w, double* dot_out, double* stat_work, lb_stream_t s) {
  LB_REQUIRE(gathered_bf16 && dense_bf16 && dwn && g && work);
  LB_REQUIRE((w && dot_out && stat_work) || (!w && !dot_out && !stat_work));
  if (!wg_geom_ok(g)) return LB_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(gathered_bf16) & 15) || (reinterpret_cast<uintptr_t>(dense_bf16) & 15)) return LB_EALIGN;
  WgPlan pl;
  if (!wg_plan(g, pl)) return LB_EUNSUPPORTED;
  const size_t numel = (size_t)pl.taps * g->out_c * g->in_c;
  LB_REQUIRE(work_floats >= numel * pl.splits);

  WgMaps maps;
  const int vs = g->stride;
  int view_empty = 0;
  const char* base = reinterpret_cast<const char*>(gathered_bf16);
  for (int v = 0; v < kMaxViews; ++v) {
    const int vv = v < vs * vs ? v : 0;
    const int qy = vv / vs, qx = vv % vs;
    int vw = (g->in_w - qx + vs - 1) / vs, vh = (g->in_h - qy + vs - 1) / vs;
    bool empty = vw <= 0 || vh <= 0;
    if (empty) { if (v < vs * vs) view_empty |= 1 << v; vw = vw > 0 ? vw : 1; vh = vh > 0 ? vh : 1; }
    const uint64_t dims[4] = {(uint64_t)g->in_c, (uint64_t)vw, (uint64_t)vh, (uint64_t)g->batch};
    const uint64_t strides[3] = {(uint64_t)vs * g->ld_in * 2, (uint64_t)vs * g->in_w * g->ld_in * 2,
                                 (uint64_t)g->in_h * g->in_w * g->ld_in * 2};
    const char* vbase = empty ? base : base + ((size_t)qy * g->in_w + qx) * g->ld_in * 2;
    uint32_t box[4] = {(uint32_t)kBox, (uint32_t)pl.tile_w, (uint32_t)pl.tile_h, (uint32_t)pl.tile_b};
    if (pl.halo) { box[1] = (uint32_t)(8 + pl.view[vv].ex); box[2] = (uint32_t)(pl.tile_h + pl.view[vv].ey); }
    int rc = tc::make_map(&maps.g[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, vbase, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)g->out_c, (uint64_t)g->out_w, (uint64_t)g->out_h, (uint64_t)g->batch};
    const uint64_t strides[3] = {(uint64_t)g->ld_out * 2, (uint64_t)g->out_w * g->ld_out * 2, (uint64_t)g->out_h * g->out_w * g->ld_out * 2};
    const uint32_t box[4] = {(uint32_t)kBox, (uint32_t)pl.tile_w, (uint32_t)pl.tile_h, (uint32_t)pl.tile_b};
    int rc = tc::make_map(&maps.d, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dense_bf16, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wgrad_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (pl.halo) {
    WhParams p;
    memset(&p, 0, sizeof(p));
    p.batch = g->batch; p.tile_h = pl.tile_h; p.tile_b = pl.tile_b;
    p.log_tile_h = 0;
    while ((1 << p.log_tile_h) < pl.tile_h) ++p.log_tile_h;
    p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.tiles_total = pl.tiles_total; p.tiles_per_split = pl.tiles_per_split;
    p.g_c = g->in_c; p.d_c = g->out_c; p.m_tiles = pl.m_tiles; p.block_n = pl.block_n; p.n_boxes = pl.n_boxes; p.m_boxes = pl.m_boxes;
    p.stages = pl.stages; p.g_box_bytes = pl.g_box_bytes; p.stage_bytes = pl.stage_bytes;
    p.view_empty = view_empty; p.tmem_cols = pl.tmem_cols;
    p.part = work; p.split_stride = numel;
    for (int v = 0; v < kMaxViews; ++v) p.view[v] = pl.view[v < vs * vs ? v : 0];
    int ng = 0;
    for (int v = 0; v < vs * vs; ++v) {
      int in_group = pl.group_taps;                           // forces a new group at the first tap of the view
      for (int t = 0; t < pl.taps; ++t) {
        const int ty = t / g->kw, tx = t % g->kw;
        const int ay = wg_floordiv(ty - g->pad, vs), ax = wg_floordiv(tx - g->pad, vs);
        if (((ty - g->pad) - ay * vs) * vs + ((tx - g->pad) - ax * vs) != v) continue;
        if (in_group == pl.group_taps) { p.group[ng].view = (short)v; p.group[ng].ntaps = 0; ++ng; in_group = 0; }
        WhGroup& gr = p.group[ng - 1];
        gr.tap[in_group] = (unsigned char)t;
        gr.dy[in_group] = (unsigned char)(ay - pl.view[v].ay0);
        gr.dx[in_group] = (unsigned char)(ax - pl.view[v].ax0);
        gr.ntaps = (short)(++in_group);
      }
    }
    LB_REQUIRE(ng == pl.ngroups);
    p.ngroups = ng;
    dim3 grid(pl.m_tiles * pl.n_tiles, ng, (unsigned)pl.splits);
    LB_REQUIRE(grid.z <= 65535);
    lb_launch(k_wgrad_halo, grid, 192, pl.smem_bytes, lb_s(s), maps, p);
    LB_LAUNCH_CHECK();
  } else {
    WgParams p;
    p.batch = g->batch; p.dst_w = g->out_w; p.dst_h = g->out_h;
    p.tile_w = pl.tile_w; p.tile_h = pl.tile_h; p.tile_b = pl.tile_b;
    p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.tiles_total = pl.tiles_total; p.tiles_per_split = pl.tiles_per_split;
    p.g_c = g->in_c; p.d_c = g->out_c; p.m_tiles = pl.m_tiles; p.block_n = pl.block_n; p.n_boxes = pl.n_boxes; p.stages = pl.stages;
    p.kw = g->kw; p.stride = g->stride; p.pad = g->pad;
    p.view_empty = view_empty;
    p.part = work; p.split_stride = numel;
    dim3 grid(pl.m_tiles * pl.n_tiles, pl.taps, (unsigned)pl.splits);
    LB_REQUIRE(grid.y <= 65535 && grid.z <= 65535);
    lb_launch(k_wgrad_tc, grid, 192, pl.smem_bytes, lb_s(s), maps, p);
    LB_LAUNCH_CHECK();
  }
  const int m_chunks = (g->in_c + 31) / 32;
  const long long tiles = (long long)m_chunks * g->out_c;
  const bool per_tap = tiles < 4096 && pl.splits >= 16;
  const long long units = per_tap ? tiles * pl.taps : tiles;
  LB_REQUIRE(units < (1ll << 31));
  long long rblocks = per_tap ? units : (units + 7) / 8;
  if (rblocks > LB_SMS * 8) rblocks = LB_SMS * 8;       // <= the statistics workspace's grid bound
#define LB_WG_REDUCE(KERNEL)                                                                                                   \
  lb_launch(KERNEL, (unsigned)rblocks, 256, 0, lb_s(s), work, dwn, w, pl.splits, numel, pl.taps, g->out_c, g->in_c, m_chunks, (int)units, \
                                                 dot_out, stat_work)
  if (per_tap) LB_WG_REDUCE(k_wgrad_reduce_tap);
  else if (pl.taps == 1) LB_WG_REDUCE(k_wgrad_reduce_tile<1>);
  else if (pl.taps == 9) LB_WG_REDUCE(k_wgrad_reduce_tile<9>);
  else if (pl.taps == 25) LB_WG_REDUCE(k_wgrad_reduce_tile<25>);
  else if (pl.taps % 16 == 0) LB_WG_REDUCE(k_wgrad_reduce_tile<16>);
  else LB_WG_REDUCE(k_wgrad_reduce_tile<0>);
#undef LB_WG_REDUCE
  LB_LAUNCH_CHECK();
  return LB_OK;
}
