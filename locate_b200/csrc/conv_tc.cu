// tcgen05 / TMEM / TMA implicit-GEMM for the convolution family (bf16 operands, fp32 accumulate).
//
//   out[pixel][n] = alpha * sum_taps sum_k A_tap[pixel][k] * Wp[tap][n][k]  (+ bias[n])
//
// * A is the bf16 channels-last activation.  For each tap the 128 rows of a tile are a DENSE box of
//   a (possibly parity-strided) view of the source tensor, so one 4-D TMA load per (tap, 64-channel
//   chunk) stages the tile, 128-byte swizzled, with out-of-range pixels/channels zero-filled by the
//   TMA unit (= the convolution padding):
//     mode 0 (Conv fwd / ConvT dgrad, stride s): in = out*s - pad + t = s*(out + a) + q,
//             (a, q) = divmod(t - pad, s): box at out + a in parity view q (s*s tensor maps);
//     mode 1 (ConvT fwd / Conv dgrad): the CTA owns one output parity phase (py,px) and visits only
//             the taps t = (py + pad) mod s (4 of 16 for 4x4/s2); in = j + (py + pad - t)/s.
// * Wp is the bf16 weight packed per forward call as [tap][n][k] (K-major rows, TMA 2-D tiles).
// * One elected thread issues tcgen05.mma (M=128, N=block_n, K=16) into a TMEM accumulator; smem ring
//   of kStages stages guarded by full/empty mbarriers; tcgen05.commit releases stages and signals the
//   epilogue, which reads TMEM with tcgen05.ld (lane = row) and writes fp32 rows.
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2-5 = epilogue.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // bf16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kMaxViews = 4;

struct TcMaps {
  CUtensorMap a[kMaxViews];
  CUtensorMap b;
};

struct TcParams {
  int batch, dst_w, dst_h;            // extents of the destination (phase) grid the tiles cover
  int out_w, out_h, out_c, ld_out;
  int sp;                             // destination parity step (stride in mode 1, else 1)
  int tile_w, tile_h, tile_b;         // box dims, product = 128
  int tiles_w, tiles_h;
  int block_n, kchunks, stages, splits;
  int kh, kw, stride, pad, mode;
  int rows_per_tap;                   // rows of Wp per tap (= out_c)
  int view_empty;                     // bit v set: parity view v has no pixels
  const float* alpha; const float* bias; void* out;
  int out_bf16;                       // `out` rows are bf16 (activation storage of the tensor-core configuration) instead of fp32
  float* part;                        // split-K partial sums [splits][rows][out_c] (raw accumulators), NULL when splits == 1
  long long part_rows;
};

__device__ __forceinline__ void tap_span(int mode, int s, int pad, int k, int parity, int& t0, int& step, int& cnt) {
  if (mode == 1) {
    t0 = (parity + pad) % s;
    step = s;
    cnt = t0 < k ? (k - t0 + s - 1) / s : 0;
  } else {
    t0 = 0; step = 1; cnt = k;
  }
}
__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

__global__ void __launch_bounds__(192, 4) k_conv_tc(const __grid_constant__ TcMaps maps, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_acc;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = kABytes + b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h; t /= p.tiles_h;
  const int x0 = tw * p.tile_w, y0 = th * p.tile_h, b0 = t * p.tile_b;
  const int n0 = blockIdx.y * p.block_n;
  const int phase = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int py = phase / p.sp, px = phase % p.sp;

  int ty0, tys, tyc, tx0, txs, txc;
  tap_span(p.mode, p.stride, p.pad, p.kh, py, ty0, tys, tyc);
  tap_span(p.mode, p.stride, p.pad, p.kw, px, tx0, txs, txc);
  // split-K: this CTA reduces the flat (tap, k-chunk) range [it_begin, it_end) into its own partial buffer; k_splitk_reduce
  // adds the splits in a fixed order (no floating-point atomics: the forward pass is bit-reproducible)
  const int iters_all = tyc * txc * p.kchunks;
  const int per_split = (iters_all + p.splits - 1) / p.splits;
  const int it_begin = min(iters_all, split * per_split);
  const int iters = min(iters_all, it_begin + per_split) - it_begin;
  const uint32_t tmem_cols = p.block_n <= 32 ? 32 : p.block_n <= 64 ? 64 : p.block_n <= 128 ? 128 : 256;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    tc::mbar_init(&bar_acc, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int v = 0; v < kMaxViews; ++v) tc::tma_prefetch_desc(&maps.a[v]);
    tc::tma_prefetch_desc(&maps.b);
  }
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
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      // (tap, k-chunk) and the ring position advance incrementally: a lone producer thread cannot afford integer
      // divisions per stage (each costs more than the MMAs it feeds)
      const int sh = p.stride == 2 ? 1 : 0;
      int kc = it_begin % p.kchunks, tap = it_begin / p.kchunks;
      int iy = txc > 0 ? tap / txc : 0, ix = txc > 0 ? tap % txc : 0;
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        const int ty = ty0 + iy * tys, tx = tx0 + ix * txs;
        int view = 0, dy, dx;
        if (p.mode == 0) {
          const int oy = ty - p.pad, ox = tx - p.pad;
          dy = oy >> sh; dx = ox >> sh;                       // arithmetic shift = floor division (stride 1 or 2)
          view = ((oy & (p.stride - 1)) << sh) + (ox & (p.stride - 1));
        } else {
          dy = (py + p.pad - ty) >> sh;
          dx = (px + p.pad - tx) >> sh;
        }
        const int cb = ((p.view_empty >> view) & 1) ? p.batch : b0;     // empty view: force the box out of range -> zeros
        const int wrow = (ty * p.kw + tx) * p.rows_per_tap + n0;
        tc::mbar_wait(&bar_empty[s], ph ^ 1u);
        uint8_t* sa = smem + s * stage_bytes;
        tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)stage_bytes);
        tc::tma_load_4d(sa, &maps.a[view], &bar_full[s], kc * kBlockK, x0 + dx, y0 + dy, cb);
        tc::tma_load_2d(sa + kABytes, &maps.b, &bar_full[s], kc * kBlockK, wrow);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
        if (++kc == p.kchunks) { kc = 0; if (++ix == txc) { ix = 0; ++iy; } }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_bf16(kBlockM, p.block_n, 0, 0);
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        tc::mbar_wait(&bar_full[s], ph);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + s * stage_bytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t ad = tc::smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t bd = tc::smem_desc_sw128(sb + k * 32, 16, 1024);
          tc::umma_bf16(tmem_base, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        tc::umma_commit(&bar_empty[s]);          // stage reusable once these MMAs have read it
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (iters > 0) tc::umma_commit(&bar_acc);   // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int wl = row % p.tile_w, hl = (row / p.tile_w) % p.tile_h, bl = row / (p.tile_w * p.tile_h);
    const int x = x0 + wl, y = y0 + hl, b = b0 + bl;
    const int oy = y * p.sp + py, ox = x * p.sp + px;
    const bool valid = x < p.dst_w && y < p.dst_h && b < p.batch && oy < p.out_h && ox < p.out_w;
    const size_t row_off = ((size_t)(b * p.out_h + oy) * p.out_w + ox) * p.ld_out;
    float* dst = reinterpret_cast<float*>(p.out) + row_off;
    __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const bool vec = p.out_bf16 ? ((p.ld_out & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                                : ((p.ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    if (iters > 0) {
      tc::mbar_wait(&bar_acc, 0);
      tc::tc_fence_after();
    }
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      float v[16];
      __syncwarp();                                // tcgen05.ld is .sync.aligned: the whole warp issues it together
      if (iters > 0) {
        tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.0f;
      }
      const int n = n0 + c0;
      if (valid && n < p.out_c) {
        if (p.splits > 1) {
          float* pr = p.part + ((size_t)split * p.part_rows + ((size_t)(b * p.out_h + oy) * p.out_w + ox)) * p.out_c + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) if (n + i < p.out_c) pr[i] = v[i];
          continue;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] * alpha + ((p.bias && n + i < p.out_c) ? __ldg(p.bias + n + i) : 0.0f);
        if (p.out_bf16) {
          if (vec && n + 15 < p.out_c) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i], v[i + 1]), h1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]), h3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(dst16 + n + i) = pk;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (n + i < p.out_c) dst16[n + i] = __float2bfloat16(v[i]);
          }
        } else if (vec && n + 15 < p.out_c) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + n + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (n + i < p.out_c) dst[n + i] = v[i];
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

// ---- host: tensor maps -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return LB_EUNSUPPORTED;
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LB_OK : LB_EINVAL;
}

int pow2_ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }

bool tc_geom_ok(const lb_conv_geom* g) {
  if (!g) return false;
  if (g->stride != 1 && g->stride != 2) return false;
  if (g->ld_in % 8) return false;                          // TMA: 16-byte global strides (pad the row, not the channels)
  if (g->in_c < 1 || g->out_c < 1 || g->kh * g->kw > 1024) return false;
  return true;
}

}  // namespace

extern "C" int lb_conv_tc_supported(const lb_conv_geom* g) { return tc_geom_ok(g) ? 1 : 0; }

// packed weight rows are padded to a multiple of 8 bf16 (16 bytes)
extern "C" size_t lb_conv_tc_packed_elems(const lb_conv_geom* g) {
  if (!g) return 0;
  const size_t kpad = (size_t)(g->in_c + 7) / 8 * 8;
  return (size_t)g->kh * g->kw * g->out_c * kpad;
}

// One thread = one (n, k) pair: its taps are contiguous in the master layout (64-100 bytes read in one go), and for each
// tap consecutive threads write consecutive k of one packed row.
template <bool kVec4>
__global__ void __launch_bounds__(256) k_pack_weight(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int taps_h, int taps_w,
                                                    int n_rows, int k, int kpad, long long w_sk, long long w_sn, long long w_sty,
                                                    long long w_stx, int items, LbFastDiv d_kpad) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
    int n, kk;
    lb_fast_divmod(d_kpad, i, n, kk);
    const float* src = w + kk * w_sk + n * w_sn;
    __nv_bfloat16* dst = out + (size_t)n * kpad + kk;
    const size_t tap_stride = (size_t)n_rows * kpad;
    if (kVec4) {                       // the taps of one (n, k) pair are one contiguous, 16-byte aligned run of the master weight
      const int taps = taps_h * taps_w;
      for (int t = 0; t < taps; t += 4) {
        const float4 v = kk < k ? __ldcs(reinterpret_cast<const float4*>(src + t)) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[(size_t)t * tap_stride] = __float2bfloat16(v.x);
        dst[(size_t)(t + 1) * tap_stride] = __float2bfloat16(v.y);
        dst[(size_t)(t + 2) * tap_stride] = __float2bfloat16(v.z);
        dst[(size_t)(t + 3) * tap_stride] = __float2bfloat16(v.w);
      }
    } else {
      for (int ty = 0; ty < taps_h; ++ty)
        for (int tx = 0; tx < taps_w; ++tx)
          dst[(size_t)(ty * taps_w + tx) * tap_stride] = __float2bfloat16(kk < k ? __ldg(src + ty * w_sty + tx * w_stx) : 0.0f);
    }
  }
}
extern "C" int lb_conv_tc_pack(const float* w, void* packed, const lb_conv_geom* g, lb_stream_t s) {
  LB_REQUIRE(w && packed && g);
  const int kpad = (g->in_c + 7) / 8 * 8;
  const long long items = (long long)g->out_c * kpad;
  LB_REQUIRE(items < (1ll << 31) - (1ll << 24));
  const int taps = g->kh * g->kw;
  const bool vec = taps % 4 == 0 && g->w_stx == 1 && g->w_sty == g->kw && g->w_sk % 4 == 0 && g->w_sn % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(w) & 15) == 0;
  if (vec)
    lb_launch(k_pack_weight<true>, lb_grid_1d((size_t)items, 256), 256, 0, lb_s(s), w, reinterpret_cast<__nv_bfloat16*>(packed), g->kh, g->kw,
                                                                            g->out_c, g->in_c, kpad, g->w_sk, g->w_sn, g->w_sty,
                                                                            g->w_stx, (int)items, lb_make_fastdiv(kpad));
  else
    lb_launch(k_pack_weight<false>, lb_grid_1d((size_t)items, 256), 256, 0, lb_s(s), w, reinterpret_cast<__nv_bfloat16*>(packed), g->kh, g->kw,
                                                                             g->out_c, g->in_c, kpad, g->w_sk, g->w_sn, g->w_sty,
                                                                             g->w_stx, (int)items, lb_make_fastdiv(kpad));
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- every weight pack of a model in ONE launch ----------------------------------------------------------------------
// The optimizer moves all weights of a model at once, so all of its packs go stale at once: ~125 k_pack_weight launches
// of a few microseconds each per step (the launch-bound share of a step at the reference's own batch sizes).  The host
// keeps one record per (weight, direction) in a device table plus a list of (record, first item) chunks; CTA = chunk.
namespace {
struct PackRec {
  const float* w; __nv_bfloat16* out;
  long long w_sk, w_sn, w_sty, w_stx;
  int taps_h, taps_w, n_rows, k, kpad, items, vec;
  LbFastDiv d_kpad;
};
constexpr int kPackChunk = 2048;       // items per CTA
__global__ void __launch_bounds__(256) k_pack_weight_batched(const PackRec* __restrict__ recs, const int2* __restrict__ chunks) {
  lb_pdl_enter();
  __shared__ PackRec r;
  const int2 ch = __ldg(chunks + blockIdx.x);
  if (threadIdx.x == 0) r = recs[ch.x];
  __syncthreads();
  const int end = min(r.items, ch.y + kPackChunk);
  const size_t tap_stride = (size_t)r.n_rows * r.kpad;
  const int taps = r.taps_h * r.taps_w;
  for (int i = ch.y + threadIdx.x; i < end; i += 256) {
    int n, kk;
    lb_fast_divmod(r.d_kpad, i, n, kk);
    const float* src = r.w + kk * r.w_sk + n * r.w_sn;
    __nv_bfloat16* dst = r.out + (size_t)n * r.kpad + kk;
    if (r.vec) {
      for (int t = 0; t < taps; t += 4) {
        const float4 v = kk < r.k ? __ldcs(reinterpret_cast<const float4*>(src + t)) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[(size_t)t * tap_stride] = __float2bfloat16(v.x);
        dst[(size_t)(t + 1) * tap_stride] = __float2bfloat16(v.y);
        dst[(size_t)(t + 2) * tap_stride] = __float2bfloat16(v.z);
        dst[(size_t)(t + 3) * tap_stride] = __float2bfloat16(v.w);
      }
    } else {
      for (int ty = 0; ty < r.taps_h; ++ty)
        for (int tx = 0; tx < r.taps_w; ++tx)
          dst[(size_t)(ty * r.taps_w + tx) * tap_stride] = __float2bfloat16(kk < r.k ? __ldg(src + ty * r.w_sty + tx * r.w_stx) : 0.0f);
    }
  }
}
}  // namespace
extern "C" int lb_pack_rec_bytes(void) { return (int)sizeof(PackRec); }
extern "C" int lb_pack_chunk_items(void) { return kPackChunk; }
// fills one host-side record (lb_pack_rec_bytes() bytes) for lb_conv_tc_pack_batched; returns the record's item count
extern "C" int lb_pack_rec_fill(const float* w, void* packed, const lb_conv_geom* g, void* rec_host) {
  if (!w || !packed || !g || !rec_host) return LB_EINVAL;
  PackRec r;
  memset(&r, 0, sizeof r);
  r.kpad = (g->in_c + 7) / 8 * 8;
  const long long items = (long long)g->out_c * r.kpad;
  if (items >= (1ll << 31) - (1ll << 24)) return LB_EINVAL;
  const int taps = g->kh * g->kw;
  r.w = w; r.out = reinterpret_cast<__nv_bfloat16*>(packed);
  r.w_sk = g->w_sk; r.w_sn = g->w_sn; r.w_sty = g->w_sty; r.w_stx = g->w_stx;
  r.taps_h = g->kh; r.taps_w = g->kw; r.n_rows = g->out_c; r.k = g->in_c; r.items = (int)items;
  r.vec = (taps % 4 == 0 && g->w_stx == 1 && g->w_sty == g->kw && g->w_sk % 4 == 0 && g->w_sn % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(w) & 15) == 0) ? 1 : 0;
  r.d_kpad = lb_make_fastdiv(r.kpad);
  memcpy(rec_host, &r, sizeof r);
  return (int)items;
}
// recs_dev: n records as filled above; chunks_dev: n_chunks pairs (record index, first item), lb_pack_chunk_items() items each
extern "C" int lb_conv_tc_pack_batched(const void* recs_dev, const void* chunks_dev, int n_chunks, lb_stream_t s) {
  LB_REQUIRE(recs_dev && chunks_dev && n_chunks > 0);
  lb_launch(k_pack_weight_batched, n_chunks, 256, 0, lb_s(s), reinterpret_cast<const PackRec*>(recs_dev),
            reinterpret_cast<const int2*>(chunks_dev));
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_conv_tc_gemm_ex(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, float* out32,
                                  void* out16, void* out16a, int ld_out16, const void* aux, int ld_aux, int aux_dtype, int flags,
                                  const lb_conv_geom* g, lb_stream_t s);
extern "C" int lb_conv_tc_ex_supported(const lb_conv_geom* g, int out32_used, int ld_out16, int ld_aux, int aux_dtype);
extern "C" int lb_conv_tc_gemm_ws(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, void* out,
                                  const lb_conv_geom* g, void* work, size_t work_bytes, int out_dtype, lb_stream_t s);

// Live taps of the first tile along one axis (host replica of the kernels' tap walk): a tap whose source box lies
// entirely in the padding contributes nothing, and the persistent kernel skips it.
static int live_taps_axis(const lb_conv_geom* g, int k, int tile, int in_extent) {
  int live = 0;
  const int s = g->stride;
  for (int t = 0; t < k; ++t) {
    int d, extent;
    if (g->mode == 0) {
      const int off = t - g->pad;
      d = off >= 0 ? off / s : -((-off + s - 1) / s);
      const int q = off - d * s;
      extent = (in_extent - q + s - 1) / s;
    } else {
      if (((t - g->pad) % s + s) % s != 0) continue;   // tap of another output phase (phase 0 is representative)
      d = (g->pad - t) / s;
      extent = in_extent;
    }
    if (d < extent && d + tile > 0) ++live;
  }
  return live;
}

// Which kernel runs a plain (fp32 out) GEMM: the persistent one wins for few-tap tiles (1x1, the 4 live taps of a 4x4/s2
// phase: the pipeline never drains between tiles) and whenever it can skip dead taps (full-extent feature-attention
// kernels, 5x5 kernels on 2x2 maps); many-tap small-channel tiles keep k_conv_tc, whose 3-4 co-resident CTAs per SM
// hide more TMA latency, and weight-bound shapes keep its split-K.
extern "C" int lb_tc2_halo_eligible(const lb_conv_geom* g);
static bool prefer_persistent(const lb_conv_geom* g) {
  if (lb_tc2_halo_eligible(g)) return true;        // many-tap tiles served from one halo tile: the persistent kernel wins
  const int sp = g->mode == 1 ? g->stride : 1;
  const int dst_w = (g->out_w + sp - 1) / sp, dst_h = (g->out_h + sp - 1) / sp;
  const int tw = pow2_ceil(dst_w) < kBlockM ? pow2_ceil(dst_w) : kBlockM;
  const int th = pow2_ceil(dst_h) < kBlockM / tw ? pow2_ceil(dst_h) : kBlockM / tw;
  const int taps_eff = g->mode == 1 ? ((g->kh + sp - 1) / sp) * ((g->kw + sp - 1) / sp) : g->kh * g->kw;
  if (taps_eff <= 4) return true;
  const int ly = live_taps_axis(g, g->kh, th, g->in_h), lx = live_taps_axis(g, g->kw, tw, g->in_w);
  const int live = ly * lx;
  return live * 2 <= taps_eff;
}

// out[row][n] = alpha * (part[0][row][n] + part[1][row][n] + ...) + bias[n], splits added in index order
template <typename T>
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ part, int splits, size_t rows, int out_c, int ld_out,
                                                      const float* __restrict__ alpha, const float* __restrict__ bias,
                                                      T* __restrict__ out) {
  lb_pdl_enter();
  const float a = alpha ? __ldg(alpha) : 1.0f;
  const size_t n = rows * (size_t)out_c;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const size_t r = i / out_c;
    const int c = (int)(i - r * out_c);
    float acc = part[i];
    for (int sp = 1; sp < splits; ++sp) acc += part[(size_t)sp * n + i];
    lb_st1(out + r * ld_out + c, fmaf(acc, a, bias ? __ldg(bias + c) : 0.0f));
  }
}

// split-K factor of the non-persistent kernel for weight-bound layers whose output tiling cannot fill the GPU
// (e.g. 5x5 1024->1024 at 2x2); 1 = no split
static int tc_splits(const lb_conv_geom* g) {
  const int sp = g->mode == 1 ? g->stride : 1;
  const int dst_w = (g->out_w + sp - 1) / sp, dst_h = (g->out_h + sp - 1) / sp;
  const int tw = pow2_ceil(dst_w) < kBlockM ? pow2_ceil(dst_w) : kBlockM;
  const int rest = kBlockM / tw;
  const int th = pow2_ceil(dst_h) < rest ? pow2_ceil(dst_h) : rest;
  const int tb = rest / th;
  int bn = (g->out_c + 15) / 16 * 16;
  if (bn > 128) bn = 128;
  const int n_tiles = (g->out_c + bn - 1) / bn;
  const long long ctas = (long long)((dst_w + tw - 1) / tw) * ((dst_h + th - 1) / th) * ((g->batch + tb - 1) / tb) * n_tiles * sp * sp;
  const int taps_eff = g->mode == 1 ? ((g->kh + sp - 1) / sp) * ((g->kw + sp - 1) / sp) : g->kh * g->kw;
  const int iters_est = taps_eff * ((g->in_c + kBlockK - 1) / kBlockK);
  if (ctas * 2 <= LB_SMS && iters_est >= 8) {
    long long want = (LB_SMS * 2 + ctas - 1) / ctas;
    if (want > iters_est / 4) want = iters_est / 4;
    if (want > 64) want = 64;
    if (want > 1) return (int)want;
  }
  return 1;
}
extern "C" size_t lb_conv_tc_workspace_bytes(const lb_conv_geom* g) {
  if (!tc_geom_ok(g)) return 0;
  const int splits = tc_splits(g);
  return splits > 1 ? (size_t)splits * g->batch * g->out_h * g->out_w * g->out_c * sizeof(float) : 0;
}

extern "C" int lb_conv_tc_gemm(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, float* out,
                               const lb_conv_geom* g, lb_stream_t s) {
  return lb_conv_tc_gemm_ws(in_bf16, w_packed, alpha, bias, out, g, nullptr, 0, LB_F32, s);
}

extern "C" int lb_conv_tc_gemm_ws(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, void* out,
                                  const lb_conv_geom* g, void* work, size_t work_bytes, int out_dtype, lb_stream_t s) {
  LB_REQUIRE(in_bf16 && w_packed && out && g);
  LB_REQUIRE(out_dtype == LB_F32 || out_dtype == LB_BF16);
  if (!tc_geom_ok(g)) return LB_EUNSUPPORTED;
  static const bool v1_only = getenv("LB_TC_V1_ONLY") != nullptr;      // debugging switches are read once, not per launch
  const bool o16 = out_dtype == LB_BF16;
  if (!v1_only && prefer_persistent(g) && lb_conv_tc_ex_supported(g, o16 ? 0 : 1, o16 ? g->ld_out : 0, 0, LB_F32) == 1 &&
      !(reinterpret_cast<uintptr_t>(out) & 15)) {
    const int rc = lb_conv_tc_gemm_ex(in_bf16, w_packed, alpha, bias, o16 ? nullptr : reinterpret_cast<float*>(out), o16 ? out : nullptr,
                                      nullptr, o16 ? g->ld_out : 0, nullptr, 0, LB_F32, 0, g, s);
    if (rc != LB_EUNSUPPORTED) return rc;
  }
  if (reinterpret_cast<uintptr_t>(in_bf16) & 15 || reinterpret_cast<uintptr_t>(w_packed) & 15) return LB_EALIGN;
  TcMaps maps;
  TcParams p;
  p.mode = g->mode; p.stride = g->stride; p.pad = g->pad; p.kh = g->kh; p.kw = g->kw;
  p.sp = g->mode == 1 ? g->stride : 1;
  p.batch = g->batch;
  p.out_w = g->out_w; p.out_h = g->out_h; p.out_c = g->out_c; p.ld_out = g->ld_out;
  p.dst_w = (g->out_w + p.sp - 1) / p.sp; p.dst_h = (g->out_h + p.sp - 1) / p.sp;
  p.tile_w = pow2_ceil(p.dst_w) < kBlockM ? pow2_ceil(p.dst_w) : kBlockM;
  int rest = kBlockM / p.tile_w;
  p.tile_h = pow2_ceil(p.dst_h) < rest ? pow2_ceil(p.dst_h) : rest;
  p.tile_b = rest / p.tile_h;
  p.tiles_w = (p.dst_w + p.tile_w - 1) / p.tile_w;
  p.tiles_h = (p.dst_h + p.tile_h - 1) / p.tile_h;
  const int tiles_b = (g->batch + p.tile_b - 1) / p.tile_b;
  int bn = (g->out_c + 15) / 16 * 16;
  if (bn > 128) bn = 128;
  p.block_n = bn;
  p.kchunks = (g->in_c + kBlockK - 1) / kBlockK;
  p.rows_per_tap = g->out_c;
  p.alpha = alpha; p.bias = bias; p.out = out; p.out_bf16 = o16 ? 1 : 0;
  const int stage_bytes = kABytes + bn * kBlockK * 2;

  // source views: mode 0 with stride 2 -> 4 parity views; otherwise one dense view
  const int nv = (g->mode == 0) ? g->stride * g->stride : 1;
  const int vs = (g->mode == 0) ? g->stride : 1;
  p.view_empty = 0;
  const char* base = reinterpret_cast<const char*>(in_bf16);
  for (int v = 0; v < kMaxViews; ++v) {
    const int vv = v < nv ? v : 0;
    const int qy = vv / vs, qx = vv % vs;
    int vw = (g->in_w - qx + vs - 1) / vs, vh = (g->in_h - qy + vs - 1) / vs;
    if (vw <= 0 || vh <= 0) { if (v < nv) p.view_empty |= 1 << v; vw = vw > 0 ? vw : 1; vh = vh > 0 ? vh : 1; }
    const uint64_t dims[4] = {(uint64_t)g->in_c, (uint64_t)vw, (uint64_t)vh, (uint64_t)g->batch};
    const uint64_t strides[3] = {(uint64_t)vs * g->ld_in * 2, (uint64_t)vs * g->in_w * g->ld_in * 2,
                                 (uint64_t)g->in_h * g->in_w * g->ld_in * 2};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.tile_w, (uint32_t)p.tile_h, (uint32_t)p.tile_b};
    const char* vbase = ((p.view_empty >> v) & 1) ? base : base + ((size_t)qy * g->in_w + qx) * g->ld_in * 2;
    int rc = make_map(&maps.a[v], vbase, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const int kpad = (g->in_c + 7) / 8 * 8;
    const uint64_t dims[2] = {(uint64_t)g->in_c, (uint64_t)g->kh * g->kw * g->out_c};
    const uint64_t strides[1] = {(uint64_t)kpad * 2};
    const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)bn};
    int rc = make_map(&maps.b, w_packed, 2, dims, strides, box);
    if (rc) return rc;
  }
  // Several CTAs per SM hide each other's fixed latencies (TMEM alloc, first TMA round trip, epilogue stores).
  // Shallow-K tiles (few taps x channel chunks): 2-3 stages, up to 4 CTAs per SM (4 x 128 TMEM columns);
  // deep-K tiles: >= 3 stages, 2 CTAs per SM.
  const int taps_eff = g->mode == 1 ? ((g->kh + p.sp - 1) / p.sp) * ((g->kw + p.sp - 1) / p.sp) : g->kh * g->kw;
  const int iters_est = taps_eff * p.kchunks;
  {
    const int budget = iters_est >= 24 ? 100 * 1024 : 56 * 1024;
    int st = budget / stage_bytes;
    if (st < 2) st = 2;
    if (st > 6) st = 6;
    static const int env_stages = getenv("LB_TC_STAGES") ? atoi(getenv("LB_TC_STAGES")) : 0;
    p.stages = env_stages ? env_stages : st;
  }
  const int smem_bytes = p.stages * stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // split-K (tc_splits) needs the caller's workspace for the per-split partial sums; without one the layer runs unsplit
  const int n_tiles = (g->out_c + bn - 1) / bn;
  const size_t rows = (size_t)g->batch * g->out_h * g->out_w;
  p.splits = tc_splits(g);
  if (p.splits > 1 && (!work || work_bytes < (size_t)p.splits * rows * g->out_c * sizeof(float) || (reinterpret_cast<uintptr_t>(work) & 3)))
    p.splits = 1;
  p.part = p.splits > 1 ? reinterpret_cast<float*>(work) : nullptr;
  p.part_rows = (long long)rows;
  dim3 grid(p.tiles_w * p.tiles_h * tiles_b, n_tiles, p.sp * p.sp * p.splits);
  LB_REQUIRE(grid.y <= 65535 && grid.z <= 65535);
  lb_launch(k_conv_tc, grid, 192, smem_bytes, lb_s(s), maps, p);
  LB_LAUNCH_CHECK();
  if (p.splits > 1) {
    LB_DISPATCH(out_dtype, T, lb_launch(k_splitk_reduce<T>, lb_grid_1d(rows * g->out_c, 256), 256, 0, lb_s(s), p.part, p.splits, rows, g->out_c,
                                                                                                    g->ld_out, alpha, bias, lb_p<T>(out)));
    LB_LAUNCH_CHECK();
  }
  return LB_OK;
}

// ---- fp32 -> bf16 producers ---------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256) k_to_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n, F f) {
  lb_pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = lb_ld4(x + 4 * i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(f(v.x), f(v.y)), hi = __floats2bfloat162_rn(f(v.z), f(v.w));
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y + 4 * i) = pk;
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __float2bfloat16(f(x[i]));
}
struct IdentF { __device__ float operator()(float v) const { return v; } };
struct RootTanh4F { __device__ float operator()(float v) const { return lb_roottanh(v); } };
struct RootTanhGF { float ig; __device__ float operator()(float v) const { return lb_roottanh_g(v, ig); } };

extern "C" int lb_cast_bf16(const float* x, void* y, size_t n, lb_stream_t s) {
  LB_REQUIRE(x && y);
  if (n == 0) return LB_OK;
  if (!lb_aligned16(x) || (reinterpret_cast<uintptr_t>(y) & 7)) return LB_EALIGN;
  lb_launch(k_to_bf16<IdentF>, lb_grid_1d((n + 3) / 4, 256), 256, 0, lb_s(s), x, reinterpret_cast<__nv_bfloat16*>(y), n, IdentF{});
  LB_LAUNCH_CHECK();
  return LB_OK;
}
// row-strided producers: dst[r*ld_dst + c] = bf16(f(src[r*ld_src + c])), c < cols.  Used for the gradient of a concat
// slice and for operands whose channel count is not a multiple of 8 (rows padded to 16 bytes for TMA).
template <typename F>
__global__ void __launch_bounds__(256) k_rows_to_bf16(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst,
                                                     int ld_dst, size_t rows, int cols, int vec, F f) {
  lb_pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const int cols4 = cols >> 2;
    const size_t total = rows * (size_t)cols4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const size_t r = i / cols4;
      const int c = (int)(i % cols4) * 4;
      const float4 v = lb_ld4(src + r * ld_src + c);
      __nv_bfloat162 lo = __floats2bfloat162_rn(f(v.x), f(v.y)), hi = __floats2bfloat162_rn(f(v.z), f(v.w));
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = pk;
    }
  } else {
    const size_t total = rows * (size_t)cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const size_t r = i / cols;
      const int c = (int)(i % cols);
      dst[r * ld_dst + c] = __float2bfloat16(f(src[r * ld_src + c]));
    }
  }
}
// growth = 0: plain cast; growth >= 1: RootTanh fused with the cast
extern "C" int lb_cast_bf16_rows(const float* src, int ld_src, void* dst, int ld_dst, int64_t rows, int cols, int growth, lb_stream_t s) {
  LB_REQUIRE(src && dst && rows >= 0 && cols > 0 && ld_src >= cols && ld_dst >= cols && growth >= 0);
  if (rows == 0) return LB_OK;
  const int vec = (!(cols & 3) && !(ld_src & 3) && !(ld_dst & 3) && lb_aligned16(src) && !(reinterpret_cast<uintptr_t>(dst) & 7)) ? 1 : 0;
  const int grid = lb_grid_1d((size_t)rows * (vec ? cols / 4 : cols), 256);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  if (growth == 0) lb_launch(k_rows_to_bf16<IdentF>, grid, 256, 0, lb_s(s), src, ld_src, d, ld_dst, (size_t)rows, cols, vec, IdentF{});
  else if (growth == 4) lb_launch(k_rows_to_bf16<RootTanh4F>, grid, 256, 0, lb_s(s), src, ld_src, d, ld_dst, (size_t)rows, cols, vec, RootTanh4F{});
  else lb_launch(k_rows_to_bf16<RootTanhGF>, grid, 256, 0, lb_s(s), src, ld_src, d, ld_dst, (size_t)rows, cols, vec, RootTanhGF{1.0f / growth});
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_roottanh_fwd_bf16(const float* x, void* y, size_t n, int growth, lb_stream_t s) {
  LB_REQUIRE(x && y && growth >= 1);
  if (n == 0) return LB_OK;
  if (!lb_aligned16(x) || (reinterpret_cast<uintptr_t>(y) & 7)) return LB_EALIGN;
  if (growth == 4)
    lb_launch(k_to_bf16<RootTanh4F>, lb_grid_1d((n + 3) / 4, 256), 256, 0, lb_s(s), x, reinterpret_cast<__nv_bfloat16*>(y), n, RootTanh4F{});
  else
    lb_launch(k_to_bf16<RootTanhGF>, lb_grid_1d((n + 3) / 4, 256), 256, 0, lb_s(s), x, reinterpret_cast<__nv_bfloat16*>(y), n, RootTanhGF{1.0f / growth});
  LB_LAUNCH_CHECK();
  return LB_OK;
}
