// tcgen05 weight-gradient GEMM: dwp[tap][n][m] += sum_pixels G_tap[pixel][m] * D[pixel][n]
//   G = gathered operand (x for Conv, dy for ConvTranspose), D = dense operand, both bf16 channels-last,
//   mode-0 gather around D's pixel grid (gathered pixel = dense*stride - pad + tap).
// The reduction axis is the PIXEL axis, which is the slow axis of channels-last data, so both operands
// are MN-major for the tensor core: a TMA box {64 channels, 64 pixels} lands as 64 rows (pixels = K) of
// 128 swizzled bytes (64 channels = M or N), which is exactly the SWIZZLE_128B MN-major canonical layout
// (8-row K atoms 1024 B apart = SBO, 64-channel groups one box apart = LBO).  a_major = b_major = MN.
// Output is a tap-major packed fp32 buffer [tap][d0][d1] (= the master layout with the taps moved
// outermost), lanes (M) running over the contiguous index, accumulated with red.global.add (split-K).
// Warp roles as conv_tc.cu: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2-5 = epilogue.
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
  float* dwp;                           // [taps][d_c][g_c]
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
    float* base = p.dwp + (size_t)tap * p.d_c * p.g_c + m;
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      float v[16];
      __syncwarp();
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (m < p.g_c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + c0 + i;
          if (n < p.d_c) atomicAdd(base + (size_t)n * p.g_c, v[i]);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg_encode_fn() {
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
int wg_make_map(CUtensorMap* m, const void* base, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = wg_encode_fn();
  if (!fn) return LB_EUNSUPPORTED;
  cuuint64_t gdim[4]; cuuint64_t gstr[3]; cuuint32_t bx[4]; cuuint32_t es[4];
  for (int i = 0; i < 4; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < 3; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LB_OK : LB_EINVAL;
}
int wg_pow2_ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }

bool wg_geom_ok(const lb_conv_geom* g) {
  if (!g || g->mode != 0) return false;
  if (g->stride != 1 && g->stride != 2) return false;
  if (g->ld_in % 8 || g->ld_out % 8) return false;         // TMA: 16-byte global strides (rows padded, channels arbitrary)
  if (g->in_c < 1 || g->out_c < 1 || g->kh * g->kw > 1024) return false;
  return true;
}

}  // namespace

extern "C" int lb_wgrad_tc_supported(const lb_conv_geom* g) { return wg_geom_ok(g) ? 1 : 0; }

// geom as lb_conv_wgrad: in_* = gathered operand, out_* = dense operand.  dwp: fp32 [kh*kw][out_c][in_c], zeroed by the caller.
extern "C" int lb_wgrad_tc(const void* gathered_bf16, const void* dense_bf16, float* dwp, const lb_conv_geom* g, lb_stream_t s) {
  LB_REQUIRE(gathered_bf16 && dense_bf16 && dwp && g);
  if (!wg_geom_ok(g)) return LB_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(gathered_bf16) & 15) || (reinterpret_cast<uintptr_t>(dense_bf16) & 15)) return LB_EALIGN;
  WgMaps maps;
  WgParams p;
  p.batch = g->batch; p.dst_w = g->out_w; p.dst_h = g->out_h;
  p.tile_w = wg_pow2_ceil(p.dst_w) < kBK ? wg_pow2_ceil(p.dst_w) : kBK;
  int rest = kBK / p.tile_w;
  p.tile_h = wg_pow2_ceil(p.dst_h) < rest ? wg_pow2_ceil(p.dst_h) : rest;
  p.tile_b = rest / p.tile_h;
  p.tiles_w = (p.dst_w + p.tile_w - 1) / p.tile_w;
  p.tiles_h = (p.dst_h + p.tile_h - 1) / p.tile_h;
  const int tiles_b = (g->batch + p.tile_b - 1) / p.tile_b;
  p.tiles_total = p.tiles_w * p.tiles_h * tiles_b;
  p.g_c = g->in_c; p.d_c = g->out_c;
  p.m_tiles = (g->in_c + kM - 1) / kM;
  int bn = (g->out_c + 63) / 64 * 64;
  if (bn > 256) bn = 256;
  p.block_n = bn; p.n_boxes = bn / kBox;
  const int n_tiles = (g->out_c + bn - 1) / bn;
  p.kw = g->kw; p.stride = g->stride; p.pad = g->pad;
  p.dwp = dwp;
  const int stage_bytes = (2 + p.n_boxes) * kBoxBytes;
  p.stages = (100 * 1024) / stage_bytes;                   // <= ~100 KB: two CTAs per SM
  if (p.stages < 2) p.stages = 2;
  if (p.stages > 6) p.stages = 6;
  const int smem_bytes = p.stages * stage_bytes + 1024;
  const int taps = g->kh * g->kw;
  const long long ctas = (long long)p.m_tiles * n_tiles * taps;
  long long splits = (LB_SMS * 2 + ctas - 1) / ctas;
  if (splits > p.tiles_total) splits = p.tiles_total;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (int)((p.tiles_total + splits - 1) / splits);
  splits = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;

  const int vs = g->stride;
  p.view_empty = 0;
  const char* base = reinterpret_cast<const char*>(gathered_bf16);
  const uint32_t box[4] = {(uint32_t)kBox, (uint32_t)p.tile_w, (uint32_t)p.tile_h, (uint32_t)p.tile_b};
  for (int v = 0; v < kMaxViews; ++v) {
    const int vv = v < vs * vs ? v : 0;
    const int qy = vv / vs, qx = vv % vs;
    int vw = (g->in_w - qx + vs - 1) / vs, vh = (g->in_h - qy + vs - 1) / vs;
    bool empty = vw <= 0 || vh <= 0;
    if (empty) { if (v < vs * vs) p.view_empty |= 1 << v; vw = vw > 0 ? vw : 1; vh = vh > 0 ? vh : 1; }
    const uint64_t dims[4] = {(uint64_t)g->in_c, (uint64_t)vw, (uint64_t)vh, (uint64_t)g->batch};
    const uint64_t strides[3] = {(uint64_t)vs * g->ld_in * 2, (uint64_t)vs * g->in_w * g->ld_in * 2,
                                 (uint64_t)g->in_h * g->in_w * g->ld_in * 2};
    const char* vbase = empty ? base : base + ((size_t)qy * g->in_w + qx) * g->ld_in * 2;
    int rc = wg_make_map(&maps.g[v], vbase, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)g->out_c, (uint64_t)g->out_w, (uint64_t)g->out_h, (uint64_t)g->batch};
    const uint64_t strides[3] = {(uint64_t)g->ld_out * 2, (uint64_t)g->out_w * g->ld_out * 2, (uint64_t)g->out_h * g->out_w * g->ld_out * 2};
    int rc = wg_make_map(&maps.d, dense_bf16, dims, strides, box);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(p.m_tiles * n_tiles, taps, (unsigned)splits);
  LB_REQUIRE(grid.y <= 65535 && grid.z <= 65535);
  k_wgrad_tc<<<grid, 192, smem_bytes, lb_s(s)>>>(maps, p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
