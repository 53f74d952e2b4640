// fp32 SIMT gather-GEMM: the general-shape path of the convolution family (any channel count,
// any kernel/stride/pad, Conv and ConvTranspose, fwd / dgrad / wgrad, Linear as the 1x1 case).
// The tcgen05 path (conv_tc.cu) takes over the tensor-core-bound shapes; this kernel keeps the
// odd ones (Cin = 3, Cout = 1/3/29, the style MLP) and is the fp32 parity anchor for the model.
//
// C[m][n] = alpha * sum_taps sum_k A_tap[m][k] * W(tap,k,n) (+bias[n]),  m = output pixel.
//  mode 0: input pixel = out*stride - pad + tap
//  mode 1: input pixel = (out + pad - tap)/stride when exact.  For stride 2 the CTA's 64 rows all have
//          the same output parity ("phase", blockIdx.z) so only the taps of that parity are visited:
//          no multiply-by-zero work (4 of 16 taps for the 4x4/s2 transposed conv).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, register-staged double buffering.
#include "common.cuh"

#define BM 64
#define BN 64
#define BK 16
#define APAD 4

struct GemmP {
  const float* in; const float* w; const float* alpha; const float* bias; float* out;
  int batch, in_h, in_w, in_c, out_h, out_w, out_c;
  int kh, kw, stride, pad, mode, ld_in, ld_out;
  long long w_sk, w_sn, w_sty, w_stx;
  int sp;            // output parity step (stride in mode 1, else 1)
  int ph_h, ph_w;    // output rows / cols per phase
  int m_phase;       // rows of the GEMM per phase = batch*ph_h*ph_w
  int vec_a, vec_c, b_kfast;
};

__device__ __forceinline__ void tap_range(int mode, int s, int pad, int k, int parity, int& t0, int& step, int& cnt) {
  if (mode == 1) {
    t0 = (parity + pad) % s;
    step = s;
    cnt = t0 < k ? (k - t0 + s - 1) / s : 0;
  } else {
    t0 = 0; step = 1; cnt = k;
  }
}

__global__ void __launch_bounds__(256) k_conv_gemm(const GemmP p) {
  lb_pdl_enter();
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int py = blockIdx.z / p.sp, px = blockIdx.z % p.sp;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  int ty0, tys, tyc, tx0, txs, txc;
  tap_range(p.mode, p.stride, p.pad, p.kh, py, ty0, tys, tyc);
  tap_range(p.mode, p.stride, p.pad, p.kw, px, tx0, txs, txc);
  const int kchunks = (p.in_c + BK - 1) / BK;
  const int iters = tyc * txc * kchunks;

  // A-load role: row ar (one output pixel), 4 consecutive k starting at akq
  const int ar = tid >> 2, akq = (tid & 3) * 4;
  const int am = m0 + ar;
  const bool a_row_ok = am < p.m_phase;
  int ab = 0, aoy = 0, aox = 0;
  bool a_pix_ok = a_row_ok;
  if (a_row_ok) {
    ab = am / (p.ph_h * p.ph_w);
    const int rem = am - ab * (p.ph_h * p.ph_w);
    aoy = (rem / p.ph_w) * p.sp + py;
    aox = (rem % p.ph_w) * p.sp + px;
    a_pix_ok = aoy < p.out_h && aox < p.out_w;     // ragged last phase row/col when out % stride != 0
  }

  float a_reg[4], b_reg[4];

  auto load_tiles = [&](int it) {
    const int kc = it % kchunks;
    const int tap = it / kchunks;
    const int ty = ty0 + (tap / txc) * tys, tx = tx0 + (tap % txc) * txs;
    const int k0 = kc * BK;
    // ---- A
    bool ok = a_pix_ok;
    int iy, ix;
    if (p.mode == 0) {
      iy = aoy * p.stride - p.pad + ty;
      ix = aox * p.stride - p.pad + tx;
    } else {
      const int vy = aoy + p.pad - ty, vx = aox + p.pad - tx;
      ok = ok && vy >= 0 && vx >= 0;
      iy = vy / p.stride;
      ix = vx / p.stride;
    }
    ok = ok && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
    const int k = k0 + akq;
    if (ok && p.vec_a && k + 3 < p.in_c) {
      const float4 v = lb_ld4(p.in + ((size_t)(ab * p.in_h + iy) * p.in_w + ix) * p.ld_in + k);
      a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
    } else {
      const float* src = ok ? p.in + ((size_t)(ab * p.in_h + iy) * p.in_w + ix) * p.ld_in : p.in;
#pragma unroll
      for (int i = 0; i < 4; ++i) a_reg[i] = (ok && k + i < p.in_c) ? __ldg(src + k + i) : 0.0f;
    }
    // ---- B
    const float* wt = p.w + ty * p.w_sty + tx * p.w_stx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk, nn;
      if (p.b_kfast) { kk = tid & 15; nn = (tid >> 4) + 16 * i; } else { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int gk = k0 + kk, gn = n0 + nn;
      b_reg[i] = (gk < p.in_c && gn < p.out_c) ? __ldg(wt + gk * p.w_sk + gn * p.w_sn) : 0.0f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][akq + i][ar] = a_reg[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk, nn;
      if (p.b_kfast) { kk = tid & 15; nn = (tid >> 4) + 16 * i; } else { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
      Bs[buf][kk][nn] = b_reg[i];
    }
  };

  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  if (iters > 0) {
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
      const int buf = it & 1;
      if (it + 1 < iters) load_tiles(it + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][tm]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (it + 1 < iters) store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= p.m_phase) continue;
    const int b = m / (p.ph_h * p.ph_w);
    const int rem = m - b * (p.ph_h * p.ph_w);
    const int oy = (rem / p.ph_w) * p.sp + py, ox = (rem % p.ph_w) * p.sp + px;
    if (oy >= p.out_h || ox >= p.out_w) continue;
    float* dst = p.out + ((size_t)(b * p.out_h + oy) * p.out_w + ox) * p.ld_out;
    const int n = n0 + tn;
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = acc[i][j] * alpha + ((p.bias && n + j < p.out_c) ? __ldg(p.bias + n + j) : 0.0f);
    if (p.vec_c && n + 3 < p.out_c) {
      lb_st4(dst + n, make_float4(r[0], r[1], r[2], r[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (n + j < p.out_c) dst[n + j] = r[j];
    }
  }
}

static int check_geom(const lb_conv_geom* g) {
  LB_REQUIRE(g);
  LB_REQUIRE(g->batch > 0 && g->in_h > 0 && g->in_w > 0 && g->in_c > 0 && g->out_h > 0 && g->out_w > 0 && g->out_c > 0);
  LB_REQUIRE(g->kh > 0 && g->kw > 0 && g->stride > 0 && g->pad >= 0 && (g->mode == 0 || g->mode == 1));
  LB_REQUIRE(g->ld_in >= g->in_c && g->ld_out >= g->out_c);
  return LB_OK;
}

extern "C" int lb_conv_gemm(const float* in, const float* w, const float* alpha, const float* bias, float* out,
                            const lb_conv_geom* g, lb_stream_t s) {
  LB_REQUIRE(in && w && out);
  int rc = check_geom(g);
  if (rc) return rc;
  GemmP p;
  p.in = in; p.w = w; p.alpha = alpha; p.bias = bias; p.out = out;
  p.batch = g->batch; p.in_h = g->in_h; p.in_w = g->in_w; p.in_c = g->in_c;
  p.out_h = g->out_h; p.out_w = g->out_w; p.out_c = g->out_c;
  p.kh = g->kh; p.kw = g->kw; p.stride = g->stride; p.pad = g->pad; p.mode = g->mode;
  p.ld_in = g->ld_in; p.ld_out = g->ld_out;
  p.w_sk = g->w_sk; p.w_sn = g->w_sn; p.w_sty = g->w_sty; p.w_stx = g->w_stx;
  p.sp = g->mode == 1 ? g->stride : 1;
  p.ph_h = (g->out_h + p.sp - 1) / p.sp; p.ph_w = (g->out_w + p.sp - 1) / p.sp;
  const long long mph = (long long)g->batch * p.ph_h * p.ph_w;
  LB_REQUIRE(mph < (1ll << 31) && (long long)g->batch * g->in_h * g->in_w < (1ll << 31));
  p.m_phase = (int)mph;
  p.vec_a = (g->in_c % 4 == 0 && g->ld_in % 4 == 0 && lb_aligned16(in)) ? 1 : 0;
  p.vec_c = (g->ld_out % 4 == 0 && lb_aligned16(out)) ? 1 : 0;
  p.b_kfast = g->w_sk <= g->w_sn ? 1 : 0;
  dim3 grid((p.m_phase + BM - 1) / BM, (g->out_c + BN - 1) / BN, p.sp * p.sp);
  LB_REQUIRE(grid.y <= 65535);
  lb_launch(k_conv_gemm, grid, 256, 0, lb_s(s), p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- weight gradient --------------------------------------------------------------------------
// dw(ty,tx,kg,kd) += sum_m G[m@tap][kg] * D[m][kd]; tile = 64 kg x 64 kd, reduction over pixels m.
struct WgradP {
  const float* gath; const float* dense; float* dw;
  int batch, g_h, g_w, g_c, d_h, d_w, d_c;
  int kh, kw, stride, pad, ld_g, ld_d;
  long long w_sk, w_sn, w_sty, w_stx;
  int m_total, m_split, tiles_g, vec_g, vec_d;
};

__global__ void __launch_bounds__(256) k_conv_wgrad(const WgradP p) {
  lb_pdl_enter();
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tile_g = blockIdx.x % p.tiles_g, tile_d = blockIdx.x / p.tiles_g;
  const int g0 = tile_g * BM, d0 = tile_d * BN;
  const int ty = blockIdx.y / p.kw, tx = blockIdx.y % p.kw;
  const int ms = blockIdx.z * p.m_split;
  const int me = min(p.m_total, ms + p.m_split);
  const int iters = (me - ms + BK - 1) / BK;

  const int lr = tid >> 4, lc = (tid & 15) * 4;     // load role: pixel lr of the chunk, 4 channels at lc
  float a_reg[4], b_reg[4];

  auto load_tiles = [&](int it) {
    const int m = ms + it * BK + lr;
    bool ok = m < me;
    int b = 0, oy = 0, ox = 0;
    if (ok) {
      b = m / (p.d_h * p.d_w);
      const int rem = m - b * (p.d_h * p.d_w);
      oy = rem / p.d_w; ox = rem % p.d_w;
    }
    // dense row
    {
      const int c = d0 + lc;
      const float* src = p.dense + (size_t)m * p.ld_d;
      if (ok && p.vec_d && c + 3 < p.d_c) {
        const float4 v = lb_ld4(src + c);
        b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) b_reg[i] = (ok && c + i < p.d_c) ? __ldg(src + c + i) : 0.0f;
      }
    }
    // gathered row
    {
      const int iy = oy * p.stride - p.pad + ty, ix = ox * p.stride - p.pad + tx;
      const bool gok = ok && iy >= 0 && iy < p.g_h && ix >= 0 && ix < p.g_w;
      const int c = g0 + lc;
      const float* src = gok ? p.gath + ((size_t)(b * p.g_h + iy) * p.g_w + ix) * p.ld_g : p.gath;
      if (gok && p.vec_g && c + 3 < p.g_c) {
        const float4 v = lb_ld4(src + c);
        a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) a_reg[i] = (gok && c + i < p.g_c) ? __ldg(src + c + i) : 0.0f;
      }
    }
  };
  auto store_tiles = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = make_float4(a_reg[0], a_reg[1], a_reg[2], a_reg[3]);
    *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  if (iters > 0) {
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
      const int buf = it & 1;
      if (it + 1 < iters) load_tiles(it + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][tm]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (it + 1 < iters) store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
  float* base = p.dw + ty * p.w_sty + tx * p.w_stx;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kg = g0 + tm + i;
    if (kg >= p.g_c) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kd = d0 + tn + j;
      if (kd < p.d_c) atomicAdd(base + kg * p.w_sk + kd * p.w_sn, acc[i][j]);
    }
  }
}

extern "C" int lb_conv_wgrad(const float* gathered, const float* dense, float* dw, const lb_conv_geom* g, lb_stream_t s) {
  LB_REQUIRE(gathered && dense && dw);
  int rc = check_geom(g);
  if (rc) return rc;
  LB_REQUIRE(g->mode == 0);
  WgradP p;
  p.gath = gathered; p.dense = dense; p.dw = dw;
  p.batch = g->batch; p.g_h = g->in_h; p.g_w = g->in_w; p.g_c = g->in_c;
  p.d_h = g->out_h; p.d_w = g->out_w; p.d_c = g->out_c;
  p.kh = g->kh; p.kw = g->kw; p.stride = g->stride; p.pad = g->pad;
  p.ld_g = g->ld_in; p.ld_d = g->ld_out;
  p.w_sk = g->w_sk; p.w_sn = g->w_sn; p.w_sty = g->w_sty; p.w_stx = g->w_stx;
  const long long mt = (long long)g->batch * g->out_h * g->out_w;
  LB_REQUIRE(mt < (1ll << 31) && (long long)g->batch * g->in_h * g->in_w < (1ll << 31));
  p.m_total = (int)mt;
  p.tiles_g = (g->in_c + BM - 1) / BM;
  const int tiles_d = (g->out_c + BN - 1) / BN;
  const int taps = g->kh * g->kw;
  LB_REQUIRE(taps <= 65535);
  const long long tiles = (long long)p.tiles_g * tiles_d * taps;
  long long splits = (LB_SMS * 4 + tiles - 1) / tiles;
  const long long max_splits = (mt + BK * 4 - 1) / (BK * 4);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long per = (mt + splits - 1) / splits;
  per = (per + BK - 1) / BK * BK;
  splits = (mt + per - 1) / per;
  p.m_split = (int)per;
  p.vec_g = (g->in_c % 4 == 0 && g->ld_in % 4 == 0 && lb_aligned16(gathered)) ? 1 : 0;
  p.vec_d = (g->out_c % 4 == 0 && g->ld_out % 4 == 0 && lb_aligned16(dense)) ? 1 : 0;
  dim3 grid(p.tiles_g * tiles_d, taps, (unsigned)splits);
  lb_launch(k_conv_wgrad, grid, 256, 0, lb_s(s), p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- column sums (bias gradients) ---------------------------------------------------------------
template <typename T>
__global__ void k_colsum(const T* __restrict__ x, long long rows, int cols, int ld, float* __restrict__ out, int chunk, int tc, int tp) {
  lb_pdl_enter();
  if (threadIdx.x >= tc * tp) return;
  const int cl = threadIdx.x % tc, pl = threadIdx.x / tc;
  const long long r0 = (long long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
  for (int c = cl; c < cols; c += tc) {
    float acc = 0.0f;
    for (long long r = r0 + pl; r < r1; r += tp) acc += lb_ld1(x + r * ld + c);
    atomicAdd(out + c, acc);
  }
}
// vector form (16-byte loads; the scalar kernel keeps 2 bytes per thread in flight and ran at a third of the HBM rate).
// `x` is the 16-byte aligned start of the rows' summed span of `cols` columns (a multiple of the vector width); the wanted
// columns are [c_first, c_first + c_count) of it -- a concat slice (the D stem's 29 channels at column 3 of 32-wide rows) is
// summed through its aligned superset.
template <typename T>
__global__ void __launch_bounds__(256) k_colsum_v(const T* __restrict__ x, long long rows, int cols, int ld, float* __restrict__ out,
                                                 int chunk, int cv, int tp, int c_first, int c_count) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  extern __shared__ float s_part[];
  const int cl = threadIdx.x % cv, pl = threadIdx.x / cv;
  const long long r0 = (long long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
  float acc[N];
#pragma unroll
  for (int k = 0; k < N; ++k) acc[k] = 0.0f;
  const bool active = pl < tp;
  if (active) {
#pragma unroll 4
    for (long long r = r0 + pl; r < r1; r += tp) {
      float v[N];
      lb_ldv(x + r * ld + cl * N, v);
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] += v[k];
    }
  }
  lb_colsum_flush<N>(acc, active, s_part, cl, pl, tp, cols, 1.0f, out, c_first, c_count);
}
// aligned superset of columns [0, cols) of rows starting at x: *shift elements before x, *span columns wide; false if none
template <typename T>
static bool colsum_vec_span(const T* x, int cols, int ld, int* shift, int* span) {
  constexpr int N = LbV<T>::N;
  if (ld % N) return false;
  const uintptr_t a = reinterpret_cast<uintptr_t>(x);
  if (a % sizeof(T)) return false;
  *shift = (int)((a & 15) / sizeof(T));
  *span = (*shift + cols + N - 1) / N * N;
  return *span <= ld && *span / N <= 256;          // the span never leaves the row it starts in (shift + cols <= ld by construction of a slice)
}
extern "C" int lb_colsum(const void* x, int64_t rows, int cols, int ld, float* out, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && out && rows > 0 && cols > 0 && ld >= cols);
  const LbColShape sh = lb_col_shape(cols);
  long long chunks = LB_SMS * 4;
  long long chunk = (rows + chunks - 1) / chunks;
  if (chunk < sh.tp) chunk = sh.tp;
  chunks = (rows + chunk - 1) / chunk;
  LB_DISPATCH(dtype, T, {
    int shift = 0, span = 0;
    if (colsum_vec_span(lb_cp<T>(x), cols, ld, &shift, &span)) {
      const int cv = span / LbV<T>::N, tp = 256 / cv;
      chunks = LB_SMS * 4;
      chunk = (rows + chunks - 1) / chunks;
      if (chunk < 4 * tp) chunk = 4 * tp;
      chunks = (rows + chunk - 1) / chunk;
      lb_launch(k_colsum_v<T>, (unsigned)chunks, 256, (size_t)tp * span * sizeof(float), lb_s(s), lb_cp<T>(x) - shift, rows, span, ld, out,
                (int)chunk, cv, tp, shift, cols);
    } else {
      lb_launch(k_colsum<T>, (unsigned)chunks, sh.threads, 0, lb_s(s), lb_cp<T>(x), rows, cols, ld, out, (int)chunk, sh.tc, sh.tp);
    }
  });
  LB_LAUNCH_CHECK();
  return LB_OK;
}
