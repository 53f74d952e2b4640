// Direct fp32 kernels for the layers with a tiny channel count on one side -- the discriminator stem (5x5/s2 3->3,
// 1x1 3->32, 1x1 3->29 + concat) in all three directions.  A 128x64 tensor-core tile makes no sense for K = 3, and the
// layers are HBM-bound on their wide side (AI <= 16 flop/B), so the job is to touch every activation once with the
// neighbouring elementwise ops fused: RootTanh on the input (forward), RootTanh'(x) on the output (input gradient), the
// CatModule copy.
//
//   lb_conv_small:       out[p][n] = alpha * sum_{tap,k} act(in[p@tap][k]) * W(tap,k,n) (+bias[n])   [* act'(xpre[p][n])]
//   lb_conv_small_wgrad: dw(tap,kg,kd) += sum_p act(gathered[p@tap][kg]) * dense[p][kd]
// Geometry and weight addressing are those of lb_conv_gemm / lb_conv_wgrad (include/locate_b200.h).
//
// Every kernel is "one thread = one pixel, everything in registers": the narrow side (<= 4 channels) is a handful of
// scalars, the wide side (<= 64) lives in up to 64 accumulators, the tiny weight sits in shared memory (alpha folded in)
// and is read as warp-wide broadcasts.  There is no shared-memory staging of activations and no barrier after the weight
// load: earlier tile-staged versions of these kernels spent their time in block barriers and index arithmetic
// (0.6-2.3 TB/s, profiles/r1_small_micro_b192.txt).
#include "common.cuh"

#define SMALL_THREADS 256
#define SMALL_MAX_W 4096      // floats of weight held in shared memory

// `in` / `out` / `xpre` are untyped here: the NARROW side (<= 4 channels) is always fp32, the wide side has the storage
// type the kernel is instantiated for (fp32 or bf16).
struct SmallP {
  const void* in; const float* w; const float* alpha; const float* bias; const void* xpre; void* out;
  int batch, in_h, in_w, in_c, out_h, out_w, out_c;
  int kh, kw, stride, pad, mode, ld_in, ld_out, ld_xpre;
  long long w_sk, w_sn, w_sty, w_stx;
  int growth_in;      // > 0: RootTanh(growth) applied to every input element on load
  int growth_out;     // > 0: result multiplied by RootTanh'(xpre[p][n]); -1: by xpre[p][n] itself (a stored derivative)
  int cat;            // rows of `out` are [in (in_c, copied) | conv (out_c)]  (CatModule, merge.py:10-16)
  int pointwise;      // 1x1, stride 1, pad 0: input pixel == output pixel
  int pixels;         // batch*out_h*out_w (< 2^31, checked by the host)
  LbFastDiv d_w, d_h;
};

__device__ __forceinline__ float small_act(float v, int growth) {
  return growth == 4 ? lb_roottanh(v) : (growth > 0 ? lb_roottanh_g(v, 1.0f / growth) : v);
}
__device__ __forceinline__ float small_dact(float v, int growth) {
  if (growth < 0) return v;                            // xpre already holds RootTanh' (stored by the forward pass)
  return growth == 4 ? lb_roottanh_grad(v) : lb_roottanh_grad_g(v, 1.0f / growth);
}
// source pixel of tap t along one axis (stride 1 or 2): false if the tap does not reach a source pixel
__device__ __forceinline__ bool small_src(int mode, int stride, int sh, int pad, int o, int t, int extent, int& i) {
  if (mode == 0) {
    i = o * stride - pad + t;
  } else {
    const int v = o + pad - t;
    if (v < 0 || (v & (stride - 1))) return false;
    i = v >> sh;
  }
  return i >= 0 && i < extent;
}

// ---- narrow INPUT (in_c <= 4): forward of the stem layers, input gradient of a wide -> 3 layer -------------------
// NC output columns per thread (4 for the RGB -> RGB layers, else a multiple of 16).  Shared weight wsm[tap][k][NC] holds
// alpha * W (and, for a concat, identity columns that copy the input), bsm the bias per column.
template <int NC, typename TO>
__global__ void __launch_bounds__(SMALL_THREADS) k_small_narrow_in(const SmallP p) {
  lb_pdl_enter();
  const float* p_in = reinterpret_cast<const float*>(p.in);
  TO* p_out = reinterpret_cast<TO*>(p.out);
  const TO* p_xpre = reinterpret_cast<const TO*>(p.xpre);
  extern __shared__ float4 sm4[];
  float* wsm = reinterpret_cast<float*>(sm4);
  const int taps = p.kh * p.kw;
  float* bsm = wsm + taps * p.in_c * NC;
  const int c0 = p.cat ? p.in_c : 0;                 // first conv column
  const int cols = c0 + p.out_c;
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
  for (int i = threadIdx.x; i < taps * p.in_c * NC; i += blockDim.x) {
    const int n = i % NC, k = (i / NC) % p.in_c, tap = i / (NC * p.in_c);
    float v = 0.0f;
    if (n < c0) v = (n == k) ? 1.0f : 0.0f;          // concat: column n copies input channel n (1x1 layers only)
    else if (n < cols) v = alpha * __ldg(p.w + k * p.w_sk + (n - c0) * p.w_sn + (tap / p.kw) * p.w_sty + (tap % p.kw) * p.w_stx);
    wsm[i] = v;
  }
  for (int n = threadIdx.x; n < NC; n += blockDim.x) bsm[n] = (p.bias && n >= c0 && n < cols) ? __ldg(p.bias + n - c0) : 0.0f;
  __syncthreads();
  const int sh = p.stride == 2 ? 1 : 0;
  const int gi = p.growth_in, go = p.growth_out;
  const bool vec_out = !(p.ld_out & 3) && lb_vec4_ok(p_out);
  const bool vec16_out = NC % LbV<TO>::N == 0 && !(cols % LbV<TO>::N) && !(p.ld_out % LbV<TO>::N) && lb_vec_ok(p_out);
  const int stride_t = gridDim.x * blockDim.x;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < p.pixels; pix += stride_t) {
    float acc[NC];
#pragma unroll
    for (int n = 0; n < NC; ++n) acc[n] = bsm[n];
    auto tap_fma = [&](const float* src, const float* wt) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < p.in_c) {
          const float a = small_act(__ldg(src + k), gi);
          const float4* w4 = reinterpret_cast<const float4*>(wt + k * NC);
#pragma unroll
          for (int j = 0; j < NC / 4; ++j) {
            const float4 wv = w4[j];
            acc[4 * j + 0] = fmaf(a, wv.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(a, wv.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(a, wv.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(a, wv.w, acc[4 * j + 3]);
          }
        }
      }
    };
    if (p.pointwise) {
      tap_fma(p_in + (size_t)pix * p.ld_in, wsm);
    } else {
      int t, ox, oy, b;
      lb_fast_divmod(p.d_w, pix, t, ox);
      lb_fast_divmod(p.d_h, t, b, oy);
      for (int ty = 0; ty < p.kh; ++ty) {
        int iy;
        if (!small_src(p.mode, p.stride, sh, p.pad, oy, ty, p.in_h, iy)) continue;
        for (int tx = 0; tx < p.kw; ++tx) {
          int ix;
          if (!small_src(p.mode, p.stride, sh, p.pad, ox, tx, p.in_w, ix)) continue;
          tap_fma(p_in + ((size_t)(b * p.in_h + iy) * p.in_w + ix) * p.ld_in, wsm + (ty * p.kw + tx) * p.in_c * NC);
        }
      }
    }
    if (go != 0) {
      const TO* xr = p_xpre + (size_t)pix * p.ld_xpre;
#pragma unroll
      for (int n = 0; n < NC; ++n)
        if (n < p.out_c) acc[n] *= small_dact(lb_ld1(xr + n), go);      // no concat on this path (checked by the host)
    }
    TO* dst = p_out + (size_t)pix * p.ld_out;
    if (vec16_out) {                                   // whole 16-byte stores (8 bf16): half the store transactions of the 4-wide form
      constexpr int V = LbV<TO>::N;
#pragma unroll
      for (int j = 0; j < NC / V; ++j) {
        if (V * j < cols) {                            // cols % V == 0 on this path
          float o[V];
#pragma unroll
          for (int i = 0; i < V; ++i) o[i] = acc[V * j + i];
          lb_stv(dst + V * j, o);
        }
      }
    } else if (vec_out) {
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) {
        if (4 * j + 3 < cols) lb_st4(dst + 4 * j, make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
        else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (4 * j + i < cols) lb_st1(dst + 4 * j + i, acc[4 * j + i]);
        }
      }
    } else {
#pragma unroll
      for (int n = 0; n < NC; ++n) if (n < cols) lb_st1(dst + n, acc[n]);
    }
  }
}

// ---- narrow OUTPUT (out_c <= 4): input gradient of the stem layers, forward of a wide -> 3 layer --------------------
// wsm[tap][k] = alpha * W(tap, k, 0..3) as one float4
template <typename TI>
__global__ void __launch_bounds__(SMALL_THREADS) k_small_narrow_out(const SmallP p) {
  lb_pdl_enter();
  extern __shared__ float4 sm4[];
  const TI* p_in = reinterpret_cast<const TI*>(p.in);
  float* p_out = reinterpret_cast<float*>(p.out);
  const float* p_xpre = reinterpret_cast<const float*>(p.xpre);
  const int taps = p.kh * p.kw;
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
  for (int i = threadIdx.x; i < taps * p.in_c; i += blockDim.x) {
    const int k = i % p.in_c, tap = i / p.in_c;
    float v[4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
      v[n] = n < p.out_c ? alpha * __ldg(p.w + k * p.w_sk + n * p.w_sn + (tap / p.kw) * p.w_sty + (tap % p.kw) * p.w_stx) : 0.0f;
    sm4[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  const int sh = p.stride == 2 ? 1 : 0;
  const int gi = p.growth_in, go = p.growth_out;
  const bool vec_in = !(p.in_c & 3) && !(p.ld_in & 3) && lb_vec4_ok(p_in);
  float b4[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) b4[n] = (p.bias && n < p.out_c) ? __ldg(p.bias + n) : 0.0f;
  const int stride_t = gridDim.x * blockDim.x;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < p.pixels; pix += stride_t) {
    float a0 = b4[0], a1 = b4[1], a2 = b4[2], a3 = b4[3];
    auto tap_fma = [&](const TI* src, const float4* wt) {
      if (vec_in) {
        for (int k = 0; k < p.in_c; k += 4) {
          const float4 x4 = lb_ld4(src + k);
          const float xs[4] = {small_act(x4.x, gi), small_act(x4.y, gi), small_act(x4.z, gi), small_act(x4.w, gi)};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 wv = wt[k + i];
            a0 = fmaf(xs[i], wv.x, a0); a1 = fmaf(xs[i], wv.y, a1); a2 = fmaf(xs[i], wv.z, a2); a3 = fmaf(xs[i], wv.w, a3);
          }
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < p.in_c; ++k) {
          const float xv = small_act(lb_ld1(src + k), gi);
          const float4 wv = wt[k];
          a0 = fmaf(xv, wv.x, a0); a1 = fmaf(xv, wv.y, a1); a2 = fmaf(xv, wv.z, a2); a3 = fmaf(xv, wv.w, a3);
        }
      }
    };
    if (p.pointwise) {
      tap_fma(p_in + (size_t)pix * p.ld_in, sm4);
    } else {
      int t, ox, oy, b;
      lb_fast_divmod(p.d_w, pix, t, ox);
      lb_fast_divmod(p.d_h, t, b, oy);
      for (int ty = 0; ty < p.kh; ++ty) {
        int iy;
        if (!small_src(p.mode, p.stride, sh, p.pad, oy, ty, p.in_h, iy)) continue;
        for (int tx = 0; tx < p.kw; ++tx) {
          int ix;
          if (!small_src(p.mode, p.stride, sh, p.pad, ox, tx, p.in_w, ix)) continue;
          tap_fma(p_in + ((size_t)(b * p.in_h + iy) * p.in_w + ix) * p.ld_in, sm4 + (ty * p.kw + tx) * p.in_c);
        }
      }
    }
    float r[4] = {a0, a1, a2, a3};
    float* dst = p_out + (size_t)pix * p.ld_out;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      if (n < p.out_c) {
        if (go != 0) r[n] *= small_dact(__ldg(p_xpre + (size_t)pix * p.ld_xpre + n), go);
        dst[n] = r[n];
      }
    }
  }
}

static bool small_pointwise(const lb_conv_geom* g) { return g->kh == 1 && g->kw == 1 && g->stride == 1 && g->pad == 0; }

extern "C" int lb_conv_small_supported(const lb_conv_geom* g) {
  if (!g) return 0;
  if (g->stride != 1 && g->stride != 2) return 0;
  if (g->in_c < 1 || g->out_c < 1 || g->in_c > 64 || g->out_c > 64) return 0;
  if (g->in_c > 4 && g->out_c > 4) return 0;                   // one side must be tiny
  const long long pixels = (long long)g->batch * g->out_h * g->out_w;
  if (pixels >= (1ll << 31) - (1ll << 24)) return 0;
  const long long taps = (long long)g->kh * g->kw;
  // (a 1024 -> 1 head on a 1x1 map is a dot product per sample and stays a GEMM: in_c > 64)
  if (g->in_c <= 4) return taps * g->in_c * 64 + 64 <= 3 * SMALL_MAX_W ? 1 : 0;
  return taps * g->in_c * 4 <= 3 * SMALL_MAX_W ? 1 : 0;
}

// cat_input != 0: `out` points at the START of rows of in_c + out_c floats (row stride g->ld_out); the kernel writes the
// input copy and the conv result of a CatModule in one pass (1x1 stride-1 layers with in_c <= 4, no fused activations).
// wide_dtype: storage of the WIDE side (out and xpre when in_c <= 4, else in); the narrow side is fp32.  A layer with
// both sides narrow (3 -> 3) is all fp32: pass LB_F32.
extern "C" int lb_conv_small(const void* in, const float* w, const float* alpha, const float* bias, void* out,
                             const lb_conv_geom* g, int growth_in, const void* xpre, int ld_xpre, int growth_out,
                             int cat_input, int wide_dtype, lb_stream_t s) {
  LB_REQUIRE(in && w && out && g && growth_in >= 0 && growth_out >= -1 && (growth_out == 0 || xpre));
  LB_REQUIRE(wide_dtype == LB_F32 || wide_dtype == LB_BF16);
  if (!lb_conv_small_supported(g)) return LB_EUNSUPPORTED;
  SmallP p;
  p.in = in; p.w = w; p.alpha = alpha; p.bias = bias; p.xpre = xpre; p.out = out;
  p.batch = g->batch; p.in_h = g->in_h; p.in_w = g->in_w; p.in_c = g->in_c;
  p.out_h = g->out_h; p.out_w = g->out_w; p.out_c = g->out_c;
  p.kh = g->kh; p.kw = g->kw; p.stride = g->stride; p.pad = g->pad; p.mode = g->mode;
  p.ld_in = g->ld_in; p.ld_out = g->ld_out; p.ld_xpre = ld_xpre;
  p.w_sk = g->w_sk; p.w_sn = g->w_sn; p.w_sty = g->w_sty; p.w_stx = g->w_stx;
  p.growth_in = growth_in; p.growth_out = growth_out;
  p.pointwise = small_pointwise(g) ? 1 : 0;
  p.cat = cat_input ? 1 : 0;
  p.pixels = (int)((long long)g->batch * g->out_h * g->out_w);
  p.d_w = lb_make_fastdiv(g->out_w); p.d_h = lb_make_fastdiv(g->out_h);
  const int taps = g->kh * g->kw;
  const int grid = lb_grid_1d((size_t)p.pixels, SMALL_THREADS, 8);
  if (g->in_c <= 4) {
    if (p.cat) LB_REQUIRE(p.pointwise && growth_in == 0 && growth_out == 0 && g->ld_out >= g->in_c + g->out_c);
    const int cols = g->out_c + (p.cat ? g->in_c : 0);
    if (cols > 64) return LB_EUNSUPPORTED;
    // columns per thread: 4 for an RGB -> RGB layer (16 accumulators and 4 weight loads per input value for 3 live
    // columns made the 5x5 / stride-2 image convs instruction-bound at 0.3 TB/s), else whole 16-column groups
    const int nc = cols <= 4 ? 4 : (cols + 15) / 16 * 16;
    const size_t smem = ((size_t)taps * g->in_c * nc + nc) * sizeof(float);
    LB_DISPATCH(wide_dtype, T, {
      switch (nc) {
        case 4: lb_launch(k_small_narrow_in<4, T>, grid, SMALL_THREADS, smem, lb_s(s), p); break;
        case 16: lb_launch(k_small_narrow_in<16, T>, grid, SMALL_THREADS, smem, lb_s(s), p); break;
        case 32: lb_launch(k_small_narrow_in<32, T>, grid, SMALL_THREADS, smem, lb_s(s), p); break;
        case 48: lb_launch(k_small_narrow_in<48, T>, grid, SMALL_THREADS, smem, lb_s(s), p); break;
        default: lb_launch(k_small_narrow_in<64, T>, grid, SMALL_THREADS, smem, lb_s(s), p); break;
      }
    });
  } else {
    if (p.cat) return LB_EUNSUPPORTED;
    const size_t smem = (size_t)taps * g->in_c * sizeof(float4);
    LB_DISPATCH(wide_dtype, T, lb_launch(k_small_narrow_out<T>, grid, SMALL_THREADS, smem, lb_s(s), p));
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- weight gradient: tiny output, reduction over every pixel ------------------------------------------------
// One thread = one dense pixel, 16 x 4 accumulators in registers: "narrow" is the side with <= 4 channels, "wide" a
// 16-wide chunk of the other side (blockIdx.y).  acc[w][n] += wide[w] * narrow[n]; at the end a recursive-halving warp
// reduction, a shared-memory reduction across the CTA's warps and 64 global atomics per CTA.
//   kNarrowDense = false: 1x1 layers with g_c <= 4: narrow = gathered pixel (g_c), wide = dense channels [16*chunk, +16)
//   kNarrowDense = true : d_c <= 4: narrow = dense pixel, wide = im2col entries [16*chunk, +16) of (tap, kg) (tap-major)
struct SmallWgP {
  const void* gath; const void* dense; float* dw;      // the narrow operand is fp32, the wide one has storage TW
  int g_h, g_w, g_c, d_h, d_w, d_c, kh, kw, stride, pad, ld_g, ld_d;
  long long w_sk, w_sn, w_sty, w_stx;
  int pixels, rows, growth_g;
  int d_shift, d_vec;   // dense rows start d_shift elements before `dense` on a 4-element boundary; d_vec: 4-element loads usable
  LbFastDiv f_gc, f_kw, f_w, f_h;
};
template <bool kNarrowDense, typename TW>
__global__ void __launch_bounds__(SMALL_THREADS) k_small_wgrad(const SmallWgP p) {
  lb_pdl_enter();
  __shared__ float s_red[64];
  if (threadIdx.x < 64) s_red[threadIdx.x] = 0.0f;
  __syncthreads();
  const int w0 = blockIdx.y * 16;                    // first wide index of this chunk
  const int gg = p.growth_g;
  float acc[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[i][n] = 0.0f;
  // wide-entry decode for the im2col form is per chunk, not per pixel: entry w = (tap, kg), tap = (ty, tx)
  int e_ty[16], e_tx[16], e_kg[16];
  if (kNarrowDense) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int tap, kg, ty, tx;
      lb_fast_divmod(p.f_gc, min(w0 + i, p.rows - 1), tap, kg);
      lb_fast_divmod(p.f_kw, tap, ty, tx);
      e_ty[i] = ty; e_tx[i] = tx; e_kg[i] = kg;
    }
  }
  const int stride_t = gridDim.x * blockDim.x;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < p.pixels; pix += stride_t) {
    float nv[4], wv[16];
    if (!kNarrowDense) {
      // wide index = column of the 16-byte aligned dense row (a concat slice starts d_shift floats into it); columns
      // outside [d_shift, d_shift + d_c) are dropped when the sums are written
      const float* gs = reinterpret_cast<const float*>(p.gath) + (size_t)pix * p.ld_g;
      const TW* ds = reinterpret_cast<const TW*>(p.dense) - p.d_shift + (size_t)pix * p.ld_d + w0;
#pragma unroll
      for (int n = 0; n < 4; ++n) nv[n] = n < p.g_c ? small_act(__ldg(gs + n), gg) : 0.0f;
      if (p.d_vec) {                                   // a thread reads its own row: 16-byte loads touch each sector twice,
#pragma unroll                                         // 4-byte loads eight times (the L1 tag rate was the limiter)
        for (int j = 0; j < 4; ++j) {
          const float4 d4 = w0 + 4 * j < p.ld_d ? lb_ld4(ds + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          wv[4 * j] = d4.x; wv[4 * j + 1] = d4.y; wv[4 * j + 2] = d4.z; wv[4 * j + 3] = d4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) wv[i] = w0 + i < p.d_shift + p.d_c ? lb_ld1(ds + i) : 0.0f;
      }
    } else {
      const float* ds = reinterpret_cast<const float*>(p.dense) + (size_t)pix * p.ld_d;
#pragma unroll
      for (int n = 0; n < 4; ++n) nv[n] = n < p.d_c ? __ldg(ds + n) : 0.0f;
      int t, ox, oy, b;
      lb_fast_divmod(p.f_w, pix, t, ox);
      lb_fast_divmod(p.f_h, t, b, oy);
      const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int iy = iy0 + e_ty[i], ix = ix0 + e_tx[i];
        float v = 0.0f;
        if (w0 + i < p.rows && iy >= 0 && iy < p.g_h && ix >= 0 && ix < p.g_w)
          v = small_act(lb_ld1(reinterpret_cast<const TW*>(p.gath) + ((size_t)(b * p.g_h + iy) * p.g_w + ix) * p.ld_g + e_kg[i]), gg);   // act(0) = 0
        wv[i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[i][n] = fmaf(wv[i], nv[n], acc[i][n]);
  }
  // Warp reduction of the 64 per-thread sums by recursive halving: at step s a lane keeps half of its values and trades
  // the other half with lane ^ (16 >> s), so 32+16+8+4+2 = 62 shuffles (not 64 x 5) leave lane L with the warp totals of
  // entries 2L and 2L+1.
  const int lane = threadIdx.x & 31;
  float v[64];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int n = 0; n < 4; ++n) v[i * 4 + n] = acc[i][n];
#pragma unroll
  for (int half = 32, bit = 16; half >= 2; half >>= 1, bit >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  if (v[0] != 0.0f) atomicAdd(&s_red[2 * lane], v[0]);
  if (v[1] != 0.0f) atomicAdd(&s_red[2 * lane + 1], v[1]);
  __syncthreads();
  if (threadIdx.x < 64) {
    const int i = threadIdx.x >> 2, n = threadIdx.x & 3;
    const int wi = w0 + i;
    const float v = s_red[threadIdx.x];
    if (v != 0.0f) {
      if (!kNarrowDense) {                             // wide = dense row column, narrow = gathered channel kg (one tap)
        const int kd = wi - p.d_shift;
        if (kd >= 0 && kd < p.d_c && n < p.g_c) atomicAdd(p.dw + n * p.w_sk + kd * p.w_sn, v);
      } else if (wi < p.rows && n < p.d_c) {           // wide = (tap, kg), narrow = dense channel kd
        int tap, kg, ty, tx;
        lb_fast_divmod(p.f_gc, wi, tap, kg);
        lb_fast_divmod(p.f_kw, tap, ty, tx);
        atomicAdd(p.dw + kg * p.w_sk + n * p.w_sn + ty * p.w_sty + tx * p.w_stx, v);
      }
    }
  }
}

extern "C" int lb_conv_small_wgrad_supported(const lb_conv_geom* g) {
  if (!g || g->mode != 0) return 0;
  if (g->stride != 1 && g->stride != 2) return 0;
  const long long pixels = (long long)g->batch * g->out_h * g->out_w;
  if (pixels >= (1ll << 31) - (1ll << 24)) return 0;
  const bool pointwise = small_pointwise(g);
  if (pointwise && g->in_c <= 4 && g->out_c <= 1024) return 1;
  if (g->out_c <= 4 && (long long)g->kh * g->kw * g->in_c <= 1024) return 1;
  return 0;
}
// geom as lb_conv_wgrad: in_* = gathered operand, out_* = dense operand; dw in the master layout (+=, caller zeroes);
// growth_gathered > 0 applies RootTanh to the gathered operand on load (the layer's pre-activation)
// wide_dtype: storage of the wide operand (dense for 1x1 layers with in_c <= 4, else gathered); the narrow one is fp32
extern "C" int lb_conv_small_wgrad(const void* gathered, const void* dense, float* dw, const lb_conv_geom* g, int growth_gathered,
                                   int wide_dtype, lb_stream_t s) {
  LB_REQUIRE(gathered && dense && dw && g && growth_gathered >= 0);
  LB_REQUIRE(wide_dtype == LB_F32 || wide_dtype == LB_BF16);
  if (!lb_conv_small_wgrad_supported(g)) return LB_EUNSUPPORTED;
  SmallWgP p;
  p.gath = gathered; p.dense = dense; p.dw = dw;
  p.g_h = g->in_h; p.g_w = g->in_w; p.g_c = g->in_c; p.d_h = g->out_h; p.d_w = g->out_w; p.d_c = g->out_c;
  p.kh = g->kh; p.kw = g->kw; p.stride = g->stride; p.pad = g->pad; p.ld_g = g->ld_in; p.ld_d = g->ld_out;
  p.w_sk = g->w_sk; p.w_sn = g->w_sn; p.w_sty = g->w_sty; p.w_stx = g->w_stx;
  p.pixels = (int)((long long)g->batch * g->out_h * g->out_w);
  p.rows = g->kh * g->kw * g->in_c;
  p.growth_g = growth_gathered;
  p.f_gc = lb_make_fastdiv(g->in_c); p.f_kw = lb_make_fastdiv(g->kw);
  p.f_w = lb_make_fastdiv(g->out_w); p.f_h = lb_make_fastdiv(g->out_h);
  const bool narrow_gathered = small_pointwise(g) && g->in_c <= 4;
  p.d_shift = 0; p.d_vec = 0;
  if (narrow_gathered && !(g->ld_out & 3)) {           // dense rows are aligned for 4-element loads up to a fixed offset of the base
    const uintptr_t esz = wide_dtype == LB_BF16 ? 2 : 4;
    p.d_shift = (int)((reinterpret_cast<uintptr_t>(dense) & (4 * esz - 1)) / esz);
    p.d_vec = (reinterpret_cast<uintptr_t>(dense) & (esz - 1)) == 0 && p.d_shift + g->out_c <= g->ld_out ? 1 : 0;
    if (!p.d_vec) p.d_shift = 0;
  }
  const int chunks = narrow_gathered ? (p.d_shift + g->out_c + 15) / 16 : (p.rows + 15) / 16;
  int gx = lb_grid_1d((size_t)p.pixels, SMALL_THREADS, chunks >= 4 ? 1 : 2);   // many pixels per thread: the final reduction is a fixed cost
  if (chunks > 65535) return LB_EUNSUPPORTED;
  dim3 grid(gx, chunks);
  LB_DISPATCH(wide_dtype, T, {
    if (narrow_gathered) lb_launch(k_small_wgrad<false, T>, grid, SMALL_THREADS, 0, lb_s(s), p);
    else lb_launch(k_small_wgrad<true, T>, grid, SMALL_THREADS, 0, lb_s(s), p);
  });
  LB_LAUNCH_CHECK();
  return LB_OK;
}
