// Softmax kernels (libs/attention.py:35,47).
//  * over PIXELS for every (b,c) of a channels-last [B][P][C] tensor: SelfAttention's softmax over HW (dim=-1 of
//    [B,F,HW]).  Rows of the logical problem are strided by C in memory, so a CTA owns (b, 32-channel slab, pixel
//    split): 8 lanes run over channel quads (4 consecutive channels per access) and 32 lanes over pixels.  Long rows at
//    small batch (HW = 65 536 at 256x256) are split over CTAs: per-split (max, sum exp) partials, then a second kernel
//    combines them and normalises -- otherwise B * C/32 CTAs would have to cover 148 SMs.
//  * over the contiguous last axis of [rows][cols]: feature attention's Softmax(dim=1) on [B,F,1,1].
// Storage type T (fp32 / bf16), fp32 arithmetic: the values are ~1/HW (6e-5 at 128^2); bf16 keeps 8 significant bits of
// them, and the gate that consumes them evaluates (gamma*att + 1) in fp32.
#include "common.cuh"

namespace {

constexpr int kCQ = 8;          // channel quads per CTA (32 channels)
constexpr int kPL = 32;         // pixel lanes per CTA
constexpr int kThreads = kCQ * kPL;

struct ML { float4 m, l; };

__device__ __forceinline__ void online1(float& m, float& l, float v) {
  const float nm = fmaxf(m, v);
  l = l * __expf(m - nm) + __expf(v - nm);
  m = nm;
}
__device__ __forceinline__ void online4(float& m, float& l, float a, float b, float c, float d) {
  const float nm = fmaxf(fmaxf(m, fmaxf(a, b)), fmaxf(c, d));
  l = l * __expf(m - nm) + ((__expf(a - nm) + __expf(b - nm)) + (__expf(c - nm) + __expf(d - nm)));
  m = nm;
}
__device__ __forceinline__ void merge1(float& m, float& l, float m2, float l2) {
  const float nm = fmaxf(m, m2);
  const float a = (m == -INFINITY) ? 0.0f : l * __expf(m - nm);
  const float b = (m2 == -INFINITY) ? 0.0f : l2 * __expf(m2 - nm);
  m = nm; l = a + b;
}

// (max, sum exp) of pixels [p0, p1) for the 4 channels of this thread's quad, reduced over the CTA's pixel lanes.
// Result valid in every thread (per channel quad).
template <typename T>
__device__ __forceinline__ ML slab_stats(const T* __restrict__ x, size_t base, int channels, int p0, int p1, int cq, int pl, bool ok,
                                         float4 (*s_m)[kCQ], float4 (*s_l)[kCQ]) {
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), l = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    // 4 pixels per step: four independent loads in flight, ONE rescale of the running sum per step (the per-pixel online
    // update is a loop-carried chain of two exps: 128 dependent steps per thread at 64x64 left the pass latency-bound)
    int p = p0 + pl;
    for (; p + 3 * kPL < p1; p += 4 * kPL) {
      const float4 v0 = lb_ld4(x + base + (size_t)p * channels), v1 = lb_ld4(x + base + (size_t)(p + kPL) * channels);
      const float4 v2 = lb_ld4(x + base + (size_t)(p + 2 * kPL) * channels), v3 = lb_ld4(x + base + (size_t)(p + 3 * kPL) * channels);
      online4(m.x, l.x, v0.x, v1.x, v2.x, v3.x); online4(m.y, l.y, v0.y, v1.y, v2.y, v3.y);
      online4(m.z, l.z, v0.z, v1.z, v2.z, v3.z); online4(m.w, l.w, v0.w, v1.w, v2.w, v3.w);
    }
    for (; p < p1; p += kPL) {
      const float4 v = lb_ld4(x + base + (size_t)p * channels);
      online1(m.x, l.x, v.x); online1(m.y, l.y, v.y); online1(m.z, l.z, v.z); online1(m.w, l.w, v.w);
    }
  }
  s_m[pl][cq] = m;
  s_l[pl][cq] = l;
  __syncthreads();
  ML r;
  r.m = s_m[0][cq]; r.l = s_l[0][cq];
#pragma unroll 4
  for (int i = 1; i < kPL; ++i) {
    const float4 m2 = s_m[i][cq], l2 = s_l[i][cq];
    merge1(r.m.x, r.l.x, m2.x, l2.x); merge1(r.m.y, r.l.y, m2.y, l2.y);
    merge1(r.m.z, r.l.z, m2.z, l2.z); merge1(r.m.w, r.l.w, m2.w, l2.w);
  }
  return r;
}

// splits == 1: whole softmax in one CTA per (slab, b).  splits > 1, phase 0: write partials[b][split][c] = (m, l);
// phase 1: combine the partials of all splits, normalise this CTA's pixel range.
template <typename T>
__global__ void __launch_bounds__(kThreads) k_softmax_pixels_fwd(const T* __restrict__ x, T* __restrict__ y, int pixels, int channels,
                                                                 int splits, int chunk, int phase, float2* __restrict__ part) {
  lb_pdl_enter();
  __shared__ float4 s_m[kPL][kCQ], s_l[kPL][kCQ];
  const int cq = threadIdx.x & (kCQ - 1), pl = threadIdx.x >> 3;
  const int c = blockIdx.x * (4 * kCQ) + 4 * cq;
  const bool ok = c < channels;
  const int b = blockIdx.y, split = blockIdx.z;
  const size_t base = (size_t)b * pixels * channels + (ok ? c : 0);
  const int p0 = split * chunk, p1 = min(pixels, p0 + chunk);
  ML r;
  if (splits == 1 || phase == 0) {
    r = slab_stats(x, base, channels, p0, p1, cq, pl, ok, s_m, s_l);
    if (splits > 1) {
      if (ok && pl == 0) {
        float2* dst = part + ((size_t)b * splits + split) * channels + c;
        dst[0] = make_float2(r.m.x, r.l.x); dst[1] = make_float2(r.m.y, r.l.y);
        dst[2] = make_float2(r.m.z, r.l.z); dst[3] = make_float2(r.m.w, r.l.w);
      }
      return;
    }
  } else {
    r.m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); r.l = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      for (int sp = 0; sp < splits; ++sp) {                 // fixed order: reproducible
        const float2* src = part + ((size_t)b * splits + sp) * channels + c;
        const float2 a0 = src[0], a1 = src[1], a2 = src[2], a3 = src[3];
        merge1(r.m.x, r.l.x, a0.x, a0.y); merge1(r.m.y, r.l.y, a1.x, a1.y);
        merge1(r.m.z, r.l.z, a2.x, a2.y); merge1(r.m.w, r.l.w, a3.x, a3.y);
      }
    }
  }
  if (!ok) return;
  const float4 inv = make_float4(1.0f / r.l.x, 1.0f / r.l.y, 1.0f / r.l.z, 1.0f / r.l.w);
#pragma unroll 4
  for (int p = p0 + pl; p < p1; p += kPL) {
    const size_t i = base + (size_t)p * channels;
    const float4 v = lb_ld4(x + i);
    lb_st4(y + i, make_float4(__expf(v.x - r.m.x) * inv.x, __expf(v.y - r.m.y) * inv.y, __expf(v.z - r.m.z) * inv.z,
                              __expf(v.w - r.m.w) * inv.w));
  }
}

// dx = y * (g - sum_p y*g); same CTA shape.  splits > 1: phase 0 writes partial dots part[b][split][c], phase 1 combines.
template <typename T>
__global__ void __launch_bounds__(kThreads) k_softmax_pixels_bwd(const T* __restrict__ y, const T* __restrict__ g, T* __restrict__ dx,
                                                                 int pixels, int channels, int splits, int chunk, int phase,
                                                                 float* __restrict__ part) {
  lb_pdl_enter();
  __shared__ float4 s_d[kPL][kCQ];
  const int cq = threadIdx.x & (kCQ - 1), pl = threadIdx.x >> 3;
  const int c = blockIdx.x * (4 * kCQ) + 4 * cq;
  const bool ok = c < channels;
  const int b = blockIdx.y, split = blockIdx.z;
  const size_t base = (size_t)b * pixels * channels + (ok ? c : 0);
  const int p0 = split * chunk, p1 = min(pixels, p0 + chunk);
  float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
  if (splits == 1 || phase == 0) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      for (int p = p0 + pl; p < p1; p += kPL) {
        const size_t i = base + (size_t)p * channels;
        const float4 yv = lb_ld4(y + i), gv = lb_ld4(g + i);
        d.x = fmaf(yv.x, gv.x, d.x); d.y = fmaf(yv.y, gv.y, d.y); d.z = fmaf(yv.z, gv.z, d.z); d.w = fmaf(yv.w, gv.w, d.w);
      }
    }
    s_d[pl][cq] = d;
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < kPL; ++i) {
      const float4 e = s_d[i][cq];
      tot.x += e.x; tot.y += e.y; tot.z += e.z; tot.w += e.w;
    }
    if (splits > 1) {
      if (ok && pl == 0) *reinterpret_cast<float4*>(part + ((size_t)b * splits + split) * channels + c) = tot;
      return;
    }
  } else if (ok) {
    for (int sp = 0; sp < splits; ++sp) {
      const float4 e = *reinterpret_cast<const float4*>(part + ((size_t)b * splits + sp) * channels + c);
      tot.x += e.x; tot.y += e.y; tot.z += e.z; tot.w += e.w;
    }
  }
  if (!ok) return;
  for (int p = p0 + pl; p < p1; p += kPL) {
    const size_t i = base + (size_t)p * channels;
    const float4 yv = lb_ld4(y + i), gv = lb_ld4(g + i);
    lb_st4(dx + i, make_float4(yv.x * (gv.x - tot.x), yv.y * (gv.y - tot.y), yv.z * (gv.z - tot.z), yv.w * (gv.w - tot.w)));
  }
}

// scalar fallbacks (channels % 4 != 0 or unaligned views): CTA = (b, 32-channel slab), lanes over channels
constexpr int kWarpsS = 16;
template <typename T>
__global__ void __launch_bounds__(32 * kWarpsS) k_softmax_pixels_fwd_s(const T* __restrict__ x, T* __restrict__ y, int pixels, int channels) {
  lb_pdl_enter();
  __shared__ float s_max[kWarpsS][33];
  __shared__ float s_sum[kWarpsS][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < channels;
  const size_t base = (size_t)blockIdx.y * pixels * channels + (ok ? c : 0);
  float m = -INFINITY, l = 0.0f;
  if (ok)
    for (int p = w; p < pixels; p += kWarpsS) online1(m, l, lb_ld1(x + base + (size_t)p * channels));
  s_max[w][lane] = m;
  s_sum[w][lane] = l;
  __syncthreads();
  float gm = -INFINITY, gl = 0.0f;
#pragma unroll
  for (int i = 0; i < kWarpsS; ++i) merge1(gm, gl, s_max[i][lane], s_sum[i][lane]);
  const float inv = 1.0f / gl;
  if (ok)
    for (int p = w; p < pixels; p += kWarpsS) {
      const size_t i = base + (size_t)p * channels;
      lb_st1(y + i, __expf(lb_ld1(x + i) - gm) * inv);
    }
}
template <typename T>
__global__ void __launch_bounds__(32 * kWarpsS) k_softmax_pixels_bwd_s(const T* __restrict__ y, const T* __restrict__ g, T* __restrict__ dx,
                                                                       int pixels, int channels) {
  lb_pdl_enter();
  __shared__ float s_dot[kWarpsS][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < channels;
  const size_t base = (size_t)blockIdx.y * pixels * channels + (ok ? c : 0);
  float d = 0.0f;
  if (ok)
    for (int p = w; p < pixels; p += kWarpsS) {
      const size_t i = base + (size_t)p * channels;
      d = fmaf(lb_ld1(y + i), lb_ld1(g + i), d);
    }
  s_dot[w][lane] = d;
  __syncthreads();
  float tot = 0.0f;
#pragma unroll
  for (int i = 0; i < kWarpsS; ++i) tot += s_dot[i][lane];
  if (ok)
    for (int p = w; p < pixels; p += kWarpsS) {
      const size_t i = base + (size_t)p * channels;
      lb_st1(dx + i, lb_ld1(y + i) * (lb_ld1(g + i) - tot));
    }
}

// pixel splits so that ~2 waves of CTAs exist; 1 when the batch alone fills the GPU
int pixel_splits(int batch, int pixels, int channels, int* chunk) {
  const long long ctas = (long long)batch * ((channels + 4 * kCQ - 1) / (4 * kCQ));
  long long splits = 1;
  if (ctas < 2 * LB_SMS) splits = (2 * LB_SMS + ctas - 1) / ctas;
  const long long max_splits = (pixels + 4 * kPL - 1) / (4 * kPL);          // at least 4 pixels per lane
  if (splits > max_splits) splits = max_splits;
  if (splits > 256) splits = 256;
  if (splits < 1) splits = 1;
  *chunk = (int)((pixels + splits - 1) / splits);
  return (int)((pixels + *chunk - 1) / *chunk);
}

template <typename T>
int softmax_pixels_fwd_t(const T* x, T* y, int batch, int pixels, int channels, float* work, size_t work_floats, lb_stream_t s) {
  if ((channels & 3) || !lb_vec4_ok(x) || !lb_vec4_ok(y)) {
    lb_launch(k_softmax_pixels_fwd_s<T>, dim3((channels + 31) / 32, batch), 32 * kWarpsS, 0, lb_s(s), x, y, pixels, channels);
    LB_LAUNCH_CHECK();
    return LB_OK;
  }
  int chunk;
  int splits = pixel_splits(batch, pixels, channels, &chunk);
  if (splits > 1 && (!work || work_floats < (size_t)2 * batch * splits * channels)) { splits = 1; chunk = pixels; }
  const dim3 grid((channels + 4 * kCQ - 1) / (4 * kCQ), batch, splits);
  LB_REQUIRE(grid.z <= 65535);
  float2* part = reinterpret_cast<float2*>(work);
  lb_launch(k_softmax_pixels_fwd<T>, grid, kThreads, 0, lb_s(s), x, y, pixels, channels, splits, chunk, 0, part);
  LB_LAUNCH_CHECK();
  if (splits > 1) {
    lb_launch(k_softmax_pixels_fwd<T>, grid, kThreads, 0, lb_s(s), x, y, pixels, channels, splits, chunk, 1, part);
    LB_LAUNCH_CHECK();
  }
  return LB_OK;
}
template <typename T>
int softmax_pixels_bwd_t(const T* y, const T* g, T* dx, int batch, int pixels, int channels, float* work, size_t work_floats,
                         lb_stream_t s) {
  if ((channels & 3) || !lb_vec4_ok(y) || !lb_vec4_ok(g) || !lb_vec4_ok(dx)) {
    lb_launch(k_softmax_pixels_bwd_s<T>, dim3((channels + 31) / 32, batch), 32 * kWarpsS, 0, lb_s(s), y, g, dx, pixels, channels);
    LB_LAUNCH_CHECK();
    return LB_OK;
  }
  int chunk;
  int splits = pixel_splits(batch, pixels, channels, &chunk);
  if (splits > 1 && (!work || work_floats < (size_t)batch * splits * channels || (reinterpret_cast<uintptr_t>(work) & 15))) {
    splits = 1; chunk = pixels;
  }
  const dim3 grid((channels + 4 * kCQ - 1) / (4 * kCQ), batch, splits);
  LB_REQUIRE(grid.z <= 65535);
  lb_launch(k_softmax_pixels_bwd<T>, grid, kThreads, 0, lb_s(s), y, g, dx, pixels, channels, splits, chunk, 0, work);
  LB_LAUNCH_CHECK();
  if (splits > 1) {
    lb_launch(k_softmax_pixels_bwd<T>, grid, kThreads, 0, lb_s(s), y, g, dx, pixels, channels, splits, chunk, 1, work);
    LB_LAUNCH_CHECK();
  }
  return LB_OK;
}

}  // namespace

// floats of scratch for the split-row path (0: the batch alone fills the GPU); forward needs 2x this for (max, sum) pairs
extern "C" size_t lb_softmax_pixels_work_floats(int batch, int pixels, int channels) {
  if (batch <= 0 || pixels <= 0 || channels <= 0 || (channels & 3)) return 0;
  int chunk;
  const int splits = pixel_splits(batch, pixels, channels, &chunk);
  return splits > 1 ? (size_t)2 * batch * splits * channels : 0;
}
extern "C" int lb_softmax_pixels_fwd(const void* x, void* y, int batch, int pixels, int channels, float* work, size_t work_floats,
                                     int dtype, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && pixels > 0 && channels > 0 && batch <= 65535);
  LB_DISPATCH(dtype, T, return softmax_pixels_fwd_t(lb_cp<T>(x), lb_p<T>(y), batch, pixels, channels, work, work_floats, s));
}
extern "C" int lb_softmax_pixels_bwd(const void* y, const void* g, void* dx, int batch, int pixels, int channels, float* work,
                                     size_t work_floats, int dtype, lb_stream_t s) {
  LB_REQUIRE(y && g && dx && batch > 0 && pixels > 0 && channels > 0 && batch <= 65535);
  LB_DISPATCH(dtype, T, return softmax_pixels_bwd_t(lb_cp<T>(y), lb_cp<T>(g), lb_p<T>(dx), batch, pixels, channels, work, work_floats, s));
}

// contiguous rows: one warp per row
template <typename T>
__global__ void __launch_bounds__(128) k_softmax_rows_fwd(const T* __restrict__ x, T* __restrict__ y, int rows, int cols) {
  lb_pdl_enter();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + (size_t)row * cols;
  float m = -INFINITY;
  for (int j = lane; j < cols; j += 32) m = fmaxf(m, lb_ld1(xr + j));
  m = lb_warp_max(m);
  float l = 0.0f;
  for (int j = lane; j < cols; j += 32) l += __expf(lb_ld1(xr + j) - m);
  l = lb_warp_sum(l);
  const float inv = 1.0f / l;
  for (int j = lane; j < cols; j += 32) lb_st1(y + (size_t)row * cols + j, __expf(lb_ld1(xr + j) - m) * inv);
}
template <typename T>
__global__ void __launch_bounds__(128) k_softmax_rows_bwd(const T* __restrict__ y, const T* __restrict__ g, T* __restrict__ dx, int rows,
                                                         int cols) {
  lb_pdl_enter();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t o = (size_t)row * cols;
  float d = 0.0f;
  for (int j = lane; j < cols; j += 32) d = fmaf(lb_ld1(y + o + j), lb_ld1(g + o + j), d);
  d = lb_warp_sum(d);
  for (int j = lane; j < cols; j += 32) lb_st1(dx + o + j, lb_ld1(y + o + j) * (lb_ld1(g + o + j) - d));
}
extern "C" int lb_softmax_rows_fwd(const void* x, void* y, int rows, int cols, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && y && rows > 0 && cols > 0);
  LB_DISPATCH(dtype, T, lb_launch(k_softmax_rows_fwd<T>, (rows + 3) / 4, 128, 0, lb_s(s), lb_cp<T>(x), lb_p<T>(y), rows, cols));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_softmax_rows_bwd(const void* y, const void* g, void* dx, int rows, int cols, int dtype, lb_stream_t s) {
  LB_REQUIRE(y && g && dx && rows > 0 && cols > 0);
  LB_DISPATCH(dtype, T, lb_launch(k_softmax_rows_bwd<T>, (rows + 3) / 4, 128, 0, lb_s(s), lb_cp<T>(y), lb_cp<T>(g), lb_p<T>(dx), rows, cols));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
