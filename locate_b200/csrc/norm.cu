// Whole-tensor normalisation ("InPlaceNorm", libs/inplace_norm.py:7-45): ONE mean and ONE unbiased
// std over all B*C*H*W elements, then per-channel (or per-sample-per-channel) gain and per-channel bias.
// HBM-bound.  Statistics are accumulated in double so N ~ 1e8 elements survive; the (sum, sumsq) pair
// and the two backward scalars are exposed so data parallel can all-reduce them between phases.
#include <cuda_bf16.h>
#include "common.cuh"

template <typename T>
__global__ void __launch_bounds__(256) k_norm_stats(const T* __restrict__ x, size_t n, double* __restrict__ sums,
                                                   double* __restrict__ work, int vec) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double s1 = 0.0, s2 = 0.0;
  if (vec) {
    constexpr int N = LbV<T>::N;
    const size_t nv = n / N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
      float v[N];
      lb_ldv(x + N * i, v);
      float p1 = 0.0f, p2 = 0.0f;
#pragma unroll
      for (int k = 0; k < N; ++k) { p1 += v[k]; p2 = fmaf(v[k], v[k], p2); }
      s1 += (double)p1;
      s2 += (double)p2;
    }
    for (size_t i = nv * N + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float v = lb_ld1(x + i);
      s1 += v; s2 += (double)v * v;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float v = lb_ld1(x + i);
      s1 += v; s2 += (double)v * v;
    }
  }
  s1 = lb_block_sum(s1, scratch);
  s2 = lb_block_sum(s2, scratch);
  lb_grid_sum2_ordered(s1, s2, work, sums, scratch);
}

extern "C" size_t lb_stat_work_doubles(void) { return LB_STAT_WORK_DOUBLES; }

extern "C" int lb_norm_stats(const void* x, size_t n, double* sums, double* work, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && sums && work && n > 0);
  LB_DISPATCH(dtype, T, lb_launch(k_norm_stats<T>, lb_grid_1d((n + 3) / 4, 256, 8), 256, 0, lb_s(s), lb_cp<T>(x), n, sums, work,
                                                                                          lb_vec_ok(lb_cp<T>(x)) ? 1 : 0));
  LB_LAUNCH_CHECK();
  return LB_OK;
}

__global__ void k_norm_finalize(const double* __restrict__ sums, double n, float* __restrict__ stats) {
  lb_pdl_enter();
  const double mean = sums[0] / n;
  double var = (sums[1] - sums[0] * mean) / (n - 1.0);      // unbiased, torch.std default
  if (var < 0.0) var = 0.0;
  const double sd = sqrt(var);
  stats[0] = (float)mean;
  stats[1] = (float)sd;
  stats[2] = (float)(1.0 / sd);                             // no eps in the reference: inf if sd == 0
  stats[3] = (float)n;
}
extern "C" int lb_norm_finalize(const double* sums, double n_total, float* stats, lb_stream_t s) {
  LB_REQUIRE(sums && stats && n_total > 1.0);
  lb_launch(k_norm_finalize, 1, 1, 0, lb_s(s), sums, n_total, stats);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// y = (x - mean) * gain[b?,c] * rstd + bias[c]; one thread = 4 consecutive channels of one pixel.  Optionally also (or
// only) act = RootTanh(y): the operand of the convolution that follows (conv.py:23-24), produced in the same pass.
template <typename T, bool kAct>
__global__ void __launch_bounds__(256) k_norm_apply4(const T* __restrict__ x, const float* __restrict__ stats,
                                                    const float* __restrict__ gain, int gain_bs, const float* __restrict__ bias,
                                                    T* __restrict__ y, T* __restrict__ act, T* __restrict__ dact, size_t nv,
                                                    LbFastDiv d_pcv, LbFastDiv d_cv) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    int b, rem, q, c;
    lb_fast_divmod(d_pcv, (int)i, b, rem);      // nv < 2^31 (checked by the host)
    lb_fast_divmod(d_cv, rem, q, c);
    c *= N;
    float v[N], gn[N], bs[N];
    lb_ldv(x + N * i, v);
    lb_ldf<N>(gain + (size_t)b * gain_bs + c, gn);
    lb_ldf<N>(bias + c, bs);
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = fmaf((v[k] - mean) * rstd, gn[k], bs[k]);
    if (y) lb_stv(y + N * i, v);
    if (kAct) {
      if (dact) {                               // RootTanh and its derivative from shared intermediates
#pragma unroll
        for (int k = 0; k < N; ++k) lb_roottanh_both_as<T>(v[k], v[k], gn[k]);
        lb_stv(dact + N * i, gn);
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = lb_roottanh_as<T>(v[k]);
      }
      lb_stv(act + N * i, v);
    }
  }
}
template <typename T>
__global__ void __launch_bounds__(256) k_norm_apply1(const T* __restrict__ x, const float* __restrict__ stats,
                                                    const float* __restrict__ gain, int gain_bs, const float* __restrict__ bias,
                                                    T* __restrict__ y, size_t n, int pc, int channels) {
  lb_pdl_enter();
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const size_t b = i / (size_t)pc;
    const int c = (int)(i % (size_t)channels);
    lb_st1(y + i, fmaf((lb_ld1(x + i) - mean) * rstd, __ldg(gain + b * gain_bs + c), __ldg(bias + c)));
  }
}

template <typename T>
static int norm_apply_t(const T* x, const float* stats, const float* gain, int gbs, const float* bias, T* y, int batch, int pixels,
                        int channels, lb_stream_t s) {
  const size_t n = (size_t)batch * pixels * channels;
  constexpr int N = LbV<T>::N;
  if ((channels % N) == 0 && n / N < ((size_t)1 << 31) - ((size_t)1 << 24) && lb_vec_ok(x) && lb_vec_ok(y) && lb_aligned16(gain) &&
      lb_aligned16(bias)) {
    lb_launch(k_norm_apply4<T, false>, lb_grid_1d(n / N, 256), 256, 0, lb_s(s), x, stats, gain, gbs, bias, y, nullptr, nullptr, n / N,
                                                                        lb_make_fastdiv((uint32_t)((size_t)pixels * channels / N)),
                                                                        lb_make_fastdiv(channels / N));
  } else {
    lb_launch(k_norm_apply1<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, stats, gain, gbs, bias, y, n, pixels * channels, channels);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_norm_apply(const void* x, const float* stats, const float* gain, int gain_batch_stride, const float* bias,
                             void* y, int batch, int pixels, int channels, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && stats && gain && bias && y && batch > 0 && pixels > 0 && channels > 0);
  LB_REQUIRE(gain_batch_stride == 0 || gain_batch_stride == channels);
  LB_DISPATCH(dtype, T, return norm_apply_t(lb_cp<T>(x), stats, gain, gain_batch_stride, bias, lb_p<T>(y), batch, pixels, channels, s));
}

template <typename T>
static int norm_apply_ex_t(const T* x, const float* stats, const float* gain, int gbs, const float* bias, T* y, T* act, T* dact,
                           int batch, int pixels, int channels, lb_stream_t s) {
  const size_t n = (size_t)batch * pixels * channels;
  constexpr int N = LbV<T>::N;
  if ((channels % N) || n / N >= ((size_t)1 << 31) - ((size_t)1 << 24) || !lb_vec_ok(x) || (y && !lb_vec_ok(y)) || !lb_vec_ok(act) ||
      (dact && !lb_vec_ok(dact)) || !lb_aligned16(gain) || !lb_aligned16(bias))
    return LB_EALIGN;
  lb_launch(k_norm_apply4<T, true>, lb_grid_1d(n / N, 256), 256, 0, lb_s(s), x, stats, gain, gbs, bias, y, act, dact, n / N,
                                                                     lb_make_fastdiv((uint32_t)((size_t)pixels * channels / N)),
                                                                     lb_make_fastdiv(channels / N));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
// y (may be NULL), act = RootTanh(y) and dact = RootTanh'(y) (may be NULL), all in storage `dtype`
extern "C" int lb_norm_apply_ex(const void* x, const float* stats, const float* gain, int gain_batch_stride, const float* bias,
                                void* y, void* act, void* dact, int batch, int pixels, int channels, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && stats && gain && bias && act && batch > 0 && pixels > 0 && channels > 0);
  LB_REQUIRE(gain_batch_stride == 0 || gain_batch_stride == channels);
  LB_DISPATCH(dtype, T, return norm_apply_ex_t(lb_cp<T>(x), stats, gain, gain_batch_stride, bias, lb_p<T>(y), lb_p<T>(act),
                                               lb_p<T>(dact), batch, pixels, channels, s));
}

// backward phase 1: per-(b,c) column sums over pixels.  CTA = (pixel chunk, b); thread = (channel lane, pixel lane)
template <typename T>
__global__ void k_norm_bwd_reduce(const T* __restrict__ x, const T* __restrict__ g, const float* __restrict__ stats,
                                  float* __restrict__ p1, float* __restrict__ p2, int pixels, int channels, int chunk, int tc, int tp) {
  lb_pdl_enter();
  if (threadIdx.x >= tc * tp) return;
  const float mean = __ldg(stats);
  const int cl = threadIdx.x % tc, pl = threadIdx.x / tc;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * chunk, q1 = min(pixels, q0 + chunk);
  const size_t base = (size_t)b * pixels * channels;
  for (int c = cl; c < channels; c += tc) {
    float a1 = 0.0f, a2 = 0.0f;
#pragma unroll 4
    for (int p = q0 + pl; p < q1; p += tp) {
      const size_t i = base + (size_t)p * channels + c;
      const float gv = lb_ld1(g + i);
      a1 += gv;
      a2 = fmaf(lb_ld1(x + i) - mean, gv, a2);
    }
    atomicAdd(p1 + (size_t)b * channels + c, a1);
    atomicAdd(p2 + (size_t)b * channels + c, a2);
  }
}
// vector form: 16-byte loads of x and g, one atomic per (b, c) and CTA
template <typename T>
__global__ void __launch_bounds__(256) k_norm_bwd_reduce_v(const T* __restrict__ x, const T* __restrict__ g, const float* __restrict__ stats,
                                                          float* __restrict__ p1, float* __restrict__ p2, int pixels, int channels,
                                                          int chunk, int cv, int tp) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  extern __shared__ float s_part[];
  const float mean = __ldg(stats);
  const int cl = threadIdx.x % cv, pl = threadIdx.x / cv;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * chunk, q1 = min(pixels, q0 + chunk);
  const size_t base = (size_t)b * pixels * channels + (size_t)cl * N;
  float a1[N], a2[N];
#pragma unroll
  for (int k = 0; k < N; ++k) a1[k] = a2[k] = 0.0f;
  const bool active = pl < tp;
  if (active) {
#pragma unroll 2
    for (int p = q0 + pl; p < q1; p += tp) {
      float xv[N], gv[N];
      lb_ldv(x + base + (size_t)p * channels, xv);
      lb_ldv(g + base + (size_t)p * channels, gv);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        a1[k] += gv[k];
        a2[k] = fmaf(xv[k] - mean, gv[k], a2[k]);
      }
    }
  }
  lb_colsum_flush<N>(a1, active, s_part, cl, pl, tp, channels, 1.0f, p1 + (size_t)b * channels);
  lb_colsum_flush<N>(a2, active, s_part, cl, pl, tp, channels, 1.0f, p2 + (size_t)b * channels);
}
extern "C" int lb_norm_bwd_reduce(const void* x, const void* g, const float* stats, float* p1, float* p2, int batch,
                                  int pixels, int channels, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && g && stats && p1 && p2 && batch > 0 && pixels > 0 && channels > 0);
  const LbColShape sh = lb_col_shape(channels);
  int chunks = (LB_SMS * 4 + batch - 1) / batch;
  int chunk = (pixels + chunks - 1) / chunks;
  if (chunk < sh.tp) chunk = sh.tp;
  chunks = (pixels + chunk - 1) / chunk;
  LB_DISPATCH(dtype, T, {
    constexpr int N = LbV<T>::N;
    if (!(channels % N) && channels / N <= 256 && lb_vec_ok(lb_cp<T>(x)) && lb_vec_ok(lb_cp<T>(g))) {
      const int cv = channels / N, tp = 256 / cv;
      chunks = (LB_SMS * 4 + batch - 1) / batch;
      chunk = (pixels + chunks - 1) / chunks;
      if (chunk < 4 * tp) chunk = 4 * tp;
      chunks = (pixels + chunk - 1) / chunk;
      lb_launch(k_norm_bwd_reduce_v<T>, dim3(chunks, batch), 256, (size_t)tp * channels * sizeof(float), lb_s(s), lb_cp<T>(x), lb_cp<T>(g),
                stats, p1, p2, pixels, channels, chunk, cv, tp);
    } else {
      lb_launch(k_norm_bwd_reduce<T>, dim3(chunks, batch), sh.threads, 0, lb_s(s), lb_cp<T>(x), lb_cp<T>(g), stats, p1, p2, pixels,
                channels, chunk, sh.tc, sh.tp);
    }
  });
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// backward phase 2 (small): scalars + parameter gradients.  CTA = 32 channels x 8 batch lanes, so the [B][C]
// partials are read coalesced and in parallel (a single CTA looping over the batch is latency-bound).
__global__ void __launch_bounds__(256) k_norm_bwd_finalize(const float* __restrict__ p1, const float* __restrict__ p2,
                                                          const float* __restrict__ gain, int gain_bs,
                                                          const float* __restrict__ stats, int batch, int channels, int bchunk,
                                                          float* __restrict__ dgain, float* __restrict__ dbias,
                                                          double* __restrict__ sout) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  __shared__ float s_db[8][33], s_dg[8][33];
  const float rstd = stats[2];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int b0 = blockIdx.y * bchunk, b1 = min(batch, b0 + bchunk);
  double s1 = 0.0, s2 = 0.0;
  float db = 0.0f, dg_shared = 0.0f;
  if (c < channels) {
    for (int b = b0 + bl; b < b1; b += 8) {
      const float a1 = p1[(size_t)b * channels + c], a2 = p2[(size_t)b * channels + c];
      const float gn = gain[(size_t)b * gain_bs + c];
      s1 += (double)gn * a1;
      s2 += (double)gn * a2;
      db += a1;
      if (gain_bs) dgain[(size_t)b * channels + c] += a2 * rstd; else dg_shared += a2;
    }
  }
  s_db[bl][cl] = db;
  s_dg[bl][cl] = dg_shared;
  __syncthreads();
  if (bl == 0 && c < channels) {
    float tb = 0.0f, tg = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { tb += s_db[i][cl]; tg += s_dg[i][cl]; }
    if (dbias) atomicAdd(dbias + c, tb);                     // batch chunks of one channel land on the same element
    if (!gain_bs && dgain) atomicAdd(dgain + c, tg * rstd);
  }
  s1 = lb_block_sum(s1, scratch);
  s2 = lb_block_sum(s2, scratch);
  if (threadIdx.x == 0) { atomicAdd(sout, s1); atomicAdd(sout + 1, s2); }
}
extern "C" int lb_norm_bwd_finalize(const float* p1, const float* p2, const float* gain, int gain_batch_stride,
                                    const float* stats, int batch, int channels, float* dgain, float* dbias, double* sout,
                                    lb_stream_t s) {
  LB_REQUIRE(p1 && p2 && gain && stats && sout && batch > 0 && channels > 0);
  LB_REQUIRE(gain_batch_stride == 0 || (gain_batch_stride == channels && dgain));
  cudaError_t e = cudaMemsetAsync(sout, 0, 2 * sizeof(double), lb_s(s));
  if (e != cudaSuccess) return (int)e;
  const int cblocks = (channels + 31) / 32;
  int bchunks = (LB_SMS + cblocks - 1) / cblocks;            // ~one wave of CTAs
  if (bchunks > (batch + 7) / 8) bchunks = (batch + 7) / 8;
  if (bchunks < 1) bchunks = 1;
  const int bchunk = (batch + bchunks - 1) / bchunks;
  bchunks = (batch + bchunk - 1) / bchunk;
  lb_launch(k_norm_bwd_finalize, dim3(cblocks, bchunks), 256, 0, lb_s(s), p1, p2, gain, gain_batch_stride, stats, batch, channels, bchunk,
                                                                   dgain, dbias, sout);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// backward phase 3: dx = gain*g*rstd - s0*rstd/N - s1*(x-mean)*rstd^3/(N-1)  (+ add, the gradient that reaches x through
// another branch -- the block's skip path -- summed here instead of in a pass of its own)
template <typename T>
__global__ void __launch_bounds__(256) k_norm_bwd_apply(const T* __restrict__ x, const T* __restrict__ g,
                                                       const float* __restrict__ stats, const float* __restrict__ gain,
                                                       int gain_bs, const double* __restrict__ sc, const T* __restrict__ add,
                                                       T* __restrict__ dx, size_t n, int pc, int channels) {
  lb_pdl_enter();
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const double nn = (double)__ldg(stats + 3);
  const double r = (double)rstd;
  const float k0 = (float)(sc[0] * r / nn);
  const float k1 = (float)(sc[1] * r * r * r / (nn - 1.0));
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const size_t b = i / (size_t)pc;
    const int c = (int)(i % (size_t)channels);
    const float gn = __ldg(gain + b * gain_bs + c);
    float o = fmaf(gn * rstd, lb_ld1(g + i), -k0) - k1 * (lb_ld1(x + i) - mean);
    if (add) o += lb_ld1(add + i);
    lb_st1(dx + i, o);
  }
}
// one thread = 4 consecutive channels; index decode by multiply-shift (a 64-bit divide per element made the scalar
// version instruction-bound at half the HBM rate)
template <typename T>
__global__ void __launch_bounds__(256) k_norm_bwd_apply4(const T* __restrict__ x, const T* __restrict__ g,
                                                        const float* __restrict__ stats, const float* __restrict__ gain,
                                                        int gain_bs, const double* __restrict__ sc, const T* __restrict__ add,
                                                        T* __restrict__ dx, int nv, LbFastDiv d_pcv, LbFastDiv d_cv) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const double nn = (double)__ldg(stats + 3);
  const double r = (double)rstd;
  const float k0 = (float)(sc[0] * r / nn);
  const float k1 = (float)(sc[1] * r * r * r / (nn - 1.0));
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    int b, rem, q, cv;
    lb_fast_divmod(d_pcv, i, b, rem);
    lb_fast_divmod(d_cv, rem, q, cv);
    float gn[N], gv[N], xv[N];
    lb_ldf<N>(gain + (size_t)b * gain_bs + N * cv, gn);
    lb_ldv(g + (size_t)N * i, gv);
    lb_ldv(x + (size_t)N * i, xv);
#pragma unroll
    for (int k = 0; k < N; ++k) gv[k] = fmaf(gn[k] * rstd, gv[k], -k0) - k1 * (xv[k] - mean);
    if (add) {
      lb_ldv(add + (size_t)N * i, xv);
#pragma unroll
      for (int k = 0; k < N; ++k) gv[k] += xv[k];
    }
    lb_stv(dx + (size_t)N * i, gv);
  }
}
template <typename T>
static int norm_bwd_apply_t(const T* x, const T* g, const float* stats, const float* gain, int gbs, const double* sc, const T* add, T* dx,
                            int batch, int pixels, int channels, lb_stream_t s) {
  const size_t n = (size_t)batch * pixels * channels;
  constexpr int N = LbV<T>::N;
  if (!(channels % N) && n / N < ((size_t)1 << 31) - ((size_t)1 << 24) && lb_vec_ok(x) && lb_vec_ok(g) && lb_vec_ok(dx) &&
      lb_aligned16(gain) && (!add || lb_vec_ok(add))) {
    lb_launch(k_norm_bwd_apply4<T>, lb_grid_1d(n / N, 256), 256, 0, lb_s(s), x, g, stats, gain, gbs, sc, add, dx, (int)(n / N),
                                                                  lb_make_fastdiv((uint32_t)((size_t)pixels * channels / N)),
                                                                  lb_make_fastdiv(channels / N));
  } else {
    lb_launch(k_norm_bwd_apply<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, g, stats, gain, gbs, sc, add, dx, n, pixels * channels, channels);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_norm_bwd_apply(const void* x, const void* g, const float* stats, const float* gain, int gain_batch_stride,
                                 const double* sc, const void* add, void* dx, int batch, int pixels, int channels, int dtype,
                                 lb_stream_t s) {
  LB_REQUIRE(x && g && stats && gain && sc && dx && batch > 0 && pixels > 0 && channels > 0);
  LB_DISPATCH(dtype, T, return norm_bwd_apply_t(lb_cp<T>(x), lb_cp<T>(g), stats, gain, gain_batch_stride, sc, lb_cp<T>(add), lb_p<T>(dx),
                                                batch, pixels, channels, s));
}
