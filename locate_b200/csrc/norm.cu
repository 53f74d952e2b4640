// Whole-tensor normalisation ("InPlaceNorm", libs/inplace_norm.py:7-45): ONE mean and ONE unbiased
// std over all B*C*H*W elements, then per-channel (or per-sample-per-channel) gain and per-channel bias.
// HBM-bound.  Statistics are accumulated in double so N ~ 1e8 elements survive; the (sum, sumsq) pair
// and the two backward scalars are exposed so data parallel can all-reduce them between phases.
#include <cuda_bf16.h>
#include "common.cuh"

__global__ void __launch_bounds__(256) k_norm_stats(const float* __restrict__ x, size_t n, double* __restrict__ sums,
                                                   double* __restrict__ work, int vec) {
  __shared__ double scratch[32];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double s1 = 0.0, s2 = 0.0;
  if (vec) {
    const size_t n4 = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = lb_ld4(x + 4 * i);
      s1 += (double)((v.x + v.y) + (v.z + v.w));
      s2 += (double)(fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w));
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float v = x[i];
      s1 += v; s2 += (double)v * v;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float v = x[i];
      s1 += v; s2 += (double)v * v;
    }
  }
  s1 = lb_block_sum(s1, scratch);
  s2 = lb_block_sum(s2, scratch);
  lb_grid_sum2_ordered(s1, s2, work, sums, scratch);
}

extern "C" size_t lb_stat_work_doubles(void) { return LB_STAT_WORK_DOUBLES; }

extern "C" int lb_norm_stats(const float* x, size_t n, double* sums, double* work, lb_stream_t s) {
  LB_REQUIRE(x && sums && work && n > 0);
  k_norm_stats<<<lb_grid_1d((n + 3) / 4, 256, 4), 256, 0, lb_s(s)>>>(x, n, sums, work, lb_aligned16(x) ? 1 : 0);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

__global__ void k_norm_finalize(const double* __restrict__ sums, double n, float* __restrict__ stats) {
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
  k_norm_finalize<<<1, 1, 0, lb_s(s)>>>(sums, n_total, stats);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// y = (x - mean) * gain[b?,c] * rstd + bias[c]; one thread = 4 consecutive channels of one pixel
__global__ void __launch_bounds__(256) k_norm_apply4(const float* __restrict__ x, const float* __restrict__ stats,
                                                    const float* __restrict__ gain, int gain_bs, const float* __restrict__ bias,
                                                    float* __restrict__ y, size_t n4, LbFastDiv d_pc4, LbFastDiv d_c4) {
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    int b, rem, q, c;
    lb_fast_divmod(d_pc4, (int)i, b, rem);      // n4 < 2^31 (checked by the host)
    lb_fast_divmod(d_c4, rem, q, c);
    c *= 4;
    const float4 v = lb_ld4(x + 4 * i);
    const float4 gn = lb_ld4(gain + (size_t)b * gain_bs + c);
    const float4 bs = lb_ld4(bias + c);
    float4 r;
    r.x = fmaf((v.x - mean) * rstd, gn.x, bs.x);
    r.y = fmaf((v.y - mean) * rstd, gn.y, bs.y);
    r.z = fmaf((v.z - mean) * rstd, gn.z, bs.z);
    r.w = fmaf((v.w - mean) * rstd, gn.w, bs.w);
    lb_st4(y + 4 * i, r);
  }
}
// same, also (or only) emitting the bf16 GEMM operand of the consumer: y16 = bf16(RootTanh?(y)).  y may be NULL when
// nothing reads the fp32 result (a conv follows directly and its backward needs only the bf16 operand).
template <bool kAct>
__global__ void __launch_bounds__(256) k_norm_apply4_ex(const float* __restrict__ x, const float* __restrict__ stats,
                                                       const float* __restrict__ gain, int gain_bs, const float* __restrict__ bias,
                                                       float* __restrict__ y, __nv_bfloat16* __restrict__ y16, size_t n4, LbFastDiv d_pc4,
                                                       LbFastDiv d_c4) {
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    int b, rem, q, c;
    lb_fast_divmod(d_pc4, (int)i, b, rem);      // n4 < 2^31 (checked by the host)
    lb_fast_divmod(d_c4, rem, q, c);
    c *= 4;
    const float4 v = lb_ld4(x + 4 * i);
    const float4 gn = lb_ld4(gain + (size_t)b * gain_bs + c);
    const float4 bs = lb_ld4(bias + c);
    float4 r;
    r.x = fmaf((v.x - mean) * rstd, gn.x, bs.x);
    r.y = fmaf((v.y - mean) * rstd, gn.y, bs.y);
    r.z = fmaf((v.z - mean) * rstd, gn.z, bs.z);
    r.w = fmaf((v.w - mean) * rstd, gn.w, bs.w);
    if (y) lb_st4(y + 4 * i, r);
    if (kAct) { r.x = lb_roottanh(r.x); r.y = lb_roottanh(r.y); r.z = lb_roottanh(r.z); r.w = lb_roottanh(r.w); }
    __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y16 + 4 * i) = pk;
  }
}
__global__ void __launch_bounds__(256) k_norm_apply1(const float* __restrict__ x, const float* __restrict__ stats,
                                                    const float* __restrict__ gain, int gain_bs, const float* __restrict__ bias,
                                                    float* __restrict__ y, size_t n, int pc, int channels) {
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const size_t b = i / (size_t)pc;
    const int c = (int)(i % (size_t)channels);
    y[i] = fmaf((x[i] - mean) * rstd, __ldg(gain + b * gain_bs + c), __ldg(bias + c));
  }
}

extern "C" int lb_norm_apply(const float* x, const float* stats, const float* gain, int gain_batch_stride, const float* bias,
                             float* y, int batch, int pixels, int channels, lb_stream_t s) {
  LB_REQUIRE(x && stats && gain && bias && y && batch > 0 && pixels > 0 && channels > 0);
  LB_REQUIRE(gain_batch_stride == 0 || gain_batch_stride == channels);
  const size_t n = (size_t)batch * pixels * channels;
  if ((channels & 3) == 0 && n / 4 < ((size_t)1 << 31) - ((size_t)1 << 24) && lb_aligned16(x) && lb_aligned16(y) && lb_aligned16(gain) &&
      lb_aligned16(bias)) {
    k_norm_apply4<<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(x, stats, gain, gain_batch_stride, bias, y, n / 4,
                                                              lb_make_fastdiv((uint32_t)((size_t)pixels * channels / 4)),
                                                              lb_make_fastdiv(channels / 4));
  } else {
    k_norm_apply1<<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(x, stats, gain, gain_batch_stride, bias, y, n, pixels * channels, channels);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}

extern "C" int lb_norm_apply_ex(const float* x, const float* stats, const float* gain, int gain_batch_stride, const float* bias,
                                float* y, void* y16, int act16, int batch, int pixels, int channels, lb_stream_t s) {
  LB_REQUIRE(x && stats && gain && bias && y16 && batch > 0 && pixels > 0 && channels > 0);
  LB_REQUIRE(gain_batch_stride == 0 || gain_batch_stride == channels);
  const size_t n = (size_t)batch * pixels * channels;
  if ((channels & 3) || n / 4 >= ((size_t)1 << 31) - ((size_t)1 << 24) || !lb_aligned16(x) || (y && !lb_aligned16(y)) ||
      !lb_aligned16(gain) || !lb_aligned16(bias) || (reinterpret_cast<uintptr_t>(y16) & 7))
    return LB_EALIGN;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(y16);
  if (act16)
    k_norm_apply4_ex<true><<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(x, stats, gain, gain_batch_stride, bias, y, o16, n / 4,
                                                                       lb_make_fastdiv((uint32_t)((size_t)pixels * channels / 4)),
                                                                       lb_make_fastdiv(channels / 4));
  else
    k_norm_apply4_ex<false><<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(x, stats, gain, gain_batch_stride, bias, y, o16, n / 4,
                                                                        lb_make_fastdiv((uint32_t)((size_t)pixels * channels / 4)),
                                                                        lb_make_fastdiv(channels / 4));
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// backward phase 1: per-(b,c) column sums over pixels.  CTA = (pixel chunk, b); thread = (channel lane, pixel lane)
__global__ void k_norm_bwd_reduce(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ stats,
                                  float* __restrict__ p1, float* __restrict__ p2, int pixels, int channels, int chunk, int tc, int tp) {
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
      const float gv = g[i];
      a1 += gv;
      a2 = fmaf(x[i] - mean, gv, a2);
    }
    atomicAdd(p1 + (size_t)b * channels + c, a1);
    atomicAdd(p2 + (size_t)b * channels + c, a2);
  }
}
extern "C" int lb_norm_bwd_reduce(const float* x, const float* g, const float* stats, float* p1, float* p2, int batch,
                                  int pixels, int channels, lb_stream_t s) {
  LB_REQUIRE(x && g && stats && p1 && p2 && batch > 0 && pixels > 0 && channels > 0);
  const LbColShape sh = lb_col_shape(channels);
  int chunks = (LB_SMS * 4 + batch - 1) / batch;
  int chunk = (pixels + chunks - 1) / chunks;
  if (chunk < sh.tp) chunk = sh.tp;
  chunks = (pixels + chunk - 1) / chunk;
  k_norm_bwd_reduce<<<dim3(chunks, batch), sh.threads, 0, lb_s(s)>>>(x, g, stats, p1, p2, pixels, channels, chunk, sh.tc, sh.tp);
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
  k_norm_bwd_finalize<<<dim3(cblocks, bchunks), 256, 0, lb_s(s)>>>(p1, p2, gain, gain_batch_stride, stats, batch, channels, bchunk,
                                                                   dgain, dbias, sout);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// backward phase 3: dx = gain*g*rstd - s0*rstd/N - s1*(x-mean)*rstd^3/(N-1)
__global__ void __launch_bounds__(256) k_norm_bwd_apply(const float* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ stats, const float* __restrict__ gain,
                                                       int gain_bs, const double* __restrict__ sc, float* __restrict__ dx,
                                                       size_t n, int pc, int channels) {
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
    dx[i] = fmaf(gn * rstd, g[i], -k0) - k1 * (x[i] - mean);
  }
}
// one thread = 4 consecutive channels; index decode by multiply-shift (a 64-bit divide per element made the scalar
// version instruction-bound at half the HBM rate)
__global__ void __launch_bounds__(256) k_norm_bwd_apply4(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ stats, const float* __restrict__ gain,
                                                        int gain_bs, const double* __restrict__ sc, float* __restrict__ dx,
                                                        int n4, LbFastDiv d_pc4, LbFastDiv d_c4) {
  const float mean = __ldg(stats), rstd = __ldg(stats + 2);
  const double nn = (double)__ldg(stats + 3);
  const double r = (double)rstd;
  const float k0 = (float)(sc[0] * r / nn);
  const float k1 = (float)(sc[1] * r * r * r / (nn - 1.0));
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    int b, rem, q, c4;
    lb_fast_divmod(d_pc4, i, b, rem);
    lb_fast_divmod(d_c4, rem, q, c4);
    const float4 gn = lb_ld4(gain + (size_t)b * gain_bs + 4 * c4);
    const float4 gv = lb_ld4(g + 4 * (size_t)i), xv = lb_ld4(x + 4 * (size_t)i);
    float4 o;
    o.x = fmaf(gn.x * rstd, gv.x, -k0) - k1 * (xv.x - mean);
    o.y = fmaf(gn.y * rstd, gv.y, -k0) - k1 * (xv.y - mean);
    o.z = fmaf(gn.z * rstd, gv.z, -k0) - k1 * (xv.z - mean);
    o.w = fmaf(gn.w * rstd, gv.w, -k0) - k1 * (xv.w - mean);
    lb_st4(dx + 4 * (size_t)i, o);
  }
}
extern "C" int lb_norm_bwd_apply(const float* x, const float* g, const float* stats, const float* gain, int gain_batch_stride,
                                 const double* sc, float* dx, int batch, int pixels, int channels, lb_stream_t s) {
  LB_REQUIRE(x && g && stats && gain && sc && dx && batch > 0 && pixels > 0 && channels > 0);
  const size_t n = (size_t)batch * pixels * channels;
  if (!(channels & 3) && n / 4 < ((size_t)1 << 31) - ((size_t)1 << 24) && lb_aligned16(x) && lb_aligned16(g) && lb_aligned16(dx) &&
      lb_aligned16(gain)) {
    k_norm_bwd_apply4<<<lb_grid_1d(n / 4, 256), 256, 0, lb_s(s)>>>(x, g, stats, gain, gain_batch_stride, sc, dx, (int)(n / 4),
                                                                  lb_make_fastdiv((uint32_t)((size_t)pixels * channels / 4)),
                                                                  lb_make_fastdiv(channels / 4));
  } else {
    k_norm_bwd_apply<<<lb_grid_1d(n, 256), 256, 0, lb_s(s)>>>(x, g, stats, gain, gain_batch_stride, sc, dx, n, pixels * channels, channels);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
