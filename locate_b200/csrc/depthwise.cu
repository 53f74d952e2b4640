// SEPARABLE = True (libs/config.py:53): the two grouped convolutions the reference builds then.
//   * depthwise k x k conv / transposed conv, groups = channels (libs/conv.py:17, FEATURE_MULTIPLIER = 1):
//       out[b,oy,ox,c] = alpha * sum_{ty,tx} in[b,iy,ix,c] * w[c][ty][tx]        (mode 0: iy = oy*s - pad + ty;
//                                                                                 mode 1: iy = (oy + pad - ty)/s if exact)
//     forward of one, input gradient of the other (same weights, mode flipped), and the weight gradient.
//   * the feature-attention "full-extent" grouped conv (libs/attention.py:15-21): [B,F,S,S] -> [B,F/r,1,1] with
//     groups = F/r, i.e. out[b,o] = alpha * sum_{j<r} sum_p in[b,p,o*r+j] * w[o][j][p].
// One multiply-add per byte moved: HBM-bound streaming kernels, no tensor cores.  Channels are the contiguous axis of
// channels-last data, so consecutive threads take consecutive channels.  Storage type T (fp32 / bf16), fp32 arithmetic.
#include "common.cuh"

namespace {

struct DwP {
  int batch, in_h, in_w, out_h, out_w, c, kh, kw, stride, pad, mode;
  LbFastDiv d_c, d_w, d_h;
};

__device__ __forceinline__ bool dw_src(int mode, int stride, int pad, int o, int t, int extent, int& i) {
  if (mode == 0) {
    i = o * stride - pad + t;
  } else {
    const int v = o + pad - t;
    if (v < 0 || v % stride) return false;
    i = v / stride;
  }
  return i >= 0 && i < extent;
}

// w: fp32 [c][kh][kw] (Conv2d (C,1,k,k) and ConvTranspose2d (C,1,k,k) alike)
template <typename T>
__global__ void __launch_bounds__(256) k_dw_conv(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ alpha,
                                                 T* __restrict__ out, int n, const DwP p) {
  lb_pdl_enter();
  const float a = alpha ? __ldg(alpha) : 1.0f;
  const int taps = p.kh * p.kw;
  const int stride_t = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride_t) {
    int t, c, ox, oy, b;
    lb_fast_divmod(p.d_c, i, t, c);
    lb_fast_divmod(p.d_w, t, t, ox);
    lb_fast_divmod(p.d_h, t, b, oy);
    const float* wc = w + (size_t)c * taps;
    float acc = 0.0f;
    for (int ty = 0; ty < p.kh; ++ty) {
      int iy;
      if (!dw_src(p.mode, p.stride, p.pad, oy, ty, p.in_h, iy)) continue;
      for (int tx = 0; tx < p.kw; ++tx) {
        int ix;
        if (!dw_src(p.mode, p.stride, p.pad, ox, tx, p.in_w, ix)) continue;
        acc = fmaf(lb_ld1(in + ((size_t)(b * p.in_h + iy) * p.in_w + ix) * p.c + c), __ldg(wc + ty * p.kw + tx), acc);
      }
    }
    lb_st1(out + i, acc * a);
  }
}

// dw[c][ty][tx] += sum over dense pixels of gathered[pixel @ tap][c] * dense[pixel][c]   (mode-0 gather around the dense grid)
// grid = (channel blocks, taps, pixel splits); fp32 atomics across the pixel splits.
template <typename T>
__global__ void __launch_bounds__(256) k_dw_wgrad(const T* __restrict__ gath, const T* __restrict__ dense, float* __restrict__ dw,
                                                  int g_h, int g_w, int d_h, int d_w, int batch, int c, int kh, int kw, int stride,
                                                  int pad, int rows_per_split) {
  lb_pdl_enter();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const int tap = blockIdx.y, ty = tap / kw, tx = tap % kw;
  const int rows = batch * d_h;                       // (b, oy) pairs
  const int r0 = blockIdx.z * rows_per_split, r1 = min(rows, r0 + rows_per_split);
  float acc = 0.0f;
  for (int r = r0; r < r1; ++r) {
    const int b = r / d_h, oy = r - b * d_h;
    const int iy = oy * stride - pad + ty;
    if (iy < 0 || iy >= g_h) continue;
    for (int ox = 0; ox < d_w; ++ox) {
      const int ix = ox * stride - pad + tx;
      if (ix < 0 || ix >= g_w) continue;
      acc = fmaf(lb_ld1(gath + ((size_t)(b * g_h + iy) * g_w + ix) * c + ch), lb_ld1(dense + ((size_t)(b * d_h + oy) * d_w + ox) * c + ch), acc);
    }
  }
  atomicAdd(dw + (size_t)ch * kh * kw + tap, acc);
}

// ---- grouped full-extent conv ---------------------------------------------------------------------------------
// forward: CTA = one sample, thread = one input channel; r consecutive channels form a group (one output)
template <typename T>
__global__ void __launch_bounds__(1024) k_gfull_fwd(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ alpha,
                                                    T* __restrict__ out, int pixels, int f, int r) {
  lb_pdl_enter();
  extern __shared__ float s_acc[];
  const int b = blockIdx.x;
  const float a = alpha ? __ldg(alpha) : 1.0f;
  for (int c = threadIdx.x; c < f; c += blockDim.x) {
    const T* src = in + (size_t)b * pixels * f + c;
    const float* wc = w + (size_t)c * pixels;            // w[o][j][p] with c = o*r + j
    float acc = 0.0f;
    for (int p = 0; p < pixels; ++p) acc = fmaf(lb_ld1(src + (size_t)p * f), __ldg(wc + p), acc);
    s_acc[c] = acc;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < f / r; o += blockDim.x) {
    float tot = 0.0f;
    for (int j = 0; j < r; ++j) tot += s_acc[o * r + j];
    lb_st1(out + (size_t)b * (f / r) + o, tot * a);
  }
}
// input gradient: din[b,p,c] = alpha * g[b, c/r] * w[c][p]
template <typename T>
__global__ void __launch_bounds__(256) k_gfull_dgrad(const T* __restrict__ g, const float* __restrict__ w, const float* __restrict__ alpha,
                                                     T* __restrict__ din, size_t n, int pixels, int f, int r) {
  lb_pdl_enter();
  const float a = alpha ? __ldg(alpha) : 1.0f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % f);
    const size_t bp = i / f;
    const int p = (int)(bp % pixels);
    const size_t b = bp / pixels;
    lb_st1(din + i, a * lb_ld1(g + b * (f / r) + c / r) * __ldg(w + (size_t)c * pixels + p));
  }
}
// weight gradient: dw[c][p] += sum_b in[b,p,c] * g[b, c/r]; thread = (p, c), fixed order over the batch (no atomics)
template <typename T>
__global__ void __launch_bounds__(256) k_gfull_wgrad(const T* __restrict__ in, const T* __restrict__ g, float* __restrict__ dw, int batch,
                                                     int pixels, int f, int r) {
  lb_pdl_enter();
  const size_t n = (size_t)pixels * f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % f), p = (int)(i / f);
    float acc = 0.0f;
    for (int b = 0; b < batch; ++b) acc = fmaf(lb_ld1(in + ((size_t)b * pixels + p) * f + c), lb_ld1(g + (size_t)b * (f / r) + c / r), acc);
    dw[(size_t)c * pixels + p] += acc;
  }
}

}  // namespace

extern "C" int lb_dw_conv(const void* in, const float* w, const float* alpha, void* out, int batch, int in_h, int in_w, int out_h,
                          int out_w, int channels, int kh, int kw, int stride, int pad, int mode, int dtype, lb_stream_t s) {
  LB_REQUIRE(in && w && out && batch > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && channels > 0 && kh > 0 && kw > 0);
  LB_REQUIRE(stride >= 1 && pad >= 0 && (mode == 0 || mode == 1));
  const size_t n = (size_t)batch * out_h * out_w * channels;
  LB_REQUIRE(n < ((size_t)1 << 31) - ((size_t)1 << 24));
  DwP p;
  p.batch = batch; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.c = channels;
  p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.mode = mode;
  p.d_c = lb_make_fastdiv(channels); p.d_w = lb_make_fastdiv(out_w); p.d_h = lb_make_fastdiv(out_h);
  LB_DISPATCH(dtype, T, lb_launch(k_dw_conv<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), lb_cp<T>(in), w, alpha, lb_p<T>(out), (int)n, p));
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// gathered / dense as lb_conv_wgrad: for a depthwise Conv gathered = x, dense = dy; for a depthwise ConvTranspose
// gathered = dy, dense = x.  dw: fp32 [c][kh][kw], accumulated (+=).
extern "C" int lb_dw_wgrad(const void* gathered, const void* dense, float* dw, int batch, int g_h, int g_w, int d_h, int d_w,
                           int channels, int kh, int kw, int stride, int pad, int dtype, lb_stream_t s) {
  LB_REQUIRE(gathered && dense && dw && batch > 0 && g_h > 0 && g_w > 0 && d_h > 0 && d_w > 0 && channels > 0 && kh > 0 && kw > 0);
  const int rows = batch * d_h;
  const int cblocks = (channels + 255) / 256;
  long long splits = (LB_SMS * 4 + (long long)cblocks * kh * kw - 1) / ((long long)cblocks * kh * kw);
  if (splits > rows) splits = rows;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const int rps = (int)((rows + splits - 1) / splits);
  splits = (rows + rps - 1) / rps;
  LB_REQUIRE(kh * kw <= 65535);
  dim3 grid(cblocks, kh * kw, (unsigned)splits);
  LB_DISPATCH(dtype, T, lb_launch(k_dw_wgrad<T>, grid, 256, 0, lb_s(s), lb_cp<T>(gathered), lb_cp<T>(dense), dw, g_h, g_w, d_h, d_w, batch,
                                                             channels, kh, kw, stride, pad, rps));
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// in: [B][P][F] channels-last; w: fp32 [F/r][r][P]; out: [B][F/r]
extern "C" int lb_gfull_fwd(const void* in, const float* w, const float* alpha, void* out, int batch, int pixels, int features,
                            int group_in, int dtype, lb_stream_t s) {
  LB_REQUIRE(in && w && out && batch > 0 && pixels > 0 && features > 0 && group_in > 0 && features % group_in == 0);
  LB_REQUIRE(features <= 12 * 1024);
  const int threads = features < 1024 ? (features + 31) / 32 * 32 : 1024;
  LB_DISPATCH(dtype, T, lb_launch(k_gfull_fwd<T>, batch, threads, (size_t)features * sizeof(float), lb_s(s), lb_cp<T>(in), w, alpha, lb_p<T>(out),
                                                                                                  pixels, features, group_in));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_gfull_dgrad(const void* g, const float* w, const float* alpha, void* din, int batch, int pixels, int features,
                              int group_in, int dtype, lb_stream_t s) {
  LB_REQUIRE(g && w && din && batch > 0 && pixels > 0 && features > 0 && group_in > 0 && features % group_in == 0);
  const size_t n = (size_t)batch * pixels * features;
  LB_DISPATCH(dtype, T, lb_launch(k_gfull_dgrad<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), lb_cp<T>(g), w, alpha, lb_p<T>(din), n, pixels, features,
                                                                              group_in));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_gfull_wgrad(const void* in, const void* g, float* dw, int batch, int pixels, int features, int group_in, int dtype,
                              lb_stream_t s) {
  LB_REQUIRE(in && g && dw && batch > 0 && pixels > 0 && features > 0 && group_in > 0 && features % group_in == 0);
  const size_t n = (size_t)pixels * features;
  LB_DISPATCH(dtype, T, lb_launch(k_gfull_wgrad<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), lb_cp<T>(in), lb_cp<T>(g), dw, batch, pixels, features,
                                                                              group_in));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
