// Elementwise / streaming kernels: RootTanh, tanh, gated residual, Nadam, copies.
// All HBM-bound: 128-bit accesses, grid-stride over a multiple of the SM count, fp32 math.
#include "common.cuh"

int g_lb_launches = 0;
int g_lb_pdl = -1;

extern "C" int lb_version(void) { return 100; }
extern "C" int lb_sm_arch(void) { return 100; }
extern "C" int lb_last_launch_count(void) { return g_lb_launches; }
extern "C" void lb_reset_launch_count(void) { g_lb_launches = 0; }
extern "C" int lb_set_pdl(int on) { const int was = lb_pdl_on() ? 1 : 0; g_lb_pdl = on ? 1 : 0; return was; }

// ------------------------------------------------------------------------------------------
// generic unary / binary streaming kernels (storage type T: fp32 or bf16, fp32 arithmetic)
// ------------------------------------------------------------------------------------------
template <typename T, typename F>
__global__ void __launch_bounds__(256) k_unary(const T* __restrict__ x, T* __restrict__ y, size_t n, F f) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nv = n / N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float v[N];
    lb_ldv(x + N * i, v);
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = f(v[k]);
    lb_stv(y + N * i, v);
  }
  for (size_t i = nv * N + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) lb_st1(y + i, f(lb_ld1(x + i)));
}

template <typename T, typename F>
__global__ void __launch_bounds__(256) k_binary(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, size_t n, F f) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nv = n / N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float u[N], v[N];
    lb_ldv(a + N * i, u);
    lb_ldv(b + N * i, v);
#pragma unroll
    for (int k = 0; k < N; ++k) u[k] = f(u[k], v[k]);
    lb_stv(y + N * i, u);
  }
  for (size_t i = nv * N + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    lb_st1(y + i, f(lb_ld1(a + i), lb_ld1(b + i)));
}

// scalar fallbacks for unaligned views
template <typename T, typename F>
__global__ void k_unary_s(const T* __restrict__ x, T* __restrict__ y, size_t n, F f) {
  lb_pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) lb_st1(y + i, f(lb_ld1(x + i)));
}
template <typename T, typename F>
__global__ void k_binary_s(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, size_t n, F f) {
  lb_pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) lb_st1(y + i, f(lb_ld1(a + i), lb_ld1(b + i)));
}

template <typename T, typename F>
static int launch_unary_t(const T* x, T* y, size_t n, F f, lb_stream_t s) {
  LB_REQUIRE(x && y);
  if (n == 0) return LB_OK;
  if (lb_vec_ok(x) && lb_vec_ok(y)) {
    lb_launch(k_unary<T, F>, lb_grid_1d((n + LbV<T>::N - 1) / LbV<T>::N, 256), 256, 0, lb_s(s), x, y, n, f);
  } else {
    lb_launch(k_unary_s<T, F>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, y, n, f);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
template <typename T, typename F>
static int launch_binary_t(const T* a, const T* b, T* y, size_t n, F f, lb_stream_t s) {
  LB_REQUIRE(a && b && y);
  if (n == 0) return LB_OK;
  if (lb_vec_ok(a) && lb_vec_ok(b) && lb_vec_ok(y)) {
    lb_launch(k_binary<T, F>, lb_grid_1d((n + LbV<T>::N - 1) / LbV<T>::N, 256), 256, 0, lb_s(s), a, b, y, n, f);
  } else {
    lb_launch(k_binary_s<T, F>, lb_grid_1d(n, 256), 256, 0, lb_s(s), a, b, y, n, f);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
template <typename F>
static int launch_unary(const void* x, void* y, size_t n, F f, int dtype, lb_stream_t s) {
  LB_DISPATCH(dtype, T, return launch_unary_t(lb_cp<T>(x), lb_p<T>(y), n, f, s));
}
template <typename F>
static int launch_binary(const void* a, const void* b, void* y, size_t n, F f, int dtype, lb_stream_t s) {
  LB_DISPATCH(dtype, T, return launch_binary_t(lb_cp<T>(a), lb_cp<T>(b), lb_p<T>(y), n, f, s));
}

struct RootTanh4 { __device__ float operator()(float x) const { return lb_roottanh(x); } };
struct RootTanhG { float ig; __device__ float operator()(float x) const { return lb_roottanh_g(x, ig); } };
struct RootTanhBwd4 { __device__ float operator()(float x, float g) const { return g * lb_roottanh_grad(x); } };
// bf16 storage: the GEMM epilogue's 3-MUFU formulas (common.cuh), the result is rounded to bf16 anyway
struct RootTanh4Fast { __device__ float operator()(float x) const { return lb_roottanh_fast(x); } };
struct RootTanhBwd4Fast { __device__ float operator()(float x, float g) const { return g * lb_roottanh_grad_fast(x); } };
struct RootTanhBwdG { float ig; __device__ float operator()(float x, float g) const { return g * lb_roottanh_grad_g(x, ig); } };
struct TanhF { __device__ float operator()(float x) const { return tanhf(x); } };
struct TanhB { __device__ float operator()(float y, float g) const { return g * (1.0f - y * y); } };
struct HingeF { __device__ float operator()(float x) const { return fmaxf(1.0f - x, 0.0f); } };
struct HingeB { __device__ float operator()(float x, float g) const { return x < 1.0f ? -g : 0.0f; } };
struct ScaleF { float f; __device__ float operator()(float x) const { return x * f; } };
struct AddF { __device__ float operator()(float a, float b) const { return a + b; } };
struct MulF { __device__ float operator()(float a, float b) const { return a * b; } };

extern "C" int lb_roottanh_fwd(const void* x, void* y, size_t n, int growth, int dtype, lb_stream_t s) {
  LB_REQUIRE(growth >= 1);
  if (growth == 4 && dtype == LB_BF16) return launch_unary(x, y, n, RootTanh4Fast{}, dtype, s);
  if (growth == 4) return launch_unary(x, y, n, RootTanh4{}, dtype, s);
  return launch_unary(x, y, n, RootTanhG{1.0f / growth}, dtype, s);
}
extern "C" int lb_roottanh_bwd(const void* x, const void* g, void* dx, size_t n, int growth, int dtype, lb_stream_t s) {
  LB_REQUIRE(growth >= 1);
  if (growth == 4 && dtype == LB_BF16) return launch_binary(x, g, dx, n, RootTanhBwd4Fast{}, dtype, s);
  if (growth == 4) return launch_binary(x, g, dx, n, RootTanhBwd4{}, dtype, s);
  return launch_binary(x, g, dx, n, RootTanhBwdG{1.0f / growth}, dtype, s);
}
extern "C" int lb_tanh_fwd(const void* x, void* y, size_t n, int dtype, lb_stream_t s) { return launch_unary(x, y, n, TanhF{}, dtype, s); }
extern "C" int lb_tanh_bwd(const void* y, const void* g, void* dx, size_t n, int dtype, lb_stream_t s) {
  return launch_binary(y, g, dx, n, TanhB{}, dtype, s);
}
extern "C" int lb_hinge_fwd(const float* x, float* y, size_t n, lb_stream_t s) { return launch_unary(x, y, n, HingeF{}, LB_F32, s); }
extern "C" int lb_hinge_bwd(const float* x, const float* g, float* dx, size_t n, lb_stream_t s) {
  return launch_binary(x, g, dx, n, HingeB{}, LB_F32, s);
}
extern "C" int lb_scale(float* x, size_t n, float factor, lb_stream_t s) { return launch_unary(x, x, n, ScaleF{factor}, LB_F32, s); }
// y = a + b: the sum of the gradients that reach a tensor consumed by two branches (skip path + gated branch of a block)
extern "C" int lb_add(const void* a, const void* b, void* y, size_t n, int dtype, lb_stream_t s) {
  return launch_binary(a, b, y, n, AddF{}, dtype, s);
}
// y = a * b: a gradient times a stored activation derivative (RootTanh' kept by the forward pass)
extern "C" int lb_mul(const void* a, const void* b, void* y, size_t n, int dtype, lb_stream_t s) {
  return launch_binary(a, b, y, n, MulF{}, dtype, s);
}

__global__ void k_fill(float* __restrict__ x, size_t n, float v) {
  lb_pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}
extern "C" int lb_fill(float* x, size_t n, float value, lb_stream_t s) {
  LB_REQUIRE(x);
  if (n == 0) return LB_OK;
  lb_launch(k_fill, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, n, value);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------
// gated residual (libs/merge.py:19-39) on channels-last [B][P][C]
// y: full tensor (storage T) or a per-(b,c) gate [B][C] (storage T) broadcast over pixels
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_gate_fwd(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ gamma,
                           T* __restrict__ out, size_t n, int pc, int channels, int y_bcast) {
  lb_pdl_enter();
  const float gm = __ldg(gamma);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (!y_bcast) {
    const size_t n4 = n >> 2;   // caller guarantees n % 4 == 0 and alignment on this path, else n4 = 0
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 a = lb_ld4(x + 4 * i), b = lb_ld4(y + 4 * i), r;
      r.x = fmaf(gm, b.x, 1.0f) * a.x; r.y = fmaf(gm, b.y, 1.0f) * a.y;
      r.z = fmaf(gm, b.z, 1.0f) * a.z; r.w = fmaf(gm, b.w, 1.0f) * a.w;
      lb_st4(out + 4 * i, r);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const size_t b = i / (size_t)pc;
      const int c = (int)(i % (size_t)channels);
      lb_st1(out + i, fmaf(gm, lb_ld1(y + b * channels + c), 1.0f) * lb_ld1(x + i));
    }
  }
}
template <typename T>
__global__ void k_gate_fwd_s(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ gamma,
                             T* __restrict__ out, size_t n) {
  lb_pdl_enter();
  const float gm = __ldg(gamma);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    lb_st1(out + i, fmaf(gm, lb_ld1(y + i), 1.0f) * lb_ld1(x + i));
}

template <typename T>
static int gate_fwd_t(const T* x, const T* y, const float* gamma, T* out, int batch, int pixels, int channels, int y_bcast, lb_stream_t s) {
  const size_t n = (size_t)batch * pixels * channels;
  if (!y_bcast && !((n & 3) == 0 && lb_vec4_ok(x) && lb_vec4_ok(y) && lb_vec4_ok(out))) {
    lb_launch(k_gate_fwd_s<T>, lb_grid_1d(n, 256), 256, 0, lb_s(s), x, y, gamma, out, n);
  } else {
    lb_launch(k_gate_fwd<T>, lb_grid_1d(y_bcast ? n : n / 4, 256), 256, 0, lb_s(s), x, y, gamma, out, n, pixels * channels, channels, y_bcast);
  }
  LB_LAUNCH_CHECK();
  return LB_OK;
}
extern "C" int lb_gate_fwd(const void* x, const void* y, const float* gamma, void* out, int batch, int pixels,
                           int channels, int y_bcast, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && y && gamma && out && batch > 0 && pixels > 0 && channels > 0);
  LB_DISPATCH(dtype, T, return gate_fwd_t(lb_cp<T>(x), lb_cp<T>(y), gamma, lb_p<T>(out), batch, pixels, channels, y_bcast, s));
}

// Same gate, also accumulating (sum, sum of squares) of the OUTPUT (as stored, i.e. after rounding to T) in fp64: every gate
// output feeds a whole-tensor norm (block.py:46-51), whose statistics pass would otherwise re-read the tensor just written.
template <typename T>
__global__ void __launch_bounds__(256) k_gate_fwd_stats(const T* __restrict__ x, const T* __restrict__ y,
                                                       const float* __restrict__ gamma, T* __restrict__ out, int nv,
                                                       LbFastDiv d_pcv, LbFastDiv d_cv, int channels, int y_bcast,
                                                       double* __restrict__ sums, double* __restrict__ work) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  __shared__ double scratch[32];
  const float gm = __ldg(gamma);
  const int stride = gridDim.x * blockDim.x;
  double s1 = 0.0, s2 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float a[N], b[N];
    lb_ldv(x + (size_t)N * i, a);
    if (y_bcast) {
      int bi, rem, q, cv;
      lb_fast_divmod(d_pcv, i, bi, rem);
      lb_fast_divmod(d_cv, rem, q, cv);
      lb_ldv(y + (size_t)bi * channels + N * cv, b);
    } else {
      lb_ldv(y + (size_t)N * i, b);
    }
    float p1 = 0.0f, p2 = 0.0f;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      a[k] = fmaf(gm, b[k], 1.0f) * a[k];
      const float r = lb_round_as<T>(a[k]);     // statistics of what the norm will read back
      p1 += r;
      p2 = fmaf(r, r, p2);
    }
    lb_stv(out + (size_t)N * i, a);
    s1 += (double)p1;
    s2 += (double)p2;
  }
  s1 = lb_block_sum(s1, scratch);
  s2 = lb_block_sum(s2, scratch);
  lb_grid_sum2_ordered(s1, s2, work, sums, scratch);
}
// sums[2] (fp64) = (sum out, sum out^2), reduced in a fixed order (work: see lb_norm_stats).  Needs channels % (16 bytes of
// elements) == 0 and 16-byte aligned pointers (LB_EALIGN otherwise: use lb_gate_fwd + lb_norm_stats).
extern "C" int lb_gate_fwd_stats(const void* x, const void* y, const float* gamma, void* out, double* sums, double* work,
                                 int batch, int pixels, int channels, int y_bcast, int dtype, lb_stream_t s) {
  LB_REQUIRE(x && y && gamma && out && sums && work && batch > 0 && pixels > 0 && channels > 0);
  const size_t n = (size_t)batch * pixels * channels;
  LB_DISPATCH(dtype, T, {
    constexpr int N = LbV<T>::N;
    if ((channels % N) || n / N >= ((size_t)1 << 31) - ((size_t)1 << 24)) return LB_EALIGN;
    if (!lb_vec_ok(lb_cp<T>(x)) || !lb_vec_ok(lb_cp<T>(y)) || !lb_vec_ok(lb_cp<T>(out))) return LB_EALIGN;
    lb_launch(k_gate_fwd_stats<T>, lb_grid_1d(n / N, 256, 8), 256, 0, lb_s(s), lb_cp<T>(x), lb_cp<T>(y), gamma, lb_p<T>(out), (int)(n / N),
                                                                    lb_make_fastdiv((uint32_t)((size_t)pixels * channels / N)),
                                                                    lb_make_fastdiv(channels / N), channels, y_bcast, sums, work);
  });
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// backward: CTA handles batch b = blockIdx.y, pixel chunk blockIdx.x; thread (cl, pl).  dy of the broadcast gate is fp32
// [B][C] (atomically accumulated over pixel chunks; the caller zeroes it and converts).
template <typename T>
__global__ void k_gate_bwd(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ gamma,
                           const T* __restrict__ g, T* __restrict__ dx, T* __restrict__ dy, float* __restrict__ dy_bcast,
                           float* __restrict__ dgamma, int pixels, int channels, int chunk, int tc, int tp,
                           int y_bcast, int strict) {
  lb_pdl_enter();
  __shared__ float scratch[32];
  const float gm = __ldg(gamma);
  const int cl = threadIdx.x % tc, pl = threadIdx.x / tc;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(pixels, p0 + chunk);
  const size_t base = (size_t)b * pixels * channels;
  float acc_gamma = 0.0f;
  for (int c = (threadIdx.x < tc * tp ? cl : channels); c < channels; c += tc) {
    float acc_dy = 0.0f;
    const float yb = y_bcast ? lb_ld1(y + (size_t)b * channels + c) : 0.0f;
#pragma unroll 4
    for (int p = p0 + pl; p < p1; p += tp) {
      const size_t i = base + (size_t)p * channels + c;
      const float xv = lb_ld1(x + i), gv = lb_ld1(g + i);
      const float yv = y_bcast ? yb : lb_ld1(y + i);
      const float xg = xv * gv;
      lb_st1(dx + i, fmaf(gm, yv, 1.0f) * gv);
      if (y_bcast) acc_dy += xg; else lb_st1(dy + i, xg * gm);
      acc_gamma = fmaf(xg, strict ? xv : yv, acc_gamma);
    }
    if (y_bcast) atomicAdd(dy_bcast + (size_t)b * channels + c, acc_dy * gm);
  }
  if (dgamma) {
    const float tot = lb_block_sum(acc_gamma, scratch);
    if (threadIdx.x == 0) atomicAdd(dgamma, tot);
  }
}

// broadcast gate (feature attention: y is [B][C]), vector form: a thread owns one 16-byte channel vector (8 bf16 / 4 fp32)
// and walks the CTA's pixel chunk; dy[b][c] = gamma * sum_p x*g is reduced over the CTA's pixel lanes in shared memory
// before one atomic per (b, c).  (The scalar kernel above keeps 2 bytes per thread in flight: 1.5 TB/s.)
template <typename T>
__global__ void __launch_bounds__(256) k_gate_bwd_bcast_v(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ gamma,
                                                         const T* __restrict__ g, T* __restrict__ dx, float* __restrict__ dy_bcast,
                                                         float* __restrict__ dgamma, int pixels, int channels, int chunk, int cv, int tp,
                                                         int strict) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  extern __shared__ float s_dy[];                   // [tp][channels]
  __shared__ float scratch[32];
  const float gm = __ldg(gamma);
  const int cl = threadIdx.x % cv, pl = threadIdx.x / cv;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(pixels, p0 + chunk);
  const size_t base = (size_t)b * pixels * channels + (size_t)cl * N;
  float acc_gamma = 0.0f;
  if (pl < tp) {
    float yb[N], fac[N], acc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      yb[k] = lb_ld1(y + (size_t)b * channels + cl * N + k);
      fac[k] = fmaf(gm, yb[k], 1.0f);
      acc[k] = 0.0f;
    }
#pragma unroll 2
    for (int p = p0 + pl; p < p1; p += tp) {
      const size_t i = base + (size_t)p * channels;
      float xv[N], gv[N];
      lb_ldv(x + i, xv);
      lb_ldv(g + i, gv);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float xg = xv[k] * gv[k];
        acc[k] += xg;
        acc_gamma = fmaf(xg, strict ? xv[k] : yb[k], acc_gamma);
        gv[k] *= fac[k];                             // dx
      }
      lb_stv(dx + i, gv);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) s_dy[pl * channels + cl * N + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < channels; c += blockDim.x) {
    float t = 0.0f;
    for (int q = 0; q < tp; ++q) t += s_dy[q * channels + c];
    atomicAdd(dy_bcast + (size_t)b * channels + c, t * gm);
  }
  if (dgamma) {
    const float tot = lb_block_sum(acc_gamma, scratch);
    if (threadIdx.x == 0) atomicAdd(dgamma, tot);
  }
}

// full-shape gate (y has the shape of x): everything is elementwise except d-gamma, so the kernel streams 4-vectors and
// reduces one scalar per CTA
template <typename T>
__global__ void __launch_bounds__(256) k_gate_bwd4(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ gamma,
                                                  const T* __restrict__ g, T* __restrict__ dx, T* __restrict__ dy,
                                                  float* __restrict__ dgamma, size_t nv, int strict) {
  lb_pdl_enter();
  constexpr int N = LbV<T>::N;
  __shared__ float scratch[32];
  const float gm = __ldg(gamma);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float xv[N], yv[N], gv[N];
    lb_ldv(x + N * i, xv);
    lb_ldv(y + N * i, yv);
    lb_ldv(g + N * i, gv);
    float part = 0.0f;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float xg = xv[k] * gv[k];
      part = fmaf(xg, strict ? xv[k] : yv[k], part);
      yv[k] = fmaf(gm, yv[k], 1.0f) * gv[k];      // dx
      xv[k] = xg * gm;                           // dy
    }
    lb_stv(dx + N * i, yv);
    lb_stv(dy + N * i, xv);
    acc += part;
  }
  if (dgamma) {
    const float tot = lb_block_sum(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(dgamma, tot);
  }
}

template <typename T>
static int gate_bwd_t(const T* x, const T* y, const float* gamma, const T* g, T* dx, T* dy, float* dy_bcast, float* dgamma, int batch,
                      int pixels, int channels, int y_bcast, int strict_reference, lb_stream_t s) {
  const size_t n = (size_t)batch * pixels * channels;
  if (!y_bcast && !(n % LbV<T>::N) && lb_vec_ok(x) && lb_vec_ok(y) && lb_vec_ok(g) && lb_vec_ok(dx) && lb_vec_ok(dy)) {
    lb_launch(k_gate_bwd4<T>, lb_grid_1d(n / LbV<T>::N, 256, 8), 256, 0, lb_s(s), x, y, gamma, g, dx, dy, dgamma, n / LbV<T>::N, strict_reference);
    LB_LAUNCH_CHECK();
    return LB_OK;
  }
  if (y_bcast && !(channels % LbV<T>::N) && channels / LbV<T>::N <= 256 && lb_vec_ok(x) && lb_vec_ok(g) && lb_vec_ok(dx)) {
    const int cv = channels / LbV<T>::N, tp = 256 / cv;
    int chunks = (LB_SMS * 4 + batch - 1) / batch;
    int chunk = (pixels + chunks - 1) / chunks;
    if (chunk < 4 * tp) chunk = 4 * tp;
    chunks = (pixels + chunk - 1) / chunk;
    lb_launch(k_gate_bwd_bcast_v<T>, dim3(chunks, batch), 256, (size_t)tp * channels * sizeof(float), lb_s(s), x, y, gamma, g, dx, dy_bcast,
              dgamma, pixels, channels, chunk, cv, tp, strict_reference);
    LB_LAUNCH_CHECK();
    return LB_OK;
  }
  const LbColShape sh = lb_col_shape(channels);
  // chunk: aim for ~4 waves of CTAs, at least tp pixels each
  int chunks = (LB_SMS * 4 + batch - 1) / batch;
  int chunk = (pixels + chunks - 1) / chunks;
  if (chunk < sh.tp) chunk = sh.tp;
  chunks = (pixels + chunk - 1) / chunk;
  dim3 grid(chunks, batch);
  lb_launch(k_gate_bwd<T>, grid, sh.threads, 0, lb_s(s), x, y, gamma, g, dx, dy, dy_bcast, dgamma, pixels, channels, chunk, sh.tc, sh.tp,
                                                  y_bcast, strict_reference);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
// dy: storage T, shape of x (y_bcast = 0);  dy_bcast: fp32 [B][C], zeroed by the caller (y_bcast = 1)
extern "C" int lb_gate_bwd(const void* x, const void* y, const float* gamma, const void* g, void* dx, void* dy, float* dy_bcast,
                           float* dgamma, int batch, int pixels, int channels, int y_bcast, int strict_reference, int dtype,
                           lb_stream_t s) {
  LB_REQUIRE(x && y && gamma && g && dx && batch > 0 && pixels > 0 && channels > 0);
  LB_REQUIRE(y_bcast ? dy_bcast != nullptr : dy != nullptr);
  LB_DISPATCH(dtype, T, return gate_bwd_t(lb_cp<T>(x), lb_cp<T>(y), gamma, lb_cp<T>(g), lb_p<T>(dx), lb_p<T>(dy), dy_bcast, dgamma,
                                          batch, pixels, channels, y_bcast, strict_reference, s));
}

// ------------------------------------------------------------------------------------------
// Nadam over a flat arena (libs/nadam.py:75-87)
// ------------------------------------------------------------------------------------------
// The step-dependent scalars live on the device so a captured CUDA graph of the whole training step replays
// correctly: k_nadam_schedule advances state = {t, m_schedule} and writes hyper = {c_grad, c_mom, 1/bias2}
// (nadam.py:62-73,78,82,85) in double precision; k_nadam reads them.
__global__ void k_nadam_schedule(double* __restrict__ state, float* __restrict__ hyper, double lr, double b1, double b2, double decay) {
  lb_pdl_enter();
  const double t = state[0] + 1.0;
  const double mu_t = b1 * (1.0 - 0.5 * pow(0.96, t * decay));
  const double mu_next = b1 * (1.0 - 0.5 * pow(0.96, (t + 1.0) * decay));
  const double sched_new = state[1] * mu_t;
  const double sched_next = sched_new * mu_next;
  state[0] = t;
  state[1] = sched_new;
  hyper[0] = (float)(lr * (1.0 - mu_t) / (1.0 - sched_new));
  hyper[1] = (float)(lr * mu_next / (1.0 - sched_next));
  hyper[2] = (float)(1.0 / (1.0 - pow(b2, t)));
}
extern "C" int lb_nadam_schedule(double* state, float* hyper, double lr, double beta1, double beta2, double schedule_decay,
                                 lb_stream_t s) {
  LB_REQUIRE(state && hyper);
  lb_launch(k_nadam_schedule, 1, 1, 0, lb_s(s), state, hyper, lr, beta1, beta2, schedule_decay);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

__global__ void __launch_bounds__(256) k_nadam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                               float* __restrict__ v, size_t n, float b1, float b2, float eps,
                                               const float* __restrict__ hyper) {
  lb_pdl_enter();
  const float c_grad = __ldg(hyper), c_mom = __ldg(hyper + 1), inv_bias2 = __ldg(hyper + 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gv = g[i];
    const float mv = fmaf(b1, m[i], (1.0f - b1) * gv);
    const float vv = fmaf(b2, v[i], (1.0f - b2) * gv * gv);
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv * inv_bias2) + eps;
    float pv = p[i];
    pv -= c_grad * gv / denom;
    pv -= c_mom * mv / denom;
    p[i] = pv;
  }
}
extern "C" int lb_nadam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float beta1,
                             float beta2, float eps, const float* hyper, lb_stream_t s) {
  LB_REQUIRE(param && grad && exp_avg && exp_avg_sq && hyper);
  if (n == 0) return LB_OK;
  lb_launch(k_nadam, lb_grid_1d(n, 256), 256, 0, lb_s(s), param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, hyper);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ------------------------------------------------------------------------------------------
// strided row copy (channel concat / slice) and layout changes
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void k_copy_rows(const TI* __restrict__ src, int ld_src, TO* __restrict__ dst, int ld_dst, size_t rows,
                            int cols, int accumulate) {
  lb_pdl_enter();
  const size_t n = rows * (size_t)cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const size_t r = i / cols;
    const int c = (int)(i % cols);
    const float v = lb_ld1(src + r * ld_src + c);
    TO* d = dst + r * ld_dst + c;
    lb_st1(d, accumulate ? lb_ld1(d) + v : v);
  }
}
// 4-element version (cols, both leading dimensions multiples of 4, aligned pointers), 32-bit multiply-shift decode
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_copy_rows4(const TI* __restrict__ src, int ld_src, TO* __restrict__ dst, int ld_dst,
                                                   int n4, LbFastDiv d_c4, int accumulate) {
  lb_pdl_enter();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    int r, c4;
    lb_fast_divmod(d_c4, i, r, c4);
    float4 v = lb_ld4(src + (size_t)r * ld_src + 4 * c4);
    TO* d = dst + (size_t)r * ld_dst + 4 * c4;
    if (accumulate) {
      const float4 o = lb_ld4(d);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    lb_st4(d, v);
  }
}
template <typename TI, typename TO>
static int copy_rows_t(const TI* src, int ld_src, TO* dst, int ld_dst, int64_t rows, int cols, int accumulate, lb_stream_t s) {
  const size_t n = (size_t)rows * cols;
  if (!(cols & 3) && !(ld_src & 3) && !(ld_dst & 3) && lb_vec4_ok(src) && lb_vec4_ok(dst) && n / 4 < ((size_t)1 << 31) - ((size_t)1 << 24))
    lb_launch(k_copy_rows4<TI, TO>, lb_grid_1d(n / 4, 256), 256, 0, lb_s(s), src, ld_src, dst, ld_dst, (int)(n / 4), lb_make_fastdiv(cols / 4), accumulate);
  else
    lb_launch(k_copy_rows<TI, TO>, lb_grid_1d(n, 256), 256, 0, lb_s(s), src, ld_src, dst, ld_dst, (size_t)rows, cols, accumulate);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
// src / dst may differ in storage type (a concat of a 3-channel fp32 image into a wide bf16 tensor, and its backward)
extern "C" int lb_copy_rows(const void* src, int ld_src, void* dst, int ld_dst, int64_t rows, int cols, int accumulate, int src_dtype,
                            int dst_dtype, lb_stream_t s) {
  LB_REQUIRE(src && dst && rows >= 0 && cols > 0 && ld_src >= cols && ld_dst >= cols);
  if (rows == 0) return LB_OK;
  LB_DISPATCH(src_dtype, TI, LB_DISPATCH(dst_dtype, TO, return copy_rows_t(lb_cp<TI>(src), ld_src, lb_p<TO>(dst), ld_dst, rows, cols,
                                                                          accumulate, s)));
}

// [B][C][HW] <-> [B][HW][C] through a 32x33 shared tile (coalesced on both sides); TI / TO: storage of source / destination
template <typename TI, typename TO>
__global__ void k_transpose_batched(const TI* __restrict__ x, TO* __restrict__ y, int rows, int cols) {
  lb_pdl_enter();
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = lb_ld1(x + base + (size_t)r * cols + c);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) lb_st1(y + base + (size_t)c * rows + r, tile[threadIdx.x][j]);
  }
}
template <typename TI, typename TO>
static int transpose_batched(const TI* x, TO* y, int batch, int rows, int cols, lb_stream_t s) {
  LB_REQUIRE(x && y && batch > 0 && rows > 0 && cols > 0 && batch <= 65535);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  LB_REQUIRE(grid.y <= 65535);
  lb_launch(k_transpose_batched<TI, TO>, grid, dim3(32, 8), 0, lb_s(s), x, y, rows, cols);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
// Few channels (RGB images at the model boundary): a 32x32 tile transpose would run 3 of its 32 rows.  One thread = one
// pixel: the C plane reads / writes are coalesced across threads, the C interleaved values are one short contiguous run.
template <bool kToNhwc>
__global__ void __launch_bounds__(256) k_layout_small_c(const float* __restrict__ x, float* __restrict__ y, int c, int hw, size_t pixels,
                                                       LbFastDiv d_hw) {
  lb_pdl_enter();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += stride) {
    int b, p;
    lb_fast_divmod(d_hw, (int)i, b, p);                 // pixels < 2^31 (checked by the host)
    const size_t plane0 = (size_t)b * c * hw + p;       // NCHW offset of (b, 0, p)
    const size_t inter0 = i * c;                        // NHWC offset of (b, p, 0)
    for (int k = 0; k < c; ++k) {
      if (kToNhwc) y[inter0 + k] = __ldg(x + plane0 + (size_t)k * hw);
      else y[plane0 + (size_t)k * hw] = __ldg(x + inter0 + k);
    }
  }
}
static int layout_small_c(const float* x, float* y, int batch, int c, int hw, bool to_nhwc, lb_stream_t s) {
  const size_t pixels = (size_t)batch * hw;
  LB_REQUIRE(x && y && batch > 0 && c > 0 && hw > 0 && pixels < ((size_t)1 << 31) - ((size_t)1 << 24));
  if (to_nhwc) lb_launch(k_layout_small_c<true>, lb_grid_1d(pixels, 256), 256, 0, lb_s(s), x, y, c, hw, pixels, lb_make_fastdiv(hw));
  else lb_launch(k_layout_small_c<false>, lb_grid_1d(pixels, 256), 256, 0, lb_s(s), x, y, c, hw, pixels, lb_make_fastdiv(hw));
  LB_LAUNCH_CHECK();
  return LB_OK;
}
// x: fp32 NCHW (what the reference's loaders and callers hand in); y: channels-last with storage `out_dtype`
extern "C" int lb_nchw_to_nhwc(const float* x, void* y, int batch, int c, int hw, int out_dtype, lb_stream_t s) {
  if (out_dtype == LB_F32 && c <= 8 && (size_t)batch * hw < ((size_t)1 << 31) - ((size_t)1 << 24))
    return layout_small_c(x, lb_p<float>(y), batch, c, hw, true, s);
  LB_DISPATCH(out_dtype, T, return transpose_batched(x, lb_p<T>(y), batch, c, hw, s));
}
// x: channels-last with storage `in_dtype`; y: fp32 NCHW
extern "C" int lb_nhwc_to_nchw(const void* x, float* y, int batch, int c, int hw, int in_dtype, lb_stream_t s) {
  if (in_dtype == LB_F32 && c <= 8 && (size_t)batch * hw < ((size_t)1 << 31) - ((size_t)1 << 24))
    return layout_small_c(lb_cp<float>(x), y, batch, c, hw, false, s);
  LB_DISPATCH(in_dtype, T, return transpose_batched(lb_cp<T>(x), y, batch, hw, c, s));
}
