// Direct fp32 kernels for the layers whose channel counts are too small for a 128x64 tensor-core tile to make
// sense: the discriminator stem (5x5/s2 3->3, 1x1 3->32, 1x1 3->29) and the generator's final 1x1 48->3, in all
// three directions.  They are HBM-bound on the wide side of the layer (AI <= 16 flop/B), so the job is to touch
// every activation once, coalesced, with the tiny weight held in shared memory and the neighbouring elementwise
// ops fused: RootTanh on the input (forward), multiplication by RootTanh'(x) on the output (input gradient).
//
//   lb_conv_small:       out[p][n] = alpha * sum_{tap,k} act(in[p@tap][k]) * W(tap,k,n) (+bias[n])   [* act'(xpre[p][n])]
//   lb_conv_small_wgrad: dw(tap,kg,kd) += sum_p gathered[p@tap][kg] * dense[p][kd]
// Geometry and weight addressing are those of lb_conv_gemm / lb_conv_wgrad (include/locate_b200.h).
#include "common.cuh"

#define SMALL_MAX_W 4096      // floats of weight held in shared memory (taps*K*N)

struct SmallP {
  const float* in; const float* w; const float* alpha; const float* bias; const float* xpre; float* out;
  int batch, in_h, in_w, in_c, out_h, out_w, out_c;
  int kh, kw, stride, pad, mode, ld_in, ld_out, ld_xpre;
  long long w_sk, w_sn, w_sty, w_stx;
  int growth_in;      // > 0: RootTanh(growth) applied to every input element on load
  int growth_out;     // > 0: result multiplied by RootTanh'(xpre[p][n])
  int ngroups;        // ceil(out_c / 4)
  long long pixels;   // batch*out_h*out_w
  LbFastDiv d_grp, d_w, d_h;
};

// one thread = one output pixel x 4 consecutive output channels
__global__ void __launch_bounds__(256) k_conv_small(const SmallP p) {
  extern __shared__ float ws[];                       // [tap][k][n_pad], n_pad = ngroups*4
  const int taps = p.kh * p.kw, npad = p.ngroups * 4;
  for (int i = threadIdx.x; i < taps * p.in_c * npad; i += blockDim.x) {
    const int n = i % npad, k = (i / npad) % p.in_c, tap = i / (npad * p.in_c);
    ws[i] = n < p.out_c ? __ldg(p.w + k * p.w_sk + n * p.w_sn + (tap / p.kw) * p.w_sty + (tap % p.kw) * p.w_stx) : 0.0f;
  }
  __syncthreads();
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
  const int total = (int)(p.pixels * p.ngroups);          // < 2^31 (checked by the host)
  const int stride = gridDim.x * blockDim.x;
  const bool vec_in = (p.in_c & 3) == 0 && (p.ld_in & 3) == 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int t, grp, ox, oy, b;
    lb_fast_divmod(p.d_grp, i, t, grp);
    lb_fast_divmod(p.d_w, t, t, ox);
    lb_fast_divmod(p.d_h, t, b, oy);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ty = 0; ty < p.kh; ++ty) {
      int iy;
      if (p.mode == 0) iy = oy * p.stride - p.pad + ty;
      else { const int v = oy + p.pad - ty; if (v < 0 || v % p.stride) continue; iy = v / p.stride; }
      if (iy < 0 || iy >= p.in_h) continue;
      for (int tx = 0; tx < p.kw; ++tx) {
        int ix;
        if (p.mode == 0) ix = ox * p.stride - p.pad + tx;
        else { const int v = ox + p.pad - tx; if (v < 0 || v % p.stride) continue; ix = v / p.stride; }
        if (ix < 0 || ix >= p.in_w) continue;
        const float* src = p.in + ((size_t)(b * p.in_h + iy) * p.in_w + ix) * p.ld_in;
        const float* wt = ws + (size_t)(ty * p.kw + tx) * p.in_c * npad + grp * 4;
        if (vec_in) {
          for (int k = 0; k < p.in_c; k += 4) {
            const float4 a4 = lb_ld4(src + k);
            float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = p.growth_in == 4 ? lb_roottanh(av[j]) : (p.growth_in > 0 ? lb_roottanh_g(av[j], 1.0f / p.growth_in) : av[j]);
              const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)(k + j) * npad);
              acc[0] = fmaf(a, w4.x, acc[0]); acc[1] = fmaf(a, w4.y, acc[1]);
              acc[2] = fmaf(a, w4.z, acc[2]); acc[3] = fmaf(a, w4.w, acc[3]);
            }
          }
        } else {
          for (int k = 0; k < p.in_c; ++k) {
            float a = __ldg(src + k);
            if (p.growth_in == 4) a = lb_roottanh(a); else if (p.growth_in > 0) a = lb_roottanh_g(a, 1.0f / p.growth_in);
            const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)k * npad);
            acc[0] = fmaf(a, w4.x, acc[0]); acc[1] = fmaf(a, w4.y, acc[1]);
            acc[2] = fmaf(a, w4.z, acc[2]); acc[3] = fmaf(a, w4.w, acc[3]);
          }
        }
      }
    }
    const size_t pix = (size_t)(b * p.out_h + oy) * p.out_w + ox;
    float* dst = p.out + pix * p.ld_out + grp * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = grp * 4 + j;
      if (n >= p.out_c) break;
      float r = acc[j] * alpha + (p.bias ? __ldg(p.bias + n) : 0.0f);
      if (p.growth_out > 0) {
        const float xv = __ldg(p.xpre + pix * p.ld_xpre + n);
        r *= p.growth_out == 4 ? lb_roottanh_grad(xv) : lb_roottanh_grad_g(xv, 1.0f / p.growth_out);
      }
      dst[j] = r;
    }
  }
}

extern "C" int lb_conv_small_supported(const lb_conv_geom* g) {
  if (!g) return 0;
  const long long wf = (long long)g->kh * g->kw * g->in_c * ((g->out_c + 3) / 4 * 4);
  // one side of the layer is tiny and the whole weight fits in shared memory
  return (wf <= SMALL_MAX_W && (g->in_c <= 4 || g->out_c <= 4 || g->in_c * g->out_c <= 128)) ? 1 : 0;
}

extern "C" int lb_conv_small(const float* in, const float* w, const float* alpha, const float* bias, float* out,
                             const lb_conv_geom* g, int growth_in, const float* xpre, int ld_xpre, int growth_out,
                             lb_stream_t s) {
  LB_REQUIRE(in && w && out && g && growth_in >= 0 && growth_out >= 0 && (growth_out == 0 || xpre));
  if (!lb_conv_small_supported(g)) return LB_EUNSUPPORTED;
  SmallP p;
  p.in = in; p.w = w; p.alpha = alpha; p.bias = bias; p.xpre = xpre; p.out = out;
  p.batch = g->batch; p.in_h = g->in_h; p.in_w = g->in_w; p.in_c = g->in_c;
  p.out_h = g->out_h; p.out_w = g->out_w; p.out_c = g->out_c;
  p.kh = g->kh; p.kw = g->kw; p.stride = g->stride; p.pad = g->pad; p.mode = g->mode;
  p.ld_in = g->ld_in; p.ld_out = g->ld_out; p.ld_xpre = ld_xpre;
  p.w_sk = g->w_sk; p.w_sn = g->w_sn; p.w_sty = g->w_sty; p.w_stx = g->w_stx;
  p.growth_in = growth_in; p.growth_out = growth_out;
  p.ngroups = (g->out_c + 3) / 4;
  p.pixels = (long long)g->batch * g->out_h * g->out_w;
  if (p.pixels * p.ngroups >= (1ll << 31) - (1ll << 24)) return LB_EUNSUPPORTED;
  p.d_grp = lb_make_fastdiv(p.ngroups); p.d_w = lb_make_fastdiv(g->out_w); p.d_h = lb_make_fastdiv(g->out_h);
  const size_t smem = (size_t)g->kh * g->kw * g->in_c * p.ngroups * 4 * sizeof(float);
  k_conv_small<<<lb_grid_1d((size_t)p.pixels * p.ngroups, 256, 16), 256, smem, lb_s(s)>>>(p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- weight gradient: tiny output, reduction over every pixel ------------------------------------------------
// thread e owns one weight element (tap, kg, kd); a CTA walks a chunk of DENSE pixels; the dense row of the current
// pixel is broadcast from shared memory, the gathered element from L1.
struct SmallWgP {
  const float* gath; const float* dense; float* dw;
  int g_h, g_w, g_c, d_h, d_w, d_c, kh, kw, stride, pad, ld_g, ld_d;
  long long w_sk, w_sn, w_sty, w_stx;
  long long pixels; int chunk, n_elems, growth_g;
};
#define WG_TILE 64
__global__ void __launch_bounds__(256) k_conv_small_wgrad(const SmallWgP p) {
  __shared__ float sd[WG_TILE][33];        // dense rows of the tile (d_c <= 32)
  __shared__ int s_iy[WG_TILE], s_ix[WG_TILE], s_b[WG_TILE];   // gather origin of every dense pixel of the tile
  const int e = threadIdx.x;
  const bool live = e < p.n_elems;
  int kd = 0, kg = 0, ty = 0, tx = 0;
  if (live) {
    kd = e % p.d_c;
    int r = e / p.d_c;
    kg = r % p.g_c; r /= p.g_c;
    tx = r % p.kw; ty = r / p.kw;
  }
  const long long p0 = (long long)blockIdx.x * p.chunk;
  const long long p1 = min(p.pixels, p0 + p.chunk);
  float acc = 0.0f;
  for (long long base = p0; base < p1; base += WG_TILE) {
    const int cnt = (int)min((long long)WG_TILE, p1 - base);
    __syncthreads();
    if (threadIdx.x < cnt) {
      const long long pix = base + threadIdx.x;
      const int ox = (int)(pix % p.d_w);
      const int oy = (int)((pix / p.d_w) % p.d_h);
      s_b[threadIdx.x] = (int)(pix / ((long long)p.d_w * p.d_h));
      s_iy[threadIdx.x] = oy * p.stride - p.pad;
      s_ix[threadIdx.x] = ox * p.stride - p.pad;
    }
    for (int i = threadIdx.x; i < cnt * p.d_c; i += blockDim.x) {
      const int r = i / p.d_c, c = i % p.d_c;
      sd[r][c] = __ldg(p.dense + (size_t)(base + r) * p.ld_d + c);
    }
    __syncthreads();
    if (live) {
      for (int r = 0; r < cnt; ++r) {
        const int iy = s_iy[r] + ty, ix = s_ix[r] + tx;
        if (iy < 0 || iy >= p.g_h || ix < 0 || ix >= p.g_w) continue;
        float gv = __ldg(p.gath + ((size_t)(s_b[r] * p.g_h + iy) * p.g_w + ix) * p.ld_g + kg);
        if (p.growth_g == 4) gv = lb_roottanh(gv); else if (p.growth_g > 0) gv = lb_roottanh_g(gv, 1.0f / p.growth_g);
        acc = fmaf(gv, sd[r][kd], acc);
      }
    }
  }
  if (live) atomicAdd(p.dw + kg * p.w_sk + kd * p.w_sn + ty * p.w_sty + tx * p.w_stx, acc);
}

extern "C" int lb_conv_small_wgrad_supported(const lb_conv_geom* g) {
  if (!g || g->mode != 0) return 0;
  return ((long long)g->kh * g->kw * g->in_c * g->out_c <= 256 && g->out_c <= 32) ? 1 : 0;
}
// geom as lb_conv_wgrad: in_* = gathered operand, out_* = dense operand; dw in the master layout (+=, caller zeroes);
// growth_gathered > 0 applies RootTanh to the gathered operand on load (the layer's pre-activation)
extern "C" int lb_conv_small_wgrad(const float* gathered, const float* dense, float* dw, const lb_conv_geom* g, int growth_gathered,
                                   lb_stream_t s) {
  LB_REQUIRE(gathered && dense && dw && g && growth_gathered >= 0);
  if (!lb_conv_small_wgrad_supported(g)) return LB_EUNSUPPORTED;
  SmallWgP p;
  p.gath = gathered; p.dense = dense; p.dw = dw;
  p.g_h = g->in_h; p.g_w = g->in_w; p.g_c = g->in_c; p.d_h = g->out_h; p.d_w = g->out_w; p.d_c = g->out_c;
  p.kh = g->kh; p.kw = g->kw; p.stride = g->stride; p.pad = g->pad; p.ld_g = g->ld_in; p.ld_d = g->ld_out;
  p.w_sk = g->w_sk; p.w_sn = g->w_sn; p.w_sty = g->w_sty; p.w_stx = g->w_stx;
  p.pixels = (long long)g->batch * g->out_h * g->out_w;
  p.n_elems = g->kh * g->kw * g->in_c * g->out_c;
  p.growth_g = growth_gathered;
  long long ctas = LB_SMS * 8;
  long long chunk = (p.pixels + ctas - 1) / ctas;
  chunk = (chunk + WG_TILE - 1) / WG_TILE * WG_TILE;
  ctas = (p.pixels + chunk - 1) / chunk;
  p.chunk = (int)chunk;
  k_conv_small_wgrad<<<(unsigned)ctas, 256, 0, lb_s(s)>>>(p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
