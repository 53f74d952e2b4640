// Direct fp32 kernels for the layers whose channel counts are too small for a 128x64 tensor-core tile to make
// sense: the discriminator stem (5x5/s2 3->3, 1x1 3->32, 1x1 3->29) and the generator's final 1x1 48->3, in all
// three directions.  They are HBM-bound on the wide side of the layer (AI <= 16 flop/B), so the job is to touch
// every activation once, with full-line accesses, the tiny weight held in shared memory and the neighbouring
// elementwise ops fused: RootTanh on the input (forward), multiplication by RootTanh'(x) on the output (input gradient).
//
//   lb_conv_small:       out[p][n] = alpha * sum_{tap,k} act(in[p@tap][k]) * W(tap,k,n) (+bias[n])   [* act'(xpre[p][n])]
//   lb_conv_small_wgrad: dw(tap,kg,kd) += sum_p act(gathered[p@tap][kg]) * dense[p][kd]
// Geometry and weight addressing are those of lb_conv_gemm / lb_conv_wgrad (include/locate_b200.h).
//
// Both kernels walk tiles of consecutive pixels and move every activation between global and shared memory with
// consecutive threads on consecutive addresses (a channels-last row of 3 or 29 floats per thread would waste 3/4 of
// every 32-byte sector request); the arithmetic then reads shared memory only.
#include "common.cuh"

#define SMALL_MAX_W 4096      // floats of weight held in shared memory (taps*K*N_pad)
#define SMALL_TP 128          // pixels per tile (forward / input gradient)
#define SMALL_THREADS 256

struct SmallP {
  const float* in; const float* w; const float* alpha; const float* bias; const float* xpre; float* out;
  int batch, in_h, in_w, in_c, out_h, out_w, out_c;
  int kh, kw, stride, pad, mode, ld_in, ld_out, ld_xpre;
  long long w_sk, w_sn, w_sty, w_stx;
  int growth_in;      // > 0: RootTanh(growth) applied to every input element on load
  int growth_out;     // > 0: result multiplied by RootTanh'(xpre[p][n])
  int ngroups;        // ceil(out_c / 4)
  int stage_in;       // 1x1 stride 1: the input tile is staged (activated once) in shared memory
  int vec_out;        // 16-byte aligned output rows
  int tiles;
  long long pixels;   // batch*out_h*out_w
  LbFastDiv d_grp, d_w, d_h, d_oc, d_ic;
};

__device__ __forceinline__ float small_act(float v, int growth) {
  return growth == 4 ? lb_roottanh(v) : (growth > 0 ? lb_roottanh_g(v, 1.0f / growth) : v);
}
__device__ __forceinline__ float small_dact(float v, int growth) {
  return growth == 4 ? lb_roottanh_grad(v) : lb_roottanh_grad_g(v, 1.0f / growth);
}

__global__ void __launch_bounds__(SMALL_THREADS) k_conv_small(const SmallP p) {
  extern __shared__ float sm[];
  const int taps = p.kh * p.kw, npad = p.ngroups * 4;
  float* ws = sm;                                           // [tap][k][npad]
  float* s_out = ws + taps * p.in_c * npad;                 // [SMALL_TP][out_c]   (first holds xpre, then the result)
  float* s_in = s_out + SMALL_TP * p.out_c;                 // [SMALL_TP][in_c + 1] (stage_in only)
  const int in_pitch = p.in_c + 1;
  for (int i = threadIdx.x; i < taps * p.in_c * npad; i += blockDim.x) {
    const int n = i % npad, k = (i / npad) % p.in_c, tap = i / (npad * p.in_c);
    ws[i] = n < p.out_c ? __ldg(p.w + k * p.w_sk + n * p.w_sn + (tap / p.kw) * p.w_sty + (tap % p.kw) * p.w_stx) : 0.0f;
  }
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long pix0 = (long long)tile * SMALL_TP;
    const int cnt = (int)min((long long)SMALL_TP, p.pixels - pix0);
    __syncthreads();                                        // previous tile's copy-out has left s_out / s_in
    // ---- stage: activated input rows (1x1), RootTanh pre-activations of the output
    if (p.stage_in) {
      const float* src = p.in + pix0 * p.ld_in;
      for (int e = threadIdx.x; e < cnt * p.in_c; e += blockDim.x) {
        int r, c;
        lb_fast_divmod(p.d_ic, e, r, c);
        s_in[r * in_pitch + c] = small_act(__ldg(src + (size_t)r * p.ld_in + c), p.growth_in);
      }
    }
    if (p.growth_out > 0) {
      const float* src = p.xpre + pix0 * p.ld_xpre;
      for (int e = threadIdx.x; e < cnt * p.out_c; e += blockDim.x) {
        int r, c;
        lb_fast_divmod(p.d_oc, e, r, c);
        s_out[e] = __ldg(src + (size_t)r * p.ld_xpre + c);
      }
    }
    __syncthreads();
    // ---- compute: one item = one pixel x 4 consecutive output channels
    for (int it = threadIdx.x; it < cnt * p.ngroups; it += blockDim.x) {
      int pl, grp;
      lb_fast_divmod(p.d_grp, it, pl, grp);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.stage_in) {
        const float* a = s_in + pl * in_pitch;
        const float* wt = ws + grp * 4;
        for (int k = 0; k < p.in_c; ++k) {
          const float av = a[k];
          const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)k * npad);
          acc[0] = fmaf(av, w4.x, acc[0]); acc[1] = fmaf(av, w4.y, acc[1]);
          acc[2] = fmaf(av, w4.z, acc[2]); acc[3] = fmaf(av, w4.w, acc[3]);
        }
      } else {
        int t, ox, oy, b;
        lb_fast_divmod(p.d_w, (int)(pix0 + pl), t, ox);     // pixels < 2^31 (checked by the host)
        lb_fast_divmod(p.d_h, t, b, oy);
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
            for (int k = 0; k < p.in_c; ++k) {
              const float av = small_act(__ldg(src + k), p.growth_in);
              const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)k * npad);
              acc[0] = fmaf(av, w4.x, acc[0]); acc[1] = fmaf(av, w4.y, acc[1]);
              acc[2] = fmaf(av, w4.z, acc[2]); acc[3] = fmaf(av, w4.w, acc[3]);
            }
          }
        }
      }
      float* o = s_out + pl * p.out_c + grp * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = grp * 4 + j;
        if (n >= p.out_c) break;
        float r = acc[j] * alpha + (p.bias ? __ldg(p.bias + n) : 0.0f);
        if (p.growth_out > 0) r *= small_dact(o[j], p.growth_out);
        o[j] = r;
      }
    }
    __syncthreads();
    // ---- copy out: consecutive threads on consecutive addresses of the output rows
    float* dst = p.out + pix0 * p.ld_out;
    if (p.vec_out) {
      const int oc4 = p.out_c >> 2;
      for (int e = threadIdx.x; e < cnt * oc4; e += blockDim.x) {
        const int r = e / oc4, c = (e - r * oc4) * 4;
        lb_st4(dst + (size_t)r * p.ld_out + c, *reinterpret_cast<const float4*>(s_out + r * p.out_c + c));
      }
    } else {
      for (int e = threadIdx.x; e < cnt * p.out_c; e += blockDim.x) {
        int r, c;
        lb_fast_divmod(p.d_oc, e, r, c);
        dst[(size_t)r * p.ld_out + c] = s_out[e];
      }
    }
  }
}

static size_t small_smem_bytes(const lb_conv_geom* g) {
  const int ngroups = (g->out_c + 3) / 4;
  const bool stage = g->kh == 1 && g->kw == 1 && g->stride == 1 && g->pad == 0;
  return ((size_t)g->kh * g->kw * g->in_c * ngroups * 4 + (size_t)SMALL_TP * g->out_c +
          (stage ? (size_t)SMALL_TP * (g->in_c + 1) : 0)) * sizeof(float);
}

extern "C" int lb_conv_small_supported(const lb_conv_geom* g) {
  if (!g) return 0;
  const long long wf = (long long)g->kh * g->kw * g->in_c * ((g->out_c + 3) / 4 * 4);
  // one side of the layer is tiny and the whole weight fits in shared memory (a 1024 -> 1 head on a 1x1 map is a dot
  // product per sample: that stays a GEMM)
  if (wf > SMALL_MAX_W || g->in_c > 64 || g->out_c > 64) return 0;
  if (!(g->in_c <= 4 || g->out_c <= 4 || g->in_c * g->out_c <= 128)) return 0;
  const long long pixels = (long long)g->batch * g->out_h * g->out_w;
  if (pixels >= (1ll << 31) - (1ll << 24)) return 0;
  return small_smem_bytes(g) <= 48 * 1024 ? 1 : 0;
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
  p.tiles = (int)((p.pixels + SMALL_TP - 1) / SMALL_TP);
  p.stage_in = (g->kh == 1 && g->kw == 1 && g->stride == 1 && g->pad == 0) ? 1 : 0;
  p.vec_out = (!(g->out_c & 3) && !(g->ld_out & 3) && lb_aligned16(out)) ? 1 : 0;
  p.d_grp = lb_make_fastdiv(p.ngroups); p.d_w = lb_make_fastdiv(g->out_w); p.d_h = lb_make_fastdiv(g->out_h);
  p.d_oc = lb_make_fastdiv(g->out_c); p.d_ic = lb_make_fastdiv(g->in_c);
  const int grid = p.tiles < LB_SMS * 8 ? p.tiles : LB_SMS * 8;
  k_conv_small<<<grid, SMALL_THREADS, small_smem_bytes(g), lb_s(s)>>>(p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// ---- weight gradient: tiny output, reduction over every pixel ------------------------------------------------
// A CTA walks tiles of WG_TP dense pixels.  Per tile the dense rows and the im2col rows of the gathered operand
// (taps x g_c values per pixel, activated on the way in) are staged in shared memory with coalesced loads; thread e
// then owns one weight element (tap, kg, kd) and accumulates over the tile's pixels from shared memory only.
struct SmallWgP {
  const float* gath; const float* dense; float* dw;
  int g_h, g_w, g_c, d_h, d_w, d_c, kh, kw, stride, pad, ld_g, ld_d;
  long long w_sk, w_sn, w_sty, w_stx;
  long long pixels; int tiles, n_elems, rows, growth_g;
  LbFastDiv f_rows, f_dc, f_gc, f_kw, f_w, f_h;
};
#define WG_TP 64
#define WG_MAX_ROWS 80
__global__ void __launch_bounds__(256) k_conv_small_wgrad(const SmallWgP p) {
  __shared__ float sd[WG_TP][33];                    // dense rows (d_c <= 32)
  __shared__ float sg[WG_TP][WG_MAX_ROWS + 1];       // im2col rows of the gathered operand (taps*g_c <= 80)
  __shared__ int s_iy[WG_TP], s_ix[WG_TP], s_b[WG_TP];
  const int e = threadIdx.x;
  const bool live = e < p.n_elems;
  int kd = 0, row = 0;
  if (live) lb_fast_divmod(p.f_dc, e, row, kd);      // e = row * d_c + kd, row = (ty*kw + tx)*g_c + kg
  float acc = 0.0f;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long base = (long long)tile * WG_TP;
    const int cnt = (int)min((long long)WG_TP, p.pixels - base);
    __syncthreads();
    if (threadIdx.x < cnt) {
      int t, ox, oy, b;
      lb_fast_divmod(p.f_w, (int)(base + threadIdx.x), t, ox);
      lb_fast_divmod(p.f_h, t, b, oy);
      s_b[threadIdx.x] = b;
      s_iy[threadIdx.x] = oy * p.stride - p.pad;
      s_ix[threadIdx.x] = ox * p.stride - p.pad;
    }
    for (int i = threadIdx.x; i < cnt * p.d_c; i += blockDim.x) {
      int r, c;
      lb_fast_divmod(p.f_dc, i, r, c);
      sd[r][c] = __ldg(p.dense + (size_t)(base + r) * p.ld_d + c);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * p.rows; i += blockDim.x) {
      int r, rw, tap, kg, ty, tx;
      lb_fast_divmod(p.f_rows, i, r, rw);
      lb_fast_divmod(p.f_gc, rw, tap, kg);
      lb_fast_divmod(p.f_kw, tap, ty, tx);
      const int iy = s_iy[r] + ty, ix = s_ix[r] + tx;
      float gv = 0.0f;
      if (iy >= 0 && iy < p.g_h && ix >= 0 && ix < p.g_w)
        gv = small_act(__ldg(p.gath + ((size_t)(s_b[r] * p.g_h + iy) * p.g_w + ix) * p.ld_g + kg), p.growth_g);
      sg[r][rw] = gv;
    }
    __syncthreads();
    if (live) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int r = 0;
      for (; r + 4 <= cnt; r += 4) {
        a0 = fmaf(sg[r][row], sd[r][kd], a0);
        a1 = fmaf(sg[r + 1][row], sd[r + 1][kd], a1);
        a2 = fmaf(sg[r + 2][row], sd[r + 2][kd], a2);
        a3 = fmaf(sg[r + 3][row], sd[r + 3][kd], a3);
      }
      for (; r < cnt; ++r) a0 = fmaf(sg[r][row], sd[r][kd], a0);
      acc += (a0 + a1) + (a2 + a3);
    }
  }
  if (live) {
    int tap, kg, ty, tx;
    lb_fast_divmod(p.f_gc, row, tap, kg);
    lb_fast_divmod(p.f_kw, tap, ty, tx);
    atomicAdd(p.dw + kg * p.w_sk + kd * p.w_sn + ty * p.w_sty + tx * p.w_stx, acc);
  }
}

extern "C" int lb_conv_small_wgrad_supported(const lb_conv_geom* g) {
  if (!g || g->mode != 0) return 0;
  const long long rows = (long long)g->kh * g->kw * g->in_c;
  const long long pixels = (long long)g->batch * g->out_h * g->out_w;
  return (rows * g->out_c <= 256 && rows <= WG_MAX_ROWS && g->out_c <= 32 && pixels < (1ll << 31) - (1ll << 24)) ? 1 : 0;
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
  p.rows = g->kh * g->kw * g->in_c;
  p.n_elems = p.rows * g->out_c;
  p.growth_g = growth_gathered;
  p.tiles = (int)((p.pixels + WG_TP - 1) / WG_TP);
  p.f_rows = lb_make_fastdiv(p.rows); p.f_dc = lb_make_fastdiv(g->out_c); p.f_gc = lb_make_fastdiv(g->in_c);
  p.f_kw = lb_make_fastdiv(g->kw); p.f_w = lb_make_fastdiv(g->out_w); p.f_h = lb_make_fastdiv(g->out_h);
  const int grid = p.tiles < LB_SMS * 6 ? p.tiles : LB_SMS * 6;
  k_conv_small_wgrad<<<grid, 256, 0, lb_s(s)>>>(p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
