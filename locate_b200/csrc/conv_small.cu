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
  int cat;            // rows of `out` are [in (in_c, copied) | conv (out_c)]  (CatModule, merge.py:10-16)
  int out_pitch;      // floats per staged output row: out_c (+ in_c when cat)
  int tiles;
  long long pixels;   // batch*out_h*out_w
  LbFastDiv d_grp, d_w, d_h, d_ic, d_oc, d_pitch, d_pitch4;
};

__device__ __forceinline__ float small_act(float v, int growth) {
  return growth == 4 ? lb_roottanh(v) : (growth > 0 ? lb_roottanh_g(v, 1.0f / growth) : v);
}
__device__ __forceinline__ float small_dact(float v, int growth) {
  return growth == 4 ? lb_roottanh_grad(v) : lb_roottanh_grad_g(v, 1.0f / growth);
}

// Tile copies between global rows (`cols` floats every `ld`) and shared rows (`cols` floats every `pitch`, starting at
// column `col0` of the shared row).  Every thread keeps several independent accesses in flight: with one load per thread
// per round trip these kernels sat at ~1 TB/s waiting on the long scoreboard (ncu, profiles/r1_small_kernels.txt).
__device__ __forceinline__ bool tile_is_flat(const float* g, int ld, int pitch, int col0, int rows, int cols) {
  return ld == cols && pitch == cols && col0 == 0 && !((rows * cols) & 3) && lb_aligned16(g);
}
template <typename F>
__device__ __forceinline__ void tile_load(float* sdst, int pitch, int col0, const float* gsrc, int ld, int rows, int cols,
                                          const LbFastDiv& dc, F f) {
  if (tile_is_flat(gsrc, ld, pitch, col0, rows, cols)) {          // one contiguous block on both sides
    const int n4 = (rows * cols) >> 2;
    for (int base = threadIdx.x; base < n4; base += 4 * SMALL_THREADS) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base + j * SMALL_THREADS < n4) v[j] = lb_ld4(gsrc + 4 * (size_t)(base + j * SMALL_THREADS));
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base + j * SMALL_THREADS < n4) {
          float4 r = v[j];
          r.x = f(r.x); r.y = f(r.y); r.z = f(r.z); r.w = f(r.w);
          *reinterpret_cast<float4*>(sdst + 4 * (size_t)(base + j * SMALL_THREADS)) = r;
        }
    }
    return;
  }
  const int n = rows * cols;
  for (int base = threadIdx.x; base < n; base += 4 * SMALL_THREADS) {
    float v[4];
    int r[4], c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = base + j * SMALL_THREADS;
      if (e < n) { lb_fast_divmod(dc, e, r[j], c[j]); v[j] = __ldg(gsrc + (size_t)r[j] * ld + c[j]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (base + j * SMALL_THREADS < n) sdst[r[j] * pitch + col0 + c[j]] = f(v[j]);
  }
}
__device__ __forceinline__ void tile_store(float* gdst, int ld, const float* ssrc, int rows, int cols, const LbFastDiv& dc,
                                           const LbFastDiv& dc4) {   // shared pitch == cols
  if (ld == cols && !((rows * cols) & 3) && lb_aligned16(gdst)) {
    const int n4 = (rows * cols) >> 2;
    for (int e = threadIdx.x; e < n4; e += SMALL_THREADS)
      lb_st4(gdst + 4 * (size_t)e, *reinterpret_cast<const float4*>(ssrc + 4 * (size_t)e));
    return;
  }
  if (!(cols & 3) && !(ld & 3) && lb_aligned16(gdst)) {
    const int n4 = rows * (cols >> 2);
    for (int e = threadIdx.x; e < n4; e += SMALL_THREADS) {
      int r, c;
      lb_fast_divmod(dc4, e, r, c);
      lb_st4(gdst + (size_t)r * ld + 4 * c, *reinterpret_cast<const float4*>(ssrc + 4 * (size_t)e));
    }
    return;
  }
  const int n = rows * cols;
  for (int e = threadIdx.x; e < n; e += SMALL_THREADS) {
    int r, c;
    lb_fast_divmod(dc, e, r, c);
    gdst[(size_t)r * ld + c] = ssrc[e];
  }
}


__global__ void __launch_bounds__(SMALL_THREADS, 4) k_conv_small(const SmallP p) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  const int taps = p.kh * p.kw, npad = p.ngroups * 4;
  float* ws = sm;                                           // [tap][k][npad]
  float* s_out = ws + taps * p.in_c * npad;                 // [SMALL_TP][out_pitch]  (first holds xpre, then the result)
  float* s_in = s_out + ((SMALL_TP * p.out_pitch + 3) & ~3);  // [SMALL_TP][in_c] (stage_in only)
  for (int i = threadIdx.x; i < taps * p.in_c * npad; i += blockDim.x) {
    const int n = i % npad, k = (i / npad) % p.in_c, tap = i / (npad * p.in_c);
    ws[i] = n < p.out_c ? __ldg(p.w + k * p.w_sk + n * p.w_sn + (tap / p.kw) * p.w_sty + (tap % p.kw) * p.w_stx) : 0.0f;
  }
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
  const int oc0 = p.cat ? p.in_c : 0;                       // first conv column of a staged output row
  const int gi = p.growth_in, go = p.growth_out;
  const int sh = p.stride == 2 ? 1 : 0;
  auto act_in = [gi](float v) { return small_act(v, gi); };
  auto ident = [](float v) { return v; };

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long pix0 = (long long)tile * SMALL_TP;
    const int cnt = (int)min((long long)SMALL_TP, p.pixels - pix0);
    __syncthreads();                                        // previous tile's copy-out has left s_out / s_in
    // ---- stage: activated input rows (1x1), the copied input columns of a concat, RootTanh pre-activations
    if (p.stage_in) tile_load(s_in, p.in_c, 0, p.in + pix0 * p.ld_in, p.ld_in, cnt, p.in_c, p.d_ic, act_in);
    if (p.cat) tile_load(s_out, p.out_pitch, 0, p.in + pix0 * p.ld_in, p.ld_in, cnt, p.in_c, p.d_ic, ident);
    if (go > 0) tile_load(s_out, p.out_pitch, oc0, p.xpre + pix0 * p.ld_xpre, p.ld_xpre, cnt, p.out_c, p.d_oc, ident);
    __syncthreads();
    // ---- compute: one item = one pixel x 4 consecutive output channels
    for (int it = threadIdx.x; it < cnt * p.ngroups; it += blockDim.x) {
      int pl, grp;
      lb_fast_divmod(p.d_grp, it, pl, grp);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.stage_in) {
        const float* a = s_in + pl * p.in_c;
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
          else { const int v = oy + p.pad - ty; if (v < 0 || (v & (p.stride - 1))) continue; iy = v >> sh; }   // stride 1 | 2
          if (iy < 0 || iy >= p.in_h) continue;
          for (int tx = 0; tx < p.kw; ++tx) {
            int ix;
            if (p.mode == 0) ix = ox * p.stride - p.pad + tx;
            else { const int v = ox + p.pad - tx; if (v < 0 || (v & (p.stride - 1))) continue; ix = v >> sh; }
            if (ix < 0 || ix >= p.in_w) continue;
            const float* src = p.in + ((size_t)(b * p.in_h + iy) * p.in_w + ix) * p.ld_in;
            const float* wt = ws + (size_t)(ty * p.kw + tx) * p.in_c * npad + grp * 4;
            for (int k = 0; k < p.in_c; ++k) {
              const float av = small_act(__ldg(src + k), gi);
              const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)k * npad);
              acc[0] = fmaf(av, w4.x, acc[0]); acc[1] = fmaf(av, w4.y, acc[1]);
              acc[2] = fmaf(av, w4.z, acc[2]); acc[3] = fmaf(av, w4.w, acc[3]);
            }
          }
        }
      }
      float* o = s_out + pl * p.out_pitch + oc0 + grp * 4;
      const int nv = min(4, p.out_c - grp * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nv) {
          float r = acc[j] * alpha;
          if (p.bias) r += __ldg(p.bias + grp * 4 + j);
          if (go > 0) r *= small_dact(o[j], go);
          o[j] = r;
        }
      }
    }
    __syncthreads();
    // ---- copy out: full staged rows, consecutive threads on consecutive addresses
    tile_store(p.out + pix0 * p.ld_out, p.ld_out, s_out, cnt, p.out_pitch, p.d_pitch, p.d_pitch4);
  }
}

static size_t small_smem_bytes(const lb_conv_geom* g, int cat) {
  const int ngroups = (g->out_c + 3) / 4;
  const bool stage = g->kh == 1 && g->kw == 1 && g->stride == 1 && g->pad == 0;
  const size_t out_floats = ((size_t)SMALL_TP * (g->out_c + (cat ? g->in_c : 0)) + 3) & ~(size_t)3;
  return ((size_t)g->kh * g->kw * g->in_c * ngroups * 4 + out_floats + (stage ? (size_t)SMALL_TP * g->in_c : 0)) * sizeof(float);
}

extern "C" int lb_conv_small_supported(const lb_conv_geom* g) {
  if (!g) return 0;
  if (g->stride != 1 && g->stride != 2) return 0;
  const long long wf = (long long)g->kh * g->kw * g->in_c * ((g->out_c + 3) / 4 * 4);
  // one side of the layer is tiny and the whole weight fits in shared memory (a 1024 -> 1 head on a 1x1 map is a dot
  // product per sample: that stays a GEMM)
  if (wf > SMALL_MAX_W || g->in_c > 64 || g->out_c > 64) return 0;
  if (!(g->in_c <= 4 || g->out_c <= 4 || g->in_c * g->out_c <= 128)) return 0;
  const long long pixels = (long long)g->batch * g->out_h * g->out_w;
  if (pixels >= (1ll << 31) - (1ll << 24)) return 0;
  return small_smem_bytes(g, 0) <= 48 * 1024 ? 1 : 0;
}

// cat_input != 0: `out` points at the START of rows of in_c + out_c floats (row stride g->ld_out); the kernel writes the
// input copy and the conv result of a CatModule in one pass (1x1 stride-1 layers only, no fused activations).
extern "C" int lb_conv_small(const float* in, const float* w, const float* alpha, const float* bias, float* out,
                             const lb_conv_geom* g, int growth_in, const float* xpre, int ld_xpre, int growth_out,
                             int cat_input, lb_stream_t s) {
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
  p.cat = cat_input ? 1 : 0;
  if (p.cat) LB_REQUIRE(p.stage_in && growth_in == 0 && growth_out == 0 && g->ld_out >= g->in_c + g->out_c);
  if (small_smem_bytes(g, p.cat) > 48 * 1024) return LB_EUNSUPPORTED;
  p.out_pitch = g->out_c + (p.cat ? g->in_c : 0);
  p.d_grp = lb_make_fastdiv(p.ngroups); p.d_w = lb_make_fastdiv(g->out_w); p.d_h = lb_make_fastdiv(g->out_h);
  p.d_ic = lb_make_fastdiv(g->in_c); p.d_oc = lb_make_fastdiv(g->out_c);
  p.d_pitch = lb_make_fastdiv(p.out_pitch); p.d_pitch4 = lb_make_fastdiv(p.out_pitch >= 4 ? p.out_pitch / 4 : 1);
  const int grid = p.tiles < LB_SMS * 8 ? p.tiles : LB_SMS * 8;
  k_conv_small<<<grid, SMALL_THREADS, small_smem_bytes(g, p.cat), lb_s(s)>>>(p);
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
__global__ void __launch_bounds__(256, 4) k_conv_small_wgrad(const SmallWgP p) {
  __shared__ __align__(16) float sd[WG_TP * 32];                // dense rows, pitch d_c (<= 32)
  __shared__ __align__(16) float sg[WG_TP * WG_MAX_ROWS];       // im2col rows of the gathered operand, pitch rows (<= 80)
  __shared__ int s_iy[WG_TP], s_ix[WG_TP], s_b[WG_TP];
  const int e = threadIdx.x;
  const bool live = e < p.n_elems;
  int kd = 0, row = 0;
  if (live) lb_fast_divmod(p.f_dc, e, row, kd);      // e = row * d_c + kd, row = (ty*kw + tx)*g_c + kg
  const bool pointwise = p.kh == 1 && p.kw == 1 && p.stride == 1 && p.pad == 0;   // gathered pixel == dense pixel
  const int gg = p.growth_g;
  auto act_g = [gg](float v) { return small_act(v, gg); };
  auto ident = [](float v) { return v; };
  float acc = 0.0f;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long base = (long long)tile * WG_TP;
    const int cnt = (int)min((long long)WG_TP, p.pixels - base);
    __syncthreads();
    tile_load(sd, p.d_c, 0, p.dense + base * p.ld_d, p.ld_d, cnt, p.d_c, p.f_dc, ident);
    if (pointwise) {
      tile_load(sg, p.rows, 0, p.gath + base * p.ld_g, p.ld_g, cnt, p.g_c, p.f_gc, act_g);
    } else {
      if (threadIdx.x < cnt) {
        int t, ox, oy, b;
        lb_fast_divmod(p.f_w, (int)(base + threadIdx.x), t, ox);
        lb_fast_divmod(p.f_h, t, b, oy);
        s_b[threadIdx.x] = b;
        s_iy[threadIdx.x] = oy * p.stride - p.pad;
        s_ix[threadIdx.x] = ox * p.stride - p.pad;
      }
      __syncthreads();
      const int n = cnt * p.rows;
      for (int i0 = threadIdx.x; i0 < n; i0 += 2 * 256) {
        float v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {                // 2 independent gathers in flight per thread
          const int i = i0 + j * 256;
          v[j] = 0.0f;
          if (i < n) {
            int r, rw, tap, kg, ty, tx;
            lb_fast_divmod(p.f_rows, i, r, rw);
            lb_fast_divmod(p.f_gc, rw, tap, kg);
            lb_fast_divmod(p.f_kw, tap, ty, tx);
            const int iy = s_iy[r] + ty, ix = s_ix[r] + tx;
            if (iy >= 0 && iy < p.g_h && ix >= 0 && ix < p.g_w)
              v[j] = __ldg(p.gath + ((size_t)(s_b[r] * p.g_h + iy) * p.g_w + ix) * p.ld_g + kg);
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (i0 + j * 256 < n) sg[i0 + j * 256] = small_act(v[j], gg);     // act(0) = 0: padding stays zero
      }
    }
    __syncthreads();
    if (live) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* gp = sg + row;
      const float* dp = sd + kd;
      int r = 0;
      for (; r + 4 <= cnt; r += 4) {
        a0 = fmaf(gp[(r + 0) * p.rows], dp[(r + 0) * p.d_c], a0);
        a1 = fmaf(gp[(r + 1) * p.rows], dp[(r + 1) * p.d_c], a1);
        a2 = fmaf(gp[(r + 2) * p.rows], dp[(r + 2) * p.d_c], a2);
        a3 = fmaf(gp[(r + 3) * p.rows], dp[(r + 3) * p.d_c], a3);
      }
      for (; r < cnt; ++r) a0 = fmaf(gp[r * p.rows], dp[r * p.d_c], a0);
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
