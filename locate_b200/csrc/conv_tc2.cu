// Persistent tcgen05 / TMEM / TMA implicit-GEMM for the convolution family with a fused, TMA-stored epilogue.
//
//   acc[pixel][n] = alpha * sum_taps sum_k A_tap[pixel][k] * Wp[tap][n][k]  (+ bias[n])
//   if aux:    acc *= RootTanh'(aux[pixel][n])           (the activation backward that follows a dgrad GEMM; aux fp32 or bf16)
//              or, LB_EX_AUX_IS_FACTOR: acc *= aux[pixel][n]   (aux already holds RootTanh' -- see below)
//   if out32:  out32[pixel][n]  = acc                     (fp32)
//   if out16:  out16[pixel][n]  = bf16(acc)               (bf16 activation storage: pre-activation / gradient)
//              or, LB_EX_OUT16_IS_DACT: bf16(RootTanh'(acc)): the forward pass stores the activation's DERIVATIVE instead
//              of its argument, so the backward epilogue multiplies by a loaded factor instead of evaluating it -- the
//              transcendental work (4 MUFU ops per element, the limiter of the C <= 96 layers' epilogues) is done once,
//              next to RootTanh itself with which it shares every intermediate
//   if out16a: out16a[pixel][n] = bf16(RootTanh(acc'))    (the next GEMM's operand, produced in place; acc' = the value
//                                                          stored by out16 when a pre-activation is written, so that the
//                                                          backward's RootTanh'(out16) belongs to exactly this value)
//
// Operand staging and tap handling are those of conv_tc.cu (per-tap dense TMA boxes of parity views, zero fill =
// padding).  What is different:
//  * one CTA per SM walks a static list of (pixel tile, channel tile, phase) work items;
//  * the TMEM accumulator is double buffered: the MMA warp fills stage (t+1)&1 while the epilogue drains t&1,
//    and the TMA producer runs ahead across tile boundaries, so per-tile fixed latencies overlap with math;
//  * the epilogue (16 warps: 4 per TMEM lane quarter, each taking every 4th 32-column slab; 8 warps when the staging
//    buffers of 16 would not fit) moves
//    tcgen05.ld -> registers -> 128B-swizzled shared slab -> cp.async.bulk.tensor store.  Every global write is a
//    full-line TMA store of a 4-D box of the (possibly parity-strided) output view, clipped by the tensor map at the
//    tensor edges; `aux` slabs arrive the same way through per-warp mbarriers, double buffered;
//  * the last K chunk issues only the k16 steps that hold real channels.
// Warp roles (up to 576 threads): 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2.. = epilogue.  The epilogue of the
// C <= 96 layers costs more than their MMAs (RootTanh: ~25 instructions, 4 of them MUFU, per output element against
// 2*K flops on the tensor core), and with 2 warps per scheduler it ran latency-bound at a third of its issue rate.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // bf16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kMaxViews = 4;
constexpr int kSlab = 32;                        // fp32 columns per epilogue slab (128 B per row)
constexpr int kEpiWarps = 16;                    // at most: 4 per TMEM lane quarter; launches use 8 or 16 (Tc2Params::epi_warps)
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemLimit = 227 * 1024 - 1024;      // dynamic; 1 KB left for the static barriers

struct Tc2Maps {
  CUtensorMap a[kMaxViews];
  CUtensorMap b;
  CUtensorMap o32[kMaxViews];
  CUtensorMap o16[kMaxViews];
  CUtensorMap o16a[kMaxViews];
  CUtensorMap aux[kMaxViews];
};

struct Tc2Params {
  int batch, out_c, in_c, in_h, in_w;
  int sp;                             // destination parity step (stride in mode 1, else 1)
  int sh;                             // log2(stride)
  LbFastDiv d_nt, d_tw, d_th, d_tb;     // divisors n_tiles, tiles_w, tiles_h, tiles_b
  int tile_w, tile_h, tile_b;         // box dims, product = 128
  int tiles_w, tiles_h, tiles_b, n_tiles, total_tiles;
  int block_n, acc_stride, kchunks, stages;
  int kh, kw, stride, pad, mode;
  int rows_per_tap, view_empty;
  int ebw, ebh, ebb;                  // 32-row sub-box of one epilogue warp
  int epi_warps;                      // 8 or 16
  // Epilogue warp GROUPS: a tile whose column slabs need only 4 or 8 warps (<= 64 channels) leaves the others idle, and one
  // warp's trip through a slab (accumulator wait, tcgen05.ld, ~700 dependent instructions, fence, TMA store) is a ~2.5 us
  // latency chain whatever the tile width -- with every warp visiting every tile, that chain WAS the tile period of all
  // layers up to 96 channels.  Groups take tiles round robin (tile t -> group t % epi_groups, accumulator t % acc_depth),
  // so 2 or 4 tiles drain at once.
  int epi_groups;                     // 1, 2 or 4
  int acc_shift;                      // log2 of the TMEM accumulator ring depth (2 or 4 stages)
  int has_o32, has_o16, has_o16a, has_aux;
  int aux_bf16;                       // aux slabs are bf16 (64-byte rows) instead of fp32 (128-byte rows)
  int aux_factor;                     // aux holds the factor itself (RootTanh' precomputed by the forward pass)
  int o16_dact;                       // out16 receives RootTanh'(acc) instead of acc
  uint32_t tmem_cols;
  uint32_t tab_base; int max_tp;      // per-phase tap table in shared memory: max_tp entries per phase
  // halo mode: ONE A box per (source view, 64-channel chunk) covers every tap shift of that view; each tap's operand
  // is the same shared-memory tile addressed through the UMMA descriptor (start row = shift, 8-row groups 16 rows apart).
  int halo;                           // 0 / 1
  int a_stages; uint32_t a_stage_bytes, a_base;   // ring of halo tiles (bytes; a_base relative to the 1 KB aligned base)
  int n_views;
  uint32_t htab_base;                 // view / tap tables of the halo mode (after the tap table)
  int resident;                       // 1: the weights of one output phase stay in shared memory (res_base), stages carry A only
  int phase_inner;                    // tile order: output phase fastest after the channel tile (0 with resident weights)
  uint32_t res_base, b_tile_bytes;
  int slab;                           // columns per epilogue slab: 32, or 24 / 16 (bf16-only epilogues) when that spreads the tile over more warps
  uint32_t epi_base, epi_per_warp, off_o32, off_o16, off_o16a, off_aux, aux_bytes;   // bytes; epi_base relative to the 1 KB aligned base
  const float* alpha; const float* bias;
};

struct TileCoord { int x0, y0, b0, n0, py, px, phase; };
__device__ __forceinline__ TileCoord decode_tile(const Tc2Params& p, int tile) {
  TileCoord c;
  int t = tile, nt, tw, th, tb;
  lb_fast_divmod(p.d_nt, t, t, nt);
  if (p.phase_inner) {               // the output phases of one source region run back to back: the region is read from DRAM once
    c.phase = t & (p.sp * p.sp - 1);
    t >>= (p.sp == 2 ? 2 : 0);           // sp is 1 or 2 (sh = log2(stride) also covers mode-0 stride-2 layers, whose sp is 1)
    lb_fast_divmod(p.d_tw, t, t, tw);
    lb_fast_divmod(p.d_th, t, tb, th);
  } else {                           // resident weights: one phase's weight set at a time
    lb_fast_divmod(p.d_tw, t, t, tw);
    lb_fast_divmod(p.d_th, t, t, th);
    lb_fast_divmod(p.d_tb, t, t, tb);
    c.phase = t;
  }
  c.x0 = tw * p.tile_w; c.y0 = th * p.tile_h; c.b0 = tb * p.tile_b; c.n0 = nt * p.block_n;
  c.py = c.phase >> p.sh; c.px = c.phase & (p.sp - 1);
  return c;
}

// The taps one tile visits along one axis, and for each tap where its source box sits.  stride is 1 or 2, so every
// division is a shift.  mode 1: taps t0, t0+s, .. (t = (parity + pad) mod s); box shift d = (parity + pad - t) / s in
// the single dense view.  mode 0: every tap; (a, q) = divmod(t - pad, s): box shift a in parity view q.
struct AxisTaps { int t0, step, cnt; };
__device__ __forceinline__ AxisTaps axis_taps(const Tc2Params& p, int k, int parity) {
  AxisTaps a;
  if (p.mode == 1) {
    a.t0 = (parity + p.pad) & (p.stride - 1);
    a.step = p.stride;
    a.cnt = a.t0 < k ? (k - a.t0 + p.stride - 1) >> p.sh : 0;
  } else {
    a.t0 = 0; a.step = 1; a.cnt = k;
  }
  return a;
}
// Index range [lo, hi) of axis_taps() entries that can be live for a tile at o0 (a superset: tap_live() still decides).
// Scanning every tap costs the lone producer / MMA threads ~25 instructions per tap; a full-extent 64x1 kernel has two
// live taps out of 64 per tile.
__device__ __forceinline__ void axis_range(const Tc2Params& p, const AxisTaps& a, int parity, int o0, int tile, int in_extent,
                                           int& lo, int& hi) {
  if (p.mode == 1) {                     // d_i = D0 - i
    const int D0 = (parity + p.pad - a.t0) >> p.sh;
    lo = o0 + D0 - in_extent + 1;
    hi = o0 + D0 + tile;
  } else if (p.stride == 1) {            // d_t = t - pad
    lo = p.pad - o0 - tile + 1;
    hi = in_extent + p.pad - o0;
  } else {                               // d_t = floor((t - pad) / 2), view extent <= ceil(in_extent / 2)
    lo = p.pad + 2 * (1 - o0 - tile);
    hi = p.pad + 2 * (((in_extent + 1) >> 1) - o0);
  }
  lo = max(lo, 0);
  hi = min(hi, a.cnt);
}

// ---- fast elementwise math for the epilogue (results are rounded to bf16 or multiplied into an fp32 gradient) ----
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// RootTanh(x) = (x^2+1)^(1/4) tanh(x)
__device__ __forceinline__ float roottanh_fast(float x) {
  const float e = ex2_approx(-2.8853900817779268f * fabsf(x));      // exp(-2|x|)
  const float th = copysignf((1.0f - e) * rcp_approx(1.0f + e), x);
  const float q = fmaf(x, x, 1.0f);
  return sqrt_approx(sqrt_approx(q)) * th;
}
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// RootTanh only (no derivative wanted): tanh.approx + two square roots = 3 MUFU ops; relative error ~5e-4, below the bf16
// rounding of the stored result
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ float roottanh_only_fast(float x) {
  return sqrt_approx(sqrt_approx(fmaf(x, x, 1.0f))) * tanh_approx(x);
}
// RootTanh and RootTanh' together from shared intermediates, 3 MUFU ops (tanh, rsqrt, sqrt) -- the MUFU pipe (16 lanes
// per clock per SM) is what bounds the epilogue of the C <= 96 layers.  sech^2 = 1 - tanh^2 carries the absolute error
// of tanh.approx (~5e-4) times 2, harmless while the sech^2 term matters (|x| < ~4: relative error of RootTanh' <= 0.9 %,
// the bf16 rounding of the stored factor is 0.4 %); beyond |x| = 4.5, where it contributes < 0.5 %, the term is dropped.
__device__ __forceinline__ void roottanh_both_fast(float x, float& f, float& df) {
  const float th = tanh_approx(x);
  const float q = fmaf(x, x, 1.0f);
  const float rs = rsqrt_approx(q);                                  // q^(-1/2)
  const float q34 = rs * sqrt_approx(rs);                            // q^(-3/4)
  const float s2 = fabsf(x) > 4.5f ? 0.0f : fmaf(-th, th, 1.0f);
  f = q * q34 * th;                                                  // q^(1/4) tanh
  df = fmaf(2.0f * q, s2, x * th) * 0.5f * q34;
}
// RootTanh'(x) = (2 q sech^2 + x tanh) q^(1/4) / (2q) = (2 q sech^2 + x tanh) * 0.5 q^(-3/4)
__device__ __forceinline__ float roottanh_grad_fast(float x) {
  const float e = ex2_approx(-2.8853900817779268f * fabsf(x));
  const float r = rcp_approx(1.0f + e);
  const float th = copysignf((1.0f - e) * r, x);
  const float s2 = 4.0f * e * r * r;
  const float q = fmaf(x, x, 1.0f);
  const float rs = rsqrt_approx(q);                                  // q^(-1/2)
  const float q34 = rs * sqrt_approx(rs);                            // q^(-3/4)
  return fmaf(2.0f * q, s2, x * th) * 0.5f * q34;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(tc::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, float* v) {      // no wait: the caller issues tmem_ld_wait()
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- lean helpers for the single-thread producer / MMA loops -------------------------------------------------------
// Those loops ARE the critical path of the small-channel layers (ncu: epilogue warps wait on the accumulator barrier,
// the tensor pipe is 15 % busy, producer / MMA warps never wait -- they are executing ~100 scalar instructions per MMA).
// Everything here works on 32-bit shared-window addresses computed once, and on precomputed descriptor halves.
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ int2 lds_int2(uint32_t addr) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 lds_int4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// descriptor = {hi, lo}: lo = start >> 4 | LBO(16 B) << 16; hi = SBO >> 4 | version 1 << 14 | SWIZZLE_128B << 29
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void umma_bf16_hl(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Per-phase tap table (built once per CTA): the lone producer / MMA threads were spending ~500 dependent scalar
// instructions per tile re-deriving each tap's shift, view and weight row (ncu: both warps busy, barely ever waiting),
// which capped the small-channel layers at ~3 us per tile.  Entry e of phase ph: tapv = {dy, dx, view, weight row of the
// tap}, tape = {view extent y, x} for the liveness test.
constexpr int kMaxTapsPhase = 64;
constexpr int kHaloPitch = 16;                   // pixels per halo-tile row: a multiple of 8, so every 8-row group of every tap
                                                 // shift starts at the same swizzle phase (descriptor base offset)
// halo-mode tables (built once per CTA).  hview[phase * 4 + view] = {dxmin, dymin, taps of this view, index of its first tap};
// htap[phase * 64 + j] = {row offset of the tap's shift inside the halo tile, weight row of the tap}, taps sorted by view.
__device__ __forceinline__ void build_halo_tables(const Tc2Params& p, const int4* tapv, int4* hview, int2* htap) {
  const int phases = p.sp * p.sp;
  if ((int)threadIdx.x >= phases) return;
  const int ph = threadIdx.x;
  const int py = ph >> p.sh, px = ph & (p.sp - 1);
  const AxisTaps ay = axis_taps(p, p.kh, py), ax = axis_taps(p, p.kw, px);
  const int nt = ay.cnt * ax.cnt;
  int first = 0;
  for (int v = 0; v < kMaxViews; ++v) {
    int dxmin = 1 << 20, dymin = 1 << 20, cnt = 0;
    for (int l = 0; l < nt; ++l) {
      const int4 t = tapv[ph * p.max_tp + l];
      if (t.z != v) continue;
      dxmin = min(dxmin, t.y); dymin = min(dymin, t.x); ++cnt;
    }
    int j = first;
    for (int l = 0; l < nt; ++l) {
      const int4 t = tapv[ph * p.max_tp + l];
      if (t.z != v) continue;
      htap[ph * kMaxTapsPhase + j] = make_int2((t.x - dymin) * kHaloPitch + (t.y - dxmin), t.w);
      ++j;
    }
    hview[ph * kMaxViews + v] = make_int4(cnt ? dxmin : 0, cnt ? dymin : 0, cnt, first);
    first = j;
  }
}
// Halo-mode A descriptors: K-major SWIZZLE_128B whose 8-row groups are 16 rows (2 KB) apart and whose first row sits
// `row` rows (of 128 bytes) into a 1 KB aligned tile.  The start address is then not aligned to the 1 KB swizzle pattern;
// measured on B200: the tensor core applies the swizzle to the absolute shared-memory address (as TMA did when it wrote
// the tile), so the descriptor's base-offset field (bits [49,52)) stays 0 -- setting it to (start >> 7) & 7 gives wrong
// products (tests/test_gpu_tc.py halo_* cases).
// (Descriptors are assembled from precomputed halves in the MMA loop: desc_lo / desc_hi above.)
// Per-phase tap table (built once per CTA): entry e of phase ph: tapv = {dy, dx, view, weight row of the tap},
// tape = {view extent y, x} for the liveness test of the per-tap path.
__device__ __forceinline__ void build_tap_table(const Tc2Params& p, int4* tapv, int2* tape) {
  const int phases = p.sp * p.sp;
  for (int idx = threadIdx.x; idx < phases * p.max_tp; idx += blockDim.x) {
    const int ph = idx / p.max_tp, l = idx - ph * p.max_tp;
    const int py = ph >> p.sh, px = ph & (p.sp - 1);
    const AxisTaps ay = axis_taps(p, p.kh, py), ax = axis_taps(p, p.kw, px);
    if (l >= ay.cnt * ax.cnt) continue;
    const int iy = l / ax.cnt, ix = l - iy * ax.cnt;
    const int ty = ay.t0 + iy * ay.step, tx = ax.t0 + ix * ax.step;
    int dy, qy, dx, qx, eh, ew;
    if (p.mode == 0) {
      const int oy = ty - p.pad, ox = tx - p.pad;
      dy = oy >> p.sh; qy = oy & (p.stride - 1); eh = (p.in_h - qy + p.stride - 1) >> p.sh;
      dx = ox >> p.sh; qx = ox & (p.stride - 1); ew = (p.in_w - qx + p.stride - 1) >> p.sh;
    } else {
      dy = (py + p.pad - ty) >> p.sh; qy = 0; eh = p.in_h;
      dx = (px + p.pad - tx) >> p.sh; qx = 0; ew = p.in_w;
    }
    tapv[idx] = make_int4(dy, dx, (qy << p.sh) + qx, (ty * p.kw + tx) * p.rows_per_tap);
    tape[idx] = make_int2(eh, ew);
  }
}
__device__ __forceinline__ bool tap_live(const Tc2Params& p, const TileCoord& c, const int4& v, const int2& e) {
  return c.y0 + v.x < e.x && c.y0 + v.x + p.tile_h > 0 && c.x0 + v.y < e.y && c.x0 + v.y + p.tile_w > 0;
}

__global__ void __launch_bounds__(kThreads, 1) k_conv_tc2(const __grid_constant__ Tc2Maps maps, const Tc2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_tfull[4], bar_tempty[4], bar_aux[kEpiWarps][2];
  __shared__ __align__(8) uint64_t bar_bfull, bar_bfree;          // resident weights: loaded / no longer read
  __shared__ __align__(8) uint64_t bar_afull[4], bar_aempty[4];   // halo mode: ring of halo tiles
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = p.halo ? (int)p.b_tile_bytes : (p.resident ? kABytes : kABytes + ((b_bytes + 1023) & ~1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int4* tapv = reinterpret_cast<int4*>(smem + p.tab_base);
  int2* tape = reinterpret_cast<int2*>(smem + p.tab_base + 4 * kMaxTapsPhase * sizeof(int4));
  build_tap_table(p, tapv, tape);
  int4* hview = reinterpret_cast<int4*>(smem + p.htab_base);
  int2* htap = reinterpret_cast<int2*>(smem + p.htab_base + 4 * kMaxViews * sizeof(int4));
  if (p.halo) {
    __syncthreads();
    build_halo_tables(p, tapv, hview, htap);
  }

  if (threadIdx.x == 0) {
    for (int a = 0; a < 4; ++a) { tc::mbar_init(&bar_afull[a], 1); tc::mbar_init(&bar_aempty[a], 1); }
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 4; ++a) { tc::mbar_init(&bar_tfull[a], 1); tc::mbar_init(&bar_tempty[a], p.epi_warps / p.epi_groups); }
    for (int w = 0; w < p.epi_warps; ++w) { tc::mbar_init(&bar_aux[w][0], 1); tc::mbar_init(&bar_aux[w][1], 1); }
    tc::mbar_init(&bar_bfull, 1); tc::mbar_init(&bar_bfree, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    for (int v = 0; v < kMaxViews; ++v) tc::tma_prefetch_desc(&maps.a[v]);
    tc::tma_prefetch_desc(&maps.b);
  }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // programmatic dependent launch: the successor may be scheduled only now that this CTA owns its TMEM columns (a
  // successor CTA allocating first, then waiting for this grid, would deadlock the SM's allocator); everything above ran
  // while the predecessor was still draining, nothing below may start before it has completed
  lb_pdl_trigger();
  lb_pdl_wait();
  // shared-window addresses, computed once (the compiler otherwise re-derives them from generic pointers at every use)
  // (pinned: ptxas would rather rematerialise them -- S2UR SR_CgaCtaId + two ULEAs in front of every barrier operation of the
  // single-thread producer / MMA loops, which are the critical path of the small-channel layers)
  const uint32_t smem_a = pin_u32(tc::smem_u32(smem)), htab_a = smem_a + p.htab_base;
  const uint32_t bar_full_a = pin_u32(tc::smem_u32(&bar_full[0])), bar_empty_a = pin_u32(tc::smem_u32(&bar_empty[0]));
  const uint32_t bar_afull_a = pin_u32(tc::smem_u32(&bar_afull[0])), bar_aempty_a = pin_u32(tc::smem_u32(&bar_aempty[0]));
  const uint32_t bar_tfull_a = pin_u32(tc::smem_u32(&bar_tfull[0])), bar_tempty_a = pin_u32(tc::smem_u32(&bar_tempty[0]));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      int st = 0; uint32_t ph = 0;                        // ring position, runs on across tiles
      int res_phase = -1; uint32_t epochs = 0;            // resident weights: phase they belong to, sets loaded so far
      int ast = 0; uint32_t aph = 0;                      // halo ring position
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord c = decode_tile(p, tile);
        if (p.halo) {
          // per (view, channel chunk): one halo box, then the weight tile of every tap of that view
          const uint32_t hv_a = htab_a + (uint32_t)(c.phase * kMaxViews) * 16u, ht_a = htab_a + 4u * kMaxViews * 16u + (uint32_t)(c.phase * kMaxTapsPhase) * 8u;
          if (p.resident && c.phase != res_phase) {
            // single-phase layer whose whole weight set fits next to the halo ring: loaded once per CTA, in the order the
            // MMA loop walks it (view, channel chunk, tap) -- no weight barrier or commit per tap afterwards
            uint32_t cnt = 0;
            for (int v = 0; v < p.n_views; ++v) cnt += (uint32_t)(lds_int4(hv_a + (uint32_t)v * 16u).z * p.kchunks);
            tc::mbar_arrive_expect_tx(&bar_bfull, cnt * (uint32_t)b_bytes);
            uint32_t dst = smem_a + p.res_base;
            for (int v = 0; v < p.n_views; ++v) {
              const int4 hv = lds_int4(hv_a + (uint32_t)v * 16u);
              for (int kc = 0; kc < p.kchunks; ++kc)
                for (int j = 0; j < hv.z; ++j, dst += p.b_tile_bytes)
                  tma_load_2d_a(dst, &maps.b, tc::smem_u32(&bar_bfull), kc * kBlockK, lds_int2(ht_a + (uint32_t)(hv.w + j) * 8u).y + c.n0);
            }
            res_phase = c.phase;
            ++epochs;
          }
          for (int v = 0; v < p.n_views; ++v) {
            const int4 hv = lds_int4(hv_a + (uint32_t)v * 16u);
            if (hv.z == 0) continue;
            const int cb = ((p.view_empty >> v) & 1) ? p.batch : c.b0;
            const int cx = c.x0 + hv.x, cy = c.y0 + hv.y;
            for (int kc = 0; kc < p.kchunks; ++kc) {
              mbar_wait_a(bar_aempty_a + 8u * ast, aph ^ 1u);
              mbar_expect_tx_a(bar_afull_a + 8u * ast, p.a_stage_bytes);
              tma_load_4d_a(smem_a + p.a_base + ast * p.a_stage_bytes, &maps.a[v], bar_afull_a + 8u * ast, kc * kBlockK, cx, cy, cb);
              if (++ast == p.a_stages) { ast = 0; aph ^= 1u; }
              if (p.resident) continue;
              for (int j = 0; j < hv.z; ++j) {
                const int2 ht = lds_int2(ht_a + (uint32_t)(hv.w + j) * 8u);
                mbar_wait_a(bar_empty_a + 8u * st, ph ^ 1u);
                mbar_expect_tx_a(bar_full_a + 8u * st, (uint32_t)b_bytes);
                tma_load_2d_a(smem_a + st * stage_bytes, &maps.b, bar_full_a + 8u * st, kc * kBlockK, ht.y + c.n0);
                if (++st == p.stages) { st = 0; ph ^= 1u; }
              }
            }
          }
          continue;
        }
        const AxisTaps ay = axis_taps(p, p.kh, c.py), ax = axis_taps(p, p.kw, c.px);
        if (p.resident && c.phase != res_phase) {
          // every tap x channel chunk of this phase, once; the previous set must have been read by its last MMA
          if (epochs > 0) tc::mbar_wait(&bar_bfree, (epochs - 1) & 1u);
          tc::mbar_arrive_expect_tx(&bar_bfull, (uint32_t)(ay.cnt * ax.cnt * p.kchunks * b_bytes));
          for (int iy = 0; iy < ay.cnt; ++iy)
            for (int ix = 0; ix < ax.cnt; ++ix) {
              const int wrow = ((ay.t0 + iy * ay.step) * p.kw + ax.t0 + ix * ax.step) * p.rows_per_tap + c.n0;
              for (int kc = 0; kc < p.kchunks; ++kc)
                tc::tma_load_2d(smem + p.res_base + (uint32_t)((iy * ax.cnt + ix) * p.kchunks + kc) * p.b_tile_bytes, &maps.b,
                                &bar_bfull, kc * kBlockK, wrow);
            }
          res_phase = c.phase;
          ++epochs;
        }
        int ylo, yhi, xlo, xhi;
        axis_range(p, ay, c.py, c.y0, p.tile_h, p.in_h, ylo, yhi);
        axis_range(p, ax, c.px, c.x0, p.tile_w, p.in_w, xlo, xhi);
        const int tbase = c.phase * p.max_tp;
        bool issued = false;
        auto load_tap = [&](const int4& v) {
          const int cb = ((p.view_empty >> v.z) & 1) ? p.batch : c.b0;   // empty view: box out of range -> zeros
          const int wrow = v.w + c.n0;
          const int cx = c.x0 + v.y, cy = c.y0 + v.x;
          const uint32_t tx_bytes = (uint32_t)(p.resident ? kABytes : kABytes + b_bytes);
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait_a(bar_empty_a + 8u * st, ph ^ 1u);
            const uint32_t sa = smem_a + st * stage_bytes, bf = bar_full_a + 8u * st;
            mbar_expect_tx_a(bf, tx_bytes);
            tma_load_4d_a(sa, &maps.a[v.z], bf, kc * kBlockK, cx, cy, cb);
            if (!p.resident) tma_load_2d_a(sa + kABytes, &maps.b, bf, kc * kBlockK, wrow);
            if (++st == p.stages) { st = 0; ph ^= 1u; }
          }
        };
        for (int iy = ylo; iy < yhi; ++iy) {
          for (int ix = xlo; ix < xhi; ++ix) {
            const int e = tbase + iy * ax.cnt + ix;
            const int4 v = tapv[e];
            if (!tap_live(p, c, v, tape[e])) continue;       // a tap whose box is all padding adds nothing
            issued = true;
            load_tap(v);
          }
        }
        if (!issued) load_tap(tapv[tbase + ay.cnt * ax.cnt - 1]);   // every tap is padding: one of them still defines the zeros
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_bf16(kBlockM, p.block_n, 0, 0);
      const int k16_last = (p.in_c - (p.kchunks - 1) * kBlockK + 15) / 16;
      int st = 0; uint32_t ph = 0;
      int lt = 0;
      int res_phase = -1; uint32_t epochs = 0;
      int ast = 0; uint32_t aph = 0;                      // halo ring position
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const TileCoord c = decode_tile(p, tile);
        const AxisTaps ay = axis_taps(p, p.kh, c.py), ax = axis_taps(p, p.kw, c.px);
        if (p.resident && c.phase != res_phase) {
          tc::mbar_wait(&bar_bfull, epochs & 1u);
          tc::tc_fence_after();
          res_phase = c.phase;
          ++epochs;
        }
        const int acc = lt & ((1 << p.acc_shift) - 1);
        mbar_wait_a(bar_tempty_a + 8u * acc, (((uint32_t)lt >> p.acc_shift) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.acc_stride);
        if (p.halo) {
          uint32_t accum = 0;
          const uint32_t hv_a = htab_a + (uint32_t)(c.phase * kMaxViews) * 16u, ht_a = htab_a + 4u * kMaxViews * 16u + (uint32_t)(c.phase * kMaxTapsPhase) * 8u;
          const uint32_t a_hi = desc_hi(kHaloPitch * 128u), b_hi = desc_hi(1024u);
          uint32_t b_res = desc_lo(smem_a + p.res_base);                       // resident weights: walked in load order
          for (int v = 0; v < p.n_views; ++v) {
            const int4 hv = lds_int4(hv_a + (uint32_t)v * 16u);
            if (hv.z == 0) continue;
            for (int kc = 0; kc < p.kchunks; ++kc) {
              mbar_wait_a(bar_afull_a + 8u * ast, aph);
              tc::tc_fence_after();
              const uint32_t a_lo0 = desc_lo(smem_a + p.a_base + ast * p.a_stage_bytes);
              const bool full_chunk = kc != p.kchunks - 1 || k16_last == kBlockK / 16;
              for (int j = 0; j < hv.z; ++j) {
                const int2 ht = lds_int2(ht_a + (uint32_t)(hv.w + j) * 8u);
                if (!p.resident) {
                  mbar_wait_a(bar_full_a + 8u * st, ph);
                  tc::tc_fence_after();
                }
                const uint32_t a_lo = a_lo0 + (uint32_t)ht.x * 8u;                  // + row * 128 B
                const uint32_t b_lo = p.resident ? b_res : desc_lo(smem_a + st * stage_bytes);
                b_res += p.b_tile_bytes >> 4;
                if (full_chunk) {
                  umma_bf16_hl(tmem_d, a_hi, a_lo, b_hi, b_lo, idesc, accum);
                  umma_bf16_hl(tmem_d, a_hi, a_lo + 2u, b_hi, b_lo + 2u, idesc, 1u);
                  umma_bf16_hl(tmem_d, a_hi, a_lo + 4u, b_hi, b_lo + 4u, idesc, 1u);
                  umma_bf16_hl(tmem_d, a_hi, a_lo + 6u, b_hi, b_lo + 6u, idesc, 1u);
                } else {
                  for (int k = 0; k < k16_last; ++k) umma_bf16_hl(tmem_d, a_hi, a_lo + 2u * k, b_hi, b_lo + 2u * k, idesc, k ? 1u : accum);
                }
                accum = 1u;
                if (!p.resident) {
                  umma_commit_a(bar_empty_a + 8u * st);
                  if (++st == p.stages) { st = 0; ph ^= 1u; }
                }
              }
              umma_commit_a(bar_aempty_a + 8u * ast);
              if (++ast == p.a_stages) { ast = 0; aph ^= 1u; }
            }
          }
          umma_commit_a(bar_tfull_a + 8u * acc);
          continue;
        }
        int ylo, yhi, xlo, xhi;
        axis_range(p, ay, c.py, c.y0, p.tile_h, p.in_h, ylo, yhi);
        axis_range(p, ax, c.px, c.x0, p.tile_w, p.in_w, xlo, xhi);
        bool issued = false;
        auto mma_tap = [&](int tap_idx) {
          const uint32_t d_hi = desc_hi(1024u);
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait_a(bar_full_a + 8u * st, ph);
            tc::tc_fence_after();
            const uint32_t sa = smem_a + st * stage_bytes;
            const uint32_t sb = p.resident ? smem_a + p.res_base + (uint32_t)(tap_idx * p.kchunks + kc) * p.b_tile_bytes : sa + kABytes;
            const uint32_t a_lo = desc_lo(sa), b_lo = desc_lo(sb);
            if (kc != p.kchunks - 1 || k16_last == kBlockK / 16) {
              umma_bf16_hl(tmem_d, d_hi, a_lo, d_hi, b_lo, idesc, issued ? 1u : 0u);
              umma_bf16_hl(tmem_d, d_hi, a_lo + 2u, d_hi, b_lo + 2u, idesc, 1u);
              umma_bf16_hl(tmem_d, d_hi, a_lo + 4u, d_hi, b_lo + 4u, idesc, 1u);
              umma_bf16_hl(tmem_d, d_hi, a_lo + 6u, d_hi, b_lo + 6u, idesc, 1u);
            } else {
              for (int k = 0; k < k16_last; ++k) umma_bf16_hl(tmem_d, d_hi, a_lo + 2u * k, d_hi, b_lo + 2u * k, idesc, (issued || k != 0) ? 1u : 0u);
            }
            issued = true;
            umma_commit_a(bar_empty_a + 8u * st);
            if (++st == p.stages) { st = 0; ph ^= 1u; }
          }
        };
        const int tbase = c.phase * p.max_tp;
        for (int iy = ylo; iy < yhi; ++iy) {
          for (int ix = xlo; ix < xhi; ++ix) {
            const int e = iy * ax.cnt + ix;
            if (!tap_live(p, c, tapv[tbase + e], tape[tbase + e])) continue;
            mma_tap(e);
          }
        }
        if (!issued) mma_tap(ay.cnt * ax.cnt - 1);
        umma_commit_a(bar_tfull_a + 8u * acc);
        if (p.resident) {                                   // last tile of this weight set: the producer may overwrite it
          const int next = tile + gridDim.x;
          if (next >= p.total_tiles || decode_tile(p, next).phase != c.phase) tc::umma_commit(&bar_bfree);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..) =====================
    const int ew = warp - 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int wpg = p.epi_warps / p.epi_groups;   // warps per group: a multiple of 4, so a group covers the four lane quarters
    const int grp = ew / wpg;                     // this warp's group takes tiles grp, grp + epi_groups, ... of the CTA's list
    const int half = (ew - grp * wpg) >> 2;       // inside its group the warp takes slabs half, half + ngrp, ...
    const int ngrp = wpg >> 2;
    const int tile_step = p.epi_groups * (int)gridDim.x;
    const int r0 = q * 32;                        // first tile row of this warp
    const int w_off = r0 % p.tile_w, h_off = (r0 / p.tile_w) % p.tile_h, b_off = r0 / (p.tile_w * p.tile_h);
    uint8_t* ebase = smem + p.epi_base + (uint32_t)ew * p.epi_per_warp;
    uint8_t* s_o32 = ebase + p.off_o32;
    uint8_t* s_o16 = ebase + p.off_o16;
    uint8_t* s_o16a = ebase + p.off_o16a;
    uint8_t* s_aux = ebase + p.off_aux;           // 2 x (4 KB fp32 | 2 KB bf16)
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const int sw = lane & 7;                      // 128B swizzle: 16-byte chunk j of row r lives at chunk j ^ (r & 7)
    const int sw64 = p.slab == 32 ? (lane >> 1) & 3 : 0;   // 64B swizzle: chunk j of row r lives at chunk j ^ ((r >> 1) & 3); narrow slabs: dense rows
    const int nch = p.slab >> 3;                  // 16-byte bf16 chunks (8 columns) per slab row: 4, 3 or 2
    const int row16 = p.slab * 2;                 // bytes per bf16 staging row

    // aux prefetch runs one job ahead of the consumer
    int a_tile = blockIdx.x + grp * (int)gridDim.x, a_slab = half;
    uint32_t a_issued = 0, a_done = 0;
    auto aux_advance = [&]() {                    // skip to the next (tile, slab) this warp owns
      while (a_tile < p.total_tiles) {
        int tq, nt;
        lb_fast_divmod(p.d_nt, a_tile, tq, nt);
        const int n0 = nt * p.block_n;
        const int nsl = (min(p.block_n, p.out_c - n0) + p.slab - 1) / p.slab;
        if (a_slab < nsl) return;
        a_tile += tile_step; a_slab = half;
      }
    };
    auto aux_issue = [&]() {
      aux_advance();
      if (a_tile >= p.total_tiles) return;
      if (lane == 0) {
        const TileCoord c = decode_tile(p, a_tile);
        const uint32_t buf = a_issued & 1u;
        tc::mbar_arrive_expect_tx(&bar_aux[ew][buf], p.aux_bytes);
        tc::tma_load_4d(s_aux + buf * p.aux_bytes, &maps.aux[c.phase], &bar_aux[ew][buf], c.n0 + a_slab * p.slab, c.x0 + w_off,
                        c.y0 + h_off, c.b0 + b_off);
      }
      ++a_issued;
      a_slab += ngrp;
    };
    if (p.has_aux) aux_issue();

    int lt = grp;
    bool stores_pending = false;
    for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < p.total_tiles; tile += tile_step, lt += p.epi_groups) {
      const TileCoord c = decode_tile(p, tile);
      const int acc = lt & ((1 << p.acc_shift) - 1);
      const int nsl = (min(p.block_n, p.out_c - c.n0) + p.slab - 1) / p.slab;
      mbar_wait_a(bar_tfull_a + 8u * acc, ((uint32_t)lt >> p.acc_shift) & 1u);
      tc::tc_fence_after();
      bool released = false;
      for (int slab = half; slab < nsl; slab += ngrp) {
        if (p.has_aux) aux_issue();               // next job's aux while this one is processed
        float v[32];
        __syncwarp();
        {
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride + slab * p.slab);
          if (nch == 4) {
            tmem_ld32(ta, v);
          } else {                                // exact-width loads: the columns behind a narrow slab belong to other warps / nobody
            tmem_ld16_nw(ta, v);
            if (nch == 3) tmem_ld8_nw(ta + 16u, v + 16);
            tmem_ld_wait();
#pragma unroll
            for (int i = 16; i < 32; ++i) if (i >= 8 * nch) v[i] = 0.0f;
          }
        }
        if (slab + ngrp >= nsl) {                 // this warp's last read of the accumulator: hand it back before the math
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(bar_tempty_a + 8u * acc);
          released = true;
        }
        const int n = c.n0 + slab * p.slab;
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], alpha, (n + i < p.out_c) ? __ldg(p.bias + n + i) : 0.0f);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= alpha;
        }
        if (p.has_aux) {
          const uint32_t buf = a_done & 1u;
          tc::mbar_wait(&bar_aux[ew][buf], (a_done >> 1) & 1u);
          if (p.aux_bf16) {
            const uint8_t* row = s_aux + buf * p.aux_bytes + lane * row16;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j >= nch) break;
              const uint4 pk = *reinterpret_cast<const uint4*>(row + ((j ^ sw64) << 4));
              const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                if (p.aux_factor) {
                  v[8 * j + 2 * i] *= f.x;
                  v[8 * j + 2 * i + 1] *= f.y;
                } else {
                  v[8 * j + 2 * i] *= roottanh_grad_fast(f.x);
                  v[8 * j + 2 * i + 1] *= roottanh_grad_fast(f.y);
                }
              }
            }
          } else if (p.aux_factor) {
            const uint8_t* row = s_aux + buf * 4096 + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x4 = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
              v[4 * j + 0] *= x4.x; v[4 * j + 1] *= x4.y; v[4 * j + 2] *= x4.z; v[4 * j + 3] *= x4.w;
            }
          } else {
            const uint8_t* row = s_aux + buf * 4096 + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x4 = *reinterpret_cast<const float4*>(row + ((j ^ sw) << 4));
              v[4 * j + 0] *= roottanh_grad_fast(x4.x);
              v[4 * j + 1] *= roottanh_grad_fast(x4.y);
              v[4 * j + 2] *= roottanh_grad_fast(x4.z);
              v[4 * j + 3] *= roottanh_grad_fast(x4.w);
            }
          }
          ++a_done;
          __syncwarp();                           // every lane has read the buffer before it is refilled
        }
        // staging buffers are single: the previous slab's stores must have read them
        if (stores_pending) {
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        if (p.has_o32) {
          uint8_t* row = s_o32 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(row + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (p.has_o16 || p.has_o16a) {
          uint8_t* row = s_o16 + lane * row16;
          uint8_t* rowa = s_o16a + lane * row16;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j >= nch) break;
            __nv_bfloat162 h[4], a[4];
            if (p.o16_dact) {                    // RootTanh -> out16a, RootTanh' -> out16, all intermediates shared
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float f0, d0, f1, d1;
                roottanh_both_fast(v[8 * j + 2 * i], f0, d0);
                roottanh_both_fast(v[8 * j + 2 * i + 1], f1, d1);
                a[i] = __floats2bfloat162_rn(f0, f1);
                h[i] = __floats2bfloat162_rn(d0, d1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
              if (p.has_o16a) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  // with a stored pre-activation the function value belongs to the STORED (rounded) argument
                  const float2 f = p.has_o16 ? __bfloat1622float2(h[i]) : make_float2(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
                  a[i] = p.has_o16 ? __floats2bfloat162_rn(roottanh_fast(f.x), roottanh_fast(f.y))
                                   : __floats2bfloat162_rn(roottanh_only_fast(f.x), roottanh_only_fast(f.y));
                }
              }
            }
            if (p.has_o16) {
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h[0]); pk.y = *reinterpret_cast<uint32_t*>(&h[1]);
              pk.z = *reinterpret_cast<uint32_t*>(&h[2]); pk.w = *reinterpret_cast<uint32_t*>(&h[3]);
              *reinterpret_cast<uint4*>(row + ((j ^ sw64) << 4)) = pk;
            }
            if (p.has_o16a) {
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&a[0]); pk.y = *reinterpret_cast<uint32_t*>(&a[1]);
              pk.z = *reinterpret_cast<uint32_t*>(&a[2]); pk.w = *reinterpret_cast<uint32_t*>(&a[3]);
              *reinterpret_cast<uint4*>(rowa + ((j ^ sw64) << 4)) = pk;
            }
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.has_o32) tma_store_4d(&maps.o32[c.phase], s_o32, n, c.x0 + w_off, c.y0 + h_off, c.b0 + b_off);
          if (p.has_o16) tma_store_4d(&maps.o16[c.phase], s_o16, n, c.x0 + w_off, c.y0 + h_off, c.b0 + b_off);
          if (p.has_o16a) tma_store_4d(&maps.o16a[c.phase], s_o16a, n, c.x0 + w_off, c.y0 + h_off, c.b0 + b_off);
          bulk_commit();
        }
        stores_pending = true;
      }
      if (!released) {                            // a warp without a slab in this tile
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar_tempty_a + 8u * acc);
      }
    }
    if (lane == 0) bulk_wait0();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int pow2_ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }

}  // namespace

// Halo mode: can every tap of a source view be served from ONE shared-memory tile of that view?  Returns the largest
// vertical shift span over (phase, view) pairs, or -1 when the geometry does not qualify (single-tap layers gain nothing;
// full-extent feature-attention kernels have mostly dead taps, which only the per-tap path skips; small maps need
// batch-spanning tiles, whose 8-row groups are not equidistant in a halo tile).
static int halo_span_y(const lb_conv_geom* g) {
  static const int env_halo = getenv("LB_TC2_HALO") ? atoi(getenv("LB_TC2_HALO")) : 1;
  if (!env_halo) return -1;
  const int sp = g->mode == 1 ? g->stride : 1, s = g->stride, sh = s == 2 ? 1 : 0;
  const int dst_w = (g->out_w + sp - 1) / sp, dst_h = (g->out_h + sp - 1) / sp;
  if (g->kh > 8 || g->kw > 8 || dst_w < 8 || dst_h < 16) return -1;
  int span_y = 0, span_x = 0, max_taps = 0;
  for (int ph = 0; ph < sp * sp; ++ph) {
    const int py = ph / sp, px = ph % sp;
    int lo_y[kMaxViews], hi_y[kMaxViews], lo_x[kMaxViews], hi_x[kMaxViews], cnt[kMaxViews] = {0, 0, 0, 0};
    int taps = 0;
    for (int ty = 0; ty < g->kh; ++ty)
      for (int tx = 0; tx < g->kw; ++tx) {
        int dy, dx, view = 0;
        if (g->mode == 1) {
          if (((py + g->pad - ty) & (s - 1)) || ((px + g->pad - tx) & (s - 1))) continue;   // tap of another phase
          dy = (py + g->pad - ty) >> sh; dx = (px + g->pad - tx) >> sh;
        } else {
          const int oy = ty - g->pad, ox = tx - g->pad;
          dy = oy >> sh; dx = ox >> sh;
          view = ((oy & (s - 1)) << sh) + (ox & (s - 1));
        }
        if (!cnt[view]) { lo_y[view] = hi_y[view] = dy; lo_x[view] = hi_x[view] = dx; }
        lo_y[view] = dy < lo_y[view] ? dy : lo_y[view]; hi_y[view] = dy > hi_y[view] ? dy : hi_y[view];
        lo_x[view] = dx < lo_x[view] ? dx : lo_x[view]; hi_x[view] = dx > hi_x[view] ? dx : hi_x[view];
        ++cnt[view]; ++taps;
      }
    for (int v = 0; v < kMaxViews; ++v)
      if (cnt[v]) {
        span_y = hi_y[v] - lo_y[v] > span_y ? hi_y[v] - lo_y[v] : span_y;
        span_x = hi_x[v] - lo_x[v] > span_x ? hi_x[v] - lo_x[v] : span_x;
      }
    max_taps = taps > max_taps ? taps : max_taps;
  }
  if (max_taps < 2 || span_x > kHaloPitch - 8) return -1;
  return span_y;
}
extern "C" int lb_tc2_halo_eligible(const lb_conv_geom* g) { return g && halo_span_y(g) >= 0 ? 1 : 0; }

// Does the persistent kernel cover this geometry and these operands?  (Otherwise lb_conv_tc_gemm stays on k_conv_tc.)
static bool tc2_ok(const lb_conv_geom* g, const void* out32, const void* out16, const void* out16a, int ld16, const void* aux, int ld_aux,
                   int aux_dtype) {
  if (g->stride != 1 && g->stride != 2) return false;
  if (g->ld_in % 8 || g->in_c < 1 || g->out_c < 1 || g->kh * g->kw > 1024) return false;
  if (g->mode == 1 && (g->kh < g->stride || g->kw < g->stride)) return false;          // a phase without taps
  {
    const int sp_ = g->mode == 1 ? g->stride : 1;
    if ((g->mode == 1 ? ((g->kh + sp_ - 1) / sp_) * ((g->kw + sp_ - 1) / sp_) : g->kh * g->kw) > 64) return false;   // tap table size
  }
  if (out32 && ((g->ld_out & 3) || (reinterpret_cast<uintptr_t>(out32) & 15))) return false;
  if (out16 && ((ld16 & 7) || (reinterpret_cast<uintptr_t>(out16) & 15))) return false;
  if (out16a && ((ld16 & 7) || (reinterpret_cast<uintptr_t>(out16a) & 15))) return false;
  if (aux && ((ld_aux & (aux_dtype == LB_BF16 ? 7 : 3)) || (reinterpret_cast<uintptr_t>(aux) & 15))) return false;
  if (aux_dtype != LB_F32 && aux_dtype != LB_BF16) return false;
  if (!out32 && !out16 && !out16a) return false;
  // weight-bound layers whose output tiling cannot fill the GPU stay on k_conv_tc's split-K path
  const int sp = g->mode == 1 ? g->stride : 1;
  const int dst_w = (g->out_w + sp - 1) / sp, dst_h = (g->out_h + sp - 1) / sp;
  const int tw = pow2_ceil(dst_w) < kBlockM ? pow2_ceil(dst_w) : kBlockM;
  const int th = pow2_ceil(dst_h) < kBlockM / tw ? pow2_ceil(dst_h) : kBlockM / tw;
  const int tb = kBlockM / (tw * th);
  const long long m_tiles = (long long)((dst_w + tw - 1) / tw) * ((dst_h + th - 1) / th) * ((g->batch + tb - 1) / tb) * sp * sp;
  const int taps_eff = g->mode == 1 ? ((g->kh + sp - 1) / sp) * ((g->kw + sp - 1) / sp) : g->kh * g->kw;
  const int iters_est = taps_eff * ((g->in_c + kBlockK - 1) / kBlockK);
  if (m_tiles * ((g->out_c + 127) / 128) * 2 <= LB_SMS && iters_est >= 8) return false;
  return true;
}

extern "C" int lb_conv_tc_gemm_ex(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, float* out32,
                                  void* out16, void* out16a, int ld_out16, const void* aux, int ld_aux, int aux_dtype, int flags,
                                  const lb_conv_geom* g, lb_stream_t s) {
  LB_REQUIRE(in_bf16 && w_packed && g);
  LB_REQUIRE(!(flags & ~(LB_EX_AUX_IS_FACTOR | LB_EX_OUT16_IS_DACT)));
  LB_REQUIRE(!(flags & LB_EX_OUT16_IS_DACT) || (out16 && out16a));
  if (!tc2_ok(g, out32, out16, out16a, ld_out16, aux, ld_aux, aux_dtype)) return LB_EUNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(in_bf16) & 15 || reinterpret_cast<uintptr_t>(w_packed) & 15) return LB_EALIGN;
  Tc2Maps maps;
  Tc2Params p;
  p.mode = g->mode; p.stride = g->stride; p.pad = g->pad; p.kh = g->kh; p.kw = g->kw;
  p.sp = g->mode == 1 ? g->stride : 1;
  p.batch = g->batch; p.out_c = g->out_c; p.in_c = g->in_c; p.in_h = g->in_h; p.in_w = g->in_w;
  p.sh = g->stride == 2 ? 1 : 0;
  const int dst_w = (g->out_w + p.sp - 1) / p.sp, dst_h = (g->out_h + p.sp - 1) / p.sp;
  p.tile_w = pow2_ceil(dst_w) < kBlockM ? pow2_ceil(dst_w) : kBlockM;
  int rest = kBlockM / p.tile_w;
  p.tile_h = pow2_ceil(dst_h) < rest ? pow2_ceil(dst_h) : rest;
  p.tile_b = rest / p.tile_h;
  int span_y = halo_span_y(g);
  p.halo = 0;
  if (span_y >= 0) {                   // 8 x 16 pixel tiles of one image: 8-row groups = rows of the tile, 16 halo rows apart
    static const int env_halo = getenv("LB_TC2_HALO") ? atoi(getenv("LB_TC2_HALO")) : 1;
    p.halo = 1;
    p.tile_w = 8; p.tile_h = 16; p.tile_b = 1;
  }
  p.tiles_w = (dst_w + p.tile_w - 1) / p.tile_w;
  p.tiles_h = (dst_h + p.tile_h - 1) / p.tile_h;
  p.tiles_b = (g->batch + p.tile_b - 1) / p.tile_b;
  p.ebw = p.tile_w < 32 ? p.tile_w : 32;
  p.ebh = p.tile_h < 32 / p.ebw ? p.tile_h : 32 / p.ebw;
  p.ebb = 32 / (p.ebw * p.ebh);
  p.kchunks = (g->in_c + kBlockK - 1) / kBlockK;
  p.rows_per_tap = g->out_c;
  p.alpha = alpha; p.bias = bias;
  p.has_o32 = out32 ? 1 : 0; p.has_o16 = out16 ? 1 : 0; p.has_o16a = out16a ? 1 : 0; p.has_aux = aux ? 1 : 0;
  p.aux_bf16 = aux_dtype == LB_BF16 ? 1 : 0;
  p.aux_factor = (flags & LB_EX_AUX_IS_FACTOR) ? 1 : 0;
  p.o16_dact = (flags & LB_EX_OUT16_IS_DACT) ? 1 : 0;
  p.slab = kSlab;                      // narrowed below once the channel tile is known

  // Slab width.  A warp reads only its own TMEM lane quarter, so a 128-row tile is spread over the 16 epilogue warps by
  // COLUMN slabs: 4 warps per quarter.  With 32-column slabs a 96-channel tile keeps 12 warps busy and a 48-channel one 8;
  // bf16-only epilogues may use 24 / 16-column slabs (dense 48 / 32-byte staging rows, exact-width tcgen05.ld), which puts
  // 96 channels on 16 warps x 24 columns and 48 channels on 12 warps x 16 columns.  The epilogue (RootTanh + RootTanh' per
  // element) is the critical path of exactly these small-channel layers.
  {
    static const int env_slab = getenv("LB_TC2_SLAB") ? atoi(getenv("LB_TC2_SLAB")) : 0;
    const int bn1 = (g->out_c + 15) / 16 * 16;              // the single channel tile this applies to
    if (!out32 && (!aux || p.aux_bf16) && g->out_c <= 256 && env_slab != 32) {
      // Single-tap layers (their tile period is the epilogue's latency chain): first the width that lets the most warp groups
      // work on different tiles (1 slab: 4 groups, 2 slabs: 2), then the fewest columns per warp.  Multi-tap layers are
      // paced by the MMA-issuing thread and want their shared memory for halo stages / resident weights (3x3 48 -> 48 lost
      // its resident weights to 24-column staging and ran 18 % slower): fewest columns per warp, as before.
      const int taps_phase = g->mode == 1 ? ((g->kh + p.sp - 1) / p.sp) * ((g->kw + p.sp - 1) / p.sp) : g->kh * g->kw;
      const bool by_groups = taps_phase == 1;
      auto groups_of = [&](int nsl) { return !by_groups ? 1 : nsl <= 1 ? 4 : nsl == 2 ? 2 : 1; };
      int best = kSlab, best_grp = groups_of((bn1 + kSlab - 1) / kSlab), best_cols = kSlab * ((((bn1 + kSlab - 1) / kSlab) + 3) / 4);
      const int cand[2] = {24, 16};
      for (int i = 0; i < 2; ++i) {
        const int nsl = (bn1 + cand[i] - 1) / cand[i];
        const int cols = cand[i] * ((nsl + 3) / 4), grp = groups_of(nsl);
        if (grp > best_grp || (grp == best_grp && cols < best_cols)) { best = cand[i]; best_grp = grp; best_cols = cols; }
      }
      p.slab = best;
      if (env_slab == 16 || env_slab == 24) p.slab = env_slab;
    }
  }
  const int tap_tab_bytes = 4 * kMaxTapsPhase * (int)(sizeof(int4) + sizeof(int2));  // 6 KB
  const int tab_bytes = tap_tab_bytes + 4096;                                         // + the halo-mode view / tap tables
  p.max_tp = g->mode == 1 ? ((g->kh + p.sp - 1) / p.sp) * ((g->kw + p.sp - 1) / p.sp) : g->kh * g->kw;
  if (p.max_tp > kMaxTapsPhase) return LB_EUNSUPPORTED;
  int epi_bytes = 0, nt = 0, bn = 0, stage_bytes = 0, stages = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    p.aux_bytes = (uint32_t)(p.aux_bf16 ? 32 * p.slab * 2 : 32 * p.slab * 4);
    // epilogue staging per warp: fp32 slab 4 KB, bf16 slabs 2 KB each (1 KB for 16 columns), aux 2 x (4 | 2) KB, all 1 KB aligned
    const uint32_t o16_bytes = (uint32_t)((32 * p.slab * 2 + 1023) & ~1023);
    const uint32_t aux_slot = (p.aux_bytes + 1023) & ~1023u;
    uint32_t off = 0;
    p.off_o32 = off; if (out32) off += 4096;
    p.off_o16 = off; if (out16) off += o16_bytes;
    p.off_o16a = off; if (out16a) off += o16_bytes;
    p.off_aux = off; if (aux) off += 2 * aux_slot;
    p.epi_per_warp = (off + 1023) & ~1023u;
    p.epi_warps = p.epi_per_warp * 16 <= 100 * 1024 ? 16 : 8;
    epi_bytes = (int)p.epi_per_warp * p.epi_warps + tab_bytes;

    // channel tile: as wide as TMEM double buffering allows (<= 256) while leaving >= 3 ring stages
    nt = (g->out_c + 255) / 256;
    for (;; ++nt) {
      // several channel tiles: a multiple of the 32-column slab, so no tile's last store reaches into its neighbour
      bn = nt > 1 ? ((g->out_c + nt - 1) / nt + 31) / 32 * 32 : (g->out_c + 15) / 16 * 16;
      stage_bytes = kABytes + ((bn * kBlockK * 2 + 1023) & ~1023);
      stages = (kSmemLimit - 1024 - epi_bytes) / stage_bytes;
      if (stages >= 3 || bn <= 64) break;
    }
    if (nt == 1 || bn % p.slab == 0 || p.slab == kSlab) break;
    p.slab = kSlab;                    // the tile had to be split and the narrow slab does not divide it: plan again with 32
  }
  if (stages < 2) return LB_EUNSUPPORTED;
  if (stages > 8) stages = 8;
  // Optional (LB_TC2_RESIDENT=1): all taps x chunks of one output phase's weights (<= ~100 KB) are loaded once per phase
  // and stay resident, so ring stages carry A only.  Measured on B200 (profiles/README.md): no gain -- the small-channel
  // layers are bound by the rate at which the TMA unit processes the 128 strided rows of each A box (~380 ns per box),
  // not by the bytes of the weight tile -- so it stays off by default.
  p.resident = 0; p.res_base = 0;
  p.b_tile_bytes = (uint32_t)((bn * kBlockK * 2 + 1023) & ~1023);
  {
    const int taps_phase = g->mode == 1 ? ((g->kh + p.sp - 1) / p.sp) * ((g->kw + p.sp - 1) / p.sp) : g->kh * g->kw;
    const long long res_bytes = (long long)taps_phase * p.kchunks * p.b_tile_bytes;
    const int a_stages = (int)((kSmemLimit - 1024 - epi_bytes - res_bytes) / kABytes);
    static const bool env_resident = getenv("LB_TC2_RESIDENT") != nullptr;   // debugging switches are read once, not per launch
    if (nt == 1 && res_bytes <= 100 * 1024 && a_stages >= 4 && env_resident) {
      p.resident = 1;
      stage_bytes = kABytes;
      stages = a_stages;
      if (stages > 8) stages = 8;
      p.res_base = (uint32_t)(stages * stage_bytes);
    }
  }
  {
    static const int env_stages = getenv("LB_TC2_STAGES") ? atoi(getenv("LB_TC2_STAGES")) : 0;
    if (env_stages >= 2 && env_stages < stages) { stages = env_stages; if (p.resident) p.res_base = (uint32_t)(stages * stage_bytes); }
  }
  p.a_stages = 0; p.a_stage_bytes = 0; p.a_base = 0;
  if (p.halo) {
    // B ring (one weight tile per stage) + A ring (one halo tile per stage)
    p.a_stage_bytes = (uint32_t)(kHaloPitch * (16 + span_y) * 128);
    const int budget = kSmemLimit - 1024 - epi_bytes;
    int a_st = 3, b_st = (budget - a_st * (int)p.a_stage_bytes) / (int)p.b_tile_bytes;
    if (b_st < 4) { a_st = 2; b_st = (budget - a_st * (int)p.a_stage_bytes) / (int)p.b_tile_bytes; }
    bool halo_res = false;
    {
      // single-phase layers whose whole weight set fits next to >= 3 halo stages (3x3 48 -> 48: 54 KB): weights resident,
      // the MMA thread's per-tap barrier wait + commit disappear (it is the critical path there: 9 taps x ~500 clocks of
      // scalar issue work per 128-pixel tile against 650 clocks of tensor time)
      static const int env_hres = getenv("LB_TC2_HALO_RESIDENT") ? atoi(getenv("LB_TC2_HALO_RESIDENT")) : 1;
      const long long res_bytes = (long long)p.max_tp * p.kchunks * p.b_tile_bytes;
      const long long a_fit = ((long long)budget - res_bytes) / (long long)p.a_stage_bytes;
      if (env_hres && !p.resident && nt == 1 && p.sp == 1 && a_fit >= 3) {
        halo_res = true;
        p.resident = 1;
        p.a_stages = a_fit > 4 ? 4 : (int)a_fit;
        stages = 0; stage_bytes = 0;
        p.a_base = 0;
        p.res_base = (uint32_t)p.a_stages * p.a_stage_bytes;
      }
    }
    if (halo_res) {
      // layout: [halo ring][resident weights][epilogue staging]
    } else if (b_st < 3 || p.resident) {
      p.halo = 0;                      // does not fit: back to per-tap boxes with the generic tile shape
      p.tile_w = pow2_ceil(dst_w) < kBlockM ? pow2_ceil(dst_w) : kBlockM;
      rest = kBlockM / p.tile_w;
      p.tile_h = pow2_ceil(dst_h) < rest ? pow2_ceil(dst_h) : rest;
      p.tile_b = rest / p.tile_h;
      p.tiles_w = (dst_w + p.tile_w - 1) / p.tile_w;
      p.tiles_h = (dst_h + p.tile_h - 1) / p.tile_h;
      p.tiles_b = (g->batch + p.tile_b - 1) / p.tile_b;
      p.ebw = p.tile_w < 32 ? p.tile_w : 32;
      p.ebh = p.tile_h < 32 / p.ebw ? p.tile_h : 32 / p.ebw;
      p.ebb = 32 / (p.ebw * p.ebh);
    } else {
      if (b_st > 8) b_st = 8;
      stages = b_st; stage_bytes = (int)p.b_tile_bytes;
      p.a_stages = a_st;
      p.a_base = (uint32_t)(stages * stage_bytes);
    }
  }
  p.block_n = bn; p.n_tiles = (g->out_c + bn - 1) / bn; p.stages = stages;
  p.acc_stride = ((bn + p.slab - 1) / p.slab * p.slab + 31) / 32 * 32;   // whole slabs, so that exact-width loads stay inside
  // epilogue warp groups / accumulator ring depth (Tc2Params::epi_groups): only for a single channel tile whose slabs leave
  // whole groups of 4 warps idle; the ring is 4 deep then, so that the MMA warp stays ahead of every group
  {
    static const int env_groups = getenv("LB_TC2_GROUPS") ? atoi(getenv("LB_TC2_GROUPS")) : 4;
    const int nsl = (bn + p.slab - 1) / p.slab;
    int groups = p.n_tiles == 1 ? p.epi_warps / (4 * nsl) : 1;
    groups = groups >= 4 ? 4 : groups >= 2 ? 2 : 1;
    if (groups > env_groups) groups = env_groups < 1 ? 1 : env_groups;
    if (groups == 3) groups = 2;
    // wide tiles (3-4 slabs): two groups of 8 warps, every warp takes two slabs of every second tile -- the same work per
    // warp, but the two groups' accumulator waits, TMEM loads and MUFU bursts are staggered instead of synchronous
    static const int env_split = getenv("LB_TC2_SPLIT_GROUPS") ? atoi(getenv("LB_TC2_SPLIT_GROUPS")) : 1;   // measured: 1x1 48->96 -7 %, 96->96 -7..-10 %, step -0.8 %
    if (env_split && groups == 1 && env_groups >= 2 && p.n_tiles == 1 && p.epi_warps == 16 && nsl >= 3 && nsl <= 4 && p.acc_stride * 4 <= 512)
      groups = 2;
    p.acc_shift = groups > 1 ? 2 : 1;
    if ((p.acc_stride << p.acc_shift) > 512) { p.acc_shift = 1; if (groups > 2) groups = 2; }
    p.epi_groups = groups;
  }
  p.tmem_cols = (uint32_t)pow2_ceil((p.acc_stride << p.acc_shift) < 32 ? 32 : (p.acc_stride << p.acc_shift));
  p.epi_base = (uint32_t)(stages * stage_bytes) + (uint32_t)p.a_stages * p.a_stage_bytes + (p.resident ? (uint32_t)(((g->mode == 1 ? ((g->kh + p.sp - 1) / p.sp) * ((g->kw + p.sp - 1) / p.sp) : g->kh * g->kw)) * p.kchunks) * p.b_tile_bytes : 0u);
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.n_tiles * p.sp * p.sp;
  static const int env_phase_inner = getenv("LB_TC2_PHASE_INNER") ? atoi(getenv("LB_TC2_PHASE_INNER")) : 1;
  p.phase_inner = (p.resident || !env_phase_inner) ? 0 : 1;
  p.d_nt = lb_make_fastdiv(p.n_tiles); p.d_tw = lb_make_fastdiv(p.tiles_w); p.d_th = lb_make_fastdiv(p.tiles_h); p.d_tb = lb_make_fastdiv(p.tiles_b);

  // source views: mode 0 with stride 2 -> 4 parity views; otherwise one dense view
  const int nv = (g->mode == 0) ? g->stride * g->stride : 1;
  const int vs = (g->mode == 0) ? g->stride : 1;
  p.n_views = nv;
  p.view_empty = 0;
  const char* base = reinterpret_cast<const char*>(in_bf16);
  for (int v = 0; v < kMaxViews; ++v) {
    const int vv = v < nv ? v : 0;
    const int qy = vv / vs, qx = vv % vs;
    int vw = (g->in_w - qx + vs - 1) / vs, vh = (g->in_h - qy + vs - 1) / vs;
    if (vw <= 0 || vh <= 0) { if (v < nv) p.view_empty |= 1 << v; vw = vw > 0 ? vw : 1; vh = vh > 0 ? vh : 1; }
    const uint64_t dims[4] = {(uint64_t)g->in_c, (uint64_t)vw, (uint64_t)vh, (uint64_t)g->batch};
    const uint64_t strides[3] = {(uint64_t)vs * g->ld_in * 2, (uint64_t)vs * g->in_w * g->ld_in * 2,
                                 (uint64_t)g->in_h * g->in_w * g->ld_in * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.tile_w, (uint32_t)p.tile_h, (uint32_t)p.tile_b};
    if (p.halo) { box[1] = kHaloPitch; box[2] = (uint32_t)(16 + span_y); box[3] = 1; }   // one box serves every tap shift of the view
    const char* vbase = ((p.view_empty >> v) & 1) ? base : base + ((size_t)qy * g->in_w + qx) * g->ld_in * 2;
    int rc = tc::make_map(&maps.a[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, vbase, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const int kpad = (g->in_c + 7) / 8 * 8;
    const uint64_t dims[2] = {(uint64_t)g->in_c, (uint64_t)g->kh * g->kw * g->out_c};
    const uint64_t strides[1] = {(uint64_t)kpad * 2};
    const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)bn};
    int rc = tc::make_map(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_packed, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  // destination views (one per output phase): 32-row x 32-column boxes
  for (int v = 0; v < kMaxViews; ++v) {
    const int vv = v < p.sp * p.sp ? v : 0;
    const int py = vv / p.sp, px = vv % p.sp;
    int vw = (g->out_w - px + p.sp - 1) / p.sp, vh = (g->out_h - py + p.sp - 1) / p.sp;
    const bool empty = vw <= 0 || vh <= 0;      // no tile ever stores there (tiles are clipped), keep the map valid
    if (vw < 1) vw = 1;
    if (vh < 1) vh = 1;
    const uint64_t dims[4] = {(uint64_t)g->out_c, (uint64_t)vw, (uint64_t)vh, (uint64_t)g->batch};
    const uint32_t box[4] = {(uint32_t)p.slab, (uint32_t)p.ebw, (uint32_t)p.ebh, (uint32_t)p.ebb};
    const CUtensorMapSwizzle sw16 = p.slab == kSlab ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;   // narrow slabs: dense rows
    const size_t pix = empty ? 0 : (size_t)py * g->out_w + px;
    if (out32) {
      const uint64_t st[3] = {(uint64_t)p.sp * g->ld_out * 4, (uint64_t)p.sp * g->out_w * g->ld_out * 4,
                              (uint64_t)g->out_h * g->out_w * g->ld_out * 4};
      int rc = tc::make_map(&maps.o32[v], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, reinterpret_cast<char*>(out32) + pix * g->ld_out * 4, 4,
                            dims, st, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    if (out16 || out16a) {
      const uint64_t st[3] = {(uint64_t)p.sp * ld_out16 * 2, (uint64_t)p.sp * g->out_w * ld_out16 * 2,
                              (uint64_t)g->out_h * g->out_w * ld_out16 * 2};
      if (out16) {
        int rc = tc::make_map(&maps.o16[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, reinterpret_cast<char*>(out16) + pix * ld_out16 * 2, 4,
                              dims, st, box, sw16);
        if (rc) return rc;
      }
      if (out16a) {
        int rc = tc::make_map(&maps.o16a[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, reinterpret_cast<char*>(out16a) + pix * ld_out16 * 2, 4,
                              dims, st, box, sw16);
        if (rc) return rc;
      }
    }
    if (aux) {
      const uint64_t es = p.aux_bf16 ? 2 : 4;
      const uint64_t st[3] = {(uint64_t)p.sp * ld_aux * es, (uint64_t)p.sp * g->out_w * ld_aux * es,
                              (uint64_t)g->out_h * g->out_w * ld_aux * es};
      int rc = tc::make_map(&maps.aux[v], p.aux_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (int)es,
                            const_cast<char*>(reinterpret_cast<const char*>(aux)) + pix * ld_aux * es, 4, dims, st, box,
                            p.aux_bf16 ? sw16 : CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  p.tab_base = p.epi_base + p.epi_per_warp * p.epi_warps;
  p.htab_base = p.tab_base + (uint32_t)tap_tab_bytes;
  const int smem_bytes = (int)p.epi_base + epi_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = p.total_tiles < LB_SMS ? p.total_tiles : LB_SMS;
  lb_launch(k_conv_tc2, grid, 64 + 32 * p.epi_warps, smem_bytes, lb_s(s), maps, p);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// out32_used: an fp32 output (row stride g->ld_out) is written; ld_out16 / ld_aux: 0 = not used
extern "C" int lb_conv_tc_ex_supported(const lb_conv_geom* g, int out32_used, int ld_out16, int ld_aux, int aux_dtype) {
  if (!g) return 0;
  // alignment of the pointers themselves is checked at call time; any 16-byte aligned base passes here
  const void* ok = reinterpret_cast<const void*>(16);
  return tc2_ok(g, out32_used ? ok : nullptr, ld_out16 ? ok : nullptr, nullptr, ld_out16, ld_aux ? ok : nullptr, ld_aux, aux_dtype) ? 1 : 0;
}
