/*
 * locate_b200 -- C ABI of the B200-native LocAtE hot path (generator/discriminator block
 * forward + backward).  Plain C: borrowed device pointers, sizes, a cudaStream_t; no torch
 * types.  Every entry point returns 0 on success, a positive cudaError_t on a CUDA failure or
 * a negative LB_E* code on an argument error.  No entry point allocates or synchronises;
 * work is enqueued on `stream` and all pointers must stay valid until it drains.
 *
 * The reference (ClashLuke/LocAtE, /root/reference) has no FFI: its operator boundary is the
 * Python API of libs/ (SURVEY.md section 8b).  Each entry point below names the reference
 * arithmetic (file:line) it replaces; locate_b200/*.py rebuilds the reference's nn.Module tree
 * on top of these calls and INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Activation layout: channels-last.  A logical [B,C,H,W] tensor is stored [B][H][W][C]
 * (fp32 unless stated); "rows" below are pixels (b,h,w) and "ld" is the channel stride of a
 * row, so a channel slice of a wider tensor is (ptr + channel_offset, ld = full C).
 */
#ifndef LOCATE_B200_H
#define LOCATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* lb_stream_t; /* == cudaStream_t */

/* storage type of activations and their gradients (the `dtype` argument of the elementwise / reduction entry points):
 * arithmetic is fp32 either way; bf16 storage is the tensor-core configuration (half the HBM bytes of every pass). */
enum { LB_F32 = 0, LB_BF16 = 1 };

enum {
  LB_OK = 0,
  LB_EINVAL = -1,   /* bad size / null pointer / unsupported combination */
  LB_EALIGN = -2,   /* pointer or leading dimension not aligned as the kernel needs */
  LB_EUNSUPPORTED = -3
};

/* Library / device info. lb_version returns e.g. 100 for 0.1.0; lb_sm_arch the compiled arch (100). */
int lb_version(void);
int lb_sm_arch(void);
int lb_last_launch_count(void);   /* kernels launched by this library since lb_reset_launch_count */
void lb_reset_launch_count(void);
/* Programmatic dependent launch of every kernel of the library (default on; environment LB_PDL=0 switches it off):
 * kernels are launched with the programmatic-stream-serialization attribute and wait for their predecessor on the
 * device (griddepcontrol.wait), so launch latency and kernel prologues overlap the predecessor's tail -- in a stream
 * and, as programmatic edges, in a captured CUDA graph.  Returns the previous setting.  (No reference counterpart:
 * the reference leaves launch scheduling to PyTorch eager mode, main.py:146-172.) */
int lb_set_pdl(int on);

/* ---- RootTanh activation: y = (x^2+1)^(1/growth) * tanh(x)          libs/activation.py:9-16
 *      bwd: dx = g * (2(x^2+1) sech^2 x + x tanh x) / (2 (x^2+1)^((growth-1)/growth))  :20-36 */
int lb_roottanh_fwd(const void* x, void* y, size_t n, int growth, int dtype, lb_stream_t stream);
int lb_roottanh_bwd(const void* x, const void* g, void* dx, size_t n, int growth, int dtype, lb_stream_t stream);
/* tanh on the generator output                                          libs/models.py:66 */
int lb_tanh_fwd(const void* x, void* y, size_t n, int dtype, lb_stream_t stream);
int lb_tanh_bwd(const void* y, const void* g, void* dx, size_t n, int dtype, lb_stream_t stream);
/* y = a + b (the sum autograd forms when a tensor feeds two branches, e.g. block.py:44-46 skip path + gated branch) */
int lb_add(const void* a, const void* b, void* y, size_t n, int dtype, lb_stream_t stream);
/* y = a * b (a gradient times the stored RootTanh' factor, where no GEMM epilogue can apply it) */
int lb_mul(const void* a, const void* b, void* y, size_t n, int dtype, lb_stream_t stream);

/* hinge(t) = max(1 - t, 0) elementwise                                   libs/utils.py:133-134 */
int lb_hinge_fwd(const float* x, float* y, size_t n, lb_stream_t stream);
int lb_hinge_bwd(const float* x, const float* g, float* dx, size_t n, lb_stream_t stream);

/* ---- whole-tensor norm ("InPlaceNorm")                                libs/inplace_norm.py:7-45
 * stats pipeline: lb_norm_stats writes (sum, sum of squares) of x to sums[2] (double; in data parallel the
 * caller all-reduces sums across ranks), then lb_norm_finalize turns (sums, n_total) into
 * stats[4] = {mean, std(unbiased), 1/std, n_total}.
 * The reduction uses no floating-point atomics (per-CTA partials added in index order by the last CTA), so the
 * statistics -- and with them the whole forward pass -- are bit-reproducible like the reference's.  `work`:
 * lb_stat_work_doubles() doubles, zero-filled ONCE when allocated (the kernels leave it zeroed); one per stream. */
size_t lb_stat_work_doubles(void);
int lb_norm_stats(const void* x, size_t n, double* sums, double* work, int dtype, lb_stream_t stream);
int lb_norm_finalize(const double* sums, double n_total, float* stats, lb_stream_t stream);
/* y = (x-mean)*gain/std + bias ; gain is [C] (gain_batch_stride = 0) or [B][C] (stride = C); x, y in storage `dtype`,
 * statistics / gain / bias always fp32. */
int lb_norm_apply(const void* x, const float* stats, const float* gain, int gain_batch_stride,
                  const float* bias, void* y, int batch, int pixels, int channels, int dtype, lb_stream_t stream);
/* same, also emitting act = RootTanh(y), the operand of the convolution that consumes the result (conv.py:23-24), in
 * the same pass, and optionally dact = RootTanh'(y) (activation.py:20-36), the factor that convolution's input gradient
 * is multiplied by (lb_conv_tc_gemm_ex, LB_EX_AUX_IS_FACTOR); y may be NULL when nothing else reads it.  Needs
 * channels % (16 bytes of elements) == 0 and 16-byte aligned pointers (LB_EALIGN otherwise: lb_norm_apply + lb_roottanh_fwd). */
int lb_norm_apply_ex(const void* x, const float* stats, const float* gain, int gain_batch_stride, const float* bias,
                     void* y, void* act, void* dact, int batch, int pixels, int channels, int dtype, lb_stream_t stream);
/* backward, 3 steps (inplace_norm.py:17-27 composed with d std/dx):
 *  1. lb_norm_bwd_reduce: p1[b][c] += sum_hw g, p2[b][c] += sum_hw (x-mean)*g      (caller zeroes p1,p2)
 *  2. lb_norm_bwd_finalize: dgain (+=, [C] or [B][C]), dbias (+=, [C]), s[2] = {sum gain*p1, sum gain*p2}
 *     (double; in data parallel the caller all-reduces s)
 *  3. lb_norm_bwd_apply: dx = gain*g/std - s0/(std*N) - s1*(x-mean)/((N-1)*std^3) (+ add, optional: the gradient
 *     reaching x through another branch, summed in the same pass)                                     */
int lb_norm_bwd_reduce(const void* x, const void* g, const float* stats, float* p1, float* p2,
                       int batch, int pixels, int channels, int dtype, lb_stream_t stream);
int lb_norm_bwd_finalize(const float* p1, const float* p2, const float* gain, int gain_batch_stride,
                         const float* stats, int batch, int channels, float* dgain, float* dbias,
                         double* s, lb_stream_t stream);
int lb_norm_bwd_apply(const void* x, const void* g, const float* stats, const float* gain,
                      int gain_batch_stride, const double* s, const void* add, void* dx, int batch, int pixels,
                      int channels, int dtype, lb_stream_t stream);

/* ---- gated residual: out = (gamma*y + 1)*x                            libs/merge.py:19-39
 * y is a full tensor (y_bcast = 0) or a per-(b,c) gate [B][C] broadcast over pixels (y_bcast = 1;
 * the Expand of feature attention, libs/util_modules.py:6-12, never materialised).  x, y, out, g, dx, dy share one
 * storage type (`dtype`).
 * bwd: dx = (gamma*y+1)*g ; dy = gamma*x*g, or -- y_bcast -- dy_bcast[b][c] += gamma * sum_pixels x*g (fp32, zeroed by the
 *      caller); dgamma += sum x*x*g when strict_reference (the reference's formula, merge.py:33-38)
 *                or sum x*y*g otherwise.  gamma is a device scalar. */
int lb_gate_fwd(const void* x, const void* y, const float* gamma, void* out, int batch, int pixels,
                int channels, int y_bcast, int dtype, lb_stream_t stream);
/* same, also writing sums[2] (fp64) = (sum out, sum out^2): the statistics of the whole-tensor norm that consumes every
 * gate output (block.py:46-51), without re-reading it (work: as lb_norm_stats).  channels % 4 == 0, aligned pointers. */
int lb_gate_fwd_stats(const void* x, const void* y, const float* gamma, void* out, double* sums, double* work, int batch,
                      int pixels, int channels, int y_bcast, int dtype, lb_stream_t stream);
int lb_gate_bwd(const void* x, const void* y, const float* gamma, const void* g, void* dx, void* dy, float* dy_bcast,
                float* dgamma, int batch, int pixels, int channels, int y_bcast, int strict_reference, int dtype,
                lb_stream_t stream);

/* ---- spectral norm power iteration, one per forward                   libs/spectral_norm.py:21-32
 * W is the row-major [height][width] view of weight_bar.  u,v updated in place; sigma_out[0] = sigma,
 * sigma_out[1] = 1/sigma.  work: lb_sn_power_iter_work_floats(height, width) floats of scratch (per-row-split partial
 * sums of W^T u, added in a fixed order: the iteration is bit-reproducible like the reference's torch.mv). */
size_t lb_sn_power_iter_work_floats(int height, int width);
int lb_sn_power_iter(const float* w, int height, int width, float* u, float* v, float* sigma_out,
                     float* work, lb_stream_t stream);
/* The same iteration for ALL spectral-normed layers of a model in 4 launches.  layers_dev: device array of
 * n_layers records {const float* w; float* u; float* v; float* dv; int height, width, t_off, s_off, nsplit,
 * rows_per_split;} (56 bytes each; t_off indexes the float `scratch` where the nsplit x width partial sums of W^T u of
 * that layer are staged -- one row per block of rows_per_split rows, added in a fixed order -- and s_off indexes
 * `s_out`, which receives s = W v of every layer, the pre-normalisation u, kept by the caller for the backward of
 * trainable u; dv: see lb_sn_uv_grad_batched, may be NULL).  items1_dev: int4 {layer, col0, row0, rows} tiles of 256
 * columns for pass 1 (row0 a multiple of rows_per_split); items3_dev: int2 {layer, row0} groups of 8 rows for pass 2.
 * sigma_out: n_layers x {sigma, 1/sigma}.  No floating-point atomics: two runs from the same (W, u) give bit-identical
 * u, v, sigma. */
int lb_sn_power_iter_batched(const void* layers_dev, int n_layers, const void* items1_dev, int n_items1,
                             const void* items3_dev, int n_items3, float* scratch, float* s_out,
                             float* sigma_out, lb_stream_t stream);
/* weight-gradient epilogue: with dwn = dL/d(W/sigma) and the LIVE u,v:
 *   grad += dwn/sigma - (sum dwn*W)/sigma^2 * u v^T       (SURVEY.md section 8c identity)
 * dwn has W's layout (packed_taps = 0) or is the tap-major [taps][d0][d1] buffer of lb_wgrad_tc
 * (packed_taps = kh*kw).  dot_out: 2 doubles (receives sum dwn*W); stat_work: as lb_norm_stats (ordered grid sum).
 * w = NULL (packed_taps = 0 only): dot_out[0] already holds sum dwn*W (lb_wgrad_tc computes it as it writes dwn).
 * Trainable u / v -- the reference's training loop calls dis.requires_grad_(True) (main.py:172), which also switches
 * on the discriminator's weight_u / weight_v (requires_grad=False Parameters, spectral_norm.py:45-46); from the second
 * discriminator step on sigma = u.(W v) (spectral_norm.py:31) hands them gradients and Nadam moves them.  Pass
 * s_fwd (= W v of the forward pass this backward belongs to, from lb_sn_power_iter_batched's s_out), du and cacc (all
 * three or none): du += c * s_fwd, cacc[0] += c with c = dL/dsigma = -(sum dwn*W)/sigma^2; lb_sn_uv_grad_batched then
 * turns the accumulated c into dv. */
int lb_sn_weight_grad(const float* dwn, const float* w, const float* u, const float* v,
                      const float* sigma, float* grad, int height, int width, int packed_taps,
                      double* dot_out, double* stat_work, const float* s_fwd, float* du, float* cacc,
                      lb_stream_t stream);
/* dv += cacc[layer] * W^T u (live u) for every layer record with dv != NULL, then cacc[layer] = 0: the gradient of the
 * trainable weight_v (see lb_sn_weight_grad); run once after the backward passes of a step, before the optimizer. */
int lb_sn_uv_grad_batched(const void* layers_dev, int n_layers, const void* items1_dev, int n_items1, float* scratch,
                          float* cacc, lb_stream_t stream);

/* ---- convolution family as a gather-GEMM                              libs/conv.py:11-24, libs/attention.py:9-54,
 *                                                                       libs/scale.py:25-34, libs/linear.py:10
 * out[b,oy,ox,n] = alpha * sum_{ty,tx,k} in[b,iy,ix,k] * W(ty,tx,k,n)  (+ bias[n])
 *   mode 0 (strided):    iy = oy*stride - pad + ty            -> Conv fwd, ConvTranspose dgrad
 *   mode 1 (transposed): iy = (oy + pad - ty)/stride if exact -> ConvTranspose fwd, Conv dgrad
 * W(ty,tx,k,n) = w[k*w_sk + n*w_sn + ty*w_sty + tx*w_stx]: the master weight is read in place in
 * whatever layout the reference stores it ((Cout,Cin,kh,kw) or (Cin,Cout,kh,kw)).
 * alpha is a device scalar (1/sigma) or NULL for 1. */
typedef struct {
  int batch, in_h, in_w, in_c;     /* gathered operand */
  int out_h, out_w, out_c;         /* produced operand */
  int kh, kw, stride, pad, mode;
  int ld_in, ld_out;               /* channel strides of the two operands' rows */
  int64_t w_sk, w_sn, w_sty, w_stx;
} lb_conv_geom;

int lb_conv_gemm(const float* in, const float* w, const float* alpha, const float* bias, float* out,
                 const lb_conv_geom* g, lb_stream_t stream);
/* weight gradient: dw(ty,tx,kg,kd) += sum_m gathered[m@tap, kg] * dense[m, kd] with mode-0 gather
 * around the DENSE operand's pixel grid (dense = dy for Conv, x for ConvTranspose).
 * geom: in_* describes the gathered operand, out_* the dense one; w_sk strides the gathered
 * channel, w_sn the dense channel.  dw must be zeroed by the caller (atomic accumulation). */
int lb_conv_wgrad(const float* gathered, const float* dense, float* dw, const lb_conv_geom* g,
                  lb_stream_t stream);
/* Direct fp32 kernels for layers with a tiny channel count on one side (the discriminator stem 3->3 / 3->32 / 3->29 and
 * the generator's final 48->3, conv.py:14-20): the weight sits in shared memory, every activation is touched once, and
 * the neighbouring RootTanh is fused: growth_in > 0 applies RootTanh to the input on load (conv.py:23-24); growth_out > 0
 * multiplies the result by RootTanh'(xpre[pixel][n]) (activation.py:18-36) -- the input-gradient direction; growth_out = -1
 * multiplies by xpre[pixel][n] itself (the derivative the forward pass stored, LB_EX_OUT16_IS_DACT).  Geometry and
 * weight addressing are those of lb_conv_gemm / lb_conv_wgrad; growth_gathered applies RootTanh to the gathered operand. */
int lb_conv_small_supported(const lb_conv_geom* g);
int lb_conv_small(const void* in, const float* w, const float* alpha, const float* bias, void* out, const lb_conv_geom* g,
                  int growth_in, const void* xpre, int ld_xpre, int growth_out, int cat_input, int wide_dtype, lb_stream_t stream);
/* The side with <= 4 channels is always fp32; wide_dtype is the storage of the other side (out / xpre when in_c <= 4,
 * else in; for the weight gradient: dense when it is a 1x1 layer with in_c <= 4, else gathered). */
/* cat_input != 0 (1x1 stride-1 layers, no fused activation): `out` is the START of rows [in_c copied | out_c conv] with
 * row stride g->ld_out -- CatModule's concat (merge.py:10-16) written in the same pass. */
int lb_conv_small_wgrad_supported(const lb_conv_geom* g);
int lb_conv_small_wgrad(const void* gathered, const void* dense, float* dw, const lb_conv_geom* g, int growth_gathered,
                        int wide_dtype, lb_stream_t stream);
/* ---- GPU-side input pipeline (SURVEY "next" row N4)                                  libs/utils.py:92-113, main.py:122-129
 * RandomHorizontalFlip + ColorJitter(brightness, contrast, saturation) + RandomResizedCrop (antialiased bilinear) +
 * ToTensor + Normalize(0.5, 0.5) for a batch of decoded uint8 images [B][Hs][Ws][3] resident in HBM (already resized
 * to 2 x image size, as transforms.Resize does).  The random draws are the host's (torchvision's get_params), per sample
 * 12 floats: crop top, left, height, width, flip, brightness / contrast / saturation factors, three operation codes in
 * application order (0 brightness, 1 contrast, 2 saturation, -1 none), one spare.  torchvision's float-tensor arithmetic.
 * mean_work: batch floats.  dst: fp32 channels-last [B][size][size][3] in [-1, 1]. */
int lb_augment(const void* src_u8, const float* params, float* mean_work, float* dst, int batch, int src_h, int src_w,
               int size, lb_stream_t stream);

/* ---- SEPARABLE = True (config.py:53): the grouped convolutions of that configuration                libs/conv.py:17, libs/attention.py:15-21
 * Depthwise k x k conv / transposed conv (groups = channels): out[b,oy,ox,c] = alpha * sum_taps in[b,iy,ix,c] * w[c][ty][tx]
 * with the mode conventions above (forward of one = input gradient of the other, same weights); w is the fp32 master
 * weight (C,1,kh,kw); in / out channels-last in storage `dtype`.  One multiply-add per byte: HBM-bound direct kernels. */
int lb_dw_conv(const void* in, const float* w, const float* alpha, void* out, int batch, int in_h, int in_w, int out_h,
               int out_w, int channels, int kh, int kw, int stride, int pad, int mode, int dtype, lb_stream_t stream);
/* dw[c][ty][tx] += sum_pixels gathered[pixel@tap][c] * dense[pixel][c] (mode-0 gather around the dense grid: gathered = x,
 * dense = dy for a conv; gathered = dy, dense = x for a transposed conv). */
int lb_dw_wgrad(const void* gathered, const void* dense, float* dw, int batch, int g_h, int g_w, int d_h, int d_w,
                int channels, int kh, int kw, int stride, int pad, int dtype, lb_stream_t stream);
/* Feature attention's grouped full-extent conv: in [B][P][F] -> out [B][F/r], out[b,o] = alpha * sum_{j<r,p} in[b,p,o*r+j] * w[o][j][p]
 * (w fp32 (F/r, r, S, S)); its input gradient din[b,p,c] = alpha * g[b,c/r] * w[c][p]; its weight gradient
 * dw[c][p] += sum_b in[b,p,c] * g[b,c/r] (fixed order over the batch). */
int lb_gfull_fwd(const void* in, const float* w, const float* alpha, void* out, int batch, int pixels, int features,
                 int group_in, int dtype, lb_stream_t stream);
int lb_gfull_dgrad(const void* g, const float* w, const float* alpha, void* din, int batch, int pixels, int features,
                   int group_in, int dtype, lb_stream_t stream);
int lb_gfull_wgrad(const void* in, const void* g, float* dw, int batch, int pixels, int features, int group_in, int dtype,
                   lb_stream_t stream);

/* ---- the same GEMM on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulator, TMA tiles), bf16 operands,
 * fp32 accumulate/output.  `in` is the bf16 channels-last activation, `w_packed` the weight packed by
 * lb_conv_tc_pack as [tap][n][k] bf16 (lb_conv_tc_packed_elems elements).  lb_conv_tc_supported says
 * whether a geometry is covered (channels % 8, stride 1|2, <= 32 taps); everything else stays on
 * lb_conv_gemm. */
int lb_conv_tc_supported(const lb_conv_geom* g);
size_t lb_conv_tc_packed_elems(const lb_conv_geom* g);
int lb_conv_tc_pack(const float* w, void* packed, const lb_conv_geom* g, lb_stream_t stream);
/* Every weight pack of a model in one launch (the optimizer moves all weights at once, so all packs go stale at once; the
 * reference re-reads its fp32 weights in every conv call, conv.py:14-20, and has no counterpart).  The host fills one
 * opaque record of lb_pack_rec_bytes() bytes per (weight, direction) with lb_pack_rec_fill (returns the record's item
 * count, < 0 on error), uploads the records and a list of int pairs (record index, first item) covering every record in
 * chunks of lb_pack_chunk_items() items, and calls lb_conv_tc_pack_batched whenever the weights have moved. */
int lb_pack_rec_bytes(void);
int lb_pack_chunk_items(void);
int lb_pack_rec_fill(const float* w, void* packed, const lb_conv_geom* g, void* rec_host);
int lb_conv_tc_pack_batched(const void* recs_dev, const void* chunks_dev, int n_chunks, lb_stream_t stream);
int lb_conv_tc_gemm(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, float* out,
                    const lb_conv_geom* g, lb_stream_t stream);
/* Same with a caller-provided workspace of lb_conv_tc_workspace_bytes(g) bytes (0 for most shapes): weight-bound layers
 * whose output tiling cannot fill the GPU (5x5 1024->1024 on a 2x2 map, style linears at small batch) split the
 * reduction over taps x channel chunks; every split writes its own partial tile there and a second kernel adds the
 * splits in index order -- no floating-point atomics, so the forward pass is bit-reproducible.  Without a workspace
 * (lb_conv_tc_gemm) such layers run unsplit. */
size_t lb_conv_tc_workspace_bytes(const lb_conv_geom* g);
/* out_dtype: LB_F32 or LB_BF16 rows of `out` (bf16 = the activation storage of the tensor-core configuration) */
int lb_conv_tc_gemm_ws(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, void* out,
                       const lb_conv_geom* g, void* work, size_t work_bytes, int out_dtype, lb_stream_t stream);
/* Persistent variant with a fused epilogue (TMEM double buffering, TMA-stored outputs):
 *   acc = alpha * GEMM (+ bias);  if aux: acc *= RootTanh'(aux[pixel][n]) (activation.py:18-36, growth 4; aux has the
 *   geometry of the output, row stride ld_aux, storage aux_dtype);
 *   out32 (fp32, row stride g->ld_out), out16 (bf16 of acc) and / or out16a (bf16 of RootTanh(acc), the operand of the
 *   convolution that follows, conv.py:23-24), both bf16 outputs with row stride ld_out16.  Any subset, not none.
 *   flags: LB_EX_OUT16_IS_DACT -- out16 receives RootTanh'(acc) instead of acc (needs out16 and out16a): the forward
 *   pass stores the activation's derivative, computed with RootTanh from shared intermediates, and the backward pass
 *   passes it back as aux with LB_EX_AUX_IS_FACTOR -- acc *= aux[pixel][n], no transcendental in the backward epilogue.
 * Returns LB_EUNSUPPORTED when the geometry / alignment is outside the kernel (lb_conv_tc_ex_supported tells in
 * advance: out32_used 0/1, ld_out16 = 0 / ld_aux = 0 for "not used"); callers then use lb_conv_tc_gemm_ws + elementwise
 * kernels. */
int lb_conv_tc_ex_supported(const lb_conv_geom* g, int out32_used, int ld_out16, int ld_aux, int aux_dtype);
enum { LB_EX_AUX_IS_FACTOR = 1, LB_EX_OUT16_IS_DACT = 2 };
int lb_conv_tc_gemm_ex(const void* in_bf16, const void* w_packed, const float* alpha, const float* bias, float* out32,
                       void* out16, void* out16a, int ld_out16, const void* aux, int ld_aux, int aux_dtype, int flags,
                       const lb_conv_geom* g, lb_stream_t stream);
/* weight gradient on the tensor cores: dwn[n][m][tap] = sum_pixels gathered[pixel@tap][m] * dense[pixel][n]
 * (geometry as lb_conv_wgrad: in_* = gathered, out_* = dense; both operands bf16 channels-last).  dwn is fp32 in the
 * master weight's own layout (Conv2d [Cout][Cin][kh][kw] with gathered = x; ConvTranspose2d [Cin][Cout][kh][kw] with
 * gathered = dy) and is OVERWRITTEN: feed it to lb_sn_weight_grad with packed_taps = 0.  Split-K over the pixel axis is
 * deterministic: every split stores its own partial in `work` (lb_wgrad_tc_workspace_floats(g) floats, contents
 * irrelevant on entry) and a second kernel adds the splits in index order -- no floating-point atomics. */
int lb_wgrad_tc_supported(const lb_conv_geom* g);
size_t lb_wgrad_tc_workspace_floats(const lb_conv_geom* g);
/* w / dot_out / stat_work (all three or none): the reduction pass also leaves dot_out[0] = sum dwn * w (w = the master
 * weight; fixed-order grid sum through stat_work, lb_stat_work_doubles() zeroed doubles) -- lb_sn_weight_grad called with
 * w = NULL then takes dot_out as given instead of making its own pass over dwn and w. */
int lb_wgrad_tc(const void* gathered_bf16, const void* dense_bf16, float* dwn, const lb_conv_geom* g, float* work,
                size_t work_floats, const float* w, double* dot_out, double* stat_work, lb_stream_t stream);
/* fp32 -> bf16 producers of GEMM operands: plain cast, and RootTanh fused with the cast (activation.py:9-16) */
int lb_cast_bf16(const float* x, void* y, size_t n, lb_stream_t stream);
/* row-strided variant: dst[r*ld_dst + c] = bf16(f(src[r*ld_src + c])); f = identity (growth 0) or RootTanh (growth >= 1) */
int lb_cast_bf16_rows(const float* src, int ld_src, void* dst, int ld_dst, int64_t rows, int cols, int growth,
                      lb_stream_t stream);
int lb_roottanh_fwd_bf16(const float* x, void* y, size_t n, int growth, lb_stream_t stream);

/* column sum: out[c] += sum_rows x[row*ld + c]  (bias gradients; nn.Conv2d's bias of libs/spectral_norm.py-wrapped layers);
 * x in storage `dtype`, out fp32.  x may be a column slice of ld-wide rows (a concat output): when ld is a multiple of
 * the 16-byte vector width the kernel reads the slice's 16-byte aligned superset inside each row (up to 15 bytes before
 * x and after x + cols of the SAME row) and discards the extra columns. */
int lb_colsum(const void* x, int64_t rows, int cols, int ld, float* out, int dtype, lb_stream_t stream);

/* ---- softmax                                                          libs/attention.py:35,47 */
/* over pixels for every (b,c) of a channels-last [B][P][C] tensor (SelfAttention, dim=-1 of [B,F,HW]); storage `dtype`.
 * Long rows at small batch (HW = 65 536 at 256x256) are split over CTAs through `work`
 * (lb_softmax_pixels_work_floats floats; 0 = not needed; NULL = run unsplit): per-split (max, sum exp) partials combined
 * in a fixed order. */
size_t lb_softmax_pixels_work_floats(int batch, int pixels, int channels);
int lb_softmax_pixels_fwd(const void* x, void* y, int batch, int pixels, int channels, float* work, size_t work_floats,
                          int dtype, lb_stream_t stream);
int lb_softmax_pixels_bwd(const void* y, const void* g, void* dx, int batch, int pixels, int channels, float* work,
                          size_t work_floats, int dtype, lb_stream_t stream);
/* over the contiguous last axis of [rows][cols] (feature attention, Softmax(dim=1) on [B,F,1,1]) */
int lb_softmax_rows_fwd(const void* x, void* y, int rows, int cols, int dtype, lb_stream_t stream);
int lb_softmax_rows_bwd(const void* y, const void* g, void* dx, int rows, int cols, int dtype, lb_stream_t stream);

/* ---- skip path resampling                                             libs/scale.py:7-45, libs/merge.py:4-16 */
/* FeaturePooling: flat NCHW-memory regrouping mean (scale.py:12-16) evaluated on channels-last data; storage `dtype` */
int lb_featpool_fwd(const void* x, void* y, int batch, int h, int w, int c_in, int c_out, int dtype, lb_stream_t stream);
int lb_featpool_bwd(const void* g, void* dx, int batch, int h, int w, int c_in, int c_out, int dtype, lb_stream_t stream);
/* bilinear x2, align_corners = False (scale.py:37-38) */
int lb_upsample2x_fwd(const void* x, void* y, int batch, int h, int w, int c, int dtype, lb_stream_t stream);
int lb_upsample2x_bwd(const void* g, void* dx, int batch, int h, int w, int c, int dtype, lb_stream_t stream);
/* AvgPool 2x2 stride 2 (scale.py:40); h,w are the INPUT sizes (even) */
int lb_avgpool2_fwd(const void* x, void* y, int batch, int h, int w, int c, int dtype, lb_stream_t stream);
int lb_avgpool2_bwd(const void* g, void* dx, int batch, int h, int w, int c, int dtype, lb_stream_t stream);
/* strided row copy dst[row*ld_dst + c] (=|+=) src[row*ld_src + c]: channel concat / slice (merge.py:15) */
int lb_copy_rows(const void* src, int ld_src, void* dst, int ld_dst, int64_t rows, int cols, int accumulate, int src_dtype,
                 int dst_dtype, lb_stream_t stream);
/* layout change at the model boundary: fp32 NCHW (what the reference's callers hold) <-> channels-last in storage `dtype` */
int lb_nchw_to_nhwc(const float* x, void* y, int batch, int c, int hw, int out_dtype, lb_stream_t stream);
int lb_nhwc_to_nchw(const void* x, float* y, int batch, int c, int hw, int in_dtype, lb_stream_t stream);

/* ---- losses                                                           libs/utils.py:133-134, libs/grad_penalty.py:1-2, main.py:149-156,616-621
 * D step: loss = mean(hinge(d_true) + hinge(-d_fake)) + gamma*(mean(d_true) - mean(d_aug))^2.
 * means[2] = {sum d_true, sum d_aug} over the GLOBAL batch n_global (lb_loss_sums computes the local
 * sums; data parallel all-reduces them).  out[3] = {hinge part (local mean), penalty, 0};
 * grads are dL/d(d_true), dL/d(d_fake), dL/d(d_aug) for a loss averaged over n_global. */
int lb_loss_sums(const float* d_true, const float* d_aug, int n_local, double* sums, lb_stream_t stream);
int lb_d_loss(const float* d_true, const float* d_fake, const float* d_aug, const double* sums, int n_local,
              double n_global, float gamma, float* out, float* g_true, float* g_fake, float* g_aug,
              lb_stream_t stream);
/* G step: loss = mean(hinge(d_fake)); out[1]; grad dL/d(d_fake) */
int lb_g_loss(const float* d_fake, int n_local, double n_global, float* out, float* g_fake, lb_stream_t stream);

/* ---- Nadam over a flat parameter arena                                libs/nadam.py:56-87
 * hyper (device float[3]) = {c_grad = lr*(1-mu_t)/(1-m_schedule_new), c_mom = lr*mu_{t+1}/(1-m_schedule_next),
 * 1/(1 - beta2^t)}, produced by lb_nadam_schedule. */
int lb_nadam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                  float beta1, float beta2, float eps, const float* hyper, lb_stream_t stream);
/* advances the device-resident schedule state = {t, m_schedule} (double[2], start {0, 1}) by one step and writes
 * hyper = {c_grad, c_mom, 1/bias2} (float[3]) -- on the device so a captured CUDA graph of the step replays. */
int lb_nadam_schedule(double* state, float* hyper, double lr, double beta1, double beta2, double schedule_decay,
                      lb_stream_t stream);
/* fill / scale helpers used around the step */
int lb_fill(float* x, size_t n, float value, lb_stream_t stream);
int lb_scale(float* x, size_t n, float factor, lb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LOCATE_B200_H */
