// Hinge + real/augmented consistency-penalty losses and their gradients
// (libs/utils.py:133-134, libs/grad_penalty.py:1-2, main.py:149-156,616-621).  [B] scalars: one CTA.
#include "common.cuh"

__global__ void __launch_bounds__(256) k_loss_sums(const float* __restrict__ d_true, const float* __restrict__ d_aug, int n,
                                                  double* __restrict__ sums) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { a += d_true[i]; b += d_aug[i]; }
  a = lb_block_sum(a, scratch);
  b = lb_block_sum(b, scratch);
  if (threadIdx.x == 0) { sums[0] = a; sums[1] = b; }
}
extern "C" int lb_loss_sums(const float* d_true, const float* d_aug, int n_local, double* sums, lb_stream_t s) {
  LB_REQUIRE(d_true && d_aug && sums && n_local > 0);
  lb_launch(k_loss_sums, 1, 256, 0, lb_s(s), d_true, d_aug, n_local, sums);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

// loss = (1/N) sum_i [max(0, 1 - t_i) + max(0, 1 + f_i)] + gamma * (mean t - mean a)^2
//   d/dt_i = (-[t_i < 1] + 2 gamma diff) / N ; d/df_i = [f_i > -1] / N ; d/da_i = -2 gamma diff / N
__global__ void __launch_bounds__(256) k_d_loss(const float* __restrict__ t, const float* __restrict__ f, const float* __restrict__ a,
                                               const double* __restrict__ sums, int n, double n_global, float gamma,
                                               float* __restrict__ out, float* __restrict__ gt, float* __restrict__ gf,
                                               float* __restrict__ ga) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  const double diff = (sums[0] - sums[1]) / n_global;
  const float inv_n = (float)(1.0 / n_global);
  const float pen_g = (float)(2.0 * gamma * diff) * inv_n;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float tv = t[i], fv = f[i];
    const float ht = 1.0f - tv, hf = 1.0f + fv;         // hinge(d_true), hinge(-d_fake)
    acc += (double)(fmaxf(ht, 0.0f) + fmaxf(hf, 0.0f));
    gt[i] = (ht > 0.0f ? -inv_n : 0.0f) + pen_g;
    gf[i] = (hf > 0.0f ? inv_n : 0.0f);
    ga[i] = -pen_g;
  }
  acc = lb_block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    out[0] = (float)(acc / n_global);
    out[1] = (float)(gamma * diff * diff);
    out[2] = 0.0f;
  }
}
extern "C" int lb_d_loss(const float* d_true, const float* d_fake, const float* d_aug, const double* sums, int n_local,
                         double n_global, float gamma, float* out, float* g_true, float* g_fake, float* g_aug, lb_stream_t s) {
  LB_REQUIRE(d_true && d_fake && d_aug && sums && out && g_true && g_fake && g_aug && n_local > 0 && n_global >= n_local);
  lb_launch(k_d_loss, 1, 256, 0, lb_s(s), d_true, d_fake, d_aug, sums, n_local, n_global, gamma, out, g_true, g_fake, g_aug);
  LB_LAUNCH_CHECK();
  return LB_OK;
}

__global__ void __launch_bounds__(256) k_g_loss(const float* __restrict__ f, int n, double n_global, float* __restrict__ out,
                                               float* __restrict__ gf) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  const float inv_n = (float)(1.0 / n_global);
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float h = 1.0f - f[i];
    acc += (double)fmaxf(h, 0.0f);
    gf[i] = h > 0.0f ? -inv_n : 0.0f;
  }
  acc = lb_block_sum(acc, scratch);
  if (threadIdx.x == 0) out[0] = (float)(acc / n_global);
}
extern "C" int lb_g_loss(const float* d_fake, int n_local, double n_global, float* out, float* g_fake, lb_stream_t s) {
  LB_REQUIRE(d_fake && out && g_fake && n_local > 0 && n_global >= n_local);
  lb_launch(k_g_loss, 1, 256, 0, lb_s(s), d_fake, n_local, n_global, out, g_fake);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
