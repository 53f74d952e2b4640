// GPU-side input pipeline (SURVEY.md "next" row N4): what the reference's loaders do on the host with PIL / torchvision
// (libs/utils.py:92-113) -- RandomHorizontalFlip, ColorJitter(brightness, contrast, saturation), RandomResizedCrop
// (square crop of 75-100 % of the area, bilinear resize with antialiasing), ToTensor, Normalize(0.5, 0.5) -- for a whole
// batch of decoded uint8 images already resident in HBM, in two launches.  The random draws stay on the host
// (locate_b200/augment.py mirrors torchvision's get_params); per sample they arrive as 12 floats:
//   [0] crop top  [1] crop left  [2] crop height  [3] crop width  [4] flip (0/1)
//   [5] brightness factor  [6] contrast factor  [7] saturation factor  [8..10] operation order (0 = brightness,
//   1 = contrast, 2 = saturation, -1 = none)  [11] unused
// Arithmetic follows torchvision's float-tensor path: brightness x*f, contrast f*x + (1-f)*mean(gray(x)), saturation
// f*x + (1-f)*gray(x), each clamped to [0,1]; gray = 0.2989 r + 0.587 g + 0.114 b; resize = ATen's antialiased
// bilinear (triangle filter widened by the scale).  No intermediate rounding to uint8 (PIL's path rounds after each op).
#include "common.cuh"

namespace {

constexpr int kAugParams = 12;

__device__ __forceinline__ float aug_gray(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }
__device__ __forceinline__ float aug_clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// applies the operations order[0..upto) to one pixel
__device__ __forceinline__ void aug_jitter(float& r, float& g, float& b, const float* __restrict__ prm, float mean, int upto, int stop_at) {
  for (int i = 0; i < upto; ++i) {
    const int op = (int)prm[8 + i];
    if (op == stop_at) return;
    if (op == 0) {
      const float f = prm[5];
      r = aug_clamp01(r * f); g = aug_clamp01(g * f); b = aug_clamp01(b * f);
    } else if (op == 1) {
      const float f = prm[6], m = (1.0f - f) * mean;
      r = aug_clamp01(fmaf(f, r, m)); g = aug_clamp01(fmaf(f, g, m)); b = aug_clamp01(fmaf(f, b, m));
    } else if (op == 2) {
      const float f = prm[7], m = (1.0f - f) * aug_gray(r, g, b);
      r = aug_clamp01(fmaf(f, r, m)); g = aug_clamp01(fmaf(f, g, m)); b = aug_clamp01(fmaf(f, b, m));
    }
  }
}

// mean over the whole source image of gray(image after the operations that precede the contrast adjustment)
// one CTA per image, fixed-order block reduction (deterministic)
__global__ void __launch_bounds__(512) k_aug_gray_mean(const uint8_t* __restrict__ src, const float* __restrict__ params, float* __restrict__ mean,
                                                      int hs, int ws) {
  lb_pdl_enter();
  __shared__ double scratch[32];
  const int b = blockIdx.x;
  const float* prm = params + (size_t)b * kAugParams;
  const uint8_t* img = src + (size_t)b * hs * ws * 3;
  const int n = hs * ws;
  double acc = 0.0;
  float part = 0.0f;
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float r = img[3 * i] * (1.0f / 255.0f), g = img[3 * i + 1] * (1.0f / 255.0f), bl = img[3 * i + 2] * (1.0f / 255.0f);
    aug_jitter(r, g, bl, prm, 0.0f, 3, 1);                 // stops at the contrast operation
    part += aug_gray(r, g, bl);
    if (++cnt == 32) { acc += (double)part; part = 0.0f; cnt = 0; }
  }
  acc += (double)part;
  acc = lb_block_sum(acc, scratch);
  if (threadIdx.x == 0) mean[b] = (float)(acc / (double)n);
}

// ATen's antialiased bilinear footprint along one axis: first source index, tap count (<= kMaxTaps), normalised weights
constexpr int kMaxTaps = 12;
__device__ __forceinline__ void aug_axis(float crop0, float crop_len, int out_i, int out_n, int& first, int& count, float* wts) {
  const float scale = crop_len / (float)out_n;
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const float center = scale * ((float)out_i + 0.5f);
  int lo = (int)(center - support + 0.5f);
  if (lo < 0) lo = 0;
  int hi = (int)(center + support + 0.5f);
  if (hi > (int)crop_len) hi = (int)crop_len;
  count = hi - lo;
  if (count > kMaxTaps) count = kMaxTaps;
  float total = 0.0f;
  for (int j = 0; j < count; ++j) {
    const float x = ((float)(j + lo) - center + 0.5f) * invscale;
    const float w = fmaxf(0.0f, 1.0f - fabsf(x));
    wts[j] = w;
    total += w;
  }
  const float inv = total != 0.0f ? 1.0f / total : 0.0f;
  for (int j = 0; j < count; ++j) wts[j] *= inv;
  first = (int)crop0 + lo;
}

// one thread per output pixel: dst is channels-last fp32 [B][S][S][3]
__global__ void __launch_bounds__(256) k_aug_apply(const uint8_t* __restrict__ src, const float* __restrict__ params, const float* __restrict__ mean,
                                                  float* __restrict__ dst, int batch, int hs, int ws, int size) {
  lb_pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * size * size) return;
  const int x = i % size, y = (i / size) % size, b = i / (size * size);
  const float* prm = params + (size_t)b * kAugParams;
  const uint8_t* img = src + (size_t)b * hs * ws * 3;
  float wy[kMaxTaps], wx[kMaxTaps];
  int y0, ny, x0, nx;
  aug_axis(prm[0], prm[2], y, size, y0, ny, wy);
  aug_axis(prm[1], prm[3], x, size, x0, nx, wx);
  const bool flip = prm[4] != 0.0f;
  const float m = mean[b];
  float ar = 0.0f, ag = 0.0f, ab = 0.0f;
  for (int jy = 0; jy < ny; ++jy) {
    const int sy = min(y0 + jy, hs - 1);
    float rr = 0.0f, rg = 0.0f, rb = 0.0f;
    for (int jx = 0; jx < nx; ++jx) {
      int sx = min(x0 + jx, ws - 1);
      if (flip) sx = ws - 1 - sx;                          // the flip precedes the crop: crop coordinates address the flipped image
      const uint8_t* px = img + ((size_t)sy * ws + sx) * 3;
      float r = px[0] * (1.0f / 255.0f), g = px[1] * (1.0f / 255.0f), bl = px[2] * (1.0f / 255.0f);
      aug_jitter(r, g, bl, prm, m, 3, -2);
      rr = fmaf(wx[jx], r, rr); rg = fmaf(wx[jx], g, rg); rb = fmaf(wx[jx], bl, rb);
    }
    ar = fmaf(wy[jy], rr, ar); ag = fmaf(wy[jy], rg, ag); ab = fmaf(wy[jy], rb, ab);
  }
  float* o = dst + (size_t)i * 3;
  o[0] = (ar - 0.5f) * 2.0f; o[1] = (ag - 0.5f) * 2.0f; o[2] = (ab - 0.5f) * 2.0f;      // Normalize((0.5,)*3, (0.5,)*3)
}

}  // namespace

// src: uint8 [B][Hs][Ws][3] (decoded, already resized to 2 x image size by the host as transforms.Resize does);
// params: fp32 [B][12] (see the top of this file); mean_work: B floats; dst: fp32 channels-last [B][size][size][3].
extern "C" int lb_augment(const void* src_u8, const float* params, float* mean_work, float* dst, int batch, int src_h, int src_w,
                          int size, lb_stream_t s) {
  LB_REQUIRE(src_u8 && params && mean_work && dst && batch > 0 && src_h > 0 && src_w > 0 && size > 0);
  LB_REQUIRE((long long)batch * size * size < (1ll << 31) && (long long)src_h * src_w < (1ll << 30));
  lb_launch(k_aug_gray_mean, batch, 512, 0, lb_s(s), reinterpret_cast<const uint8_t*>(src_u8), params, mean_work, src_h, src_w);
  LB_LAUNCH_CHECK();
  const int n = batch * size * size;
  lb_launch(k_aug_apply, (n + 255) / 256, 256, 0, lb_s(s), reinterpret_cast<const uint8_t*>(src_u8), params, mean_work, dst, batch, src_h, src_w,
                                                    size);
  LB_LAUNCH_CHECK();
  return LB_OK;
}
