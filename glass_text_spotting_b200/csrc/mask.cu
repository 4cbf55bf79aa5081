// mask.cu -- the two mask-branch kernels that are not convolutions (sm_100a).
//
// The mask head itself (detectron2 MaskRCNNConvUpsampleHead as configured by glass_finetune_totaltext.yaml: 4 x conv3x3
// + ReLU, 2x2/s2 deconv + ReLU, 1x1 predictor) runs on glass_conv_gemm: the deconv is a GEMM with N = 4 x 256 (one
// 256-block per output sub-pixel (dy, dx)) and the predictor a block-diagonal GEMM over those 1024 columns, so a pooled
// pixel's row ends up holding its 2 x 2 output logits.  Here:
//   * mask_finalize: mask_rcnn_inference (sigmoid) + the sub-pixel scatter -> pred_masks [K, 28, 28];
//   * paste_masks_rotated: the reference's paste_masks_in_image / _do_paste_mask for rotated boxes
//     (glass/postprocess/post_processor_academic.py:187-335): every image pixel is rotated into the box frame,
//     normalised to [-1, 1], and the 28 x 28 mask is sampled like F.grid_sample(bilinear, zeros, align_corners=False),
//     then thresholded -- one pass over [K, H, W], no [K, H, W, 2] grid tensor and no per-box Python loop.
#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"

namespace glass {

__global__ void mask_finalize_kernel(const float* __restrict__ logits, int ld, int k_words, int h, int w,
                                     float* __restrict__ masks) {
  const int hp = h + 2, wp = w + 2, oh = 2 * h, ow = 2 * w;
  const int64_t total = (int64_t)k_words * oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow), oy = (int)((i / ow) % oh), k = (int)(i / ((int64_t)ow * oh));
    const int x = ox >> 1, dx = ox & 1, y = oy >> 1, dy = oy & 1;
    const int64_t row = ((int64_t)k * hp + y + 1) * wp + x + 1;
    const float v = logits[row * ld + dy * 2 + dx];
    masks[i] = 1.0f / (1.0f + expf(-v));
  }
}

struct PasteParams {
  const float* masks;
  const float* boxes;
  int k, m, img_h, img_w;
  float threshold;
  uint8_t* out;
  float* out_soft;
};

__device__ __forceinline__ float paste_sample(const float* __restrict__ mk, int m, float cx, float cy, float cs, float sn,
                                              float x0, float x1, float y0, float y1, int X, int Y) {
  // the reference's fp32 sequence: centre, rotate by rot = [[cos, sin], [-sin, cos]] (row vector x matrix), recentre,
  // normalise with the UN-rotated extents, un-normalise for grid_sample(align_corners=False)
  const float gx = ((float)X + 0.5f) - cx, gy = ((float)Y + 0.5f) - cy;
  float rx = gx * cs + gy * (-sn);
  float ry = gx * sn + gy * cs;
  rx += cx;
  ry += cy;
  const float u = (rx - x0) / (x1 - x0) * 2.f - 1.f;
  const float v = (ry - y0) / (y1 - y0) * 2.f - 1.f;
  const float ix = ((u + 1.f) * (float)m - 1.f) / 2.f;
  const float iy = ((v + 1.f) * (float)m - 1.f) / 2.f;
  if (!(ix > -1.f && ix < (float)m && iy > -1.f && iy < (float)m)) return 0.f;  // all four taps outside (also NaN)
  const float fx = floorf(ix), fy = floorf(iy);
  const int ix0 = (int)fx, iy0 = (int)fy, ix1 = ix0 + 1, iy1 = iy0 + 1;
  const float w_nw = ((float)ix1 - ix) * ((float)iy1 - iy), w_ne = (ix - fx) * ((float)iy1 - iy);
  const float w_sw = ((float)ix1 - ix) * (iy - fy), w_se = (ix - fx) * (iy - fy);
  float acc = 0.f;
  const bool xin0 = ix0 >= 0 && ix0 < m, xin1 = ix1 >= 0 && ix1 < m, yin0 = iy0 >= 0 && iy0 < m, yin1 = iy1 >= 0 && iy1 < m;
  if (yin0 && xin0) acc += mk[iy0 * m + ix0] * w_nw;
  if (yin0 && xin1) acc += mk[iy0 * m + ix1] * w_ne;
  if (yin1 && xin0) acc += mk[iy1 * m + ix0] * w_sw;
  if (yin1 && xin1) acc += mk[iy1 * m + ix1] * w_se;
  return acc;
}

template <int VEC>  // pixels per thread along X (4: one 32-bit store)
__global__ void __launch_bounds__(256) paste_masks_rotated_kernel(const PasteParams p) {
  const int wv = p.img_w / VEC;
  const int64_t total = (int64_t)p.k * p.img_h * wv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xg = (int)(i % wv), Y = (int)((i / wv) % p.img_h), k = (int)(i / ((int64_t)wv * p.img_h));
    const float* b = p.boxes + (int64_t)k * 5;
    const float cx = b[0], cy = b[1], bw = b[2], bh = b[3];
    const float a = b[4] * 0.017453292519943295f;  // torch.deg2rad
    const float cs = (float)cos((double)a), sn = (float)sin((double)a);
    const float x0 = cx - bw / 2.f, x1 = cx + bw / 2.f, y0 = cy - bh / 2.f, y1 = cy + bh / 2.f;
    const float* mk = p.masks + (int64_t)k * p.m * p.m;
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int X = xg * VEC + e;
      const float s = paste_sample(mk, p.m, cx, cy, cs, sn, x0, x1, y0, y1, X, Y);
      const int64_t o = ((int64_t)k * p.img_h + Y) * p.img_w + X;
      if (p.out_soft) p.out_soft[o] = s;
      if (VEC == 1) p.out[o] = s >= p.threshold ? 1 : 0;
      else packed |= (s >= p.threshold ? 1u : 0u) << (8 * e);
    }
    if (VEC == 4) *reinterpret_cast<uint32_t*>(p.out + ((int64_t)k * p.img_h + Y) * p.img_w + xg * 4) = packed;
  }
}

}  // namespace glass

using namespace glass;

extern "C" int glass_mask_finalize(const float* logits, int ld, int k_words, int h, int w, float* masks, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(logits && masks, "null pointer");
  GLASS_CHECK(ld >= 4 && k_words >= 0 && h > 0 && w > 0, "bad shape");
  if (k_words == 0) return 0;
  const int64_t total = (int64_t)k_words * 4 * h * w;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)num_sms() * 32) blocks = (int64_t)num_sms() * 32;
  mask_finalize_kernel<<<(int)blocks, 256, 0, stream>>>(logits, ld, k_words, h, w, masks);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_paste_masks_rotated(const float* masks, const float* boxes, int k, int m, int img_h, int img_w,
                                         float threshold, uint8_t* out, float* out_soft, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(masks && boxes && out, "null pointer");
  GLASS_CHECK(k >= 0 && m > 0 && img_h > 0 && img_w > 0, "bad shape");
  if (k == 0) return 0;
  PasteParams p{masks, boxes, k, m, img_h, img_w, threshold, out, out_soft};
  const bool vec4 = img_w % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  const int64_t total = (int64_t)k * img_h * (vec4 ? img_w / 4 : img_w);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)num_sms() * 64) blocks = (int64_t)num_sms() * 64;
  if (vec4) paste_masks_rotated_kernel<4><<<(int)blocks, 256, 0, stream>>>(p);
  else paste_masks_rotated_kernel<1><<<(int)blocks, 256, 0, stream>>>(p);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
