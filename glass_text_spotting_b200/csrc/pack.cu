// pack.cu -- device-side glue between the detector and the recognizer / the gather, and weight pre-packing.
//   glass_pack_rois        : per-image detections + counts -> the recognizer's RoI list, word offsets and the live word
//                            count, all left ON THE DEVICE (the recognizer's kernels read the count themselves), which
//                            removes the one host sync of the step (recognizers_hybrid_head.py:176-181 runs the
//                            recognizer on the detected boxes; the reference sizes it on the host).
//   glass_pack_detections  : the fixed-size per-image record of the end-of-loop all-gather (replaces the
//                            pickled comm.gather of glass/evaluation/text_evaluator.py:246-249).
//   glass_prepack_weights  : fp32 weight matrix -> K-major split-fp16 hi/lo planes with a per-row power-of-two pre-scale
//                            (SURVEY.md 8b "pre-packed weights ... produced once into caller-owned buffers").
#include <math_constants.h>

#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"

namespace glass {

__global__ void __launch_bounds__(256) pack_rois_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ counts,
                                                        int n_img, int max_det, float* __restrict__ rois,
                                                        int32_t* __restrict__ word_start, int32_t* __restrict__ total) {
  pdl_prologue();
  extern __shared__ int32_t s_start[];   // [n_img + 1]
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < n_img; ++i) {
      s_start[i] = acc;
      acc += min(max(counts[i], 0), max_det);
    }
    s_start[n_img] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= n_img; i += blockDim.x) word_start[i] = s_start[i];
  if (threadIdx.x == 0) *total = s_start[n_img];
  const int cap = n_img * max_det;
  for (int r = threadIdx.x; r < cap; r += blockDim.x) {
    // row r of the compacted list belongs to the image whose [start, next start) range holds it
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < s_start[n_img]) {
      int img = 0;
      while (s_start[img + 1] <= r) ++img;
      const float* b = boxes + ((int64_t)img * max_det + (r - s_start[img])) * 5;
      v[0] = (float)img;
#pragma unroll
      for (int c = 0; c < 5; ++c) v[1 + c] = b[c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) rois[(int64_t)r * 6 + c] = v[c];
  }
}

// one CTA per (image, detection slot): (valid, box 5, score, class, orientation 2, text probabilities) or zeros
__global__ void __launch_bounds__(128) pack_detections_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                              const float* __restrict__ orient,
                                                              const int32_t* __restrict__ counts,
                                                              const float* __restrict__ probs,
                                                              const int32_t* __restrict__ word_start, int max_det, int tp,
                                                              float* __restrict__ rec) {
  pdl_prologue();
  const int img = blockIdx.x / max_det, j = blockIdx.x - img * max_det;
  const int64_t row = (int64_t)blockIdx.x * (10 + tp);
  const bool live = j < min(max(counts[img], 0), max_det);
  if (threadIdx.x < 10) {
    float v = 0.f;
    if (live) {
      const int64_t d = (int64_t)img * max_det + j;
      const int c = threadIdx.x;
      if (c == 0) v = 1.f;
      else if (c <= 5) v = boxes[d * 5 + c - 1];
      else if (c == 6) v = scores[d];
      else if (c == 7) v = 0.f;                                  // pred_classes: one foreground class
      else v = orient ? orient[d * 2 + c - 8] : 0.f;
    }
    rec[row + threadIdx.x] = v;
  }
  const float* src = live ? probs + (int64_t)(word_start[img] + j) * tp : nullptr;
  for (int i = threadIdx.x; i < tp; i += blockDim.x) rec[row + 10 + i] = live ? src[i] : 0.f;
}

// one CTA per weight row: amax -> power-of-two pre-scale putting the row's largest magnitude in [256, 512) (both fp16
// planes stay normal), hi = fp16(w * 2^s), lo = fp16(w * 2^s - hi); scale[row] *= 2^-s / kActScale (undone in the GEMM
// epilogue together with the activation pre-scale).  Bit-identical to packing._pack_rows.
__global__ void __launch_bounds__(256) prepack_rows_kernel(const float* __restrict__ w, int k, __half* __restrict__ hi,
                                                           __half* __restrict__ lo, float* __restrict__ scale) {
  __shared__ float s_red[8];
  __shared__ float s_mul;
  const int row = blockIdx.x;
  const float* wr = w + (int64_t)row * k;
  float amax = 0.f;
  for (int i = threadIdx.x; i < k; i += blockDim.x) amax = fmaxf(amax, fabsf(wr[i]));
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, s_red[i]);
    int s = 0;
    if (m > 0.f) {
      // floor(log2(512 / m)) from the exponent: m = f * 2^e with f in [0.5, 1)  ->  512 / m in (2^(9-e), 2^(10-e)]
      int e;
      const float f = frexpf(m, &e);
      s = (f == 0.5f) ? 10 - e : 9 - e;
      s = max(-24, min(24, s));
    }
    s_mul = exp2f((float)s);
    scale[row] = scale[row] * exp2f((float)-s) / kActScale;
  }
  __syncthreads();
  const float mul = s_mul;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const float x = fminf(fmaxf(wr[i] * mul, -60000.f), 60000.f);
    const __half h = __float2half_rn(x);
    hi[(int64_t)row * k + i] = h;
    lo[(int64_t)row * k + i] = __float2half_rn(x - __half2float(h));
  }
}

}  // namespace glass

using namespace glass;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int glass_pack_rois(const float* boxes, const int32_t* counts, int n_img, int max_det, float* rois,
                               int32_t* word_start, int32_t* total, void* stream) {
  GLASS_CHECK(boxes && counts && rois && word_start && total, "null pointer");
  GLASS_CHECK(n_img > 0 && n_img <= 4096 && max_det > 0, "bad shape (n_img <= 4096)");
  GLASS_CUDA(launch_pdl(pack_rois_kernel, dim3(1), dim3(256), (n_img + 1) * sizeof(int32_t), STREAM, boxes, counts, n_img, max_det, rois, word_start, total));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_pack_detections(const float* boxes, const float* scores, const float* orient, const int32_t* counts,
                                     const float* probs, const int32_t* word_start, int n_img, int max_det, int steps,
                                     int classes, float* rec, void* stream) {
  GLASS_CHECK(boxes && scores && counts && probs && word_start && rec, "null pointer");
  GLASS_CHECK(n_img > 0 && max_det > 0 && steps > 0 && classes > 0, "bad shape");
  GLASS_CUDA(launch_pdl(pack_detections_kernel, dim3(n_img * max_det), dim3(128), 0, STREAM, boxes, scores, orient, counts, probs, word_start, max_det,
                                                              steps * classes, rec));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_prepack_weights(const float* w, int n, int k, void* hi, void* lo, float* scale, void* stream) {
  GLASS_CHECK(w && hi && lo && scale, "null pointer");
  GLASS_CHECK(n > 0 && k > 0, "bad shape");
  prepack_rows_kernel<<<n, 256, 0, STREAM>>>(w, k, (__half*)hi, (__half*)lo, scale);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
