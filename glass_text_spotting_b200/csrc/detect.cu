// detect.cu -- the latency-bound "decision" kernels of the detector, all on device with no host
// round trip (the reference syncs to the host in sort / nonzero / the NMS greedy scan):
//   glass_rpn_topk_decode  : per (image, level) exact top-k of the objectness logits (radix select +
//                            bitonic sort, ties by index like a stable sort), anchor generation and
//                            Box2BoxTransformRotated.apply_deltas for the selected anchors only.
//   glass_nms_rotated      : per image: clip / validity filter, batched-NMS offsets, stable sort by
//                            score, greedy rotated NMS with early exit after max_keep survivors.
//   glass_box_decode       : box-predictor logits -> (box, score, orientation) candidates.
// Semantics follow detectron2 v0.6 find_top_rrpn_proposals / batched_nms_rotated /
// Box2BoxTransformRotated and glass/modeling/roi_heads/rotated_fast_rcnn.py:88-148, 344-373, 480-491
// (SURVEY.md A.4, A.7).  Built with --fmad=false: the arithmetic rounds like the fp32 CPU path.
#include <math_constants.h>

#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"
#include "rotated_iou.cuh"

namespace glass {

__device__ __forceinline__ uint32_t float_to_sortable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending uint order == ascending float order
}

// In-place ascending bitonic sort of `n` (power of two) 64-bit keys in shared memory.
__device__ void bitonic_sort_u64(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float py_mod(float v, float m) {  // torch.remainder: sign of the divisor
  float r = fmodf(v, m);
  if (r != 0.f && ((r < 0.f) != (m < 0.f))) r += m;
  return r;
}

#define GLASS_SCALE_CLAMP 4.135166556742356f /* log(1000/16) */

__device__ __forceinline__ void apply_deltas_rotated(const float* d, float ax, float ay, float aw, float ah, float aa,
                                                     const float* wts, float* out) {
  const float dx = d[0] / wts[0], dy = d[1] / wts[1];
  float dw = d[2] / wts[2], dh = d[3] / wts[3];
  const float da = d[4] / wts[4];
  dw = fminf(dw, GLASS_SCALE_CLAMP);
  dh = fminf(dh, GLASS_SCALE_CLAMP);
  out[0] = dx * aw + ax;
  out[1] = dy * ah + ay;
  out[2] = expf(dw) * aw;
  out[3] = expf(dh) * ah;
  const float pa = da * 180.0f / 3.14159265358979323846f + aa;
  out[4] = py_mod(pa + 180.0f, 360.0f) - 180.0f;
}

// ------------------------------------------------------------------------------------------ RPN top-k
struct RpnTopkKernelParams {
  const float* pred;
  int n_img, h, w, ld, A, stride;
  float aw[16], ah[16], aa[16];
  float wts[5];
  int topk, level, num_levels;
  float* out_boxes;
  float* out_scores;
  unsigned int* keys;  // [n_img, h*w*A] dense sortable keys
  unsigned int* hist;  // [n_img, 4096] histogram of the top 12 key bits
};
constexpr int RPN_HIST0_BITS = 12, RPN_HIST0_BINS = 1 << RPN_HIST0_BITS;

// Pass 0, many CTAs per image: gather the strided logits into a dense key array and histogram the top 12 bits
// (sign, exponent, 3 mantissa bits: fine enough that the later passes only touch a few thousand candidates).
__global__ void __launch_bounds__(256) rpn_keys_kernel(const RpnTopkKernelParams p) {
  pdl_prologue();
  __shared__ unsigned int hist[RPN_HIST0_BINS];
  const int img = blockIdx.y;
  const int npix = p.h * p.w;
  for (int i = threadIdx.x; i < RPN_HIST0_BINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const float* base = p.pred + (int64_t)img * npix * p.ld;
  unsigned int* keys = p.keys + (int64_t)img * npix * p.A;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const float* row = base + (int64_t)pix * p.ld;
    for (int a = 0; a < p.A; ++a) {
      const unsigned key = float_to_sortable(__ldg(row + a));
      keys[(int64_t)pix * p.A + a] = key;
      atomicAdd(&hist[key >> (32 - RPN_HIST0_BITS)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RPN_HIST0_BINS; i += blockDim.x)
    if (hist[i]) atomicAdd(&p.hist[img * RPN_HIST0_BINS + i], hist[i]);
}

// Passes 1..3 + compaction + sort + decode, one CTA per image, over the dense keys (coalesced).
__global__ void __launch_bounds__(1024) rpn_topk_decode_kernel(const RpnTopkKernelParams p) {
  pdl_prologue();
  __shared__ unsigned int hist[RPN_HIST0_BINS];
  __shared__ unsigned int coarse[32];
  __shared__ unsigned int s_prefix, s_mask, s_remaining, s_cnt_gt, s_cnt_eq;
  __shared__ unsigned long long sel[1024];
  const int img = blockIdx.x;
  const int npix = p.h * p.w;
  const int n = npix * p.A;
  const int want = min(p.topk, n);
  const float* base = p.pred + (int64_t)img * npix * p.ld;
  const unsigned int* keys = p.keys + (int64_t)img * n;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;  // every image's slice is 16-byte aligned iff n % 4 == 0

  if (threadIdx.x == 0) {
    s_prefix = 0;
    s_mask = 0;
    s_remaining = (unsigned)want;
    s_cnt_gt = 0;
    s_cnt_eq = 0;
  }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sel[i] = ~0ull;
  __syncthreads();

  // radix select (MSB first) of the want-th largest key in digits of 12 + 10 + 10 bits; the 12-bit
  // histogram comes from pass 0, the two remaining passes only count keys inside the threshold bin
  for (int pass = 0; pass < 3; ++pass) {
    const int bits = pass == 0 ? RPN_HIST0_BITS : 10;
    const int shift = pass == 0 ? 20 : (pass == 1 ? 10 : 0);
    const int bins = 1 << bits;
    const unsigned prefix = s_prefix, mask = s_mask;
    if (pass == 0) {
      for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = p.hist[img * RPN_HIST0_BINS + i];
    } else {
      for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      // one CTA walks the whole key array: 16-byte loads, four of them in flight per thread (the scalar loop was
      // latency-bound at ~6 B/clk: 0.6 ms for the 786 k keys of the p2 level)
      auto count = [&](unsigned key) {
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & (unsigned)(bins - 1)], 1u);
      };
      const uint4* k4 = reinterpret_cast<const uint4*>(keys);
      const int n4 = vec_ok ? n >> 2 : 0;
#pragma unroll 4
      for (int e = threadIdx.x; e < n4; e += blockDim.x) {
        const uint4 v = k4[e];
        count(v.x); count(v.y); count(v.z); count(v.w);
      }
      for (int e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x) count(keys[e]);
    }
    __syncthreads();
    {  // two-level scan from the top bin down: 32 warp sums, then one thread walks <= 32 + bins/32 entries
      const int per_warp = bins >> 5;
      const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
      unsigned part = 0;
      for (int i = ln; i < per_warp; i += 32) part += hist[wid * per_warp + i];
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (ln == 0) coarse[wid] = part;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int per_warp = bins >> 5;
      unsigned remaining = s_remaining;
      int wsel = 31;
      for (; wsel > 0; --wsel) {
        if (coarse[wsel] >= remaining) break;
        remaining -= coarse[wsel];
      }
      int d = (wsel + 1) * per_warp - 1;
      for (; d > wsel * per_warp; --d) {
        const unsigned c = hist[d];
        if (c >= remaining) break;
        remaining -= c;
      }
      s_remaining = remaining;
      s_prefix = prefix | ((unsigned)d << shift);
      s_mask = mask | ((unsigned)(bins - 1) << shift);
    }
    __syncthreads();
  }
  const unsigned thr = s_prefix;           // key of the want-th largest element
  const unsigned need_eq = s_remaining;    // how many elements equal to thr are still needed
  const unsigned n_gt = (unsigned)want - need_eq;

  // compaction: everything above the threshold, then `need_eq` elements equal to it
  auto take = [&](unsigned key, int e) {
    if (key > thr) {
      const unsigned slot = atomicAdd(&s_cnt_gt, 1u);
      sel[slot] = ((unsigned long long)(~key) << 32) | (unsigned)e;
    } else if (key == thr) {
      const unsigned slot = atomicAdd(&s_cnt_eq, 1u);
      if (slot < need_eq) sel[n_gt + slot] = ((unsigned long long)(~key) << 32) | (unsigned)e;
    }
  };
  {
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    const int n4 = vec_ok ? n >> 2 : 0;
#pragma unroll 4
    for (int e = threadIdx.x; e < n4; e += blockDim.x) {
      const uint4 v = k4[e];
      take(v.x, 4 * e); take(v.y, 4 * e + 1); take(v.z, 4 * e + 2); take(v.w, 4 * e + 3);
    }
    for (int e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x) take(keys[e], e);
  }
  __syncthreads();
  bitonic_sort_u64(sel, 1024);  // ascending (~key, index) == descending score, ties by index

  for (int r = threadIdx.x; r < p.topk; r += blockDim.x) {
    float* ob = p.out_boxes + ((int64_t)img * p.num_levels * p.topk + (int64_t)p.level * p.topk + r) * 5;
    float* os = p.out_scores + (int64_t)img * p.num_levels * p.topk + (int64_t)p.level * p.topk + r;
    if (r < want) {
      const unsigned e = (unsigned)(sel[r] & 0xffffffffull);
      const int pix = (int)(e / (unsigned)p.A), a = (int)(e % (unsigned)p.A);
      const int y = pix / p.w, x = pix - y * p.w;
      const float* row = base + (int64_t)pix * p.ld;
      float d[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) d[j] = row[p.A + a * 5 + j];
      float box[5];
      apply_deltas_rotated(d, (float)(x * p.stride), (float)(y * p.stride), p.aw[a], p.ah[a], p.aa[a], p.wts, box);
#pragma unroll
      for (int j = 0; j < 5; ++j) ob[j] = box[j];
      *os = row[a];
    } else {
#pragma unroll
      for (int j = 0; j < 5; ++j) ob[j] = 0.f;
      *os = -CUDART_INF_F;
    }
  }
}

// ------------------------------------------------------------------------------------------ rotated NMS
struct NmsKernelParams {
  const float* boxes;
  const float* scores;
  const int* group;
  int group_size;
  const int* m_dev;
  int n_img, m, sort_n;
  const float* img_hw;
  int clip, filter_empty;
  float score_thresh;  // candidates with score <= score_thresh are dropped (use -inf to keep all)
  float iou_thresh;
  int max_keep;
  float* out_boxes;
  float* out_scores;
  int* out_index;
  int* out_count;
  float* ws_boxes;  // [n_img, m, 5] cleaned boxes
  RBox* ws_rbox;    // [n_img, m] offset boxes with rotation products
};

__device__ __forceinline__ void clip_rotated(float* b, float h, float w) {
  // d2 RotatedBoxes.clip(clip_angle_threshold=1.0): normalise the angle; near-horizontal boxes only
  b[4] = py_mod(b[4] + 180.0f, 360.0f) - 180.0f;
  if (fabsf(b[4]) <= 1.0f) {
    float x1 = b[0] - b[2] / 2.0f, y1 = b[1] - b[3] / 2.0f;
    float x2 = b[0] + b[2] / 2.0f, y2 = b[1] + b[3] / 2.0f;
    x1 = fminf(fmaxf(x1, 0.f), w);
    y1 = fminf(fmaxf(y1, 0.f), h);
    x2 = fminf(fmaxf(x2, 0.f), w);
    y2 = fminf(fmaxf(y2, 0.f), h);
    b[0] = (x1 + x2) / 2.0f;
    b[1] = (y1 + y2) / 2.0f;
    b[2] = fminf(b[2], x2 - x1);
    b[3] = fminf(b[3], y2 - y1);
  }
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* scratch) {
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, other) : fminf(v, other);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float x = (lane < (int)(blockDim.x >> 5)) ? scratch[lane] : (is_max ? -CUDART_INF_F : CUDART_INF_F);
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, x, o);
      x = is_max ? fmaxf(x, other) : fminf(x, other);
    }
    if (lane == 0) scratch[0] = x;
  }
  __syncthreads();
  const float r = scratch[0];
  __syncthreads();
  return r;
}

constexpr int NMS_CHUNK = 64;
constexpr int NMS_MAX_KEEP = 128;

__global__ void __launch_bounds__(1024) nms_rotated_kernel(const NmsKernelParams p) {
  pdl_prologue();
  extern __shared__ unsigned long long keys[];  // [sort_n]
  __shared__ float scratch[32];
  __shared__ RBox kept[NMS_MAX_KEEP];
  __shared__ int kept_idx[NMS_MAX_KEEP];
  __shared__ RBox chunk[NMS_CHUNK];
  __shared__ int chunk_idx[NMS_CHUNK];
  __shared__ unsigned long long pair_mask[NMS_CHUNK];
  __shared__ unsigned int supp[2];  // 64-bit "suppressed by an earlier survivor" mask as 2 x u32
  __shared__ int s_nkept, s_nvalid;

  const int img = blockIdx.x;
  const int m = p.m_dev ? min(p.m_dev[img], p.m) : p.m;
  const float ih = p.img_hw ? p.img_hw[2 * img] : 0.f, iw = p.img_hw ? p.img_hw[2 * img + 1] : 0.f;
  const float* boxes = p.boxes + (int64_t)img * p.m * 5;
  const float* scores = p.scores + (int64_t)img * p.m;
  float* wsb = p.ws_boxes + (int64_t)img * p.m * 5;
  RBox* wsr = p.ws_rbox + (int64_t)img * p.m;

  // ---- clean, validate, and find the coordinate range of the valid boxes
  float vmax = -CUDART_INF_F, vmin = CUDART_INF_F;
  for (int i = threadIdx.x; i < p.sort_n; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < m) {
      float b[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) b[j] = boxes[(int64_t)i * 5 + j];
      const float s = scores[i];
      bool ok = isfinite(b[0]) && isfinite(b[1]) && isfinite(b[2]) && isfinite(b[3]) && isfinite(b[4]) &&
                isfinite(s) && (s > p.score_thresh);
      if (ok && p.clip) clip_rotated(b, ih, iw);
      if (ok && p.filter_empty) ok = (b[2] > 0.f) && (b[3] > 0.f);
#pragma unroll
      for (int j = 0; j < 5; ++j) wsb[(int64_t)i * 5 + j] = b[j];
      if (ok) {
        key = ((unsigned long long)(~float_to_sortable(s)) << 32) | (unsigned)i;
        vmax = fmaxf(vmax, fmaxf(b[0], b[1]) + fmaxf(b[2], b[3]) / 2.0f);
        vmin = fminf(vmin, fminf(b[0], b[1]) - fmaxf(b[2], b[3]) / 2.0f);
      }
    }
    keys[i] = key;
  }
  const float cmax = block_reduce(vmax, true, scratch);
  const float cmin = block_reduce(vmin, false, scratch);
  const float unit = cmax - cmin + 1.0f;  // batched_nms_rotated: offsets = idx * (max - min + 1)
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int g = p.group ? p.group[(int64_t)img * p.m + i] : (p.group_size > 0 ? i / p.group_size : 0);
    const float off = (float)g * unit;
    wsr[i] = make_rbox(wsb[(int64_t)i * 5 + 0] + off, wsb[(int64_t)i * 5 + 1] + off, wsb[(int64_t)i * 5 + 2],
                       wsb[(int64_t)i * 5 + 3], wsb[(int64_t)i * 5 + 4]);
  }
  if (threadIdx.x == 0) s_nkept = 0;
  __syncthreads();
  bitonic_sort_u64(keys, p.sort_n);
  // number of valid candidates = first index with the sentinel key
  if (threadIdx.x == 0) s_nvalid = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < p.sort_n; i += blockDim.x) {
    if (keys[i] != ~0ull && (i + 1 == p.sort_n || keys[i + 1] == ~0ull)) s_nvalid = i + 1;
  }
  __syncthreads();
  const int nvalid = s_nvalid;

  // ---- greedy NMS over the sorted candidates in chunks of 64, stopping at max_keep survivors
  for (int c0 = 0; c0 < nvalid; c0 += NMS_CHUNK) {
    const int cn = min(NMS_CHUNK, nvalid - c0);
    if (threadIdx.x < NMS_CHUNK) {
      pair_mask[threadIdx.x] = 0ull;
      if (threadIdx.x < cn) {
        const int idx = (int)(keys[c0 + threadIdx.x] & 0xffffffffull);
        chunk_idx[threadIdx.x] = idx;
        chunk[threadIdx.x] = wsr[idx];
      }
    }
    if (threadIdx.x < 2) supp[threadIdx.x] = 0u;
    __syncthreads();
    const int nk = s_nkept;
    // (a) chunk candidates vs earlier survivors
    for (int t = threadIdx.x; t < cn * nk; t += blockDim.x) {
      const int ci = t % cn, ki = t / cn;
      if (rotated_iou(kept[ki], chunk[ci]) > p.iou_thresh) atomicOr(&supp[ci >> 5], 1u << (ci & 31));
    }
    // (b) pairs inside the chunk: bit j of pair_mask[i] (j > i) = "i suppresses j"
    for (int t = threadIdx.x; t < cn * cn; t += blockDim.x) {
      const int i = t / cn, j = t - i * cn;
      if (j > i && rotated_iou(chunk[i], chunk[j]) > p.iou_thresh) atomicOr(&pair_mask[i], 1ull << j);
    }
    __syncthreads();
    // (c) sequential resolve (cheap bit operations)
    if (threadIdx.x == 0) {
      unsigned long long dead = (unsigned long long)supp[0] | ((unsigned long long)supp[1] << 32);
      int k = nk;
      for (int i = 0; i < cn && k < p.max_keep; ++i) {
        if ((dead >> i) & 1ull) continue;
        kept[k] = chunk[i];
        kept_idx[k] = chunk_idx[i];
        ++k;
        dead |= pair_mask[i];
      }
      s_nkept = k;
    }
    __syncthreads();
    if (s_nkept >= p.max_keep) break;
  }
  const int nk = s_nkept;
  if (threadIdx.x == 0) p.out_count[img] = nk;
  for (int k = threadIdx.x; k < p.max_keep; k += blockDim.x) {
    float* ob = p.out_boxes + ((int64_t)img * p.max_keep + k) * 5;
    if (k < nk) {
      const int idx = kept_idx[k];
#pragma unroll
      for (int j = 0; j < 5; ++j) ob[j] = wsb[(int64_t)idx * 5 + j];
      p.out_scores[(int64_t)img * p.max_keep + k] = scores[idx];
      p.out_index[(int64_t)img * p.max_keep + k] = idx;
    } else {
#pragma unroll
      for (int j = 0; j < 5; ++j) ob[j] = 0.f;
      p.out_scores[(int64_t)img * p.max_keep + k] = 0.f;
      p.out_index[(int64_t)img * p.max_keep + k] = -1;
    }
  }
}

// ------------------------------------------------------------------------------------------ box predictor decode
__global__ void box_decode_kernel(const float* __restrict__ pred, int ld, const float* __restrict__ proposals,
                                  const int* __restrict__ counts, int n_img, int per_img, float w0, float w1,
                                  float w2, float w3, float w4, float* __restrict__ out_boxes,
                                  float* __restrict__ out_scores, float* __restrict__ out_orient) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * per_img) return;
  const int img = i / per_img, r = i - img * per_img;
  float* ob = out_boxes + (int64_t)i * 5;
  if (counts && r >= counts[img]) {
#pragma unroll
    for (int j = 0; j < 5; ++j) ob[j] = 0.f;
    out_scores[i] = -CUDART_INF_F;
    out_orient[2 * i] = 0.f;
    out_orient[2 * i + 1] = 0.f;
    return;
  }
  const float* row = pred + (int64_t)i * ld;  // [0,2) class scores (fg, bg); [2,7) deltas; [7,11) orientation
  const float* pb = proposals + (int64_t)i * 5;
  const float wts[5] = {w0, w1, w2, w3, w4};
  float box[5];
  apply_deltas_rotated(row + 2, pb[0], pb[1], pb[2], pb[3], pb[4], wts, box);
  // softmax over (fg, bg)
  const float mx = fmaxf(row[0], row[1]);
  const float e0 = expf(row[0] - mx), e1 = expf(row[1] - mx);
  const float p_fg = e0 / (e0 + e1), p_bg = e1 / (e0 + e1);
  // orientation: softmax over 4, (argmax, max prob)
  float om = row[7];
  for (int j = 1; j < 4; ++j) om = fmaxf(om, row[7 + j]);
  float oe[4], osum = 0.f;
  for (int j = 0; j < 4; ++j) {
    oe[j] = expf(row[7 + j] - om);
    osum += oe[j];
  }
  int oarg = 0;
  float obest = oe[0] / osum;
  for (int j = 1; j < 4; ++j) {
    const float v = oe[j] / osum;
    if (v > obest) {
      obest = v;
      oarg = j;
    }
  }
  const bool finite = isfinite(box[0]) && isfinite(box[1]) && isfinite(box[2]) && isfinite(box[3]) &&
                      isfinite(box[4]) && isfinite(p_fg) && isfinite(p_bg);
#pragma unroll
  for (int j = 0; j < 5; ++j) ob[j] = box[j];
  out_scores[i] = finite ? p_fg : -CUDART_INF_F;
  out_orient[2 * i] = (float)oarg;
  out_orient[2 * i + 1] = obest;
}

}  // namespace glass

using namespace glass;

extern "C" int64_t glass_rpn_topk_workspace_bytes(int n_img, int h, int w, int num_anchors) {
  return (int64_t)n_img * (RPN_HIST0_BINS + (int64_t)h * w * num_anchors) * (int64_t)sizeof(unsigned int);
}

extern "C" int glass_rpn_topk_decode(const GlassRpnTopkParams* p, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(p != nullptr && p->pred && p->out_boxes && p->out_scores, "null pointer");
  GLASS_CHECK(p->n_img > 0 && p->h > 0 && p->w > 0, "bad shape");
  GLASS_CHECK(p->num_anchors >= 1 && p->num_anchors <= 16, "num_anchors must be in [1,16]");
  GLASS_CHECK(p->ld >= 6 * p->num_anchors, "ld must hold A logits + 5A deltas");
  GLASS_CHECK(p->topk >= 1 && p->topk <= 1024, "topk must be in [1,1024]");
  GLASS_CHECK(p->level >= 0 && p->level < p->num_levels, "bad level slot");
  RpnTopkKernelParams k{};
  k.pred = p->pred; k.n_img = p->n_img; k.h = p->h; k.w = p->w; k.ld = p->ld; k.A = p->num_anchors;
  k.stride = p->stride;
  for (int a = 0; a < p->num_anchors; ++a) {
    k.aw[a] = p->anchor_w[a]; k.ah[a] = p->anchor_h[a]; k.aa[a] = p->anchor_angle[a];
  }
  for (int j = 0; j < 5; ++j) k.wts[j] = p->weights[j];
  k.topk = p->topk; k.level = p->level; k.num_levels = p->num_levels;
  k.out_boxes = p->out_boxes; k.out_scores = p->out_scores;
  GLASS_CHECK(p->workspace && p->workspace_bytes >= glass_rpn_topk_workspace_bytes(p->n_img, p->h, p->w, p->num_anchors),
              "workspace too small");
  k.hist = reinterpret_cast<unsigned int*>(p->workspace);
  k.keys = k.hist + (size_t)p->n_img * RPN_HIST0_BINS;
  GLASS_CUDA(cudaMemsetAsync(k.hist, 0, (size_t)p->n_img * RPN_HIST0_BINS * sizeof(unsigned int), stream));
  const int npix = p->h * p->w;
  int chunks = (npix + 256 * 8 - 1) / (256 * 8);  // ~8 pixels per thread
  if (chunks > 64) chunks = 64;
  rpn_keys_kernel<<<dim3(chunks, p->n_img), 256, 0, stream>>>(k);
  rpn_topk_decode_kernel<<<p->n_img, 1024, 0, stream>>>(k);
  count_launch(2);
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t glass_nms_workspace_bytes(int n_img, int m) {
  return (int64_t)n_img * m * (5 * sizeof(float) + sizeof(RBox)) + 256;
}

extern "C" int glass_nms_rotated(const GlassNmsParams* p, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(p != nullptr && p->boxes && p->scores, "null pointer");
  GLASS_CHECK(p->out_boxes && p->out_scores && p->out_index && p->out_count, "null output");
  GLASS_CHECK(p->n_img > 0 && p->m > 0 && p->m <= 8192, "m must be in [1, 8192]");
  GLASS_CHECK(p->max_keep >= 1 && p->max_keep <= NMS_MAX_KEEP, "max_keep must be in [1,128]");
  GLASS_CHECK(!p->clip || p->img_hw, "clip needs img_hw");
  GLASS_CHECK(p->workspace && p->workspace_bytes >= glass_nms_workspace_bytes(p->n_img, p->m), "workspace too small");
  NmsKernelParams k{};
  k.boxes = p->boxes; k.scores = p->scores; k.group = p->group; k.group_size = p->group_size; k.m_dev = p->m_dev;
  k.n_img = p->n_img; k.m = p->m;
  int sort_n = 64;
  while (sort_n < p->m) sort_n <<= 1;
  k.sort_n = sort_n;
  k.img_hw = p->img_hw; k.clip = p->clip; k.filter_empty = p->filter_empty;
  k.score_thresh = p->score_thresh; k.iou_thresh = p->iou_thresh; k.max_keep = p->max_keep;
  k.out_boxes = p->out_boxes; k.out_scores = p->out_scores; k.out_index = p->out_index; k.out_count = p->out_count;
  uintptr_t ws = (reinterpret_cast<uintptr_t>(p->workspace) + 127) & ~uintptr_t(127);
  k.ws_rbox = reinterpret_cast<RBox*>(ws);
  k.ws_boxes = reinterpret_cast<float*>(ws + (size_t)p->n_img * p->m * sizeof(RBox));
  const int smem = sort_n * (int)sizeof(unsigned long long);
  GLASS_CUDA(cudaFuncSetAttribute(nms_rotated_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
  GLASS_CUDA(launch_pdl(nms_rotated_kernel, dim3(p->n_img), dim3(1024), smem, stream, k));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_box_decode(const float* pred, int ld, const float* proposals, const int32_t* counts, int n_img,
                                int per_img, const float* host_weights, float* out_boxes, float* out_scores,
                                float* out_orient, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(pred && proposals && host_weights && out_boxes && out_scores && out_orient, "null pointer");
  GLASS_CHECK(ld >= 11 && n_img > 0 && per_img > 0, "bad shape");
  const int total = n_img * per_img;
  GLASS_CUDA(launch_pdl(box_decode_kernel, dim3((total + 127) / 128), dim3(128), 0, stream, pred, ld, proposals, counts, n_img, per_img,
                                                            host_weights[0], host_weights[1], host_weights[2],
                                                            host_weights[3], host_weights[4], out_boxes, out_scores,
                                                            out_orient));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
