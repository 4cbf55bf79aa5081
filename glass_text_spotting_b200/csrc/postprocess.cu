// postprocess.cu -- the reference's word post-processor as ONE kernel, one CTA per image (sm_100a).
//
// Replaces the host loop of PostProcessorRotatedBoxes.__call__ (glass/postprocess/post_processor_rotated_boxes.py:
// 66-184: size / score filters, then "merge until nothing merges": N x N rotated intersection-over-min-area, pair
// masks, cv2.minAreaRect merge of each valid pair on the host, nms_rotated(0.99), repeat) and the text-score filter
// of PostProcessorAcademic.__call__ (post_processor_academic.py:26-34).  In the reference every iteration costs a
// device->host copy, a Python loop over pairs with an OpenCV call each, and a host->device copy; here the <= 128
// detections of an image live in shared memory for the whole loop and nothing returns to the host.
//
// Pair order, write-back order and comparisons follow the reference exactly:
//  * pairs = nonzero(triu(ioa) >= 0.01) is row-major (i < j); boxes[pairs[:,0]] = merged, then boxes[pairs[:,1]] = merged
//    with repeated indices means: box x takes the merge of the LAST pair naming it as second member (largest i < x),
//    else of the last pair naming it as first member (largest j > x);
//  * the merge orientation is the higher-scored box's angle in RADIANS compared against degrees (:203-205 vs :266) --
//    a quirk of the reference that is reproduced, not fixed;
//  * nms_rotated returns survivors in descending-score order, which becomes the box order of the next round.
// The min-area rectangle of the 8 corner points is found by brute force: a minimal rectangle has a side collinear
// with a hull edge, every hull edge joins two of the points, and a non-hull direction can only give a larger
// rectangle -- so the minimum over the 28 point-pair directions is the minimum (double precision).
// Built with --fmad=false: threshold decisions round like the fp32 CPU path.
#include <math_constants.h>

#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"
#include "rotated_iou.cuh"

namespace glass {

constexpr int PP_MAX = 128;  // detections per image (DETECTIONS_PER_IMAGE = 100)
constexpr int PP_WORDS = PP_MAX / 32;
constexpr int PP_THREADS = 512;  // the pair loops (IoA matrix, NMS relation) are spread over 4x more threads than boxes

struct PostprocessKernelParams {
  const float* boxes;
  const float* scores;
  const float* text_scores;
  const int32_t* counts;
  int32_t n_img, m;
  float min_box_dim, valid_score, detect_threshold, text_threshold;
  float merge_ioa_thresh, height_ratio_lo, height_ratio_hi, max_angle_diff, minimal_ioa_thresh, nms_iou;
  int32_t max_iters;
  float* out_boxes;
  float* out_scores;
  float* out_polygons;
  int32_t* out_index;
  int32_t* out_count;
  int32_t* out_iters;
};

// torch.remainder for floats: fmod, then shifted into the divisor's sign
__device__ __forceinline__ float torch_remainder(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && ((b < 0.f) != (m < 0.f))) m += b;
  return m;
}
__device__ __forceinline__ double py_mod(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0 && ((b < 0.0) != (m < 0.0))) m += b;
  return m;
}

// boxes_to_polygons (post_processor_rotated_boxes.py:221-249): vertex k of box (cx, cy, w, h, a)
__device__ inline void box_polygon(const float* b, float (&px)[4], float (&py)[4]) {
  const float cx = b[0], cy = b[1], w = b[2], h = b[3], a = b[4];
  const float t = (-a / 180.f) * 3.14159265358979323846f;
  const float st = (float)sin((double)t), ct = (float)cos((double)t);
  px[0] = cx + (h * st - w * ct) / 2.f;
  px[1] = cx + (h * st + w * ct) / 2.f;
  px[2] = cx - (h * st - w * ct) / 2.f;
  px[3] = cx - (h * st + w * ct) / 2.f;
  py[0] = cy - (h * ct + w * st) / 2.f;
  py[1] = cy - (h * ct - w * st) / 2.f;
  py[2] = cy + (h * ct + w * st) / 2.f;
  py[3] = cy + (h * ct - w * st) / 2.f;
}

// _merge_rotated_boxes + polygons_to_rotated_boxes (:186-218, :251-286) for one pair
__device__ inline void merge_pair(const float* b1, const float* b2, float s1, float s2, float* out) {
  float qx[8], qy[8];
  {
    float ax[4], ay[4], bx[4], by[4];
    box_polygon(b1, ax, ay);
    box_polygon(b2, bx, by);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      qx[k] = ax[k]; qy[k] = ay[k];
      qx[4 + k] = bx[k]; qy[4 + k] = by[k];
    }
  }
  // orientation handed to polygons_to_rotated_boxes: radians of the higher-scored box (fp32 tensor arithmetic)
  const float orient = (s1 >= s2) ? (b1[4] * 3.14159265358979323846f / 180.f) : (b2[4] * 3.14159265358979323846f / 180.f);

  double best_area = CUDART_INF, bux = 1.0, buy = 0.0, bumin = 0, bumax = 0, bvmin = 0, bvmax = 0;
  for (int p = 0; p < 8; ++p) {
    for (int q = p + 1; q < 8; ++q) {
      const double dx = (double)qx[q] - (double)qx[p], dy = (double)qy[q] - (double)qy[p];
      const double len2 = dx * dx + dy * dy;
      if (!(len2 > 1e-12)) continue;
      const double inv = 1.0 / sqrt(len2);
      const double ux = dx * inv, uy = dy * inv;
      double umin = CUDART_INF, umax = -CUDART_INF, vmin = CUDART_INF, vmax = -CUDART_INF;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double pu = (double)qx[k] * ux + (double)qy[k] * uy;
        const double pv = -(double)qx[k] * uy + (double)qy[k] * ux;
        umin = fmin(umin, pu); umax = fmax(umax, pu);
        vmin = fmin(vmin, pv); vmax = fmax(vmax, pv);
      }
      const double area = (umax - umin) * (vmax - vmin);
      if (area < best_area) {
        best_area = area; bux = ux; buy = uy;
        bumin = umin; bumax = umax; bvmin = vmin; bvmax = vmax;
      }
    }
  }
  // rectangle: axes u = (bux, buy), v = (-buy, bux); a box of angle a has its width axis along (cos a, -sin a)
  const double cu = 0.5 * (bumin + bumax), cv = 0.5 * (bvmin + bvmax);
  const double cxr = cu * bux - cv * buy, cyr = cu * buy + cv * bux;
  const double eu = bumax - bumin, ev = bvmax - bvmin;
  const double a0 = atan2(-buy, bux) * 180.0 / 3.14159265358979323846;
  // the reference turns cv2's angle by multiples of 90 until (orient - angle) wraps into (-45, 45]
  double ang = a0, wd = eu, ht = ev;
  for (int k = 0; k < 4; ++k) {
    const double cand = a0 + 90.0 * k;
    const float diff0 = orient - (float)cand;  // fp32 tensor minus python float
    const float diff = torch_remainder(diff0 + 180.f, 360.f) - 180.f;
    if (diff > -45.f && diff <= 45.f) {
      ang = cand;
      wd = (k & 1) ? ev : eu;
      ht = (k & 1) ? eu : ev;
      break;
    }
  }
  ang = py_mod(ang + 180.0, 360.0) - 180.0;
  out[0] = (float)cxr; out[1] = (float)cyr; out[2] = (float)wd; out[3] = (float)ht; out[4] = (float)ang;
}

__global__ void __launch_bounds__(PP_THREADS) postprocess_merge_kernel(const PostprocessKernelParams p) {
  __shared__ float bx[PP_MAX][5];
  __shared__ float nb[PP_MAX][5];
  __shared__ float sc[PP_MAX];
  __shared__ int oi[PP_MAX];
  __shared__ RBox rb[PP_MAX];
  __shared__ unsigned int vmask[PP_MAX][PP_WORDS];  // bit j of row i: pair (i, j), i < j, passes every merge test
  __shared__ int order[PP_MAX];
  __shared__ unsigned char supp[PP_MAX];
  __shared__ float tb[PP_MAX][5];
  __shared__ float ts[PP_MAX];
  __shared__ int ti[PP_MAX];
  __shared__ int s_n, s_any;

  const int img = blockIdx.x;
  const int t = threadIdx.x;
  const int m = min(p.counts ? p.counts[img] : p.m, min(p.m, PP_MAX));
  const float* boxes = p.boxes + (int64_t)img * p.m * 5;
  const float* scores = p.scores + (int64_t)img * p.m;

  // ---- filter_small_boxes (:87-92) and scores >= valid_score (:98), order preserved
  if (t == 0) {
    int n = 0;
    for (int i = 0; i < m; ++i) {
      const float w = boxes[i * 5 + 2], h = boxes[i * 5 + 3], s = scores[i];
      if (fminf(w, h) >= p.min_box_dim && s >= p.valid_score) {
#pragma unroll
        for (int k = 0; k < 5; ++k) bx[n][k] = boxes[i * 5 + k];
        sc[n] = s;
        oi[n] = i;
        ++n;
      }
    }
    s_n = n;
  }
  __syncthreads();

  int iters = 0;
  while (iters < p.max_iters) {
    const int n = s_n;
    if (n == 0) break;
    if (t < n) rb[t] = make_rbox(bx[t][0], bx[t][1], bx[t][2], bx[t][3], bx[t][4]);
    for (int k = t; k < PP_MAX * PP_WORDS; k += blockDim.x) (&vmask[0][0])[k] = 0u;
    if (t == 0) s_any = 0;
    __syncthreads();

    // ---- pair tests over the upper triangle (:117-158)
    for (int idx = t; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      if (i >= j) continue;
      // boxes whose circumscribed circles are apart cannot intersect: IoU is exactly 0 on both paths
      const float dx = bx[i][0] - bx[j][0], dy = bx[i][1] - bx[j][1];
      const float ri = 0.5f * sqrtf(bx[i][2] * bx[i][2] + bx[i][3] * bx[i][3]);
      const float rj = 0.5f * sqrtf(bx[j][2] * bx[j][2] + bx[j][3] * bx[j][3]);
      if (dx * dx + dy * dy > (ri + rj + 1.f) * (ri + rj + 1.f)) continue;
      const float iou = rotated_iou(rb[i], rb[j]);
      const float a1 = bx[i][2] * bx[i][3], a2 = bx[j][2] * bx[j][3];
      const float inter = (a1 + a2) * iou / (1.f + iou);  // glass/structures/boxes.py:44
      const float ioa = inter / fminf(a1, a2);
      if (!(ioa >= p.minimal_ioa_thresh)) continue;
      float dang = bx[j][4] - bx[i][4];
      dang = fabsf(torch_remainder(dang + 180.f, 360.f) - 180.f);
      const bool similar_angle = (dang < p.max_angle_diff) || (dang > (180.f - p.max_angle_diff));
      const float ratio = bx[j][3] / bx[i][3];
      const bool similar_height = (p.height_ratio_lo < ratio) && (ratio < p.height_ratio_hi);
      const bool valid_score = fminf(sc[i], sc[j]) >= p.valid_score;
      const bool ioa_ok = ioa >= p.merge_ioa_thresh;
      if (similar_angle && similar_height && valid_score && ioa_ok) {
        atomicOr(&vmask[i][j >> 5], 1u << (j & 31));
        s_any = 1;
      }
    }
    __syncthreads();
    if (!s_any) break;  // both stopping conditions of the reference (:123, :161) reduce to "no valid pair"

    // ---- merged boxes, with the reference's write-back order (:176-178)
    if (t < n) {
      int pi = -1, pj = -1;
      for (int i = t - 1; i >= 0; --i) {
        if (vmask[i][t >> 5] & (1u << (t & 31))) { pi = i; pj = t; break; }
      }
      if (pi < 0) {
        for (int j = n - 1; j > t; --j) {
          if (vmask[t][j >> 5] & (1u << (j & 31))) { pi = t; pj = j; break; }
        }
      }
      if (pi >= 0) {
        merge_pair(bx[pi], bx[pj], sc[pi], sc[pj], nb[t]);
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) nb[t][k] = bx[t][k];
      }
    }
    __syncthreads();
    if (t < n) {
#pragma unroll
      for (int k = 0; k < 5; ++k) bx[t][k] = nb[t][k];
      rb[t] = make_rbox(nb[t][0], nb[t][1], nb[t][2], nb[t][3], nb[t][4]);
      // stable descending rank by score
      int rank = 0;
      const float s = sc[t];
      for (int u = 0; u < n; ++u) rank += (sc[u] > s) || (sc[u] == s && u < t);
      order[rank] = t;
      supp[t] = 0;
    }
    __syncthreads();

    // ---- nms_rotated(boxes, scores, 0.99) (:181).  The suppression relation (IoU > thr between sorted positions
    // a < b) is evaluated for all pairs in parallel into bit rows; one thread then replays the greedy scan on the bits.
    // IoU <= min(area) / max(area), so pairs whose areas differ by more than the threshold allows are skipped unseen.
    for (int k = t; k < PP_MAX * PP_WORDS; k += blockDim.x) (&vmask[0][0])[k] = 0u;
    __syncthreads();
    for (int idx = t; idx < n * n; idx += blockDim.x) {
      const int a = idx / n, b = idx - a * n;
      if (a >= b) continue;
      const int ia = order[a], jb = order[b];
      const float aa = bx[ia][2] * bx[ia][3], ab = bx[jb][2] * bx[jb][3];
      if (fminf(aa, ab) < (p.nms_iou - 0.01f) * fmaxf(aa, ab)) continue;
      const float dx = bx[ia][0] - bx[jb][0], dy = bx[ia][1] - bx[jb][1];
      const float ra = 0.5f * sqrtf(bx[ia][2] * bx[ia][2] + bx[ia][3] * bx[ia][3]);
      const float rj = 0.5f * sqrtf(bx[jb][2] * bx[jb][2] + bx[jb][3] * bx[jb][3]);
      if (dx * dx + dy * dy > (ra + rj + 1.f) * (ra + rj + 1.f)) continue;
      if (rotated_iou(rb[ia], rb[jb]) > p.nms_iou) atomicOr(&vmask[a][b >> 5], 1u << (b & 31));
    }
    __syncthreads();
    if (t == 0) {
      unsigned int dead[PP_WORDS];
#pragma unroll
      for (int w = 0; w < PP_WORDS; ++w) dead[w] = 0u;
      for (int a = 0; a < n; ++a) {
        if (dead[a >> 5] & (1u << (a & 31))) continue;
#pragma unroll
        for (int w = 0; w < PP_WORDS; ++w) dead[w] |= vmask[a][w];
      }
      for (int a = 0; a < n; ++a) supp[order[a]] = (dead[a >> 5] >> (a & 31)) & 1u;
    }
    __syncthreads();
    // survivors, in descending-score order, become the next round's list
    if (t == 0) {
      int k = 0;
      for (int a = 0; a < n; ++a) {
        const int ia = order[a];
        if (!supp[ia]) {
#pragma unroll
          for (int c = 0; c < 5; ++c) tb[k][c] = bx[ia][c];
          ts[k] = sc[ia];
          ti[k] = oi[ia];
          ++k;
        }
      }
      s_n = k;
    }
    __syncthreads();
    if (t < s_n) {
#pragma unroll
      for (int c = 0; c < 5; ++c) bx[t][c] = tb[t][c];
      sc[t] = ts[t];
      oi[t] = ti[t];
    }
    __syncthreads();
    ++iters;
  }
  __syncthreads();

  // ---- scores >= detect_threshold (:103), text score >= text_threshold (post_processor_academic.py:31-32), polygons
  float* ob = p.out_boxes + (int64_t)img * p.m * 5;
  float* os = p.out_scores + (int64_t)img * p.m;
  float* op = p.out_polygons ? p.out_polygons + (int64_t)img * p.m * 8 : nullptr;
  int32_t* ox = p.out_index + (int64_t)img * p.m;
  if (t == 0) {
    const int n = s_n;
    int k = 0;
    for (int i = 0; i < n; ++i) {
      bool keep = sc[i] >= p.detect_threshold;
      if (keep && p.text_scores) keep = p.text_scores[(int64_t)img * p.m + oi[i]] >= p.text_threshold;
      if (keep) {
        order[k++] = i;
      }
    }
    s_n = k;
    p.out_count[img] = k;
    if (p.out_iters) p.out_iters[img] = iters;
  }
  __syncthreads();
  const int k = s_n;
  for (int i = t; i < p.m; i += blockDim.x) {
    if (i < k) {
      const int src = order[i];
#pragma unroll
      for (int c = 0; c < 5; ++c) ob[i * 5 + c] = bx[src][c];
      os[i] = sc[src];
      ox[i] = oi[src];
      if (op) {
        float px[4], py[4];
        box_polygon(bx[src], px, py);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          op[i * 8 + 2 * v] = px[v];
          op[i * 8 + 2 * v + 1] = py[v];
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 5; ++c) ob[i * 5 + c] = 0.f;
      os[i] = 0.f;
      ox[i] = -1;
      if (op) {
#pragma unroll
        for (int v = 0; v < 8; ++v) op[i * 8 + v] = 0.f;
      }
    }
  }
}

// get_instances_text's numeric part (glass/evaluation/text_evaluator.py:323-331 -> TextEncoder.decode_attention,
// glass/modeling/recognition/text_encoder.py:80-151): per step max / argmax over the classes; the word score is the
// product of the max probabilities up to AND including the first stop symbol (or all steps if there is none).
__global__ void __launch_bounds__(128) text_scores_kernel(const float* __restrict__ probs, int n_words, int steps,
                                                          int classes, int stop_index, float* __restrict__ score,
                                                          int32_t* __restrict__ out_idx, float* __restrict__ out_maxp,
                                                          const int32_t* __restrict__ n_dev) {
  if (n_dev) n_words = min(n_words, max(*n_dev, 0));
  __shared__ float s_p[4][64];
  __shared__ int s_i[4][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int word = blockIdx.x * 4 + warp;
  if (word >= n_words) return;
  const float* pw = probs + (int64_t)word * steps * classes;
  for (int st = 0; st < steps; ++st) {
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int c = lane; c < classes; c += 32) {
      const float v = pw[(int64_t)st * classes + c];
      if (v > best) { best = v; bi = c; }  // first maximum wins within a lane (ascending c)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0 && st < 64) { s_p[warp][st] = best; s_i[warp][st] = bi; }
  }
  __syncwarp();
  if (lane == 0) {
    int first_stop = steps;
    for (int st = 0; st < steps; ++st)
      if (s_i[warp][st] == stop_index) { first_stop = st; break; }
    const int last = min(first_stop, steps - 1);  // mask[word_length] = True, word_length = min(#before stop, T - 1)
    float prod = 1.f;
    for (int st = 0; st < steps; ++st) {
      const float v = st <= last ? s_p[warp][st] : 1.f;  // masked entries are set to 1 before the product
      prod = st == 0 ? v : prod * v;
    }
    score[word] = prod;
  }
  for (int st = lane; st < steps; st += 32) {
    if (out_idx) out_idx[(int64_t)word * steps + st] = s_i[warp][st];
    if (out_maxp) out_maxp[(int64_t)word * steps + st] = s_p[warp][st];
  }
}

}  // namespace glass

using namespace glass;

extern "C" int glass_text_scores(const float* probs, int n_words, int steps, int classes, int stop_index, float* score,
                                 int32_t* out_idx, float* out_maxp, const int32_t* n_dev, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(probs && score, "null pointer");
  GLASS_CHECK(n_words >= 0 && steps > 0 && steps <= 64 && classes > 0, "steps must be in 1..64");
  if (n_words == 0) return 0;
  text_scores_kernel<<<(n_words + 3) / 4, 128, 0, stream>>>(probs, n_words, steps, classes, stop_index, score, out_idx,
                                                           out_maxp, n_dev);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_postprocess_merge(const GlassPostprocessParams* p, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(p != nullptr, "null params");
  GLASS_CHECK(p->boxes && p->scores && p->out_boxes && p->out_scores && p->out_index && p->out_count, "null pointer");
  GLASS_CHECK(p->n_img >= 0 && p->m > 0 && p->m <= PP_MAX, "m must be in 1..128");
  GLASS_CHECK(p->pairs_height_ratio_thresh >= 0.f && p->max_iters > 0, "bad thresholds");
  if (p->n_img == 0) return 0;
  PostprocessKernelParams k{};
  k.boxes = p->boxes; k.scores = p->scores; k.text_scores = p->text_scores; k.counts = p->counts;
  k.n_img = p->n_img; k.m = p->m;
  k.min_box_dim = p->min_box_dim; k.valid_score = p->valid_score; k.detect_threshold = p->detect_threshold;
  k.text_threshold = p->text_threshold; k.merge_ioa_thresh = p->merge_ioa_thresh;
  // the reference compares an fp32 tensor with python floats: thr and 1 / (thr + 1e-6) rounded to fp32
  k.height_ratio_lo = p->pairs_height_ratio_thresh;
  k.height_ratio_hi = (float)(1.0 / ((double)p->pairs_height_ratio_thresh + 1e-6));
  k.max_angle_diff = p->max_angle_diff; k.minimal_ioa_thresh = p->minimal_ioa_thresh; k.nms_iou = p->nms_iou;
  k.max_iters = p->max_iters;
  k.out_boxes = p->out_boxes; k.out_scores = p->out_scores; k.out_polygons = p->out_polygons;
  k.out_index = p->out_index; k.out_count = p->out_count; k.out_iters = p->out_iters;
  postprocess_merge_kernel<<<p->n_img, PP_THREADS, 0, stream>>>(k);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
