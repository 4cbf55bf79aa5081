/*
 * oracle/d2_ops.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement, in plain C, of the three detectron2 v0.6 native operators that
 * sit on GLASS's inference path.  detectron2 is a third-party dependency of the
 * reference (pinned at v0.6: /root/reference/demo/glass_demo.ipynb cell 4,
 * /root/reference/README.md:36) and is NOT vendored under /root/reference, so the
 * published algorithm is restated here [d2-recall] and pinned by detectron2's own
 * upstream known-answer tests (tests/test_oracle_d2_ops.py):
 *
 *   roi_align_rotated_forward  <- detectron2/layers/csrc/ROIAlignRotated/ROIAlignRotated_cpu.cpp
 *       reference call sites: glass/modeling/fusion/recognizers_hybrid_head.py:200-205,320
 *       (box pooler), :464-469,550 (recognizer pooler), :495-500,556 (image pooler)
 *   box_iou_rotated            <- detectron2/layers/csrc/box_iou_rotated/box_iou_rotated_utils.h
 *   nms_rotated                <- detectron2/layers/csrc/nms_rotated/nms_rotated_cpu.cpp
 *       reference call sites: glass/modeling/roi_heads/rotated_fast_rcnn.py:131
 *       (batched_nms_rotated) and d2 find_top_rrpn_proposals (RotatedRPN inherits it,
 *       glass/modeling/proposal_generator/rotated_rpn.py:17)
 *
 * All arithmetic is fp32 as in the reference (T = float); compile with
 * -ffp-contract=off so no FMA contraction changes the rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ */
/* ROIAlignRotated forward, NCHW fp32.                                  */
/* rois: [n_rois, 6] = (batch_idx, cx, cy, w, h, angle_deg)             */
/* ------------------------------------------------------------------ */
void oracle_roi_align_rotated_forward(const float* input, const float* rois, float* output,
                                      int n_rois, int channels, int height, int width,
                                      int pooled_h, int pooled_w, float spatial_scale,
                                      int sampling_ratio) {
  for (int n = 0; n < n_rois; ++n) {
    const float* roi = rois + (size_t)n * 6;
    const int batch = (int)roi[0];
    const float offset = 0.5f;
    const float roi_center_w = roi[1] * spatial_scale - offset;
    const float roi_center_h = roi[2] * spatial_scale - offset;
    const float roi_width = roi[3] * spatial_scale;
    const float roi_height = roi[4] * spatial_scale;
    const float theta = (float)((double)roi[5] * M_PI / 180.0);
    const float cos_theta = cosf(theta);
    const float sin_theta = sinf(theta);
    const float bin_size_h = roi_height / (float)pooled_h;
    const float bin_size_w = roi_width / (float)pooled_w;
    const int grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / (float)pooled_h);
    const int grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / (float)pooled_w);
    const int cnt = grid_h * grid_w;
    const float count = (float)(cnt > 1 ? cnt : 1);
    const float roi_start_h = -roi_height / 2.0f;
    const float roi_start_w = -roi_width / 2.0f;

    for (int ph = 0; ph < pooled_h; ++ph) {
      for (int pw = 0; pw < pooled_w; ++pw) {
        float* out = output + (((size_t)n * channels) * pooled_h + ph) * pooled_w + pw;
        for (int c = 0; c < channels; ++c) out[(size_t)c * pooled_h * pooled_w] = 0.f;
        for (int iy = 0; iy < grid_h; ++iy) {
          const float yy = roi_start_h + ph * bin_size_h +
                           ((float)iy + .5f) * bin_size_h / (float)grid_h;
          for (int ix = 0; ix < grid_w; ++ix) {
            const float xx = roi_start_w + pw * bin_size_w +
                             ((float)ix + .5f) * bin_size_w / (float)grid_w;
            float y = yy * cos_theta - xx * sin_theta + roi_center_h;
            float x = yy * sin_theta + xx * cos_theta + roi_center_w;
            if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) continue;
            if (y < 0.f) y = 0.f;
            if (x < 0.f) x = 0.f;
            int y_low = (int)y, x_low = (int)x, y_high, x_high;
            if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; }
            else y_high = y_low + 1;
            if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; }
            else x_high = x_low + 1;
            const float ly = y - (float)y_low, lx = x - (float)x_low;
            const float hy = 1.f - ly, hx = 1.f - lx;
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            for (int c = 0; c < channels; ++c) {
              const float* f = input + ((size_t)batch * channels + c) * height * width;
              out[(size_t)c * pooled_h * pooled_w] +=
                  w1 * f[y_low * width + x_low] + w2 * f[y_low * width + x_high] +
                  w3 * f[y_high * width + x_low] + w4 * f[y_high * width + x_high];
            }
          }
        }
        for (int c = 0; c < channels; ++c) out[(size_t)c * pooled_h * pooled_w] /= count;
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/* Rotated IoU (box_iou_rotated_utils.h).                               */
/* ------------------------------------------------------------------ */
typedef struct { float x, y; } pt_t;

static inline float dot2(pt_t a, pt_t b) { return a.x * b.x + a.y * b.y; }
static inline float cross2(pt_t a, pt_t b) { return a.x * b.y - b.x * a.y; }
static inline pt_t sub2(pt_t a, pt_t b) { pt_t r = {a.x - b.x, a.y - b.y}; return r; }

static void rotated_vertices(const float* b, pt_t* p) {
  /* b = (cx, cy, w, h, angle_deg); M_PI/180 constant as in the d2 header */
  const double theta = (double)b[4] * 0.01745329251;
  const float c2 = (float)cos(theta) * 0.5f;
  const float s2 = (float)sin(theta) * 0.5f;
  p[0].x = b[0] + s2 * b[3] + c2 * b[2];
  p[0].y = b[1] + c2 * b[3] - s2 * b[2];
  p[1].x = b[0] - s2 * b[3] + c2 * b[2];
  p[1].y = b[1] - c2 * b[3] - s2 * b[2];
  p[2].x = 2 * b[0] - p[0].x;
  p[2].y = 2 * b[1] - p[0].y;
  p[3].x = 2 * b[0] - p[1].x;
  p[3].y = 2 * b[1] - p[1].y;
}

static int intersection_points(const pt_t* p1, const pt_t* p2, pt_t* out) {
  pt_t v1[4], v2[4];
  for (int i = 0; i < 4; ++i) {
    v1[i] = sub2(p1[(i + 1) % 4], p1[i]);
    v2[i] = sub2(p2[(i + 1) % 4], p2[i]);
  }
  const double EPS = 1e-5;
  int num = 0;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
      const float det = cross2(v2[j], v1[i]);
      if (fabs(det) <= 1e-14) continue;
      const pt_t v12 = sub2(p2[j], p1[i]);
      const float t1 = cross2(v2[j], v12) / det;
      const float t2 = cross2(v1[i], v12) / det;
      if (t1 > -EPS && t1 < 1.0f + EPS && t2 > -EPS && t2 < 1.0f + EPS) {
        out[num].x = p1[i].x + v1[i].x * t1;
        out[num].y = p1[i].y + v1[i].y * t1;
        ++num;
      }
    }
  }
  {
    const pt_t AB = v2[0], DA = v2[3];
    const float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
    for (int i = 0; i < 4; ++i) {
      const pt_t AP = sub2(p1[i], p2[0]);
      const float APdotAB = dot2(AP, AB);
      const float APdotAD = -dot2(AP, DA);
      if (APdotAB > -EPS && APdotAD > -EPS && APdotAB < ABdotAB + EPS && APdotAD < ADdotAD + EPS)
        out[num++] = p1[i];
    }
  }
  {
    const pt_t AB = v1[0], DA = v1[3];
    const float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
    for (int i = 0; i < 4; ++i) {
      const pt_t AP = sub2(p2[i], p1[0]);
      const float APdotAB = dot2(AP, AB);
      const float APdotAD = -dot2(AP, DA);
      if (APdotAB > -EPS && APdotAD > -EPS && APdotAB < ABdotAB + EPS && APdotAD < ADdotAD + EPS)
        out[num++] = p2[i];
    }
  }
  return num;
}

static int hull_less(pt_t a, pt_t b) {
  const float t = cross2(a, b);
  if (fabsf(t) < 1e-6f) return dot2(a, a) < dot2(b, b);
  return t > 0;
}

static int convex_hull_graham(const pt_t* p, int n, pt_t* q) {
  int t = 0;
  for (int i = 1; i < n; ++i)
    if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
  const pt_t start = p[t];
  for (int i = 0; i < n; ++i) q[i] = sub2(p[i], start);
  pt_t tmp = q[0]; q[0] = q[t]; q[t] = tmp;
  /* insertion sort of q[1..n) by polar angle (std::sort in d2; same comparator) */
  for (int i = 2; i < n; ++i) {
    pt_t key = q[i];
    int j = i - 1;
    while (j >= 1 && hull_less(key, q[j])) { q[j + 1] = q[j]; --j; }
    q[j + 1] = key;
  }
  int k;
  for (k = 1; k < n; ++k)
    if (dot2(q[k], q[k]) > 1e-8f) break;
  if (k == n) { q[0] = p[t]; return 1; }
  q[1] = q[k];
  int m = 2;
  for (int i = k + 1; i < n; ++i) {
    while (m > 1 && cross2(sub2(q[i], q[m - 2]), sub2(q[m - 1], q[m - 2])) >= 0) --m;
    q[m++] = q[i];
  }
  return m; /* shift_to_zero = true: leave the hull relative to `start` */
}

static float polygon_area(const pt_t* q, int m) {
  if (m <= 2) return 0.f;
  float area = 0.f;
  for (int i = 1; i < m - 1; ++i)
    area += fabsf(cross2(sub2(q[i], q[0]), sub2(q[i + 1], q[0])));
  return area / 2.0f;
}

float oracle_single_box_iou_rotated(const float* b1_raw, const float* b2_raw) {
  float b1[5], b2[5];
  const float sx = (b1_raw[0] + b2_raw[0]) / 2.0f;
  const float sy = (b1_raw[1] + b2_raw[1]) / 2.0f;
  b1[0] = b1_raw[0] - sx; b1[1] = b1_raw[1] - sy; b1[2] = b1_raw[2]; b1[3] = b1_raw[3]; b1[4] = b1_raw[4];
  b2[0] = b2_raw[0] - sx; b2[1] = b2_raw[1] - sy; b2[2] = b2_raw[2]; b2[3] = b2_raw[3]; b2[4] = b2_raw[4];
  const float area1 = b1[2] * b1[3], area2 = b2[2] * b2[3];
  if (area1 < 1e-14 || area2 < 1e-14) return 0.f;
  pt_t p1[4], p2[4], inter[24], ordered[24];
  rotated_vertices(b1, p1);
  rotated_vertices(b2, p2);
  const int num = intersection_points(p1, p2, inter);
  float intersection = 0.f;
  if (num > 2) {
    const int m = convex_hull_graham(inter, num, ordered);
    intersection = polygon_area(ordered, m);
  }
  return intersection / (area1 + area2 - intersection);
}

void oracle_box_iou_rotated(const float* boxes1, int n1, const float* boxes2, int n2, float* ious) {
  for (int i = 0; i < n1; ++i)
    for (int j = 0; j < n2; ++j)
      ious[(size_t)i * n2 + j] = oracle_single_box_iou_rotated(boxes1 + 5 * i, boxes2 + 5 * j);
}

/* ------------------------------------------------------------------ */
/* nms_rotated: `order` = indices sorted by score descending (caller).  */
/* Suppression test is `iou > thr` (the CUDA kernel's comparison, the   */
/* path the reference runs on).  Returns number kept.    */
/* ------------------------------------------------------------------ */
int oracle_nms_rotated(const float* boxes, const int64_t* order, int n, float iou_threshold,
                       int64_t* keep) {
  uint8_t* suppressed = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
  int num_keep = 0;
  for (int _i = 0; _i < n; ++_i) {
    const int64_t i = order[_i];
    if (suppressed[i]) continue;
    keep[num_keep++] = i;
    for (int _j = _i + 1; _j < n; ++_j) {
      const int64_t j = order[_j];
      if (suppressed[j]) continue;
      if (oracle_single_box_iou_rotated(boxes + 5 * i, boxes + 5 * j) > iou_threshold)
        suppressed[j] = 1;
    }
  }
  free(suppressed);
  return num_keep;
}
