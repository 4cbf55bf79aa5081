// rotated_iou.cuh -- IoU of two rotated rectangles (cx, cy, w, h, angle_deg), device side.
//
// Algorithm: detectron2 v0.6 layers/csrc/box_iou_rotated/box_iou_rotated_utils.h:
// centre both boxes on their midpoint, take the corner points, collect edge-edge crossings and the
// corners of one box lying inside the other (one-sided tolerance 1e-5), order the point cloud by a
// Graham scan and take the polygon area.  Reference call sites: batched_nms_rotated at
// glass/modeling/roi_heads/rotated_fast_rcnn.py:131 and inside d2 find_top_rrpn_proposals
// (RotatedRPN inherits it, glass/modeling/proposal_generator/rotated_rpn.py:17).
// Compiled with --fmad=false so every product/sum rounds like the fp32 CPU path (threshold decisions
// in NMS depend on the last bits).
#pragma once
#include <cuda_runtime.h>

namespace glass {

struct Pt {
  float x, y;
};

// Per-box quantities that do not depend on the partner box: half-extent products of the rotation.
struct RBox {
  float cx, cy, w, h;
  float sh, cw, ch, sw;  // (sin/2)*h, (cos/2)*w, (cos/2)*h, (sin/2)*w
};

__device__ __forceinline__ RBox make_rbox(float cx, float cy, float w, float h, float angle_deg) {
  RBox b;
  b.cx = cx; b.cy = cy; b.w = w; b.h = h;
  const double theta = (double)angle_deg * 0.01745329251;
  const float c2 = (float)cos(theta) * 0.5f;
  const float s2 = (float)sin(theta) * 0.5f;
  b.sh = s2 * h; b.cw = c2 * w; b.ch = c2 * h; b.sw = s2 * w;
  return b;
}

__device__ __forceinline__ float pt_dot(Pt a, Pt b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float pt_cross(Pt a, Pt b) { return a.x * b.y - b.x * a.y; }
__device__ __forceinline__ Pt pt_sub(Pt a, Pt b) { return Pt{a.x - b.x, a.y - b.y}; }

__device__ __forceinline__ void rbox_corners(const RBox& b, float cx, float cy, Pt* p) {
  p[0] = Pt{cx + b.sh + b.cw, cy + b.ch - b.sw};
  p[1] = Pt{cx - b.sh + b.cw, cy - b.ch - b.sw};
  p[2] = Pt{2 * cx - p[0].x, 2 * cy - p[0].y};
  p[3] = Pt{2 * cx - p[1].x, 2 * cy - p[1].y};
}

__device__ __forceinline__ bool corner_inside(Pt pt, Pt origin, Pt ab, Pt da, float abab, float adad) {
  const double EPS = 1e-5;
  const Pt ap = pt_sub(pt, origin);
  const float u = pt_dot(ap, ab);
  const float v = -pt_dot(ap, da);
  return (u > -EPS) && (v > -EPS) && (u < abab + EPS) && (v < adad + EPS);
}

__device__ __forceinline__ bool polar_before(Pt a, Pt b) {
  const float t = pt_cross(a, b);
  if (fabsf(t) < 1e-6f) return pt_dot(a, a) < pt_dot(b, b);
  return t > 0;
}

__device__ inline float rotated_iou(const RBox& a, const RBox& b) {
  const float area_a = a.w * a.h, area_b = b.w * b.h;
  if (area_a < 1e-14 || area_b < 1e-14) return 0.f;
  const float mx = (a.cx + b.cx) / 2.0f, my = (a.cy + b.cy) / 2.0f;
  Pt pa[4], pb[4], ea[4], eb[4];
  rbox_corners(a, a.cx - mx, a.cy - my, pa);
  rbox_corners(b, b.cx - mx, b.cy - my, pb);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ea[i] = pt_sub(pa[(i + 1) & 3], pa[i]);
    eb[i] = pt_sub(pb[(i + 1) & 3], pb[i]);
  }
  Pt cloud[24];
  int n = 0;
  const double EPS = 1e-5;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
      const float det = pt_cross(eb[j], ea[i]);
      if (fabs((double)det) <= 1e-14) continue;
      const Pt d = pt_sub(pb[j], pa[i]);
      const float t1 = pt_cross(eb[j], d) / det;
      const float t2 = pt_cross(ea[i], d) / det;
      if (t1 > -EPS && t1 < 1.0f + EPS && t2 > -EPS && t2 < 1.0f + EPS) {
        cloud[n++] = Pt{pa[i].x + ea[i].x * t1, pa[i].y + ea[i].y * t1};
      }
    }
  }
  {
    const float abab = pt_dot(eb[0], eb[0]), adad = pt_dot(eb[3], eb[3]);
    for (int i = 0; i < 4; ++i)
      if (corner_inside(pa[i], pb[0], eb[0], eb[3], abab, adad)) cloud[n++] = pa[i];
  }
  {
    const float abab = pt_dot(ea[0], ea[0]), adad = pt_dot(ea[3], ea[3]);
    for (int i = 0; i < 4; ++i)
      if (corner_inside(pb[i], pa[0], ea[0], ea[3], abab, adad)) cloud[n++] = pb[i];
  }
  float inter = 0.f;
  if (n > 2) {
    // Graham scan: pivot = lowest (then leftmost) point; others sorted by polar angle about it
    int piv = 0;
    for (int i = 1; i < n; ++i)
      if (cloud[i].y < cloud[piv].y || (cloud[i].y == cloud[piv].y && cloud[i].x < cloud[piv].x)) piv = i;
    const Pt origin = cloud[piv];
    Pt q[24];
    for (int i = 0; i < n; ++i) q[i] = pt_sub(cloud[i], origin);
    const Pt tmp = q[0];
    q[0] = q[piv];
    q[piv] = tmp;
    for (int i = 2; i < n; ++i) {  // insertion sort of q[1..n)
      const Pt key = q[i];
      int j = i - 1;
      while (j >= 1 && polar_before(key, q[j])) {
        q[j + 1] = q[j];
        --j;
      }
      q[j + 1] = key;
    }
    int k = 1;
    while (k < n && !(pt_dot(q[k], q[k]) > 1e-8f)) ++k;
    int m = 1;
    if (k < n) {
      q[1] = q[k];
      m = 2;
      for (int i = k + 1; i < n; ++i) {
        while (m > 1 && pt_cross(pt_sub(q[i], q[m - 2]), pt_sub(q[m - 1], q[m - 2])) >= 0) --m;
        q[m++] = q[i];
      }
    }
    if (m > 2) {
      float area = 0.f;
      for (int i = 1; i < m - 1; ++i) area += fabsf(pt_cross(pt_sub(q[i], q[0]), pt_sub(q[i + 1], q[0])));
      inter = area / 2.0f;
    }
  }
  return inter / (area_a + area_b - inter);
}

}  // namespace glass
