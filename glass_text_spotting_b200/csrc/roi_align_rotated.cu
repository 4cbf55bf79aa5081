// roi_align_rotated.cu -- multi-level rotated RoIAlign for sm_100a (HBM/L2 gather-bound).
//
// Algorithm = detectron2 v0.6 ROIAlignRotated (layers/csrc/ROIAlignRotated) + ROIPooler level
// assignment (modeling/poolers.py), as used by the reference at recognizers_hybrid_head.py:320,
// :550, :556.  Layout is ours: feature maps are fp32 NHWC so that the four
// bilinear taps of a sample are four contiguous C-vectors; one warp owns one output bin and its
// lanes sweep the channel vector with 16-byte loads (C = 256 -> two float4 per lane per tap, a
// fully coalesced 1 KB request).  The geometry (cos/sin, bin size, sampling grid) is warp-uniform.
#include <stdlib.h>

#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"

namespace glass {

struct RoiKernelParams {
  int num_levels;
  const void* feat[GLASS_MAX_LEVELS];
  const void* feat_lo[GLASS_MAX_LEVELS];
  int feat_h[GLASS_MAX_LEVELS], feat_w[GLASS_MAX_LEVELS];
  float scale[GLASS_MAX_LEVELS];
  int border, ld, channels, min_level;
  const float* rois;
  const int* n_rois_dev;
  int n_rois, ph, pw, sampling;
  float* out_f32;
  __half* out_hi;
  __half* out_lo;
  int out_hp, out_wp, out_border, out_coff, ld_out;
  int bins_per_warp;
};

__device__ __forceinline__ int assign_level(float w, float h, int min_level, int num_levels) {
  // d2 assign_boxes_to_levels: floor(4 + log2(sqrt(area)/224 + 1e-8)), clamped to the pooler's levels
  const float size = sqrtf(w * h);
  float lvl = floorf(4.f + log2f(size / 224.f + 1e-8f));
  const float lo = (float)min_level, hi = (float)(min_level + num_levels - 1);
  lvl = fminf(fmaxf(lvl, lo), hi);
  return (int)lvl - min_level;
}

// 4 channels of one pixel: fp32 map -> one 16 B load; split-fp16 map -> two 8 B loads (hi + lo).
template <bool SPLIT_IN>
__device__ __forceinline__ float4 load4(const void* base, const void* base_lo, int64_t pix_off, int ci) {
  if (SPLIT_IN) {
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base) + pix_off) + ci);
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base_lo) + pix_off) + ci);
    const float2 a = unpack16x2(h.x, l.x), b = unpack16x2(h.y, l.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + pix_off) + ci);
}

template <int NV, bool SPLIT_IN>  // float4 vectors per lane: channels <= 128*NV
__global__ void __launch_bounds__(256) roi_align_rotated_kernel(const RoiKernelParams p) {
  pdl_prologue();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_rois = p.n_rois_dev ? min(*p.n_rois_dev, p.n_rois) : p.n_rois;
  const int bins = p.ph * p.pw;
  const int64_t total = (int64_t)n_rois * bins;
  const int cvecs = p.channels >> 2;
  for (int64_t wid = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total;
       wid += (int64_t)gridDim.x * warps_per_block) {
    const int roi_idx = (int)(wid / bins);
    const int bin = (int)(wid - (int64_t)roi_idx * bins);
    const int bph = bin / p.pw, bpw = bin - bph * p.pw;
    const float* roi = p.rois + (int64_t)roi_idx * 6;
    const int batch = (int)roi[0];
    const float bx = roi[1], by = roi[2], bw = roi[3], bh = roi[4], ba = roi[5];
    const int lvl = p.num_levels > 1 ? assign_level(bw, bh, p.min_level, p.num_levels) : 0;
    const int H = p.feat_h[lvl], W = p.feat_w[lvl];
    const float s = p.scale[lvl];
    const int Hp = H + 2 * p.border, Wp = W + 2 * p.border;
    const void* f = p.feat[lvl];
    const void* flo = p.feat_lo[lvl];
    const int64_t img_off = (int64_t)batch * Hp * Wp * p.ld;

    const float cw = bx * s - 0.5f, chh = by * s - 0.5f;
    const float rw = bw * s, rh = bh * s;
    // the sampling coordinates must round like the fp32 CPU path (a 1-ulp difference in cos/sin moves a
    // sample by ~1e-5 px, visible on high-frequency inputs): correctly rounded sin/cos via double
    const float theta = (float)((double)ba * 3.14159265358979323846 / 180.0);
    const float sn = (float)sin((double)theta), cs = (float)cos((double)theta);
    const float bsh = rh / (float)p.ph, bsw = rw / (float)p.pw;
    const int gh = p.sampling > 0 ? p.sampling : (int)ceilf(rh / (float)p.ph);
    const int gw = p.sampling > 0 ? p.sampling : (int)ceilf(rw / (float)p.pw);
    const float count = (float)max(gh * gw, 1);
    const float sh0 = -rh / 2.0f, sw0 = -rw / 2.0f;

    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int iy = 0; iy < gh; ++iy) {
      const float yy = sh0 + bph * bsh + ((float)iy + .5f) * bsh / (float)gh;
      for (int ix = 0; ix < gw; ++ix) {
        const float xx = sw0 + bpw * bsw + ((float)ix + .5f) * bsw / (float)gw;
        float y = yy * cs - xx * sn + chh;
        float x = yy * sn + xx * cs + cw;
        if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
        y = fmaxf(y, 0.f);
        x = fmaxf(x, 0.f);
        int yl = (int)y, xl = (int)x, yh, xh;
        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
        const float ly = y - (float)yl, lx = x - (float)xl;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
        const int64_t o1 = img_off + ((int64_t)(yl + p.border) * Wp + xl + p.border) * p.ld;
        const int64_t o2 = img_off + ((int64_t)(yl + p.border) * Wp + xh + p.border) * p.ld;
        const int64_t o3 = img_off + ((int64_t)(yh + p.border) * Wp + xl + p.border) * p.ld;
        const int64_t o4 = img_off + ((int64_t)(yh + p.border) * Wp + xh + p.border) * p.ld;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int ci = lane + 32 * v;
          if (ci < cvecs) {
            const float4 a = load4<SPLIT_IN>(f, flo, o1, ci), b = load4<SPLIT_IN>(f, flo, o2, ci);
            const float4 c = load4<SPLIT_IN>(f, flo, o3, ci), d = load4<SPLIT_IN>(f, flo, o4, ci);
            acc[v].x += w1 * a.x + w2 * b.x + w3 * c.x + w4 * d.x;
            acc[v].y += w1 * a.y + w2 * b.y + w3 * c.y + w4 * d.y;
            acc[v].z += w1 * a.z + w2 * b.z + w3 * c.z + w4 * d.z;
            acc[v].w += w1 * a.w + w2 * b.w + w3 * c.w + w4 * d.w;
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int ci = lane + 32 * v;
      if (ci < cvecs) {
        const float4 o = make_float4(acc[v].x / count, acc[v].y / count, acc[v].z / count, acc[v].w / count);
        if (p.out_f32) {
          reinterpret_cast<float4*>(p.out_f32 + ((int64_t)roi_idx * bins + bin) * p.channels)[ci] = o;
        }
        if (p.out_hi) {
          const int64_t row = ((int64_t)roi_idx * p.out_hp + bph + p.out_border) * p.out_wp + bpw + p.out_border;
          __half h0, l0, h1, l1, h2, l2, h3, l3;
          split16(o.x, h0, l0);
          split16(o.y, h1, l1);
          split16(o.z, h2, l2);
          split16(o.w, h3, l3);
          const int64_t off = row * p.ld_out + p.out_coff + ci * 4;
          *reinterpret_cast<uint2*>(p.out_hi + off) = make_uint2(pack16x2(h0, h1), pack16x2(h2, h3));
          *reinterpret_cast<uint2*>(p.out_lo + off) = make_uint2(pack16x2(l0, l1), pack16x2(l2, l3));
        }
      }
    }
  }
}

// Split-fp16 input, 8 channels per lane: a tap is one 16-byte load per plane per lane (C = 256 -> the whole
// channel vector in one pass, 512 B per plane per warp request), 8 accumulators, 16-byte stores.
//
// The first version of this kernel was ISSUE-bound, not memory-bound (ncu: 73 % issue slots busy, 13 % DRAM;
// ~500 warp instructions per sample because the file is built with --fmad=false for the coordinate math).
// Per element and tap it now costs 3 instructions: HADD2.F32 (hi -> fp32), FHADD (mixed-precision add of the lo
// half, PTX add.rn.f32.f16) and one explicit FFMA with the 1/16 storage scale folded into the bilinear weight
// (a power of two: exact).  The sampling coordinates keep the unfused fp32 arithmetic of the CPU path.
struct TapSet {
  int o1, o2, o3, o4;   // element offsets of the four taps inside the image plane (32-bit)
  float w1, w2, w3, w4;  // bilinear weights * kActScaleInv; all zero for a sample outside the map
};

__device__ __forceinline__ TapSet make_taps(float y, float x, int H, int W, int Wp, int border, int ld) {
  TapSet t;
  const bool inside = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  y = fmaxf(y, 0.f);
  x = fmaxf(x, 0.f);
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
  if (!inside) { yl = yh = xl = xh = 0; }  // keep the (unused) addresses in range
  const float ly = y - (float)yl, lx = x - (float)xl;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float sc = inside ? kActScaleInv : 0.f;
  t.w1 = hy * hx * sc; t.w2 = hy * lx * sc; t.w3 = ly * hx * sc; t.w4 = ly * lx * sc;
  const int r0 = (yl + border) * Wp + border, r1 = (yh + border) * Wp + border;
  t.o1 = (r0 + xl) * ld; t.o2 = (r0 + xh) * ld; t.o3 = (r1 + xl) * ld; t.o4 = (r1 + xh) * ld;
  return t;
}

// acc += w * (hi + lo) for 8 packed channels, 2.5 instructions per element:
//   hi plane: HADD2.F32 (fp16 -> fp32) and one packed FFMA2 (fma.rn.f32x2) per channel pair;
//   lo plane: one FHFMA (fma.rn.f32.f16: fp16 x fp16 + fp32, the product is exact) with the weight rounded to
//   fp16 -- |lo| <= 2^-11 |hi|, so the weight's 2^-12 rounding error enters at 2^-23 relative, below fp32's own.
__device__ __forceinline__ void fma8(const uint4& h, const uint4& l, float w, float (&acc)[8]) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
  const unsigned short wh = __half_as_ushort(__float2half_rn(w));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
    asm("{\n"
        ".reg .b64 ra, rb, rc;\n"
        ".reg .b16 l0, l1;\n"
        ".reg .f32 t0, t1;\n"
        "mov.b64 ra, {%2, %2};\n"
        "mov.b64 rb, {%3, %4};\n"
        "mov.b64 rc, {%0, %1};\n"
        "fma.rn.f32x2 rc, ra, rb, rc;\n"
        "mov.b64 {t0, t1}, rc;\n"
        "mov.b32 {l0, l1}, %5;\n"
        "fma.rn.f32.f16 %0, l0, %6, t0;\n"
        "fma.rn.f32.f16 %1, l1, %6, t1;\n"
        "}"
        : "+f"(acc[2 * j]), "+f"(acc[2 * j + 1])
        : "f"(w), "f"(f.x), "f"(f.y), "r"(lw[j]), "h"(wh));
  }
}

#define GLASS_LD16(ptr) __ldg(reinterpret_cast<const uint4*>(ptr))

static constexpr int kBinsPerWarp = 4;  // default consecutive bins per warp task: the per-RoI set-up is amortised over them

template <int S>  // S = 2: the sampling grid is 2 x 2 (all four samples' loads are in flight together); 0: generic
__global__ void __launch_bounds__(256, 2) roi_align_rotated_split8_kernel(const RoiKernelParams p) {
  pdl_prologue();
  // One warp per task of kBinsPerWarp consecutive output bins (a warp per whole row of bins was measured slower:
  // too few warps in flight); the grid is NOT persistent so that the block scheduler balances the tail.
  const int lane = threadIdx.x & 31;
  const int n_rois = p.n_rois_dev ? min(*p.n_rois_dev, p.n_rois) : p.n_rois;
  const int bins = p.ph * p.pw;
  const int64_t total = (int64_t)n_rois * bins;
  const bool active = lane * 8 < p.channels;
  const int lane_off = active ? lane * 8 : 0;
  const int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t w_begin = task * p.bins_per_warp;
  const int64_t w_end = min(total, w_begin + p.bins_per_warp);

  // per-RoI state (warp-uniform), recomputed only when the task crosses into another RoI
  int cur_roi = -1;
  int H = 0, W = 0, Wp = 0, gh = 1, gw = 1;
  const __half* fhi = nullptr;
  const __half* flo = nullptr;
  float cw = 0.f, chh = 0.f, sn = 0.f, cs = 1.f, bsh = 0.f, bsw = 0.f, sh0 = 0.f, sw0 = 0.f, count = 1.f, inv_count = 0.f;

  for (int64_t wid = w_begin; wid < w_end; ++wid) {
    const int roi_idx = (int)(wid / bins);
    const int bin = (int)(wid - (int64_t)roi_idx * bins);
    const int bph = bin / p.pw, bpw = bin - bph * p.pw;
    if (roi_idx != cur_roi) {
      cur_roi = roi_idx;
      const float* roi = p.rois + (int64_t)roi_idx * 6;
      const int batch = (int)roi[0];
      const float bx = roi[1], by = roi[2], bw = roi[3], bh = roi[4], ba = roi[5];
      const int lvl = p.num_levels > 1 ? assign_level(bw, bh, p.min_level, p.num_levels) : 0;
      H = p.feat_h[lvl];
      W = p.feat_w[lvl];
      const float s = p.scale[lvl];
      const int Hp = H + 2 * p.border;
      Wp = W + 2 * p.border;
      const int64_t img_off = (int64_t)batch * Hp * Wp * p.ld + lane_off;
      fhi = reinterpret_cast<const __half*>(p.feat[lvl]) + img_off;
      flo = reinterpret_cast<const __half*>(p.feat_lo[lvl]) + img_off;
      cw = bx * s - 0.5f;
      chh = by * s - 0.5f;
      const float rw = bw * s, rh = bh * s;
      // correctly rounded sin/cos (via double) so that the samples land exactly where the fp32 CPU path puts them
      const float theta = (float)((double)ba * 3.14159265358979323846 / 180.0);
      sn = (float)sin((double)theta);
      cs = (float)cos((double)theta);
      bsh = rh / (float)p.ph;
      bsw = rw / (float)p.pw;
      gh = S > 0 ? S : (p.sampling > 0 ? p.sampling : (int)ceilf(rh / (float)p.ph));
      gw = S > 0 ? S : (p.sampling > 0 ? p.sampling : (int)ceilf(rw / (float)p.pw));
      const int cnt = max(gh * gw, 1);
      count = (float)cnt;
      // x / 2^k == x * 2^-k exactly: skip the division sequence for power-of-two sample counts
      inv_count = (cnt & (cnt - 1)) == 0 ? 1.0f / count : 0.f;
      sh0 = -rh / 2.0f;
      sw0 = -rw / 2.0f;
    }

    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (S == 2) {
      // same expressions as the generic path with gh = gw = 2 (x / 2.0f == x * 0.5f exactly)
      const float xx0 = sw0 + bpw * bsw + (0.f + .5f) * bsw * 0.5f;
      const float xx1 = sw0 + bpw * bsw + (1.f + .5f) * bsw * 0.5f;
      const float xs0 = xx0 * sn, xc0 = xx0 * cs, xs1 = xx1 * sn, xc1 = xx1 * cs;
      // (a 24-warp / 80-register variant with only one row of samples in flight was measured 12 % slower)
#pragma unroll
      for (int iy = 0; iy < 2; ++iy) {
        const float yy = sh0 + bph * bsh + ((float)iy + .5f) * bsh * 0.5f;
        const float yc = yy * cs, ys = yy * sn;
        const TapSet t0 = make_taps(yc - xs0 + chh, ys + xc0 + cw, H, W, Wp, p.border, p.ld);
        const TapSet t1 = make_taps(yc - xs1 + chh, ys + xc1 + cw, H, W, Wp, p.border, p.ld);
        // 16 independent 16-byte loads in flight per lane before the first use
        const uint4 a0h = GLASS_LD16(fhi + t0.o1), a0l = GLASS_LD16(flo + t0.o1);
        const uint4 b0h = GLASS_LD16(fhi + t0.o2), b0l = GLASS_LD16(flo + t0.o2);
        const uint4 c0h = GLASS_LD16(fhi + t0.o3), c0l = GLASS_LD16(flo + t0.o3);
        const uint4 d0h = GLASS_LD16(fhi + t0.o4), d0l = GLASS_LD16(flo + t0.o4);
        const uint4 a1h = GLASS_LD16(fhi + t1.o1), a1l = GLASS_LD16(flo + t1.o1);
        const uint4 b1h = GLASS_LD16(fhi + t1.o2), b1l = GLASS_LD16(flo + t1.o2);
        const uint4 c1h = GLASS_LD16(fhi + t1.o3), c1l = GLASS_LD16(flo + t1.o3);
        const uint4 d1h = GLASS_LD16(fhi + t1.o4), d1l = GLASS_LD16(flo + t1.o4);
        fma8(a0h, a0l, t0.w1, acc); fma8(b0h, b0l, t0.w2, acc); fma8(c0h, c0l, t0.w3, acc); fma8(d0h, d0l, t0.w4, acc);
        fma8(a1h, a1l, t1.w1, acc); fma8(b1h, b1l, t1.w2, acc); fma8(c1h, c1l, t1.w3, acc); fma8(d1h, d1l, t1.w4, acc);
      }
    } else {
      for (int iy = 0; iy < gh; ++iy) {
        const float yy = sh0 + bph * bsh + ((float)iy + .5f) * bsh / (float)gh;
        for (int ix = 0; ix < gw; ++ix) {
          const float xx = sw0 + bpw * bsw + ((float)ix + .5f) * bsw / (float)gw;
          const TapSet t = make_taps(yy * cs - xx * sn + chh, yy * sn + xx * cs + cw, H, W, Wp, p.border, p.ld);
          const uint4 ah = GLASS_LD16(fhi + t.o1), al = GLASS_LD16(flo + t.o1);
          const uint4 bh2 = GLASS_LD16(fhi + t.o2), bl = GLASS_LD16(flo + t.o2);
          const uint4 ch = GLASS_LD16(fhi + t.o3), cl = GLASS_LD16(flo + t.o3);
          const uint4 dh = GLASS_LD16(fhi + t.o4), dl = GLASS_LD16(flo + t.o4);
          fma8(ah, al, t.w1, acc); fma8(bh2, bl, t.w2, acc); fma8(ch, cl, t.w3, acc); fma8(dh, dl, t.w4, acc);
        }
      }
    }
    if (active) {
      if (inv_count != 0.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = acc[j] * inv_count;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = acc[j] / count;
      }
      if (p.out_f32) {
        float4* o = reinterpret_cast<float4*>(p.out_f32 + ((int64_t)roi_idx * bins + bin) * p.channels + lane * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      if (p.out_hi) {
        const int64_t row = ((int64_t)roi_idx * p.out_hp + bph + p.out_border) * p.out_wp + bpw + p.out_border;
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split16x2(acc[2 * j], acc[2 * j + 1], hw[j], lw[j]);
        const int64_t off = row * p.ld_out + p.out_coff + lane * 8;
        *reinterpret_cast<uint4*>(p.out_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(p.out_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  }
}

// ------------------------------------------------------------------ image pooler (3 channels, NCHW raw image)
struct ImgRoiKernelParams {
  const float* img;
  int n, h, w, h_pad, w_pad;
  float mean[3], inv_std[3];
  const float* rois;
  const int* n_rois_dev;
  int n_rois, ph, pw, sampling;
  float* out_f32;
  __half* out_hi;
  __half* out_lo;
  int out_border, ld_out;
  const float4* img4;   // optional: the NORMALISED image as pixel-interleaved float4 (c0, c1, c2, 0) [n, h, w]
};

// raw fp32 NCHW [n,3,h,w] -> normalised pixel-interleaved float4 [n,h,w] ((x - mean) * inv_std, the fp32 operation the
// pooler would otherwise repeat for every tap): one 16-byte load per bilinear tap instead of three scalar ones
__global__ void image_to_nhwc4_kernel(const float* __restrict__ img, int n, int h, int w, float m0, float m1, float m2,
                                      float is0, float is1, float is2, float4* __restrict__ out) {
  pdl_prologue();
  const int64_t plane = (int64_t)h * w, total = (int64_t)n * plane;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / plane, r = i - b * plane;
    const float* f = img + b * 3 * plane + r;
    out[i] = make_float4((__ldg(f) - m0) * is0, (__ldg(f + plane) - m1) * is1, (__ldg(f + 2 * plane) - m2) * is2, 0.f);
  }
}

// One CTA = one RoI x a block of IMG_ROWS output rows.  The per-RoI constants -- above all the correctly rounded
// sine / cosine, which are DOUBLE-precision library calls -- are computed ONCE per CTA by one thread (round 1's kernel
// evaluated them in every one of the 16384 bins of a crop: the fp64 pipe, not memory, was its bound), then every thread
// walks its bins with the reference's fp32 arithmetic, operation for operation.
constexpr int IMG_ROWS = 8;
__global__ void __launch_bounds__(256) image_roi_align_rotated_kernel(const ImgRoiKernelParams p) {
  pdl_prologue();
  const int n_rois = p.n_rois_dev ? min(*p.n_rois_dev, p.n_rois) : p.n_rois;
  const int row_blocks = (p.ph + IMG_ROWS - 1) / IMG_ROWS;
  const int roi_idx = blockIdx.x / row_blocks;
  if (roi_idx >= n_rois) return;
  const int row0 = (blockIdx.x - roi_idx * row_blocks) * IMG_ROWS;
  __shared__ float s_c[10];   // batch, cw, chh, rw, rh, sn, cs
  if (threadIdx.x == 0) {
    const float* roi = p.rois + (int64_t)roi_idx * 6;
    const float theta = (float)((double)roi[5] * 3.14159265358979323846 / 180.0);
    s_c[0] = roi[0];
    s_c[1] = roi[1] - 0.5f;
    s_c[2] = roi[2] - 0.5f;
    s_c[3] = roi[3];
    s_c[4] = roi[4];
    s_c[5] = (float)sin((double)theta);
    s_c[6] = (float)cos((double)theta);
  }
  __syncthreads();
  const int batch = (int)s_c[0];
  const float cw = s_c[1], chh = s_c[2], rw = s_c[3], rh = s_c[4], sn = s_c[5], cs = s_c[6];
  const int H = p.h_pad, W = p.w_pad;
  const int bins = p.ph * p.pw;
  const float bsh = rh / (float)p.ph, bsw = rw / (float)p.pw;
  const int gh = p.sampling > 0 ? p.sampling : (int)ceilf(rh / (float)p.ph);
  const int gw = p.sampling > 0 ? p.sampling : (int)ceilf(rw / (float)p.pw);
  const float count = (float)max(gh * gw, 1);
  const float sh0 = -rh / 2.0f, sw0 = -rw / 2.0f;
  const float* f = p.img + (int64_t)batch * 3 * p.h * p.w;
  const float4* f4 = p.img4 ? p.img4 + (int64_t)batch * p.h * p.w : nullptr;
  const int64_t cstride = (int64_t)p.h * p.w;
  const int rows = min(IMG_ROWS, p.ph - row0);
  for (int i = threadIdx.x; i < rows * p.pw; i += blockDim.x) {
    const int bph = row0 + i / p.pw, bpw = i % p.pw;
    const int bin = bph * p.pw + bpw;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int iy = 0; iy < gh; ++iy) {
      const float yy = sh0 + bph * bsh + ((float)iy + .5f) * bsh / (float)gh;
      for (int ix = 0; ix < gw; ++ix) {
        const float xx = sw0 + bpw * bsw + ((float)ix + .5f) * bsw / (float)gw;
        float y = yy * cs - xx * sn + chh;
        float x = yy * sn + xx * cs + cw;
        if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
        y = fmaxf(y, 0.f);
        x = fmaxf(x, 0.f);
        int yl = (int)y, xl = (int)x, yh, xh;
        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
        const float ly = y - (float)yl, lx = x - (float)xl;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
        const bool in1 = yl < p.h && xl < p.w, in2 = yl < p.h && xh < p.w;
        const bool in3 = yh < p.h && xl < p.w, in4 = yh < p.h && xh < p.w;
        if (f4 != nullptr) {   // pre-normalised float4 pixels: same values, one vector load per tap
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 a = in1 ? __ldg(f4 + (int64_t)yl * p.w + xl) : z4;
          const float4 b = in2 ? __ldg(f4 + (int64_t)yl * p.w + xh) : z4;
          const float4 cc = in3 ? __ldg(f4 + (int64_t)yh * p.w + xl) : z4;
          const float4 d = in4 ? __ldg(f4 + (int64_t)yh * p.w + xh) : z4;
          acc[0] += w1 * a.x + w2 * b.x + w3 * cc.x + w4 * d.x;
          acc[1] += w1 * a.y + w2 * b.y + w3 * cc.y + w4 * d.y;
          acc[2] += w1 * a.z + w2 * b.z + w3 * cc.z + w4 * d.z;
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* fc = f + c * cstride;
            const float a = in1 ? (__ldg(fc + (int64_t)yl * p.w + xl) - p.mean[c]) * p.inv_std[c] : 0.f;
            const float b = in2 ? (__ldg(fc + (int64_t)yl * p.w + xh) - p.mean[c]) * p.inv_std[c] : 0.f;
            const float cc = in3 ? (__ldg(fc + (int64_t)yh * p.w + xl) - p.mean[c]) * p.inv_std[c] : 0.f;
            const float d = in4 ? (__ldg(fc + (int64_t)yh * p.w + xh) - p.mean[c]) * p.inv_std[c] : 0.f;
            acc[c] += w1 * a + w2 * b + w3 * cc + w4 * d;
          }
        }
      }
    }
    const float o0 = acc[0] / count, o1 = acc[1] / count, o2 = acc[2] / count;
    if (p.out_f32) {
      float* o = p.out_f32 + (int64_t)roi_idx * 3 * bins + bin;
      o[0] = o0;
      o[bins] = o1;
      o[2 * bins] = o2;
    }
    if (p.out_hi) {
      const int hp = p.ph + 2 * p.out_border, wp = p.pw + 2 * p.out_border;
      const int64_t row = ((int64_t)roi_idx * hp + bph + p.out_border) * wp + bpw + p.out_border;
      __half h0, l0, h1, l1, h2, l2;
      split16(o0, h0, l0);
      split16(o1, h1, l1);
      split16(o2, h2, l2);
      const __half z = __float2half_rn(0.f);
      *reinterpret_cast<uint4*>(p.out_hi + row * p.ld_out) = make_uint4(pack16x2(h0, h1), pack16x2(h2, z), 0, 0);
      *reinterpret_cast<uint4*>(p.out_lo + row * p.ld_out) = make_uint4(pack16x2(l0, l1), pack16x2(l2, z), 0, 0);
    }
  }
}

}  // namespace glass

using namespace glass;

extern "C" int glass_roi_align_rotated(const GlassRoiAlignParams* p, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(p != nullptr, "null params");
  GLASS_CHECK(p->num_levels >= 1 && p->num_levels <= GLASS_MAX_LEVELS, "num_levels out of range");
  GLASS_CHECK(p->channels > 0 && p->channels % 4 == 0 && p->channels <= 256, "channels must be a multiple of 4, <= 256");
  GLASS_CHECK(p->feat_ld >= p->channels && p->feat_ld % 4 == 0, "feat_ld must be >= channels and a multiple of 4");
  GLASS_CHECK(p->rois != nullptr && p->n_rois >= 0, "rois missing");
  GLASS_CHECK(p->pooled_h > 0 && p->pooled_w > 0 && p->sampling_ratio >= 0, "bad pooled size");
  GLASS_CHECK(p->out_f32 || (p->out_hi && p->out_lo), "no output requested");
  if (p->out_hi) GLASS_CHECK(p->ld_out % 4 == 0 && p->out_coff % 4 == 0 && p->out_hp > 0 && p->out_wp > 0, "bad split-output geometry");
  if (p->n_rois == 0) return 0;
  RoiKernelParams k{};
  k.num_levels = p->num_levels;
  for (int l = 0; l < p->num_levels; ++l) {
    GLASS_CHECK(p->feat[l] != nullptr && p->feat_h[l] > 0 && p->feat_w[l] > 0, "bad feature level");
    GLASS_CHECK((reinterpret_cast<uintptr_t>(p->feat[l]) & 15) == 0, "feature maps must be 16-byte aligned");
    k.feat[l] = p->feat[l];
    k.feat_lo[l] = p->feat_lo[l];
    if (p->feat_is_split) GLASS_CHECK(p->feat_lo[l] != nullptr, "feat_lo missing for split feature map");
    k.feat_h[l] = p->feat_h[l];
    k.feat_w[l] = p->feat_w[l];
    k.scale[l] = p->spatial_scale[l];
  }
  k.border = p->feat_border; k.ld = p->feat_ld; k.channels = p->channels; k.min_level = p->min_level;
  k.rois = p->rois; k.n_rois_dev = p->n_rois_dev; k.n_rois = p->n_rois;
  k.ph = p->pooled_h; k.pw = p->pooled_w; k.sampling = p->sampling_ratio;
  k.out_f32 = p->out_f32; k.out_hi = (__half*)p->out_hi; k.out_lo = (__half*)p->out_lo;
  k.out_hp = p->out_hp; k.out_wp = p->out_wp; k.out_border = GLASS_BORDER_LO(p->out_border); k.out_coff = p->out_coff;
  k.ld_out = p->ld_out;
  const int64_t warps = (int64_t)p->n_rois * p->pooled_h * p->pooled_w;
  int64_t blocks = (warps + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8 * 8;  // grid-stride beyond ~8 waves of resident CTAs
  if (blocks > cap) blocks = cap;
  const bool split8 = p->feat_is_split && p->channels % 8 == 0 && p->feat_ld % 8 == 0 &&
                      (!p->out_hi || (p->ld_out % 8 == 0 && p->out_coff % 8 == 0));
  if (split8) {
    GLASS_CHECK((int64_t)(p->feat_h[0] + 2 * p->feat_border) * (p->feat_w[0] + 2 * p->feat_border) * p->feat_ld < ((int64_t)1 << 31),
                "feature plane too large for 32-bit tap offsets");
    static const int bpw_env = getenv("GLASS_ROI_BPW") ? atoi(getenv("GLASS_ROI_BPW")) : 0;  // A/B knob
    k.bins_per_warp = bpw_env > 0 ? bpw_env : kBinsPerWarp;
    const int64_t tasks = (warps + k.bins_per_warp - 1) / k.bins_per_warp;
    const int64_t sblocks = (tasks + 7) / 8;
    GLASS_CHECK(sblocks < ((int64_t)1 << 31), "too many bins");
    if (p->sampling_ratio == 2) roi_align_rotated_split8_kernel<2><<<(int)sblocks, 256, 0, stream>>>(k);
    else roi_align_rotated_split8_kernel<0><<<(int)sblocks, 256, 0, stream>>>(k);
  } else if (p->channels <= 128) {
    if (p->feat_is_split) roi_align_rotated_kernel<1, true><<<(int)blocks, 256, 0, stream>>>(k);
    else roi_align_rotated_kernel<1, false><<<(int)blocks, 256, 0, stream>>>(k);
  } else {
    if (p->feat_is_split) roi_align_rotated_kernel<2, true><<<(int)blocks, 256, 0, stream>>>(k);
    else roi_align_rotated_kernel<2, false><<<(int)blocks, 256, 0, stream>>>(k);
  }
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_image_roi_align_rotated(const GlassImageRoiAlignParams* p, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(p != nullptr, "null params");
  GLASS_CHECK(p->img && p->rois, "null pointer");
  GLASS_CHECK(p->n > 0 && p->h > 0 && p->w > 0 && p->h_pad >= p->h && p->w_pad >= p->w, "bad image shape");
  GLASS_CHECK(p->pooled_h > 0 && p->pooled_w > 0 && p->sampling_ratio >= 0, "bad pooled size");
  GLASS_CHECK(p->out_f32 || (p->out_hi && p->out_lo), "no output requested");
  if (p->out_hi) GLASS_CHECK(p->ld_out >= 8 && p->ld_out % 8 == 0, "ld_out must be a multiple of 8");
  if (p->n_rois == 0) return 0;
  ImgRoiKernelParams k{};
  k.img = p->img; k.n = p->n; k.h = p->h; k.w = p->w; k.h_pad = p->h_pad; k.w_pad = p->w_pad;
  for (int c = 0; c < 3; ++c) { k.mean[c] = p->mean[c]; k.inv_std[c] = p->inv_std[c]; }
  k.rois = p->rois; k.n_rois_dev = p->n_rois_dev; k.n_rois = p->n_rois;
  k.ph = p->pooled_h; k.pw = p->pooled_w; k.sampling = p->sampling_ratio;
  k.out_f32 = p->out_f32; k.out_hi = (__half*)p->out_hi; k.out_lo = (__half*)p->out_lo;
  k.out_border = p->out_border; k.ld_out = p->ld_out;
  k.img4 = nullptr;
  if (p->workspace != nullptr) {   // normalise once per pixel into the caller's workspace, then gather float4 taps
    const int64_t need = (int64_t)p->n * p->h * p->w * 16;
    GLASS_CHECK(p->workspace_bytes >= need, "image pooler workspace too small (n*h*w*16 bytes)");
    GLASS_CHECK((reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0, "workspace must be 16-byte aligned");
    const int64_t px = (int64_t)p->n * p->h * p->w;
    int64_t nb = (px + 255) / 256;
    if (nb > (int64_t)num_sms() * 16) nb = (int64_t)num_sms() * 16;
    image_to_nhwc4_kernel<<<(int)nb, 256, 0, stream>>>(p->img, p->n, p->h, p->w, p->mean[0], p->mean[1], p->mean[2],
                                                       p->inv_std[0], p->inv_std[1], p->inv_std[2],
                                                       reinterpret_cast<float4*>(p->workspace));
    count_launch();
    k.img4 = reinterpret_cast<const float4*>(p->workspace);
  }
  const int64_t blocks = (int64_t)p->n_rois * ((p->pooled_h + IMG_ROWS - 1) / IMG_ROWS);
  GLASS_CHECK(blocks < ((int64_t)1 << 31), "too many RoIs");
  image_roi_align_rotated_kernel<<<(int)blocks, 256, 0, stream>>>(k);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
