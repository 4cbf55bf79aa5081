// layout.cu -- HBM-bound glue kernels around the tcgen05 GEMM: layout conversion between fp32 NCHW
// (the reference's tensors) and split-fp16 padded NHWC (ours), the stem im2col with fused pixel
// normalisation, the generic tap gather for strided convs, and max-pooling.
// All are streaming kernels: 16-byte vector accesses along the channel (innermost) dimension.
#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"

namespace glass {

static inline int grid_for(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------ NCHW fp32 -> split padded NHWC
// One thread per (pixel, 8-channel group): reads 8 strided floats (coalesced across the warp along
// x), writes one 16 B vector per plane.
__global__ void pack_nchw_kernel(const float* __restrict__ src, int n, int c, int h, int w,
                                 __half* __restrict__ dhi, __half* __restrict__ dlo, int cp,
                                 int border) {
  const int groups = (c + 7) / 8;
  const int64_t total = (int64_t)n * groups * h * w;
  const int hp = h + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border), wp = w + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border);
  border = GLASS_BORDER_LO(border);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    int64_t t = i / w;
    const int y = (int)(t % h);
    t /= h;
    const int g = (int)(t % groups);
    const int img = (int)(t / groups);
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v0 = 0.f, v1 = 0.f;
      const int c0 = g * 8 + 2 * j;
      if (c0 < c) v0 = src[(((int64_t)img * c + c0) * h + y) * w + x];
      if (c0 + 1 < c) v1 = src[(((int64_t)img * c + c0 + 1) * h + y) * w + x];
      __half h0, l0, h1, l1;
      split16(v0, h0, l0);
      split16(v1, h1, l1);
      hw[j] = pack16x2(h0, h1);
      lw[j] = pack16x2(l0, l1);
    }
    const int64_t row = ((int64_t)img * hp + y + border) * wp + x + border;
    *reinterpret_cast<uint4*>(dhi + row * cp + g * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(dlo + row * cp + g * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

__global__ void unpack_nchw_kernel(const __half* __restrict__ shi, const __half* __restrict__ slo,
                                   int n, int c, int h, int w, int cp, int border, float* __restrict__ dst) {
  const int64_t total = (int64_t)n * c * h * w;
  const int hp = h + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border), wp = w + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border);
  border = GLASS_BORDER_LO(border);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    int64_t t = i / w;
    const int y = (int)(t % h);
    t /= h;
    const int ch = (int)(t % c);
    const int img = (int)(t / c);
    const int64_t row = ((int64_t)img * hp + y + border) * wp + x + border;
    dst[i] = (__half2float(shi[row * cp + ch]) + __half2float(slo[row * cp + ch])) * kActScaleInv;
  }
}

__global__ void nhwc_f32_to_nchw_kernel(const float* __restrict__ src, int n, int c, int h, int w, int ld,
                                        int border, float* __restrict__ dst) {
  const int64_t total = (int64_t)n * c * h * w;
  const int hp = h + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border), wp = w + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border);
  border = GLASS_BORDER_LO(border);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    int64_t t = i / w;
    const int y = (int)(t % h);
    t /= h;
    const int ch = (int)(t % c);
    const int img = (int)(t / c);
    const int64_t row = ((int64_t)img * hp + y + border) * wp + x + border;
    dst[i] = src[row * ld + ch];
  }
}

// ------------------------------------------------------------------ generic tap gather
__global__ void gather_taps_kernel(const uint4* __restrict__ shi, const uint4* __restrict__ slo, int n, int h, int w,
                                   int cv /*cp/8*/, int border, int kh, int kw, int sh, int sw, int ph, int pw,
                                   int ho, int wo, uint4* __restrict__ dhi, uint4* __restrict__ dlo, int dst_border,
                                   const int32_t* __restrict__ n_dev) {
  pdl_prologue();
  if (n_dev) n = min(n, max(*n_dev, 0));  // live image / word count left on the device by an earlier kernel
  const int taps = kh * kw;
  const int64_t total = (int64_t)n * ho * wo * taps * cv;
  // destination rows: dense [n, ho, wo], or the positions of a padded / shared-border plane of the OUTPUT geometry (the
  // GEMM's M space then is the output plane itself: TMA-store epilogue; border rows are never written)
  const int dlo_b = GLASS_BORDER_LO(dst_border);
  const int hpo = ho + dlo_b + GLASS_BORDER_HI(dst_border), wpo = wo + dlo_b + GLASS_BORDER_HI(dst_border);
  const int hp = h + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border), wp = w + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border);
  border = GLASS_BORDER_LO(border);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    int64_t t = i / cv;
    const int tap = (int)(t % taps);
    const int64_t pix = t / taps;
    const int x = (int)(pix % wo);
    const int y = (int)((pix / wo) % ho);
    const int b = (int)(pix / ((int64_t)wo * ho));
    const int r = tap / kw, s = tap - r * kw;
    const int iy = y * sh - ph + r, ix = x * sw - pw + s;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
      const int64_t row = ((int64_t)b * hp + iy + border) * wp + ix + border;
      vh = __ldg(shi + row * cv + c);
      vl = __ldg(slo + row * cv + c);
    }
    const int64_t o = dst_border == 0 ? i : ((((int64_t)b * hpo + y + dlo_b) * wpo + x + dlo_b) * taps + tap) * cv + c;
    dhi[o] = vh;
    dlo[o] = vl;
  }
}

// ------------------------------------------------------------------ stem: normalise + space-to-depth
// raw fp32 NCHW [n,3,h,w] -> split-fp16 NHWC [n, h/2+2b, w/2+2b, 16] with a b-pixel zero border: pixel (Y, X) holds
// channel (dy*2+dx)*3 + c = (img[c, 2Y+dy, 2X+dx] - mean[c]) * inv_std[c]; channels 12..15 are zero.  The 7x7/s2/p3
// stem conv then is a 4x4 stride-1 conv over this map (taps Y-2..Y+1, X-2..X+1), i.e. 4 k-blocks of 4 pixels x 16
// channels in the GEMM's compact-channel mode -- no 192-wide im2col matrix is materialised (805 MB at bs = 4).
// b = 1 is enough although the taps reach 2 pixels up / left: in the flattened pixel order the cell left of a row's left
// border IS the previous row's right border, and the row above an image's top border IS the previous image's bottom
// border (before the first image: the TMA's out-of-bounds zeros) -- and with b = 1 the map has the geometry of the conv's
// output plane, i.e. the GEMM is flat and stores through TMA.
__global__ void stem_s2d_kernel(const float* __restrict__ img, int n, int h, int w, float m0, float m1, float m2,
                                float is0, float is1, float is2, uint4* __restrict__ dhi, uint4* __restrict__ dlo,
                                int border) {
  pdl_prologue();
  const int ho = h / 2, wo = w / 2, hp = ho + 2 * border, wp = wo + 2 * border;
  const int64_t total = (int64_t)n * ho * wo;
  const float mean[3] = {m0, m1, m2}, istd[3] = {is0, is1, is2};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo);
    const int y = (int)((i / wo) % ho);
    const int b = (int)(i / ((int64_t)wo * ho));
    float v[16];
#pragma unroll
    for (int j = 12; j < 16; ++j) v[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(img + (((int64_t)b * 3 + c) * h + 2 * y + dy) * w + 2 * x));
        v[(dy * 2 + 0) * 3 + c] = (r.x - mean[c]) * istd[c];
        v[(dy * 2 + 1) * 3 + c] = (r.y - mean[c]) * istd[c];
      }
    }
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __half h0, l0, h1, l1;
      split16(v[2 * j], h0, l0);
      split16(v[2 * j + 1], h1, l1);
      hw[j] = pack16x2(h0, h1);
      lw[j] = pack16x2(l0, l1);
    }
    const int64_t o = ((((int64_t)b * hp + y + border) * wp) + x + border) * 2;  // 2 x uint4 per 16-channel pixel
    dhi[o] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    dhi[o + 1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    dlo[o] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    dlo[o + 1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  }
}

// ------------------------------------------------------------------ border re-zero (border = 1)
__global__ void zero_border_kernel(uint4* __restrict__ hi, uint4* __restrict__ lo, int n, int hp, int wp, int cv,
                                   const int32_t* __restrict__ n_dev) {
  pdl_prologue();
  if (n_dev) n = min(n, max(*n_dev, 0));
  const int per_img = 2 * wp + 2 * (hp - 2);
  const int64_t total = (int64_t)n * per_img * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const int64_t t = i / cv;
    const int b = (int)(t % per_img);
    const int img = (int)(t / per_img);
    int y, x;
    if (b < wp) { y = 0; x = b; }
    else if (b < 2 * wp) { y = hp - 1; x = b - wp; }
    else { const int r = b - 2 * wp; y = 1 + (r >> 1); x = (r & 1) ? wp - 1 : 0; }
    const int64_t off = (((int64_t)img * hp + y) * wp + x) * cv + v;
    hi[off] = make_uint4(0, 0, 0, 0);
    lo[off] = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------ max pool (padding = -inf)
__global__ void maxpool_kernel(const uint4* __restrict__ shi, const uint4* __restrict__ slo, int n, int h, int w,
                               int cv, int border, int kh, int kw, int sh, int sw, int ph, int pw, int ho, int wo,
                               uint4* __restrict__ dhi, uint4* __restrict__ dlo, int dborder,
                               const int32_t* __restrict__ n_dev) {
  pdl_prologue();
  if (n_dev) n = min(n, max(*n_dev, 0));
  const int64_t total = (int64_t)n * ho * wo * cv;
  const int hp = h + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border), wp = w + GLASS_BORDER_LO(border) + GLASS_BORDER_HI(border);
  border = GLASS_BORDER_LO(border);
  const int dhp = ho + GLASS_BORDER_LO(dborder) + GLASS_BORDER_HI(dborder), dwp = wo + GLASS_BORDER_LO(dborder) + GLASS_BORDER_HI(dborder);
  dborder = GLASS_BORDER_LO(dborder);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const int64_t pix = i / cv;
    const int x = (int)(pix % wo);
    const int y = (int)((pix / wo) % ho);
    const int b = (int)(pix / ((int64_t)wo * ho));
    float best[8];
    uint32_t bh[8], bl[8];  // fp16 bit patterns of the winning (hi, lo)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      bh[j] = 0;
      bl[j] = 0;
    }
    for (int r = 0; r < kh; ++r) {
      const int iy = y * sh - ph + r;
      if (iy < 0 || iy >= h) continue;
      for (int s = 0; s < kw; ++s) {
        const int ix = x * sw - pw + s;
        if (ix < 0 || ix >= w) continue;
        const int64_t row = ((int64_t)b * hp + iy + border) * wp + ix + border;
        const uint4 a = __ldg(shi + row * cv + c);
        const uint4 d = __ldg(slo + row * cv + c);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 vv = unpack16x2(aw[j], dw[j]);
          const float v0 = vv.x, v1 = vv.y;
          if (v0 > best[2 * j]) {
            best[2 * j] = v0;
            bh[2 * j] = aw[j] & 0xffffu;
            bl[2 * j] = dw[j] & 0xffffu;
          }
          if (v1 > best[2 * j + 1]) {
            best[2 * j + 1] = v1;
            bh[2 * j + 1] = aw[j] >> 16;
            bl[2 * j + 1] = dw[j] >> 16;
          }
        }
      }
    }
    const int64_t orow = ((int64_t)b * dhp + y + dborder) * dwp + x + dborder;
    dhi[orow * cv + c] = make_uint4(bh[0] | (bh[1] << 16), bh[2] | (bh[3] << 16), bh[4] | (bh[5] << 16),
                                    bh[6] | (bh[7] << 16));
    dlo[orow * cv + c] = make_uint4(bl[0] | (bl[1] << 16), bl[2] | (bl[3] << 16), bl[4] | (bl[5] << 16),
                                    bl[6] | (bl[7] << 16));
  }
}

// ------------------------------------------------------------------ bilinear resize, uint8 HWC -> fp32 CHW
// torch.nn.functional.interpolate(mode="bilinear", align_corners=False, size=(ho, wo)) semantics:
// src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out, neighbours clamped at the last pixel.
__global__ void resize_bilinear_u8_kernel(const uint8_t* __restrict__ src, int h, int w, int flip,
                                          float* __restrict__ dst, int ho, int wo) {
  const float sy = (float)h / (float)ho, sx = (float)w / (float)wo;
  const int64_t total = (int64_t)ho * wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo), y = (int)(i / wo);
    const float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = flip ? 2 - c : c;
      const float p00 = (float)__ldg(src + ((int64_t)y0 * w + x0) * 3 + cs);
      const float p01 = (float)__ldg(src + ((int64_t)y0 * w + x1) * 3 + cs);
      const float p10 = (float)__ldg(src + ((int64_t)y1 * w + x0) * 3 + cs);
      const float p11 = (float)__ldg(src + ((int64_t)y1 * w + x1) * 3 + cs);
      dst[(int64_t)c * total + i] = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
    }
  }
}

}  // namespace glass

using namespace glass;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int glass_resize_bilinear_u8(const uint8_t* src_hwc, int h, int w, int flip_channels, float* dst_chw, int ho,
                                        int wo, void* stream) {
  GLASS_CHECK(src_hwc && dst_chw, "null pointer");
  GLASS_CHECK(h > 0 && w > 0 && ho > 0 && wo > 0, "bad shape");
  resize_bilinear_u8_kernel<<<grid_for((int64_t)ho * wo, 256), 256, 0, STREAM>>>(src_hwc, h, w, flip_channels, dst_chw,
                                                                                  ho, wo);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_pack_nchw(const float* src, int n, int c, int h, int w, void* dst_hi, void* dst_lo, int cp,
                               int border, void* stream) {
  GLASS_CHECK(src && dst_hi && dst_lo, "null pointer");
  GLASS_CHECK(n > 0 && c > 0 && h > 0 && w > 0 && cp >= c && cp % 8 == 0 && border >= 0, "bad shape");
  const int64_t total = (int64_t)n * ((c + 7) / 8) * h * w;
  pack_nchw_kernel<<<grid_for(total, 256), 256, 0, STREAM>>>(src, n, c, h, w, (__half*)dst_hi,
                                                             (__half*)dst_lo, cp, border);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_unpack_nchw(const void* src_hi, const void* src_lo, int n, int c, int h, int w, int cp,
                                 int border, float* dst, void* stream) {
  GLASS_CHECK(src_hi && src_lo && dst, "null pointer");
  GLASS_CHECK(n > 0 && c > 0 && h > 0 && w > 0 && cp >= c && border >= 0, "bad shape");
  const int64_t total = (int64_t)n * c * h * w;
  unpack_nchw_kernel<<<grid_for(total, 256), 256, 0, STREAM>>>((const __half*)src_hi,
                                                               (const __half*)src_lo, n, c, h, w, cp, border,
                                                               dst);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_nhwc_f32_to_nchw(const float* src, int n, int c, int h, int w, int ld, int border, float* dst,
                                      void* stream) {
  GLASS_CHECK(src && dst, "null pointer");
  GLASS_CHECK(n > 0 && c > 0 && h > 0 && w > 0 && ld >= c && border >= 0, "bad shape");
  const int64_t total = (int64_t)n * c * h * w;
  nhwc_f32_to_nchw_kernel<<<grid_for(total, 256), 256, 0, STREAM>>>(src, n, c, h, w, ld, border, dst);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_stem_s2d(const float* img, int n, int h, int w, const float* mean, const float* inv_std,
                              void* dst_hi, void* dst_lo, int border, void* stream) {
  GLASS_CHECK(img && mean && inv_std && dst_hi && dst_lo, "null pointer");
  GLASS_CHECK(border == 1 || border == 2, "border must be 1 or 2");
  GLASS_CHECK(n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "image size must be even");
  GLASS_CHECK((reinterpret_cast<uintptr_t>(img) & 7) == 0, "image must be 8-byte aligned");
  const int64_t total = (int64_t)n * (h / 2) * (w / 2);
  GLASS_CUDA(launch_pdl(stem_s2d_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM, img, n, h, w, mean[0], mean[1], mean[2], inv_std[0],
                                                            inv_std[1], inv_std[2], (uint4*)dst_hi, (uint4*)dst_lo,
                                                            border));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_gather_taps(const void* src_hi, const void* src_lo, int n, int h, int w, int cp, int border,
                                 int kh, int kw, int sh, int sw, int ph, int pw, int ho, int wo, void* dst_hi,
                                 void* dst_lo, int dst_border, const int32_t* n_dev, void* stream) {
  GLASS_CHECK(src_hi && src_lo && dst_hi && dst_lo, "null pointer");
  GLASS_CHECK(n > 0 && h > 0 && w > 0 && cp % 8 == 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && ho > 0 && wo > 0,
              "bad shape");
  const int64_t total = (int64_t)n * ho * wo * kh * kw * (cp / 8);
  GLASS_CUDA(launch_pdl(gather_taps_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM, (const uint4*)src_hi, (const uint4*)src_lo, n, h, w,
                                                               cp / 8, border, kh, kw, sh, sw, ph, pw, ho, wo,
                                                               (uint4*)dst_hi, (uint4*)dst_lo, dst_border, n_dev));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_zero_border(void* hi, void* lo, int n, int h, int w, int cp, const int32_t* n_dev, void* stream) {
  GLASS_CHECK(hi && lo, "null pointer");
  GLASS_CHECK(n > 0 && h > 0 && w > 0 && cp > 0 && cp % 8 == 0, "bad shape");
  const int hp = h + 2, wp = w + 2;
  const int64_t total = (int64_t)n * (2 * wp + 2 * (hp - 2)) * (cp / 8);
  GLASS_CUDA(launch_pdl(zero_border_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM, (uint4*)hi, (uint4*)lo, n, hp, wp, cp / 8, n_dev));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_maxpool(const void* src_hi, const void* src_lo, int n, int h, int w, int cp, int border, int kh,
                             int kw, int sh, int sw, int ph, int pw, int ho, int wo, void* dst_hi, void* dst_lo,
                             int dst_border, const int32_t* n_dev, void* stream) {
  GLASS_CHECK(src_hi && src_lo && dst_hi && dst_lo, "null pointer");
  GLASS_CHECK(n > 0 && h > 0 && w > 0 && cp % 8 == 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && ho > 0 && wo > 0,
              "bad shape");
  const int64_t total = (int64_t)n * ho * wo * (cp / 8);
  GLASS_CUDA(launch_pdl(maxpool_kernel, dim3(grid_for(total, 256)), dim3(256), 0, STREAM, (const uint4*)src_hi, (const uint4*)src_lo, n, h, w,
                                                           cp / 8, border, kh, kw, sh, sw, ph, pw, ho, wo,
                                                           (uint4*)dst_hi, (uint4*)dst_lo, dst_border, n_dev));
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
