// baseline_d2.cu -- BENCHMARK BASELINE, not on the product path.
//
// A faithful restatement of how detectron2 v0.6 runs rotated RoIAlign on a GPU (layers/csrc/ROIAlignRotated/
// ROIAlignRotated_cuda.cu, restated from its published algorithm -- detectron2 is not installable offline, SURVEY.md
// 8d cfg 3): fp32 NCHW feature map, ONE THREAD PER OUTPUT ELEMENT (n, c, ph, pw) in a grid-stride loop, every thread
// recomputing the RoI geometry and doing 4 scalar loads per sample.  ROIPooler calls it once per FPN level on the RoIs
// assigned to that level.  bench.py --workload roialign_512 times it next to glass_roi_align_rotated ("HBM GB/s vs
// detectron2 CUDA op", BASELINE.json configs[2]); tests/test_gpu_kernels.py checks that both agree with the oracle.
#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"

namespace glass {

__device__ __forceinline__ float d2_bilinear(const float* __restrict__ plane, int height, int width, float y, float x) {
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) return 0.f;
  if (y < 0.f) y = 0.f;
  if (x < 0.f) x = 0.f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = y - (float)y_low, lx = x - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
  const float v1 = plane[y_low * width + x_low], v2 = plane[y_low * width + x_high];
  const float v3 = plane[y_high * width + x_low], v4 = plane[y_high * width + x_high];
  return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

__global__ void d2_roi_align_rotated_forward(int64_t nthreads, const float* __restrict__ input, float spatial_scale,
                                             int channels, int height, int width, int pooled_h, int pooled_w,
                                             int sampling_ratio, const float* __restrict__ rois,
                                             float* __restrict__ top) {
  for (int64_t index = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; index < nthreads;
       index += (int64_t)gridDim.x * blockDim.x) {
    const int pw = (int)(index % pooled_w);
    const int ph = (int)((index / pooled_w) % pooled_h);
    const int c = (int)((index / pooled_w / pooled_h) % channels);
    const int n = (int)(index / pooled_w / pooled_h / channels);
    const float* roi = rois + (int64_t)n * 6;
    const int batch = (int)roi[0];
    const float center_w = roi[1] * spatial_scale - 0.5f, center_h = roi[2] * spatial_scale - 0.5f;
    const float roi_w = roi[3] * spatial_scale, roi_h = roi[4] * spatial_scale;
    const float theta = roi[5] * 3.14159265358979323846f / 180.0f;
    const float cos_t = cosf(theta), sin_t = sinf(theta);
    const float bin_h = roi_h / (float)pooled_h, bin_w = roi_w / (float)pooled_w;
    const float* plane = input + ((int64_t)batch * channels + c) * height * width;
    const int grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / (float)pooled_h);
    const int grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / (float)pooled_w);
    const float start_h = -roi_h / 2.0f, start_w = -roi_w / 2.0f;
    const float count = (float)max(grid_h * grid_w, 1);
    float acc = 0.f;
    for (int iy = 0; iy < grid_h; ++iy) {
      const float yy = start_h + ph * bin_h + ((float)iy + .5f) * bin_h / (float)grid_h;
      for (int ix = 0; ix < grid_w; ++ix) {
        const float xx = start_w + pw * bin_w + ((float)ix + .5f) * bin_w / (float)grid_w;
        const float y = yy * cos_t - xx * sin_t + center_h;
        const float x = yy * sin_t + xx * cos_t + center_w;
        acc += d2_bilinear(plane, height, width, y, x);
      }
    }
    top[index] = acc / count;
  }
}

}  // namespace glass

using namespace glass;

extern "C" int glass_baseline_roi_align_rotated_d2(const float* input_nchw, int n, int channels, int height, int width,
                                                   const float* rois, int n_rois, float spatial_scale, int pooled_h,
                                                   int pooled_w, int sampling_ratio, float* out_nchw, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(input_nchw && rois && out_nchw, "null pointer");
  GLASS_CHECK(n > 0 && channels > 0 && height > 0 && width > 0 && n_rois >= 0 && pooled_h > 0 && pooled_w > 0, "bad shape");
  if (n_rois == 0) return 0;
  const int64_t total = (int64_t)n_rois * channels * pooled_h * pooled_w;
  // detectron2: grid(min(ceil_div(output_size, 512), 4096)), block(512)
  int64_t blocks = (total + 511) / 512;
  if (blocks > 4096) blocks = 4096;
  d2_roi_align_rotated_forward<<<(int)blocks, 512, 0, stream>>>(total, input_nchw, spatial_scale, channels, height,
                                                               width, pooled_h, pooled_w, sampling_ratio, rois, out_nchw);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
