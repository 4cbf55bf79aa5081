// d2_ops.cu -- the detectron2 operator surface the reference reaches outside the fused hot path
// (SURVEY.md 8b "operator-level surface"): pairwise rotated IoU / IoA and the general rotated NMS that returns
// EVERY survivor (the fused glass_nms_rotated stops at max_keep <= 128, which is all the hot path needs).
//
//   torch.ops.detectron2.box_iou_rotated  <- pairwise_iou_rotated at glass/structures/boxes.py:34
//   glass.structures.boxes.pairwise_ioa_rotated (:24-49; post_processor_rotated_boxes.py:128)
//   torch.ops.detectron2.nms_rotated      <- glass/postprocess/post_processor_rotated_boxes.py:181,
//                                            glass/modeling/roi_heads/rotated_fast_rcnn.py:131 (batched)
//
// NMS design = detectron2's own (layers/csrc/nms_rotated/nms_rotated_cuda.cu): a 64x64-blocked suppression bitmask
// over the score-sorted boxes, then a sequential scan -- which detectron2 runs on the HOST after a device-to-host copy
// of the mask and which stays on the device here (one CTA, the removed-set in shared memory), so the op never syncs.
// Compiled with --fmad=false like the other decision kernels (threshold tests depend on the last bits).
#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"
#include "rotated_iou.cuh"

namespace glass {

// ---------------------------------------------------------------------------------------------- pairwise IoU / IoA
constexpr int IOU_TX = 32, IOU_TY = 8;

__global__ void __launch_bounds__(IOU_TX* IOU_TY)
    box_iou_rotated_kernel(const float* __restrict__ b1, int n1, const float* __restrict__ b2, int n2, int mode,
                           float* __restrict__ out) {
  __shared__ RBox rows[IOU_TY], cols[IOU_TX];
  const int t = threadIdx.y * IOU_TX + threadIdx.x;
  const int i0 = blockIdx.y * IOU_TY, j0 = blockIdx.x * IOU_TX;
  if (t < IOU_TX) {
    const int j = j0 + t;
    if (j < n2) cols[t] = make_rbox(b2[j * 5 + 0], b2[j * 5 + 1], b2[j * 5 + 2], b2[j * 5 + 3], b2[j * 5 + 4]);
  } else if (t < IOU_TX + IOU_TY) {
    const int i = i0 + t - IOU_TX;
    if (i < n1) rows[t - IOU_TX] = make_rbox(b1[i * 5 + 0], b1[i * 5 + 1], b1[i * 5 + 2], b1[i * 5 + 3], b1[i * 5 + 4]);
  }
  __syncthreads();
  const int i = i0 + threadIdx.y, j = j0 + threadIdx.x;
  if (i >= n1 || j >= n2) return;
  const RBox a = rows[threadIdx.y], b = cols[threadIdx.x];
  const float iou = rotated_iou(a, b);
  float v = iou;
  if (mode == 1) {  // glass/structures/boxes.py:37-47, same operation order in fp32
    const float a1 = a.w * a.h, a2 = b.w * b.h;
    const float inter = (a1 + a2) * iou / (1.0f + iou);
    v = inter / fminf(a1, a2);
  }
  out[(int64_t)i * n2 + j] = v;
}

// ---------------------------------------------------------------------------------------------- general rotated NMS
constexpr int NMS_BLK = 64;

// sorted[r] = RBox of boxes[order[r]]
__global__ void nms_all_gather_kernel(const float* __restrict__ boxes, const int64_t* __restrict__ order, int n,
                                      RBox* __restrict__ sorted) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* b = boxes + order[r] * 5;
  sorted[r] = make_rbox(b[0], b[1], b[2], b[3], b[4]);
}

// mask[r * col_blocks + cb] bit c set <=> iou(sorted[r], sorted[cb*64 + c]) > thr, for cb*64 + c > r.
// Blocks below the diagonal are skipped (never read by the scan).
__global__ void __launch_bounds__(NMS_BLK)
    nms_all_mask_kernel(const RBox* __restrict__ sorted, int n, float thr, int col_blocks,
                        unsigned long long* __restrict__ mask) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (rb > cb) return;
  __shared__ RBox cols[NMS_BLK];
  const int ncol = min(n - cb * NMS_BLK, NMS_BLK);
  if ((int)threadIdx.x < ncol) cols[threadIdx.x] = sorted[cb * NMS_BLK + threadIdx.x];
  __syncthreads();
  const int r = rb * NMS_BLK + threadIdx.x;
  if (r >= n) return;
  const RBox cur = sorted[r];
  unsigned long long bits = 0;
  const int start = (rb == cb) ? threadIdx.x + 1 : 0;
  // circumscribed circles more than a pixel apart => the rectangles are disjoint and the IoU is exactly 0
  const float rad = 0.5f * sqrtf(cur.w * cur.w + cur.h * cur.h);
  for (int c = start; c < ncol; ++c) {
    const RBox o = cols[c];
    if (thr >= 0.f) {
      const float dx = o.cx - cur.cx, dy = o.cy - cur.cy;
      const float reach = rad + 0.5f * sqrtf(o.w * o.w + o.h * o.h) + 1.0f;
      if (dx * dx + dy * dy > reach * reach) continue;
    }
    if (rotated_iou(cur, o) > thr) bits |= 1ULL << c;
  }
  mask[(int64_t)r * col_blocks + cb] = bits;
}

// One CTA, col_blocks <= 128 threads: walk the sorted boxes, keep the ones not yet removed, OR their mask rows in.
__global__ void __launch_bounds__(128)
    nms_all_scan_kernel(const unsigned long long* __restrict__ mask, const int64_t* __restrict__ order, int n,
                        int col_blocks, int64_t* __restrict__ keep, int32_t* __restrict__ keep_count) {
  __shared__ unsigned long long remv[128];
  const int t = threadIdx.x;
  remv[t] = 0;
  __syncthreads();
  int kept = 0;
  for (int r = 0; r < n; ++r) {
    const int blk = r / NMS_BLK;
    const bool alive = !((remv[blk] >> (r % NMS_BLK)) & 1ULL);
    __syncthreads();  // everyone has read remv[blk] before anyone updates it
    if (alive) {
      if (t == 0) keep[kept] = order[r];
      ++kept;
      if (t >= blk && t < col_blocks) remv[t] |= mask[(int64_t)r * col_blocks + t];
    }
    __syncthreads();
  }
  if (t == 0) *keep_count = kept;
  for (int r = kept + t; r < n; r += blockDim.x) keep[r] = -1;
}

}  // namespace glass

using namespace glass;

extern "C" int glass_box_iou_rotated(const float* boxes1, int n1, const float* boxes2, int n2, int mode, float* out,
                                     void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(n1 >= 0 && n2 >= 0, "negative box count");
  GLASS_CHECK(mode == 0 || mode == 1, "mode must be 0 (IoU) or 1 (intersection over the smaller area)");
  if (n1 == 0 || n2 == 0) return 0;
  GLASS_CHECK(boxes1 && boxes2 && out, "null pointer");
  GLASS_CHECK((n1 + IOU_TY - 1) / IOU_TY <= 65535, "n1 too large");
  dim3 grid((n2 + IOU_TX - 1) / IOU_TX, (n1 + IOU_TY - 1) / IOU_TY), block(IOU_TX, IOU_TY);
  box_iou_rotated_kernel<<<grid, block, 0, stream>>>(boxes1, n1, boxes2, n2, mode, out);
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

static inline int64_t nms_all_sorted_bytes(int n) { return ((int64_t)n * sizeof(RBox) + 255) / 256 * 256; }

extern "C" int64_t glass_nms_rotated_all_workspace_bytes(int n) {
  const int64_t cb = (n + NMS_BLK - 1) / NMS_BLK;
  return nms_all_sorted_bytes(n) + (int64_t)n * cb * (int64_t)sizeof(unsigned long long) + 256;
}

extern "C" int glass_nms_rotated_all(const float* boxes, const int64_t* order, int n, float iou_thresh, int64_t* keep,
                                     int32_t* keep_count, void* workspace, int64_t workspace_bytes, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  GLASS_CHECK(n >= 0 && n <= 8192, "n must be in [0, 8192]");
  GLASS_CHECK(keep_count != nullptr, "null keep_count");
  if (n == 0) {
    GLASS_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), stream));
    return 0;
  }
  GLASS_CHECK(boxes && order && keep, "null pointer");
  GLASS_CHECK(workspace && workspace_bytes >= glass_nms_rotated_all_workspace_bytes(n), "workspace too small");
  GLASS_CHECK((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  const int col_blocks = (n + NMS_BLK - 1) / NMS_BLK;
  RBox* sorted = reinterpret_cast<RBox*>(workspace);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + nms_all_sorted_bytes(n));
  nms_all_gather_kernel<<<(n + 127) / 128, 128, 0, stream>>>(boxes, order, n, sorted);
  nms_all_mask_kernel<<<dim3(col_blocks, col_blocks), NMS_BLK, 0, stream>>>(sorted, n, iou_thresh, col_blocks, mask);
  nms_all_scan_kernel<<<1, 128, 0, stream>>>(mask, order, n, col_blocks, keep, keep_count);
  count_launch(3);
  GLASS_CUDA(cudaGetLastError());
  return 0;
}
