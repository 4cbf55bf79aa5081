// common.cuh -- shared device helpers for libglass_b200 (sm_100a only).
//
// Hand-written PTX wrappers for the Blackwell primitives the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// UMMA shared-memory / instruction descriptors.  No CUTLASS: the bit layouts follow
// the PTX ISA (cross-checked against cute/arch/mma_sm100_desc.hpp).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

namespace glass {

// ---------------------------------------------------------------- host-side error state
void set_last_error(const std::string& msg);
int fail(const std::string& msg);  // sets last error, returns -1

#define GLASS_CHECK(cond, msg)                                        \
  do {                                                                \
    if (!(cond)) return ::glass::fail(std::string(__func__) + ": " + (msg)); \
  } while (0)

#define GLASS_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return ::glass::fail(std::string(__func__) + ": " #expr ": " + cudaGetErrorString(_e)); \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// Programmatic dependent launch along the step's chain of ~170 launches.  The GEMM launches carry the
// programmatic-stream-serialization attribute (conv_gemm.cu); every other kernel of the chain starts with pdl_prologue(),
// whose griddepcontrol.launch_dependents lets a GEMM that FOLLOWS it be scheduled -- barrier init, TMEM allocation,
// descriptor prefetch -- while this kernel still runs, instead of after it has drained (+3 % end to end, also inside a CUDA
// graph).  Giving the small kernels the attribute as well (GLASS_PDL_ALL=1: their CTAs park at griddepcontrol.wait beside
// the running GEMM) was measured slower (191 vs 197 images/s), so by default they launch plainly.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  static const bool on = getenv("GLASS_PDL_ALL") && atoi(getenv("GLASS_PDL_ALL")) != 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// First statement of such a kernel: let the NEXT kernel of the stream be scheduled as this one's CTAs retire, then (a
// no-op unless this launch itself is programmatic) wait until the PREVIOUS one has completed and its writes are visible.
// Nothing written by an earlier kernel (device-side counts included) may be read before it.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// Plane geometry code carried by every `border` argument of the ABI (include/glass_b200.h): low byte = LEADING zero rows /
// columns of a padded plane; GLASS_BORDER_SHARED set = no trailing ones (the next row's / plane's leading border serves as
// the trailing border of this one).
#define GLASS_BORDER_LO(code) ((code) & 0xff)
#define GLASS_BORDER_HI(code) (((code) & GLASS_BORDER_SHARED) ? 0 : ((code) & 0xff))

#ifdef __CUDACC__
// ---------------------------------------------------------------- small utilities
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// "split-fp16" storage: an fp32 value x is held as two fp16 planes (hi, lo) with
//   kActScale * x ~= hi + lo      (22 mantissa bits; |lo| <= 2^-11 |hi|).
// The power-of-two pre-scale keeps the lo plane out of fp16's subnormal range for |x| >~ 8e-3 and
// leaves head-room up to |x| = 60000/16; it is undone exactly in the GEMM epilogue / by readers.
static constexpr float kActScale = 16.0f;
static constexpr float kActScaleInv = 0.0625f;

__device__ __forceinline__ void split16(float x, __half& hi, __half& lo) {
  const float xs = fminf(fmaxf(x * kActScale, -60000.f), 60000.f);
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}

__device__ __forceinline__ uint32_t pack16x2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// two packed fp16 words (hi plane, lo plane) -> the two represented fp32 values (scale removed)
__device__ __forceinline__ float2 unpack16x2(uint32_t hi, uint32_t lo) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return make_float2((h.x + l.x) * kActScaleInv, (h.y + l.y) * kActScaleInv);
}

// hi + lo of two packed fp16 words as fp32 (storage scale NOT removed): HADD2.F32 + FHADD per element
// (add.rn.f32.f16 is sm_100's mixed-precision add: fp16 operand + fp32 operand -> fp32)
__device__ __forceinline__ float2 unpack16x2_sum(uint32_t hi, uint32_t lo) {
  float2 r;
  asm("{\n"
      ".reg .b16 h0, h1, l0, l1;\n"
      ".reg .f32 f0, f1;\n"
      "mov.b32 {h0, h1}, %2;\n"
      "mov.b32 {l0, l1}, %3;\n"
      "cvt.f32.f16 f0, h0;\n"
      "cvt.f32.f16 f1, h1;\n"
      "add.rn.f32.f16 %0, l0, f0;\n"
      "add.rn.f32.f16 %1, l1, f1;\n"
      "}"
      : "=f"(r.x), "=f"(r.y)
      : "r"(hi), "r"(lo));
  return r;
}

// two fp32 values -> packed (hi, hi) and (lo, lo) words: same values as split16 on each, with the two
// roundings to fp16 done by one packed conversion each
__device__ __forceinline__ void split16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float a = fminf(fmaxf(x0 * kActScale, -60000.f), 60000.f);
  const float b = fminf(fmaxf(x1 * kActScale, -60000.f), 60000.f);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// two fp32 values ALREADY in the storage scale (kActScale * x) -> packed (hi, hi) and (lo, lo) words.  One packed
// saturating conversion per plane (F2FP.SATFINITE.F16.F32.PACK_AB): values beyond fp16's range saturate at +-65504
// instead of turning into inf - inf; 6 instructions per pair against 14 for two split16 calls.
__device__ __forceinline__ void split16x2_scaled(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("{\n"
      ".reg .b16 h0, h1;\n"
      ".reg .f32 f0, f1;\n"
      "cvt.rn.satfinite.f16x2.f32 %0, %3, %2;\n"
      "mov.b32 {h0, h1}, %0;\n"
      "cvt.f32.f16 f0, h0;\n"
      "cvt.f32.f16 f1, h1;\n"
      "sub.rn.f32 f0, %2, f0;\n"
      "sub.rn.f32 f1, %3, f1;\n"
      "cvt.rn.satfinite.f16x2.f32 %1, f1, f0;\n"
      "}"
      : "=&r"(hi), "=r"(lo)
      : "f"(a), "f"(b));
}
// hi + lo of two packed fp16 words as fp32, storage scale kept (the counterpart of split16x2_scaled)
__device__ __forceinline__ float2 unpack16x2_scaled(uint32_t hi, uint32_t lo) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return make_float2(h.x + l.x, h.y + l.y);
}

// one lane of a converged warp (the compiler keeps warp-uniform operands in uniform registers around it)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Explicit shared-state-space vector accesses.  A pointer derived from the dynamic shared-memory base through integer
// arithmetic (the 1024-byte alignment of the operand ring) loses its state space: the compiler then emits GENERIC
// LD.E / ST.E, which go through the L1TEX address path and sit on the long scoreboard (ncu, round 2: those were the
// hottest stall sites of the GEMM epilogue).  These force LDS / STS.
__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a broken pipeline traps (-> launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      __trap();
    }
  }
}

// Shared-space ADDRESS variants (uint32_t, as returned by smem_u32): a generic pointer costs a cvta sequence at every use,
// which the compiler re-materialises freely inside register-starved loops (the GEMM epilogue spent a fifth of its issue
// slots that way).  opaque_u32 launders a value through volatile asm so that it is KEPT in a register (or spilled) instead
// of being recomputed from special registers.
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_a(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_a(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_remote_a(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// pull a tensor box into L2 ahead of the TMA load that will want it
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c0), "r"(c1)
               : "memory");
}

// TMA store (shared -> global, bulk async-group completion) and its fences
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the source shared memory of every committed store has been READ (it may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... of all but the newest `pending` committed groups (pending in 0..2)
__device__ __forceinline__ void tma_store_wait_read_but(int pending) {
  if (pending <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (pending == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], fp16 x fp16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Warp-collective: lane i of the warp gets 16 consecutive fp32 columns of TMEM lane
// (quadrant*32 + i).  taddr = base | (lane_base << 16) | column.
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- CTA-pair (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier (even CTA of the pair)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
      "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows from each CTA] * B[N/2 rows from each CTA]; issued by the leader CTA only
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair when the issued MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// UMMA shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 16-bit elements.
// A tile is rows x 64 fp16 (128 B per row), 8-row groups 1024 B apart (SBO), as
// written by a TMA box with CU_TENSOR_MAP_SWIZZLE_128B into 1024 B-aligned smem.
//   [0,14) start address >> 4   [16,30) LBO >> 4 (ignored for swizzled K-major)
//   [32,46) SBO >> 4            [46,48) version = 1 (sm_100)   [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor, kind::f16: A = B = fp16 (K-major), D = fp32, dense.
//   [4,6) c_format = 1 (F32)  [7,10) a_format = 0 (F16)  [10,13) b_format = 0 (F16)
//   [15] a_major = 0 (K)  [16] b_major = 0 (K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ inline uint32_t umma_idesc_f16_f32(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace glass
