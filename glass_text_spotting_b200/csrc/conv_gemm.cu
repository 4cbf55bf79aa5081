// conv_gemm.cu -- persistent, warp-specialised tcgen05 implicit-GEMM (sm_100a).
//
// One kernel serves every conv / Linear on GLASS's dense path (see include/glass_b200.h).
//   A: activations, split-fp16 padded NHWC flattened to rows [pixels, C]; a conv tap (r,s) is a
//      constant row shift, so each k-block is ONE 2-D TMA box (64 channels x 128 pixels) at row
//      m0 + shift -- image borders are real zero pixels in memory, tensor edges are TMA OOB zeros.
//   B: weights packed [Cout, taps*C] K-major.
//   D: 128 x BN fp32 partial sums in TMEM, double buffered (2 x 256 columns).  The tensor core's
//      accumulator TRUNCATES (round-toward-zero, measured: error grows linearly with the chain length,
//      tools/accuracy_probe.py), so a chain is cut every kb_per_chunk k-blocks: the MMA warp switches
//      to the other TMEM buffer and the epilogue warps drain the finished chunk into fp32 registers
//      with round-to-nearest adds (Ootomo & Yokota's "accumulate outside the tensor core").
// Roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM
// allocator, warps 4..19 = epilogue (TMEM lane quadrant = warp % 4, four warps per quadrant split the columns).
// Precision modes: fp16x3 split (3 MMAs per k-step into the same accumulator: hi*hi, hi*lo, lo*hi)
// gives fp32-grade results (22-bit operands, fp32 accumulate); mode 1 issues only hi*hi.
#include <cuda.h>
#include <stdlib.h>

#include <new>

#include "common.cuh"
#include "glass_b200.h"
#include "host_util.h"

namespace glass {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int TMEM_COLS = 512;     // all of the SM's tensor memory
static constexpr int MAX_ACC_BUFS = 8;    // accumulator ring: 512 / max(64, tile width rounded to a power of two)
static constexpr int A_TILE_BYTES = BM * BK * 2;
static constexpr int A_BLOCK_ROWS = BM + 8;                 // tap-row mode: one block serves the 3 s-taps (row offsets 0,1,2)
static constexpr int A_BLOCK_BYTES = A_BLOCK_ROWS * BK * 2;  // 17408 = 17 x 1024
static constexpr int MAX_A_STAGES = 4;
static constexpr int EPI_WARPS = 16;
static constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;  // 4 control warps + 16 epilogue warps
static constexpr int MAX_STAGES = 12;
static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 2048;  // per epilogue warp: 32 rows x 16 fp32 columns
static constexpr int EPI_BOX_BYTES = 2 * BM * BK * 2;      // TMA epilogue: one pair (hi, lo) of 128-row x 64-column boxes
static constexpr int MAX_EPI_BOXES = 3;
static constexpr int MAX_CHUNKS_PER_WARP = 4;  // 256 columns / 16 per chunk / 4 column parts

struct GemmKernelParams {
  int64_t rows_m;
  int32_t tiles_m, tiles_n, bn;
  int32_t kblocks_per_tap, ntaps;
  int32_t a_col0;  // pixel-grouped mode: first column of the k-block boxes inside the (overlapping) tensor row
  int32_t tap_shift[GLASS_MAX_TAPS];
  int32_t num_stages, stage_bytes, b_tile_bytes;
  // tap-row mode (3x3 convs): per (tap row r, channel block) ONE activation block of 136 rows is staged and its three
  // s-taps are addressed by advancing the UMMA descriptor start by s rows (the 128B swizzle is a function of the
  // absolute shared-memory address, verified by tools/probes/umma_rowoffset_probe.cu); weights stream per tap.
  int32_t group3, a_stages, a_stage_bytes, b_stages, b_stage_bytes, ring_bytes;
  // resident weights (tap-row mode, one N tile, all k-blocks of the weight tile fit the ring: the 64-channel 3x3 convs):
  // the weight tile is loaded ONCE per CTA and every M tile re-uses it -- those layers were bound by re-streaming it from
  // L2 for every tile (74 KB per tile beside 104 KB of activations)
  int32_t resident_b;
  int32_t kb_per_chunk;  // k-blocks accumulated inside the tensor core before a drain to registers
  int32_t acc_cols, acc_bufs;  // TMEM accumulator ring: acc_bufs (a power of two) buffers of acc_cols columns (product <= 512)
  int32_t acc_shift;           // log2(acc_bufs): ring position / lap parity by mask and shift (no runtime division per chunk)
  int32_t m_h, m_w, m_border;
  int32_t m_border_hi;  // trailing border rows / columns of the M-space planes (0 for shared-border planes)
  const float* scale;
  const float* bias;
  int32_t relu_pre, relu_post;
  const __half* res_hi;
  const __half* res_lo;
  int32_t res_hp, res_wp, res_border, res_shift;
  __half* out_hi;
  __half* out_lo;
  float* out_f32;
  int32_t out_hp, out_wp, out_border;
  int32_t ld_out, ld_f32, n_store;
  // live M extent on the device: rows_m = min(rows_m, *m_count_dev * m_rows_per_count) (detections of this step)
  const int32_t* m_count_dev;
  int32_t m_rows_per_count;
  int32_t* sat_count;  // debug: counts outputs beyond the split-fp16 storage range (|16 y| > 60000), or nullptr
  int32_t tma_res;  // TMA epilogue: the residual tile is TMA-loaded into the staging boxes (same geometry as the output)
  // TMA epilogue: a ring of epi_boxes pairs of staging boxes (32 KB each; epi_bytes in all).  One pair serialises
  // residual load -> arithmetic -> store read per 64-column part (HBM latency exposed: the K <= 256 layers with a residual
  // ran at 0.55-0.6 of the copy bandwidth); with three, two residual loads and a store are in flight behind the arithmetic.
  int32_t epi_boxes, epi_bytes;
};

// PAIR = true: two CTAs of a cluster (one TPC) cooperate through tcgen05 cta_group::2 -- an M = 256 tile pair
// shares ONE weight tile, each CTA staging only half of it (N/2 rows), which halves the dominant L2->SMEM
// traffic of the wide layers.  The leader (even) CTA issues the MMAs for both; each CTA keeps its own 128 rows
// of the accumulator in its own TMEM and runs its own epilogue.
// TMAEPI = true (flat layers: the output plane IS the M space, split-fp16 output, bn a multiple of 64): the epilogue warps
// only do TMEM -> registers -> folded norm / ReLU / residual -> fp16 hi/lo into two 128-row x 64-column SWIZZLE_128B
// boxes in shared memory; one thread then issues cp.async.bulk.tensor STORES of the boxes, and the residual tile arrives
// by a TMA LOAD into the same boxes (added in place).  No per-row address arithmetic, no transposing shuffles, fully
// coalesced 128-byte rows.  Border rows of the padded plane are written as the zeros they already are.
template <bool SPLIT, bool PAIR, bool TMAEPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                 const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
                 const GemmKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A_hi, A_lo?, B_hi, B_lo?)] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stage_base = reinterpret_cast<float*>(smem + (size_t)p.ring_bytes);  // epilogue staging, 2 KB per warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.ring_bytes + (size_t)p.epi_bytes);
  uint64_t* full_bar = bars;                      // [MAX_STAGES]
  uint64_t* empty_bar = bars + MAX_STAGES;        // [MAX_STAGES]
  uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES;   // [MAX_ACC_BUFS]
  uint64_t* tmem_empty_bar = bars + 2 * MAX_STAGES + MAX_ACC_BUFS;  // [MAX_ACC_BUFS]
  uint64_t* a_full_bar = bars + 2 * MAX_STAGES + 2 * MAX_ACC_BUFS;                  // [MAX_A_STAGES]
  uint64_t* a_empty_bar = bars + 2 * MAX_STAGES + 2 * MAX_ACC_BUFS + MAX_A_STAGES;  // [MAX_A_STAGES]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 2 * MAX_ACC_BUFS + 2 * MAX_A_STAGES);
  uint64_t* res_bar = bars + 2 * MAX_STAGES + 2 * MAX_ACC_BUFS + 2 * MAX_A_STAGES + 1;   // [MAX_EPI_BOXES] residual boxes landed

  // Programmatic dependent launch: let the next kernel of the stream start its prologue as soon as SMs free up
  // (it blocks at its own griddepcontrol.wait until this grid has completed and flushed).
  asm volatile("griddepcontrol.launch_dependents;");
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kblocks = p.kblocks_per_tap * p.ntaps;
  // work decomposition: a "unit" is one (M tile, N tile) for a single CTA, or one (M tile PAIR, N tile) for a
  // CTA pair; units are strided over the CTAs (pairs) of the persistent grid
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int b_rows = PAIR ? p.bn / 2 : p.bn;  // weight-tile rows staged by this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (SPLIT) {
      tma_prefetch_desc(&map_a_lo);
      tma_prefetch_desc(&map_b_lo);
    }
    if (TMAEPI) {
      tma_prefetch_desc(&map_o_hi);
      tma_prefetch_desc(&map_o_lo);
      if (p.tma_res) {
        tma_prefetch_desc(&map_r_hi);
        tma_prefetch_desc(&map_r_lo);
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < MAX_A_STAGES; ++i) {
      mbar_init(&a_full_bar[i], 1);
      mbar_init(&a_empty_bar[i], 1);
    }
    for (int i = 0; i < MAX_ACC_BUFS; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < MAX_EPI_BOXES; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(tmem_holder, TMEM_COLS);
    else tmem_alloc(tmem_holder, TMEM_COLS);
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive / TMA signal
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // global memory written by it may only be touched from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // M extent: the host's worst case, or the live count a previous kernel of the stream left on the device (the number
  // of detected words) -- the persistent grid simply walks fewer tiles, nothing is re-planned on the host
  int64_t rows_m = p.rows_m;
  int tiles_m = p.tiles_m;
  if (p.m_count_dev != nullptr) {
    const int64_t live = (int64_t)max(*reinterpret_cast<const volatile int32_t*>(p.m_count_dev), 0) * p.m_rows_per_count;
    if (live < rows_m) {
      rows_m = live;
      tiles_m = (int)((rows_m + BM - 1) / BM);
    }
  }
  const int units_m = PAIR ? (tiles_m + 1) / 2 : tiles_m;
  const int num_units = units_m * p.tiles_n;

  // Register re-balancing between the warpgroups (the launch gives every thread 168): the control
  // warpgroup needs few registers, the two epilogue warpgroups hold the whole 128 x BN fp32 tile.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  }

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0 && p.group3) {
      int as = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      uint8_t* b_ring = smem + (size_t)p.a_stages * p.a_stage_bytes;
      for (int unit = worker; unit < num_units; unit += num_workers) {
        const int um = p.tiles_n == 1 ? unit : unit / p.tiles_n;
        const int tn = unit - um * p.tiles_n;
        const int m0 = (PAIR ? um * 2 + (int)rank : um) * BM;
        const int n0 = tn * p.bn + (int)rank * b_rows;
        for (int r = 0; r < 3; ++r) {
          const int arow = m0 + p.tap_shift[3 * r];  // first row of the block = rows of tap (r, s = 0)
          for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
            mbar_wait(&a_empty_bar[as], aphase ^ 1);
            uint8_t* sa = smem + (size_t)as * p.a_stage_bytes;
            if (leader) mbar_arrive_expect_tx(&a_full_bar[as], (uint32_t)p.a_stage_bytes * (PAIR ? 2u : 1u));
            if (PAIR) {
              tma_load_2d_pair(sa, &map_a_hi, &a_full_bar[as], kb * BK, arow);
              if (SPLIT) tma_load_2d_pair(sa + A_BLOCK_BYTES, &map_a_lo, &a_full_bar[as], kb * BK, arow);
            } else {
              tma_load_2d(sa, &map_a_hi, &a_full_bar[as], kb * BK, arow);
              if (SPLIT) tma_load_2d(sa + A_BLOCK_BYTES, &map_a_lo, &a_full_bar[as], kb * BK, arow);
            }
            if (++as == p.a_stages) {
              as = 0;
              aphase ^= 1;
            }
            for (int sx = 0; sx < 3; ++sx) {
              if (!(p.resident_b && unit != worker)) {   // resident weights: loaded by this CTA's first tile only
                mbar_wait(&empty_bar[bs], bphase ^ 1);
                uint8_t* sb = b_ring + (size_t)bs * p.b_stage_bytes;
                if (leader) mbar_arrive_expect_tx(&full_bar[bs], (uint32_t)p.b_stage_bytes * (PAIR ? 2u : 1u));
                const int kw = ((3 * r + sx) * p.kblocks_per_tap + kb) * BK;
                if (PAIR) {
                  tma_load_2d_pair(sb, &map_b_hi, &full_bar[bs], kw, n0);
                  if (SPLIT) tma_load_2d_pair(sb + p.b_tile_bytes, &map_b_lo, &full_bar[bs], kw, n0);
                } else {
                  tma_load_2d(sb, &map_b_hi, &full_bar[bs], kw, n0);
                  if (SPLIT) tma_load_2d(sb + p.b_tile_bytes, &map_b_lo, &full_bar[bs], kw, n0);
                }
              }
              if (++bs == p.b_stages) {
                bs = 0;
                bphase ^= 1;
              }
            }
          }
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = worker; unit < num_units; unit += num_workers) {
        const int um = p.tiles_n == 1 ? unit : unit / p.tiles_n;
        const int tn = unit - um * p.tiles_n;
        const int m0 = (PAIR ? um * 2 + (int)rank : um) * BM;
        const int n0 = tn * p.bn + (int)rank * b_rows;  // this CTA's half of the weight tile
        for (int t = 0; t < p.ntaps; ++t) {
          const int arow = m0 + p.tap_shift[t];
          for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* s = smem + (size_t)stage * p.stage_bytes;
            // the leader's barrier collects the bytes of BOTH CTAs' loads
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.stage_bytes * (PAIR ? 2u : 1u));
            const int ka = p.a_col0 + kb * BK;
            const int kw = (t * p.kblocks_per_tap + kb) * BK;
            auto load = [&](void* dst, const CUtensorMap* map, int c0, int c1) {
              if (PAIR) tma_load_2d_pair(dst, map, &full_bar[stage], c0, c1);
              else tma_load_2d(dst, map, &full_bar[stage], c0, c1);
            };
            load(s, &map_a_hi, ka, arow);
            if (SPLIT) {
              load(s + A_TILE_BYTES, &map_a_lo, ka, arow);
              load(s + 2 * A_TILE_BYTES, &map_b_hi, kw, n0);
              load(s + 2 * A_TILE_BYTES + p.b_tile_bytes, &map_b_lo, kw, n0);
            } else {
              load(s + A_TILE_BYTES, &map_b_hi, kw, n0);
            }
            if (++stage == p.num_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // The whole warp walks the (warp-uniform) loop so that descriptors and addresses stay in uniform registers;
    // only the elected lane issues tcgen05.mma / tcgen05.commit.  With one lane doing everything the issue path
    // costs ~80 clk per MMA (R2UR traffic), which starves tiles narrower than 256 columns (32-64 clk per MMA).
    if (leader) {
      const uint32_t idesc = umma_idesc_f16_f32(PAIR ? 2 * BM : BM, p.bn);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t chunk = 0;  // running chunk counter of this CTA: TMEM buffer = chunk & 1
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if (PAIR) umma_f16_ss_pair(d, da, db, idesc, acc);
        else umma_f16_ss(d, da, db, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if (PAIR) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t b_ring = smem_u32(smem) + (uint32_t)(p.group3 ? p.a_stages * p.a_stage_bytes : 0);
      for (int unit = worker; unit < num_units; unit += num_workers) {
        uint32_t d_tmem = 0;
        uint32_t a_block = 0;  // tap-row mode: shared-memory address of the current activation block
        for (int kb = 0; kb < kblocks; ++kb) {
          const int kin = kb % p.kb_per_chunk;  // position inside the accumulation chunk
          if (kin == 0) {
            // a new chunk starts from zero in the next TMEM buffer once the epilogue has drained it
            const uint32_t buf = chunk & (uint32_t)(p.acc_bufs - 1);
            mbar_wait(&tmem_empty_bar[buf], ((chunk >> p.acc_shift) & 1u) ^ 1u);
            tcgen05_fence_after();
            d_tmem = tmem_base + buf * (uint32_t)p.acc_cols;
          }
          // operand addresses of this k-block
          uint32_t sa_hi, sa_lo, sb_hi, sb_lo;
          int sx = 0;
          if (p.group3) {
            sx = kb % 3;  // k-block order: (tap row r, channel block) outer, s-tap inner
            if (sx == 0) {
              mbar_wait(&a_full_bar[as], aphase);
              a_block = smem_u32(smem) + (uint32_t)(as * p.a_stage_bytes);
            }
            if (!(p.resident_b && unit != worker)) mbar_wait(&full_bar[stage], phase);   // (resident: landed with the first tile)
            sa_hi = a_block + (uint32_t)(sx * BK * 2);  // descriptor start advanced by sx rows of 128 bytes
            sa_lo = sa_hi + A_BLOCK_BYTES;
            sb_hi = b_ring + (uint32_t)(stage * p.b_stage_bytes);
            sb_lo = sb_hi + (uint32_t)p.b_tile_bytes;
          } else {
            mbar_wait(&full_bar[stage], phase);
            const uint32_t s0 = smem_u32(smem + (size_t)stage * p.stage_bytes);
            sa_hi = s0;
            sa_lo = s0 + A_TILE_BYTES;
            sb_hi = s0 + (SPLIT ? 2 : 1) * A_TILE_BYTES;
            sb_lo = sb_hi + (uint32_t)p.b_tile_bytes;
          }
          tcgen05_fence_after();
          const uint64_t da_hi = umma_smem_desc_sw128(sa_hi), da_lo = umma_smem_desc_sw128(sa_lo);
          const uint64_t db_hi = umma_smem_desc_sw128(sb_hi), db_lo = umma_smem_desc_sw128(sb_lo);
          const bool a_done = p.group3 && sx == 2;
          const bool chunk_done = kin == p.kb_per_chunk - 1 || kb == kblocks - 1;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t koff = (uint64_t)(k * 2);  // 32 bytes >> 4
              const uint32_t acc_flag = (kin > 0 || k > 0) ? 1u : 0u;
              if (SPLIT) {
                mma(d_tmem, da_lo + koff, db_hi + koff, acc_flag);
                mma(d_tmem, da_hi + koff, db_lo + koff, 1u);
                mma(d_tmem, da_hi + koff, db_hi + koff, 1u);
              } else {
                mma(d_tmem, da_hi + koff, db_hi + koff, acc_flag);
              }
            }
            if (!p.resident_b) commit(&empty_bar[stage]);  // frees the (weight) slot in both CTAs when these MMAs retire
            if (a_done) commit(&a_empty_bar[as]);  // all three s-taps have consumed the activation block
            if (chunk_done) commit(&tmem_full_bar[chunk & (uint32_t)(p.acc_bufs - 1)]);  // -> the epilogue(s) drain it
          }
          __syncwarp();
          if (a_done && ++as == p.a_stages) {
            as = 0;
            aphase ^= 1;
          }
          if (chunk_done) ++chunk;
          if (++stage == (p.group3 ? p.b_stages : p.num_stages)) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    if constexpr (TMAEPI) {
      // ===================================================== epilogue through TMA (16 warps)
      // warp e = warp-4: TMEM lane quadrant q = e & 3 -> tile rows 32q..32q+31 (one per lane); sub = e >> 2 -> the
      // 16-column chunk `sub` of every 64-column PART of the tile.  The tile is finished part by part through a ring
      // of p.epi_boxes pairs of staging boxes (hi / lo, 16 KB each; one pair = the 32 KB the classic epilogue uses as
      // its transpose stage).
      const int e = warp - 4;
      const int q = e & 3;
      const int sub = e >> 2;
      const int row_in_tile = q * 32 + lane;
      const int plane = p.m_h * p.m_w;
      const int parts = p.bn >> 6;
      const bool has_res = p.tma_res != 0;
      const bool issuer = threadIdx.x == 128;   // warp 4, lane 0: issues the TMA stores / residual loads
      // Everything below is addressed through 32-bit shared-space addresses derived from ONE laundered base, and the
      // per-thread offsets are laundered too: left to itself the compiler re-derived them from %tid / the generic smem
      // pointer (cvta) at every use inside the register-starved part loop -- ~300 of the ~550 warp instructions per part.
      const uint32_t sbase = opaque_u32(smem_u32(smem));
      const uint32_t boxes = sbase + (uint32_t)p.ring_bytes;   // ring of p.epi_boxes pairs: hi box at + 32 KB * b, lo box 16 KB behind it
      const uint32_t bars_a = boxes + (uint32_t)p.epi_bytes;
      const uint32_t tmem_full_a = bars_a + 8u * (2 * MAX_STAGES);
      const uint32_t tmem_empty_a = tmem_full_a + 8u * MAX_ACC_BUFS;
      const uint32_t res_bar_a = bars_a + 8u * (2 * MAX_STAGES + 2 * MAX_ACC_BUFS + 2 * MAX_A_STAGES + 1);
      // this thread's two 16-byte chunks of its 128-byte box row (SWIZZLE_128B: chunk index ^ (row & 7))
      const uint32_t row_off = (uint32_t)row_in_tile * 128u;
      const uint32_t off0 = opaque_u32(row_off + (uint32_t)(((2 * sub) ^ (lane & 7)) << 4));
      const uint32_t off1 = opaque_u32(row_off + (uint32_t)(((2 * sub + 1) ^ (lane & 7)) << 4));
      const uint32_t col_off = opaque_u32((uint32_t)sub * 16u);   // first of this thread's 16 columns inside a part
      const int chunks_per_tile = (kblocks + p.kb_per_chunk - 1) / p.kb_per_chunk;
      uint32_t chunk = 0;
      const int nb = p.epi_boxes;
      const int ahead = nb > 1 ? nb - 1 : 1;   // residual loads run `ahead` parts in front of the arithmetic
      int box = 0;                             // ring position of the current part
      uint32_t box_phase = 0;                  // parity of res_bar[box] for the current lap
      auto unit_origin = [&](int unit, int& m0, int& n0) {
        const int um = p.tiles_n == 1 ? unit : unit / p.tiles_n;
        const int tn = unit - um * p.tiles_n;
        m0 = (PAIR ? um * 2 + (int)rank : um) * BM;
        n0 = tn * p.bn;
      };
      // issuer only: the part the next residual load is for, and the box it goes to
      int la_unit = worker, la_pt = 0, la_box = 0;
      auto res_load_next = [&]() {
        if (la_unit < num_units) {
          int rm0, rn0;
          unit_origin(la_unit, rm0, rn0);
          const uint32_t dst = boxes + (uint32_t)la_box * (uint32_t)EPI_BOX_BYTES;
          const uint32_t bar = res_bar_a + 8u * (uint32_t)la_box;
          mbar_arrive_expect_tx_a(bar, (uint32_t)EPI_BOX_BYTES);
          tma_load_2d_a(dst, &map_r_hi, bar, rn0 + la_pt * 64, rm0);
          tma_load_2d_a(dst + (uint32_t)(EPI_BOX_BYTES / 2), &map_r_lo, bar, rn0 + la_pt * 64, rm0);
        }
        if (++la_pt == parts) {
          la_pt = 0;
          la_unit += num_workers;
        }
        if (++la_box == nb) la_box = 0;
      };
      if (has_res && issuer) {   // the first parts' residuals: no dependence on the MMA
        for (int i = 0; i < ahead; ++i) res_load_next();
      }
      for (int unit = worker; unit < num_units; unit += num_workers) {
        int m0, n0;
        unit_origin(unit, m0, n0);
        const int64_t m = (int64_t)m0 + row_in_tile;
        bool valid = m < rows_m;
        if (valid && (p.m_border | p.m_border_hi) != 0) {   // (planes without a border: every row is a pixel)
          const uint32_t rem = (uint32_t)m % (uint32_t)plane;   // rows_m < 2^31 (checked by the host): 32-bit division
          const int yy = (int)(rem / (uint32_t)p.m_w);
          const int xx = (int)rem - yy * p.m_w;
          valid = (yy >= p.m_border) && (xx >= p.m_border) && (yy < p.m_h - p.m_border_hi) && (xx < p.m_w - p.m_border_hi);
        }
        // ---- drain every accumulation chunk of this tile from TMEM into fp32 registers (round-to-nearest adds)
        float accv[MAX_CHUNKS_PER_WARP][16];
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        for (int g = 0; g < chunks_per_tile; ++g, ++chunk) {
          const uint32_t buf = chunk & (uint32_t)(p.acc_bufs - 1);
          mbar_wait_a(tmem_full_a + 8u * buf, (chunk >> p.acc_shift) & 1u);
          tcgen05_fence_after();
          const uint32_t t_row = tmem_base + buf * (uint32_t)p.acc_cols + lane_sel + col_off;
#pragma unroll
          for (int pt = 0; pt < MAX_CHUNKS_PER_WARP; ++pt) {
            if (pt < parts) {
              uint32_t r[16];
              tmem_ld_32x32b_x16(t_row + (uint32_t)(pt * 64), r);
              tmem_ld_wait();
              if (g == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) accv[pt][j] = __uint_as_float(r[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) accv[pt][j] += __uint_as_float(r[j]);
              }
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_remote_a(tmem_empty_a + 8u * buf, 0u);
            else mbar_arrive_a(tmem_empty_a + 8u * buf);
          }
        }
        // ---- part by part: math in registers -> staging boxes -> TMA store.  All arithmetic runs in the STORAGE scale
        // (kActScale * y, a power of two: the same bits as scaling at the end): y16 = D * (16 scale) + 16 bias; the
        // residual planes already hold 16 r.
#pragma unroll
        for (int pt = 0; pt < MAX_CHUNKS_PER_WARP; ++pt) {
          if (pt < parts) {
            const uint32_t n = (uint32_t)(n0 + pt * 64) + col_off;
            float v[16];
            // y16 = 16 (D scale + bias): one fused multiply-add per element like the direct-store epilogue, then the exact
            // power-of-two factor (the same bits as fma(D, 16 scale, 16 bias), half the multiplies)
            if (p.scale != nullptr && p.bias != nullptr) {
              const float4* sp = reinterpret_cast<const float4*>(p.scale + n);
              const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 s4 = __ldg(sp + (j >> 2)), b4 = __ldg(bp + (j >> 2));
                v[j] = fmaf(accv[pt][j], s4.x, b4.x) * kActScale; v[j + 1] = fmaf(accv[pt][j + 1], s4.y, b4.y) * kActScale;
                v[j + 2] = fmaf(accv[pt][j + 2], s4.z, b4.z) * kActScale; v[j + 3] = fmaf(accv[pt][j + 3], s4.w, b4.w) * kActScale;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.scale != nullptr) s4 = __ldg(reinterpret_cast<const float4*>(p.scale + n + j));
                if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                v[j] = fmaf(accv[pt][j], s4.x, b4.x) * kActScale; v[j + 1] = fmaf(accv[pt][j + 1], s4.y, b4.y) * kActScale;
                v[j + 2] = fmaf(accv[pt][j + 2], s4.z, b4.z) * kActScale; v[j + 3] = fmaf(accv[pt][j + 3], s4.w, b4.w) * kActScale;
              }
            }
            if (p.relu_pre) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            const uint32_t box_hi = boxes + (uint32_t)box * (uint32_t)EPI_BOX_BYTES, box_lo = box_hi + (uint32_t)(EPI_BOX_BYTES / 2);
            if (has_res) {
              mbar_wait_a(res_bar_a + 8u * (uint32_t)box, box_phase);   // this part's residual has landed in its pair of boxes
              const uint4 rh0 = lds_u4(box_hi + off0), rh1 = lds_u4(box_hi + off1);
              const uint4 rl0 = lds_u4(box_lo + off0), rl1 = lds_u4(box_lo + off1);
              const uint32_t aw[8] = {rh0.x, rh0.y, rh0.z, rh0.w, rh1.x, rh1.y, rh1.z, rh1.w};
              const uint32_t bw[8] = {rl0.x, rl0.y, rl0.z, rl0.w, rl1.x, rl1.y, rl1.z, rl1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 rv = unpack16x2_scaled(aw[j], bw[j]);
                v[2 * j] += rv.x;
                v[2 * j + 1] += rv.y;
              }
            }
            if (p.relu_post) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.sat_count != nullptr && valid) {   // debug only (GLASS_DEBUG_SAT): silent saturation made visible
              int over = 0;
#pragma unroll
              for (int j = 0; j < 16; ++j) over += fabsf(v[j]) > 60000.f;
              if (over) atomicAdd(p.sat_count, over);
            }
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)   // border rows / rows past the M extent: zeros
              split16x2_scaled(valid ? v[2 * j] : 0.f, valid ? v[2 * j + 1] : 0.f, hw[j], lw[j]);
            if (!has_res && nb == 1) {
              // single pair, no residual: the boxes are free once the previous part's stores have READ them (the issuer
              // waits only now, after this part's math: the store's shared-memory reads overlap the arithmetic)
              if (issuer) tma_store_wait_read();
              named_bar_sync(2, 32 * EPI_WARPS);
            }
            sts_u4(box_hi + off0, make_uint4(hw[0], hw[1], hw[2], hw[3]));
            sts_u4(box_hi + off1, make_uint4(hw[4], hw[5], hw[6], hw[7]));
            sts_u4(box_lo + off0, make_uint4(lw[0], lw[1], lw[2], lw[3]));
            sts_u4(box_lo + off1, make_uint4(lw[4], lw[5], lw[6], lw[7]));
            fence_proxy_async_smem();
            // ring without residual: the NEXT part's pair was last read by the store nb - 1 parts back; the issuer makes
            // sure of that before it joins the barrier, so passing the barrier also means "the next pair is free"
            if (!has_res && nb > 1 && issuer) tma_store_wait_read_but(nb - 2);
            named_bar_sync(1, 32 * EPI_WARPS);   // this pair of boxes is complete
            if (issuer) {
              tma_store_2d(&map_o_hi, box_hi, n0 + pt * 64, m0);
              tma_store_2d(&map_o_lo, box_lo, n0 + pt * 64, m0);
              tma_store_commit();
              if (has_res) {
                // the residual `ahead` parts on goes into the pair the PREVIOUS part's stores are reading (this part's
                // own pair when the ring has one); the other threads cannot touch a pair before its residual has landed
                // (they wait on res_bar), so no second barrier
                tma_store_wait_read_but(nb > 1 ? 1 : 0);
                res_load_next();
              }
            }
            if (++box == nb) {
              box = 0;
              box_phase ^= 1u;
            }
          }
        }
      }
      if (issuer) tma_store_wait_all();
    } else {
    // ===================================================== epilogue (16 warps, direct stores)
    // warp e = warp-4: TMEM lane quadrant q = e & 3 (== warp % 4, the tcgen05.ld lane rule), column half
    // e >> 2.  Each thread owns one output row of the tile and walks its half of the 16-column chunks.
    // The residual does not depend on the MMA, so its loads run two chunks ahead (and, for the first
    // chunks of a tile, are issued before the accumulator is even ready).
    const int e = warp - 4;
    const int q = e & 3;
    const int part = e >> 2;  // which quarter of the tile's columns
    const int row_in_tile = q * 32 + lane;
    const int plane = p.m_h * p.m_w;
    const int nchunks = p.bn / 16;
    const int cbase = nchunks >> 2, crem = nchunks & 3;
    const int c_begin = part * cbase + min(part, crem);
    const int c_end = c_begin + cbase + (part < crem ? 1 : 0);
    const bool has_res = p.res_hi != nullptr;
    const uint32_t stage = smem_u32(stage_base) + (uint32_t)e * 2048u;   // this warp's 32 rows x 16 fp32 staging block
    const int chunks_per_tile = (kblocks + p.kb_per_chunk - 1) / p.kb_per_chunk;
    uint32_t chunk = 0;  // mirrors the MMA warp's running chunk counter
    for (int unit = worker; unit < num_units; unit += num_workers) {
      const int um = p.tiles_n == 1 ? unit : unit / p.tiles_n;
      const int tn = unit - um * p.tiles_n;
      const int tm = PAIR ? um * 2 + (int)rank : um;
      const int n0 = tn * p.bn;
      const int64_t m = (int64_t)tm * BM + row_in_tile;
      bool valid = m < rows_m;
      int64_t out_row = 0, res_row = 0;
      if (valid) {
        const int img = (int)((uint32_t)m / (uint32_t)plane);   // rows_m < 2^31 (checked by the host): 32-bit division
        const int rem = (int)((uint32_t)m - (uint32_t)img * (uint32_t)plane);
        const int yy = rem / p.m_w;
        const int xx = rem - yy * p.m_w;
        const int y = yy - p.m_border, x = xx - p.m_border;
        valid = (y >= 0) && (x >= 0) && (yy < p.m_h - p.m_border_hi) && (xx < p.m_w - p.m_border_hi);
        out_row = ((int64_t)img * p.out_hp + y + p.out_border) * p.out_wp + x + p.out_border;
        res_row = ((int64_t)img * p.res_hp + (y >> p.res_shift) + p.res_border) * p.res_wp + (x >> p.res_shift) +
                  p.res_border;
      }
      // the residual does not depend on the MMA: pull this lane's row segment into L2 now, so the coalesced
      // loads of the final phase find it there instead of waiting on HBM
      if (has_res && valid && c_begin < c_end && n0 + c_begin * 16 < p.n_store) {
        const __half* rp_hi = p.res_hi + res_row * p.ld_out + n0 + c_begin * 16;
        const __half* rp_lo = p.res_lo + res_row * p.ld_out + n0 + c_begin * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp_hi));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp_lo));
      }

      // ---- drain every accumulation chunk of this tile from TMEM into fp32 registers (round-to-nearest
      // adds): the tensor core's own accumulator truncates, so chains are kept to kb_per_chunk k-blocks.
      float accv[MAX_CHUNKS_PER_WARP][16];
      const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
      for (int g = 0; g < chunks_per_tile; ++g, ++chunk) {
        const uint32_t buf = chunk & (uint32_t)(p.acc_bufs - 1);
        mbar_wait(&tmem_full_bar[buf], (chunk >> p.acc_shift) & 1u);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + buf * (uint32_t)p.acc_cols + lane_sel;
#pragma unroll
        for (int ci = 0; ci < MAX_CHUNKS_PER_WARP; ++ci) {
          if (c_begin + ci < c_end) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(t_row + (uint32_t)((c_begin + ci) * 16), r);
            tmem_ld_wait();
            if (g == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) accv[ci][j] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) accv[ci][j] += __uint_as_float(r[j]);
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_remote(&tmem_empty_bar[buf], 0u);  // the leader's MMA thread owns the wait
          else mbar_arrive(&tmem_empty_bar[buf]);
        }
      }

      // ---- final math + stores.  Each lane owns one output ROW in registers, but a warp store that
      // writes 32 different rows touches 32 cache lines (LSU-bound, measured 31 sectors/request), so
      // the tile goes through a per-warp XOR-swizzled shared-memory stage, 32 columns at a time, and is
      // written back in the transposed mapping: 4 lanes x 16 B = 64 contiguous bytes of one row per
      // plane, 8 rows per instruction.  The residual is read in that same coalesced mapping.
#pragma unroll
      for (int ci = 0; ci < MAX_CHUNKS_PER_WARP; ++ci) {
        const int c = c_begin + ci;  // 16-column chunk (warp-uniform)
        if (c < c_end) {
          {
            const int n = n0 + c * 16;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = accv[ci][j];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {   // folded norm: ONE fused multiply-add per element (y = D * scale + bias)
              float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.scale != nullptr) s4 = __ldg(reinterpret_cast<const float4*>(p.scale + n + j));
              if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              v[j] = fmaf(v[j], s4.x, b4.x); v[j + 1] = fmaf(v[j + 1], s4.y, b4.y);
              v[j + 2] = fmaf(v[j + 2], s4.z, b4.z); v[j + 3] = fmaf(v[j + 3], s4.w, b4.w);
            }
            if (p.relu_pre) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int phys = jj ^ ((lane >> 1) & 3);
              sts_f4(stage + (uint32_t)(lane * 16 + phys * 4) * 4u, v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
            }
          }
          __syncwarp();
          const int cg = lane & 1;  // 8-column half of the 16-column stage
          const int n = n0 + c * 16 + cg * 8;
#pragma unroll
          for (int ps = 0; ps < 2; ++ps) {
            const int R = ps * 16 + (lane >> 1);
            const int valid_r = __shfl_sync(0xffffffffu, (int)valid, R);
            const long long out_row_r = __shfl_sync(0xffffffffu, (long long)out_row, R);
            const long long res_row_r = __shfl_sync(0xffffffffu, (long long)res_row, R);
            if (valid_r && n < p.n_store) {
              const int sw = (R >> 1) & 3;
              const float4 a = lds_f4(stage + (uint32_t)(R * 16 + (((2 * cg) ^ sw) * 4)) * 4u);
              const float4 b = lds_f4(stage + (uint32_t)(R * 16 + (((2 * cg + 1) ^ sw) * 4)) * 4u);
              float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
              if (has_res) {
                const uint4 rh = __ldg(reinterpret_cast<const uint4*>(p.res_hi + res_row_r * p.ld_out + n));
                const uint4 rl = __ldg(reinterpret_cast<const uint4*>(p.res_lo + res_row_r * p.ld_out + n));
                const uint32_t aw[4] = {rh.x, rh.y, rh.z, rh.w}, bw[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 rv = unpack16x2(aw[j], bw[j]);
                  v[2 * j] += rv.x;
                  v[2 * j + 1] += rv.y;
                }
              }
              if (p.relu_post) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if (p.out_f32 != nullptr) {
                float4* o = reinterpret_cast<float4*>(p.out_f32 + out_row_r * p.ld_f32 + n);
                o[0] = make_float4(v[0], v[1], v[2], v[3]);
                o[1] = make_float4(v[4], v[5], v[6], v[7]);
              }
              if (p.out_hi != nullptr) {
                if (p.sat_count != nullptr) {   // debug only (GLASS_DEBUG_SAT)
                  int over = 0;
#pragma unroll
                  for (int j = 0; j < 8; ++j) over += fabsf(v[j]) * kActScale > 60000.f;
                  if (over) atomicAdd(p.sat_count, over);
                }
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  __half h0, l0, h1, l1;
                  split16(v[2 * j], h0, l0);
                  split16(v[2 * j + 1], h1, l1);
                  hw[j] = pack16x2(h0, h1);
                  lw[j] = pack16x2(l0, l1);
                }
                *reinterpret_cast<uint4*>(p.out_hi + out_row_r * p.ld_out + n) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                if (p.out_lo != nullptr)
                  *reinterpret_cast<uint4*>(p.out_lo + out_row_r * p.ld_out + n) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
            }
          }
          __syncwarp();
        }
      }
    }
    }  // classic epilogue
  }

  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host
static int make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                       uint32_t box_outer, uint64_t row_stride_elems = 0) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {(row_stride_elems ? row_stride_elems : inner) * 2};  // bytes, dim 1 (may overlap)
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

}  // namespace glass

using namespace glass;

// A launch plan: everything the host derives from a GlassConvGemmParams -- validation, the four TMA descriptors, the
// tile / pipeline geometry, the grid -- so that a caller that repeats a launch (same buffers every step) pays for it
// once (glass_plan_create) and the per-step cost is one cudaLaunchKernelEx (glass_plan_launch).
struct GlassGemmPlan {
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  CUtensorMap mo_hi, mo_lo, mr_hi, mr_lo;   // TMA epilogue: output / residual boxes (64 columns x 128 rows)
  GemmKernelParams k;
  int split, pair, tma_epi, grid, smem_req;
};

static int build_plan(const GlassConvGemmParams* p, GlassGemmPlan* plan) {
  GLASS_CHECK(p != nullptr, "null params");
  GLASS_CHECK(p->a_hi && p->b_hi, "a_hi / b_hi must be set");
  const bool split = (p->mode == 0);
  GLASS_CHECK(p->mode == 0 || p->mode == 1, "mode must be 0 (fp16x3 split) or 1 (single fp16 pass)");
  if (split) GLASS_CHECK(p->a_lo && p->b_lo, "split mode needs a_lo / b_lo");
  GLASS_CHECK(p->k_per_tap > 0 && p->k_per_tap % BK == 0, "k_per_tap must be a positive multiple of 64");
  GLASS_CHECK(p->ntaps >= 1 && p->ntaps <= GLASS_MAX_TAPS, "ntaps out of range");
  GLASS_CHECK(p->n >= 16 && p->n % 16 == 0, "n must be a multiple of 16");
  int bn = p->n;
  if (p->n > 256) {
    GLASS_CHECK(p->n % 128 == 0, "n > 256 must be a multiple of 128");
    bn = (p->n % 256 == 0) ? 256 : 128;
  }
  const int64_t rows_m = (int64_t)p->m_imgs * p->m_h * p->m_w;
  // Small-M GEMMs (the box head's FCs: 400 rows x 2048 columns x K = 12544; the RPN head on p5 / p6) would occupy a
  // fraction of the 148 SMs with 256-wide tiles, and each of those CTAs is bound by its own TMA feed: narrower tiles
  // put more SMs to work and shrink the weight stream per CTA (per-output arithmetic does not depend on the tile width).
  {
    const int sms_now = num_sms();
    const int64_t tm = (rows_m + BM - 1) / BM;
    while (bn > 64 && bn % 2 == 0 && (bn / 2) % 16 == 0 && p->n % (bn / 2) == 0 && tm * (p->n / bn) * 2 <= sms_now) bn /= 2;
  }
  GLASS_CHECK(rows_m > 0 && rows_m < (int64_t)1 << 31, "bad M space");
  GLASS_CHECK(p->rows_a > 0 && p->rows_a < (int64_t)1 << 31, "bad rows_a");
  const int m_b = GLASS_BORDER_LO(p->m_border), m_bh = GLASS_BORDER_HI(p->m_border);
  GLASS_CHECK(p->m_border >= 0 && m_b + m_bh < p->m_h && m_b + m_bh < p->m_w, "bad m_border");
  GLASS_CHECK(p->out_hi || p->out_f32, "no output requested");
  const int n_store = p->n_store > 0 ? p->n_store : p->n;
  GLASS_CHECK(n_store <= p->n, "n_store > n");
  if (p->out_hi || p->res_hi) GLASS_CHECK(p->ld_out >= n_store && p->ld_out % 8 == 0, "ld_out must be >= n_store, multiple of 8");
  if (p->out_f32) GLASS_CHECK(p->ld_f32 >= n_store && p->ld_f32 % 4 == 0, "ld_f32 must be >= n_store, multiple of 4");
  GLASS_CHECK(n_store % 16 == 0, "n_store must be a multiple of 16");
  if (p->res_hi) GLASS_CHECK(p->res_lo != nullptr, "res_lo missing");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  GLASS_CHECK(al16(p->a_hi) && al16(p->a_lo) && al16(p->b_hi) && al16(p->b_lo) && al16(p->out_hi) &&
                  al16(p->out_lo) && al16(p->out_f32) && al16(p->res_hi) && al16(p->res_lo) && al16(p->scale) &&
                  al16(p->bias),
              "all pointers must be 16-byte aligned");

  // CTA-pair (cta_group::2) mode: worth it when the weight tile is wide and there is enough work to pair up
  const int sms = num_sms();
  GLASS_CHECK(sms > 0, "no CUDA device");
  const int tiles_m_all = (int)((rows_m + BM - 1) / BM);
  bool pair = (bn % 32 == 0) && bn >= 64 && (int64_t)tiles_m_all * (p->n / bn) >= sms;
  // Small layers (a handful of tile rounds: res5, the p5 / p6 ends of FPN and RPN): the rule above can land just past a
  // round boundary -- 37 M tiles x 4 N tiles = 76 pair units on 74 pairs = TWO rounds where 148 single tiles fill the 148
  // SMs in one.  For them the tile width and the pairing are chosen together by rounds x (tile width / the width's
  // measured MMA efficiency: 0.34 / 0.62 / 0.90 at 64 / 128 / 256 columns), pairs 3 % ahead for their shared weight tile.
  static const int tile_model = getenv("GLASS_TILE_MODEL") ? atoi(getenv("GLASS_TILE_MODEL")) : 1;
  if (tile_model && p->pair_mode != 2) {
    auto rounds = [&](int w, bool pr) {
      const int64_t um = pr ? (tiles_m_all + 1) / 2 : tiles_m_all;
      const int64_t units = um * (p->n / w);
      const int workers = pr ? sms / 2 : sms;
      return (int)((units + workers - 1) / workers);
    };
    auto cost = [&](int w, bool pr) {
      const double eff = w >= 256 ? 0.90 : (w >= 128 ? 0.62 + (w - 128) * (0.28 / 128.0) : 0.34 + (w - 64) * (0.28 / 64.0));
      return rounds(w, pr) * (w / (eff > 0.1 ? eff : 0.1)) * (pr ? 0.97 : 1.0);
    };
    if (rounds(bn, pair) <= 3) {
      int best_bn = bn;
      bool best_pair = pair;
      double best = cost(bn, pair);
      int w = p->n > 256 ? ((p->n % 256 == 0) ? 256 : 128) : p->n;   // widest admissible tile, then its halvings
      for (;;) {
        for (int pr = 0; pr < 2; ++pr) {
          if (pr && (p->pair_mode == 1 || w % 32 != 0 || w < 64)) continue;
          const double c = cost(w, pr != 0);
          if (c < best - 1e-9) {
            best = c;
            best_bn = w;
            best_pair = pr != 0;
          }
        }
        if (!(w > 64 && w % 2 == 0 && (w / 2) % 16 == 0 && p->n % (w / 2) == 0)) break;
        w /= 2;
      }
      bn = best_bn;
      pair = best_pair;
    }
  }
  if (p->pair_mode == 1) pair = false;
  if (p->pair_mode == 2) {
    GLASS_CHECK(bn % 32 == 0, "pair mode needs a tile width that is a multiple of 32");
    pair = true;
  }
  const int b_rows = pair ? bn / 2 : bn;

  CUtensorMap& ma_hi = plan->ma_hi;
  CUtensorMap& ma_lo = plan->ma_lo;
  CUtensorMap& mb_hi = plan->mb_hi;
  CUtensorMap& mb_lo = plan->mb_lo;
  const uint64_t ktot = (uint64_t)p->k_per_tap * p->ntaps;
  // compact-channel mode: rows overlap (stride a_ld < 64 elements); the last 64/a_ld - 1 rows would read past
  // the tensor, so they are left to the TMA's out-of-bounds zero fill (they start inside the zero border)
  const int a_ld = p->a_ld > 0 ? p->a_ld : p->k_per_tap;
  uint64_t a_rows = (uint64_t)p->rows_a;
  const bool grouped = p->a_inner > 0;
  int a_inner = p->k_per_tap;
  if (grouped) {
    // pixel-grouped mode: rows_a counts GROUP rows of a_ld elements; the tensor row is a window of a_inner elements
    GLASS_CHECK(p->a_ld > 0 && a_ld % 8 == 0 && p->a_col0 >= 0 && p->a_col0 % 8 == 0, "grouped mode: a_ld / a_col0 must be multiples of 8");
    GLASS_CHECK(p->a_inner >= p->a_col0 + p->k_per_tap && p->a_inner % 8 == 0, "grouped mode: a_inner too small");
    a_inner = p->a_inner;
    const int64_t total = p->rows_a * (int64_t)a_ld;
    GLASS_CHECK(total > a_inner, "too few rows for grouped mode");
    a_rows = (uint64_t)((total - a_inner) / a_ld + 1);  // later rows would read past the tensor: TMA zero fill
  } else if (a_ld != p->k_per_tap) {
    GLASS_CHECK(p->k_per_tap == BK && (a_ld == 8 || a_ld == 16 || a_ld == 32), "compact mode needs k_per_tap 64, a_ld 8/16/32");
    GLASS_CHECK(p->rows_a > BK / a_ld, "too few rows for compact mode");
    a_rows = (uint64_t)p->rows_a - (uint64_t)(BK / a_ld) + 1;
  }
  // TMA epilogue: flat layers only -- the output (and residual) plane IS the M space, so tile rows are contiguous
  // tensor rows and one 2-D box per 64 columns covers them
  static const int tma_env = getenv("GLASS_TMA_EPI") ? atoi(getenv("GLASS_TMA_EPI")) : 1;
  const bool flat_out = split && p->out_hi && p->out_lo && !p->out_f32 && p->out_hp == p->m_h && p->out_wp == p->m_w &&
                        p->out_border == p->m_border && bn % 64 == 0 && p->ld_out >= n_store &&
                        (n_store == p->n || !p->res_hi);   // pad columns (n_store < n): clipped by the TMA store
  const bool flat_res = !p->res_hi || (p->res_shift == 0 && p->res_hp == p->m_h && p->res_wp == p->m_w &&
                                       p->res_border == p->m_border);
  bool tma_epi = flat_out && flat_res && tma_env != 0 && p->epi_mode != 1;
  if (p->epi_mode == 2) GLASS_CHECK(flat_out && flat_res, "epi_mode 2 (TMA epilogue) needs a flat split-fp16 output (and residual)");
  // staging ring of the TMA epilogue: the layers with at most 4 k-blocks (1x1 convs with K <= 256: HBM-bound) trade
  // pipeline stages they do not need for two more pairs of boxes (one more without a residual)
  const int split_mul = split ? 2 : 1;
  int epi_boxes = 1;
  if (tma_epi) {
    static const int nb_small = getenv("GLASS_EPI_BOXES_SMALLK") ? atoi(getenv("GLASS_EPI_BOXES_SMALLK")) : 0;
    static const int nb_big = getenv("GLASS_EPI_BOXES") ? atoi(getenv("GLASS_EPI_BOXES")) : 1;
    static const int small_kb = getenv("GLASS_EPI_SMALLK") ? atoi(getenv("GLASS_EPI_SMALLK")) : 4;
    const int kblocks_all = p->ntaps * (p->k_per_tap / BK);
    if (kblocks_all <= small_kb) epi_boxes = nb_small > 0 ? nb_small : (p->res_hi ? 3 : 2);
    else epi_boxes = p->res_hi ? nb_big : 1;
    if (epi_boxes < 1) epi_boxes = 1;
    if (epi_boxes > MAX_EPI_BOXES) epi_boxes = MAX_EPI_BOXES;
    // never at the price of a ring below two stages
    const int stage_b = (A_TILE_BYTES + (pair ? bn / 2 : bn) * BK * 2) * split_mul;
    while (epi_boxes > 1 && (227 * 1024 - epi_boxes * EPI_BOX_BYTES - 1536) / stage_b < 2) --epi_boxes;
  }
  const int epi_bytes = tma_epi ? epi_boxes * EPI_BOX_BYTES : EPI_STAGE_BYTES;
  // tap-row mode for 3x3 'same' convs (taps ordered (r, s), s-taps = consecutive rows): one 136-row block per tap row
  const int smem_budget = 227 * 1024 - epi_bytes - 1024 /*align slack*/ - 512 /*barriers*/;
  const int b_stage_bytes = b_rows * BK * 2 * split_mul;
  bool group3 = p->ntaps == 9 && a_ld == p->k_per_tap && p->tap_mode != 1 && !grouped;
  for (int r = 0; r < 3 && group3; ++r)
    group3 = p->tap_shift[3 * r + 1] == p->tap_shift[3 * r] + 1 && p->tap_shift[3 * r + 2] == p->tap_shift[3 * r] + 2;
  int a_stages = 2, b_stages = 0;
  bool resident_b = false;
  if (group3) {
    static const int res_env = getenv("GLASS_RESIDENT_B") ? atoi(getenv("GLASS_RESIDENT_B")) : 1;
    const int kblocks_all = p->ntaps * (p->k_per_tap / BK);
    if (res_env && p->n == bn && kblocks_all <= MAX_STAGES &&
        kblocks_all * b_stage_bytes + 2 * A_BLOCK_BYTES * split_mul <= smem_budget) {
      resident_b = true;
      b_stages = kblocks_all;
      a_stages = (smem_budget - b_stages * b_stage_bytes) / (A_BLOCK_BYTES * split_mul);
      if (a_stages > MAX_A_STAGES) a_stages = MAX_A_STAGES;
    }
  }
  if (group3 && !resident_b) {
    b_stages = (smem_budget - a_stages * A_BLOCK_BYTES * split_mul) / b_stage_bytes;
    if (b_stages > MAX_STAGES) {
      b_stages = MAX_STAGES;
      a_stages = (smem_budget - b_stages * b_stage_bytes) / (A_BLOCK_BYTES * split_mul);
      if (a_stages > MAX_A_STAGES) a_stages = MAX_A_STAGES;
    }
    if (b_stages < 3) group3 = false;  // not enough room to keep the weight stream pipelined
  }
  const int a_box_rows = group3 ? A_BLOCK_ROWS : BM;
  if (make_map_2d(&ma_hi, p->a_hi, a_inner, a_rows, BK, a_box_rows, a_ld)) return -1;
  if (make_map_2d(&mb_hi, p->b_hi, ktot, p->n, BK, b_rows)) return -1;
  if (split) {
    if (make_map_2d(&ma_lo, p->a_lo, a_inner, a_rows, BK, a_box_rows, a_ld)) return -1;
    if (make_map_2d(&mb_lo, p->b_lo, ktot, p->n, BK, b_rows)) return -1;
  } else {
    ma_lo = ma_hi;
    mb_lo = mb_hi;
  }

  GemmKernelParams& k = plan->k;
  k = GemmKernelParams{};
  k.rows_m = rows_m;
  k.tiles_m = (int)((rows_m + BM - 1) / BM);
  k.tiles_n = p->n / bn;
  k.bn = bn;
  k.kblocks_per_tap = p->k_per_tap / BK;
  k.ntaps = p->ntaps;
  k.a_col0 = grouped ? p->a_col0 : 0;
  for (int i = 0; i < GLASS_MAX_TAPS; ++i) k.tap_shift[i] = p->tap_shift[i];
  k.b_tile_bytes = b_rows * BK * 2;
  k.stage_bytes = (A_TILE_BYTES + k.b_tile_bytes) * split_mul;
  k.num_stages = smem_budget / k.stage_bytes;
  if (k.num_stages > MAX_STAGES) k.num_stages = MAX_STAGES;
  GLASS_CHECK(k.num_stages >= 2, "stage too large");
  k.group3 = group3 ? 1 : 0;
  k.a_stages = a_stages;
  k.a_stage_bytes = A_BLOCK_BYTES * split_mul;
  k.b_stages = b_stages;
  k.resident_b = (group3 && resident_b) ? 1 : 0;
  k.b_stage_bytes = b_stage_bytes;
  k.ring_bytes = group3 ? a_stages * k.a_stage_bytes + b_stages * b_stage_bytes : k.num_stages * k.stage_bytes;
  // default: drain every 2 k-blocks (split) / 4 (single pass): a drain reads the whole 128 x BN fp32 tile from TMEM
  // (64 B/clk/SM), which hides behind two k-blocks of MMA work but not behind one
  k.kb_per_chunk = p->kb_per_chunk > 0 ? p->kb_per_chunk : (split ? 2 : 4);
  k.acc_cols = bn <= 64 ? 64 : (bn <= 128 ? 128 : 256);
  k.acc_bufs = TMEM_COLS / k.acc_cols;  // narrow tiles get a deeper ring: the chunk hand-shake latency hides behind it
  if (const char* e = getenv("GLASS_ACC_BUFS")) {  // tuning / A-B knob (rounded down to a power of two)
    const int v = atoi(e);
    if (v >= 2 && v <= k.acc_bufs) k.acc_bufs = v >= 8 ? 8 : (v >= 4 ? 4 : 2);
  }
  k.acc_shift = k.acc_bufs == 8 ? 3 : (k.acc_bufs == 4 ? 2 : 1);
  k.m_h = p->m_h; k.m_w = p->m_w; k.m_border = m_b; k.m_border_hi = m_bh;
  k.scale = p->scale; k.bias = p->bias;
  k.relu_pre = p->relu_pre; k.relu_post = p->relu_post;
  k.res_hi = (const __half*)p->res_hi; k.res_lo = (const __half*)p->res_lo;
  k.res_hp = p->res_hp; k.res_wp = p->res_wp; k.res_border = GLASS_BORDER_LO(p->res_border); k.res_shift = p->res_shift;
  k.out_hi = (__half*)p->out_hi; k.out_lo = (__half*)p->out_lo; k.out_f32 = p->out_f32;
  k.out_hp = p->out_hp; k.out_wp = p->out_wp; k.out_border = GLASS_BORDER_LO(p->out_border);
  k.ld_out = p->ld_out; k.ld_f32 = p->ld_f32; k.n_store = n_store;
  k.m_count_dev = p->m_count_dev;
  k.m_rows_per_count = p->m_rows_per_count;
  k.sat_count = p->sat_count;

  plan->tma_epi = tma_epi ? 1 : 0;
  k.tma_res = (tma_epi && p->res_hi) ? 1 : 0;
  k.epi_boxes = epi_boxes;
  k.epi_bytes = epi_bytes;
  if (tma_epi) {
    if (make_map_2d(&plan->mo_hi, p->out_hi, n_store, rows_m, BK, BM, p->ld_out)) return -1;
    if (make_map_2d(&plan->mo_lo, p->out_lo, n_store, rows_m, BK, BM, p->ld_out)) return -1;
    if (p->res_hi) {
      if (make_map_2d(&plan->mr_hi, p->res_hi, p->n, rows_m, BK, BM, p->ld_out)) return -1;
      if (make_map_2d(&plan->mr_lo, p->res_lo, p->n, rows_m, BK, BM, p->ld_out)) return -1;
    } else {
      plan->mr_hi = plan->mo_hi;
      plan->mr_lo = plan->mo_lo;
    }
  } else {
    plan->mo_hi = plan->mo_lo = plan->mr_hi = plan->mr_lo = ma_hi;   // unused by the classic epilogue
  }
  if (p->m_count_dev) GLASS_CHECK(p->m_rows_per_count > 0, "m_count_dev needs m_rows_per_count > 0");

  // always ask for > half of the SM's shared memory: exactly one CTA per SM owns all 512 TMEM columns
  const int smem_bytes = k.ring_bytes + epi_bytes + 1024 /*align slack*/ + 512 /*barriers*/;
  const int smem_req = smem_bytes < 120 * 1024 ? 120 * 1024 : smem_bytes;
  plan->split = split ? 1 : 0;
  plan->pair = pair ? 1 : 0;
  plan->smem_req = smem_req;
  if (pair) {
    const int units = ((k.tiles_m + 1) / 2) * k.tiles_n;
    const int pairs = units < sms / 2 ? units : sms / 2;
    plan->grid = 2 * pairs;
  } else {
    const int total_tiles = k.tiles_m * k.tiles_n;
    plan->grid = total_tiles < sms ? total_tiles : sms;
  }
  return 0;
}

// the largest dynamic shared-memory request any plan makes: set once per kernel variant, not per launch
static constexpr int kMaxSmem = 227 * 1024;
template <bool SPLIT, bool PAIR, bool TMAEPI>   // (the variants share one function-pointer TYPE: key the static on the variant)
static cudaError_t ensure_smem_attr() {
  static cudaError_t once = cudaFuncSetAttribute(conv_gemm_kernel<SPLIT, PAIR, TMAEPI>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  return once;
}

template <bool SPLIT, bool PAIR, bool TMAEPI>
static cudaError_t launch_variant(const cudaLaunchConfig_t* cfg, const GlassGemmPlan* plan) {
  cudaError_t e = ensure_smem_attr<SPLIT, PAIR, TMAEPI>();
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(cfg, conv_gemm_kernel<SPLIT, PAIR, TMAEPI>, plan->ma_hi, plan->ma_lo, plan->mb_hi, plan->mb_lo,
                            plan->mo_hi, plan->mo_lo, plan->mr_hi, plan->mr_lo, plan->k);
}

static int launch_plan(const GlassGemmPlan* plan, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  int nattr = 0;
  attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[nattr].val.programmaticStreamSerializationAllowed = 1;
  ++nattr;
  if (plan->pair) {
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = 2;
    attr[nattr].val.clusterDim.y = 1;
    attr[nattr].val.clusterDim.z = 1;
    ++nattr;
  }
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = plan->smem_req;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  const int v = (plan->split ? 4 : 0) | (plan->pair ? 2 : 0) | (plan->tma_epi ? 1 : 0);
  switch (v) {
    case 7: GLASS_CUDA((launch_variant<true, true, true>(&cfg, plan))); break;
    case 6: GLASS_CUDA((launch_variant<true, true, false>(&cfg, plan))); break;
    case 5: GLASS_CUDA((launch_variant<true, false, true>(&cfg, plan))); break;
    case 4: GLASS_CUDA((launch_variant<true, false, false>(&cfg, plan))); break;
    case 2: GLASS_CUDA((launch_variant<false, true, false>(&cfg, plan))); break;
    default: GLASS_CUDA((launch_variant<false, false, false>(&cfg, plan))); break;   // (the TMA epilogue needs split output)
  }
  count_launch();
  GLASS_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int glass_conv_gemm(const GlassConvGemmParams* p, void* stream_v) {
  GlassGemmPlan plan;
  if (build_plan(p, &plan)) return -1;
  return launch_plan(&plan, reinterpret_cast<cudaStream_t>(stream_v));
}

extern "C" int glass_plan_create(const GlassConvGemmParams* p, void** plan_out) {
  GLASS_CHECK(plan_out != nullptr, "null plan_out");
  *plan_out = nullptr;
  GlassGemmPlan* plan = new (std::nothrow) GlassGemmPlan;
  GLASS_CHECK(plan != nullptr, "out of host memory");
  if (build_plan(p, plan)) {
    delete plan;
    return -1;
  }
  *plan_out = plan;
  return 0;
}

extern "C" int glass_plan_launch(const void* plan, void* stream_v) {
  GLASS_CHECK(plan != nullptr, "null plan");
  return launch_plan(static_cast<const GlassGemmPlan*>(plan), reinterpret_cast<cudaStream_t>(stream_v));
}

extern "C" int glass_plan_destroy(void* plan) {
  delete static_cast<GlassGemmPlan*>(plan);
  return 0;
}
