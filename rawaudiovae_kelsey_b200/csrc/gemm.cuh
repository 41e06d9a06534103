// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]        (bf16 operands, fp32 accumulation in TMEM)
//
// One CTA per SM loops over 128 x BLOCK_N output tiles (x split-K slices for the weight-gradient GEMMs):
//   warp 0 (one lane)  : TMA producer   - fills a STAGES-deep ring of {A tile, B tile} in 128B-swizzled smem
//   warp 1 (one lane)  : UMMA issuer    - tcgen05.mma 128 x BLOCK_N x 16 into one of two TMEM accumulators
//   warps 2..5         : epilogue       - tcgen05.ld the finished accumulator, apply the fused epilogue
//                                         (bias/ReLU/tanh, reparameterisation+KL, tanh+MSE+dL/da, ReLU mask, ...)
// so the epilogue of tile i overlaps the MMAs of tile i+1 (double-buffered TMEM, 2 x BLOCK_N columns).
//
// Operands may be K-major (row = m or n, K contiguous: activations / weights in the forward pass) or
// MN-major (row = k, M or N contiguous: weights in dgrad, activations in wgrad) - no transposed copies are
// ever materialised; only the TMA box and the UMMA descriptor change.
//
// "fp32 mode" runs the same kernel with num_passes = 3 over split operands (x = hi + lo, both bf16):
// hi*hi + hi*lo + lo*hi accumulated in the same fp32 TMEM accumulator (error ~2^-16 relative).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace rvae {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two teams of four)
constexpr int kEpiTeams = 2;
constexpr int kMaxPasses = 3;
constexpr int kOutSlotBytes = 128 * 128;  // one epilogue staging slot: 128 rows x 128 bytes (64 bf16 / 32 fp32 columns)
constexpr int kOutSlots = 2;

enum : int { MAJOR_K = 0, MAJOR_MN = 1 };
enum : int { EPI_LINEAR = 0, EPI_HEAD = 1, EPI_OUT = 2, EPI_DRELU = 3, EPI_DZ = 4, EPI_WGRAD = 5 };
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_TANH_APPROX = 3 };

// Epilogue arguments. Slot meaning per epilogue kind:
//   LINEAR : v = acc + bias[n]; act(v) -> out_hi (bf16) [, out_lo (bf16 residual)] [, out_f32]
//   HEAD   : tile columns [0,half) are mu, [half,BLOCK_N) are logvar of the same latent columns.
//            mu -> out_f32, logvar -> out_f32_b, eps <- in0 (f32, NULL = 0), z = mu + eps*exp(lv/2) -> out_hi/out_lo,
//            aux0 <- eps*sigma/2 (f32, for backward), aux1 <- c0*mu, aux2 <- c0*(e^lv-1)/2 (KL gradients),
//            loss_acc[0] += sum(1 + lv - mu^2 - e^lv)
//   OUT    : xh = tanh(acc + bias[n]) -> out_f32 (optional); x <- in0 (bf16 hi) [+ in1 (bf16 lo)];
//            loss_acc[0] += sum((xh-x)^2); da = c0*(xh-x)*(1-xh^2) -> out_hi/out_lo
//   DRELU  : v = acc * [in0[m,n] > 0] (in0 bf16, optional) -> out_hi/out_lo
//   DZ     : dz = acc; dmu = dz + in1[m,n]; dlv = dz*in0[m,n] + in2[m,n] (all f32 [M, L]);
//            dmu -> out_hi[m, n], dlv -> out_hi[m, L + n] (ldo = 2L) [, out_lo likewise]
//   WGRAD  : out_f32[m, n] (+)= acc   (red.add when accumulate != 0, plain store otherwise)
//   OUT / DRELU / DZ additionally accumulate the column sums of what they emit into `colsum` (the bias gradient of
//   the layer whose pre-activation gradient this is: db = sum_b da), so no separate reduction kernel is needed.
struct EpiArgs {
  const float* bias;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;
  float* out_f32_b;
  const void* in0;
  const void* in1;
  const void* in2;
  float* aux0;
  float* aux1;
  float* aux2;
  double* loss_acc;
  float* colsum;  // OUT / DRELU / DZ: colsum[c] += sum over rows of the bf16 stream's fp32 values (bias gradients)
  int ldo;   // leading dimension (elements) of out_hi/out_lo/out_f32 and of bf16 inputs in0/in1
  int act;   // LINEAR / OUT activation
  int L;     // HEAD / DZ latent width
  int accumulate;
  float c0;
};

struct alignas(64) GemmParams {
  CUtensorMap tmA[kMaxPasses];
  CUtensorMap tmB[kMaxPasses];
  CUtensorMap tmOutHi;   // bf16 output plane (box 64 cols x 128 rows, 128B swizzle)
  CUtensorMap tmOutLo;   // bf16 residual plane (fp32 emulation)
  CUtensorMap tmOutF32;  // fp32 output (box 32 cols x 128 rows, 128B swizzle)
  int M, N, K;
  int num_passes;
  int m_blocks, n_blocks;  // m_blocks counts 128*CG-row tiles
  int k_splits, kb_per_split, kb_total;
  int b_tile_stride;  // K-major B: row advance per n-block
  int b_half_stride;  // K-major B: row offset of the second half-tile load
  int debug;          // experiments only (env RVAE_DEBUG, CG == 1): 1 = no MMA issue, 2 = no TMA loads
  EpiArgs epi;
};

// CG = CTAs cooperating on one tile (tcgen05 cta_group): 1 -> 128 x BLOCK_N tile per CTA; 2 -> a CTA pair computes a
// 256 x BLOCK_N tile, each CTA staging its own 128 rows of A and HALF of the B tile (the pair's tensor cores read
// both halves), which cuts L2->smem traffic per FLOP by a third and doubles the MMA's M.
template <int BLOCK_N, int CG>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = (BLOCK_N / CG) * kBlockK * 2;  // per CTA
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = kOutSlots * kOutSlotBytes;  // epilogue staging for TMA stores
  static constexpr int kBarrierBytes = 256 + kEpiTeams * 128 * 4;  // mbarriers + TMEM slot + per-team strips
  static constexpr int kBudget = 232448 - 1024 - kOutBytes - kBarrierBytes;  // 227 KB usable per CTA
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = (kAccStages * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kOutBytes + kBarrierBytes;
};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Store COUNT consecutive fp32 values of one output row as bf16 (hi plane and optional residual plane).
template <int COUNT>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* hi_row, __nv_bfloat16* lo_row, const float (&v)[COUNT]) {
  uint4* dh = reinterpret_cast<uint4*>(hi_row);
#pragma unroll
  for (int i = 0; i < COUNT / 8; ++i) {
    dh[i] = make_uint4(ptx::pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                       ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  }
  if (lo_row != nullptr) {
    uint4* dl = reinterpret_cast<uint4*>(lo_row);
#pragma unroll
    for (int i = 0; i < COUNT / 8; ++i) {
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = v[8 * i + j] - __bfloat162float(__float2bfloat16_rn(v[8 * i + j]));
      dl[i] = make_uint4(ptx::pack_bf16x2(r[0], r[1]), ptx::pack_bf16x2(r[2], r[3]), ptx::pack_bf16x2(r[4], r[5]),
                         ptx::pack_bf16x2(r[6], r[7]));
    }
  }
}

template <int COUNT>
__device__ __forceinline__ void store_row_f32(float* row, const float (&v)[COUNT]) {
  float4* d = reinterpret_cast<float4*>(row);
#pragma unroll
  for (int i = 0; i < COUNT / 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

template <int COUNT>
__device__ __forceinline__ void load_row_f32(const float* row, float (&v)[COUNT]) {
  const float4* s = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int i = 0; i < COUNT / 4; ++i) {
    float4 t = __ldg(s + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

template <int COUNT>
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* row, float (&v)[COUNT]) {
  const uint4* s = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int i = 0; i < COUNT / 8; ++i) {
    uint4 t = __ldg(s + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __bfloat1622float2(h[j]);
      v[8 * i + 2 * j] = f.x;
      v[8 * i + 2 * j + 1] = f.y;
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Epilogue staging. Eight epilogue warps form two TEAMS of four (one warp per TMEM lane quarter, one output row per
// thread). The work units of a tile (64-column bf16 sub-tiles, 32-column fp32 sub-tiles) alternate between the
// teams, so one team's TMEM loads / global side loads / barrier waits overlap the other's math and smem writes.
// A team stages a unit as 128 rows x 128 bytes in its own 128B-swizzled smem slot (16-byte chunk c of row r lands at
// chunk c ^ (r & 7): conflict-free for row-per-thread writes); then one thread issues a TMA store (or reduce-add)
// of the slot: full 128-byte coalesced lines to L2 instead of 16-byte row-strided stores, and rows / columns beyond
// the tensor are clipped by the TMA unit.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void team_bar_sync(int team) {
  asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory");
}

__device__ __forceinline__ void slot_write16(uint8_t* slot, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(slot + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}
// 32 fp32 values -> bf16 into chunks [c0, c0 + 4) of row r
__device__ __forceinline__ void stage_bf16(uint8_t* slot, int r, int c0, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    slot_write16(slot, r, c0 + i,
                 make_uint4(ptx::pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), ptx::pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                            ptx::pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), ptx::pack_bf16x2(v[8 * i + 6], v[8 * i + 7])));
}
__device__ __forceinline__ void stage_bf16_residual(uint8_t* slot, int r, int c0, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = v[8 * i + j] - __bfloat162float(__float2bfloat16_rn(v[8 * i + j]));
    slot_write16(slot, r, c0 + i,
                 make_uint4(ptx::pack_bf16x2(q[0], q[1]), ptx::pack_bf16x2(q[2], q[3]), ptx::pack_bf16x2(q[4], q[5]),
                            ptx::pack_bf16x2(q[6], q[7])));
  }
}
// 32 fp32 values -> the 8 chunks of row r
__device__ __forceinline__ void stage_f32(uint8_t* slot, int r, const float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    slot_write16(slot, r, i,
                 make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                            __float_as_uint(v[4 * i + 3])));
}

// Column sums of a [32 rows (lanes) x 64 columns] register tile: butterfly reduce-scatter over the lanes (62
// shuffles); afterwards lane l holds the sums of columns 2l and 2l+1 in v[0], v[1]. Destroys v.
__device__ __forceinline__ void warp_colsum64(float* v, int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int half = 32 >> step;          // values kept per lane after this step
    const int mask = 16 >> step;
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < half) {
        // keep columns [0, half) if the mask bit of this lane is clear, [half, 2*half) otherwise
        const float send = upper ? v[j] : v[j + half];
        const float keep = upper ? v[j + half] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    }
  }
}

// One team's staging slot + bias strip. acquire(): the team's previous TMA store has finished reading the slot
// (and, as a side effect of the barrier, the bias strip written before it is visible). commit(): hand the slot to TMA.
struct TeamOut {
  uint8_t* slot;
  float* bias_s;   // 64 floats
  float* csum_s;   // 64 floats: per-unit column sums, combined across the team's four warps
  int team;
  bool issuer;
  int debug;       // experiments: 4 = no TMA store, 16 = no team barriers
  __device__ __forceinline__ void acquire() {
    if (issuer) ptx::tma_store_wait_read<0>();
    if (!(debug & 16)) team_bar_sync(team);
  }
  __device__ __forceinline__ void commit(const CUtensorMap* tm, int c0, int c1, bool reduce) {
    ptx::fence_proxy_async_smem();
    if (!(debug & 16)) team_bar_sync(team);
    if (issuer && !(debug & 4)) {
      if (reduce) ptx::tma_reduce_add_2d(tm, slot, c0, c1);
      else ptx::tma_store_2d(tm, slot, c0, c1);
      ptx::tma_store_commit();
    }
  }
};

template <int BLOCK_N, int A_MAJOR, int B_MAJOR, int EPI, int CG>
__device__ __forceinline__ void gemm_body(const GemmParams& p) {
  using Cfg = GemmCfg<BLOCK_N, CG>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kIdesc = ptx::umma_idesc_bf16(kBlockM * CG, BLOCK_N, A_MAJOR, B_MAJOR);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* out_slots = smem + kStages * Cfg::kStageBytes;  // 1024-byte aligned (stage sizes are multiples of 1 KB)
  float* bias_strips = reinterpret_cast<float*>(out_slots + Cfg::kOutBytes);  // kEpiTeams x (64 bias + 64 colsum)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_slots + Cfg::kOutBytes + kEpiTeams * 128 * 4);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + Cfg::kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int group_id = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;       // tile-owning unit: CTA or CTA pair
  const int num_groups = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.num_passes; ++i) {
      ptx::prefetch_tensormap(&p.tmA[i]);
      ptx::prefetch_tensormap(&p.tmB[i]);
    }
    ptx::prefetch_tensormap(&p.tmOutHi);
    ptx::prefetch_tensormap(&p.tmOutLo);
    ptx::prefetch_tensormap(&p.tmOutF32);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);    // leader's barrier: its producer arrives once with the pair's byte count
      ptx::mbar_init(&empty_bar[i], 1);   // per CTA: tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    for (int i = 0; i < Cfg::kAccStages; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);        // per CTA: tcgen05.commit
      ptx::mbar_init(&tmem_empty_bar[i], 8 * CG);  // leader's barrier: one arrive per epilogue warp of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<CG>(tmem_slot, Cfg::kTmemCols);
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of the previous
  // kernel of the stream; from here on we read what it wrote.
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  const int tiles = p.m_blocks * p.n_blocks;
  const int total_units = tiles * p.k_splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one per CTA)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      bool slot_free = ptx::mbar_try_wait(&empty_bar[0], 1);
      for (int u = group_id; u < total_units; u += num_groups) {
        const int ks = u / tiles;
        const int tile = u - ks * tiles;
        const int n_blk = tile / p.m_blocks;
        const int m_blk = tile - n_blk * p.m_blocks;
        const int m0 = (m_blk * CG + static_cast<int>(cta_rank)) * kBlockM;
        const int kb_begin = ks * p.kb_per_split;
        const int kb_count = min(p.kb_per_split, p.kb_total - kb_begin);
        for (int pass = 0; pass < p.num_passes; ++pass) {
          const CUtensorMap* tmA = &p.tmA[pass];
          const CUtensorMap* tmB = &p.tmB[pass];
          for (int kb = kb_begin; kb < kb_begin + kb_count; ++kb) {
            ptx::mbar_wait_probed(slot_free, &empty_bar[stage], phase ^ 1);
            {  // probe the NEXT slot now: the probe's latency overlaps the TMA issue below
              const uint32_t ns = (stage + 1 == kStages) ? 0u : stage + 1;
              const uint32_t np = (stage + 1 == kStages) ? phase ^ 1u : phase;
              slot_free = ptx::mbar_try_wait(&empty_bar[ns], np ^ 1);
            }
            uint64_t* fb = &full_bar[stage];
            if (CG == 1 && (p.debug & 2)) {  // experiment: measure the MMA side alone (operands are stale smem)
              ptx::mbar_arrive(fb);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
              continue;
            }
            if constexpr (CG == 1) {
              ptx::mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
            } else {
              // Both CTAs' loads complete_tx on the LEADER's barrier; the leader alone arrives, announcing the pair's
              // bytes. The peer sends no arrive (a cluster-scope release per k-block would throttle its producer):
              // it cannot run a phase ahead because its smem slot is only freed by the commit that follows the MMAs
              // which consumed this phase, and a complete_tx that lands before the leader's expect_tx merely leaves
              // the tx-count transiently negative while the leader's arrival is still pending.
              if (leader) ptx::mbar_arrive_expect_tx(fb, 2 * Cfg::kStageBytes);
            }
            uint8_t* sA = smem + stage * Cfg::kStageBytes;
            uint8_t* sB = sA + Cfg::kABytes;
            const int k0 = kb * kBlockK;
            if constexpr (A_MAJOR == MAJOR_K) {
              ptx::tma_load_2d_cg<CG>(sA, tmA, fb, k0, m0);
            } else {
#pragma unroll
              for (int i = 0; i < kBlockM / 64; ++i)
                ptx::tma_load_2d_cg<CG>(sA + i * (kBlockK * 128), tmA, fb, m0 + i * 64, k0);
            }
            if constexpr (B_MAJOR == MAJOR_K) {
              // the B tile is staged as two boxes of BLOCK_N/2 rows: CG == 1 loads both, CG == 2 one per CTA
              const int r0 = n_blk * p.b_tile_stride;
              if constexpr (CG == 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  ptx::tma_load_2d_cg<1>(sB + h * (Cfg::kBBytes / 2), tmB, fb, k0, r0 + h * p.b_half_stride);
              } else {
                ptx::tma_load_2d_cg<2>(sB, tmB, fb, k0, r0 + static_cast<int>(cta_rank) * p.b_half_stride);
              }
            } else {
              const int n0 = n_blk * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / CG);
#pragma unroll
              for (int i = 0; i < BLOCK_N / CG / 64; ++i)
                ptx::tma_load_2d_cg<CG>(sB + i * (kBlockK * 128), tmB, fb, n0 + i * 64, k0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer (leader CTA only)
    if (lane == 0 && leader) {
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      bool data_ready = false;
      for (int u = group_id; u < total_units; u += num_groups) {
        const int ks = u / tiles;
        const int kb_begin = ks * p.kb_per_split;
        const int kb_count = min(p.kb_per_split, p.kb_total - kb_begin);
        const int iters = kb_count * p.num_passes;
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int it = 0; it < iters; ++it) {
          ptx::mbar_wait_probed(data_ready, &full_bar[stage], phase);
          {  // probe the NEXT stage now: the MMA issue below hides the probe's latency, so the tensor pipe does not
             // drain while this thread waits on a barrier that has long completed
            const uint32_t ns = (stage + 1 == kStages) ? 0u : stage + 1;
            const uint32_t np = (stage + 1 == kStages) ? phase ^ 1u : phase;
            data_ready = ptx::mbar_try_wait(&full_bar[ns], np);
          }
          ptx::tc_fence_after();
          if (CG == 1 && (p.debug & 1)) {  // experiment: measure the TMA side alone
            ptx::mbar_arrive(&empty_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = (A_MAJOR == MAJOR_K) ? ptx::umma_desc_sw128(a_addr + k * (kUmmaK * 2), 0, 1024)
                                                        : ptx::umma_desc_sw128(a_addr + k * (kUmmaK * 128), kBlockK * 128, 1024);
            const uint64_t bdesc = (B_MAJOR == MAJOR_K) ? ptx::umma_desc_sw128(b_addr + k * (kUmmaK * 2), 0, 1024)
                                                        : ptx::umma_desc_sw128(b_addr + k * (kUmmaK * 128), kBlockK * 128, 1024);
            ptx::umma_bf16<CG>(d_tmem, adesc, bdesc, kIdesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit<CG>(&empty_bar[stage]);  // frees the smem slot (in both CTAs) once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (CG == 1 && (p.debug & 1)) ptx::mbar_arrive(&tmem_full_bar[as]);
        else ptx::umma_commit<CG>(&tmem_full_bar[as]);  // accumulator complete -> epilogue (of both CTAs)
        if (++as == Cfg::kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2 teams x 4 warps)
    const int team = (warp - 2) >> 2;
    const int quarter = warp & 3;  // tcgen05.ld: a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int row = quarter * 32 + lane;  // row of the tile owned by this thread
    const int team_tid = (((warp - 2) & 3) << 5) | lane;  // 0..127 within the team
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const EpiArgs& e = p.epi;
    TeamOut out{out_slots + team * kOutSlotBytes, bias_strips + team * 128, bias_strips + team * 128 + 64, team,
                team_tid == 0, p.debug};
    const bool dual = e.out_lo != nullptr;
    float loss_local = 0.f;
    uint32_t as = 0, aphase = 0;
    for (int u = group_id; u < total_units; u += num_groups) {
      const int ks = u / tiles;
      const int tile = u - ks * tiles;
      const int n_blk = tile / p.m_blocks;
      const int m_blk = tile - n_blk * p.m_blocks;
      const int m0 = (m_blk * CG + static_cast<int>(cta_rank)) * kBlockM;
      const int m = m0 + row;
      const bool row_ok = m < p.M;
      ptx::mbar_wait(&tmem_full_bar[as], aphase);
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem_base + lane_base + as * BLOCK_N;

      if constexpr (EPI == EPI_HEAD) {
        // tile columns [0, kHalf) = mu, [kHalf, BLOCK_N) = logvar of latent columns n_blk*kHalf ..
        constexpr int kHalf = BLOCK_N / 2;
        const int L = e.L;
        for (int sub = team; sub < kHalf / 64; sub += kEpiTeams) {
          const int col0 = n_blk * kHalf + sub * 64;
          if (e.out_hi) out.acquire();
          float z[64];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = sub * 64 + h * 32;
            const int col = col0 + h * 32;
            __syncwarp();
            uint32_t rm[32], rl[32];
            ptx::tmem_ld_32x32(t_acc + c, rm);
            ptx::tmem_ld_32x32(t_acc + kHalf + c, rl);
            const size_t off = static_cast<size_t>(m) * L + col;
            float eps[32];
            if (row_ok && e.in0) {
              load_row_f32<32>(reinterpret_cast<const float*>(e.in0) + off, eps);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) eps[j] = 0.f;
            }
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 2; ++q) {  // 16 columns at a time keeps the side arrays small
              float mu[16], lv[16], esh[16], gmu[16], glv[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int jj = 16 * q + j;
                mu[j] = __uint_as_float(rm[jj]) + __ldg(e.bias + col + jj);
                lv[j] = __uint_as_float(rl[jj]) + __ldg(e.bias + L + col + jj);
                const float sig = expf(0.5f * lv[j]);
                const float var = sig * sig;
                z[h * 32 + jj] = fmaf(eps[jj], sig, mu[j]);
                esh[j] = 0.5f * eps[jj] * sig;
                gmu[j] = e.c0 * mu[j];
                glv[j] = 0.5f * e.c0 * (var - 1.f);
                if (row_ok) loss_local += (1.f + lv[j]) - fmaf(mu[j], mu[j], var);
              }
              if (row_ok) {
                store_row_f32<16>(e.out_f32 + off + 16 * q, mu);
                store_row_f32<16>(e.out_f32_b + off + 16 * q, lv);
                if (e.aux0) store_row_f32<16>(e.aux0 + off + 16 * q, esh);
                if (e.aux1) store_row_f32<16>(e.aux1 + off + 16 * q, gmu);
                if (e.aux2) store_row_f32<16>(e.aux2 + off + 16 * q, glv);
              }
            }
          }
          if (e.out_hi) {
            stage_bf16(out.slot, row, 0, z);
            stage_bf16(out.slot, row, 4, z + 32);
            out.commit(&p.tmOutHi, col0, m0, false);
            if (dual) {
              out.acquire();
              stage_bf16_residual(out.slot, row, 0, z);
              stage_bf16_residual(out.slot, row, 4, z + 32);
              out.commit(&p.tmOutLo, col0, m0, false);
            }
          }
        }
      } else if constexpr (EPI == EPI_WGRAD) {
        for (int piece = team; piece < BLOCK_N / 32; piece += kEpiTeams) {
          const int n = n_blk * BLOCK_N + piece * 32;
          if (n >= p.N) continue;
          __syncwarp();
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_acc + piece * 32, r);
          out.acquire();
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          stage_f32(out.slot, row, v);
          out.commit(&p.tmOutF32, n, m0, e.accumulate != 0);
        }
      } else {
        // ---- bf16 stream(s): LINEAR (act), DRELU (mask), OUT (da4), DZ (dmu then dlv): 64-column units
        constexpr int kStreams = (EPI == EPI_DZ) ? 2 : 1;
        constexpr int kSubs = BLOCK_N / 64;
        if (e.out_hi) {
          for (int unit = team; unit < kStreams * kSubs; unit += kEpiTeams) {
            const int stream = unit / kSubs;
            const int sub = unit - stream * kSubs;
            const int n0 = n_blk * BLOCK_N + sub * 64;
            if (n0 >= p.N) continue;
            __syncwarp();
            // 1) accumulator: both 32-column TMEM loads in flight
            uint32_t r[64];
            if (!(p.debug & 8)) {
              ptx::tmem_ld_32x32(t_acc + sub * 64, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
              ptx::tmem_ld_32x32(t_acc + sub * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j) r[j] = 0x3f800000u + j;
            }
            // 2) side inputs from global memory, issued before anything waits
            const size_t off = static_cast<size_t>(m) * e.ldo + n0;
            float side[64];
            if constexpr (EPI == EPI_DRELU) {
              if (e.in0 && row_ok) {
                load_row_bf16<64>(reinterpret_cast<const __nv_bfloat16*>(e.in0) + off, side);
              } else {
#pragma unroll
                for (int j = 0; j < 64; ++j) side[j] = 1.f;
              }
            } else if constexpr (EPI == EPI_OUT) {
              if (row_ok) {
                load_row_bf16<64>(reinterpret_cast<const __nv_bfloat16*>(e.in0) + off, side);
                if (e.in1) {
                  float xl[64];
                  load_row_bf16<64>(reinterpret_cast<const __nv_bfloat16*>(e.in1) + off, xl);
#pragma unroll
                  for (int j = 0; j < 64; ++j) side[j] += xl[j];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 64; ++j) side[j] = 0.f;
              }
            } else if constexpr (EPI == EPI_DZ) {
              const size_t loff = static_cast<size_t>(m) * e.L + n0;
              if (row_ok) {
                load_row_f32<64>(reinterpret_cast<const float*>(stream == 0 ? e.in1 : e.in2) + loff, side);
              } else {
#pragma unroll
                for (int j = 0; j < 64; ++j) side[j] = 0.f;
              }
            }
            // 3) bias strip of this unit -> smem (published by the barrier inside acquire())
            if constexpr (EPI == EPI_LINEAR || EPI == EPI_OUT) {
              if (team_tid < 64) out.bias_s[team_tid] = e.bias ? __ldg(e.bias + n0 + team_tid) : 0.f;
            }
            const bool do_colsum = (EPI == EPI_OUT || EPI == EPI_DRELU || EPI == EPI_DZ) && e.colsum != nullptr;
            if (do_colsum && team_tid >= 64) out.csum_s[team_tid - 64] = 0.f;
            out.acquire();
            ptx::tmem_ld_wait();
            float v[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
            if constexpr (EPI == EPI_LINEAR) {
#pragma unroll
              for (int j = 0; j < 64; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(out.bias_s + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
              if (e.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
              } else if (e.act == ACT_TANH) {
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] = tanhf(v[j]);
              } else if (e.act == ACT_TANH_APPROX) {
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] = ptx::tanh_approx(v[j]);
              }
            } else if constexpr (EPI == EPI_DRELU) {
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] = (side[j] > 0.f) ? v[j] : 0.f;
            } else if constexpr (EPI == EPI_OUT) {
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                const float a = v[j] + out.bias_s[j];
                const float xh = (e.act == ACT_TANH_APPROX) ? ptx::tanh_approx(a) : tanhf(a);
                const float d = xh - side[j];
                if (row_ok) loss_local = fmaf(d, d, loss_local);
                v[j] = row_ok ? e.c0 * d * (1.f - xh * xh) : 0.f;
              }
            } else if constexpr (EPI == EPI_DZ) {
              if (stream == 0) {  // dmu = dz + g_mu
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] += side[j];
              } else {            // dlv = dz * (eps sigma / 2) + g_logvar
                float esh[64];
                const size_t loff = static_cast<size_t>(m) * e.L + n0;
                if (row_ok) {
                  load_row_f32<64>(reinterpret_cast<const float*>(e.in0) + loff, esh);
                } else {
#pragma unroll
                  for (int j = 0; j < 64; ++j) esh[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] = fmaf(v[j], esh[j], side[j]);
              }
            }
            const int c_out = (EPI == EPI_DZ) ? stream * e.L + n0 : n0;
            if (!(p.debug & 32)) {
              stage_bf16(out.slot, row, 0, v);
              stage_bf16(out.slot, row, 4, v + 32);
            } else if (v[0] == 12345.678f) {
              out.slot[row] = 1;  // keep v live
            }
            if (dual) {  // residual plane (fp32 emulation): second pass through the same slot, before v is reduced
              out.commit(&p.tmOutHi, c_out, m0, false);
              out.acquire();
              stage_bf16_residual(out.slot, row, 0, v);
              stage_bf16_residual(out.slot, row, 4, v + 32);
            }
            if (do_colsum) {
              // bias gradient: column sums of this unit over the tile's 128 rows (rows >= M contribute zeros)
              warp_colsum64(v, lane);
              atomicAdd(out.csum_s + 2 * lane, v[0]);
              atomicAdd(out.csum_s + 2 * lane + 1, v[1]);
            }
            out.commit(dual ? &p.tmOutLo : &p.tmOutHi, c_out, m0, false);
            // (the barrier inside commit() ordered the four warps' smem atomics before this read)
            if (do_colsum && team_tid < 64) atomicAdd(e.colsum + c_out + team_tid, out.csum_s[team_tid]);
          }
        }
        // ---- fp32 stream: LINEAR's fp32 copy / OUT's xhat: 32-column units
        if constexpr (EPI == EPI_LINEAR || EPI == EPI_OUT) {
          if (e.out_f32) {
            for (int piece = team; piece < BLOCK_N / 32; piece += kEpiTeams) {
              const int n = n_blk * BLOCK_N + piece * 32;
              if (n >= p.N) continue;
              __syncwarp();
              uint32_t r[32];
              ptx::tmem_ld_32x32(t_acc + piece * 32, r);
              if (team_tid < 32) out.bias_s[team_tid] = e.bias ? __ldg(e.bias + n + team_tid) : 0.f;
              out.acquire();
              ptx::tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + out.bias_s[j];
              const int act = (EPI == EPI_OUT && e.act != ACT_TANH_APPROX) ? ACT_TANH : e.act;
              if (act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              } else if (act == ACT_TANH) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
              } else if (act == ACT_TANH_APPROX) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ptx::tanh_approx(v[j]);
              }
              if constexpr (EPI == EPI_OUT) {
                if (!e.out_hi && row_ok) {  // loss-only forward: accumulate the MSE here
                  const size_t off = static_cast<size_t>(m) * e.ldo + n;
                  float x[32];
                  load_row_bf16<32>(reinterpret_cast<const __nv_bfloat16*>(e.in0) + off, x);
                  if (e.in1) {
                    float xl[32];
                    load_row_bf16<32>(reinterpret_cast<const __nv_bfloat16*>(e.in1) + off, xl);
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] += xl[j];
                  }
#pragma unroll
                  for (int j = 0; j < 32; ++j) loss_local = fmaf(v[j] - x[j], v[j] - x[j], loss_local);
                }
              }
              stage_f32(out.slot, row, v);
              out.commit(&p.tmOutF32, n, m0, false);
            }
          }
        }
      }
      // release the accumulator back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || leader) ptx::mbar_arrive(&tmem_empty_bar[as]);
        else ptx::mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      }
      if (++as == Cfg::kAccStages) { as = 0; aphase ^= 1; }
    }
    if (out.issuer) ptx::tma_store_wait<0>();  // all bulk stores issued by this thread have completed
    if constexpr (EPI == EPI_HEAD || EPI == EPI_OUT) {
      const float s = warp_sum(loss_local);
      if (lane == 0 && e.loss_acc) atomicAdd(e.loss_acc, static_cast<double>(s));
    }
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int A_MAJOR, int B_MAJOR, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  gemm_body<BLOCK_N, A_MAJOR, B_MAJOR, EPI, 1>(p);
}

template <int BLOCK_N, int A_MAJOR, int B_MAJOR, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_kernel_2cta(const __grid_constant__ GemmParams p) {
  gemm_body<BLOCK_N, A_MAJOR, B_MAJOR, EPI, 2>(p);
}

}  // namespace rvae
