// conv_pair.cuh -- the wide, long-K 3x3 convs (bottleneck ResnetBlocks, unet.py:17-39) as a CTA-PAIR implicit GEMM:
// tcgen05.mma.cta_group::2, M = 256 pixels per pair (128 per CTA), N = block_n, K blocks of 64 channels of one tap.
//
// Why a pair: with cta_group::1 every UMMA (M128 x N192 x K16) has the tensor core read a 4 KB A slice and a 6 KB B slice from the
// CTA's shared memory while TMA writes the same amount: ~213 B/clk against the 128 B/clk port, i.e. a structural ceiling of 60-65 %
// tensor-pipe activity (profiles/r1_dominant_kernel_ncu.md measured 65.8 %).  In a pair each CTA stages its own 128 rows of A but only
// HALF of the weight rows (block_n / 2); the instruction reads B from both CTAs' shared memory: 4 + 3 KB per UMMA and CTA.
//
// Protocol (one cluster = CTA 0 "leader" + CTA 1 "peer", same shared-memory layout in both):
//   warp 0 (both CTAs)  TMA producer: waits on its OWN empty[s], arms the LEADER's full[s] with its byte count and issues its loads with
//                       .cta_group::2 so that their complete_tx lands on the leader's barrier
//   warp 1 (leader)     one thread waits on full[s] (2 arrivals + both CTAs' bytes), issues 4 x tcgen05.mma.cta_group::2, commits to
//                       empty[s] of BOTH CTAs (multicast); after the last K block commits to tfull[a] of both CTAs
//   warp 2 (both)       TMEM allocation (cta_group::2: one warp of each CTA), 2 accumulator stages of block_n columns
//   warps 4..11 (both)  epilogue of the CTA's own 128 rows: tcgen05.ld -> + bias -> ReLU -> + fp16 residual -> fp16 stores; one arrival
//                       per warp on the LEADER's tempty[a] (remote for the peer)
// Every wait is bounded (mbar_wait traps), so a protocol error is a CUDA error, not a hung GPU.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace vsb {

struct ConvPairParams {
  int M, N, num_kb, c_blocks;     // pixels, output channels, K blocks (taps x channel blocks of 64), channel blocks per tap
  int S, pad;                     // kernel width, zero padding (taps advance s fastest, like the weight K order (r, s, c))
  int tile_h, tiles_per_img;      // an M tile = tile_h full rows of the map (tile_h * W == 128); tiles per image = H / tile_h
  int m_pairs, n_tiles, num_work; // M tiles / 2, N / block_n, m_pairs * n_tiles
  int block_n, stages;
  uint32_t a_bytes, b_bytes;      // per stage and CTA: 128 x 64 and (block_n / 2) x 64 fp16
  uint32_t idesc;                 // kind::f16, D = f32, A = B = f16 K-major, N = block_n, M = 256
  int relu;
  const float* bias;
  const __half* resid; int ld_res;   // added AFTER the activation (ResnetBlock: act(norm(conv)) + res), may be null
  __half* out; int ld_out;
};

constexpr int kPairThreads = 384;   // 12 warps
constexpr int kPairEpiWarps = 8;
constexpr int kPairMaxStages = 8;
constexpr uint32_t kPairTmemCols = 512;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
// Default semantics (.release.cta) on purpose: with .release.cluster the compiler puts MEMBAR.ALL.GPU + ERRBAR in front of every arrive
// (ncu of the first version: the producers spent their time in those fences and the pair ran at half the single-CTA rate).  The data
// these arrivals order travels through the async proxy (TMA complete_tx) or is fenced by tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// TMA loads of a CTA pair: the data goes to the executing CTA's shared memory, the complete_tx to `bar_cluster_addr` (the leader's barrier)
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16: 128 rows from each CTA] * B[block_n x 16: half the rows from each CTA]; leader thread only
__device__ __forceinline__ void umma2_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once all previously issued MMAs of the pair have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvPairParams p) {
  extern __shared__ uint8_t pair_smem_raw[];
  uint8_t* tiles = pair_smem_raw + ((1024u - (smem_u32(pair_smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B tiles: 1024-byte aligned
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);    // leader's copies are the live ones
  uint64_t* empty_bar = full_bar + kPairMaxStages;
  uint64_t* tfull_bar = empty_bar + kPairMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;                                                         // leader's copies are the live ones
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)cluster_id_x(), ncl = (int)cluster_nid_x();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 2);       // one arrive.expect_tx per CTA
      mbar_init(&empty_bar[s], 1);      // the leader's commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * kPairEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem2_alloc(tmem_slot, kPairTmemCols);
    tmem2_relinquish();
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                   // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================================================================== TMA producer (one lane, both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cid; w < p.num_work; w += ncl) {
        const int n_tile = w % p.n_tiles, m_tile = 2 * (w / p.n_tiles) + (int)rank;
        const int img = m_tile / p.tiles_per_img, cy = (m_tile - img * p.tiles_per_img) * p.tile_h;
        const int brow = n_tile * p.block_n + (int)rank * (p.block_n >> 1);
        int tap_r = 0, tap_s = 0, cblk = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = tiles + (size_t)stage * stage_bytes;
          const uint32_t fb = mapa_rank(smem_u32(&full_bar[stage]), 0);
          mbar_arrive_expect_tx_cluster(fb, stage_bytes);
          tma2_load_4d(&tmA, fb, sa, cblk * 64, tap_s - p.pad, cy + tap_r - p.pad, img);
          tma2_load_2d(&tmB, fb, sa + p.a_bytes, kb * 64, brow);
          if (++cblk == p.c_blocks) { cblk = 0; if (++tap_s == p.S) { tap_s = 0; ++tap_r; } }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread of the leader)
    if (rank == 0 && elect_one()) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int w = cid; w < p.num_work; w += ncl) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.block_n);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa, 128), bdesc = make_smem_desc(sa + p.a_bytes, 128);
#pragma unroll
          for (int j = 0; j < 4; ++j)       // 16 K elements = 32 bytes inside the swizzled row: +2 in the (addr >> 4) field
            umma2_f16_ss(d_tmem, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), p.idesc, (uint32_t)((kb | j) != 0));
          umma2_commit_both(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma2_commit_both(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (8 warps per CTA: lane quadrant x column half)
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int nch = p.block_n >> 4, c_beg = half * (nch >> 1), c_end = c_beg + (nch >> 1);
    int as = 0;
    uint32_t aphase = 0;
    for (int w = cid; w < p.num_work; w += ncl) {
      const int n_tile = w % p.n_tiles, m_tile = 2 * (w / p.n_tiles) + (int)rank;
      const long m = (long)m_tile * 128 + row;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n);
      for (int ch = c_beg; ch < c_end; ++ch) {
        const int n = n_tile * p.block_n + ch * 16;
        float v[16];
        tmem_ld16(t0 + (uint32_t)(ch * 16), v);
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 bq = __ldg(b4 + u);
          v[4 * u + 0] += bq.x; v[4 * u + 1] += bq.y; v[4 * u + 2] += bq.z; v[4 * u + 3] += bq.w;
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.resid != nullptr) {
          __align__(16) __half h[16];
          const uint4* rg = reinterpret_cast<const uint4*>(p.resid + m * p.ld_res + n);
          reinterpret_cast<uint4*>(h)[0] = __ldg(rg);
          reinterpret_cast<uint4*>(h)[1] = __ldg(rg + 1);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += __half2float(h[j]);
        }
        __align__(16) __half2 h2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          h2[j] = __floats2half2_rn(fmaxf(fminf(v[2 * j], 65504.f), -65504.f), fmaxf(fminf(v[2 * j + 1], 65504.f), -65504.f));
        uint4* o = reinterpret_cast<uint4*>(p.out + m * p.ld_out + n);
        o[0] = reinterpret_cast<const uint4*>(h2)[0];
        o[1] = reinterpret_cast<const uint4*>(h2)[1];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty_bar[as]), 0));     // this warp's part of the accumulator stage is drained
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  // teardown: nobody leaves (or frees TMEM) while the pair's MMAs / remote arrivals can still touch this CTA
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem2_dealloc(tmem_base, kPairTmemCols);
}

}  // namespace vsb
