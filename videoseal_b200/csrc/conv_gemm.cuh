// K1/K2: implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05.mma,
// accumulators in TMEM), persistent and warp-specialised:
//
//   warp 0        : TMA producer (weights unless resident; activations for LD_TMA; input halo tiles for LD_HALO_*)
//   warp 1        : single-thread tcgen05.mma issuer
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue (TMEM -> registers -> fused bias/BN-fold/act/residual/LN/tanh/GRN-stats -> HBM);
//                   warp w owns TMEM lane quadrant w%4 and every second 16-column chunk ((w-4)/4)
//   warps 12..15  : A-tile builders (all loaders except LD_TMA):
//       LD_HALO_CONV3 : on-chip im2col - the (8+2)x(16+2) input halo of an 8x16 output tile is TMA-loaded ONCE per channel
//                       chunk (zero fill = conv padding) and the tap tiles are copied smem->smem into the swizzled
//                       UMMA layout (input pixels cross L2->SM once instead of nine times)
//       LD_HALO_UPS   : the UBlock's bilinear x2 (align_corners=False) + ReflectionPad2d(1) + virtual skip-concat: a 6x10
//                       low-resolution halo is TMA-loaded, the 10x18 upsampled+padded halo is built from it in shared
//                       memory once, then tap tiles are copied as in LD_HALO_CONV3
//       LD_GATHER_CONV: generic strided / reflect-padded conv gathered from global memory (stride-2 3x3, k2s2 patchify,
//                       reflect-padded head conv)
//       LD_GATHER_SCALE: A[m,k] = G[m,k] * scale[sample(m), k]  (ConvNeXt pwconv2 with the GRN factor folded in)
//
// GEMM view: D[M = output pixels, N = C_out] = A[M, K = taps*C_in] * W[N, K]^T, fp16 operands, fp32 accumulation.
// BLOCK_M = 128 (one UMMA M=128 atom, cta_group::1), BLOCK_N <= 256, two TMEM accumulator stages so the epilogue of
// tile i overlaps the main loop of tile i+1.  Small weight matrices stay resident in shared memory.
//
// Replaces: nn.Conv2d / nn.Linear + BatchNorm2d(eval) + ReLU / LayerNorm / GELU / GRN / tanh call sites of
// videoseal/modules/unet.py:17-197, modules/common.py:45-52,150-169, modules/convnext.py:41-57,
// modules/pixel_decoder.py:44-55 (see DESIGN.md for the per-layer mapping).
#pragma once
#include "ptx.cuh"

namespace vsb {

enum : int { LD_TMA = 0, LD_GATHER_CONV = 1, LD_HALO_UPS = 2, LD_GATHER_SCALE = 3, LD_HALO_CONV3 = 4 };
enum : int { EPI_AFFINE = 0, EPI_LN = 1 };
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

constexpr int kBlockM = 128;
constexpr int kMaxStages = 24;
constexpr int kMaxHalo = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;     // TMEM columns between the two accumulator stages
constexpr int kHeaderBytes = 11264;  // barriers | bias x2 | LN w,b / outc rows | outc partials (3 x [128][3])
constexpr int kHaloTW = 16, kHaloTH = 8;
constexpr int kHaloW = kHaloTW + 2, kHaloH = kHaloTH + 2;   // 18 x 10 input (or upsampled) halo of an 8x16 output tile

// division by a runtime constant as multiply-high + shift (valid for 0 <= n < 2^31); the host precomputes (mul, shr)
struct FastDiv {
  uint32_t mul, shr;
  int d;
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr); }
};

struct ConvGemmParams {
  FastDiv fd_ntiles, fd_tpi, fd_tx, fd_tw, fd_group, fd_rps, fd_hw, fd_cc, fd_ow, fd_oh, fd_ct, fd_s;
  // ---- GEMM view
  int M, N, num_kb, kblk, block_n, n_tiles, m_tiles, num_tiles, stages;
  uint32_t a_stage_bytes, b_stage_bytes, stage_bytes, idesc;
  int group;                      // consecutive M tiles that share one accumulator round (amortises per-tile handshakes for small N)
  int b_resident;                 // whole [block_n x K] weight slab lives in shared memory for the kernel's lifetime
  int b_fixed_ntile;              // b_resident with several N tiles: this CTA only ever sees N tile blockIdx.x % n_tiles
  int tma_store;                  // 1: fp16, 2: fp32 output tiles leave through TMA stores: every epilogue warp stages its 16-column chunks
  int cpw;                        //   ([32 rows][16 cols] boxes) in shared memory and one lane issues a bulk store per chunk; the warps of a
                                  //   lane quadrant own consecutive chunks (cpw = the largest count; the split may be uneven)
  uint32_t ostage_off, ostage_bytes;   // staging area: [epilogue warp][boxes][32 rows][ost_cpb * 16 cols]
  int ost_cpb;                    //   chunks per staging box (cpw: one box per warp; 1: one box per chunk)
  uint32_t ost_rowb;              //   bytes per box row
  int ost_swz;                    //   0 linear, 1 / 2 / 3: 32 / 64 / 128-byte TMA swizzle of the box rows
  int bias_global;                // epilogue reads the bias straight from global memory (warp-uniform 16-byte loads, L1 hits) instead
                                  // of a shared copy: no per-tile barrier between the epilogue warps (whole tiles only)
  uint32_t bres_off;
  // ---- output-tile geometry: tile_mode 0: rows are consecutive GEMM rows; 1: tile_h x tile_w pixel patch of an H x W map
  int tile_mode, H, W, tile_w, tile_h, tiles_x, tiles_per_img;
  // ---- A via TMA (LD_TMA): a_is_conv 0: 2D map (K, M); 1: 4D map (C, W, H, B) with zero-fill halo, one load per tap
  int a_is_conv, R, S, pad, c_blocks;
  // ---- halo loaders: the input is cut into `c_blocks` channel chunks of `cc` (= min(C, 64)) channels, the first c0_blocks
  //      from source 0; each chunk contributes kb_per_c K-blocks of 64 (K order inside a chunk: (tap, channel), zero padded)
  int c0_blocks, cc, kb_per_c;
  uint32_t halo_bytes /*TMA box bytes*/, halo_stride /*buffer pitch*/, halo_off, u_off;
  int halo_bufs;                  // ring depth of the halo tiles (prefetch distance of the producer)
  // ---- residual prefetch ring (thread-private slots)
  uint32_t resid_off, resid_stride;
  int resid_depth;
  int resid_direct;               // 1: no shared-memory ring, the epilogue reads the fp16 residual straight from global memory (long-K layers:
                                  //    their epilogue is hidden behind the main loop, the 48 KB buy one more TMA pipeline stage)
  // ---- gather loaders (NHWC fp16 sources; K order = (r, s, c) with c over the virtual concat [src0 | src1])
  const __half* src0;
  const __half* src1;
  int C0, C1, ld0, ld1, IH, IW, OH, OW, stride, pad_mode /*0 zero, 1 reflect*/, Ktot;
  const float* a_scale;  // LD_GATHER_SCALE: [num_samples, ld_scale] fp32
  int rows_per_sample, ld_scale;
  int b_sample_rows;              // > 0: per-sample weights [samples][N][K] (GRN folded into pwconv2); = N
  FastDiv fd_tps;                 // M tiles per sample
  // ---- epilogue
  int epi, act;
  const float* bias;     // [N] or null
  const __half* resid16; int ld_res16;   // added AFTER the activation (ResnetBlock: act(norm(conv)) + res)
  const float* resid32;  int ld_res32;
  __half* out16; int ld_out16;
  float* out32;  int ld_out32;
  const float* ln_w; const float* ln_b; float ln_eps;     // EPI_LN (requires n_tiles == 1)
  const float* outc_w; const float* outc_b; float* delta; int n_out, hw, outc_tanh;  // fused 1x1 outc (+tanh)
  float* grn_stats;      // [num_samples, N] sum of squares of the epilogue output (GRN), or null
  int grn_part;          // 1: grn_stats is [M / 32, N]: every epilogue warp STORES the column sums of its 32 rows (no atomics: the
                         //    consumer adds the rows of a sample in a fixed order -> bit-reproducible logits); needs
                         //    rows_per_sample % 32 == 0
};

// exact-erf GELU (nn.GELU default) with ONE special-function op per element:
//   gelu(x) = 0.5 x (1 + erf(x / sqrt2)) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt2),   erfc(|x| / sqrt2) = 2^Q(|x|)
// Q = degree-5 polynomial without constant term (erfc(0) = 1), weighted minimax fit of log2 erfc on [0, 3.3 sqrt2] (oracle-side
// check in tests/test_cpu_library.py): |gelu - exact| <= 1.0e-6 for every fp32 x in [-30, 30] (the tail runs to 2^-88 -> 0 on its
// own).  The version this replaces (Abramowitz-Stegun 7.1.25) needed rcp AND ex2 per element and was MUFU-bound in the pwconv1
// epilogue (profiles/r1_cnx_kernels_ncu.md: MIO stalls 19 %); its error was 2.5e-5.
constexpr float kGeluQ0 = -1.1510894298553467f, kGeluQ1 = -0.4592657685279846f, kGeluQ2 = -0.05254320427775383f,
                kGeluQ3 = 0.007386638317257166f, kGeluQ4 = -0.0005183160537853837f;
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  float q = fmaf(kGeluQ4, ax, kGeluQ3);
  q = fmaf(q, ax, kGeluQ2);
  q = fmaf(q, ax, kGeluQ1);
  q = fmaf(q, ax, kGeluQ0);
  const float e = ex2_approx(q * ax);
  return fmaf(-0.5f * ax, e, fmaxf(x, 0.f));
}

// two GELUs at once with Blackwell's packed fp32 arithmetic (fma/mul .f32x2): the polynomial costs half the issue slots
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
// packed in, packed out; max(x, 0) as 0.5 (x + |x|) so that the tail is two packed FMAs: gelu = 0.5 x + 0.5 |x| (1 - e)
__device__ __forceinline__ float2 gelu_erf_p(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  float2 q = f2fma(make_float2(kGeluQ4, kGeluQ4), ax, make_float2(kGeluQ3, kGeluQ3));
  q = f2fma(q, ax, make_float2(kGeluQ2, kGeluQ2));
  q = f2fma(q, ax, make_float2(kGeluQ1, kGeluQ1));
  q = f2fma(q, ax, make_float2(kGeluQ0, kGeluQ0));
  q = f2mul(q, ax);
  const float2 e = make_float2(ex2_approx(q.x), ex2_approx(q.y));
  const float2 hax = f2mul(ax, make_float2(0.5f, 0.5f));
  const float2 t = f2fma(make_float2(-e.x, -e.y), hax, hax);            // 0.5 |x| (1 - e)
  return f2fma(x, make_float2(0.5f, 0.5f), t);
}
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const float2 ax = make_float2(fabsf(x0), fabsf(x1));
  float2 q = f2fma(make_float2(kGeluQ4, kGeluQ4), ax, make_float2(kGeluQ3, kGeluQ3));
  q = f2fma(q, ax, make_float2(kGeluQ2, kGeluQ2));
  q = f2fma(q, ax, make_float2(kGeluQ1, kGeluQ1));
  q = f2fma(q, ax, make_float2(kGeluQ0, kGeluQ0));
  q = f2mul(q, ax);
  const float2 e = make_float2(ex2_approx(q.x), ex2_approx(q.y));
  const float2 r = f2fma(f2mul(ax, make_float2(-0.5f, -0.5f)), e, make_float2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  x0 = r.x; x1 = r.y;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// separable bilinear tap: (wy*(wx*a + (1-wx)*b) + (1-wy)*(wx*c + (1-wx)*d)) on 8 packed halves
__device__ __forceinline__ uint4 lerp2x2_h8(uint4 a, uint4 b, uint4 c, uint4 d, float wx, float wy) {
  const __half2 hx = __float2half2_rn(wx), hx1 = __float2half2_rn(1.f - wx);
  const __half2 hy = __float2half2_rn(wy), hy1 = __float2half2_rn(1.f - wy);
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  const __half2* pc = reinterpret_cast<const __half2*>(&c);
  const __half2* pd = reinterpret_cast<const __half2*>(&d);
  uint4 o;
  __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 t0 = __hfma2(hx, pa[i], __hmul2(hx1, pb[i]));
    const __half2 t1 = __hfma2(hx, pc[i], __hmul2(hx1, pd[i]));
    po[i] = __hfma2(hy, t0, __hmul2(hy1, t1));
  }
  return o;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, uint32_t src_bytes) {   // src_bytes 16 or 0
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (counted) arrival once all prior cp.async of this thread have completed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {   // wait until at most n groups are pending
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
template <int NT> __device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
__device__ __forceinline__ void bld_bar_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// epilogue warps: 16 for the pure-TMA kernels (GELU / GRN statistics are issue-bound; the small-K affine epilogues are bound by
// the TMEM-load -> math -> store chain of each warp, profiles/r2g_gemm_ncu.md: twice the warps hide twice the latency), 8 when 4
// builder warps also need registers
template <int LOADER, int ACT> struct EpiCfg { static constexpr int kWarps = (LOADER == LD_TMA) ? 16 : 8; };

template <int LOADER, int ACT>
__global__ void __launch_bounds__(LOADER == LD_TMA ? 640 : 512, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ ConvGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // header (kHeaderBytes): barriers | tmem ptr | s_bias[2][256] | s_lnw[256] s_lnb[256] s_oc2[256] | s_dot[128][3]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* hfull_bar = tempty_bar + 2;
  uint64_t* hempty_bar = hfull_bar + kMaxHalo;
  uint64_t* bres_bar = hempty_bar + kMaxHalo;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + 1024);        // [2][256]
  float* s_lnw = reinterpret_cast<float*>(smem + 3072);         // [256]  (LN weight | outc row 0)
  float* s_lnb = reinterpret_cast<float*>(smem + 4096);         // [256]  (LN bias   | outc row 1)
  float* s_oc2 = reinterpret_cast<float*>(smem + 5120);         // [256]  (            outc row 2)
  float* s_dot = reinterpret_cast<float*>(smem + 6144);         // [128][3]
  uint8_t* tiles = smem + kHeaderBytes;
  uint8_t* halo = smem + p.halo_off;
  uint8_t* ubuf = smem + p.u_off;
  uint8_t* bres = smem + p.bres_off;
  uint8_t* rbuf = smem + p.resid_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kNumBuilders = 128;
  constexpr bool kHalo = (LOADER == LD_HALO_CONV3 || LOADER == LD_HALO_UPS);
  constexpr int kEpiWarps = EpiCfg<LOADER, ACT>::kWarps;
  constexpr int kEpiThreads = kEpiWarps * 32;
  constexpr int kEpiSplit = kEpiWarps / 4;          // warps sharing one TMEM lane quadrant (interleaved 16-column chunks)
  constexpr int kBuilderWarp0 = 4 + kEpiWarps;

  if (threadIdx.x == 0) {
    const uint32_t full_count = LOADER == LD_TMA ? 1u : kNumBuilders + (p.b_resident ? 0u : 1u);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], full_count);
      mbar_init(&empty_bar[s], 1u);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1u);
      mbar_init(&tempty_bar[s], (uint32_t)kEpiThreads);
    }
    for (int s = 0; s < kMaxHalo; ++s) {
      mbar_init(&hfull_bar[s], 1u);
      mbar_init(&hempty_bar[s], kNumBuilders);
    }
    mbar_init(bres_bar, 1u);
    fence_barrier_init();
    if (LOADER == LD_TMA || kHalo) tma_prefetch_desc(&tmA);
    if (LOADER == LD_HALO_UPS || (LOADER == LD_TMA && p.c0_blocks != 0)) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmO);
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  if (p.epi == EPI_LN) {
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) { s_lnw[i] = p.ln_w[i]; s_lnb[i] = p.ln_b[i]; }
  } else if (p.outc_w != nullptr) {  // fused 1x1 outc: N <= 256 (n_tiles == 1), up to 3 output channels
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
      s_lnw[i] = p.outc_w[i];
      s_lnb[i] = p.n_out > 1 ? p.outc_w[p.N + i] : 0.f;
      s_oc2[i] = p.n_out > 2 ? p.outc_w[2 * p.N + i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // (PDL builds) nothing above reads or writes a tensor produced by an earlier kernel: only parameters and weights
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t row_bytes = (uint32_t)p.kblk * 2u;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int hb = 0;
      uint32_t hphase = 0;
      if (p.b_resident) {  // n_tiles == 1: the whole weight slab once
        mbar_arrive_expect_tx(bres_bar, (uint32_t)p.num_kb * p.b_stage_bytes);
        const int brow0 = p.b_fixed_ntile ? ((int)blockIdx.x - p.fd_ntiles.div((int)blockIdx.x) * p.n_tiles) * p.block_n : 0;
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(&tmB, bres_bar, bres + (size_t)kb * p.b_stage_bytes, kb * p.kblk, brow0);
      }
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
       const int st = p.fd_ntiles.div(tile), n_tile = tile - st * p.n_tiles;
       for (int g = 0; g < p.group; ++g) {
        const int m_tile = st * p.group + g;
        if (m_tile >= p.m_tiles) break;
        int cb = 0, cx = 0, cy = 0;
        if (p.tile_mode == 1) {
          cb = p.fd_tpi.div(m_tile);
          const int rem = m_tile - cb * p.tiles_per_img;
          const int ty = p.fd_tx.div(rem);
          cy = ty * p.tile_h;
          cx = (rem - ty * p.tiles_x) * p.tile_w;
        }
        if (kHalo) {
          for (int c = 0; c < p.c_blocks; ++c) {
            mbar_wait(&hempty_bar[hb], hphase ^ 1u);
            mbar_arrive_expect_tx(&hfull_bar[hb], p.halo_bytes);
            if (LOADER == LD_HALO_CONV3) {
              tma_load_4d(&tmA, &hfull_bar[hb], halo + (size_t)hb * p.halo_stride, c * p.cc, cx - 1, cy - 1, cb);
            } else {
              const bool s0 = c < p.c0_blocks;
              tma_load_4d(s0 ? &tmA : &tmA2, &hfull_bar[hb], halo + (size_t)hb * p.halo_stride,
                          (s0 ? c : c - p.c0_blocks) * p.cc, (cx >> 1) - 1, (cy >> 1) - 1, cb);
            }
            if (++hb == p.halo_bufs) { hb = 0; hphase ^= 1u; }
            if (!p.b_resident) {
              int kcoord = c * p.kb_per_c * 64;
              for (int kbi = 0; kbi < p.kb_per_c; ++kbi, kcoord += 64) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&full_bar[stage], p.b_stage_bytes);
                tma_load_2d(&tmB, &full_bar[stage], tiles + (size_t)stage * p.stage_bytes + p.a_stage_bytes, kcoord, n_tile * p.block_n);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
              }
            }
          }
        } else if (LOADER == LD_TMA) {
          int tap_r = 0, tap_s = 0, cblk = 0;
          const int brow = n_tile * p.block_n + (p.b_sample_rows ? p.fd_tps.div(m_tile) * p.b_sample_rows : 0);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = tiles + (size_t)stage * p.stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], p.a_stage_bytes + (p.b_resident ? 0u : p.b_stage_bytes));
            if (p.a_is_conv) {
              tma_load_4d(&tmA, &full_bar[stage], sa, cblk * p.kblk, cx + tap_s - p.pad, cy + tap_r - p.pad, cb);
              if (++cblk == p.c_blocks) { cblk = 0; if (++tap_s == p.S) { tap_s = 0; ++tap_r; } }
            } else {
              // 2D GEMM operand, optionally the virtual concat [A0 | A1] along K (c0_blocks K-blocks come from A0)
              if (p.c0_blocks == 0 || kb < p.c0_blocks) tma_load_2d(&tmA, &full_bar[stage], sa, kb * p.kblk, m_tile * kBlockM);
              else tma_load_2d(&tmA2, &full_bar[stage], sa, (kb - p.c0_blocks) * p.kblk, m_tile * kBlockM);
            }
            if (!p.b_resident) tma_load_2d(&tmB, &full_bar[stage], sa + p.a_stage_bytes, kb * p.kblk, brow);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        } else if (!p.b_resident) {  // gather loaders: weights only
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], p.b_stage_bytes);
            tma_load_2d(&tmB, &full_bar[stage], tiles + (size_t)stage * p.stage_bytes + p.a_stage_bytes, kb * p.kblk, n_tile * p.block_n);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
       }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int nmma = p.kblk >> 4;
      if (p.b_resident) { mbar_wait(bres_bar, 0u); tc_fence_after(); }
      const uint32_t bres_addr = smem_u32(bres);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const int st = p.fd_ntiles.div(tile);
       for (int g = 0; g < p.group; ++g) {
        if (st * p.group + g >= p.m_tiles) break;
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccStride + g * p.block_n);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (LOADER == LD_GATHER_CONV) fence_proxy_async_smem();   // cp.async-written A tile -> async proxy
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * p.stage_bytes);
          const uint64_t adesc = make_smem_desc(sa, row_bytes);
          const uint64_t bdesc = make_smem_desc(p.b_resident ? bres_addr + (uint32_t)kb * p.b_stage_bytes : sa + p.a_stage_bytes, row_bytes);
          for (int j = 0; j < nmma; ++j) {
            // advance 16 K-elements = 32 bytes inside the swizzled row: +2 in the (addr >> 4) field
            umma_f16_ss(d_tmem, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), p.idesc, (uint32_t)((kb | j) != 0));
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
       }
        umma_commit(&tfull_bar[as]);  // accumulators of the whole group complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ===================================================================== epilogue (8 or 16 warps)
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;  // which 16-column chunks: chunk index modulo kEpiSplit
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..kEpiThreads-1
    const int nchunks = p.block_n >> 4;
    // 16-column chunks of this warp: interleaved (half, half + kEpiSplit, ...) or, with TMA stores, cpw consecutive ones
    int c_beg = half, c_end = nchunks, c_step = kEpiSplit;
    if (p.tma_store) {
      const int base = nchunks / kEpiSplit, rem = nchunks - base * kEpiSplit;
      c_beg = half * base + min(half, rem);
      c_end = c_beg + base + (half < rem ? 1 : 0);
      c_step = 1;
    }
    uint8_t* ost = smem + p.ostage_off + (size_t)(warp - 4) * p.ostage_bytes;   // this warp's output staging boxes
    const uint32_t ost_box = 32u * p.ost_rowb;
    const uint32_t ost_upc = p.tma_store == 2 ? 4u : 2u;                         // 16-byte units per chunk
    // 16-byte unit u of row r of a swizzled box lives at unit u ^ f(r) (CU_TENSOR_MAP_SWIZZLE_{32,64,128}B)
    const uint32_t ost_x = p.ost_swz == 3 ? (uint32_t)(lane & 7) : (p.ost_swz == 2 ? (uint32_t)((lane >> 1) & 3) : (p.ost_swz == 1 ? (uint32_t)((lane >> 2) & 1) : 0u));
    auto ost_ptr = [&](int k /*chunk of this warp*/, uint32_t t /*16-byte unit inside the chunk*/) -> uint4* {
      const int box = p.ost_cpb == 1 ? k : 0, kin = p.ost_cpb == 1 ? 0 : k;
      return reinterpret_cast<uint4*>(ost + (size_t)box * ost_box + (size_t)lane * p.ost_rowb + (((uint32_t)kin * ost_upc + t) ^ ost_x) * 16u);
    };
    const int D = p.resid_depth;
    const bool has_res = (p.resid16 != nullptr) || (p.resid32 != nullptr);

    // geometry of this thread's row in M tile `m_tile`
    auto tile_row = [&](int m_tile, long& m, bool& mvalid) {
      if (p.tile_mode == 1) {
        const int b = p.fd_tpi.div(m_tile);
        const int rem = m_tile - b * p.tiles_per_img;
        const int ty = p.fd_tx.div(rem), tx = rem - ty * p.tiles_x;
        const int ry = p.fd_tw.div(row), rx = row - ry * p.tile_w;
        m = ((long)b * p.H + (ty * p.tile_h + ry)) * p.W + (tx * p.tile_w + rx);
        mvalid = true;
      } else {
        m = (long)m_tile * kBlockM + row;
        mvalid = m < p.M;
      }
    };
    // residual row of sub-tile number `sq` (= iteration * group + g of this CTA) -> ring slot (thread-private region,
    // layout [16-byte chunk][row] = conflict-free)
    const int G = p.group;
    auto prefetch_resid = [&](int sq, int slot) {
      if (has_res && !p.resid_direct) {
        const int it = p.fd_group.div(sq), g = sq - it * G;
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int st = p.fd_ntiles.div(tile);
        const int m_tile = st * G + g;
        if (tile < p.num_tiles && m_tile < p.m_tiles) {
          const int n0 = (tile - st * p.n_tiles) * p.block_n;
          long m; bool mvalid;
          tile_row(m_tile, m, mvalid);
          if (mvalid && n0 + p.block_n <= p.N) {
            uint8_t* dst = rbuf + (size_t)slot * p.resid_stride;
            if (p.resid16 != nullptr) {
              const __half* r = p.resid16 + m * p.ld_res16 + n0;
              for (int ch = c_beg; ch < c_end; ch += c_step) {
                cp_async16(dst + ((size_t)(2 * ch) * 128 + row) * 16, r + ch * 16);
                cp_async16(dst + ((size_t)(2 * ch + 1) * 128 + row) * 16, r + ch * 16 + 8);
              }
            } else {
              const float* r = p.resid32 + m * p.ld_res32 + n0;
              for (int ch = c_beg; ch < c_end; ch += c_step) {
#pragma unroll
                for (int u = 0; u < 4; ++u) cp_async16(dst + ((size_t)(4 * ch + u) * 128 + row) * 16, r + ch * 16 + 4 * u);
              }
            }
          }
        }
      }
      cp_async_commit();
    };

    int as = 0;
    uint32_t aphase = 0;
    int bsel = 0;
    int slot = 0;
    int sq = 0;
    for (int i = 0; i < D - 1; ++i) prefetch_resid(i, i);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int st = p.fd_ntiles.div(tile);
      const int n0 = (tile - st * p.n_tiles) * p.block_n;
      // ---- prologue, overlapped with the main loop: bias -> smem
      const float* sb = s_bias + bsel * 256;
      if (p.bias_global) {
        sb = p.bias + n0;
      } else {
        if (et < p.block_n) s_bias[bsel * 256 + et] = (p.bias != nullptr && n0 + et < p.N) ? __ldg(p.bias + n0 + et) : 0.f;
        epi_bar_sync<kEpiThreads>();  // bias visible to all epilogue warps (double-buffered across tiles)
      }
     for (int g = 0; g < G; ++g, ++sq) {
      // residual of the sub-tile D-1 ahead -> ring (issued before waiting for this group's accumulators)
      {
        int pslot = slot + D - 1;
        if (pslot >= D) pslot -= D;
        prefetch_resid(sq + D - 1, pslot);
      }
      if (g == 0) {
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
      }
      cp_async_wait_pending(D - 1);
      const int m_tile = st * G + g;
      const uint8_t* rb_ = rbuf + (size_t)slot * p.resid_stride;
      if (++slot == D) slot = 0;
      if (m_tile >= p.m_tiles) continue;     // ragged last group (uniform across the CTA)
      long m; bool mvalid;
      tile_row(m_tile, m, mvalid);
      const bool res_fast = mvalid && (n0 + p.block_n <= p.N);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccStride + g * p.block_n);

      if (p.epi == EPI_LN) {
        // channels-first LayerNorm over the full C_out row (biased variance), then activation.  The kEpiSplit warps that share a
        // TMEM lane quadrant split the 16-column chunks: pass 1 = per-warp partial (sum, sum of squares) of its chunks, exchanged
        // through shared memory and added in a fixed order; pass 2 = normalise + store its chunks.  (The first version let one warp
        // per quadrant walk the whole row three times: stem GEMM 124 us at 19 TF/s.)
        {
          const int N = p.N;
          float* s_part = s_dot;                                  // [kEpiSplit][128][2]
          float sum = 0.f, sq = 0.f;
          for (int ch = half; ch < nchunks; ch += kEpiSplit) {
            const int c = ch * 16;
            float v[16];
            tmem_ld16(trow + c, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x = (c + j < N) ? v[j] + sb[c + j] : 0.f;
              sum += x;
              sq = fmaf(x, x, sq);
            }
          }
          s_part[(half * 128 + row) * 2] = sum;
          s_part[(half * 128 + row) * 2 + 1] = sq;
          epi_bar_sync<kEpiThreads>();
          sum = 0.f; sq = 0.f;
#pragma unroll
          for (int h2 = 0; h2 < kEpiSplit; ++h2) { sum += s_part[(h2 * 128 + row) * 2]; sq += s_part[(h2 * 128 + row) * 2 + 1]; }
          const float mean = sum / (float)N;
          const float rstd = 1.0f / sqrtf(fmaxf(sq / (float)N - mean * mean, 0.f) + p.ln_eps);
          for (int ch = half; ch < nchunks; ch += kEpiSplit) {
            const int c = ch * 16;
            float v[16];
            tmem_ld16(trow + c, v);
            if (c >= N) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = c + j;
              float y = 0.f;
              if (n < N) {
                y = (v[j] + sb[c + j] - mean) * rstd * s_lnw[n] + s_lnb[n];
                if (ACT == ACT_RELU) y = fmaxf(y, 0.f);
                else if (ACT == ACT_GELU) y = gelu_erf(y);
              }
              v[j] = y;
            }
            if (mvalid) {
              if (p.out32 != nullptr) {
                float* o = p.out32 + m * p.ld_out32 + c;
                if (c + 16 <= N) {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    reinterpret_cast<float4*>(o)[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                } else {
                  for (int j = 0; j < 16 && c + j < N; ++j) o[j] = v[j];
                }
              }
              if (p.out16 != nullptr) {
                __align__(16) __half h[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = __float2half_rn(v[j]);
                __half* o = p.out16 + m * p.ld_out16 + c;
                if (c + 16 <= N) {
                  reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<const uint4*>(h)[0];
                  reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<const uint4*>(h)[1];
                } else {
                  for (int j = 0; j < 16 && c + j < N; ++j) o[j] = h[j];
                }
              }
            }
          }
          epi_bar_sync<kEpiThreads>();     // s_part is rewritten by the next sub-tile
        }
      } else {
        float dot0 = 0.f, dot1 = 0.f, dot2 = 0.f;
        const bool grn_uniform =
            p.grn_stats != nullptr &&
            p.fd_rps.div(m_tile * kBlockM + q * 32) == p.fd_rps.div(m_tile * kBlockM + q * 32 + 31) &&
            (m_tile * kBlockM + q * 32 + 31) < p.M;
        float* grn_row = !grn_uniform ? nullptr
                         : (p.grn_part ? p.grn_stats + (long)(m_tile * 4 + q) * p.N
                                       : p.grn_stats + (long)p.fd_rps.div(m_tile * kBlockM + q * 32) * p.N);
        // whole tile, fp16 output only, no residual / fused 1x1, statistics (if any) uniform per warp
        const bool whole = (m_tile + 1) * kBlockM <= p.M && n0 + p.block_n <= p.N && p.out16 != nullptr && p.out32 == nullptr &&
                           p.outc_w == nullptr && (p.ld_out16 & 7) == 0;
        const bool lean32 = ACT == ACT_NONE && (m_tile + 1) * kBlockM <= p.M && n0 + p.block_n <= p.N && p.out32 != nullptr &&
                            p.outc_w == nullptr && p.grn_stats == nullptr && p.resid32 != nullptr && res_fast && (p.ld_out32 & 3) == 0 &&
                            (p.out16 == nullptr || (p.ld_out16 & 7) == 0);
        const bool lean = ACT == ACT_GELU ? (whole && !has_res && (p.grn_stats == nullptr || grn_uniform))
                                          : (whole && p.grn_stats == nullptr && (!has_res || (p.resid16 != nullptr && res_fast)));
        uint32_t vnext[16];
        if (c_beg < c_end) tmem_ld16_issue(trow + c_beg * 16, vnext);
        const bool use_ts = p.tma_store == 1 && lean;      // this warp's part of the tile leaves through its staging boxes + TMA stores
        const bool use_ts32 = p.tma_store == 2 && lean32 && p.out16 == nullptr;
        if (use_ts || use_ts32) {                          // the previous stores of this warp have finished reading the staging boxes
          if (lane == 0) bulk_wait_group_read0();
          __syncwarp();
        }
        if (ACT == ACT_GELU && lean) {
          // pwconv1 on whole tiles (every shipped card): bias + GELU + fp16 store + GRN column statistics and nothing else, in packed
          // fp32 arithmetic throughout.  The accumulators are consumed straight out of the TMEM load registers by the packed bias add,
          // so the next chunk's load can be issued without a register copy; one saturating F2FP per output pair.
          const __half* orow = p.out16 + m * p.ld_out16 + n0;
          for (int ch = c_beg; ch < c_end; ch += c_step) {
            const int c = ch * 16;
            tmem_ld_wait(vnext);
            const float4* sb4 = reinterpret_cast<const float4*>(sb + c);
            float2 w2[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 bq = sb4[u];
              w2[2 * u] = __fadd2_rn(make_float2(__uint_as_float(vnext[4 * u]), __uint_as_float(vnext[4 * u + 1])), make_float2(bq.x, bq.y));
              w2[2 * u + 1] = __fadd2_rn(make_float2(__uint_as_float(vnext[4 * u + 2]), __uint_as_float(vnext[4 * u + 3])), make_float2(bq.z, bq.w));
            }
            if (ch + c_step < c_end) tmem_ld16_issue(trow + c + 16 * c_step, vnext);   // overlaps the math below
#pragma unroll
            for (int j = 0; j < 8; ++j) w2[j] = gelu_erf_p(w2[j]);
            {
              uint32_t h[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(w2[j].y), "f"(w2[j].x));   // {hi, lo}
              if (use_ts) {
                *ost_ptr(ch - c_beg, 0) = make_uint4(h[0], h[1], h[2], h[3]);
                *ost_ptr(ch - c_beg, 1) = make_uint4(h[4], h[5], h[6], h[7]);
              } else {
                uint4* o = reinterpret_cast<uint4*>(const_cast<__half*>(orow) + c);
                o[0] = make_uint4(h[0], h[1], h[2], h[3]);
                o[1] = make_uint4(h[4], h[5], h[6], h[7]);
              }
            }
            if (grn_row != nullptr) {
              float sq[16];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 t2 = __fmul2_rn(w2[j], w2[j]);
                sq[2 * j] = t2.x; sq[2 * j + 1] = t2.y;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float send = (lane & 16) ? sq[i] : sq[i + 8];
                const float keep = (lane & 16) ? sq[i + 8] : sq[i];
                sq[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float send = (lane & 8) ? sq[i] : sq[i + 4];
                const float keep = (lane & 8) ? sq[i + 4] : sq[i];
                sq[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float send = (lane & 4) ? sq[i] : sq[i + 2];
                const float keep = (lane & 4) ? sq[i + 2] : sq[i];
                sq[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              {
                const float send = (lane & 2) ? sq[0] : sq[1];
                const float keep = (lane & 2) ? sq[1] : sq[0];
                sq[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
              }
              sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], 1);
              if ((lane & 1) == 0) {
                if (p.grn_part) grn_row[n0 + c + ((lane >> 1) & 15)] = sq[0];
                else atomicAdd(grn_row + n0 + c + ((lane >> 1) & 15), sq[0]);
              }
            }
          }
        } else
        for (int ch = c_beg; ch < c_end; ch += c_step) {
          const int c = ch * 16;
          float v[16];
          tmem_ld_wait(vnext);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(vnext[j]);
          if (ch + c_step < c_end) tmem_ld16_issue(trow + c + 16 * c_step, vnext);   // overlaps the math below
          const int n = n0 + c;
          if (ACT == ACT_NONE && lean32) {
            // pwconv2 on whole tiles: bias + fp32 residual stream from the prefetch ring, updated in place (+ fp16 copy)
            const float4* sb4 = reinterpret_cast<const float4*>(sb + c);
            float4* o = reinterpret_cast<float4*>(p.out32 + m * p.ld_out32 + n);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 bq = sb4[u];
              const float4 t = *reinterpret_cast<const float4*>(rb_ + ((size_t)(4 * ch + u) * 128 + row) * 16);
              v[4 * u + 0] += bq.x + t.x; v[4 * u + 1] += bq.y + t.y; v[4 * u + 2] += bq.z + t.z; v[4 * u + 3] += bq.w + t.w;
              float4* dst = use_ts32 ? reinterpret_cast<float4*>(ost_ptr(ch - c_beg, (uint32_t)u)) : o + u;
              *dst = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            }
            if (p.out16 != nullptr) {
              __align__(16) __half2 h2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                h2[j] = __floats2half2_rn(fminf(fmaxf(v[2 * j], -65504.f), 65504.f), fminf(fmaxf(v[2 * j + 1], -65504.f), 65504.f));
              uint4* o16 = reinterpret_cast<uint4*>(p.out16 + m * p.ld_out16 + n);
              o16[0] = reinterpret_cast<const uint4*>(h2)[0];
              o16[1] = reinterpret_cast<const uint4*>(h2)[1];
            }
            continue;
          }
          if (ACT != ACT_GELU && lean) {
            // whole-tile conv / GEMM with an fp16 output: bias, ReLU, optional fp16 residual from the prefetch ring, store
            const float4* sb4 = reinterpret_cast<const float4*>(sb + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 bq = sb4[u];
              v[4 * u + 0] += bq.x; v[4 * u + 1] += bq.y; v[4 * u + 2] += bq.z; v[4 * u + 3] += bq.w;
            }
            if (ACT == ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (has_res) {
              __align__(16) __half h[16];
              if (p.resid_direct) {
                const uint4* rg = reinterpret_cast<const uint4*>(p.resid16 + m * p.ld_res16 + n);
                reinterpret_cast<uint4*>(h)[0] = __ldg(rg);
                reinterpret_cast<uint4*>(h)[1] = __ldg(rg + 1);
              } else {
                reinterpret_cast<uint4*>(h)[0] = *reinterpret_cast<const uint4*>(rb_ + ((size_t)(2 * ch) * 128 + row) * 16);
                reinterpret_cast<uint4*>(h)[1] = *reinterpret_cast<const uint4*>(rb_ + ((size_t)(2 * ch + 1) * 128 + row) * 16);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __half2float(h[j]);
            }
            __align__(16) __half2 h2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a = fminf(v[2 * j], 65504.f), b2 = fminf(v[2 * j + 1], 65504.f);
              if (ACT == ACT_NONE || has_res) { a = fmaxf(a, -65504.f); b2 = fmaxf(b2, -65504.f); }
              h2[j] = __floats2half2_rn(a, b2);
            }
            if (use_ts) {
              *ost_ptr(ch - c_beg, 0) = reinterpret_cast<const uint4*>(h2)[0];
              *ost_ptr(ch - c_beg, 1) = reinterpret_cast<const uint4*>(h2)[1];
            } else {
              uint4* o = reinterpret_cast<uint4*>(p.out16 + m * p.ld_out16 + n);
              o[0] = reinterpret_cast<const uint4*>(h2)[0];
              o[1] = reinterpret_cast<const uint4*>(h2)[1];
            }
            continue;
          }
          if (n >= p.N) continue;  // uniform across the warp
          const int nval = min(16, p.N - n);   // 16 on every full chunk (all shipped tiny/pixelseal layers)
          {
            const float4* sb4 = reinterpret_cast<const float4*>(sb + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 bq = sb4[u];
              v[4 * u + 0] += bq.x; v[4 * u + 1] += bq.y; v[4 * u + 2] += bq.z; v[4 * u + 3] += bq.w;
            }
          }
          if (ACT == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) gelu_erf2(v[j], v[j + 1]);
          } else if (ACT == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (nval < 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j >= nval) v[j] = 0.f;   // ragged last chunk (chunky widths): keep the tail inert
          }
          if (has_res && mvalid) {
            if (res_fast && !p.resid_direct) {
              if (p.resid16 != nullptr) {
                __align__(16) __half h[16];
                reinterpret_cast<uint4*>(h)[0] = *reinterpret_cast<const uint4*>(rb_ + ((size_t)(2 * ch) * 128 + row) * 16);
                reinterpret_cast<uint4*>(h)[1] = *reinterpret_cast<const uint4*>(rb_ + ((size_t)(2 * ch + 1) * 128 + row) * 16);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __half2float(h[j]);
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float4 t = *reinterpret_cast<const float4*>(rb_ + ((size_t)(4 * ch + u) * 128 + row) * 16);
                  v[4 * u + 0] += t.x; v[4 * u + 1] += t.y; v[4 * u + 2] += t.z; v[4 * u + 3] += t.w;
                }
              }
            } else if (p.resid16 != nullptr) {
              const __half* r = p.resid16 + m * p.ld_res16 + n;
#pragma unroll
              for (int j = 0; j < 16; ++j) if (j < nval) v[j] += __half2float(r[j]);
            } else {
              const float* r = p.resid32 + m * p.ld_res32 + n;
#pragma unroll
              for (int j = 0; j < 16; ++j) if (j < nval) v[j] += r[j];
            }
          }
          if (p.outc_w != nullptr) {   // N <= 256 and a multiple of 16 here (host-checked): no tail
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              dot0 += v[j] * s_lnw[n + j];
              dot1 += v[j] * s_lnb[n + j];
              dot2 += v[j] * s_oc2[n + j];
            }
          }
          if (mvalid) {
            if (p.out16 != nullptr) {
              __half* o = p.out16 + m * p.ld_out16 + n;
              __align__(16) __half2 h[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float a = fminf(v[2 * j], 65504.f), b2 = fminf(v[2 * j + 1], 65504.f);
                if (ACT == ACT_NONE || has_res) { a = fmaxf(a, -65504.f); b2 = fmaxf(b2, -65504.f); }   // ReLU >= 0, GELU >= -0.17
                h[j] = __floats2half2_rn(a, b2);
              }
              if (nval == 16) {
                reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<const uint4*>(h)[0];
                reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<const uint4*>(h)[1];
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (2 * j < nval) o[2 * j] = __low2half(h[j]);
                  if (2 * j + 1 < nval) o[2 * j + 1] = __high2half(h[j]);
                }
              }
            }
            if (p.out32 != nullptr) {
              float* o = p.out32 + m * p.ld_out32 + n;
              if (nval == 16) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  reinterpret_cast<float4*>(o)[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < nval) o[j] = v[j];
              }
            }
          }
          if (p.grn_stats != nullptr) {
            // column sums of squares over this warp's 32 rows (GRN: ||x||_2 over H,W per (sample, channel))
            float sq[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) sq[j] = mvalid ? v[j] * v[j] : 0.f;
            if (grn_uniform) {
              // halving butterfly: after xor 16,8,4,2 each lane holds one column summed over 16 rows
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float send = (lane & 16) ? sq[i] : sq[i + 8];
                const float keep = (lane & 16) ? sq[i + 8] : sq[i];
                sq[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float send = (lane & 8) ? sq[i] : sq[i + 4];
                const float keep = (lane & 8) ? sq[i + 4] : sq[i];
                sq[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float send = (lane & 4) ? sq[i] : sq[i + 2];
                const float keep = (lane & 4) ? sq[i + 2] : sq[i];
                sq[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              {
                const float send = (lane & 2) ? sq[0] : sq[1];
                const float keep = (lane & 2) ? sq[1] : sq[0];
                sq[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
              }
              sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], 1);
              const int col = (lane >> 1) & 15;
              if ((lane & 1) == 0 && col < nval) {
                if (p.grn_part) grn_row[n + col] = sq[0];
                else atomicAdd(grn_row + n + col, sq[0]);
              }
            } else if (mvalid) {
              float* gr = p.grn_stats + (long)p.fd_rps.div((int)m) * p.N + n;
#pragma unroll
              for (int j = 0; j < 16; ++j) if (j < nval) atomicAdd(gr + j, sq[j]);
            }
          }
        }
        if (use_ts || use_ts32) {
          // staging boxes complete: make the generic-proxy writes visible to the async proxy, then ONE lane hands the 32 x 16 boxes to
          // the TMA unit (full 32-byte sectors, no LSU tag traffic: the 16-byte-per-row register stores this replaces cost 32 tag
          // cycles per warp instruction and were the bottleneck of every short-K layer)
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int k = 0; k < c_end - c_beg; k += p.ost_cpb)
              tma_store_2d(&tmO, ost + (size_t)(p.ost_cpb == 1 ? k : 0) * ost_box, n0 + (c_beg + k) * 16, (int)(m - lane));
            bulk_commit_group();
          }
        }
        if (p.outc_w != nullptr) {
          // the warps sharing a row each hold a partial dot product: combine through shared memory
          if (half != 0) {
            float* d = s_dot + ((half - 1) * 128 + row) * 3;
            d[0] = dot0; d[1] = dot1; d[2] = dot2;
          }
          epi_bar_sync<kEpiThreads>();
          if (half == 0 && mvalid) {
#pragma unroll
            for (int h2 = 1; h2 < kEpiSplit; ++h2) {
              const float* d = s_dot + ((h2 - 1) * 128 + row) * 3;
              dot0 += d[0]; dot1 += d[1]; dot2 += d[2];
            }
            const int b = p.fd_hw.div((int)m), pix = (int)m - b * p.hw;
            const float dd[3] = {dot0, dot1, dot2};
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              if (o < p.n_out) {
                float d = dd[o] + __ldg(p.outc_b + o);
                if (p.outc_tanh) d = tanhf(d);
                p.delta[((long)b * p.n_out + o) * p.hw + pix] = d;
              }
            }
          }
        }
      }
     }  // g
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
      bsel ^= 1;
    }
    if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores of this warp have landed
  } else if (LOADER != LD_TMA && warp >= kBuilderWarp0) {
    // ===================================================================== A-tile builders (4 warps)
    const int gt = threadIdx.x - kBuilderWarp0 * 32;  // 0..127
    const int j = gt & 7;              // 16-byte chunk (8 fp16 K-elements) within the 128-byte swizzled row
    const int rg = gt >> 3;            // rows rg + 16*i
    int stage = 0;
    uint32_t phase = 0;
    if (kHalo) {
      const uint32_t pb = (uint32_t)p.cc * 2u;   // bytes per halo pixel
      const int cpp = p.cc >> 3;                 // 16-byte chunks per halo pixel
      int hb = 0;
      uint32_t hphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
       for (int g = 0; g < p.group; ++g) {
        const int m_tile = p.fd_ntiles.div(tile) * p.group + g;
        if (m_tile >= p.m_tiles) break;
        int cy, cx;
        {
          const int b = p.fd_tpi.div(m_tile);
          const int rem = m_tile - b * p.tiles_per_img;
          const int ty = p.fd_tx.div(rem);
          cy = ty * p.tile_h;
          cx = (rem - ty * p.tiles_x) * p.tile_w;
        }
        for (int c = 0; c < p.c_blocks; ++c) {
          mbar_wait(&hfull_bar[hb], hphase);
          const uint8_t* hsrc = halo + (size_t)hb * p.halo_stride;
          if (LOADER == LD_HALO_UPS) {
            // phase 1: build the 10 x 18 upsampled + reflect-padded halo (once per channel chunk) from the 6 x 10
            // low-resolution tile that starts at source pixel (cy/2 - 1, cx/2 - 1)
            bld_bar_sync();  // everyone is done reading the previous U
            const int sy0 = (cy >> 1) - 1, sx0 = (cx >> 1) - 1;
            for (int e = gt; e < kHaloH * kHaloW * cpp; e += 128) {
              const int ch = e % cpp;
              const int px = e / cpp;
              const int hy = px / kHaloW, hx = px - hy * kHaloW;
              const int uy = reflect_idx(cy + hy - 1, 2 * p.IH);
              const int ux = reflect_idx(cx + hx - 1, 2 * p.IW);
              const int iy = uy >> 1, ix = ux >> 1;
              int ya, yb, xa, xb;
              float wya, wxa;
              if (uy & 1) { ya = iy; yb = min(iy + 1, p.IH - 1); wya = 0.75f; }
              else        { ya = max(iy - 1, 0); yb = iy; wya = 0.25f; }
              if (ux & 1) { xa = ix; xb = min(ix + 1, p.IW - 1); wxa = 0.75f; }
              else        { xa = max(ix - 1, 0); xb = ix; wxa = 0.25f; }
              const uint8_t* base = hsrc + ch * 16;
              const uint4 v00 = *reinterpret_cast<const uint4*>(base + (size_t)((ya - sy0) * 10 + (xa - sx0)) * pb);
              const uint4 v01 = *reinterpret_cast<const uint4*>(base + (size_t)((ya - sy0) * 10 + (xb - sx0)) * pb);
              const uint4 v10 = *reinterpret_cast<const uint4*>(base + (size_t)((yb - sy0) * 10 + (xa - sx0)) * pb);
              const uint4 v11 = *reinterpret_cast<const uint4*>(base + (size_t)((yb - sy0) * 10 + (xb - sx0)) * pb);
              *reinterpret_cast<uint4*>(ubuf + (size_t)px * pb + ch * 16) = lerp2x2_h8(v00, v01, v10, v11, wxa, wya);
            }
            mbar_arrive(&hempty_bar[hb]);  // low-res tile consumed
            bld_bar_sync();                // U complete
            hsrc = ubuf;
          }
          // phase 2: tap tiles.  K-block kbi of this chunk holds K indices [64*kbi, 64*kbi+64) of the (tap, channel)
          // order; 16-byte chunk j covers k = 64*kbi + 8*j -> tap = k / cc, channel offset = k % cc
          const int stage0 = stage;
          const uint32_t phase0 = phase;
          for (int kbi = 0; kbi < p.kb_per_c; ++kbi) {
            const int k = kbi * 64 + j * 8;
            const int tap = p.fd_cc.div(k);
            const int coff = (k - tap * p.cc) * 2;
            const int tr = tap / 3, ts = tap - tr * 3;
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = tiles + (size_t)stage * p.stage_bytes;
            if (tap < 9) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = rg + 16 * i;
                const int ry = r >> 4, rx = r & 15;
                const uint4 val = *reinterpret_cast<const uint4*>(hsrc + (size_t)((ry + tr) * kHaloW + rx + ts) * pb + coff);
                *reinterpret_cast<uint4*>(sa + r * 128 + ((j ^ (r & 7)) << 4)) = val;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = rg + 16 * i;
                *reinterpret_cast<uint4*>(sa + r * 128 + ((j ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
              }
            }
            if (p.kb_per_c > p.stages) {   // ring shorter than the chunk: publish each K-block immediately
              fence_proxy_async_smem();
              mbar_arrive(&full_bar[stage]);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          if (p.kb_per_c <= p.stages) {
            // one generic->async proxy fence for the whole chunk, then publish its K-blocks
            fence_proxy_async_smem();
            int sp = stage0;
            for (int kbi = 0; kbi < p.kb_per_c; ++kbi) {
              mbar_arrive(&full_bar[sp]);
              if (++sp == p.stages) sp = 0;
            }
          }
          (void)phase0;
          if (LOADER == LD_HALO_CONV3) mbar_arrive(&hempty_bar[hb]);
          if (++hb == p.halo_bufs) { hb = 0; hphase ^= 1u; }
        }
       }
      }
    } else {
      const int Ct = p.C0 + p.C1;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
       for (int g = 0; g < p.group; ++g) {
        const int m_tile = p.fd_ntiles.div(tile) * p.group + g;
        if (m_tile >= p.m_tiles) break;
        int pb[8], py[8], px[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m_tile * kBlockM + rg + 16 * i;      // M < 2^31 (host-checked)
          if (m < p.M) {
            if (LOADER == LD_GATHER_SCALE) {
              pb[i] = p.fd_rps.div(m);
              py[i] = 0;
              px[i] = m;
            } else {
              const int t = p.fd_ow.div(m);
              px[i] = m - t * p.OW;
              pb[i] = p.fd_oh.div(t);
              py[i] = t - pb[i] * p.OH;
            }
          } else {
            pb[i] = -1; py[i] = 0; px[i] = 0;
          }
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = tiles + (size_t)stage * p.stage_bytes;
          const int k = kb * 64 + j * 8;
          const bool kvalid = k < p.Ktot;
          if (LOADER == LD_GATHER_SCALE) {
            uint4 g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              g[i] = make_uint4(0, 0, 0, 0);
              if (kvalid && pb[i] >= 0) g[i] = __ldg(reinterpret_cast<const uint4*>(p.src0 + (long)px[i] * p.ld0 + k));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 val = make_uint4(0, 0, 0, 0);
              if (kvalid && pb[i] >= 0) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.a_scale + (long)pb[i] * p.ld_scale + k));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.a_scale + (long)pb[i] * p.ld_scale + k) + 1);
                const __half2* ph = reinterpret_cast<const __half2*>(&g[i]);
                __half2* po = reinterpret_cast<__half2*>(&val);
                float2 f;
                f = __half22float2(ph[0]); f.x *= s0.x; f.y *= s0.y; po[0] = __float22half2_rn(f);
                f = __half22float2(ph[1]); f.x *= s0.z; f.y *= s0.w; po[1] = __float22half2_rn(f);
                f = __half22float2(ph[2]); f.x *= s1.x; f.y *= s1.y; po[2] = __float22half2_rn(f);
                f = __half22float2(ph[3]); f.x *= s1.z; f.y *= s1.w; po[3] = __float22half2_rn(f);
              }
              const int r = rg + 16 * i;
              *reinterpret_cast<uint4*>(sa + r * 128 + ((j ^ (r & 7)) << 4)) = val;
            }
          } else {
            const int tap = kvalid ? p.fd_ct.div(k) : 0;
            const int c = k - tap * Ct;
            const int tr = p.fd_s.div(tap), ts = tap - tr * p.S;
            const __half* src = (c < p.C0) ? p.src0 : p.src1;
            const int ld = (c < p.C0) ? p.ld0 : p.ld1;
            const int cc = (c < p.C0) ? c : c - p.C0;
            // asynchronous gather: 16-byte cp.async per (row, chunk) with zero fill for padding / tails; the mbarrier
            // arrival fires when this thread's copies have landed, so the builder runs ahead by the ring depth
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half* gsrc = p.src0;   // always a valid address (no bytes are read when nbytes == 0)
              uint32_t nbytes = 0;
              if (kvalid && pb[i] >= 0) {
                int iy = py[i] * p.stride + tr - p.pad;
                int ix = px[i] * p.stride + ts - p.pad;
                bool ok = true;
                if (p.pad_mode == 1) {
                  iy = reflect_idx(iy, p.IH);
                  ix = reflect_idx(ix, p.IW);
                } else {
                  ok = (iy >= 0) && (iy < p.IH) && (ix >= 0) && (ix < p.IW);
                }
                if (ok) { gsrc = src + (((long)pb[i] * p.IH + iy) * p.IW + ix) * ld + cc; nbytes = 16; }
              }
              const int r = rg + 16 * i;
              cp_async16_zfill(sa + r * 128 + ((j ^ (r & 7)) << 4), gsrc, nbytes);
            }
            cp_async_mbar_arrive_noinc(&full_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
       }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace vsb
