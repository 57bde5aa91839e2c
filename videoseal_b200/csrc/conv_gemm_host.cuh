// Host-side construction + launch of conv_gemm_kernel instances (tensor maps, tiling, smem budget).
#pragma once
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "conv_gemm.cuh"
#include "conv_pair.cuh"

namespace vsb {

// status codes of include/vsb200.h (static_assert-ed equal in vsb200.cu): every throw site names its code, the C ABI returns it
enum : int { kErrInvalid = -1, kErrUnsupported = -2, kErrCuda = -3, kErrState = -4 };
struct Error : std::runtime_error {
  int code;
  explicit Error(const std::string& m, int c = kErrInvalid) : std::runtime_error(m), code(c) {}
};

#define VSB_CHECK(cond, msg)                                                                      \
  do {                                                                                            \
    if (!(cond)) throw ::vsb::Error(std::string(msg) + " [" #cond "] at " __FILE__ ":" + std::to_string(__LINE__)); \
  } while (0)

#define VSB_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      throw ::vsb::Error(std::string("CUDA error ") + cudaGetErrorString(_e) + " in " #expr " at " __FILE__ ":" + std::to_string(__LINE__), ::vsb::kErrCuda); \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    VSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    VSB_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

inline CUtensorMapSwizzle swizzle_for(int kblk) {
  return kblk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kblk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// fp16 (or fp32: f32 = true, always unswizzled) tensor map, dims innermost-first; strides (bytes) for dims 1..rank-1
inline void encode_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int kblk, bool no_swizzle = false, bool f32 = false, int swz_bytes = -1 /* >= 0: explicit 0/32/64/128 */) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  VSB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  for (int i = 0; i + 1 < rank; ++i) VSB_CHECK(gs[i] % 16 == 0, "TMA strides must be multiples of 16B");
  CUresult r = get_encode_fn()(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                               const_cast<void*>(base), gd, gs, bx, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swz_bytes >= 0 ? (swz_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : swz_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)
                                              : ((no_swizzle || f32) ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle_for(kblk)),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VSB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

inline FastDiv make_fastdiv(int d) {
  FastDiv f; f.d = d < 1 ? 1 : d; f.mul = 0; f.shr = 0;
  if (f.d > 1) {
    int l = 0; while ((1u << l) < (uint32_t)f.d) ++l;       // ceil(log2 d)
    const int pp = 31 + l;
    const unsigned long long m = ((1ull << pp) + (unsigned long long)f.d - 1) / (unsigned long long)f.d;
    f.mul = (uint32_t)m; f.shr = (uint32_t)(pp - 32);
  }
  return f;
}

struct ConvGemmOp {
  int loader = LD_TMA;
  ConvGemmParams p;
  CUtensorMap tmA, tmA2, tmB, tmO;
  int grid = 0, threads = 0;
  int w_samples = 1;   // > 1: weights are [w_samples][N][K], sample = row / rows_per_sample (finalize_op)
  bool pair = false;   // long-K conv on whole 128-pixel row tiles: CTA-pair kernel (conv_pair.cuh), parameters in pp, weights map in tmB
  ConvPairParams pp;
  size_t pair_smem = 0;
  int pair_grid = 0;
  size_t smem = 0;
  const char* name = "";
  ConvGemmOp() { memset(&pp, 0, sizeof(pp)); memset(&p, 0, sizeof(p)); memset(&tmA, 0, sizeof(tmA)); memset(&tmA2, 0, sizeof(tmA2)); memset(&tmB, 0, sizeof(tmB)); memset(&tmO, 0, sizeof(tmO)); }
};

inline int pick_block_n(int N, int max_bn = 256) {
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= max_bn) return n16;
  if (max_bn < 256) {
    for (int bn = max_bn / 16 * 16; bn >= 64; bn -= 16)
      if (n16 % bn == 0) return bn;
    return max_bn / 16 * 16;
  }
  // prefer the largest tile in {256,...,128} (multiples of 16) that divides n16; else 256 with a ragged last tile
  for (int bn = 256; bn >= 128; bn -= 16)
    if (n16 % bn == 0) return bn;
  return 256;
}

// Common finalisation: weights map, tiling, stage count, smem, grid.
// W: [N, Kw] fp16 row-major (K-major), Kw >= num_kb*kblk need not hold (TMA zero-fills out-of-bounds K).
inline void finalize_op(ConvGemmOp& op, const __half* W, int N, int Kw, int ldw, int num_sms, int block_n_override = 0) {
  ConvGemmParams& p = op.p;
  p.N = N;
  // an fp32 residual tile is prefetched into shared memory: keep it <= 64 KB
  p.block_n = block_n_override ? block_n_override : pick_block_n(N, p.resid32 ? 128 : 256);
  VSB_CHECK(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "bad block_n");
  p.n_tiles = (N + p.block_n - 1) / p.block_n;
  // small-N layers: several consecutive M tiles share one accumulator round (TMEM columns g*block_n) so that the per-tile
  // barrier handshakes / index math are amortised
  p.group = 1;
  if (!getenv("VSB_NO_GROUP")) {
    p.group = p.block_n <= 32 ? kAccStride / p.block_n : 1;   // measured: only pays off for N <= 32
    if (p.group > 4) p.group = 4;
    if (p.group < 1) p.group = 1;
    if (p.m_tiles < 4 * num_sms) p.group = 1;   // keep enough tiles for load balance
  }
  p.num_tiles = ((p.m_tiles + p.group - 1) / p.group) * p.n_tiles;
  p.fd_ntiles = make_fastdiv(p.n_tiles); p.fd_tpi = make_fastdiv(p.tiles_per_img); p.fd_tx = make_fastdiv(p.tiles_x);
  p.fd_tw = make_fastdiv(p.tile_w); p.fd_group = make_fastdiv(p.group); p.fd_rps = make_fastdiv(p.rows_per_sample);
  p.fd_hw = make_fastdiv(p.hw); p.fd_cc = make_fastdiv(p.cc); p.fd_ow = make_fastdiv(p.OW); p.fd_oh = make_fastdiv(p.OH);
  p.fd_ct = make_fastdiv(p.C0 + p.C1); p.fd_s = make_fastdiv(p.S);
  p.a_stage_bytes = (uint32_t)(kBlockM * p.kblk * 2);
  p.b_stage_bytes = (uint32_t)(p.block_n * p.kblk * 2);
  p.stage_bytes = p.a_stage_bytes + ((p.b_stage_bytes + 1023u) & ~1023u);
  VSB_CHECK(p.a_stage_bytes % 1024 == 0, "A stage must be 1024B aligned");
  // shared memory: [header | stages x (A|B) | 2 halo tiles | upsampled halo | resident weights | residual ring]
  size_t resid_one = 0;
  p.resid_direct = 0;
  if (p.resid16 && op.loader == LD_TMA && p.num_kb >= 24 && (p.ld_res16 % 8) == 0 && N % p.block_n == 0 && (long)p.m_tiles * kBlockM == (long)p.M &&
      p.out32 == nullptr && p.outc_w == nullptr && !getenv("VSB_RESID_RING"))
    p.resid_direct = 1;     // whole tiles only (the lean affine path is the one that implements the direct read)
  if (p.resid16 && !p.resid_direct) resid_one = (size_t)kBlockM * p.block_n * 2;
  if (p.resid32) resid_one = (size_t)kBlockM * p.block_n * 4;
  int depth = 1;
  if (resid_one) { depth = (int)(32768 / resid_one); if (depth > 4) depth = 4; if (depth < 1) depth = 1; }
  p.resid_depth = depth;
  p.resid_stride = (uint32_t)resid_one;
  const size_t resid_bytes = resid_one * depth;
  p.halo_stride = (uint32_t)(((size_t)p.halo_bytes + 1023) & ~size_t(1023));
  // halo ring: as deep as ~48 KB allows (2..8): the producer prefetches that many tiles ahead of the builders
  p.halo_bufs = 2;
  if (p.halo_stride) { p.halo_bufs = (int)(49152 / p.halo_stride); if (p.halo_bufs > kMaxHalo) p.halo_bufs = kMaxHalo; if (p.halo_bufs < 2) p.halo_bufs = 2; }
  const size_t halo_total = (size_t)p.halo_bufs * p.halo_stride;
  const size_t u_bytes = op.loader == LD_HALO_UPS ? (((size_t)kHaloH * kHaloW * p.cc * 2 + 1023) & ~size_t(1023)) : 0;
  // resident weights: one N tile, whole K slab <= 48 KB -> loaded once per CTA instead of once per tile
  const size_t bslab = (size_t)p.num_kb * p.b_stage_bytes;
  p.b_resident = (p.n_tiles == 1 && bslab <= 48 * 1024 && op.w_samples == 1) ? 1 : 0;
  // several N tiles: every CTA keeps ONE N tile for its whole life (grid a multiple of n_tiles, tile = blockIdx.x + k * gridDim.x),
  // so its [block_n x K] weight slab can stay resident too when it fits beside >= 4 A stages.  Without this the slab is re-streamed
  // from L2 for every M tile (res 1x1 384->384 @32^2: 151 MB of weight traffic per launch against 100 MB of activations: L2-bound).
  p.b_fixed_ntile = 0;
  if (!p.b_resident && p.n_tiles > 1 && p.n_tiles <= 8 && op.w_samples == 1 && op.loader == LD_TMA && !getenv("VSB_NO_BRES2") &&
      kHeaderBytes + resid_bytes + bslab + 4 * (size_t)p.a_stage_bytes + 64 * 1024 <= 225 * 1024 && p.num_tiles >= 2 * num_sms) {
    p.b_resident = 1;
    p.b_fixed_ntile = 1;
  }
  if (getenv("VSB_NO_BRES")) { p.b_resident = 0; p.b_fixed_ntile = 0; }
  const size_t bres_bytes = p.b_resident ? bslab : 0;
  if (p.b_resident) p.stage_bytes = p.a_stage_bytes;
  // TMA-store epilogue (short main loops only: there the epilogue is the critical path; long-K layers keep their shared memory for
  // pipeline stages): whole tiles, fp16 output (affine / GELU epilogues) or the in-place fp32 residual stream of pwconv2
  p.tma_store = 0; p.cpw = 0; p.ostage_bytes = 0;
  {
    const int nchunks = p.block_n / 16, epw = 4;     // 16 epilogue warps = 4 per TMEM lane quadrant
    const bool common = op.loader == LD_TMA && p.outc_w == nullptr && p.epi == EPI_AFFINE && (long)p.m_tiles * kBlockM == (long)p.M &&
                        N % p.block_n == 0 && p.num_kb <= 12 && !getenv("VSB_NO_TMA_STORE");
    if (common && p.out16 != nullptr && p.out32 == nullptr && (p.ld_out16 % 8) == 0 && (reinterpret_cast<uintptr_t>(p.out16) & 15) == 0)
      p.tma_store = 1;
    else if (common && p.out32 != nullptr && p.out16 == nullptr && p.act == ACT_NONE && p.resid32 != nullptr && p.grn_stats == nullptr &&
             (p.ld_out32 % 4) == 0 && (reinterpret_cast<uintptr_t>(p.out32) & 15) == 0 && !getenv("VSB_NO_TMA_STORE32"))
      p.tma_store = 2;
    if (p.tma_store) {
      p.cpw = (nchunks + epw - 1) / epw;
      const uint32_t cb = p.tma_store == 2 ? 64u : 32u;          // bytes per row of one 16-column chunk
      // staging boxes: all cpw chunks of a warp in one box when the split is even and the row is 32 / 64 / 96 / 128 bytes, else one box
      // per chunk.  Rows of 32 / 64 / 128 bytes use the matching TMA swizzle so that the 16-byte register stores of the 32 lanes (one
      // row each) are bank-conflict free; 96-byte rows stay linear (2-way conflicts).
      p.ost_cpb = (nchunks % epw == 0 && p.cpw * cb <= 128u) ? p.cpw : 1;
      p.ost_rowb = (uint32_t)p.ost_cpb * cb;
      p.ost_swz = p.ost_rowb == 128u ? 3 : (p.ost_rowb == 64u ? 2 : (p.ost_rowb == 32u ? 1 : 0));
      if (getenv("VSB_TMA_STORE_NOSWZ")) p.ost_swz = 0;
      p.ostage_bytes = (uint32_t)p.cpw * 32u * cb;               // per warp: multiples of 1024 bytes (swizzle atoms stay aligned)
    }
  }
  const size_t ostage_total = p.tma_store ? (size_t)16 * p.ostage_bytes + 1024 : 0;
  const size_t fixed = kHeaderBytes + resid_bytes + halo_total + u_bytes + bres_bytes + ostage_total;
  VSB_CHECK(fixed + 2 * (size_t)p.stage_bytes <= 225 * 1024, "tile does not fit in shared memory");
  int stages = (int)((225 * 1024 - fixed) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  p.halo_off = (uint32_t)(kHeaderBytes + (size_t)stages * p.stage_bytes);
  p.u_off = (uint32_t)(p.halo_off + halo_total);
  p.bres_off = (uint32_t)(p.u_off + u_bytes);
  p.resid_off = (uint32_t)(p.bres_off + bres_bytes);
  p.ostage_off = (uint32_t)((p.resid_off + resid_bytes + 1023) & ~size_t(1023));   // swizzle atoms of the staging boxes need 1024-byte alignment
  op.smem = 1024 /*align slack*/ + p.ostage_off + ostage_total;
  if (p.tma_store) {   // output map: [M rows][N cols], box = one staging box (32 rows x 16 columns), dense rows
    uint64_t odims[2] = {(uint64_t)N, (uint64_t)p.M};
    uint32_t obox[2] = {(uint32_t)(p.ost_cpb * 16), 32u};
    const int swzb = p.ost_swz == 3 ? 128 : (p.ost_swz == 2 ? 64 : (p.ost_swz == 1 ? 32 : 0));
    if (p.tma_store == 1) {
      uint64_t ostrides[1] = {(uint64_t)p.ld_out16 * 2};
      encode_map(&op.tmO, p.out16, 2, odims, ostrides, obox, 0, /*no_swizzle=*/true, false, swzb);
    } else {
      uint64_t ostrides[1] = {(uint64_t)p.ld_out32 * 4};
      encode_map(&op.tmO, p.out32, 2, odims, ostrides, obox, 0, /*no_swizzle=*/true, /*f32=*/true, swzb);
    }
  }
  // instruction descriptor (kind::f16): D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
  if (p.epi == EPI_LN || p.outc_w) VSB_CHECK(p.n_tiles == 1 && N <= 256, "LN / fused-outc epilogues need the full row in one tile");
  // weights: dims (K, N)
  if (op.w_samples > 1) {
    VSB_CHECK(op.loader == LD_TMA && !p.a_is_conv && p.rows_per_sample > 0 && p.rows_per_sample % kBlockM == 0 && N % p.block_n == 0,
              "per-sample weights: plain GEMM with whole tiles per sample only");
    p.b_sample_rows = N;
    p.fd_tps = make_fastdiv(p.rows_per_sample / kBlockM);
  }
  uint64_t dims[2] = {(uint64_t)Kw, (uint64_t)N * (uint64_t)op.w_samples};
  uint64_t strides[1] = {(uint64_t)ldw * 2};
  uint32_t box[2] = {(uint32_t)p.kblk, (uint32_t)p.block_n};
  encode_map(&op.tmB, W, 2, dims, strides, box, p.kblk);
  // whole tiles and an aligned bias vector: the epilogue warps read it from global memory and never meet at a barrier (the fused
  // 1x1 outc epilogue keeps its own barrier, so it keeps the shared copy)
  p.bias_global = (p.bias != nullptr && N % p.block_n == 0 && p.epi == EPI_AFFINE && p.outc_w == nullptr &&
                   (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 && !getenv("VSB_NO_BIAS_GLOBAL")) ? 1 : 0;
  op.grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (p.b_fixed_ntile) op.grid = (num_sms / p.n_tiles) * p.n_tiles;
  op.threads = op.loader == LD_TMA ? 640 : 512;
  VSB_CHECK((long)p.m_tiles * kBlockM < (1L << 31), "M too large for 32-bit row indices");
  // CTA-pair kernel for the long-K convs on whole row tiles (the 16 bottleneck convs of the U-Net): plain bias / ReLU / fp16 residual epilogue
  static const bool no_pair = getenv("VSB_NO_PAIR") != nullptr;
  if (!no_pair && op.loader == LD_TMA && p.a_is_conv && p.kblk == 64 && p.num_kb >= 24 && p.tile_w == p.W && p.m_tiles % 2 == 0 &&
      p.epi == EPI_AFFINE && (p.act == ACT_NONE || p.act == ACT_RELU) && p.out16 != nullptr && p.out32 == nullptr && p.resid32 == nullptr &&
      p.outc_w == nullptr && p.grn_stats == nullptr && op.w_samples == 1 && p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 &&
      (p.ld_out16 % 8) == 0 && (p.resid16 == nullptr || p.ld_res16 % 8 == 0) && num_sms >= 2) {
    int bn = 0;
    for (int cand : {256, 192, 128}) if (N % cand == 0) { bn = cand; break; }
    if (const char* e = getenv("VSB_PAIR_BN")) { const int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0 && N % v == 0) bn = v; }
    if (bn) {
      ConvPairParams& q = op.pp;
      q.M = p.M; q.N = N; q.num_kb = p.num_kb; q.c_blocks = p.c_blocks; q.S = p.S; q.pad = p.pad;
      q.tile_h = p.tile_h; q.tiles_per_img = p.tiles_per_img;
      q.m_pairs = p.m_tiles / 2; q.n_tiles = N / bn; q.num_work = q.m_pairs * q.n_tiles;
      q.block_n = bn; q.a_bytes = 128 * 64 * 2; q.b_bytes = (uint32_t)(bn / 2) * 64 * 2;
      q.stages = (int)std::min<size_t>(kPairMaxStages, (size_t)(200 * 1024) / (q.a_bytes + q.b_bytes));
      if (const char* e = getenv("VSB_PAIR_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= q.stages) q.stages = v; }
      q.idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      q.relu = p.act == ACT_RELU ? 1 : 0;
      q.bias = p.bias; q.resid = p.resid16; q.ld_res = p.ld_res16; q.out = p.out16; q.ld_out = p.ld_out16;
      uint32_t pbox[2] = {64u, (uint32_t)(bn / 2)};
      encode_map(&op.tmB, W, 2, dims, strides, pbox, 64);      // each CTA of the pair stages half of the tile's weight rows
      op.pair_smem = (size_t)q.stages * (q.a_bytes + q.b_bytes) + (2 * kPairMaxStages + 4) * sizeof(uint64_t) + 16 + 1024;
      op.pair_grid = 2 * std::min(num_sms / 2, q.num_work);
      op.pair = true;
    }
  }
}

// A = NHWC fp16 activation [B, H, W, C] (pixel pitch ld elements); conv RxS stride 1, zero padding `pad`
// (R = S = 1, pad = 0 for 1x1).  Weights [N][R*S*C] with K order (r, s, c).
inline void setup_tma_conv(ConvGemmOp& op, const __half* src, int B, int H, int W, int C, int ld, int R, int S, int pad) {
  ConvGemmParams& p = op.p;
  op.loader = LD_TMA;
  p.a_is_conv = 1;
  p.tile_mode = 1;
  p.kblk = C >= 64 ? 64 : C;
  VSB_CHECK(p.kblk == 64 || p.kblk == 32 || p.kblk == 16, "TMA conv: C must be 16, 32 or a multiple of 64");
  VSB_CHECK(C % p.kblk == 0, "TMA conv: C must be a multiple of the K block");
  p.c_blocks = C / p.kblk;
  p.R = R; p.S = S; p.pad = pad;
  p.H = H; p.W = W;
  p.tile_w = W < kBlockM ? W : kBlockM;
  p.tile_h = kBlockM / p.tile_w;
  VSB_CHECK(p.tile_w * p.tile_h == kBlockM && W % p.tile_w == 0 && H % p.tile_h == 0, "TMA conv: map size must tile by 128 pixels");
  p.tiles_x = W / p.tile_w;
  p.tiles_per_img = p.tiles_x * (H / p.tile_h);
  p.m_tiles = B * p.tiles_per_img;
  p.M = B * H * W;
  p.num_kb = R * S * p.c_blocks;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
  uint32_t box[4] = {(uint32_t)p.kblk, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1u};
  encode_map(&op.tmA, src, 4, dims, strides, box, p.kblk);
}


// 3x3 stride-1 zero-padded conv with on-chip im2col: the (8+2)x(16+2) input halo of each 8x16 output tile is loaded once
// per channel chunk.  Weights [N][9*C] with K order (r, s, c).
inline int halo_kpad(int C0, int C1) {   // padded K of the chunk-major halo weight layout
  const int cc = C0 < 64 ? C0 : 64;
  const int kb_per_c = (9 * cc + 63) / 64;
  return ((C0 + C1) / cc) * kb_per_c * 64;
}

// Halo-loader weight layout: [N][chunk][tap][cc] with each chunk zero-padded to kb_per_c*64 (see ConvGemmParams).
inline void setup_halo_common(ConvGemmParams& p, int C0, int C1, int B, int OH, int OW) {
  p.tile_mode = 1;
  p.kblk = 64;
  p.cc = C0 < 64 ? C0 : 64;
  VSB_CHECK(p.cc == 64 || p.cc == 32 || p.cc == 16, "halo loader: C must be 16, 32 or a multiple of 64");
  VSB_CHECK(C0 % p.cc == 0 && C1 % p.cc == 0, "halo loader: channels must be multiples of the chunk size");
  VSB_CHECK(OW % kHaloTW == 0 && OH % kHaloTH == 0, "halo loader: output map must tile by 8x16");
  p.c0_blocks = C0 / p.cc;
  p.c_blocks = (C0 + C1) / p.cc;
  p.kb_per_c = (9 * p.cc + 63) / 64;
  p.R = 3; p.S = 3; p.pad = 1;
  p.H = OH; p.W = OW; p.tile_w = kHaloTW; p.tile_h = kHaloTH;
  p.tiles_x = OW / kHaloTW;
  p.tiles_per_img = p.tiles_x * (OH / kHaloTH);
  p.m_tiles = B * p.tiles_per_img;
  p.M = B * OH * OW;
  p.num_kb = p.c_blocks * p.kb_per_c;
}

// 3x3 stride-1 zero-padded conv with on-chip im2col: the (8+2)x(16+2) input halo of each 8x16 output tile is loaded once
// per channel chunk.
inline void setup_halo_conv3(ConvGemmOp& op, const __half* src, int B, int H, int W, int C, int ld) {
  ConvGemmParams& p = op.p;
  op.loader = LD_HALO_CONV3;
  setup_halo_common(p, C, 0, B, H, W);
  p.halo_bytes = (uint32_t)(kHaloH * kHaloW * p.cc * 2);
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
  uint32_t box[4] = {(uint32_t)p.cc, (uint32_t)kHaloW, (uint32_t)kHaloH, 1u};
  encode_map(&op.tmA, src, 4, dims, strides, box, 64, /*no_swizzle=*/true);
}

// UBlock up-conv (modules/common.py:45-52 after the skip concat of unet.py:187-190): conv3x3(valid) o ReflectionPad2d(1) o
// bilinear x2 of the virtual concat [src0 (C0) | src1 (C1)], both NHWC fp16 at IH x IW; output at 2IH x 2IW.
inline void setup_halo_ups(ConvGemmOp& op, const __half* src0, int C0, int ld0, const __half* src1, int C1, int ld1, int B, int IH, int IW) {
  ConvGemmParams& p = op.p;
  op.loader = LD_HALO_UPS;
  VSB_CHECK(C1 == 0 || C1 == C0, "halo ups: both sources must have the same channel count");
  setup_halo_common(p, C0, C1, B, 2 * IH, 2 * IW);
  p.IH = IH; p.IW = IW; p.OH = 2 * IH; p.OW = 2 * IW;
  p.halo_bytes = (uint32_t)(6 * 10 * p.cc * 2);
  uint32_t box[4] = {(uint32_t)p.cc, 10u, 6u, 1u};
  {
    uint64_t dims[4] = {(uint64_t)C0, (uint64_t)IW, (uint64_t)IH, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)ld0 * 2, (uint64_t)IW * ld0 * 2, (uint64_t)IH * IW * ld0 * 2};
    encode_map(&op.tmA, src0, 4, dims, strides, box, 64, true);
  }
  if (C1 > 0) {
    uint64_t dims[4] = {(uint64_t)C1, (uint64_t)IW, (uint64_t)IH, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)ld1 * 2, (uint64_t)IW * ld1 * 2, (uint64_t)IH * IW * ld1 * 2};
    encode_map(&op.tmA2, src1, 4, dims, strides, box, 64, true);
  }
}

// A = [M, K] fp16 row-major (row pitch ld elements)
inline void setup_tma_gemm(ConvGemmOp& op, const __half* A, long M, int K, int ld) {
  ConvGemmParams& p = op.p;
  op.loader = LD_TMA;
  p.a_is_conv = 0;
  p.kblk = K >= 64 ? 64 : K;
  VSB_CHECK(p.kblk == 64 || p.kblk == 32 || p.kblk == 16, "TMA gemm: K must be 16, 32 or >= 64");
  p.M = (int)M;
  p.m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  p.num_kb = (K + p.kblk - 1) / p.kblk;
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t strides[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {(uint32_t)p.kblk, (uint32_t)kBlockM};
  encode_map(&op.tmA, A, 2, dims, strides, box, p.kblk);
}

// A = [A0 | A1] virtual concat along K of two [M, K0] / [M, K1] fp16 matrices (row pitches ld0 / ld1)
inline void setup_tma_gemm2(ConvGemmOp& op, const __half* A0, int K0, int ld0, const __half* A1, int K1, int ld1, long M) {
  ConvGemmParams& p = op.p;
  op.loader = LD_TMA;
  p.a_is_conv = 0;
  p.kblk = K0 >= 64 ? 64 : K0;
  VSB_CHECK(p.kblk == 64 || p.kblk == 32 || p.kblk == 16, "TMA gemm2: K0 must be 16, 32 or a multiple of 64");
  VSB_CHECK(K0 % p.kblk == 0 && K1 % p.kblk == 0, "TMA gemm2: K0, K1 must be multiples of the K block");
  p.M = (int)M;
  p.m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  p.c0_blocks = K0 / p.kblk;
  p.num_kb = (K0 + K1) / p.kblk;
  uint32_t box[2] = {(uint32_t)p.kblk, (uint32_t)kBlockM};
  {
    uint64_t dims[2] = {(uint64_t)K0, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ld0 * 2};
    encode_map(&op.tmA, A0, 2, dims, strides, box, p.kblk);
  }
  {
    uint64_t dims[2] = {(uint64_t)K1, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ld1 * 2};
    encode_map(&op.tmA2, A1, 2, dims, strides, box, p.kblk);
  }
}

// generic gathered conv: out[b,oy,ox,:] over taps (r,s) of the virtual concat [src0 (C0 ch) | src1 (C1 ch)]
inline void setup_gather_conv(ConvGemmOp& op, int loader, const __half* src0, int C0, int ld0, const __half* src1, int C1, int ld1,
                              int B, int IH, int IW, int OH, int OW, int R, int S, int stride, int pad, int pad_mode) {
  ConvGemmParams& p = op.p;
  op.loader = loader;
  p.kblk = 64;
  p.src0 = src0; p.src1 = src1; p.C0 = C0; p.C1 = C1; p.ld0 = ld0; p.ld1 = ld1;
  VSB_CHECK(C0 % 8 == 0 && C1 % 8 == 0 && ld0 % 8 == 0 && (C1 == 0 || ld1 % 8 == 0), "gather conv: channels must be multiples of 8");
  p.IH = IH; p.IW = IW; p.OH = OH; p.OW = OW; p.R = R; p.S = S; p.stride = stride; p.pad = pad; p.pad_mode = pad_mode;
  p.Ktot = R * S * (C0 + C1);
  p.M = B * OH * OW;
  p.m_tiles = (p.M + kBlockM - 1) / kBlockM;
  p.num_kb = (p.Ktot + 63) / 64;
}

// A[m, k] = G[m, k] * scale[m / rows_per_sample, k]   (ConvNeXt pwconv2 with the GRN factor folded in)
inline void setup_gather_scale(ConvGemmOp& op, const __half* G, long M, int K, int ld, const float* scale, int ld_scale, int rows_per_sample) {
  ConvGemmParams& p = op.p;
  op.loader = LD_GATHER_SCALE;
  p.kblk = 64;
  p.src0 = G; p.ld0 = ld; p.C0 = K; p.C1 = 0;
  VSB_CHECK(K % 8 == 0 && ld % 8 == 0 && ld_scale % 4 == 0, "gather scale: K must be a multiple of 8");
  p.Ktot = K;
  p.a_scale = scale; p.ld_scale = ld_scale; p.rows_per_sample = rows_per_sample;
  p.M = (int)M;
  p.m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  p.num_kb = (K + 63) / 64;
}

#ifdef VSB_PDL
// launch with programmatic stream serialization: the kernel may begin while its predecessor drains and synchronises
// itself with pdl_wait() (ptx.cuh).  VSB_NO_PDL=1 falls back to classic launches at run time (A/B in one build).
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool off = getenv("VSB_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = off ? 0 : 1;
  VSB_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}
#endif

template <int LOADER, int ACT>
inline void launch_one(const ConvGemmOp& op, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    VSB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<LOADER, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
#ifdef VSB_PDL
  launch_pdl(conv_gemm_kernel<LOADER, ACT>, dim3(op.grid), dim3(op.threads), op.smem, st, op.tmA, op.tmA2, op.tmB, op.tmO, op.p);
#else
  conv_gemm_kernel<LOADER, ACT><<<op.grid, op.threads, op.smem, st>>>(op.tmA, op.tmA2, op.tmB, op.tmO, op.p);
#endif
}

// instantiated (loader, activation) pairs: only what the networks need, to bound compile time
inline void launch_pair(const ConvGemmOp& op, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    VSB_CUDA(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
#ifdef VSB_PDL
  launch_pdl(conv_pair_kernel, dim3(op.pair_grid), dim3(kPairThreads), op.pair_smem, st, op.tmA, op.tmB, op.pp);
#else
  conv_pair_kernel<<<op.pair_grid, kPairThreads, op.pair_smem, st>>>(op.tmA, op.tmB, op.pp);
#endif
}

inline void launch(const ConvGemmOp& op, cudaStream_t st) {
  if (op.pair) { launch_pair(op, st); VSB_CUDA(cudaGetLastError()); return; }
  const int a = op.p.act;
  switch (op.loader) {
    case LD_TMA:
      if (a == ACT_NONE) launch_one<LD_TMA, ACT_NONE>(op, st);
      else if (a == ACT_RELU) launch_one<LD_TMA, ACT_RELU>(op, st);
      else launch_one<LD_TMA, ACT_GELU>(op, st);
      break;
    case LD_GATHER_CONV:
      if (a == ACT_NONE) launch_one<LD_GATHER_CONV, ACT_NONE>(op, st);
      else if (a == ACT_RELU) launch_one<LD_GATHER_CONV, ACT_RELU>(op, st);
      else throw Error("gather conv: unsupported activation", kErrUnsupported);
      break;
    case LD_HALO_UPS:
      if (a == ACT_RELU) launch_one<LD_HALO_UPS, ACT_RELU>(op, st);
      else throw Error("halo ups: unsupported activation", kErrUnsupported);
      break;
    case LD_HALO_CONV3:
      if (a == ACT_RELU) launch_one<LD_HALO_CONV3, ACT_RELU>(op, st);
      else throw Error("halo conv3: unsupported activation", kErrUnsupported);
      break;
    case LD_GATHER_SCALE:
      if (a == ACT_NONE) launch_one<LD_GATHER_SCALE, ACT_NONE>(op, st);
      else throw Error("gather scale: unsupported activation", kErrUnsupported);
      break;
    default: throw Error("bad loader");
  }
  VSB_CUDA(cudaGetLastError());
}

}  // namespace vsb
