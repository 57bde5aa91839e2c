// Host side of conv3_direct.cuh: geometry, tensor map, shared-memory plan, launch.
#pragma once
#include "conv3_direct.cuh"
#include "conv_gemm_host.cuh"

namespace vsb {

struct Conv3DirectOp {
  Conv3DirectParams p;
  int grid = 0;
  size_t smem = 0;
  Conv3DirectOp() { memset(&p, 0, sizeof(p)); }
};

// ring rows that fit next to the resident weights for a W-wide map (0 = channel counts not supported)
inline int conv3_direct_ring_rows(int C, int N, int W, int R = 3) {
  if (!((C == 16 || C == 32 || C == 64) && (N == 16 || N == 32 || N == 64))) return 0;
  const int rw = (W + R - 1 + 7) / 8 * 8;
  const uint32_t w_pad = ((uint32_t)((size_t)N * R * R * C * 2) + 1023u) & ~1023u;
  const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - kD3HeaderBytes - w_pad;
  int rows = (int)(budget / ((size_t)C * 2 * rw)) - (128 / rw + 1);
  rows = std::min(rows, kD3MaxRing);
  rows = std::min(rows, (kD3TileBars - 16) * 128 / rw);   // tiles in flight must stay below the tile_done barrier ring
  return rows;
}
// Fewest ring rows that cannot deadlock: a tile touches ceil(129/rw) + R rows at once, the producer refills a slot only after
// the tiles that read its previous row are done, and the issuers run in order; two more rows keep loads ahead of the MMAs.
inline int conv3_direct_min_rows(int W, int R = 3) {
  const int rw = (W + R - 1 + 7) / 8 * 8;
  return (129 + rw - 1) / rw + R + 2;
}
// W = 0: channel-count check only (weight packing time)
inline bool conv3_direct_ok(int C, int N, int ld_in, int W = 0, int R = 3) {
  static const bool off = getenv("VSB_NO_DIRECT") != nullptr;
  static const bool off1 = getenv("VSB_NO_DIRECT1") != nullptr;
  if (off || (R == 1 && off1) || ld_in != C) return false;
  if (W == 0) return conv3_direct_ring_rows(C, N, 64, R) > 0;
  return W % 8 == 0 && conv3_direct_ring_rows(C, N, W, R) >= conv3_direct_min_rows(W, R);
}

// x: dense NHWC fp16 [B,H,W,C]; wpk: weights in core-matrix layout (pack_direct_weights_kernel), N*9*C halves
inline void setup_conv3_direct(Conv3DirectOp& op, const __half* x, int B, int H, int W, int C, int N, const __half* wpk, int num_sms,
                               int R = 3) {
  Conv3DirectParams& p = op.p;
  VSB_CHECK(R == 3 || R == 1, "direct conv: 3x3 or 1x1 only");
  VSB_CHECK(conv3_direct_ring_rows(C, N, W, R) > 0, "direct conv3: unsupported channel counts");
  p.B = B; p.H = H; p.W = W; p.C = C; p.N = N; p.R = R;
  p.relu = 1;
  p.rw = (W + R - 1 + 7) / 8 * 8;
  p.hp = H + R - 1;
  const long positions = (long)B * p.hp * p.rw;
  VSB_CHECK(positions + 3L * p.rw < (1L << 31) - 1024, "direct conv3: batch too large for 32-bit positions");
  p.n_tiles = (int)((positions + 127) / 128);
  p.mirror_rows = 128 / p.rw + 1;
  p.w_bytes = (uint32_t)((size_t)N * R * R * C * 2);
  const uint32_t w_pad = (p.w_bytes + 1023u) & ~1023u;
  const int rows = conv3_direct_ring_rows(C, N, W, R);
  VSB_CHECK(rows >= conv3_direct_min_rows(W, R), "direct conv3: shared-memory ring too small");
  p.ring_rows = rows;
  p.plane_stride = (uint32_t)((size_t)(rows + p.mirror_rows) * p.rw * 16);
  p.w_off = kD3HeaderBytes;
  p.ring_off = kD3HeaderBytes + w_pad;
  op.smem = 1024 + (size_t)p.ring_off + (size_t)(C / 8) * p.plane_stride;
  p.acc_stages = kD3MaxAcc;      // 8 stages x N <= 64 columns
  p.tmem_cols = kD3MaxAcc * N;   // 128 / 256 / 512: powers of two
  p.idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.fd_rw = make_fastdiv(p.rw);
  p.fd_hp = make_fastdiv(p.hp);
  p.wpk = wpk;
  static const bool swap = getenv("VSB_DIRECT_SWAP") != nullptr;
  p.swap_lbo_sbo = swap ? 1 : 0;
  p.x = x;
  VSB_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0, "direct conv3: input must be 16-byte aligned");
  op.grid = std::min(num_sms, p.n_tiles);
}

template <int N, int KS, int R, int MODE = 0>
inline void launch_direct_nkr(const Conv3DirectOp& op, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    VSB_CUDA(cudaFuncSetAttribute(conv3_direct_kernel<N, KS, R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
#ifdef VSB_PDL
  launch_pdl(conv3_direct_kernel<N, KS, R, MODE>, dim3(op.grid), dim3(kD3Threads), op.smem, st, op.p);
#else
  conv3_direct_kernel<N, KS, R, MODE><<<op.grid, kD3Threads, op.smem, st>>>(op.p);
#endif
  VSB_CUDA(cudaGetLastError());
}
template <int N, int KS>
inline void launch_direct_nk(const Conv3DirectOp& op, cudaStream_t st) {
  if (op.p.R == 3) launch_direct_nkr<N, KS, 3>(op, st);
  else launch_direct_nkr<N, KS, 1>(op, st);
}
template <int N>
inline void launch_direct_n(const Conv3DirectOp& op, cudaStream_t st) {
  if (op.p.C == 16) launch_direct_nk<N, 1>(op, st);
  else if (op.p.C == 32) launch_direct_nk<N, 2>(op, st);
  else launch_direct_nk<N, 4>(op, st);
}

inline void launch_direct(const Conv3DirectOp& op, cudaStream_t st) {
  VSB_CHECK(op.smem <= 227 * 1024, "direct conv3: shared-memory plan too large");
  if (op.p.x2 != nullptr) {   // up-conv phase mode: 64 input channels (two sources), 4 phases x 16 output channels
    VSB_CHECK(op.p.N == 64 && op.p.C == 64 && op.p.R == 3, "up-phase direct conv: C = 64, N = 4 x 16 only");
    launch_direct_nkr<64, 4, 3, 1>(op, st);
    return;
  }
  if (op.p.s2d) {             // stride-2 conv through the space-to-depth view: C = 4 x 16 input channels, 32 outputs
    VSB_CHECK(op.p.N == 32 && op.p.C == 64 && op.p.R == 3, "stride-2 direct conv: 16 -> 32 channels only");
    launch_direct_nkr<32, 4, 3, 2>(op, st);
    return;
  }
  if (op.p.N == 16) launch_direct_n<16>(op, st);
  else if (op.p.N == 32) launch_direct_n<32>(op, st);
  else launch_direct_n<64>(op, st);
}

// UBlock up-conv (bilinear x2 -> reflect pad -> conv3x3 -> LN -> ReLU, modules/common.py:45-52 + unet.py:187-190) as ONE 3x3 conv
// on the LOW-resolution grid with 4 x 16 output columns (one group per output phase (oy&1, ox&1)) and replicate padding:
// away from the image border the bilinear weights depend only on the phase, so they fold into the conv weights
// (Model::pack_upconv_phase).  The outermost output rows/columns (reflect padding of the up-sampled map) do not follow the
// phase pattern and are recomputed exactly by up_border_fix_kernel afterwards.
inline void setup_up_phase_direct(Conv3DirectOp& op, const __half* x0, int C0, const __half* x1, int C1, int B, int H, int W,
                                  const __half* wpk, const float* ln_w, const float* ln_b, float eps, __half* out, int num_sms) {
  VSB_CHECK(C0 % 8 == 0 && C1 % 8 == 0 && C0 + C1 == 64, "up-phase direct conv: 64 input channels in two sources");
  setup_conv3_direct(op, x0, B, H, W, 64, 64, wpk, num_sms, 3);
  op.p.x2 = x1; op.p.planes0 = C0 / 8; op.p.replicate = 1;
  op.p.ln_w = ln_w; op.p.ln_b = ln_b; op.p.ln_eps = eps;
  op.p.out = out; op.p.relu = 0;
  VSB_CHECK((reinterpret_cast<uintptr_t>(x1) & 15) == 0, "direct conv3: input must be 16-byte aligned");
}

// DBlock.down (unet.py:75: 3x3, stride 2, pad 1, bias) on x [B,2H,2W,Cin] -> [B,H,W,N]: output (Y,X), tap r reads input row
// 2Y + r - 1 = row (Y-1, dy=1), (Y, dy=0), (Y, dy=1) of the space-to-depth view, i.e. a conv with offsets {-1, 0} only.
// wpk: weights re-indexed to [N][(ey,ex)][(dy,dx,c)] (Model::pack_down_s2d) in the direct layout.
inline void setup_down_s2d_direct(Conv3DirectOp& op, const __half* x, int Cin, int B, int H, int W, int N, const __half* wpk,
                                  const float* bias, __half* out, int num_sms) {
  setup_conv3_direct(op, x, B, H, W, 4 * Cin, N, wpk, num_sms, 3);
  op.p.s2d = 1; op.p.bias = bias; op.p.out = out; op.p.relu = 0;
}

// [N][9*C] fp16 on the device -> a new buffer in the direct layout
inline void pack_direct_weights(const __half* w, int N, int C, __half* out, cudaStream_t st, int T = 9) {
  pack_direct_weights_kernel<<<32, 256, 0, st>>>(w, N, C, T, out);
  VSB_CUDA(cudaGetLastError());
}

}  // namespace vsb
