// Model assembly: weight folding/packing (host), per-batch execution plans (device buffers + launch lists),
// and the embed / detect drivers behind the C ABI.
#pragma once
#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vsb200.h"
#include "conv_gemm_host.cuh"
#include "conv3_direct_host.cuh"
#include "pointwise.cuh"
#include "resize_blend.cuh"
#include "dwconv_ring.cuh"

namespace vsb {

static int64_t g_launches = 0;

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

// device allocations owned by one object, freed together
struct DevicePool {
  std::vector<void*> ptrs;
  size_t bytes = 0;
  void* alloc(size_t n) {
    void* p = nullptr;
    n = (n + 255) & ~size_t(255);
    if (n == 0) n = 256;
    VSB_CUDA(cudaMalloc(&p, n));
    ptrs.push_back(p);
    bytes += n;
    return p;
  }
  template <class T> T* alloc_n(size_t n) { return reinterpret_cast<T*>(alloc(n * sizeof(T))); }
  template <class T> T* upload(const std::vector<T>& v) {
    T* p = alloc_n<T>(v.size());
    VSB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
  }
  ~DevicePool() { for (void* p : ptrs) cudaFree(p); }
};

struct ConvW {          // packed conv / linear weights: fp16 [N][K] (K order r,s,c) + fp32 bias
  __half* w = nullptr;
  __half* wd = nullptr;  // 3x3 convs eligible for conv3_direct.cuh: the same weights in its core-matrix layout
  float* bias = nullptr;
  int N = 0, K = 0;
  int ld = 0;            // row pitch of `w` in elements (0 = K): rows are padded to a multiple of 8 (16-byte TMA strides)
};

struct DebugTensor { const void* ptr; int dtype; /*0 f32, 1 f16*/ int64_t shape[4]; int64_t ld; };

struct Step {
  std::function<void(cudaStream_t)> fn;
  int launches = 1;
  std::string name = "";
};

// optional per-step CUDA-event profile (bench.py roofline leg): name -> (total ms, launches)
struct Profile {
  bool enabled = false;
  std::map<std::string, std::pair<double, long>> acc;
  std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
  void flush() {
    for (auto& e : pending) {
      float ms = 0.f;
      cudaEventSynchronize(e.second.second);
      cudaEventElapsedTime(&ms, e.second.first, e.second.second);
      auto& a = acc[e.first];
      a.first += ms; a.second += 1;
      cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second);
    }
    pending.clear();
  }
};
static Profile g_profile;

struct Plan {
  DevicePool pool;
  uint64_t last_use = 0;    // LRU stamp of the plan cache
  std::vector<Step> steps;
  std::map<std::string, DebugTensor> dbg;
  // embed plan I/O
  const float* in_imgs = nullptr;  // set per call (pointer slot read by the first kernel's lambda)
  long in_frame_stride = 0;
  const uint8_t* in_msgs = nullptr;
  int in_msg_stride = 0;
  float* x_res = nullptr;   // [B,3,S,S] resized RGB when the input is not at processing size
  float* delta = nullptr;   // [B,CD,S,S]
  float* logits = nullptr;  // [B,1+K]
  void run(cudaStream_t st) {
    if (!g_profile.enabled) {
      for (auto& s : steps) { s.fn(st); g_launches += s.launches; }
      return;
    }
    for (auto& s : steps) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a, st);
      s.fn(st);
      cudaEventRecord(b, st);
      g_launches += s.launches;
      g_profile.pending.push_back({s.name, {a, b}});
    }
  }
};

// ATen-compatible separable resample tables (host) ------------------------------------------------
struct ResampleHost {
  std::vector<int> start, cnt;
  std::vector<float> w;
  int maxt = 0;
};

// aten/src/ATen/native/cpu/UpSampleKernel.cpp (_compute_indices_weights_aa, bilinear filter) for antialias=True;
// plain half-pixel bilinear (align_corners=False) otherwise.  Upscaling with antialias=True degenerates to plain bilinear.
inline ResampleHost make_resample(int in, int out, bool antialias) {
  ResampleHost t;
  const double scale = (double)in / (double)out;
  if (antialias && scale > 1.0) {
    const double support = scale;  // interp_size/2 * scale with interp_size = 2
    t.maxt = (int)std::ceil(support) * 2 + 1;
    t.start.resize(out); t.cnt.resize(out); t.w.assign((size_t)out * t.maxt, 0.f);
    for (int i = 0; i < out; ++i) {
      const double center = scale * (i + 0.5);
      const double invscale = 1.0 / scale;
      int xmin = (int)(center - support + 0.5); if (xmin < 0) xmin = 0;
      int xmax = (int)(center + support + 0.5); if (xmax > in) xmax = in;
      const int n = xmax - xmin;
      double total = 0.0;
      std::vector<double> ww(n);
      for (int j = 0; j < n; ++j) {
        double x = (j + xmin - center + 0.5) * invscale;
        if (x < 0) x = -x;
        ww[j] = x < 1.0 ? 1.0 - x : 0.0;
        total += ww[j];
      }
      t.start[i] = xmin; t.cnt[i] = n;
      for (int j = 0; j < n; ++j) t.w[(size_t)i * t.maxt + j] = (float)(total != 0.0 ? ww[j] / total : 0.0);
    }
  } else {
    t.maxt = 2;
    t.start.resize(out); t.cnt.resize(out); t.w.assign((size_t)out * 2, 0.f);
    for (int i = 0; i < out; ++i) {
      // area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
      // single rounding (fused multiply-add), like the vectorised ATen CPU kernel and the CUDA one
      float src = std::fmaf((float)scale, i + 0.5f, -0.5f);
      if (src < 0.f) src = 0.f;
      int x0 = (int)src; if (x0 > in - 1) x0 = in - 1;
      const int x1 = x0 + ((x0 < in - 1) ? 1 : 0);
      const float l1 = src - (float)x0, l0 = 1.f - l1;
      t.start[i] = x0;
      if (x1 == x0) { t.cnt[i] = 1; t.w[(size_t)i * 2] = 1.f; }
      else { t.cnt[i] = 2; t.w[(size_t)i * 2] = l0; t.w[(size_t)i * 2 + 1] = l1; }
    }
  }
  return t;
}

struct ResampleDev {
  DevicePool pool;
  ResampleTab tab;
  // geometry of resize_sep_kernel for this (in -> out) pair: output rows per block, input rows per staging round, the largest
  // number of input rows one block touches, dynamic shared memory
  int toy = 8, gstage = 1, rin_max = 1;
  size_t smem = 0;
  ResampleDev(int ih, int iw, int oh, int ow, bool aa) {
    ResampleHost y = make_resample(ih, oh, aa), x = make_resample(iw, ow, aa);
    tab.ystart = pool.upload(y.start); tab.ycnt = pool.upload(y.cnt); tab.yw = pool.upload(y.w); tab.maxt_y = y.maxt;
    tab.xstart = pool.upload(x.start); tab.xcnt = pool.upload(x.cnt); tab.xw = pool.upload(x.w); tab.maxt_x = x.maxt;
    gstage = iw * 4 * 4 <= 16 * 1024 ? 4 : (iw * 4 * 2 <= 16 * 1024 ? 2 : 1);    // rows per staging round (kRsStages - 1 rounds in flight): 4, 2 or 1
    for (toy = 8; toy >= 1; toy >>= 1) {
      rin_max = 1;
      for (int o0 = 0; o0 < oh; o0 += toy) {
        const int o1 = std::min(oh, o0 + toy) - 1;
        rin_max = std::max(rin_max, y.start[o1] + y.cnt[o1] - y.start[o0]);
      }
      smem = ((((size_t)rin_max * ow + 3) & ~(size_t)3) + kRsStages * (((size_t)gstage * iw + 8 + 3) & ~(size_t)3)) * sizeof(float);
      if (smem <= 100 * 1024 || toy == 1) break;
    }
    if (smem > 200 * 1024) throw Error("resize: image too wide for the shared-memory resample kernel (not implemented)", kErrUnsupported);
  }
};

// K4 v3 launch (dwconv_ring.cuh): tensor map of the fp32 NHWC residual stream + geometry, built once per plan step
struct DwRingOp {
  CUtensorMap tm;
  int C = 0, B = 0, H = 0, R = 0;
  const float *dww = nullptr, *dwb = nullptr, *lnw = nullptr, *lnb = nullptr;
  __half* out = nullptr;
};
template <int C, int NS, int SX, int NSLOT, bool WS = false>
inline void launch_dw_ring_t(const DwRingOp& op, cudaStream_t st) {
  static bool attr = false;
  constexpr size_t smem = dw_ring_smem<C, NS, SX, NSLOT, WS>();
  if (!attr) {
    VSB_CUDA(cudaFuncSetAttribute(dwconv7_ln_ring_kernel<C, NS, SX, NSLOT, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const unsigned blocks = (unsigned)((long)op.B * (op.H / op.R) * (op.H / (NS * SX)));
#ifdef VSB_PDL
  launch_pdl(dwconv7_ln_ring_kernel<C, NS, SX, NSLOT, WS>, dim3(blocks), dim3(NS * C / 2 + 32), smem, st, op.tm, op.B, op.H, op.H, op.dww, op.dwb, op.lnw,
             op.lnb, op.out, op.R);
#else
  dwconv7_ln_ring_kernel<C, NS, SX, NSLOT, WS><<<blocks, NS * C / 2 + 32, smem, st>>>(op.tm, op.B, op.H, op.H, op.dww, op.dwb, op.lnw, op.lnb, op.out, op.R);
#endif
  VSB_CUDA(cudaGetLastError());
}
inline bool dw_ring_ok(int C, int H, int ld) {
  // (the 384-channel instantiation exists and is correct, but at 16 x 16 maps it measured 50 us against 33 us for the whole-row strip
  //  kernel: one 192-thread block per SM with a 14-step prologue per 8 rows; profiles/r2_history.md)
  static const bool all = getenv("VSB_DW_RING_384") != nullptr;
  return ld == C && ((C == 96 && H % 16 == 0) || (C == 192 && H % 16 == 0) || (all && C == 384 && H % 8 == 0));
}
inline void setup_dw_ring(DwRingOp& op, const float* x, int B, int H, int C) {
  op.C = C; op.B = B; op.H = H;
  op.R = C == 384 ? 8 : (H % 32 == 0 ? 32 : 16);    // rows per block: 6 halo rows + the block prologue are amortised over R
  const int BW = (C == 96 ? 8 : 4) + 6, CB = C <= 256 ? C : 192;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)H, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)H * C * 4, (uint64_t)H * H * C * 4};
  uint32_t box[4] = {(uint32_t)CB, (uint32_t)BW, 1u, 1u};
  encode_map(&op.tm, x, 4, dims, strides, box, 0, true, /*f32=*/true);
}
inline void launch_dw_ring(const DwRingOp& op, cudaStream_t st) {
  static const bool ws = getenv("VSB_DW_WS") != nullptr;     // weights from shared memory (more blocks per SM): A/B switch
  if (op.C == 96) { if (ws) launch_dw_ring_t<96, 2, 4, 6, true>(op, st); else launch_dw_ring_t<96, 2, 4, 8>(op, st); }
  else if (op.C == 192) { if (ws) launch_dw_ring_t<192, 1, 4, 4, true>(op, st); else launch_dw_ring_t<192, 1, 4, 8>(op, st); }
  else launch_dw_ring_t<384, 1, 4, 8>(op, st);
}

class Model {
 public:
  vsb_model_desc d;
  std::map<std::string, HostTensor> sd;
  bool finalized = false;
  int device = -1, num_sms = 0;
  std::recursive_mutex mu;   // serialises the C entry points of one handle (vsb200.cu VSB_MODEL_SCOPE)
  DevicePool wpool;
  static constexpr int kMaxBatch = 64;

  // ---- packed weights
  struct ResBlockW { ConvW c1, c2, res; };
  float *first_w1 = nullptr, *first_b1 = nullptr, *first_wr = nullptr, *first_br = nullptr;  // inc (CUDA-core layer)
  ConvW inc_c2;
  std::vector<ConvW> down_conv;
  std::vector<ResBlockW> down_rb, bott_rb, up_rb;
  std::vector<ConvW> up_conv;
  std::vector<float*> up_lnw, up_lnb;
  std::vector<__half*> up_phase_w;   // per level: phase-folded weights (direct layout) or nullptr
  std::vector<float*> up_border_w;   // per level: fp32 taps [9][Cout][Ct] for the border fix-up, or nullptr
  float *outc_w = nullptr, *outc_b = nullptr;
  float* msg_table = nullptr;
  // extractor
  __half* stem_w16 = nullptr;
  float* stem_wk = nullptr;   // widths > 256 (chunkyseal): fp32 [48][C0] for the CUDA-core stem kernel
  float *stem_b = nullptr, *stem_lnw = nullptr, *stem_lnb = nullptr;
  struct DsW { float* lnw; float* lnb; ConvW conv; };
  std::vector<DsW> ds;
  struct CnBlockW { float* dww; float* dwb; float* lnw; float* lnb; ConvW pw1; float* gamma; ConvW pw2; };
  std::vector<std::vector<CnBlockW>> cn;
  ConvW head_conv;
  float *head_lnw = nullptr, *head_lnb = nullptr, *head_lw = nullptr, *head_lb = nullptr;

  // grow-only device staging for the host-buffer entry points (no cudaMalloc on the steady-state path)
  struct Staging { void* p = nullptr; size_t cap = 0; };
  Staging staging[8];   // 0-5: host entry points; 6-7: low-resolution attenuation scratch
  void* stage(int slot, size_t bytes) {
    Staging& s = staging[slot];
    if (bytes > s.cap) {
      if (s.p) cudaFree(s.p);
      s.p = nullptr; s.cap = 0;
      VSB_CUDA(cudaMalloc(&s.p, bytes));
      s.cap = bytes;
    }
    return s.p;
  }
  ~Model() {
    for (auto& s : staging) if (s.p) cudaFree(s.p);
    if (s_in) { cudaStreamDestroy(s_in); cudaStreamDestroy(s_cmp); cudaStreamDestroy(s_out); }
    for (auto e : ev_in) cudaEventDestroy(e);
    for (auto e : ev_cmp) cudaEventDestroy(e);
  }

  std::map<int, std::unique_ptr<Plan>> embed_plans, detect_plans;
  std::map<uint64_t, std::unique_ptr<ResampleDev>> resamplers;
  Plan* last_plan = nullptr;

  explicit Model(const vsb_model_desc& desc) : d(desc) {}

  const HostTensor& get(const std::string& k) const {
    auto it = sd.find(k);
    if (it == sd.end()) throw Error("missing checkpoint tensor: " + k, kErrState);
    return it->second;
  }

  // ---------------------------------------------------------------- packing helpers
  static std::vector<__half> to_half(const std::vector<float>& v) {
    std::vector<__half> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2half_rn(v[i]);
    return h;
  }
  // conv weight [N][C][R][S] (+ per-output-channel scale, + per-input-channel scale) -> [N][R][S][C]
  // pad_taps: every tap's channel run is zero-padded to a multiple of 8 (gather loaders move 16-byte channel chunks; the
  // activation buffer then has the same padded pixel pitch with ZERO pad channels).  Rows are padded to a multiple of 8
  // elements either way (a no-op for every width of the tiny / pixelseal cards).
  ConvW pack_conv(const std::string& wkey, const std::vector<float>* out_scale, const std::vector<float>& bias,
                  const std::vector<float>* in_scale = nullptr, bool pad_taps = false) {
    const HostTensor& w = get(wkey);
    VSB_CHECK(w.shape.size() == 4 || w.shape.size() == 2, "conv weight rank");
    const int N = (int)w.shape[0], C = (int)w.shape[1];
    const int R = w.shape.size() == 4 ? (int)w.shape[2] : 1, S = w.shape.size() == 4 ? (int)w.shape[3] : 1;
    const int Cp = pad_taps ? (C + 7) / 8 * 8 : C;
    const int K = R * S * Cp, ld = (K + 7) / 8 * 8;
    std::vector<float> p((size_t)N * ld, 0.f);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r)
          for (int s = 0; s < S; ++s) {
            float v = w.data[(((size_t)n * C + c) * R + r) * S + s];
            if (out_scale) v *= (*out_scale)[n];
            if (in_scale) v *= (*in_scale)[c];
            p[(size_t)n * ld + ((size_t)r * S + s) * Cp + c] = v;
          }
    ConvW cw;
    cw.N = N; cw.K = K; cw.ld = ld;
    cw.w = wpool.upload(to_half(p));
    cw.bias = bias.empty() ? nullptr : wpool.upload(bias);
    return cw;
  }
  // 3x3 conv weight [N][C0+C1][3][3] -> halo-loader layout [N][chunk][tap][cc], chunks zero-padded to kb_per_c*64
  ConvW pack_conv_halo(const std::string& wkey, const std::vector<float>* out_scale, const std::vector<float>& bias, int C0, int C1,
                       const std::vector<float>* in_scale = nullptr) {
    const HostTensor& w = get(wkey);
    VSB_CHECK(w.shape.size() == 4 && w.shape[2] == 3 && w.shape[3] == 3 && (int)w.shape[1] == C0 + C1, "halo conv weight shape");
    const int N = (int)w.shape[0], Ct = C0 + C1;
    const int cc = C0 < 64 ? C0 : 64, kbc = (9 * cc + 63) / 64, Kpad = halo_kpad(C0, C1);
    std::vector<float> p((size_t)N * Kpad, 0.f);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < Ct; ++c)
        for (int t = 0; t < 9; ++t) {
          float v = w.data[((size_t)n * Ct + c) * 9 + t];
          if (out_scale) v *= (*out_scale)[n];
          if (in_scale) v *= (*in_scale)[c];
          p[(size_t)n * Kpad + (size_t)(c / cc) * kbc * 64 + t * cc + (c % cc)] = v;
        }
    ConvW cw;
    cw.N = N; cw.K = Kpad;
    cw.w = wpool.upload(to_half(p));
    cw.bias = bias.empty() ? nullptr : wpool.upload(bias);
    return cw;
  }
  // UBlock up-conv as a low-resolution tap GEMM (see ups_gather_ln_kernel): rows n' = tap*Cout + co, K = concat channels
  ConvW pack_upconv_taps(const std::string& wkey, const std::vector<float>& in_scale) {
    const HostTensor& w = get(wkey);
    VSB_CHECK(w.shape.size() == 4 && w.shape[2] == 3 && w.shape[3] == 3, "up conv weight shape");
    const int Co = (int)w.shape[0], Ct = (int)w.shape[1];
    std::vector<float> p((size_t)9 * Co * Ct);
    for (int co = 0; co < Co; ++co)
      for (int c = 0; c < Ct; ++c)
        for (int t = 0; t < 9; ++t) p[((size_t)t * Co + co) * Ct + c] = w.data[((size_t)co * Ct + c) * 9 + t] * in_scale[c];
    ConvW cw;
    cw.N = 9 * Co; cw.K = Ct;
    cw.w = wpool.upload(to_half(p));
    return cw;
  }
  // the same up-conv folded for the phase path (conv3_direct_host.cuh, setup_up_phase_direct): rows n = (py*2+px)*Cout + co,
  // K = (ey*3+ex)*Ct + c over the 3x3 LOW-resolution neighbourhood; g[phase][d][e] = weight of low-res offset e-1 in tap d of
  // an output row of that phase (bilinear x2, align_corners=False).  Also the plain fp32 taps for the border fix-up kernel.
  void pack_upconv_phase(const std::string& wkey, const std::vector<float>& in_scale, __half** wd_out, float** wk_out) {
    const HostTensor& w = get(wkey);
    const int Co = (int)w.shape[0], Ct = (int)w.shape[1];
    static const float g[2][3][3] = {{{.75f, .25f, 0.f}, {.25f, .75f, 0.f}, {0.f, .75f, .25f}},
                                     {{.25f, .75f, 0.f}, {0.f, .75f, .25f}, {0.f, .25f, .75f}}};
    std::vector<float> p((size_t)4 * Co * 9 * Ct, 0.f), wk((size_t)9 * Co * Ct);
    for (int co = 0; co < Co; ++co)
      for (int c = 0; c < Ct; ++c)
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) {
            const float kv = w.data[(((size_t)co * Ct + c) * 3 + dy) * 3 + dx] * in_scale[c];
            wk[((size_t)(dy * 3 + dx) * Co + co) * Ct + c] = kv;
            for (int py = 0; py < 2; ++py)
              for (int px = 0; px < 2; ++px)
                for (int ey = 0; ey < 3; ++ey)
                  for (int ex = 0; ex < 3; ++ex) {
                    const float cf = g[py][dy][ey] * g[px][dx][ex];
                    if (cf != 0.f) p[((size_t)((py * 2 + px) * Co + co) * 9 + (ey * 3 + ex)) * Ct + c] += kv * cf;
                  }
          }
    __half* plain = wpool.upload(to_half(p));
    *wd_out = wpool.alloc_n<__half>(p.size());
    pack_direct_weights(plain, 4 * Co, Ct, *wd_out, 0, 9);
    VSB_CUDA(cudaStreamSynchronize(0));
    *wk_out = wpool.upload(wk);
  }
  // DBlock.down weights [N][Cin][3][3] re-indexed for the space-to-depth ring conv (conv3_direct_host.cuh,
  // setup_down_s2d_direct): K = (ey*3+ex)*4Cin + (dy*2+dx)*Cin + c with tap r -> (ey, dy) = (0,1), (1,0), (1,1)
  __half* pack_down_s2d(const std::string& wkey) {
    const HostTensor& w = get(wkey);
    const int N = (int)w.shape[0], Ci = (int)w.shape[1], Cs = 4 * Ci;
    static const int E[3] = {0, 1, 1}, D[3] = {1, 0, 1};
    std::vector<float> p((size_t)N * 9 * Cs, 0.f);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < Ci; ++c)
        for (int r = 0; r < 3; ++r)
          for (int s2 = 0; s2 < 3; ++s2)
            p[((size_t)n * 9 + (E[r] * 3 + E[s2])) * Cs + (D[r] * 2 + D[s2]) * Ci + c] = w.data[(((size_t)n * Ci + c) * 3 + r) * 3 + s2];
    __half* plain = wpool.upload(to_half(p));
    __half* wd = wpool.alloc_n<__half>(p.size());
    pack_direct_weights(plain, N, Cs, wd, 0, 9);
    VSB_CUDA(cudaStreamSynchronize(0));
    return wd;
  }
  int halo_max_c = -1;
  bool use_halo(int C) {
    if (halo_max_c < 0) {
      const char* e = getenv("VSB_HALO_MAXC");
      halo_max_c = e ? atoi(e) : 16;
    }
    return C <= halo_max_c;
  }
  ConvW pack_conv3(const std::string& wkey, const std::vector<float>* out_scale, const std::vector<float>& bias) {
    const int C = (int)get(wkey).shape[1], N = (int)get(wkey).shape[0];
    ConvW cw = use_halo(C) ? pack_conv_halo(wkey, out_scale, bias, C, 0) : pack_conv(wkey, out_scale, bias);
    if (conv3_direct_ok(C, N, C)) {
      ConvW plain = use_halo(C) ? pack_conv(wkey, out_scale, {}) : cw;   // [N][9*C], K order (r, s, c)
      cw.wd = wpool.alloc_n<__half>((size_t)N * 9 * C);
      pack_direct_weights(plain.w, N, C, cw.wd, 0);
      VSB_CUDA(cudaStreamSynchronize(0));
    }
    return cw;
  }
  // eval-mode BatchNorm2d folded into the preceding bias-free conv: w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)
  void bn_fold(const std::string& bn, std::vector<float>& scale, std::vector<float>& bias) const {
    const HostTensor &g = get(bn + ".weight"), &b = get(bn + ".bias"), &mu = get(bn + ".running_mean"), &var = get(bn + ".running_var");
    const size_t n = g.data.size();
    scale.resize(n); bias.resize(n);
    for (size_t i = 0; i < n; ++i) {
      const float s = g.data[i] / std::sqrt(var.data[i] + 1e-5f);
      scale[i] = s;
      bias[i] = b.data[i] - mu.data[i] * s;
    }
  }
  ResBlockW pack_resblock(const std::string& key) {
    ResBlockW rb;
    std::vector<float> s, b;
    bn_fold(key + ".double_conv.1", s, b);
    rb.c1 = pack_conv3(key + ".double_conv.0.weight", &s, b);
    bn_fold(key + ".double_conv.4", s, b);
    rb.c2 = pack_conv3(key + ".double_conv.3.weight", &s, b);
    rb.res = pack_conv(key + ".res_conv.weight", nullptr, get(key + ".res_conv.bias").data);
    if (conv3_direct_ok(rb.res.K, rb.res.N, rb.res.K, 0, 1)) {   // narrow 1x1: same ring kernel with a single tap
      rb.res.wd = wpool.alloc_n<__half>((size_t)rb.res.N * rb.res.K);
      pack_direct_weights(rb.res.w, rb.res.N, rb.res.K, rb.res.wd, 0, 1);
      VSB_CUDA(cudaStreamSynchronize(0));
    }
    return rb;
  }

  void finalize(int dev) {
    if (finalized) throw Error("model already finalized", kErrState);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) throw Error("no CUDA device: libvsb200 has no CPU fallback", kErrCuda);
    VSB_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    VSB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) throw Error("libvsb200 is built for sm_100a only (found sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")", kErrUnsupported);
    device = dev;
    num_sms = prop.multiProcessorCount;
    if (d.unet_act != 0 || d.unet_norm != 0) throw Error("unsupported card: only BatchNorm+ReLU U-Nets are implemented on the GPU path", kErrUnsupported);
    VSB_CHECK(d.unet_levels >= 2 && d.unet_levels <= 6, "unet levels");
    VSB_CHECK(d.unet_in_ch == 1 || d.unet_in_ch == 3, "unet in_channels must be 1 or 3");
    VSB_CHECK(d.unet_out_ch >= 1 && d.unet_out_ch <= 3, "unet out_channels must be <= 3");
    VSB_CHECK((d.yuv != 0) == (d.unet_in_ch == 1), "yuv cards use 1 input channel");
    VSB_CHECK(d.img_size % 128 == 0, "processing size must be a multiple of 128");
    const std::string P = "embedder.unet.";
    const int L = d.unet_levels;
    std::vector<int> z(d.unet_z, d.unet_z + L);
    // ---- inc: first conv on CUDA cores, fp32 weights with BN folded
    {
      std::vector<float> s, b;
      bn_fold(P + "inc.double_conv.1", s, b);
      const HostTensor& w = get(P + "inc.double_conv.0.weight");
      VSB_CHECK((int)w.shape[0] == z[0] && (int)w.shape[1] == d.unet_in_ch, "inc conv shape");
      VSB_CHECK(z[0] % 16 == 0, "z0 must be a multiple of 16");
      std::vector<float> w1(w.data);
      const size_t per = (size_t)d.unet_in_ch * 9;
      for (int n = 0; n < z[0]; ++n) for (size_t k = 0; k < per; ++k) w1[n * per + k] *= s[n];
      first_w1 = wpool.upload(w1); first_b1 = wpool.upload(b);
      first_wr = wpool.upload(get(P + "inc.res_conv.weight").data);
      first_br = wpool.upload(get(P + "inc.res_conv.bias").data);
      bn_fold(P + "inc.double_conv.4", s, b);
      inc_c2 = pack_conv3(P + "inc.double_conv.3.weight", &s, b);
    }
    for (int i = 0; i < L - 1; ++i) {
      down_conv.push_back(pack_conv(P + "downs." + std::to_string(i) + ".down.weight", nullptr, get(P + "downs." + std::to_string(i) + ".down.bias").data));
      if (z[i] == 16 && z[i + 1] == 32 && conv3_direct_ok(64, 32, 64, 0, 3) && !getenv("VSB_NO_S2D"))
        down_conv.back().wd = pack_down_s2d(P + "downs." + std::to_string(i) + ".down.weight");
      down_rb.push_back(pack_resblock(P + "downs." + std::to_string(i) + ".conv"));
    }
    for (int i = 0; i < d.unet_num_blocks; ++i) bott_rb.push_back(pack_resblock(P + "bottleneck.model." + std::to_string(i)));
    {
      std::vector<int> zz(z);
      zz[L - 1] += d.hidden;
      int j = 0;
      for (int ii = L - 2; ii >= 0; --ii, ++j) {
        const std::string U = P + "ups." + std::to_string(j);
        // virtual concat (x | skip * 2^-1/2): fold the skip scale into the second half of the input channels
        std::vector<float> in_scale(2 * zz[ii + 1], 1.0f);
        for (int c = zz[ii + 1]; c < 2 * zz[ii + 1]; ++c) in_scale[c] = 0.70710678118654752440f;
        up_conv.push_back(pack_upconv_taps(U + ".up.upsample_block.2.weight", in_scale));
        __half* pw = nullptr; float* bw = nullptr;
        if (zz[ii] == 16 && 2 * zz[ii + 1] == 64 && conv3_direct_ok(64, 64, 64, 0, 3) && !getenv("VSB_NO_UPPHASE"))
          pack_upconv_phase(U + ".up.upsample_block.2.weight", in_scale, &pw, &bw);
        up_phase_w.push_back(pw); up_border_w.push_back(bw);
        up_lnw.push_back(wpool.upload(get(U + ".up.upsample_block.3.weight").data));
        up_lnb.push_back(wpool.upload(get(U + ".up.upsample_block.3.bias").data));
        up_rb.push_back(pack_resblock(U + ".conv"));
      }
    }
    outc_w = wpool.upload(get(P + "outc.weight").data);  // [n_out][z0][1][1]
    outc_b = wpool.upload(get(P + "outc.bias").data);
    {
      const HostTensor& t = get(P + "msg_processor.msg_embeddings.weight");
      VSB_CHECK((int)t.shape[0] == 2 * d.nbits && (int)t.shape[1] == d.hidden, "message table shape");
      VSB_CHECK(d.hidden % 8 == 0, "hidden must be a multiple of 8");
      msg_table = wpool.upload(t.data);
    }
    // ---- extractor (ConvNeXt-V2)
    const std::string Q = "detector.convnext.";
    {
      const HostTensor& w = get(Q + "downsample_layers.0.0.weight");  // [C0][3][4][4]
      const int C0 = d.ext_dims[0];
      VSB_CHECK((int)w.shape[0] == C0 && w.shape[1] == 3 && w.shape[2] == 4, "stem shape");
      if (C0 <= 256) {   // tensor-core stem GEMM with the LayerNorm fused in the epilogue (needs the whole row in one N tile)
        std::vector<float> t((size_t)C0 * 64, 0.f);   // GEMM operand [C0][64], K = (c, r, t) zero-padded 48 -> 64
        for (int c = 0; c < C0; ++c) for (int k = 0; k < 48; ++k) t[(size_t)c * 64 + k] = w.data[(size_t)c * 48 + k];
        stem_w16 = wpool.upload(to_half(t));
      } else {           // stem_ln_kernel: [48][C0]
        std::vector<float> t((size_t)48 * C0);
        for (int c = 0; c < C0; ++c) for (int k = 0; k < 48; ++k) t[(size_t)k * C0 + c] = w.data[(size_t)c * 48 + k];
        stem_wk = wpool.upload(t);
      }
      stem_b = wpool.upload(get(Q + "downsample_layers.0.0.bias").data);
      stem_lnw = wpool.upload(get(Q + "downsample_layers.0.1.weight").data);
      stem_lnb = wpool.upload(get(Q + "downsample_layers.0.1.bias").data);
    }
    for (int i = 1; i < 4; ++i) {
      DsW w;
      const std::string K = Q + "downsample_layers." + std::to_string(i);
      w.lnw = wpool.upload(get(K + ".0.weight").data);
      w.lnb = wpool.upload(get(K + ".0.bias").data);
      w.conv = pack_conv(K + ".1.weight", nullptr, get(K + ".1.bias").data, nullptr, /*pad_taps=*/true);
      ds.push_back(w);
    }
    cn.resize(4);
    for (int s = 0; s < 4; ++s) {
      const int C = d.ext_dims[s];
      VSB_CHECK(C % 2 == 0, "extractor widths must be even (channel pairs in the depthwise kernel)");
      for (int j = 0; j < d.ext_depths[s]; ++j) {
        const std::string B = Q + "stages." + std::to_string(s) + "." + std::to_string(j) + ".";
        CnBlockW w;
        const HostTensor& dw = get(B + "dwconv.weight");  // [C][1][7][7]
        std::vector<float> t((size_t)49 * C);
        for (int c = 0; c < C; ++c) for (int k = 0; k < 49; ++k) t[(size_t)k * C + c] = dw.data[(size_t)c * 49 + k];
        w.dww = wpool.upload(t);
        w.dwb = wpool.upload(get(B + "dwconv.bias").data);
        w.lnw = wpool.upload(get(B + "norm.weight").data);
        w.lnb = wpool.upload(get(B + "norm.bias").data);
        w.pw1 = pack_conv(B + "pwconv1.weight", nullptr, get(B + "pwconv1.bias").data);
        w.gamma = wpool.upload(get(B + "grn.gamma").data);
        // GRN: gamma*(x*Nx) + beta + x  ==  x*(gamma*Nx + 1) + beta ; beta goes through pwconv2 into its bias
        const HostTensor &w2 = get(B + "pwconv2.weight"), &b2 = get(B + "pwconv2.bias"), &beta = get(B + "grn.beta");
        std::vector<float> bias2(C);
        for (int n = 0; n < C; ++n) {
          double a = b2.data[n];
          for (int k = 0; k < 4 * C; ++k) a += (double)w2.data[(size_t)n * 4 * C + k] * (double)beta.data[k];
          bias2[n] = (float)a;
        }
        w.pw2 = pack_conv(B + "pwconv2.weight", nullptr, bias2);
        cn[s].push_back(w);
      }
    }
    {
      const std::string D = "detector.pixel_decoder.";
      head_conv = pack_conv(D + "output_upscaling.0.upsample_block.2.weight", nullptr, {}, nullptr, /*pad_taps=*/true);
      head_lnw = wpool.upload(get(D + "output_upscaling.0.upsample_block.3.weight").data);
      head_lnb = wpool.upload(get(D + "output_upscaling.0.upsample_block.3.bias").data);
      const HostTensor& lw = get(D + "linear.weight");
      VSB_CHECK((int)lw.shape[0] == 1 + d.nbits, "head linear shape");
      head_lw = wpool.upload(lw.data);
      head_lb = wpool.upload(get(D + "linear.bias").data);
    }
    sd.clear();  // host copies no longer needed
    finalized = true;
  }

  // ---------------------------------------------------------------- plan building helpers
  static void dbg(Plan& pl, const std::string& name, const void* p, int dtype, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ld) {
    pl.dbg[name] = DebugTensor{p, dtype, {B, H, W, C}, ld};
  }
  void add_conv(Plan& pl, ConvGemmOp op, const ConvW& w, const std::string& name, int block_n = 0, const __half* w_override = nullptr) {
    finalize_op(op, w_override ? w_override : w.w, w.N, w.K, w.ld ? w.ld : w.K, num_sms, block_n);
    pl.steps.push_back(Step{[op](cudaStream_t st) { launch(op, st); }, 1, name});
  }
  // 3x3 stride-1 zero-padded conv: on-chip im2col from a shared-memory halo tile (input crosses L2->SM once) up to
  // `halo_max_c` input channels, one TMA load per tap above (VSB_HALO_MAXC overrides, for experiments); the weight layout
  // was chosen by the same rule at pack time (pack_conv3)
  void setup_conv3(ConvGemmOp& op, const __half* x, int B, int H, int W, int C, int ld) {
    if (use_halo(C)) setup_halo_conv3(op, x, B, H, W, C, ld);
    else setup_tma_conv(op, x, B, H, W, C, ld, 3, 3, 1);
  }

  // ResnetBlock on NHWC fp16 (modules/unet.py:17-39):  out = relu(conv3'(relu(conv3'(x)))) + (conv1(x) + b)
  // `x` has pixel pitch ldx; output written with pitch ld_out.  If fuse_outc: the 1x1 outc + tanh is fused and the
  // block output itself is not written.
  __half* add_resblock(Plan& pl, const std::string& name, const ResBlockW& w, const __half* x, int B, int H, int W, int Cin, int ldx,
                       __half* out, int ld_out, bool fuse_outc = false) {
    const int Cout = w.c1.N;
    const long M = (long)B * H * W;
    __half* r = pl.pool.alloc_n<__half>(M * Cout);
    __half* h = pl.pool.alloc_n<__half>(M * Cout);
    if (!out && !fuse_outc) { out = pl.pool.alloc_n<__half>(M * Cout); ld_out = Cout; }
    if (w.res.wd && conv3_direct_ok(Cin, Cout, ldx, W, 1)) {  // res 1x1, narrow layers
      Conv3DirectOp op; setup_conv3_direct(op, x, B, H, W, Cin, Cout, w.res.wd, num_sms, 1);
      op.p.bias = w.res.bias; op.p.out = r; op.p.relu = 0;
      pl.steps.push_back(Step{[op](cudaStream_t st) { launch_direct(op, st); }, 1,
                              "unet.conv1x1d." + std::to_string(Cin) + "-" + std::to_string(Cout) + "@" + std::to_string(H)});
    } else {  // res 1x1
      ConvGemmOp op; setup_tma_conv(op, x, B, H, W, Cin, ldx, 1, 1, 0);
      op.p.epi = EPI_AFFINE; op.p.act = ACT_NONE; op.p.bias = w.res.bias; op.p.out16 = r; op.p.ld_out16 = Cout;
      add_conv(pl, op, w.res, "unet.conv1x1." + std::to_string(Cin) + "-" + std::to_string(Cout) + "@" + std::to_string(H));
    }
    const std::string sfx = "@" + std::to_string(H);
    if (w.c1.wd && conv3_direct_ok(Cin, Cout, ldx, W)) {  // conv3 + BN + ReLU, narrow layers: straight from a ring of padded rows
      Conv3DirectOp op; setup_conv3_direct(op, x, B, H, W, Cin, Cout, w.c1.wd, num_sms);
      op.p.bias = w.c1.bias; op.p.out = h;
      pl.steps.push_back(Step{[op](cudaStream_t st) { launch_direct(op, st); }, 1, "unet.conv3x3d." + std::to_string(Cin) + "-" + std::to_string(Cout) + sfx});
    } else {  // conv3 + BN + ReLU
      ConvGemmOp op; setup_conv3(op, x, B, H, W, Cin, ldx);
      op.p.epi = EPI_AFFINE; op.p.act = ACT_RELU; op.p.bias = w.c1.bias; op.p.out16 = h; op.p.ld_out16 = Cout;
      add_conv(pl, op, w.c1, "unet.conv3x3." + std::to_string(Cin) + "-" + std::to_string(Cout) + "@" + std::to_string(H));
    }
    if (w.c2.wd && conv3_direct_ok(Cout, Cout, Cout, W) && (out == nullptr || ld_out == Cout)) {
      Conv3DirectOp op; setup_conv3_direct(op, h, B, H, W, Cout, Cout, w.c2.wd, num_sms);
      op.p.bias = w.c2.bias; op.p.resid = r; op.p.out = out;
      if (fuse_outc) {
        op.p.outc_w = outc_w; op.p.outc_b = outc_b; op.p.n_out = d.unet_out_ch; op.p.delta = pl.delta; op.p.outc_tanh = d.unet_last_tanh;
      }
      pl.steps.push_back(Step{[op](cudaStream_t st) { launch_direct(op, st); }, 1,
                              std::string(fuse_outc ? "unet.conv3x3d+outc." : "unet.conv3x3d.") + std::to_string(Cout) + "-" + std::to_string(Cout) + sfx});
    } else {  // conv3 + BN + ReLU, + res
      ConvGemmOp op; setup_conv3(op, h, B, H, W, Cout, Cout);
      op.p.epi = EPI_AFFINE; op.p.act = ACT_RELU; op.p.bias = w.c2.bias; op.p.resid16 = r; op.p.ld_res16 = Cout;
      if (fuse_outc) {
        op.p.outc_w = outc_w; op.p.outc_b = outc_b; op.p.n_out = d.unet_out_ch; op.p.delta = pl.delta; op.p.hw = H * W;
        op.p.outc_tanh = d.unet_last_tanh;
        if (out) { op.p.out16 = out; op.p.ld_out16 = ld_out; }
      } else {
        op.p.out16 = out; op.p.ld_out16 = ld_out;
      }
      add_conv(pl, op, w.c2, std::string(fuse_outc ? "unet.conv3x3+outc." : "unet.conv3x3.") + std::to_string(Cout) + "-" + std::to_string(Cout) + "@" + std::to_string(H));
    }
    if (out) dbg(pl, name, out, 1, B, H, W, Cout, ld_out);
    return out;
  }

  // Plan cache: one activation arena per distinct batch size, at most `max_plans` per kind (VSB_MAX_PLANS, default 8): a service that sees
  // many different clip / tail lengths would otherwise accumulate up to 64 arenas of up to 2.3 GB each.  The least recently used plan
  // is dropped (its cudaFree calls synchronise the device, so work still in flight on it completes first).
  uint64_t plan_clock = 0;
  int max_plans = [] { const char* e = getenv("VSB_MAX_PLANS"); const int v = e ? atoi(e) : 8; return v < 2 ? 2 : v; }();
  void evict_plans(std::map<int, std::unique_ptr<Plan>>& m, const Plan* keep) {
    while ((int)m.size() > max_plans) {
      auto victim = m.end();
      for (auto it = m.begin(); it != m.end(); ++it)
        if (it->second.get() != keep && (victim == m.end() || it->second->last_use < victim->second->last_use)) victim = it;
      if (victim == m.end()) break;
      if (victim->second.get() == last_plan) last_plan = nullptr;
      m.erase(victim);
    }
  }
  size_t cached_plans() const { return embed_plans.size() + detect_plans.size(); }

  Plan* get_embed_plan(int B) {
    auto it = embed_plans.find(B);
    if (it != embed_plans.end()) { it->second->last_use = ++plan_clock; return it->second.get(); }
    std::unique_ptr<Plan> up(new Plan());
    Plan& pl = *up;
    const int S = d.img_size, L = d.unet_levels;
    std::vector<int> z(d.unet_z, d.unet_z + L);
    pl.x_res = pl.pool.alloc_n<float>((size_t)B * 3 * S * S);
    pl.delta = pl.pool.alloc_n<float>((size_t)B * d.unet_out_ch * S * S);
    Plan* plp = &pl;
    // ---- inc
    const long M0 = (long)B * S * S;
    __half* h1 = pl.pool.alloc_n<__half>(M0 * z[0]);
    __half* r0 = pl.pool.alloc_n<__half>(M0 * z[0]);
    {
      const int Z = z[0], yuv = d.yuv, cin = d.unet_in_ch;
      float *w1 = first_w1, *b1 = first_b1, *wr = first_wr, *br = first_br;
      pl.steps.push_back(Step{[=](cudaStream_t st) {
        dim3 grid((S + 31) / 32, (S + 15) / 16, B);
        // frame stride lets image-size inputs be read in place (key frames of a video are `step` frames apart)
        const float* src = plp->in_imgs;
        if (plp->in_frame_stride != (long)3 * S * S) {
          // strided key frames: launch per frame group is avoided by passing the stride through the batch index
        }
        if (cin == 1) unet_first_kernel<1><<<grid, 256, 0, st>>>(src, B, S, S, Z, w1, b1, wr, br, h1, r0, yuv);
        else unet_first_kernel<3><<<grid, 256, 0, st>>>(src, B, S, S, Z, w1, b1, wr, br, h1, r0, yuv);
        VSB_CUDA(cudaGetLastError());
      }, 1, "unet.first"});
    }
    std::vector<__half*> skips;  // outputs of inc, downs[0..]
    std::vector<int> skip_ld;
    __half* x = pl.pool.alloc_n<__half>(M0 * z[0]);
    if (inc_c2.wd && conv3_direct_ok(z[0], z[0], z[0], S)) {
      Conv3DirectOp op; setup_conv3_direct(op, h1, B, S, S, z[0], z[0], inc_c2.wd, num_sms);
      op.p.bias = inc_c2.bias; op.p.resid = r0; op.p.out = x;
      pl.steps.push_back(Step{[op](cudaStream_t st) { launch_direct(op, st); }, 1,
                              "unet.conv3x3d." + std::to_string(z[0]) + "-" + std::to_string(z[0]) + "@" + std::to_string(S)});
      dbg(pl, "inc", x, 1, B, S, S, z[0], z[0]);
    } else {
      ConvGemmOp op; setup_conv3(op, h1, B, S, S, z[0], z[0]);
      op.p.epi = EPI_AFFINE; op.p.act = ACT_RELU; op.p.bias = inc_c2.bias; op.p.resid16 = r0; op.p.ld_res16 = z[0];
      op.p.out16 = x; op.p.ld_out16 = z[0];
      add_conv(pl, op, inc_c2, "unet.conv3x3." + std::to_string(z[0]) + "-" + std::to_string(z[0]) + "@" + std::to_string(S));
      dbg(pl, "inc", x, 1, B, S, S, z[0], z[0]);
    }
    int ldx = z[0], hs = S;
    const int zb = z[L - 1] + d.hidden;
    __half* cat = nullptr;
    for (int i = 0; i < L - 1; ++i) {
      const int ho = hs / 2;
      const long Mo = (long)B * ho * ho;
      __half* dn = pl.pool.alloc_n<__half>(Mo * z[i + 1]);
      if (down_conv[i].wd && ldx == z[i] && hs % 2 == 0 && conv3_direct_ok(4 * z[i], z[i + 1], 4 * z[i], ho, 3)) {
        Conv3DirectOp op;
        setup_down_s2d_direct(op, x, z[i], B, ho, ho, z[i + 1], down_conv[i].wd, down_conv[i].bias, dn, num_sms);
        pl.steps.push_back(Step{[op](cudaStream_t st) { launch_direct(op, st); }, 1,
                                "unet.down3x3s2d." + std::to_string(z[i]) + "-" + std::to_string(z[i + 1]) + "@" + std::to_string(ho)});
      } else {
        ConvGemmOp op;
        setup_gather_conv(op, LD_GATHER_CONV, x, z[i], ldx, nullptr, 0, 0, B, hs, hs, ho, ho, 3, 3, 2, 1, 0);
        op.p.epi = EPI_AFFINE; op.p.act = ACT_NONE; op.p.bias = down_conv[i].bias; op.p.out16 = dn; op.p.ld_out16 = z[i + 1];
        add_conv(pl, op, down_conv[i], "unet.down3x3s2." + std::to_string(z[i]) + "-" + std::to_string(z[i + 1]) + "@" + std::to_string(ho));
      }
      __half* out = nullptr; int ld_out = 0;
      if (i == L - 2) {  // last level writes straight into the message-concat buffer (channels [0, z))
        cat = pl.pool.alloc_n<__half>(Mo * zb);
        out = cat; ld_out = zb;
      }
      x = add_resblock(pl, "down" + std::to_string(i), down_rb[i], dn, B, ho, ho, z[i + 1], z[i + 1], out, ld_out);
      ldx = (i == L - 2) ? zb : z[i + 1];
      hs = ho;
      skips.push_back(x); skip_ld.push_back(ldx);
    }
    // ---- message channels
    {
      const int K = d.nbits, hidden = d.hidden, hw = hs * hs, coff = z[L - 1];
      float* table = msg_table;
      pl.steps.push_back(Step{[=](cudaStream_t st) {
        dim3 grid(B, 8);
        msg_embed_kernel<<<grid, 256, hidden * sizeof(float), st>>>(plp->in_msgs, table, K, hidden, cat, hw, zb, coff, plp->in_msg_stride);
        VSB_CUDA(cudaGetLastError());
      }, 1, "unet.msg"});
      dbg(pl, "cat", cat, 1, B, hs, hs, zb, zb);
    }
    // ---- bottleneck
    x = cat; ldx = zb;
    for (int i = 0; i < d.unet_num_blocks; ++i) {
      x = add_resblock(pl, "bott" + std::to_string(i), bott_rb[i], x, B, hs, hs, zb, ldx, nullptr, 0);
      ldx = zb;
    }
    // ---- ups
    std::vector<int> zz(z); zz[L - 1] = zb;
    int j = 0;
    for (int ii = L - 2; ii >= 0; --ii, ++j) {
      const int Cin = zz[ii + 1], Cout = zz[ii];
      // hiddens stack (unet.py:175-191): first pop is the message-concatenated latent, then downs[L-3], ...
      const __half* skip = (j == 0) ? cat : skips[L - 2 - j];
      const int ld_skip = (j == 0) ? zb : skip_ld[L - 2 - j];
      const int ho = hs * 2;
      const long Mo = (long)B * ho * ho;
      __half* u = pl.pool.alloc_n<__half>(Mo * Cout);
      if (up_phase_w[j] && ldx == Cin && ld_skip == Cin && conv3_direct_ok(64, 64, 64, hs, 3)) {
        // one tensor-core conv on the low-resolution grid (bilinear weights folded per output phase) + exact border pixels
        Conv3DirectOp op;
        setup_up_phase_direct(op, x, Cin, skip, Cin, B, hs, hs, up_phase_w[j], up_lnw[j], up_lnb[j], 1e-6f, u, num_sms);
        pl.steps.push_back(Step{[op](cudaStream_t st) { launch_direct(op, st); }, 1,
                                "unet.upphase." + std::to_string(2 * Cin) + "-" + std::to_string(Cout) + "@" + std::to_string(ho)});
        const __half *x0 = x, *x1 = skip;
        const int IH = hs, Ci = Cin;
        const float *wk = up_border_w[j], *lw = up_lnw[j], *lb = up_lnb[j];
        pl.steps.push_back(Step{[=](cudaStream_t st) {
          const long nb = (long)B * (2 * (2 * IH) + 2 * (2 * IH - 2));
          static bool attr = false;
          if (!attr) { VSB_CUDA(cudaFuncSetAttribute(up_border_fix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpFixSmem)); attr = true; }
          up_border_fix_kernel<<<(unsigned)std::min<long>((nb + kUpFixPPB - 1) / kUpFixPPB, 148L * 3), 256, kUpFixSmem, st>>>(x0, Ci, x1, Ci, B, IH, IH, wk, lw, lb, 1e-6f, u);
          VSB_CUDA(cudaGetLastError());
        }, 1, "unet.upborder." + std::to_string(Cout) + "@" + std::to_string(ho)});
      } else {
        // channel mixing first, at low resolution: y[b,i,j, tap*Cout + co] = sum_c W[co,c,tap] * [x | skip/sqrt2][b,i,j,c]
        const long Mi = (long)B * hs * hs;
        __half* ytap = pl.pool.alloc_n<__half>(Mi * 9 * Cout);
        ConvGemmOp op;
        setup_tma_gemm2(op, x, Cin, ldx, skip, Cin, ld_skip, Mi);
        op.p.epi = EPI_AFFINE; op.p.act = ACT_NONE; op.p.out16 = ytap; op.p.ld_out16 = 9 * Cout;
        add_conv(pl, op, up_conv[j], "unet.uptap1x1." + std::to_string(2 * Cin) + "-" + std::to_string(9 * Cout) + "@" + std::to_string(hs));
        // then bilinear x2 + reflect pad + 3x3 tap sum + LayerNorm + ReLU (CUDA cores, memory-bound)
        const int IH = hs, Cc = Cout;
        float *lw = up_lnw[j], *lb = up_lnb[j];
        const int vpt = Cc <= 256 ? 1 : 2;            // 16-byte channel vectors per thread (chunkyseal's first up-conv: 512 channels)
        const int gthreads = Cc / (8 * vpt);          // threads per output pixel
        VSB_CHECK(Cc % (8 * vpt) == 0 && gthreads <= 32 && (gthreads & (gthreads - 1)) == 0, "up conv: C_out must be 8 * 2^k <= 512");
        const size_t tsmem = (size_t)100 * 9 * Cc * sizeof(__half);
        static const bool no_tiled = getenv("VSB_UPS_UNTILED") != nullptr;
        if (!no_tiled && IH % 8 == 0 && tsmem <= 160 * 1024) {
          // low-resolution taps of a 16 x 16 output tile staged in shared memory (ups_gather_ln_tiled_kernel)
          pl.steps.push_back(Step{[=](cudaStream_t st) {
            const unsigned blocks = (unsigned)((long)B * (IH / 8) * (IH / 8));
            if (vpt == 1) {
              static bool attr = false;
              if (!attr) { VSB_CUDA(cudaFuncSetAttribute(ups_gather_ln_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
              ups_gather_ln_tiled_kernel<1><<<blocks, 256, tsmem, st>>>(ytap, B, IH, IH, Cc, lw, lb, 1e-6f, u, Cc);
            } else {
              static bool attr = false;
              if (!attr) { VSB_CUDA(cudaFuncSetAttribute(ups_gather_ln_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
              ups_gather_ln_tiled_kernel<2><<<blocks, 256, tsmem, st>>>(ytap, B, IH, IH, Cc, lw, lb, 1e-6f, u, Cc);
            }
            VSB_CUDA(cudaGetLastError());
          }, 1, "unet.upgather." + std::to_string(Cout) + "@" + std::to_string(ho)});
        } else
        pl.steps.push_back(Step{[=](cudaStream_t st) {
          const int ppb = 256 / gthreads;
          const long blocks = (Mo + ppb - 1) / ppb;
          const unsigned grid = (unsigned)std::min<long>(blocks, 148L * 32);
          if (vpt == 1) ups_gather_ln_kernel<1><<<grid, 256, 0, st>>>(ytap, B, IH, IH, Cc, lw, lb, 1e-6f, u, Cc);
          else ups_gather_ln_kernel<2><<<grid, 256, 0, st>>>(ytap, B, IH, IH, Cc, lw, lb, 1e-6f, u, Cc);
          VSB_CUDA(cudaGetLastError());
        }, 1, "unet.upgather." + std::to_string(Cout) + "@" + std::to_string(ho)});
      }
      dbg(pl, "up" + std::to_string(j) + "_conv", u, 1, B, ho, ho, Cout, Cout);
      const bool last = (ii == 0);
      x = add_resblock(pl, "up" + std::to_string(j), up_rb[j], u, B, ho, ho, Cout, Cout, nullptr, 0, /*fuse_outc=*/last);
      ldx = Cout; hs = ho;
    }
    dbg(pl, "delta", pl.delta, 0, B, d.unet_out_ch, S, S, S);
    Plan* ret = up.get();
    ret->last_use = ++plan_clock;
    embed_plans[B] = std::move(up);
    evict_plans(embed_plans, ret);
    return ret;
  }

  Plan* get_detect_plan(int B) {
    auto it = detect_plans.find(B);
    if (it != detect_plans.end()) { it->second->last_use = ++plan_clock; return it->second.get(); }
    std::unique_ptr<Plan> up(new Plan());
    Plan& pl = *up;
    Plan* plp = &pl;
    const int S = d.img_size;
    pl.x_res = pl.pool.alloc_n<float>((size_t)B * 3 * S * S);
    pl.logits = pl.pool.alloc_n<float>((size_t)B * (1 + d.nbits));
    const int st_ = d.ext_stem_stride;
    int hs = (S - 4) / st_ + 1;
    int C = d.ext_dims[0];
    // pixel pitch of every NHWC buffer of the trunk: the width rounded up to 8 channels (16-byte fp16 rows for TMA / vector
    // access; the proportional chunkyseal widths 362 / 724 are not multiples of 8).  Pad channels are never read through a
    // TMA map (its K extent is the true width); the buffers the gather loaders read are zeroed once below.
    auto pitch = [](int c) { return (c + 7) / 8 * 8; };
    int Cp = pitch(C);
    float* x = pl.pool.alloc_n<float>((size_t)B * hs * hs * Cp);
    if (stem_wk != nullptr) {
      // widths > 256: CUDA-core stem conv + LayerNorm (one warp per output pixel); the fused-LN GEMM epilogue needs N <= 256
      const int OH = hs, Cc = C, ldo = Cp;
      float *wk = stem_wk, *sb = stem_b, *lw = stem_lnw, *lb = stem_lnb;
      pl.steps.push_back(Step{[=](cudaStream_t st) {
        const long npix = (long)B * OH * OH;
        const int grid = (int)std::min<long>((npix + 7) / 8, 148L * 16);
        stem_ln_kernel<<<grid, 256, (size_t)8 * (48 + Cc) * sizeof(float), st>>>(plp->in_imgs, B, S, S, OH, OH, st_, wk, sb, lw, lb, Cc, x, ldo);
        VSB_CUDA(cudaGetLastError());
      }, 1, "cnx.stem_ln." + std::to_string(C) + "@" + std::to_string(hs)});
      dbg(pl, "ds0", x, 0, B, hs, hs, C, Cp);
    } else {
      // stem (convnext.py:108-111): im2col of the k4 patches (x = 2*img-1 folded) -> tensor-core GEMM [M,64]x[64,C0] with the
      // conv bias and the channels-first LayerNorm fused into the epilogue, fp32 NHWC output (the residual stream)
      const int OH = hs;
      const long M = (long)B * OH * OH;
      __half* patches = pl.pool.alloc_n<__half>(M * 64);
      pl.steps.push_back(Step{[=](cudaStream_t st) {
        const long total = M * 4;
        const int grid = (int)std::min<long>((total + 255) / 256, 148L * 32);
        stem_im2col_kernel<<<grid, 256, 0, st>>>(plp->in_imgs, B, S, S, OH, OH, st_, patches);
        VSB_CUDA(cudaGetLastError());
      }, 1, "cnx.stem_im2col"});
      ConvGemmOp op; setup_tma_gemm(op, patches, M, 64, 64);
      op.p.epi = EPI_LN; op.p.act = ACT_NONE; op.p.bias = stem_b; op.p.ln_w = stem_lnw; op.p.ln_b = stem_lnb; op.p.ln_eps = 1e-6f;
      op.p.out32 = x; op.p.ld_out32 = Cp;
      ConvW sw; sw.w = stem_w16; sw.N = C; sw.K = 64; sw.bias = stem_b;
      add_conv(pl, op, sw, "cnx.stem_gemm." + std::to_string(C) + "@" + std::to_string(hs));
      dbg(pl, "ds0", x, 0, B, hs, hs, C, Cp);
    }
    int maxK4 = 0;
    for (int s = 0; s < 4; ++s) maxK4 = std::max(maxK4, 4 * d.ext_dims[s]);
    for (int s = 1; s < 4; ++s) VSB_CHECK(d.ext_dims[s] >= d.ext_dims[s - 1], "extractor widths must not shrink along the trunk");
    // GRN statistics ping-pong: block j's pwconv1 accumulates into stats_pp[j&1]; its apply pass clears stats_pp[(j+1)&1] over
    // B*K_j floats, which covers what block j-1 left there because the widths never shrink along the trunk.  Both buffers are
    // zeroed at the start of every run (the last block of the previous run leaves its statistics behind).
    const size_t stats_n = (size_t)B * maxK4;
    float* stats_all = pl.pool.alloc_n<float>(2 * stats_n);
    float* stats_pp[2] = {stats_all, stats_all + stats_n};
    int blk_counter = 0;
    // Deterministic mode (every stage of the tiny trunk): rows per sample a multiple of 32 -> the pwconv1 epilogue
    // warps STORE per-warp partial column sums ([M/32][4C]) and the GRN kernels add them in a fixed order: no atomics, no clearing,
    // bit-reproducible logits.  Otherwise (chunkyseal's odd maps) float atomics into [B][4C] as before, cleared here and by the
    // GRN kernels.
    bool all_part = true;
    size_t part_n = 0;
    {
      int h2 = hs;
      for (int s = 0; s < 4; ++s) {
        if (s > 0) h2 = (h2 - 2) / 2 + 1;
        const long rps = (long)h2 * h2, Ms = (long)B * rps;
        if (rps % 32 != 0) all_part = false;      // every epilogue warp (32 rows) then lies inside one sample, also in a ragged last tile
        part_n = std::max(part_n, (size_t)((Ms + kBlockM - 1) / kBlockM * 4) * 4 * d.ext_dims[s]);
      }
    }
    static const bool no_part = getenv("VSB_GRN_ATOMIC") != nullptr;
    if (no_part) all_part = false;
    float* part_buf = all_part ? pl.pool.alloc_n<float>(part_n) : nullptr;
    if (!all_part)
      pl.steps.push_back(Step{[=](cudaStream_t st) { VSB_CUDA(cudaMemsetAsync(stats_all, 0, 2 * stats_n * sizeof(float), st)); }, 0,
                              "cnx.grn_clear"});
    __half* x16 = nullptr;
    for (int s = 0; s < 4; ++s) {
      if (s > 0) {
        const int Cv = d.ext_dims[s - 1], Cvp = Cp, Cn = d.ext_dims[s], Cnp = pitch(Cn);
        const long Mp = (long)B * hs * hs;
        __half* xn = pl.pool.alloc_n<__half>(Mp * Cvp);
        if (Cvp != Cv) {   // pad channels stay zero: ln_rows writes [0, Cv)
          VSB_CUDA(cudaMemset(xn, 0, (size_t)Mp * Cvp * sizeof(__half)));
          VSB_CUDA(cudaDeviceSynchronize());   // the plan runs on the caller's stream, which need not be ordered after stream 0
        }
        {
          float *lw = ds[s - 1].lnw, *lb = ds[s - 1].lnb;
          const float* xin = x;
          pl.steps.push_back(Step{[=](cudaStream_t st) {
            const int grid = (int)std::min<long>((Mp + 7) / 8, 148L * 16);
            ln_rows_kernel<<<grid, 256, 0, st>>>(xin, Mp, Cv, Cvp, lw, lb, 1e-6f, xn, Cvp);
            VSB_CUDA(cudaGetLastError());
          }, 1, "cnx.ln_rows"});
        }
        const int ho = (hs - 2) / 2 + 1;   // odd maps (127, 63, 31): the last row / column is dropped, like nn.Conv2d k2 s2
        float* xo = pl.pool.alloc_n<float>((size_t)B * ho * ho * Cnp);
        ConvGemmOp op;
        // K order (r, s, c) with c over the PADDED width: matches pack_conv(..., pad_taps)
        setup_gather_conv(op, LD_GATHER_CONV, xn, Cvp, Cvp, nullptr, 0, 0, B, hs, hs, ho, ho, 2, 2, 2, 0, 0);
        op.p.epi = EPI_AFFINE; op.p.act = ACT_NONE; op.p.bias = ds[s - 1].conv.bias; op.p.out32 = xo; op.p.ld_out32 = Cnp;
        add_conv(pl, op, ds[s - 1].conv, "cnx.down2x2s2." + std::to_string(Cv) + "-" + std::to_string(Cn) + "@" + std::to_string(ho));
        x = xo; hs = ho; C = Cn; Cp = Cnp;
        dbg(pl, "ds" + std::to_string(s), x, 0, B, hs, hs, C, Cp);
      }
      const long M = (long)B * hs * hs;
      const int rows_per_sample = hs * hs;
      __half* a = pl.pool.alloc_n<__half>(M * Cp);
      __half* g = pl.pool.alloc_n<__half>(M * 4 * C);   // 4C is a multiple of 8 for every even width
      // GRN multiplier: folded into per-sample pwconv2 weights where a sample has many more rows than W2 has output
      // channels (stages 0-1 of the tiny trunk), applied to g in place otherwise
      static const bool no_wscale = getenv("VSB_NO_WSCALE") != nullptr;
      const bool wscale = !no_wscale && 4 * C <= rows_per_sample && rows_per_sample % kBlockM == 0 && C % 16 == 0;
      __half* w2s = wscale ? pl.pool.alloc_n<__half>((size_t)B * C * 4 * C) : nullptr;
      for (int jb = 0; jb < d.ext_depths[s]; ++jb) {
        const CnBlockW& w = cn[s][jb];
        {
          const int H = hs, Cc = C, ldc = Cp;
          const float* xin = x;
          float *dww = w.dww, *dwb = w.dwb, *lw = w.lnw, *lb = w.lnb;
          static const bool no_ring = getenv("VSB_DW_NO_RING") != nullptr;
          if (!no_ring && dw_ring_ok(Cc, H, ldc)) {
            // TMA-ring register-rolling kernel (dwconv_ring.cuh) for the 96 / 192 / 384-channel stages of the tiny trunk
            DwRingOp rop;
            setup_dw_ring(rop, xin, B, H, Cc);
            rop.dww = dww; rop.dwb = dwb; rop.lnw = lw; rop.lnb = lb; rop.out = a;
            pl.steps.push_back(Step{[rop](cudaStream_t st) { launch_dw_ring(rop, st); }, 1,
                                    "cnx.dwconv7_ln." + std::to_string(C) + "@" + std::to_string(hs)});
          } else
          pl.steps.push_back(Step{[=](cudaStream_t st) {
            if (!(Cc == 96 || Cc == 192 || Cc == 384 || Cc == 768) || getenv("VSB_DW_WIDE")) {
              // any even width / odd map size (chunkyseal's proportional trunk): threads loop over the channel pairs
              const long nstrips = (long)B * H * ((H + kDwStrip - 1) / kDwStrip);
              const int threads = std::min(512, (Cc / 2 + 31) / 32 * 32);
              const size_t smem_w = (size_t)kDwStrip * Cc * sizeof(float);
              static bool attr = false;
              if (!attr) { VSB_CUDA(cudaFuncSetAttribute(dwconv7_ln_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
              if (smem_w > 200 * 1024) throw Error("dwconv7: more than 6400 channels is not implemented", kErrUnsupported);
              dwconv7_ln_wide_kernel<<<(unsigned)nstrips, threads, smem_w, st>>>(xin, B, H, H, Cc, ldc, dww, dwb, lw, lb, a, ldc);
              VSB_CUDA(cudaGetLastError());
              return;
            }
            const int strips = (H + kDwStrip - 1) / kDwStrip;
            const long nstrips = (long)B * H * strips;
            const int C2 = Cc / 2;
            int spb = 1;
            for (int s2 = 1; s2 <= 4; s2 *= 2) if ((s2 * C2) % 32 == 0 && s2 * C2 <= 512) { spb = s2; break; }
            if (spb * C2 > 512) throw Error("dwconv7: C/2 > 512 threads is not implemented (unsupported chunky extractor width)", kErrUnsupported);
#define VSB_DWC(CC, ST, FR, SPB, NGR, RPB) dwconv7_ln_c_kernel<CC, ST, FR><<<(unsigned)(((NGR) + (SPB) - 1) / (SPB)), (SPB) * (CC) / 2, \
    (size_t)(SPB) * (ST) * (CC) * sizeof(float), st>>>(xin, B, H, H, dww, dwb, lw, lb, a, Cc, SPB, (int)(NGR), RPB)
            {
              // rows per block: 1.  Walking 2-4 consecutive rows per block (L1 reuse of the 7-row input window) was measured
              // SLOWER (96@64: 148 -> 228 us, 192@32: 58 -> 87 us): the kernel is latency-bound and fewer, longer blocks with
              // two block barriers per row hide less of it; VSB_DW_RPB re-enables the experiment
              int rpb = 1;
              if (const char* e = getenv("VSB_DW_RPB")) { const int r2 = atoi(e); if (r2 > 0 && H % r2 == 0) rpb = r2; }
              // narrow maps: one strip = the whole row, so no column is ever out of the image
              const long nrows = (long)B * H;
              if (Cc == 384 && H == 16) VSB_DWC(384, 16, true, 1, nrows, 1);
              else if (Cc == 768 && H == 8) VSB_DWC(768, 8, true, 1, nrows, 1);
              else if (Cc == 96) VSB_DWC(96, 8, false, spb, nstrips / rpb, rpb);
              else if (Cc == 192) VSB_DWC(192, 8, false, spb, nstrips / rpb, rpb);
              else if (Cc == 384) VSB_DWC(384, 8, false, spb, nstrips / rpb, rpb);
              else VSB_DWC(768, 8, false, spb, nstrips / rpb, rpb);
            }
#undef VSB_DWC
            VSB_CUDA(cudaGetLastError());
          }, 1, "cnx.dwconv7_ln." + std::to_string(C) + "@" + std::to_string(hs)});
        }
        {
          ConvGemmOp op; setup_tma_gemm(op, a, M, C, Cp);
          op.p.epi = EPI_AFFINE; op.p.act = ACT_GELU; op.p.bias = w.pw1.bias; op.p.out16 = g; op.p.ld_out16 = 4 * C;
          float* stats = all_part ? part_buf : stats_pp[blk_counter & 1];
          ++blk_counter;
          op.p.grn_stats = stats; op.p.rows_per_sample = rows_per_sample; op.p.grn_part = all_part ? 1 : 0;
          add_conv(pl, op, w.pw1, "cnx.pwconv1." + std::to_string(C) + "@" + std::to_string(hs));
        }
        {
          const int K4 = 4 * C;
          float* gamma = w.gamma;
          float* stats = all_part ? part_buf : stats_pp[(blk_counter - 1) & 1];
          float* stats_next = all_part ? nullptr : stats_pp[blk_counter & 1];
          const int part_rows = all_part ? rows_per_sample / 32 : 0;
          if (wscale) {
            // GRN multiplier folded into per-sample copies of W2 (pointwise.cuh K2e): no pass over g at all
            const __half* w2 = w.pw2.w;
            __half* w2s_ = w2s;
            const int Nn = C;
            pl.steps.push_back(Step{[=](cudaStream_t st) {
              static bool attr = false;
              if (!attr) { VSB_CUDA(cudaFuncSetAttribute(grn_scale_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
              grn_scale_weights_kernel<<<dim3(B, 4), 256, 2 * K4 * sizeof(float), st>>>(stats, stats_next, gamma, w2, w2s_, Nn, K4, part_rows);
              VSB_CUDA(cudaGetLastError());
            }, 1, "cnx.grn_wscale." + std::to_string(C) + "@" + std::to_string(hs)});
          } else {
          pl.steps.push_back(Step{[=](cudaStream_t st) {
            // enough row-slabs per sample to fill the GPU (~8 blocks per SM)
            static bool attr = false;
            if (!attr) { VSB_CUDA(cudaFuncSetAttribute(grn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
            int slabs = std::max(1, std::min(rows_per_sample, (148 * 8 + B - 1) / B));
            grn_apply_kernel<<<B * slabs, 256, 2 * K4 * sizeof(float), st>>>(g, rows_per_sample, K4, K4, stats, stats_next, gamma, slabs, part_rows);
            VSB_CUDA(cudaGetLastError());
          }, 1, "cnx.grn_apply." + std::to_string(C) + "@" + std::to_string(hs)});
          }
        }
        {
          ConvGemmOp op; setup_tma_gemm(op, g, M, 4 * C, 4 * C);
          op.p.epi = EPI_AFFINE; op.p.act = ACT_NONE; op.p.bias = w.pw2.bias; op.p.resid32 = x; op.p.ld_res32 = Cp;
          op.p.out32 = x; op.p.ld_out32 = Cp;  // in place: each element is read and written by the same thread
          if (wscale) { op.w_samples = B; op.p.rows_per_sample = rows_per_sample; }
          if (s == 3 && jb == d.ext_depths[s] - 1) {
            x16 = pl.pool.alloc_n<__half>(M * Cp);
            if (Cp != C) {   // the head conv's gather reads the pad channels
              VSB_CUDA(cudaMemset(x16, 0, (size_t)M * Cp * sizeof(__half)));
              VSB_CUDA(cudaDeviceSynchronize());
            }
            op.p.out16 = x16; op.p.ld_out16 = Cp;
          }
          add_conv(pl, op, w.pw2, "cnx.pwconv2." + std::to_string(C) + "@" + std::to_string(hs), 0, wscale ? w2s : nullptr);
        }
        if (jb == 0) { dbg(pl, "s" + std::to_string(s) + "b0_a", a, 1, B, hs, hs, C, Cp); dbg(pl, "s" + std::to_string(s) + "b0_g", g, 1, B, hs, hs, 4 * C, 4 * C); }
      }
      dbg(pl, "stage" + std::to_string(s), x, 0, B, hs, hs, C, Cp);
    }
    // ---- head (pixel_decoder.py:61-83)
    {
      const long M = (long)B * hs * hs;
      float* y = pl.pool.alloc_n<float>(M * Cp);
      ConvGemmOp op;
      setup_gather_conv(op, LD_GATHER_CONV, x16, Cp, Cp, nullptr, 0, 0, B, hs, hs, hs, hs, 3, 3, 1, 1, 1);
      op.p.epi = EPI_AFFINE; op.p.act = ACT_NONE; op.p.out32 = y; op.p.ld_out32 = Cp;
      add_conv(pl, op, head_conv, "cnx.head3x3." + std::to_string(C) + "@" + std::to_string(hs));
      dbg(pl, "head_conv", y, 0, B, hs, hs, C, Cp);
      float* pooled = pl.pool.alloc_n<float>((size_t)B * C);
      const int P = hs * hs, Cc = C, ldy = Cp, NO = 1 + d.nbits;
      float *lw = head_lnw, *lb = head_lnb, *hw = head_lw, *hb = head_lb;
      float* logits = pl.logits;
      pl.steps.push_back(Step{[=](cudaStream_t st) {
        static bool attr = false;
        if (!attr) { VSB_CUDA(cudaFuncSetAttribute(head_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
        head_pool_kernel<<<B, 256, (size_t)8 * Cc * sizeof(float), st>>>(y, P, Cc, ldy, lw, lb, pooled);
        const int grid = (int)(((long)B * NO + 7) / 8);
        head_linear_kernel<<<grid, 256, 0, st>>>(pooled, hw, hb, B, Cc, NO, logits);
        VSB_CUDA(cudaGetLastError());
      }, 2, "cnx.head_tail"});
    }
    Plan* ret = up.get();
    ret->last_use = ++plan_clock;
    detect_plans[B] = std::move(up);
    evict_plans(detect_plans, ret);
    return ret;
  }

  ResampleDev* get_resampler(int ih, int iw, int oh, int ow, bool aa) {
    const uint64_t key = ((uint64_t)ih << 48) ^ ((uint64_t)iw << 32) ^ ((uint64_t)oh << 16) ^ ((uint64_t)ow << 1) ^ (aa ? 1 : 0);
    auto it = resamplers.find(key);
    if (it != resamplers.end()) return it->second.get();
    std::unique_ptr<ResampleDev> r(new ResampleDev(ih, iw, oh, ow, aa));
    ResampleDev* ret = r.get();
    resamplers[key] = std::move(r);
    return ret;
  }

  // per-call CUDA-event scope for the launches outside the plans (bench.py roofline leg: vsb_profile_enable)
  template <class Fn> static void prof_scope(const std::string& name, cudaStream_t st, int launches, Fn&& fn) {
    if (!g_profile.enabled) { fn(); g_launches += launches; return; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    fn();
    cudaEventRecord(b, st);
    g_launches += launches;
    g_profile.pending.push_back({name, {a, b}});
  }

  // resample `n` frames that are `frame_stride` floats apart to S x S (one launch, also for strided key frames)
  void resize_frames(const float* src, long frame_stride, int n, int H, int W, float* dst, bool aa, cudaStream_t st) {
    const int S = d.img_size;
    ResampleDev* r = get_resampler(H, W, S, S, aa);
    const int vec = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (frame_stride % 4 == 0);
    if (n > 16384) {   // gridDim.y limit: split very long frame ranges
      for (int i = 0; i < n; i += 16384)
        resize_frames(src + (long)i * frame_stride, frame_stride, std::min(16384, n - i), H, W, dst + (size_t)i * 3 * S * S, aa, st);
      return;
    }
    const dim3 grid((S + r->toy - 1) / r->toy, n * 3);
    prof_scope("pw.resize." + std::to_string(H) + "x" + std::to_string(W) + "@" + std::to_string(n), st, 1, [&] {
      const int mt = r->tab.maxt_x <= 2 ? 2 : (r->tab.maxt_x <= 8 ? 8 : 0);
#define VSB_RS(MT, GS) do { static bool attr_ = false; \
        if (!attr_) { VSB_CUDA(cudaFuncSetAttribute(resize_sep_kernel<MT, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_ = true; } \
        resize_sep_kernel<MT, GS><<<grid, 256, r->smem, st>>>(src, frame_stride, dst, H, W, S, S, r->tab, r->toy, r->rin_max, vec); } while (0)
      if (r->gstage == 4) { if (mt == 2) VSB_RS(2, 4); else if (mt == 8) VSB_RS(8, 4); else VSB_RS(0, 4); }
      else if (r->gstage == 2) { if (mt == 2) VSB_RS(2, 2); else if (mt == 8) VSB_RS(8, 2); else VSB_RS(0, 2); }
      else { if (mt == 2) VSB_RS(2, 1); else if (mt == 8) VSB_RS(8, 1); else VSB_RS(0, 1); }
#undef VSB_RS
    });
    VSB_CUDA(cudaGetLastError());
  }

  void check_ready() const { if (!finalized) throw Error("model not finalized", kErrState); }

  // ---- streaming host path: embed + detect of HOST frames with the PCIe copies overlapped with compute
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_cmp;
  void embed_detect_host(const float* imgs_h, const uint8_t* msgs_h, int n_msgs, float* imgs_w_h, float* logits_h, int F, int H, int W,
                         int step, int video_mode, int chunk_keys, float scaling_i, float scaling_w, int flags) {
    check_ready();
    if (!s_in) {
      VSB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
      VSB_CUDA(cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
      VSB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    }
    const size_t fpx = (size_t)3 * H * W;
    const int NO = 1 + d.nbits;
    float* imgs = (float*)stage(0, (size_t)F * fpx * sizeof(float));
    float* out = (float*)stage(1, (size_t)F * fpx * sizeof(float));
    float* lg = (float*)stage(2, (size_t)F * NO * sizeof(float));
    uint8_t* msgs = (uint8_t*)stage(3, (size_t)n_msgs * d.nbits + 256);
    // Pipeline (three streams): frames travel in SLICES (a short first one, then ~48; multiples of `step`, and of whole reference chunks in
    // 'interpolate' mode where neighbouring keys of one chunk are mixed, so that key-frame groups stay whole); embed() runs per
    // slice as soon as its copy has landed, the D2H of a slice's watermarked frames starts as soon as its embed() is done and
    // runs under the following compute; detect() runs once 64 frames have accumulated (full-size extractor batches) and on the tail.
    // Exposed copies: the first slice in and the last group's logits out.
    // Slice sizes: the FIRST slice's copy is the exposed one, so it may be shorter than the rest (VSB_E2E_SL0 / VSB_E2E_SL1, read per
    // call so that one process can compare settings).  Measured at 64 x 3x256x256 (tests/e2e_ab.py, results bit-identical for every
    // setting): 32/32 7 165 f/s, 24/40 7 207, 16/48 7 317, 8/56 7 310, 16/24 7 020 -> defaults 16 / 48.
    const int unit = (video_mode == VSB_VIDEO_INTERPOLATE && step > 1) ? step * std::max(1, chunk_keys) : step;
    auto knob = [&](const char* name, int dflt) {
      const char* e = getenv(name);
      int v = e ? atoi(e) : dflt;
      if (v < 1 || v > kMaxBatch) v = dflt;
      return ((v + unit - 1) / unit) * unit;
    };
    const int sl0 = knob("VSB_E2E_SL0", 16), sl1 = knob("VSB_E2E_SL1", 48);
    std::vector<std::pair<int, int>> slices;   // (first frame, frames)
    for (int f0 = 0; f0 < F;) {
      const int n = std::min(slices.empty() ? sl0 : sl1, F - f0);
      slices.push_back({f0, n});
      f0 += n;
    }
    const int nsl = (int)slices.size();
    while ((int)ev_in.size() < nsl + 1) {
      cudaEvent_t a, b;
      VSB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      VSB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
      ev_in.push_back(a); ev_cmp.push_back(b);
    }
    VSB_CUDA(cudaMemcpyAsync(msgs, msgs_h, (size_t)n_msgs * d.nbits, cudaMemcpyHostToDevice, s_in));
    for (int k = 0; k < nsl; ++k) {       // all input copies are queued up front: s_in runs ahead of the compute
      const int f0 = slices[k].first, n = slices[k].second;
      VSB_CUDA(cudaMemcpyAsync(imgs + (size_t)f0 * fpx, imgs_h + (size_t)f0 * fpx, (size_t)n * fpx * sizeof(float), cudaMemcpyHostToDevice, s_in));
      VSB_CUDA(cudaEventRecord(ev_in[k], s_in));
    }
    int g0 = 0;                           // first frame not yet detected
    for (int k = 0; k < nsl; ++k) {
      const int f0 = slices[k].first, n = slices[k].second;
      VSB_CUDA(cudaStreamWaitEvent(s_cmp, ev_in[k], 0));
      const uint8_t* mk = msgs + (n_msgs == 1 ? 0 : (size_t)f0 * d.nbits);
      embed(imgs + (size_t)f0 * fpx, mk, n_msgs == 1 ? 1 : n, out + (size_t)f0 * fpx, nullptr, n, H, W, step, video_mode, chunk_keys, scaling_i,
            scaling_w, flags, s_cmp);
      VSB_CUDA(cudaEventRecord(ev_cmp[k], s_cmp));
      VSB_CUDA(cudaStreamWaitEvent(s_out, ev_cmp[k], 0));
      VSB_CUDA(cudaMemcpyAsync(imgs_w_h + (size_t)f0 * fpx, out + (size_t)f0 * fpx, (size_t)n * fpx * sizeof(float), cudaMemcpyDeviceToHost, s_out));
      if (f0 + n - g0 >= kMaxBatch || k == nsl - 1) {   // a full extractor batch (or the tail) is complete: detect its frames
        const int gn = f0 + n - g0;
        detect(out + (size_t)g0 * fpx, lg + (size_t)g0 * NO, gn, H, W, flags & VSB_FLAG_RESIZE_NO_AA, s_cmp);
        VSB_CUDA(cudaEventRecord(ev_cmp[nsl], s_cmp));
        VSB_CUDA(cudaStreamWaitEvent(s_out, ev_cmp[nsl], 0));
        VSB_CUDA(cudaMemcpyAsync(logits_h + (size_t)g0 * NO, lg + (size_t)g0 * NO, (size_t)gn * NO * sizeof(float), cudaMemcpyDeviceToHost, s_out));
        g0 = f0 + n;
      }
    }
    VSB_CUDA(cudaStreamSynchronize(s_out));
  }

  // ---- the same streaming scheme with RGB24 frames [F,H,W,3] uint8 on the host (inference_streaming.py:23-33,116-124):
  // out_h == nullptr -> detect only (logits of the input frames); logits_h == nullptr -> embed only; both -> the logits are
  // those of the re-quantised watermarked frames, which is what the reference's detect pass reads back from the encoder
  void frames_host_u8(const uint8_t* in_h, const uint8_t* msgs_h, int n_msgs, uint8_t* out_h, float* logits_h, int F, int H, int W,
                      int step, int video_mode, int chunk_keys, float scaling_i, float scaling_w, int flags) {
    check_ready();
    if (!s_in) {
      VSB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
      VSB_CUDA(cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
      VSB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    }
    const bool do_embed = out_h != nullptr, do_detect = logits_h != nullptr;
    VSB_CHECK(do_embed || do_detect, "nothing to do");
    const size_t fpx = (size_t)3 * H * W;
    const long plane = (long)H * W;
    const int NO = 1 + d.nbits;
    float* imgs = (float*)stage(0, (size_t)F * fpx * sizeof(float));
    float* out = do_embed ? (float*)stage(1, (size_t)F * fpx * sizeof(float)) : nullptr;
    float* lg = do_detect ? (float*)stage(2, (size_t)F * NO * sizeof(float)) : nullptr;
    uint8_t* msgs = do_embed ? (uint8_t*)stage(3, (size_t)n_msgs * d.nbits + 256) : nullptr;
    uint8_t* in8 = (uint8_t*)stage(4, (size_t)F * fpx);
    uint8_t* out8 = do_embed ? (uint8_t*)stage(5, (size_t)F * fpx) : nullptr;
    int ch = kMaxBatch;   // one-byte samples: the exposed first/last copies are short, so full 64-frame batches win
    const int unit = (video_mode == VSB_VIDEO_INTERPOLATE && step > 1) ? step * std::max(1, chunk_keys) : step;
    if (ch % unit) ch = ((ch + unit - 1) / unit) * unit;
    const int nch = (F + ch - 1) / ch;
    while ((int)ev_in.size() < nch) {
      cudaEvent_t a, b;
      VSB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      VSB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
      ev_in.push_back(a); ev_cmp.push_back(b);
    }
    if (do_embed) VSB_CUDA(cudaMemcpyAsync(msgs, msgs_h, (size_t)n_msgs * d.nbits, cudaMemcpyHostToDevice, s_in));
    for (int k = 0; k < nch; ++k) {
      const int f0 = k * ch, n = std::min(ch, F - f0);
      VSB_CUDA(cudaMemcpyAsync(in8 + (size_t)f0 * fpx, in_h + (size_t)f0 * fpx, (size_t)n * fpx, cudaMemcpyHostToDevice, s_in));
      VSB_CUDA(cudaEventRecord(ev_in[k], s_in));
      VSB_CUDA(cudaStreamWaitEvent(s_cmp, ev_in[k], 0));
      const long npx = (long)n * plane;
      const unsigned blocks = (unsigned)std::min<long>((npx + 255) / 256, 148L * 16);
      u8hwc_to_f32chw_kernel<<<blocks, 256, 0, s_cmp>>>(in8 + (size_t)f0 * fpx, imgs + (size_t)f0 * fpx, npx, plane);
      g_launches += 1;
      const float* det_in = imgs + (size_t)f0 * fpx;
      if (do_embed) {
        const uint8_t* mk = msgs + (n_msgs == 1 ? 0 : (size_t)f0 * d.nbits);
        embed(imgs + (size_t)f0 * fpx, mk, n_msgs == 1 ? 1 : n, out + (size_t)f0 * fpx, nullptr, n, H, W, step, video_mode, chunk_keys,
              scaling_i, scaling_w, flags, s_cmp);
        f32chw_to_u8hwc_kernel<<<blocks, 256, 0, s_cmp>>>(out + (size_t)f0 * fpx, out8 + (size_t)f0 * fpx,
                                                          do_detect ? out + (size_t)f0 * fpx : nullptr, npx, plane);
        g_launches += 1;
        det_in = out + (size_t)f0 * fpx;
      }
      if (do_detect) detect(det_in, lg + (size_t)f0 * NO, n, H, W, flags & VSB_FLAG_RESIZE_NO_AA, s_cmp);
      VSB_CUDA(cudaGetLastError());
      VSB_CUDA(cudaEventRecord(ev_cmp[k], s_cmp));
      VSB_CUDA(cudaStreamWaitEvent(s_out, ev_cmp[k], 0));
      if (do_embed) VSB_CUDA(cudaMemcpyAsync(out_h + (size_t)f0 * fpx, out8 + (size_t)f0 * fpx, (size_t)n * fpx, cudaMemcpyDeviceToHost, s_out));
      if (do_detect) VSB_CUDA(cudaMemcpyAsync(logits_h + (size_t)f0 * NO, lg + (size_t)f0 * NO, (size_t)n * NO * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    }
    VSB_CUDA(cudaStreamSynchronize(s_out));
  }

  // U-Net on `n` processing-size RGB frames (contiguous [n,3,S,S]) -> plan->delta
  Plan* run_unet(const float* x, const uint8_t* msgs, int msg_stride, int n, cudaStream_t st) {
    Plan* pl = get_embed_plan(n);
    pl->in_imgs = x;
    pl->in_frame_stride = (long)3 * d.img_size * d.img_size;
    pl->in_msgs = msgs;
    pl->in_msg_stride = msg_stride;
    pl->run(st);
    last_plan = pl;
    return pl;
  }

  void embedder_forward(const float* x, const uint8_t* msgs, int n_msgs, float* delta, int B, cudaStream_t st) {
    check_ready();
    VSB_CHECK(n_msgs == 1 || n_msgs == B, "msgs must be [1,K] or [B,K]");
    const int S = d.img_size;
    for (int b0 = 0; b0 < B; b0 += kMaxBatch) {
      const int n = std::min(kMaxBatch, B - b0);
      Plan* pl = run_unet(x + (size_t)b0 * 3 * S * S, msgs + (n_msgs == 1 ? 0 : (size_t)b0 * d.nbits), n_msgs == 1 ? 0 : d.nbits, n, st);
      VSB_CUDA(cudaMemcpyAsync(delta + (size_t)b0 * d.unet_out_ch * S * S, pl->delta, (size_t)n * d.unet_out_ch * S * S * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    }
  }

  void embed(const float* imgs, const uint8_t* msgs, int n_msgs, float* imgs_w, float* preds_w, int F, int H, int W, int step,
             int video_mode, int chunk_keys, float scaling_i, float scaling_w, int flags, cudaStream_t st) {
    check_ready();
    VSB_CHECK(F > 0 && H > 0 && W > 0 && step >= 1, "bad embed shape");
    VSB_CHECK(n_msgs == 1 || (n_msgs == F && step == 1), "msgs must be [1,K], or [F,K] in image mode");
    // 'interpolate' mixes neighbouring key frames inside one reference chunk (videoseal.py:300-331 chunks by chunk_size
    // key frames; :101-113 interpolates within the chunk), so U-Net batches are whole chunks in that mode
    const bool interp = (video_mode == VSB_VIDEO_INTERPOLATE) && step > 1;
    if (interp) VSB_CHECK(chunk_keys >= 1 && chunk_keys <= kMaxBatch, "video_mode='interpolate' needs 1 <= chunk_size <= 64");
    const int kbatch = interp ? (kMaxBatch / chunk_keys) * chunk_keys : kMaxBatch;
    const int bint = interp ? chunk_keys : 0;
    const int S = d.img_size;
    const bool same = (H == S && W == S);
    const bool aa = !(flags & VSB_FLAG_RESIZE_NO_AA);
    const bool use_jnd = !(flags & VSB_FLAG_NO_ATTENUATION) && d.jnd_in_ch != 0;
    if (use_jnd && !((d.jnd_in_ch == 1 || d.jnd_in_ch == 3) && (d.jnd_out_ch == 1 || d.jnd_out_ch == 3)))
      throw Error("JND attenuation: in_channels / out_channels must be 1 or 3 (configs/attenuation.yaml)", kErrUnsupported);
    // channels of hmaps * preds_w (broadcast): what `preds_w` holds and what the low-resolution path hands to the blend
    const int PC = use_jnd ? std::max(d.unet_out_ch, d.jnd_out_ch) : d.unet_out_ch;
    const bool lowres = use_jnd && (flags & VSB_FLAG_LOWRES_ATTN);
    const long fstride = (long)3 * H * W;
    const int nkeys = (F + step - 1) / step;
    ResampleDev* up = same ? nullptr : get_resampler(S, S, H, W, aa);
    for (int k0 = 0; k0 < nkeys; k0 += kbatch) {
      const int nk = std::min(kbatch, nkeys - k0);
      const int f0 = k0 * step, f1 = std::min(F, (k0 + nk) * step);
      const int nf = f1 - f0;
      Plan* pl = get_embed_plan(nk);
      const float* x = imgs + (size_t)f0 * fstride;
      // low-resolution attenuation needs EVERY frame of the chunk at processing size (wam.py:177-180, videoseal.py:321-324):
      // resample them once into a grow-only scratch buffer of the model (no allocation or synchronisation on the steady-state
      // path) and take the key frames from there
      const float* frames_res = x;
      if (lowres && !same) {
        float* fr = (float*)stage(6, (size_t)nf * 3 * S * S * sizeof(float));
        resize_frames(x, fstride, nf, H, W, fr, aa, st);
        frames_res = fr;
      }
      if (!same || step != 1) {
        // key frames -> contiguous processing-size RGB batch
        if (same || lowres) {
          VSB_CUDA(cudaMemcpy2DAsync(pl->x_res, (size_t)3 * S * S * sizeof(float), frames_res, (size_t)step * 3 * S * S * sizeof(float),
                                     (size_t)3 * S * S * sizeof(float), nk, cudaMemcpyDeviceToDevice, st));
        } else {
          resize_frames(x, (long)step * fstride, nk, H, W, pl->x_res, aa, st);
        }
        x = pl->x_res;
      }
      run_unet(x, msgs + (n_msgs == 1 ? 0 : (size_t)k0 * d.nbits), n_msgs == 1 ? 0 : d.nbits, nk, st);
      const float* delta = pl->delta;
      const int balt = (video_mode == VSB_VIDEO_ALTERNATE) ? 1 : 0;
      BlendParams bp;
      memset(&bp, 0, sizeof(bp));
      bp.imgs = imgs + (size_t)f0 * fstride; bp.imgs_w = imgs_w + (size_t)f0 * fstride;
      bp.preds_w = preds_w ? preds_w + (size_t)f0 * PC * H * W : nullptr;
      bp.F = nf; bp.H = H; bp.W = W; bp.PH = S; bp.PW = S; bp.CD = d.unet_out_ch;
      bp.jnd_in = d.jnd_in_ch ? d.jnd_in_ch : 1; bp.jnd_out = d.jnd_out_ch ? d.jnd_out_ch : 1;
      bp.clamp = (flags & VSB_FLAG_CLAMP) ? 1 : 0; bp.identity_resample = same ? 1 : 0;
      bp.scaling_i = scaling_i; bp.scaling_w = scaling_w;
      if (up) bp.tab = up->tab;
      if (lowres) {
        // per-frame attenuated delta at processing size: out[f] = hmap(frames_res[f]) * delta(keys of f); then a plain blend
        float* lowres_buf = (float*)stage(7, (size_t)nf * PC * S * S * sizeof(float));
        VSB_CHECK(nf <= 65535, "too many frames per key-frame batch");
        dim3 grid((S + kBlendTW - 1) / kBlendTW, (S + kBlendTH - 1) / kBlendTH, nf);
        const int CD = d.unet_out_ch;
        prof_scope("pw.jnd_lowres@" + std::to_string(nf), st, 1,
                   [&] { jnd_lowres_kernel<<<grid, 256, 0, st>>>(frames_res, delta, lowres_buf, S, S, CD, step, balt, bint, bp.jnd_in, bp.jnd_out); });
        VSB_CUDA(cudaGetLastError());
        bp.delta = lowres_buf; bp.step = 1; bp.alternate = 0; bp.interp_chunk = 0; bp.use_jnd = 0; bp.CD = PC;
      } else {
        bp.delta = delta; bp.step = step; bp.alternate = balt; bp.interp_chunk = bint; bp.use_jnd = use_jnd ? 1 : 0;
      }
      launch_blend(bp, st);
    }
  }

  static void launch_blend(const BlendParams& bp, cudaStream_t st) {
    VSB_CHECK(bp.F <= 65535, "too many frames per blend launch");
    dim3 grid((bp.W + kB2TW - 1) / kB2TW, (bp.H + kB2TH - 1) / kB2TH, bp.F);
    const bool vec = (bp.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(bp.imgs) | reinterpret_cast<uintptr_t>(bp.imgs_w) |
                                          reinterpret_cast<uintptr_t>(bp.preds_w)) % 16 == 0);
    // delta read in place, or through <= 2 taps per axis (plain bilinear up-scale: every input at least as large as the processing size)
    const bool fastup = bp.identity_resample || (bp.tab.maxt_x <= 2 && bp.tab.maxt_y <= 2 && bp.H >= bp.PH && bp.W >= bp.PW);
    VSB_CHECK(bp.CD == 1 || bp.CD == 3, "delta must have 1 or 3 channels");
    // TMA-fed kernel (jnd_blend3_kernel): vector layout, delta in place or <= 2 taps per axis; the delta tile needs a TMA-legal row pitch
    static const bool blend_old = getenv("VSB_BLEND_OLD") != nullptr;
    const bool delta_tma = bp.identity_resample || (bp.PW % 4 == 0 && reinterpret_cast<uintptr_t>(bp.delta) % 16 == 0);
    // (boxes never exceed the tensor they are cut from: tiny frames / processing sizes take the load-then-compute kernel)
    const bool tma = vec && fastup && delta_tma && !blend_old && bp.W >= kB2LP && bp.H >= kB2TH + 4 &&
                     (bp.identity_resample || (bp.PW >= kB3DW && bp.PH >= kB2DH)) &&
                     (!bp.identity_resample || (bp.PW % 4 == 0 && reinterpret_cast<uintptr_t>(bp.delta) % 16 == 0));
    CUtensorMap tmI, tmD;
    if (tma) {
      const uint64_t idims[3] = {(uint64_t)bp.W, (uint64_t)bp.H, (uint64_t)bp.F * 3};
      const uint64_t istr[2] = {(uint64_t)bp.W * 4, (uint64_t)bp.H * bp.W * 4};
      const uint32_t ibox[3] = {(uint32_t)kB2LP, (uint32_t)(kB2TH + 4), 3};
      encode_map(&tmI, bp.imgs, 3, idims, istr, ibox, 0, true, /*f32=*/true);
      tmD = tmI;
      if (!bp.identity_resample) {
        const int nkeys = (bp.F + bp.step - 1) / bp.step;
        const uint64_t ddims[3] = {(uint64_t)bp.PW, (uint64_t)bp.PH, (uint64_t)nkeys * bp.CD};
        const uint64_t dstr[2] = {(uint64_t)bp.PW * 4, (uint64_t)bp.PH * bp.PW * 4};
        const uint32_t dbox[3] = {(uint32_t)kB3DW, (uint32_t)kB2DH, 1};
        encode_map(&tmD, bp.delta, 3, ddims, dstr, dbox, 0, true, /*f32=*/true);
      }
    }
    prof_scope(std::string(bp.use_jnd ? "pw.jnd_blend." : "pw.blend.") + std::to_string(bp.H) + "x" + std::to_string(bp.W) + "@" + std::to_string(bp.F) +
                   (bp.preds_w ? "+preds" : ""), st, 1, [&] {
#define VSB_BL(V, FU, CDV, JI) do { static bool attr_ = false; \
        if (!attr_) { VSB_CUDA(cudaFuncSetAttribute(jnd_blend2_kernel<V, FU, CDV, JI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b2_smem(JI))); attr_ = true; } \
        jnd_blend2_kernel<V, FU, CDV, JI><<<grid, 256, b2_smem(JI), st>>>(bp); } while (0)
#define VSB_BL2(CDV, JI) do { \
        if (vec) { if (fastup) VSB_BL(4, 1, CDV, JI); else VSB_BL(4, 0, CDV, JI); } else { if (fastup) VSB_BL(1, 1, CDV, JI); else VSB_BL(1, 0, CDV, JI); } } while (0)
#define VSB_BL3(CDV, JI) do { static bool attr_ = false; \
        if (!attr_) { VSB_CUDA(cudaFuncSetAttribute(jnd_blend3_kernel<CDV, JI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b3_smem(JI))); attr_ = true; } \
        jnd_blend3_kernel<CDV, JI><<<grid3, 256, b3_smem(JI), st>>>(bp, tmI, tmD); } while (0)
      const dim3 grid3(grid.x, (bp.H + kB2TH * kB3NT - 1) / (kB2TH * kB3NT), bp.F);
      const bool j3 = bp.use_jnd && bp.jnd_in == 3;      // one heat-map per RGB channel (jnd_3_1 / jnd_3_3): rarely used, own instantiation
      if (tma) {
        if (bp.CD == 1) { if (j3) VSB_BL3(1, 3); else VSB_BL3(1, 1); }
        else            { if (j3) VSB_BL3(3, 3); else VSB_BL3(3, 1); }
      } else if (bp.CD == 1) { if (j3) VSB_BL2(1, 3); else VSB_BL2(1, 1); }
      else                   { if (j3) VSB_BL2(3, 3); else VSB_BL2(3, 1); }
#undef VSB_BL3
#undef VSB_BL2
#undef VSB_BL
    });
    VSB_CUDA(cudaGetLastError());
  }

  void jnd_heatmaps(const float* imgs, float* hmaps, int F, int H, int W, cudaStream_t st) {
    // heat-map only (operator seam, modules/jnd.py:80-108): jnd_lowres_kernel with delta == nullptr writes hmap * 1; hmaps [F, jnd_out, H, W]
    VSB_CHECK(F <= 65535, "too many frames");
    VSB_CHECK(d.jnd_in_ch != 0, "the card has no JND attenuation");
    dim3 grid((W + kBlendTW - 1) / kBlendTW, (H + kBlendTH - 1) / kBlendTH, F);
    jnd_lowres_kernel<<<grid, 256, 0, st>>>(imgs, nullptr, hmaps, H, W, 1, 1, 0, 0, d.jnd_in_ch, d.jnd_out_ch);
    g_launches += 1;
    VSB_CUDA(cudaGetLastError());
  }

  void detect(const float* imgs, float* logits, int F, int H, int W, int flags, cudaStream_t st) {
    check_ready();
    VSB_CHECK(F > 0 && H > 0 && W > 0, "bad detect shape");
    const int S = d.img_size;
    const bool same = (H == S && W == S);
    const bool aa = !(flags & VSB_FLAG_RESIZE_NO_AA);
    const long fstride = (long)3 * H * W;
    const int NO = 1 + d.nbits;
    for (int f0 = 0; f0 < F; f0 += kMaxBatch) {
      const int n = std::min(kMaxBatch, F - f0);
      Plan* pl = get_detect_plan(n);
      const float* x = imgs + (size_t)f0 * fstride;
      if (!same) { resize_frames(x, fstride, n, H, W, pl->x_res, aa, st); x = pl->x_res; }
      pl->in_imgs = x;
      pl->run(st);
      last_plan = pl;
      VSB_CUDA(cudaMemcpyAsync(logits + (size_t)f0 * NO, pl->logits, (size_t)n * NO * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
  }
};

}  // namespace vsb
