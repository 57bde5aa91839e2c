// libvsb200.so: the single translation unit behind include/vsb200.h.
//   nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC vsb200.cu -o libvsb200.so
#include "model.cuh"

using namespace vsb;

struct vsb_model {
  Model impl;
  explicit vsb_model(const vsb_model_desc& d) : impl(d) {}
};

static thread_local std::string g_err;

static int fail(const std::exception& e, int code) {
  g_err = e.what();
  return code;
}

static_assert((int)kErrInvalid == (int)VSB_ERR_INVALID && (int)kErrUnsupported == (int)VSB_ERR_UNSUPPORTED && (int)kErrCuda == (int)VSB_ERR_CUDA &&
                  (int)kErrState == (int)VSB_ERR_STATE,
              "status codes of include/vsb200.h");

// Every C entry point that touches the GPU runs on the model's device and restores the caller's current device afterwards (the
// library must not change process-wide state under PyTorch), and holds the model's mutex: plans, scratch buffers, staging areas
// and the launch / profile counters belong to the handle, so calls on ONE handle are serialised (use one handle per thread or
// stream for concurrency; include/vsb200.h "Threading").
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) { VSB_CUDA(cudaSetDevice(dev)); } else { prev = -1; }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define VSB_MODEL_SCOPE(m) std::lock_guard<std::recursive_mutex> _lk((m)->impl.mu); DeviceGuard _dg((m)->impl.device)

#define VSB_API_BEGIN try {
#define VSB_API_END                                                                       \
  }                                                                                       \
  catch (const vsb::Error& e) { return fail(e, e.code); }                                 \
  catch (const std::bad_alloc& e) { return fail(e, VSB_ERR_CUDA); }                       \
  catch (const std::exception& e) { return fail(e, VSB_ERR_INVALID); }

extern "C" {

const char* vsb_last_error(void) { return g_err.c_str(); }
int vsb_version(void) { return 100; }

int vsb_model_create(const vsb_model_desc* desc, vsb_model** out) {
  VSB_API_BEGIN
  VSB_CHECK(desc != nullptr && out != nullptr, "null argument");
  *out = new vsb_model(*desc);
  return VSB_OK;
  VSB_API_END
}

int vsb_model_set_tensor(vsb_model* m, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  VSB_API_BEGIN
  VSB_CHECK(m && name && data && (shape || ndim == 0) && ndim >= 0 && ndim <= 4, "bad tensor argument");
  if (m->impl.finalized) throw Error("model already finalized", kErrState);
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  t.data.assign(data, data + t.numel());
  m->impl.sd[name] = std::move(t);
  return VSB_OK;
  VSB_API_END
}

int vsb_model_finalize(vsb_model* m, int32_t device) {
  VSB_API_BEGIN
  VSB_CHECK(m != nullptr, "null model");
  std::lock_guard<std::recursive_mutex> lk(m->impl.mu);
  DeviceGuard dg(device);
  m->impl.finalize(device);
  return VSB_OK;
  VSB_API_END
}

void vsb_model_destroy(vsb_model* m) {
  if (m == nullptr) return;
  try { DeviceGuard dg(m->impl.device); delete m; } catch (...) {}
}

int vsb_embed(vsb_model* m, const float* imgs, const uint8_t* msgs, int32_t n_msgs, float* imgs_w, float* preds_w, int32_t F, int32_t H,
              int32_t W, int32_t step, int32_t video_mode, int32_t chunk_keys, float scaling_i, float scaling_w, int32_t flags,
              void* stream) {
  VSB_API_BEGIN
  VSB_CHECK(m && imgs && msgs && imgs_w, "null argument");
  VSB_MODEL_SCOPE(m);
  m->impl.embed(imgs, msgs, n_msgs, imgs_w, preds_w, F, H, W, step, video_mode, chunk_keys, scaling_i, scaling_w, flags,
                (cudaStream_t)stream);
  return VSB_OK;
  VSB_API_END
}

int vsb_embedder_forward(vsb_model* m, const float* x, const uint8_t* msgs, int32_t n_msgs, float* delta, int32_t B, void* stream) {
  VSB_API_BEGIN
  VSB_CHECK(m && x && msgs && delta && B > 0, "null argument");
  VSB_MODEL_SCOPE(m);
  m->impl.embedder_forward(x, msgs, n_msgs, delta, B, (cudaStream_t)stream);
  return VSB_OK;
  VSB_API_END
}

int vsb_detect(vsb_model* m, const float* imgs, float* logits, int32_t F, int32_t H, int32_t W, int32_t flags, void* stream) {
  VSB_API_BEGIN
  VSB_CHECK(m && imgs && logits, "null argument");
  VSB_MODEL_SCOPE(m);
  m->impl.detect(imgs, logits, F, H, W, flags, (cudaStream_t)stream);
  return VSB_OK;
  VSB_API_END
}

int vsb_jnd_heatmaps(vsb_model* m, const float* imgs, float* hmaps, int32_t F, int32_t H, int32_t W, void* stream) {
  VSB_API_BEGIN
  VSB_CHECK(m && imgs && hmaps, "null argument");
  VSB_MODEL_SCOPE(m);
  m->impl.jnd_heatmaps(imgs, hmaps, F, H, W, (cudaStream_t)stream);
  return VSB_OK;
  VSB_API_END
}

int vsb_embed_host(vsb_model* m, const float* imgs_h, const uint8_t* msgs_h, int32_t n_msgs, float* imgs_w_h, float* preds_w_h, int32_t F,
                   int32_t H, int32_t W, int32_t step, int32_t video_mode, int32_t chunk_keys, float scaling_i, float scaling_w,
                   int32_t flags) {
  VSB_API_BEGIN
  VSB_CHECK(m && imgs_h && msgs_h && imgs_w_h, "null argument");
  VSB_MODEL_SCOPE(m);
  m->impl.check_ready();
  const size_t n = (size_t)F * 3 * H * W;
  const bool use_jnd = !(flags & VSB_FLAG_NO_ATTENUATION) && m->impl.d.jnd_in_ch != 0;
  const size_t np = (size_t)F * (use_jnd ? std::max(m->impl.d.unet_out_ch, m->impl.d.jnd_out_ch) : m->impl.d.unet_out_ch) * H * W;
  float* imgs = (float*)m->impl.stage(0, n * sizeof(float));
  float* out = (float*)m->impl.stage(1, n * sizeof(float));
  float* pw = preds_w_h ? (float*)m->impl.stage(2, np * sizeof(float)) : nullptr;
  uint8_t* msgs = (uint8_t*)m->impl.stage(3, (size_t)n_msgs * m->impl.d.nbits + 256);
  cudaStream_t st = 0;
  VSB_CUDA(cudaMemcpyAsync(imgs, imgs_h, n * sizeof(float), cudaMemcpyHostToDevice, st));
  VSB_CUDA(cudaMemcpyAsync(msgs, msgs_h, (size_t)n_msgs * m->impl.d.nbits, cudaMemcpyHostToDevice, st));
  m->impl.embed(imgs, msgs, n_msgs, out, pw, F, H, W, step, video_mode, chunk_keys, scaling_i, scaling_w, flags, st);
  VSB_CUDA(cudaMemcpyAsync(imgs_w_h, out, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (pw) VSB_CUDA(cudaMemcpyAsync(preds_w_h, pw, np * sizeof(float), cudaMemcpyDeviceToHost, st));
  VSB_CUDA(cudaStreamSynchronize(st));
  return VSB_OK;
  VSB_API_END
}

int vsb_detect_host(vsb_model* m, const float* imgs_h, float* logits_h, int32_t F, int32_t H, int32_t W, int32_t flags) {
  VSB_API_BEGIN
  VSB_CHECK(m && imgs_h && logits_h, "null argument");
  VSB_MODEL_SCOPE(m);
  m->impl.check_ready();
  const size_t n = (size_t)F * 3 * H * W;
  const size_t nl = (size_t)F * (1 + m->impl.d.nbits);
  float* imgs = (float*)m->impl.stage(0, n * sizeof(float));
  float* lg = (float*)m->impl.stage(2, nl * sizeof(float));
  cudaStream_t st = 0;
  VSB_CUDA(cudaMemcpyAsync(imgs, imgs_h, n * sizeof(float), cudaMemcpyHostToDevice, st));
  m->impl.detect(imgs, lg, F, H, W, flags, st);
  VSB_CUDA(cudaMemcpyAsync(logits_h, lg, nl * sizeof(float), cudaMemcpyDeviceToHost, st));
  VSB_CUDA(cudaStreamSynchronize(st));
  return VSB_OK;
  VSB_API_END
}

int vsb_embed_detect_host(vsb_model* m, const float* imgs_h, const uint8_t* msgs_h, int32_t n_msgs, float* imgs_w_h, float* logits_h,
                          int32_t F, int32_t H, int32_t W, int32_t step, int32_t video_mode, int32_t chunk_keys, float scaling_i,
                          float scaling_w, int32_t flags) {
  VSB_API_BEGIN
  VSB_CHECK(m && imgs_h && msgs_h && imgs_w_h && logits_h, "null argument");
  VSB_MODEL_SCOPE(m);
  m->impl.embed_detect_host(imgs_h, msgs_h, n_msgs, imgs_w_h, logits_h, F, H, W, step, video_mode, chunk_keys, scaling_i, scaling_w,
                            flags);
  return VSB_OK;
  VSB_API_END
}

int vsb_frames_host_u8(vsb_model* m, const uint8_t* frames_h, const uint8_t* msgs_h, int32_t n_msgs, uint8_t* frames_w_h, float* logits_h,
                       int32_t F, int32_t H, int32_t W, int32_t step, int32_t video_mode, int32_t chunk_keys, float scaling_i,
                       float scaling_w, int32_t flags) {
  VSB_API_BEGIN
  VSB_CHECK(m && frames_h && (frames_w_h || logits_h) && (msgs_h || !frames_w_h), "null argument");
  VSB_MODEL_SCOPE(m);
  VSB_CHECK(F > 0 && H > 0 && W > 0 && step >= 1, "bad shape");
  m->impl.frames_host_u8(frames_h, msgs_h, n_msgs, frames_w_h, logits_h, F, H, W, step, video_mode, chunk_keys, scaling_i, scaling_w,
                         flags);
  return VSB_OK;
  VSB_API_END
}

int64_t vsb_launch_count(int32_t reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int64_t vsb_debug_get_tensor(vsb_model* m, const char* name, float* host_out, int64_t capacity, int64_t* shape4) {
  try {
    VSB_CHECK(m && name && shape4, "null argument");
    Plan* pl = m->impl.last_plan;
    VSB_CHECK(pl != nullptr, "no plan has run yet");
    VSB_MODEL_SCOPE(m);
    auto it = pl->dbg.find(name);
    if (it == pl->dbg.end()) throw Error(std::string("unknown debug tensor: ") + name);
    const DebugTensor& t = it->second;
    const int64_t rows = t.shape[0] * t.shape[1] * t.shape[2], C = t.shape[3];
    for (int i = 0; i < 4; ++i) shape4[i] = t.shape[i];
    const int64_t n = rows * C;
    if (host_out == nullptr) return n;
    VSB_CHECK(capacity >= n, "debug buffer too small");
    VSB_CUDA(cudaDeviceSynchronize());
    if (t.dtype == 0) {
      VSB_CUDA(cudaMemcpy2D(host_out, C * sizeof(float), t.ptr, t.ld * sizeof(float), C * sizeof(float), rows, cudaMemcpyDeviceToHost));
    } else {
      std::vector<__half> h((size_t)n);
      VSB_CUDA(cudaMemcpy2D(h.data(), C * sizeof(__half), t.ptr, t.ld * sizeof(__half), C * sizeof(__half), rows, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < n; ++i) host_out[i] = __half2float(h[i]);
    }
    return n;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

int vsb_debug_conv(const vsb_conv_test* t, void* stream) {
  VSB_API_BEGIN
  VSB_CHECK(t != nullptr, "null argument");
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    VSB_CUDA(cudaGetDevice(&dev));
    VSB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (t->loader == 5) {
    // direct 3x3 conv (conv3_direct.cuh): weights arrive as [N][9*C0] and are re-packed here
    VSB_CHECK(((t->R == 3 && t->S == 3 && t->pad == 1) || (t->R == 1 && t->S == 1 && t->pad == 0)) && t->stride == 1 && t->C1 == 0,
              "direct conv: 3x3 pad 1 or 1x1, stride 1");
    // (one scratch buffer, never freed: test-only path; the small pack kernel runs on every call)
    static __half* wpk = nullptr;
    if (wpk == nullptr) VSB_CUDA(cudaMalloc(&wpk, (size_t)64 * 9 * 64 * sizeof(__half)));
    pack_direct_weights((const __half*)t->weights, t->N, t->C0, wpk, (cudaStream_t)stream, t->R * t->S);
    Conv3DirectOp dop;
    setup_conv3_direct(dop, (const __half*)t->src0, t->B, t->IH, t->IW, t->C0, t->N, wpk, num_sms, t->R);
    dop.p.relu = t->act == ACT_RELU ? 1 : 0;
    dop.p.bias = t->bias; dop.p.resid = (const __half*)t->resid16; dop.p.out = (__half*)t->out16;
    dop.p.outc_w = t->outc_w; dop.p.outc_b = t->outc_b; dop.p.n_out = t->n_out; dop.p.delta = t->delta; dop.p.outc_tanh = 1;
    launch_direct(dop, (cudaStream_t)stream);
    g_launches += 1;
    return VSB_OK;
  }
  ConvGemmOp op;
  const int Ct = t->C0 + t->C1;
  int K = t->R * t->S * Ct;
  if (t->loader == LD_HALO_UPS || t->loader == LD_HALO_CONV3) K = halo_kpad(t->C0, t->C1);  // weights in halo layout
  int OH, OW;
  const int ld0 = t->ld0 ? t->ld0 : t->C0, ld_out = t->ld_out ? t->ld_out : t->N;
  if (t->loader == LD_TMA) {
    OH = t->IH; OW = t->IW;
    if (t->IH == 1 && t->R == 1) setup_tma_gemm(op, (const __half*)t->src0, (long)t->B * t->IW, t->C0, ld0);
    else setup_tma_conv(op, (const __half*)t->src0, t->B, t->IH, t->IW, t->C0, ld0, t->R, t->S, t->pad);
  } else if (t->loader == LD_GATHER_CONV) {
    OH = (t->IH + 2 * t->pad - t->R) / t->stride + 1;
    OW = (t->IW + 2 * t->pad - t->S) / t->stride + 1;
    setup_gather_conv(op, LD_GATHER_CONV, (const __half*)t->src0, t->C0, ld0, (const __half*)t->src1, t->C1, t->C1, t->B, t->IH, t->IW, OH, OW,
                      t->R, t->S, t->stride, t->pad, t->pad_mode);
  } else if (t->loader == LD_HALO_UPS) {
    OH = 2 * t->IH; OW = 2 * t->IW;
    setup_halo_ups(op, (const __half*)t->src0, t->C0, t->C0, (const __half*)t->src1, t->C1, t->C1, t->B, t->IH, t->IW);
  } else if (t->loader == LD_HALO_CONV3) {
    OH = t->IH; OW = t->IW;
    setup_halo_conv3(op, (const __half*)t->src0, t->B, t->IH, t->IW, t->C0, t->C0);
  } else {
    OH = t->IH; OW = t->IW;
    setup_gather_scale(op, (const __half*)t->src0, (long)t->B * t->IH * t->IW, t->C0, t->C0, t->a_scale, t->C0, t->rows_per_sample);
  }
  ConvGemmParams& p = op.p;
  p.epi = t->epi; p.act = t->act; p.bias = t->bias;
  p.resid16 = (const __half*)t->resid16; p.ld_res16 = t->N;
  p.resid32 = t->resid32; p.ld_res32 = ld_out;
  p.out16 = (__half*)t->out16; p.ld_out16 = ld_out;
  p.out32 = t->out32; p.ld_out32 = ld_out;
  p.ln_w = t->ln_w; p.ln_b = t->ln_b; p.ln_eps = 1e-6f;
  p.outc_w = t->outc_w; p.outc_b = t->outc_b; p.n_out = t->n_out; p.delta = t->delta; p.hw = OH * OW; p.outc_tanh = 1;
  p.grn_stats = t->grn_stats;
  if (t->rows_per_sample) p.rows_per_sample = t->rows_per_sample;
  finalize_op(op, (const __half*)t->weights, t->N, K, t->ldw ? t->ldw : K, num_sms, t->block_n);
  launch(op, (cudaStream_t)stream);
  g_launches += 1;
  return VSB_OK;
  VSB_API_END
}

int vsb_debug_resample_table(int32_t in, int32_t out, int32_t antialias, int32_t* start, int32_t* cnt, float* weights, int64_t capacity) {
  try {
    VSB_CHECK(in > 0 && out > 0 && start && cnt && weights, "bad argument");
    const ResampleHost t = make_resample(in, out, antialias != 0);
    VSB_CHECK(capacity >= (int64_t)out * t.maxt, "weights buffer too small");
    for (int i = 0; i < out; ++i) { start[i] = t.start[i]; cnt[i] = t.cnt[i]; }
    for (size_t i = 0; i < t.w.size(); ++i) weights[i] = t.w[i];
    return t.maxt;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"

// ---- per-step profile (bench.py roofline leg): enable, run some steps, then read "name\ttotal_ms\tcount\n" lines
extern "C" int vsb_profile_enable(int32_t on) {
  g_profile.flush();
  if (on) g_profile.acc.clear();
  g_profile.enabled = on != 0;
  return VSB_OK;
}
extern "C" int64_t vsb_profile_read(char* buf, int64_t capacity) {
  g_profile.flush();
  std::string out;
  for (auto& kv : g_profile.acc) out += kv.first + "\t" + std::to_string(kv.second.first) + "\t" + std::to_string(kv.second.second) + "\n";
  if (buf && capacity > (int64_t)out.size()) memcpy(buf, out.c_str(), out.size() + 1);
  return (int64_t)out.size() + 1;
}
