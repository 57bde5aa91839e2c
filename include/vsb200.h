/* libvsb200.so — C ABI of the B200-native VideoSeal embed/detect hot path.
 *
 * Plain pointers and sizes only (no torch types).  All `*_dev` pointers are CUDA device pointers on the
 * device the model was finalised on; `stream` is a cudaStream_t passed as void*.  Every entry point returns
 * 0 on success or a negative vsb_status; vsb_last_error() gives the message.  There is NO CPU fallback:
 * without an sm_100 GPU vsb_model_finalize() fails.
 *
 * Each entry point names the reference interface it replaces (paths relative to facebookresearch/videoseal):
 *   vsb_model_create / set_tensor / finalize  <-  videoseal/utils/cfg.py:88-149 setup_model (build_embedder
 *        models/embedder.py:243, build_extractor models/extractor.py:170, load_state_dict of checkpoint['model'])
 *   vsb_embed                                 <-  models/wam.py:134-204 Wam.embed and models/videoseal.py:258-350
 *        Videoseal.embed (resize, RGB->Y, UnetEmbedder.forward embedder.py:151, _apply_video_mode, JND.heatmaps
 *        modules/jnd.py:80, Blender.additive_blend blender.py:61, clamp)
 *   vsb_embedder_forward                      <-  models/embedder.py:151-165 UnetEmbedder.forward (operator seam)
 *   vsb_detect                                <-  models/wam.py:206-234 Wam.detect / videoseal.py:352-388 (resize,
 *        ConvnextExtractor.forward models/extractor.py:154)
 *   vsb_jnd_heatmaps                          <-  modules/jnd.py:80-108 JND.heatmaps (operator seam)
 *   vsb_embed_host / vsb_detect_host          <-  the same calls with HOST (CPU) tensors as the reference keeps
 *        full-resolution video on the CPU (evals/full.py:117-120); H2D/D2H copies happen inside.
 */
#ifndef VSB200_H
#define VSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vsb_model vsb_model;

enum vsb_status {
  VSB_OK = 0,
  VSB_ERR_INVALID = -1,      /* bad argument / shape */
  VSB_ERR_UNSUPPORTED = -2,  /* card feature outside the hot path (raise at load() time) */
  VSB_ERR_CUDA = -3,         /* CUDA error, message in vsb_last_error() */
  VSB_ERR_STATE = -4         /* call order (e.g. embed before finalize, missing weight) */
};

enum vsb_video_mode { VSB_VIDEO_REPEAT = 0, VSB_VIDEO_ALTERNATE = 1, VSB_VIDEO_INTERPOLATE = 2 };

enum vsb_flags {
  VSB_FLAG_CLAMP = 1,            /* clamp imgs_w to [0,1]           (wam.py:196-197) */
  VSB_FLAG_LOWRES_ATTN = 2,      /* JND at processing resolution    (wam.py:177-180) */
  VSB_FLAG_NO_ATTENUATION = 4,   /* model.attenuation = None        */
  VSB_FLAG_RESIZE_NO_AA = 8      /* interpolation antialias=False   (videoseal.py:394) */
};

/* Architecture hyper-parameters taken from the model card (videoseal/cards/<card>.yaml). */
typedef struct vsb_model_desc {
  int32_t nbits;             /* args.nbits */
  int32_t hidden;            /* int(nbits * hidden_size_multiplier), embedder.py:244 */
  int32_t img_size;          /* args.img_size_proc | img_size_extractor */
  int32_t yuv;               /* 'yuv' in embedder.model, embedder.py:281 */
  int32_t unet_in_ch, unet_out_ch;
  int32_t unet_levels;       /* len(z_channels_mults), <= 6 */
  int32_t unet_z[6];         /* z_channels * mult */
  int32_t unet_num_blocks;
  int32_t unet_act;          /* 0 relu (the only value the GPU path implements; 1 = silu is rejected by vsb_model_finalize) */
  int32_t unet_norm;         /* 0 batch (the only value the GPU path implements; 1 = rms is rejected by vsb_model_finalize) */
  int32_t unet_last_tanh;
  int32_t ext_depths[4];
  int32_t ext_dims[4];       /* after proportional_dim scaling, extractor.py:193-198 */
  int32_t ext_stem_stride;
  int32_t jnd_in_ch, jnd_out_ch; /* configs/attenuation.yaml entry (jnd_1_1 / jnd_1_3 / jnd_3_1 / jnd_3_3); 0,0 = no attenuation */
} vsb_model_desc;

const char* vsb_last_error(void);
int vsb_version(void);

int vsb_model_create(const vsb_model_desc* desc, vsb_model** out);
/* one state_dict entry, fp32, host memory, contiguous; `name` is the reference state_dict key */
int vsb_model_set_tensor(vsb_model* m, const char* name, const float* data, const int64_t* shape, int32_t ndim);
/* fold BN, pack to fp16 KRSC, upload to `device`, build nothing else (plans are built lazily per batch shape) */
int vsb_model_finalize(vsb_model* m, int32_t device);
void vsb_model_destroy(vsb_model* m);

/* imgs_dev [F,3,H,W] fp32 in [0,1]; msgs_dev uint8 {0,1}: [F,nbits] (n_msgs == F) or [1,nbits] (n_msgs == 1);
 * imgs_w_dev [F,3,H,W]; preds_w_dev NULL or [F,PC,H,W] (hmap * delta_up, before scaling_w; PC = max(unet_out_ch, jnd_out_ch) with
 * attenuation on, the broadcast of `hmaps * preds_w` in wam.py:189-193, else unet_out_ch);
 * step = 1 for image mode, else frames share the key frame f/step (videoseal.py:292-340);
 * chunk_keys = the model's chunk_size (key frames per reference chunk, videoseal.py:300): it only changes the result in
 * video_mode 'interpolate', where neighbouring key frames of ONE chunk are mixed (videoseal.py:101-113); 1..64 there. */
int vsb_embed(vsb_model* m, const float* imgs_dev, const uint8_t* msgs_dev, int32_t n_msgs, float* imgs_w_dev,
              float* preds_w_dev, int32_t F, int32_t H, int32_t W, int32_t step, int32_t video_mode, int32_t chunk_keys,
              float scaling_i, float scaling_w, int32_t flags, void* stream);

/* x_dev [B, 3, S, S] RGB fp32 in [0,1] at processing size S = img_size (the Y extraction for yuv cards happens
 * inside, wam.py:168-172); delta_dev [B, unet_out_ch, S, S] */
int vsb_embedder_forward(vsb_model* m, const float* x_dev, const uint8_t* msgs_dev, int32_t n_msgs, float* delta_dev,
                         int32_t B, void* stream);

/* imgs_dev [F,3,H,W] -> logits_dev [F, 1+nbits] */
int vsb_detect(vsb_model* m, const float* imgs_dev, float* logits_dev, int32_t F, int32_t H, int32_t W, int32_t flags,
               void* stream);

/* imgs_dev [F,3,H,W] -> hmaps_dev [F,jnd_out_ch,H,W] */
int vsb_jnd_heatmaps(vsb_model* m, const float* imgs_dev, float* hmaps_dev, int32_t F, int32_t H, int32_t W, void* stream);

/* Host-buffer variants (pinned or pageable): copy in, run, copy out, synchronise. */
int vsb_embed_host(vsb_model* m, const float* imgs_host, const uint8_t* msgs_host, int32_t n_msgs, float* imgs_w_host,
                   float* preds_w_host, int32_t F, int32_t H, int32_t W, int32_t step, int32_t video_mode,
                   int32_t chunk_keys, float scaling_i, float scaling_w, int32_t flags);
int vsb_detect_host(vsb_model* m, const float* imgs_host, float* logits_host, int32_t F, int32_t H, int32_t W, int32_t flags);

/* Streaming form of embed followed by detect of the watermarked frames, HOST (ideally pinned) buffers in and out: the frames are
 * processed in chunks of ~32 with the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlapped with the compute of chunk k
 * (three internal streams).  This is how the reference's callers use the path on full-resolution video kept on the CPU
 * (evals/full.py:117-120, inference_streaming.py:83-107).  imgs_w_host [F,3,H,W], logits_host [F,1+nbits]. */
int vsb_embed_detect_host(vsb_model* m, const float* imgs_host, const uint8_t* msgs_host, int32_t n_msgs, float* imgs_w_host,
                          float* logits_host, int32_t F, int32_t H, int32_t W, int32_t step, int32_t video_mode,
                          int32_t chunk_keys, float scaling_i, float scaling_w, int32_t flags);

/* The streaming caller's form (inference_streaming.py:23-33 embed_video_clip, :116-124 detect_video_clip): RGB24 frames
 * [F,H,W,3] uint8 on the host, converted on the GPU (x = u8 / 255; u8 = (uint8) trunc(imgs_w * 255)) so that PCIe carries one byte
 * per sample instead of four; same chunked three-stream overlap as vsb_embed_detect_host.
 *   frames_w_host != NULL, logits_host == NULL : embed only            (embed_video_clip)
 *   frames_w_host == NULL, logits_host != NULL : detect the input      (detect_video_clip; msgs_host may be NULL)
 *   both != NULL                               : embed, then detect the re-quantised watermarked frames */
int vsb_frames_host_u8(vsb_model* m, const uint8_t* frames_host, const uint8_t* msgs_host, int32_t n_msgs, uint8_t* frames_w_host,
                       float* logits_host, int32_t F, int32_t H, int32_t W, int32_t step, int32_t video_mode, int32_t chunk_keys,
                       float scaling_i, float scaling_w, int32_t flags);

/* number of kernels launched by this library since the last call with reset != 0 (bench.py's gpu_launches) */
int64_t vsb_launch_count(int32_t reset);

/* per-step CUDA-event profile used by bench.py's roofline leg: enable(1), run steps, read "name\ttotal_ms\tcount\n" lines */
int vsb_profile_enable(int32_t on);
int64_t vsb_profile_read(char* buf, int64_t capacity);

/* ---- debug / test seams (used by tests/ only) ------------------------------------------------------------- */
/* copy a named intermediate activation of the most recent embed/detect sub-batch to the host as fp32.
 * shape4 receives [B, H, W, C] (NHWC) or [M, C, 1, 1]-style dims; returns the element count or <0. */
int64_t vsb_debug_get_tensor(vsb_model* m, const char* name, float* host_out, int64_t capacity, int64_t* shape4);
/* standalone conv/GEMM on the tcgen05 kernel for unit tests: see tests/test_conv_gemm_gpu.py */
typedef struct vsb_conv_test {
  int32_t loader;            /* 0 TMA per tap, 1 gather conv, 2 halo bilinear-x2+reflect (UBlock), 3 gather GRN-scale, 4 halo conv3x3 */
  int32_t B, IH, IW, C0, C1; /* input NHWC fp16 [B,IH,IW,C0] (+ optional second source [.., C1]) */
  int32_t R, S, stride, pad, pad_mode;
  int32_t N;                 /* C_out; weights fp16 [N][R*S*(C0+C1)] */
  int32_t epi, act;          /* 0 affine / 1 LN;  0 none 1 relu 2 gelu */
  int32_t rows_per_sample;   /* GRN / scale */
  int32_t block_n;           /* 0 = auto */
  const void* src0; const void* src1; const void* weights;
  const float* bias; const void* resid16; const float* resid32; const float* a_scale;
  const float* ln_w; const float* ln_b;
  const float* outc_w; const float* outc_b; int32_t n_out;
  void* out16; float* out32; float* delta; float* grn_stats;
  int32_t ld0;               /* pixel pitch of src0 in elements (0 = C0): widths that are not multiples of 8 live in padded rows */
  int32_t ldw;               /* row pitch of the weights (0 = K) */
  int32_t ld_out;            /* row pitch of out16 / out32 / resid32 (0 = N) */
} vsb_conv_test;
int vsb_debug_conv(const vsb_conv_test* t, void* stream);
/* host-only (no GPU needed): the separable table the resize kernels use along one axis of length `in` -> `out`, i.e. the library's
 * restatement of ATen's bilinear (anti-aliased or plain, align_corners=False) index/weight computation behind F.interpolate
 * (wam.py:163,184,224).  start / cnt: [out]; weights: [out * maxt] row-major, capacity in floats.  Returns maxt, or < 0. */
int vsb_debug_resample_table(int32_t in, int32_t out, int32_t antialias, int32_t* start, int32_t* cnt, float* weights, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* VSB200_H */
