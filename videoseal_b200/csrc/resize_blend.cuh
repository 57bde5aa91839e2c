// HBM-bound full-resolution stage of the path, second generation (round 2):
//   K5  resize_sep_kernel    F.interpolate(bilinear, align_corners=False, antialias on/off) as a SEPARABLE resample: input rows are
//                            staged in shared memory with 128-bit coalesced loads, the horizontal pass runs out of shared memory
//                            with the taps in registers, the vertical pass out of a second shared buffer.  One launch handles key
//                            frames that are `frame_stride` apart (video mode).  Replaces the scalar per-output-pixel gather
//                            (49 global loads per output at 768 -> 256).  (models/wam.py:161-164,222-226; videoseal.py:303-310)
//   K7  jnd_blend2_kernel    JND heat-map x up-resampled delta, additive blend, clamp (modules/jnd.py:63-108, wam.py:183-201,
//                            blender.py:61-68, videoseal.py:80-118): 128 x 32 pixel tiles, every input pixel is read from HBM once
//                            (128-bit loads) and kept in shared memory, luminance halo 16 % instead of 55 %, 5x5 / Sobel windows
//                            read as 128-bit shared loads into a rolling register window, g^2.4 as ex2(1.2 lg2 g^2) on the SFU.
//                            Widths that are not multiples of 4 run the same kernel with lane-strided scalar accesses (still fully
//                            coalesced) instead of falling back to a different kernel.
#pragma once
#include "pointwise.cuh"

namespace vsb {

// ------------------------------------------------------------------------------------------------ K5
// in: planes [n][3][IH][IW] with frames `frame_stride` floats apart; out: [n*3][OH][OW] contiguous.
// grid (ceil(OH / toy), n*3); block 256.  Shared: hbuf [rin_max][OW] | stage [gstage][IWp].
template <int MAXT>   // x taps per output held in registers (2 or 8); 0 = any count, weights re-read from the table
__global__ void __launch_bounds__(256) resize_sep_kernel(const float* __restrict__ in, long frame_stride, float* __restrict__ out, int IH,
                                                         int IW, int OH, int OW, ResampleTab t, int toy, int gstage, int rin_max,
                                                         int vec) {
  extern __shared__ __align__(16) float rs_smem[];
  float* hbuf = rs_smem;
  float* stg = rs_smem + (size_t)rin_max * OW;
  const int IWp = (IW + 3) & ~3;
  const int pl = blockIdx.y;
  const int n = pl / 3, c = pl - n * 3;
  const float* src = in + (long)n * frame_stride + (long)c * IH * IW;
  const int oy0 = blockIdx.x * toy, oy1 = min(OH, oy0 + toy);
  const int r0 = __ldg(t.ystart + oy0);
  const int nr = __ldg(t.ystart + oy1 - 1) + __ldg(t.ycnt + oy1 - 1) - r0;   // <= rin_max (host-computed from the same table)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g0 = 0; g0 < nr; g0 += gstage) {
    const int ng = min(gstage, nr - g0);
    for (int r = warp; r < ng; r += 8) {
      const float* row = src + (long)(r0 + g0 + r) * IW;
      float* d = stg + r * IWp;
      if (vec) {
        for (int i = lane; i < (IW >> 2); i += 32) reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(row) + i);
      } else {
        for (int i = lane; i < IW; i += 32) d[i] = __ldg(row + i);
      }
    }
    __syncthreads();
    for (int ox = threadIdx.x; ox < OW; ox += 256) {
      const int xs = __ldg(t.xstart + ox), xc = __ldg(t.xcnt + ox);
      const float* wp = t.xw + (long)ox * t.maxt_x;
      if (MAXT > 0) {
        float w[MAXT > 0 ? MAXT : 1];
#pragma unroll
        for (int i = 0; i < MAXT; ++i) w[i] = i < xc ? __ldg(wp + i) : 0.f;
        for (int r = 0; r < ng; ++r) {
          const float* s = stg + r * IWp + xs;
          float acc = w[0] * s[0];
#pragma unroll
          for (int i = 1; i < MAXT; ++i) acc = i < xc ? fmaf(w[i], s[i], acc) : acc;
          hbuf[(g0 + r) * OW + ox] = acc;
        }
      } else {
        for (int r = 0; r < ng; ++r) {
          const float* s = stg + r * IWp + xs;
          float acc = __ldg(wp) * s[0];
          for (int i = 1; i < xc; ++i) acc = fmaf(__ldg(wp + i), s[i], acc);
          hbuf[(g0 + r) * OW + ox] = acc;
        }
      }
    }
    __syncthreads();
  }
  float* dst = out + (long)pl * OH * OW;
  for (int oy = oy0 + warp; oy < oy1; oy += 8) {          // one warp per output row: the y taps are warp-uniform
    const int ys = __ldg(t.ystart + oy) - r0, yc = __ldg(t.ycnt + oy);
    const float* wy = t.yw + (long)oy * t.maxt_y;
    for (int ox = lane; ox < OW; ox += 32) {
      const float* h = hbuf + ys * OW + ox;
      float acc = __ldg(wy) * h[0];
      for (int j = 1; j < yc; ++j) acc = fmaf(__ldg(wy + j), h[j * OW], acc);
      dst[oy * OW + ox] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K7
constexpr int kB2TW = 128, kB2TH = 32, kB2LP = 136;   // tile, and the pitch of the luminance tile (image column x0 + j at index 4 + j)
constexpr size_t kB2Smem = (size_t)((kB2TH + 4) * kB2LP + 3 * kB2TH * kB2TW) * sizeof(float);

__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx_f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lum255(float r, float g, float b) {
  return 0.299f * (255.f * r) + 0.587f * (255.f * g) + 0.114f * (255.f * b);
}
// modules/jnd.py:63-108 on L = 255*Y: la = conv5x5(L)/32 -> piecewise; cm = .117 * 16 g^2.4 / (g^2 + 26^2), g = |Sobel|;
// hmap = max(la + cm - .3 min(la, cm), 0) / 255.  la_sum = the 5x5 weighted sum (weights 1 / 2 / 0, jnd.py:37-43).
__device__ __forceinline__ float jnd_value(float la_sum, float gx, float gy) {
  float la = la_sum * (1.f / 32.f);
  la = (la <= 127.f) ? 17.f * (1.f - sqrtf(la * (1.f / 127.f) + 1e-5f)) : (3.f / 128.f) * (la - 127.f) + 3.f;
  const float g2 = gx * gx + gy * gy;
  const float pw = ex2_approx_f(1.2f * lg2_approx(g2));    // g^2.4 = (g^2)^1.2;  g2 == 0 -> lg2 = -inf -> 0
  const float cm = 0.117f * (16.f * pw / (g2 + 676.f));
  return fmaxf(la + cm - 0.3f * fminf(la, cm), 0.f) * (1.f / 255.f);
}
// window rows r0..r4 are pointers to the element of the CENTRE column in five consecutive luminance rows
__device__ __forceinline__ float jnd_window(const float* r0, const float* r1, const float* r2, const float* r3, const float* r4) {
  const float la = (r0[-2] + r0[-1] + r0[0] + r0[1] + r0[2]) + (r1[-2] + r1[2]) + (r2[-2] + r2[2]) + (r3[-2] + r3[2]) +
                   (r4[-2] + r4[-1] + r4[0] + r4[1] + r4[2]) + 2.f * (r1[-1] + r1[0] + r1[1] + r2[-1] + r2[1] + r3[-1] + r3[0] + r3[1]);
  const float gx = (r1[1] - r1[-1]) + 2.f * (r2[1] - r2[-1]) + (r3[1] - r3[-1]);
  const float gy = (r1[-1] + 2.f * r1[0] + r1[1]) - (r3[-1] + 2.f * r3[0] + r3[1]);
  return jnd_value(la, gx, gy);
}

// VEC 4: W % 4 == 0 and 16-byte aligned tensors: thread = 4 consecutive pixels (128-bit global / shared accesses).
// VEC 1: any W / alignment: thread = pixels lane, lane+32, lane+64, lane+96 of the tile row (32-bit accesses, fully coalesced).
// FASTUP 1: the delta is read in place or through <= 2 taps per axis (plain bilinear up-scale); 0: generic separable tables.
template <int VEC, int FASTUP>
__global__ void __launch_bounds__(256, 2) jnd_blend2_kernel(const BlendParams p) {
  extern __shared__ __align__(16) float b2_smem[];
  float* lum = b2_smem;                               // [TH + 4][LP]
  float* rgb = b2_smem + (kB2TH + 4) * kB2LP;         // [3][TH][TW]
  const int f = blockIdx.z, x0 = blockIdx.x * kB2TW, y0 = blockIdx.y * kB2TH;
  const long plane = (long)p.H * p.W;
  const float* img = p.imgs + (long)f * 3 * plane;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto col_of = [&](int q) { return VEC == 4 ? lane * 4 + q : lane + 32 * q; };

  // ---- phase 1: this thread's pixels (4 rows x 4 pixels x RGB) -> shared; luminance tile with a 2-pixel zero halo
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yl = warp * 4 + k, gy = y0 + yl;
    if (VEC == 4) {
      const int gx = x0 + lane * 4;
      float4 v[3];
      v[0] = v[1] = v[2] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy < p.H && gx < p.W) {
        const float* s = img + (long)gy * p.W + gx;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __ldg(reinterpret_cast<const float4*>(s + c * plane));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(rgb + (c * kB2TH + yl) * kB2TW + lane * 4) = v[c];
      if (p.use_jnd)
        *reinterpret_cast<float4*>(lum + (yl + 2) * kB2LP + 4 + lane * 4) =
            make_float4(lum255(v[0].x, v[1].x, v[2].x), lum255(v[0].y, v[1].y, v[2].y), lum255(v[0].z, v[1].z, v[2].z),
                        lum255(v[0].w, v[1].w, v[2].w));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = lane + 32 * q, gx = x0 + col;
        float v[3] = {0.f, 0.f, 0.f};
        if (gy < p.H && gx < p.W) {
          const float* s = img + (long)gy * p.W + gx;
#pragma unroll
          for (int c = 0; c < 3; ++c) v[c] = __ldg(s + c * plane);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[(c * kB2TH + yl) * kB2TW + col] = v[c];
        if (p.use_jnd) lum[(yl + 2) * kB2LP + 4 + col] = lum255(v[0], v[1], v[2]);
      }
    }
  }
  if (p.use_jnd) {
    if (warp < 4) {   // halo rows: shared rows 0, 1, TH+2, TH+3 <-> image rows y0-2, y0-1, y0+TH, y0+TH+1
      const int s = warp < 2 ? warp : kB2TH + warp;
      const int gy = y0 + s - 2;
      const bool rowok = gy >= 0 && gy < p.H;
      if (VEC == 4) {
        const int gx = x0 + lane * 4;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rowok && gx < p.W) {
          const float* sp = img + (long)gy * p.W + gx;
          const float4 r = __ldg(reinterpret_cast<const float4*>(sp)), g = __ldg(reinterpret_cast<const float4*>(sp + plane)),
                       b = __ldg(reinterpret_cast<const float4*>(sp + 2 * plane));
          l4 = make_float4(lum255(r.x, g.x, b.x), lum255(r.y, g.y, b.y), lum255(r.z, g.z, b.z), lum255(r.w, g.w, b.w));
        }
        *reinterpret_cast<float4*>(lum + s * kB2LP + 4 + lane * 4) = l4;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = lane + 32 * q, gx = x0 + col;
          float l = 0.f;
          if (rowok && gx < p.W) {
            const float* sp = img + (long)gy * p.W + gx;
            l = lum255(__ldg(sp), __ldg(sp + plane), __ldg(sp + 2 * plane));
          }
          lum[s * kB2LP + 4 + col] = l;
        }
      }
    }
    const int ht = (int)threadIdx.x - (256 - 4 * (kB2TH + 4));   // halo columns x0-2, x0-1, x0+TW, x0+TW+1 of all TH+4 rows
    if (ht >= 0) {
      const int s = ht >> 2, j = ht & 3;
      const int col = j < 2 ? j - 2 : kB2TW + j - 2;
      const int gy = y0 + s - 2, gx = x0 + col;
      float l = 0.f;
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
        const float* sp = img + (long)gy * p.W + gx;
        l = lum255(__ldg(sp), __ldg(sp + plane), __ldg(sp + 2 * plane));
      }
      lum[s * kB2LP + 4 + col] = l;
    }
  }
  __syncthreads();

  // ---- phase 2
  const FrameKeys fk = frame_keys(f, p.F, p.step, p.alternate, p.interp_chunk);
  const bool has_delta = fk.has;
  const int nsrc = (fk.k1 != fk.k0 && fk.a != 1.f) ? 2 : 1;
  // x taps of this thread's 4 pixels for the delta up-sample (FASTUP, not the identity)
  int xs[4], xo[4];
  float xw0[4], xw1[4];
  if (FASTUP && !p.identity_resample) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ox = min(x0 + col_of(q), p.W - 1);
      xs[q] = __ldg(p.tab.xstart + ox);
      const int xc = __ldg(p.tab.xcnt + ox);
      xw0[q] = __ldg(p.tab.xw + (long)ox * p.tab.maxt_x);
      xw1[q] = xc > 1 ? __ldg(p.tab.xw + (long)ox * p.tab.maxt_x + 1) : 0.f;
      xo[q] = xc > 1 ? 1 : 0;
    }
  }
  // rolling window of luminance rows (VEC 4): row i of the window = shared row warp*4 + k + i, 12 floats from shared column lane*4
  float Lw[5][12];
  if (VEC == 4 && p.use_jnd) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4* s4 = reinterpret_cast<const float4*>(lum + (warp * 4 + i) * kB2LP + lane * 4);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const float4 t4 = s4[u];
        Lw[i][4 * u] = t4.x; Lw[i][4 * u + 1] = t4.y; Lw[i][4 * u + 2] = t4.z; Lw[i][4 * u + 3] = t4.w;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yl = warp * 4 + k, gy = y0 + yl;
    float hm[4] = {1.f, 1.f, 1.f, 1.f};
    if (p.use_jnd) {
      if (VEC == 4) {
        const float4* s4 = reinterpret_cast<const float4*>(lum + (yl + 4) * kB2LP + lane * 4);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const float4 t4 = s4[u];
          Lw[4][4 * u] = t4.x; Lw[4][4 * u + 1] = t4.y; Lw[4][4 * u + 2] = t4.z; Lw[4][4 * u + 3] = t4.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) hm[q] = jnd_window(&Lw[0][4 + q], &Lw[1][4 + q], &Lw[2][4 + q], &Lw[3][4 + q], &Lw[4][4 + q]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 12; ++j) Lw[i][j] = Lw[i + 1][j];
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* cpt = lum + yl * kB2LP + 4 + lane + 32 * q;
          hm[q] = jnd_window(cpt, cpt + kB2LP, cpt + 2 * kB2LP, cpt + 3 * kB2LP, cpt + 4 * kB2LP);
        }
      }
    }
    if (gy >= p.H) continue;                                   // uniform per warp
    // ---- delta of this row (up-resampled from PH x PW) x heat-map
    float d[3][4];
    int ys = 0, yo = 0;
    float yw0 = 1.f, yw1 = 0.f;
    if (FASTUP && !p.identity_resample) {
      ys = __ldg(p.tab.ystart + gy);
      const int yc = __ldg(p.tab.ycnt + gy);
      yw0 = __ldg(p.tab.yw + (long)gy * p.tab.maxt_y);
      yw1 = yc > 1 ? __ldg(p.tab.yw + (long)gy * p.tab.maxt_y + 1) : 0.f;
      yo = yc > 1 ? 1 : 0;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c >= p.CD) break;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = 0.f;
        const int gx = x0 + col_of(q);
        if (has_delta && gx < p.W) {
          for (int s = 0; s < nsrc; ++s) {
            const float* src = p.delta + ((long)(s ? fk.k1 : fk.k0) * p.CD + c) * p.PH * p.PW;
            float vs;
            if (p.identity_resample) {
              vs = __ldg(src + (long)gy * p.PW + gx);
            } else if (FASTUP) {
              const float* a = src + (long)ys * p.PW + xs[q];
              const float* b = a + yo * p.PW;
              const float ra = xw0[q] * __ldg(a) + xw1[q] * __ldg(a + xo[q]);
              const float rb = xw0[q] * __ldg(b) + xw1[q] * __ldg(b + xo[q]);
              vs = yw0 * ra + yw1 * rb;
            } else {
              const int ys2 = p.tab.ystart[gy], yc = p.tab.ycnt[gy], xs2 = p.tab.xstart[gx], xc = p.tab.xcnt[gx];
              vs = 0.f;
              for (int j = 0; j < yc; ++j) {
                float racc = 0.f;
                for (int i = 0; i < xc; ++i) racc += p.tab.xw[gx * p.tab.maxt_x + i] * __ldg(src + (long)(ys2 + j) * p.PW + xs2 + i);
                vs += p.tab.yw[gy * p.tab.maxt_y + j] * racc;
              }
            }
            v += (nsrc == 1 ? 1.f : (s ? 1.f - fk.a : fk.a)) * vs;
          }
        }
        d[c][q] = v * hm[q];
      }
    }
    const long o = (long)gy * p.W + x0;
    const bool vfull = VEC == 4 && (x0 + lane * 4 < p.W);      // W % 4 == 0: a 4-pixel group is inside or outside as a whole
    if (p.preds_w != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c >= p.CD) break;
        float* dst = p.preds_w + ((long)f * p.CD + c) * plane + o;
        if (VEC == 4) {
          if (vfull) *reinterpret_cast<float4*>(dst + lane * 4) = make_float4(d[c][0], d[c][1], d[c][2], d[c][3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (x0 + lane + 32 * q < p.W) dst[lane + 32 * q] = d[c][q];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float in[4], out[4], dd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) dd[q] = (p.CD == 1 || c == 0) ? d[0][q] : (c == 1 ? d[1][q] : d[2][q]);
      if (VEC == 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(rgb + (c * kB2TH + yl) * kB2TW + lane * 4);
        in[0] = t4.x; in[1] = t4.y; in[2] = t4.z; in[3] = t4.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) in[q] = rgb[(c * kB2TH + yl) * kB2TW + lane + 32 * q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = p.scaling_i * in[q] + p.scaling_w * dd[q];
        if (p.clamp) v = fminf(fmaxf(v, 0.f), 1.f);
        out[q] = v;
      }
      float* dst = p.imgs_w + ((long)f * 3 + c) * plane + o;
      if (VEC == 4) {
        if (vfull) *reinterpret_cast<float4*>(dst + lane * 4) = make_float4(out[0], out[1], out[2], out[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (x0 + lane + 32 * q < p.W) dst[lane + 32 * q] = out[q];
      }
    }
  }
}

}  // namespace vsb
