// Layout/packing kernels, the timestep MLP, and the fused final LayerNorm + 7x7 conv + DDIM update.
#pragma once
#include "../../include/cdc_b200.h"
#include "common.cuh"
#include "igemm_tc.cuh"   // FastDiv

namespace cdc {

// ------------------------------------------------------------------------------------------------
// x_t (fp32 NCHW, `cx` channels) [+ a small fp32 NCHW context with `cc` channels, cx+cc <= 8]
//   -> the packed network input in WINDOW FORM, X8 [B][H][W+8][8] fp16: one 16-byte record per pixel (x channels | context
//   channels | zeros), 3 zero pixels left and 5 right of every row.  The first 7x7 convolution (unet.py:61 /
//   network_components.py:87) reads it through an overlapping-window tensor map (ConvSeg::win8): "channel" kx*8+c of pixel x
//   = the value at (y, x+kx-3), i.e. the horizontal taps folded into K, so that the convolution runs as 7 vertical taps of 64
//   channels.  Slot kx=7 (pixel x+4) has zero weights; the 1x1 res_conv reads slot kx=3 (the pixel itself).
// ------------------------------------------------------------------------------------------------
__global__ void pack_input_kernel(const float* __restrict__ x, int cx, const float* __restrict__ ctx, int cc,
                                  int B, int H, int W, FastDiv fdWp, FastDiv fdH, __half* __restrict__ out) {
  // out: [B][H][W + 8][8] fp16 — pixel record = (x channels | folded context channels | zeros); 3 zero pixels on the left
  // and 5 on the right of every row, so that the 64 halves starting at padded pixel x are the horizontal taps x-3 .. x+4 of
  // logical pixel x (ConvSeg::win8: the convolution reads them through an overlapping-window tensor map)
  pdl_launch_dependents();
  pdl_wait();
  const int Wp = W + 8;
  const int total = B * H * Wp;   // < 2^31 for every supported shape (checked by the engine)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = fdiv(i, fdWp);
    const int xx = i - t * Wp - 3;
    const int b = fdiv(t, fdH);
    const int yy = t - b * H;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    if (xx >= 0 && xx < W) {
      const size_t row = (size_t)yy * W + xx;
      const size_t plane = (size_t)H * W;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < cx) v[c] = x[((size_t)b * cx + c) * plane + row];
        else if (c < cx + cc) v[c] = ctx[((size_t)b * cc + (c - cx)) * plane + row];
      }
    }
    uint4 o;
    o.x = pack_half2(v[0], v[1]);
    o.y = pack_half2(v[2], v[3]);
    o.z = pack_half2(v[4], v[5]);
    o.w = pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (size_t)i * 8) = o;
  }
}

// fp32 NCHW [B,C,HW] -> fp16 NHWC [B,HW,C]   (context maps, once per decode)
__global__ void nchw_to_nhwc_half_kernel(const float* __restrict__ in, int C, int HW, __half* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    tile[j][tx] = (c < C && pp < HW) ? in[((size_t)b * C + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    if (c < C && pp < HW) out[((size_t)b * HW + pp) * C + c] = __float2half_rn(tile[tx][j]);
  }
}

// fp32 NCHW [B,C,HW] -> fp16 NHWC [B,HW,C] value + rounding remainder (the quantised latent entering context_fn.decode:
// a "trunk" tensor — its 1x1 res_conv runs as three compensated passes)
__global__ void nchw_to_nhwc_hilo_kernel(const float* __restrict__ in, int C, int HW, __half* __restrict__ out,
                                         __half* __restrict__ out_lo) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    tile[j][tx] = (c < C && pp < HW) ? in[((size_t)b * C + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    if (c < C && pp < HW) {
      const float v = tile[tx][j];
      const __half h = __float2half_rn(v);
      out[((size_t)b * HW + pp) * C + c] = h;
      out_lo[((size_t)b * HW + pp) * C + c] = __float2half_rn(v - __half2float(h));
    }
  }
}

// fp16 NHWC [B,HW,C] -> fp32 NCHW [B,C,HW]   (read-back of a context map: tests / debugging)
__global__ void nhwc_half_to_nchw_kernel(const __half* __restrict__ in, int C, int HW, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    tile[j][tx] = (c < C && pp < HW) ? __half2float(in[((size_t)b * HW + pp) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    if (c < C && pp < HW) out[((size_t)b * C + c) * HW + pp] = tile[tx][j];
  }
}

// ConvTranspose2d(C_in, C_out <= 8, 4, stride 2, pad 1) with fp32 weights on a hi+lo NHWC input, fp32 NCHW output: the
// last Upsample of the eps variant's context decoder (64 -> 3 channels, compress_modules.py:150-156), once per decode.
// Phase form of SURVEY.md Appendix E: output (oy, ox) with parity (py, px) reads the 2 x 2 inputs (i+ty+py-1, j+tx+px-1)
// through kernel taps (3-2ty-py, 3-2tx-px).  One thread per output pixel; the weights [ky][kx][c][o] sit in shared memory.
struct UpSmallParams {
  const __half* in;      // [B, h, w, Cin]
  const __half* in_lo;   // may be null
  const float* wt;       // [4][4][Cin][Cout]
  const float* bias;     // [Cout]
  float* out;            // [B, Cout, 2h, 2w]
  int B, h, w, Cin, Cout;
};
__global__ void __launch_bounds__(256) upsample_small_kernel(const UpSmallParams p) {
  extern __shared__ float ws_[];
  const int nw = 16 * p.Cin * p.Cout;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < nw; i += blockDim.x) ws_[i] = p.wt[i];   // weights: constant, may precede the wait
  pdl_wait();
  __syncthreads();
  const int OW = 2 * p.w, OH = 2 * p.h;
  const long long total = (long long)p.B * OH * OW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW);
    const long long t = idx / OW;
    const int oy = (int)(t % OH), b = (int)(t / OH);
    const int py = oy & 1, px = ox & 1, i = oy >> 1, j = ox >> 1;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = o < p.Cout ? p.bias[o] : 0.f;
    for (int ty = 0; ty < 2; ++ty) {
      const int iy = i + ty + py - 1, ky = 3 - 2 * ty - py;
      if (iy < 0 || iy >= p.h) continue;
      for (int tx = 0; tx < 2; ++tx) {
        const int ix = j + tx + px - 1, kx = 3 - 2 * tx - px;
        if (ix < 0 || ix >= p.w) continue;
        const size_t base = (((size_t)b * p.h + iy) * p.w + ix) * p.Cin;
        const float* wk = ws_ + (size_t)(ky * 4 + kx) * p.Cin * p.Cout;
        for (int c0 = 0; c0 < p.Cin; c0 += 8) {
          const uint4 vh = *reinterpret_cast<const uint4*>(p.in + base + c0);
          uint4 vl = make_uint4(0u, 0u, 0u, 0u);
          if (p.in_lo) vl = *reinterpret_cast<const uint4*>(p.in_lo + base + c0);
          const uint32_t hh[4] = {vh.x, vh.y, vh.z, vh.w}, ll[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 a = unpack_half2(hh[q]), l = unpack_half2(ll[q]);
            const float x0 = a.x + l.x, x1 = a.y + l.y;
            const float* w0 = wk + (size_t)(c0 + 2 * q) * p.Cout;
#pragma unroll
            for (int o = 0; o < 8; ++o)
              if (o < p.Cout) acc[o] = fmaf(x1, w0[p.Cout + o], fmaf(x0, w0[o], acc[o]));
          }
        }
      }
    }
    const size_t plane = (size_t)OH * OW;
    for (int o = 0; o < p.Cout; ++o) p.out[((size_t)b * p.Cout + o) * plane + (size_t)oy * OW + ox] = acc[o];
  }
}

// ------------------------------------------------------------------------------------------------
// Timestep path: temb = W2 gelu_erf(W1 t + b1) + b2 (unet.py:40), then for every ResnetBlock
// shift = Wk leaky_relu_0.2(temb) + bk (network_components.py:97-101,110-111).  One CTA per image;
// all blocks' Linear layers are concatenated row-wise ([R][dim]).
// ------------------------------------------------------------------------------------------------
__global__ void time_mlp_kernel(const float* __restrict__ time, const cdc_step_coef* __restrict__ table,
                                const int* __restrict__ step_ptr, const float* __restrict__ W1,
                                const float* __restrict__ b1, const float* __restrict__ W2,
                                const float* __restrict__ b2, const float* __restrict__ Wcat,
                                const float* __restrict__ bcat, int dim, int R, float* __restrict__ shifts) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* hid = sm;            // [4*dim]
  float* act = sm + 4 * dim;  // [dim]
  const int b = blockIdx.x;
  const float t = table ? table[*step_ptr].unet_time : time[b];
  for (int j = threadIdx.x; j < 4 * dim; j += blockDim.x) {
    const float v = W1[j] * t + b1[j];
    hid[j] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  }
  __syncthreads();
  // layer 2: one warp per output, coalesced float4 rows (4*dim is a multiple of 128), shuffle reduction in a fixed order
  for (int j = threadIdx.x >> 5; j < dim; j += blockDim.x >> 5) {
    const float4* w = reinterpret_cast<const float4*>(W2 + (size_t)j * 4 * dim);
    const float4* h4 = reinterpret_cast<const float4*>(hid);
    float a = 0.f;
    for (int k = threadIdx.x & 31; k < dim; k += 32) {
      const float4 wv = w[k], hv = h4[k];
      a = fmaf(wv.x, hv.x, a);
      a = fmaf(wv.y, hv.y, a);
      a = fmaf(wv.z, hv.z, a);
      a = fmaf(wv.w, hv.w, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    a += b2[j];
    if ((threadIdx.x & 31) == 0) act[j] = a > 0.f ? a : 0.2f * a;
  }
  __syncthreads();
  // blockIdx.y selects a 256-row slice of the concatenated per-block Linear layers (temb is recomputed per CTA); a row is
  // dim floats: sixteen 16-byte loads issued together instead of 64 dependent scalar ones
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r < R) {
    const float4* w = reinterpret_cast<const float4*>(Wcat + (size_t)r * dim);
    const float4* a4 = reinterpret_cast<const float4*>(act);
    float a0 = bcat[r], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < dim / 4; ++k) {
      const float4 wv = w[k], av = a4[k];
      a0 = fmaf(wv.x, av.x, a0);
      a1 = fmaf(wv.y, av.y, a1);
      a2 = fmaf(wv.z, av.z, a2);
      a3 = fmaf(wv.w, av.w, a3);
    }
    shifts[(size_t)b * R + r] = (a0 + a1) + (a2 + a3);
  }
}

__global__ void advance_step_kernel(int* step_ptr) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) *step_ptr -= 1;
}

// ------------------------------------------------------------------------------------------------
// final_conv = LayerNorm(64) -> Conv2d(64, channels<=8, 7, pad 3) (unet.py:93) fused with the DDIM
// update (epsilonparam/modules/denoising_diffusion.py:137-152; xparam/...:152-174).
// CTA = 16x16 output pixels; the 22x22x64 halo is LayerNorm-ed once into shared memory (zeros
// outside the image: padding applies to the LayerNorm output), then M=256,N=8,K=49*64 on mma.sync.
// ------------------------------------------------------------------------------------------------
struct FinalParams {
  const __half* in;     // [B,H,W,64] fp16
  const __half* in_lo;  // optional compensation half of `in`
  const float* ln_g;    // [64]
  const float* ln_b;
  const __half* Wf;     // [8][kFinalWStride] fp16, K index = (ky*7+kx)*64 + c, rows >= channels are zero
  const float* bias;    // [channels]
  int B, H, W, channels;
  // mode 0: out = network output (Unet.forward).  mode 1: in-place DDIM update of x.
  int mode;
  float* out;           // mode 0: [B,channels,H,W] fp32
  float* x;             // mode 1: x_t in / x_{t-1} out, fp32 NCHW
  const float* z;       // optional noise
  const cdc_step_coef* table;
  const int* step_ptr;
  int variant, pred_mode, clip_mode;
  // eta != 0 inside the captured loop: z holds the noise of several consecutive steps, step i at z + (*z_first - i) * z_stride
  const int* z_first;   // null: z is this step's tensor
  long long z_stride;
};

__device__ __forceinline__ const float* final_noise(const FinalParams& p) {
  return p.z_first ? p.z + (long long)(*p.z_first - *p.step_ptr) * p.z_stride : p.z;
}

constexpr int kFinalWStride = 49 * 64 + 8;
constexpr int kFinalHalo = 22;
constexpr int kFinalSmemBytes = kFinalHalo * kFinalHalo * 128 + 8 * kFinalWStride * 2;

__global__ void __launch_bounds__(256) final_conv_kernel(const FinalParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sIn = smem;
  __half* sW = reinterpret_cast<__half*>(smem + kFinalHalo * kFinalHalo * 128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;

  pdl_launch_dependents();
  // weights -> smem with cp.async (constant data; lands while the halo is being normalised)
  {
    const uint32_t dst = smem_u32(sW);
    for (int i = tid; i < 8 * kFinalWStride * 2 / 16; i += 256)
      cp_async16(dst + i * 16, reinterpret_cast<const uint4*>(p.Wf) + i, 16);
    cp_async_commit();
  }
  pdl_wait();
  // halo load + LayerNorm: 8 threads per pixel, 8 channels each; 8 pixels per thread are fetched up front so the
  // global-load latency is paid twice per CTA instead of once per pixel batch
  {
    const int j = tid & 7;
    float g[8], bb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      g[c] = p.ln_g[j * 8 + c];
      bb[c] = p.ln_b[j * 8 + c];
    }
    constexpr int kBatch = 8;
    for (int hp00 = 0; hp00 < kFinalHalo * kFinalHalo; hp00 += 32 * kBatch) {
      uint4 raw[kBatch], rlo[kBatch];
      bool inb[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int hp = hp00 + u * 32 + (tid >> 3);
        const int hy = hp / kFinalHalo, hx = hp - hy * kFinalHalo;
        const int yy = y0 + hy - 3, xx = x0 + hx - 3;
        inb[u] = hp < kFinalHalo * kFinalHalo && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        raw[u] = make_uint4(0u, 0u, 0u, 0u);
        rlo[u] = make_uint4(0u, 0u, 0u, 0u);
        if (inb[u]) {
          const size_t off = (((size_t)b * p.H + yy) * p.W + xx) * 64 + j * 8;
          raw[u] = *reinterpret_cast<const uint4*>(p.in + off);
          if (p.in_lo) rlo[u] = *reinterpret_cast<const uint4*>(p.in_lo + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int hp = hp00 + u * 32 + (tid >> 3);
        float v[8];
        float2 f;
        f = unpack_half2(raw[u].x); v[0] = f.x; v[1] = f.y;
        f = unpack_half2(raw[u].y); v[2] = f.x; v[3] = f.y;
        f = unpack_half2(raw[u].z); v[4] = f.x; v[5] = f.y;
        f = unpack_half2(raw[u].w); v[6] = f.x; v[7] = f.y;
        f = unpack_half2(rlo[u].x); v[0] += f.x; v[1] += f.y;
        f = unpack_half2(rlo[u].y); v[2] += f.x; v[3] += f.y;
        f = unpack_half2(rlo[u].z); v[4] += f.x; v[5] += f.y;
        f = unpack_half2(rlo[u].w); v[6] += f.x; v[7] += f.y;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) s += v[c];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float mean = s * (1.f / 64.f);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float d = v[c] - mean;
          q += d * d;
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        const float rstd = 1.f / sqrtf(q * (1.f / 64.f) + 1e-5f);
        if (hp < kFinalHalo * kFinalHalo) {
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (inb[u]) {
            float y[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) y[c] = (v[c] - mean) * rstd * g[c] + bb[c];
            o.x = pack_half2(y[0], y[1]);
            o.y = pack_half2(y[2], y[3]);
            o.z = pack_half2(y[4], y[5]);
            o.w = pack_half2(y[6], y[7]);
          }
          *reinterpret_cast<uint4*>(sIn + swz128(hp, j)) = o;
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
  const uint32_t sIn32 = smem_u32(sIn), sW32 = smem_u32(sW);
  for (int tap = 0; tap < 49; ++tap) {
    const int ky = tap / 7, kx = tap - ky * 7;
#pragma unroll
    for (int ks = 0; ks < 4; ks += 2) {
      uint32_t bf[4];
      ldmatrix_x4(bf, sW32 + ((lane & 7) * kFinalWStride + tap * 64 + ks * 16 + (lane >> 3) * 8) * 2);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int ty = warp * 2 + mt;
        const int hp = (ty + ky) * kFinalHalo + (lane & 15) + kx;
        uint32_t a0[4], a1[4];
        ldmatrix_x4(a0, sIn32 + swz128(hp, ks * 2 + (lane >> 4)));
        ldmatrix_x4(a1, sIn32 + swz128(hp, (ks + 1) * 2 + (lane >> 4)));
        mma_16816(acc[mt], a0, bf[0], bf[1]);
        mma_16816(acc[mt], a1, bf[2], bf[3]);
      }
    }
  }

  // epilogue: lane holds (pixel tx = lane>>2 (+8), channels n = 2*(lane&3), +1)
  const int n_base = (lane & 3) * 2;
  cdc_step_coef cf = {};
  if (p.mode == 1) cf = p.table[*p.step_ptr];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = n_base + e;
        if (n >= p.channels) continue;
        const int yy = y0 + warp * 2 + mt, xx = x0 + (lane >> 2) + h * 8;
        if (yy >= p.H || xx >= p.W) continue;
        const float f = acc[mt][2 * h + e] + p.bias[n];
        const size_t idx = (((size_t)b * p.channels + n) * p.H + yy) * p.W + xx;
        if (p.mode == 0) {
          p.out[idx] = f;
          continue;
        }
        const float xt = p.x[idx];
        float x0v, noise;
        const bool clip = p.clip_mode == CDC_CLIP_FULL || (p.clip_mode == CDC_CLIP_HALF && b < p.B / 2);
        if (p.variant == CDC_VARIANT_EPS || p.pred_mode == CDC_PRED_NOISE) {
          x0v = cf.sqrt_recip_acp * xt - cf.sqrt_recipm1_acp * f;
          if (clip) x0v = fminf(fmaxf(x0v, -1.f), 1.f);
          noise = f;
        } else {
          x0v = (p.pred_mode == CDC_PRED_X) ? f : cf.sqrt_acp * xt - cf.sqrt_1m_acp * f;
          if (clip) x0v = fminf(fmaxf(x0v, -1.f), 1.f);
          noise = (cf.sqrt_recip_acp * xt - x0v) / cf.sqrt_recipm1_acp;
        }
        float xn = cf.sqrt_acp_prev * x0v + cf.dir_coef * noise;
        if (p.z) xn += cf.noise_coef * final_noise(p)[idx];
        p.x[idx] = xn;
      }
}

// ------------------------------------------------------------------------------------------------
// Same operator for channels <= 3 (every CDC configuration), restructured so that the tensor-core operand traffic no
// longer dominates: with N = 3 output channels an MMA re-reads its whole A fragment for 3 useful columns, and the
// 49-tap form above moves 2 MB of ldmatrix traffic per CTA (shared-memory bound, ~200 us at 8x256x256).  Here the
// HORIZONTAL taps are moved into N:
//     Y[y, hx, kx*nch + ch] = sum_{ky, c} LN(in)[y + ky - 3, hx, c] * w[ch, c, ky, kx]      (M = 16 x 22, N = 24, K = 448)
//     out[y, x, ch]         = bias[ch] + sum_kx Y[y, x + kx, kx*nch + ch]                     (7 shifted adds per channel)
// so K shrinks 7x and every A fragment feeds three n-tiles.  Halo pixel index = m + 22*ky for output-row-major m, i.e.
// the vertical taps are plain row offsets into the same LayerNorm-ed halo tile.
// ------------------------------------------------------------------------------------------------
constexpr int kFinalW2Stride = 7 * 64 + 8;     // halfs per weight row (k = ky*64 + c), padded: conflict-free ldmatrix
constexpr int kFinalYStride = 26;              // floats per Y row (24 used)
constexpr int kFinalSmemBytes2 = kFinalHalo * kFinalHalo * 128 + 24 * kFinalW2Stride * 2;

__global__ void __launch_bounds__(256) final_conv_kx_kernel(const FinalParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sIn = smem;
  float* sY = reinterpret_cast<float*>(smem);   // reuses the halo tile once the MMAs are done
  __half* sW = reinterpret_cast<__half*>(smem + kFinalHalo * kFinalHalo * 128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;

  pdl_launch_dependents();
  {
    const uint32_t dst = smem_u32(sW);
    for (int i = tid; i < 24 * kFinalW2Stride * 2 / 16; i += 256)
      cp_async16(dst + i * 16, reinterpret_cast<const uint4*>(p.Wf) + i, 16);
    cp_async_commit();
  }
  pdl_wait();
  // halo load + LayerNorm (as in final_conv_kernel): 8 threads per pixel, 8 channels each
  {
    const int j = tid & 7;
    float g[8], bb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      g[c] = p.ln_g[j * 8 + c];
      bb[c] = p.ln_b[j * 8 + c];
    }
    constexpr int kBatch = 8;
    for (int hp00 = 0; hp00 < kFinalHalo * kFinalHalo; hp00 += 32 * kBatch) {
      uint4 raw[kBatch], rlo[kBatch];
      bool inb[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int hp = hp00 + u * 32 + (tid >> 3);
        const int hy = hp / kFinalHalo, hx = hp - hy * kFinalHalo;
        const int yy = y0 + hy - 3, xx = x0 + hx - 3;
        inb[u] = hp < kFinalHalo * kFinalHalo && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        raw[u] = make_uint4(0u, 0u, 0u, 0u);
        rlo[u] = make_uint4(0u, 0u, 0u, 0u);
        if (inb[u]) {
          const size_t off = (((size_t)b * p.H + yy) * p.W + xx) * 64 + j * 8;
          raw[u] = *reinterpret_cast<const uint4*>(p.in + off);
          if (p.in_lo) rlo[u] = *reinterpret_cast<const uint4*>(p.in_lo + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int hp = hp00 + u * 32 + (tid >> 3);
        float v[8];
        float2 f;
        f = unpack_half2(raw[u].x); v[0] = f.x; v[1] = f.y;
        f = unpack_half2(raw[u].y); v[2] = f.x; v[3] = f.y;
        f = unpack_half2(raw[u].z); v[4] = f.x; v[5] = f.y;
        f = unpack_half2(raw[u].w); v[6] = f.x; v[7] = f.y;
        f = unpack_half2(rlo[u].x); v[0] += f.x; v[1] += f.y;
        f = unpack_half2(rlo[u].y); v[2] += f.x; v[3] += f.y;
        f = unpack_half2(rlo[u].z); v[4] += f.x; v[5] += f.y;
        f = unpack_half2(rlo[u].w); v[6] += f.x; v[7] += f.y;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) sum += v[c];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        const float mean = sum * (1.f / 64.f);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float d = v[c] - mean;
          q += d * d;
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        const float rstd = 1.f / sqrtf(q * (1.f / 64.f) + 1e-5f);
        if (hp < kFinalHalo * kFinalHalo) {
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (inb[u]) {
            float y[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) y[c] = (v[c] - mean) * rstd * g[c] + bb[c];
            o.x = pack_half2(y[0], y[1]);
            o.y = pack_half2(y[2], y[3]);
            o.z = pack_half2(y[4], y[5]);
            o.w = pack_half2(y[6], y[7]);
          }
          *reinterpret_cast<uint4*>(sIn + swz128(hp, j)) = o;
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- Y = halo rows x (kx, ch): 22 m-tiles of 16 (output-row-major pixels incl. the horizontal halo), 3 n-tiles ----
  float acc[3][3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][n][k] = 0.f;
  const uint32_t sIn32 = smem_u32(sIn), sW32 = smem_u32(sW);
  const int nmt = warp < 6 ? 3 : 2;   // m-tiles warp, warp + 8, warp + 16 (< 22)
  for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
    for (int ks = 0; ks < 4; ks += 2) {
      uint32_t bf[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
        ldmatrix_x4(bf[nt], sW32 + ((nt * 8 + (lane & 7)) * kFinalW2Stride + ky * 64 + ks * 16 + (lane >> 3) * 8) * 2);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < nmt) {
          const int hp = (warp + 8 * i) * 16 + (lane & 15) + kFinalHalo * ky;
          uint32_t a0[4], a1[4];
          ldmatrix_x4(a0, sIn32 + swz128(hp, ks * 2 + (lane >> 4)));
          ldmatrix_x4(a1, sIn32 + swz128(hp, (ks + 1) * 2 + (lane >> 4)));
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            mma_16816(acc[i][nt], a0, bf[nt][0], bf[nt][1]);
            mma_16816(acc[i][nt], a1, bf[nt][2], bf[nt][3]);
          }
        }
      }
    }
  }
  __syncthreads();   // every warp is done with the halo tile: its memory becomes Y
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < nmt) {
      const int m = (warp + 8 * i) * 16 + (lane >> 2);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const int n = nt * 8 + (lane & 3) * 2;
        *reinterpret_cast<float2*>(sY + m * kFinalYStride + n) = make_float2(acc[i][nt][0], acc[i][nt][1]);
        *reinterpret_cast<float2*>(sY + (m + 8) * kFinalYStride + n) = make_float2(acc[i][nt][2], acc[i][nt][3]);
      }
    }
  }
  __syncthreads();

  // ---- horizontal taps + bias + sampler update: one output pixel per thread ----
  const int ty = tid >> 4, tx = tid & 15;
  const int yy = y0 + ty, xx = x0 + tx;
  if (yy >= p.H || xx >= p.W) return;
  const int nch = p.channels;
  cdc_step_coef cf = {};
  if (p.mode == 1) cf = p.table[*p.step_ptr];
  const float* yrow = sY + (ty * kFinalHalo + tx) * kFinalYStride;
  for (int n = 0; n < nch; ++n) {
    float f = p.bias[n];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) f += yrow[kx * kFinalYStride + kx * nch + n];
    const size_t idx = (((size_t)b * nch + n) * p.H + yy) * p.W + xx;
    if (p.mode == 0) {
      p.out[idx] = f;
      continue;
    }
    const float xt = p.x[idx];
    float x0v, noise;
    const bool clip = p.clip_mode == CDC_CLIP_FULL || (p.clip_mode == CDC_CLIP_HALF && b < p.B / 2);
    if (p.variant == CDC_VARIANT_EPS || p.pred_mode == CDC_PRED_NOISE) {
      x0v = cf.sqrt_recip_acp * xt - cf.sqrt_recipm1_acp * f;
      if (clip) x0v = fminf(fmaxf(x0v, -1.f), 1.f);
      noise = f;
    } else {
      x0v = (p.pred_mode == CDC_PRED_X) ? f : cf.sqrt_acp * xt - cf.sqrt_1m_acp * f;
      if (clip) x0v = fminf(fmaxf(x0v, -1.f), 1.f);
      noise = (cf.sqrt_recip_acp * xt - x0v) / cf.sqrt_recipm1_acp;
    }
    float xn = cf.sqrt_acp_prev * x0v + cf.dir_coef * noise;
    if (p.z) xn += cf.noise_coef * final_noise(p)[idx];
    p.x[idx] = xn;
  }
}

}  // namespace cdc
