// Final LayerNorm + 7x7 convolution (64 -> channels <= 3) + DDIM update on the Blackwell tensor cores.
// Same formulation as final_conv_kx_kernel (misc.cuh: horizontal taps folded into N,
//     Y[y, hx, kx*nch + ch] = sum_{ky, c} LN(in)[y + ky - 3, hx, c] * w[ch, c, ky, kx] ),
// but the GEMM runs as tcgen05.mma from shared memory into TMEM: the mma.sync form is bound by the legacy tensor path
// (2016 m16n8k16 per CTA).  The LayerNorm-ed halo is stored with a row pitch of 24 pixels, so that
//   * the 16 x 24 = 384 output-row-major pixels are exactly three M = 128 tiles,
//   * a vertical tap is a row offset of 24*ky pixels = whole 8-row swizzle atoms: the same K-major SWIZZLE_128B tile
//     serves all 7 taps through the descriptor start address.
// 84 UMMAs (M = 128, N = 32, K = 16) per CTA issued by one thread; accumulators: 3 x 32 TMEM columns.
#pragma once
#include "igemm_tc.cuh"
#include "misc.cuh"

namespace cdc {

constexpr int kFinalPitch = 24;                                  // halo row pitch (pixels): 22 used
constexpr int kFinalInBytes = kFinalHalo * kFinalPitch * 128;    // 67584
constexpr int kFinalW3Bytes = 7 * 32 * 128;                      // [ky][32 rows n][64 c] fp16, pre-swizzled (SWIZZLE_128B)
constexpr int kFinalSmemBytes3 = 1024 + kFinalInBytes + kFinalW3Bytes + 64;   // barriers + TMEM address word at the end

// PRELN: the input is already LayerNorm-ed fp16 (written by the last Upsample's epilogue, TcConvParams::ln_out): the
// whole 22 x 24 x 64 halo tile is ONE TMA box (conv zero padding = out-of-bounds zero fill) and the kernel has no
// normalisation phase at all.
template <bool PRELN>
__global__ void __launch_bounds__(256) final_conv_tc_kernel(const FinalParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* sIn = smem;
  float* sY = reinterpret_cast<float*>(smem);   // reuses the halo tile once the MMAs are done
  const uint32_t sW32 = base + kFinalInBytes;
  const uint32_t bar = sW32 + kFinalW3Bytes;    // MMA completion
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kFinalInBytes + kFinalW3Bytes + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;

  pdl_launch_dependents();
  {   // weights: the blob already holds the swizzled shared-memory image
    for (int i = tid; i < kFinalW3Bytes / 16; i += 256)
      cp_async16(sW32 + i * 16, reinterpret_cast<const uint4*>(p.Wf) + i, 16);
    cp_async_commit();
  }
  const uint32_t bar_in = bar + 16;             // halo tile landed (PRELN)
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init(bar_in, 1);
    tc::fence_barrier_init();
    if (PRELN) tc::prefetch_tmap(&tmap);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  __syncthreads();   // barriers initialised before anyone polls them
  pdl_wait();
  if (PRELN) {
    if (tid == 0) {
      tc::mbar_expect_tx(bar_in, (uint32_t)kFinalInBytes);
      tc::tma_load_4d(base, &tmap, bar_in, 0, x0 - 3, y0 - 3, b);
    }
  } else
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
          *reinterpret_cast<uint4*>(sIn + swz128((hp / kFinalHalo) * kFinalPitch + hp % kFinalHalo, j)) = o;
        }
      }
    }
  }
  cp_async_wait<0>();
  if (PRELN) tc::mbar_wait(bar_in, 0);
  tc::fence_proxy_async();   // weights (and, without PRELN, the halo tile) were written through the generic proxy
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_f16(32);
      const uint32_t desc_hi = (uint32_t)(tc::make_desc_sw128(0) >> 32);
#pragma unroll 1
      for (int t = 0; t < 3; ++t)
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky) {
          const uint32_t a_lo = (uint32_t)tc::make_desc_sw128(base + (t * 128 + kFinalPitch * ky) * 128);
          const uint32_t b_lo = (uint32_t)tc::make_desc_sw128(sW32 + ky * 4096);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc::umma_f16_lo(tmem_base + t * 32, a_lo + ks * 2, b_lo + ks * 2, desc_hi, idesc, (ky > 0 || ks > 0) ? 1u : 0u);
        }
      tc::umma_commit(bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(bar, 0);
  tc::tc_fence_after();
  // the halo tile is dead (all MMAs retired): its memory becomes Y[384][26]
  {
    const int quad = warp & 3;
#pragma unroll 1
    for (int t = (warp >> 2); t < 3; t += 2) {   // warps 0-3: tiles 0 and 2, warps 4-7: tile 1
      uint32_t v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * 32), v);
      float* dst = sY + (t * 128 + quad * 32 + lane) * kFinalYStride;
#pragma unroll
      for (int n = 0; n < 24; n += 2) *reinterpret_cast<float2*>(dst + n) = make_float2(__uint_as_float(v[n]), __uint_as_float(v[n + 1]));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }

  // ---- horizontal taps + bias + sampler update: one output pixel per thread ----
  const int ty = tid >> 4, tx = tid & 15;
  const int yy = y0 + ty, xx = x0 + tx;
  if (yy >= p.H || xx >= p.W) return;
  const int nch = p.channels;
  cdc_step_coef cf = {};
  if (p.mode == 1) cf = p.table[*p.step_ptr];
  const float* yrow = sY + (ty * kFinalPitch + tx) * kFinalYStride;
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

// ------------------------------------------------------------------------------------------------
// Persistent form (round 2): one CTA per SM walks the 16 x 16 output tiles; the weights are loaded once, the halo tiles
// are double-buffered (the TMA load of tile k+2 starts when the MMAs of tile k retire), the accumulators are
// double-buffered in TMEM (the MMAs of tile k+1 run under the epilogue of tile k) and Y has its own staging area.
// Warps 0-7: epilogue (TMEM -> Y -> horizontal taps + bias + sampler update), warp 8: TMA producer + MMA issuer.
// The one-tile-per-CTA form above spent each CTA's life in load -> MMA -> epilogue, serialised (52 us at B = 8, 256 x 256).
// ------------------------------------------------------------------------------------------------
constexpr int kFinalPersistThreads = 288;
constexpr int kFinalYBytes = 3 * 128 * kFinalYStride * 4;
constexpr int kFinalSmemBytesP = 1024 + 2 * kFinalInBytes + kFinalW3Bytes + kFinalYBytes + 128;

__global__ void __launch_bounds__(kFinalPersistThreads, 1)
final_conv_tc_persist_kernel(const FinalParams p, const __grid_constant__ CUtensorMap tmap, int tiles_x, int tiles_y,
                             int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sW32 = base + 2 * kFinalInBytes;
  float* sY = reinterpret_cast<float*>(smem + 2 * kFinalInBytes + kFinalW3Bytes);
  const uint32_t bars = base + 2 * kFinalInBytes + kFinalW3Bytes + kFinalYBytes;
  const uint32_t bar_full = bars, bar_done = bars + 16, bar_tempty = bars + 32;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + (bars + 48 - base));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nk = total_tiles > (int)blockIdx.x ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  pdl_launch_dependents();
  for (int i = tid; i < kFinalW3Bytes / 16; i += kFinalPersistThreads)
    cp_async16(sW32 + i * 16, reinterpret_cast<const uint4*>(p.Wf) + i, 16);   // weights: constant, may precede the wait
  cp_async_commit();
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(bar_full + 8 * i, 1);
      tc::mbar_init(bar_done + 8 * i, 1);
      tc::mbar_init(bar_tempty + 8 * i, 256);
    }
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmap);
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  cp_async_wait<0>();
  tc::fence_proxy_async();   // the weights were written through the generic proxy
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  pdl_wait();

  auto tile_coords = [&](int k, int& b, int& y0, int& x0) {
    const int t = (int)blockIdx.x + k * (int)gridDim.x;
    const int bx = t % tiles_x, r = t / tiles_x;
    x0 = bx * 16;
    y0 = (r % tiles_y) * 16;
    b = r / tiles_y;
  };

  if (warp == 8) {
    // =============================== TMA producer + MMA issuer ===============================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_f16(32);
      const uint32_t desc_hi = (uint32_t)(tc::make_desc_sw128(0) >> 32);
      auto load = [&](int k) {
        int b, y0, x0;
        tile_coords(k, b, y0, x0);
        const uint32_t full = bar_full + 8 * (k & 1);
        tc::mbar_expect_tx(full, (uint32_t)kFinalInBytes);
        tc::tma_load_4d(base + (k & 1) * kFinalInBytes, &tmap, full, 0, x0 - 3, y0 - 3, b);
      };
      for (int k = 0; k < nk && k < 2; ++k) load(k);
      for (int k = 0; k < nk; ++k) {
        const int buf = k & 1;
        const uint32_t use = (uint32_t)(k >> 1);
        tc::mbar_wait(bar_full + 8 * buf, use & 1u);
        if (k >= 2) tc::mbar_wait(bar_tempty + 8 * buf, (use - 1u) & 1u);   // the epilogue has read tile k-2's accumulators
        tc::tc_fence_after();
        const uint32_t hbase = base + buf * kFinalInBytes;
#pragma unroll 1
        for (int t = 0; t < 3; ++t)
#pragma unroll 1
          for (int ky = 0; ky < 7; ++ky) {
            const uint32_t a_lo = (uint32_t)tc::make_desc_sw128(hbase + (t * 128 + kFinalPitch * ky) * 128);
            const uint32_t b_lo = (uint32_t)tc::make_desc_sw128(sW32 + ky * 4096);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::umma_f16_lo(tmem_base + buf * 96 + t * 32, a_lo + ks * 2, b_lo + ks * 2, desc_hi, idesc,
                              (ky > 0 || ks > 0) ? 1u : 0u);
          }
        tc::umma_commit(bar_done + 8 * buf);
        if (k + 2 < nk) {   // this halo buffer is free once the MMAs above retire
          tc::mbar_wait(bar_done + 8 * buf, use & 1u);
          load(k + 2);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue: 256 threads ===============================
    const int quad = warp & 3;
    const int ty = tid >> 4, tx = tid & 15;
    const int nch = p.channels;
    cdc_step_coef cf = {};
    if (p.mode == 1) cf = p.table[*p.step_ptr];
    for (int k = 0; k < nk; ++k) {
      const int buf = k & 1;
      const uint32_t use = (uint32_t)(k >> 1);
      int b, y0, x0;
      tile_coords(k, b, y0, x0);
      tc::mbar_wait(bar_done + 8 * buf, use & 1u);
      tc::tc_fence_after();
#pragma unroll 1
      for (int t = (warp >> 2); t < 3; t += 2) {   // warps 0-3: tiles 0 and 2, warps 4-7: tile 1
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 96 + t * 32), v);
        float* dst = sY + (t * 128 + quad * 32 + lane) * kFinalYStride;
#pragma unroll
        for (int n = 0; n < 24; n += 2)
          *reinterpret_cast<float2*>(dst + n) = make_float2(__uint_as_float(v[n]), __uint_as_float(v[n + 1]));
      }
      tc::tc_fence_before();
      tc::mbar_arrive(bar_tempty + 8 * buf);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // Y complete
      const int yy = y0 + ty, xx = x0 + tx;
      const float* yrow = sY + (ty * kFinalPitch + tx) * kFinalYStride;
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
      asm volatile("bar.sync 1, 256;" ::: "memory");   // Y consumed before the next tile overwrites it
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

}  // namespace cdc
