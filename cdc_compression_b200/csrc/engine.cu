// CDC denoiser engine: weight registry + repack, layer plan, workspace arena, launches, C ABI.
//
// Mirrors the reference's module structure (epsilonparam/modules/unet.py:18-124,
// xparam/modules/unet.py:19-135) as a flat list of kernel launches:
//   pack_input -> [ResnetBlock x2 -> LinearAttention -> Downsample] x levels -> mid -> [cat skip ->
//   ResnetBlock x2 -> LinearAttention -> Upsample] x levels -> LayerNorm+conv7x7 (+DDIM update)
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/cdc_b200.h"
#include "attn.cuh"
#include "igemm_hmma.cuh"
#include "igemm_tc.cuh"
#include "attn_tc.cuh"
#include "attn_alg_tc.cuh"
#include "misc.cuh"
#include "final_tc.cuh"

using namespace cdc;

namespace {

std::string g_create_error;

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

// ---------------------------------------------------------------- weight blob
struct Blob {
  std::vector<uint8_t> host;
  size_t reserve(size_t bytes) {
    size_t off = (host.size() + 255) & ~size_t(255);
    host.resize(off + bytes, 0);
    return off;
  }
  template <typename T>
  T* at(size_t off) { return reinterpret_cast<T*>(host.data() + off); }
};

struct ConvW {      // one convolution-shaped weight set in kernel layout
  size_t w = 0;     // fp16 [chunk][N][64] (x4 phases for transposed conv)
  size_t bias = 0;  // fp32 [N]
  int nchunks = 0;
  int N = 0;
  double macs_per_row = 0;  // algorithmic MACs per output row (pixel)
};
struct BlockW {
  ConvW conv;
  size_t g = 0, b = 0;
};
struct ResW {
  BlockW b1, b2;
  bool has_res = false;
  ConvW res;
  int shift_off = 0;  // offset of this block's rows in the concatenated timestep-MLP output
  int cin = 0, cout = 0;
};
struct AttnW {
  int C = 0;
  size_t wkv = 0, u = 0, c = 0;          // fp16 [C/64][2C][64], fp32 [2C], fp32 [2C]
  size_t wq = 0, woT = 0;                // fp32 [C][C] (scale folded), fp32 [C][C] transposed
  size_t wq16h = 0, wq16l = 0;           // the same two matrices as fp16 value + remainder in the MN-blocked operand layout of
  size_t wo16h = 0, wo16l = 0;           // attn_alg_tc_kernel: element (k, mn) at ((mn/64) * C + k) * 64 + mn % 64
  size_t g = 0, bln = 0, bout = 0;       // fp32 [C]
};

struct Level {
  ResW rb0, rb1;
  AttnW attn;
  bool has_resample = false;
  ConvW resample;
  int cin = 0, cout = 0;
};

// ---------------------------------------------------------------- plan
enum OpKind { OP_PACK, OP_CONV, OP_TIME, OP_ATTN_CTX, OP_COMBINE, OP_SGEMM, OP_FINISH, OP_FINAL, OP_ADVANCE, OP_LNROWS,
              OP_LATENT, OP_UPSMALL, OP_ALG };

struct Op {
  int kind = 0;
  std::string name;
  // conv
  ConvParams conv{};
  int bm = 0, bn = 0, epi = 0;
  dim3 grid{1, 1, 1};
  // attention
  AttnCtxParams actx{};
  AttnTcParams atc{};      // tcgen05 form of the context kernel (use_tc; tensor maps in maps.a[0] / maps.b[0])
  int atc_smem = 0;
  struct { const float *pc, *pm, *ps; int C, nchunks; float* out; __half *o16h, *o16l; } comb{};
  struct { const float *At, *Bm; float* Cout; int M, N, K; long long sA, sB, sC; } sg{};
  GemmFinish gfin{};       // OP_SGEMM: fused finish epilogue (second attention product)
  struct { const float *Mf, *g, *bln, *bout; int C; __half* Mg; float *um, *cm; } fin{};
  // tcgen05 path (stride-1 convolutions when the engine's mainloop is 1)
  bool use_tc = false;
  TcMaps maps;
  TcConvParams tcp{};
  TcVecs64 vecs64{};       // C_out == 64 LayerNorm layers: bias / gain / offset passed by value
  int tc_grid = 0, tc_smem = 0, tc_occ = 1;
  // stream lane: 1 = runs on the engine's side stream concurrently with the main-lane ops that follow it (the 1x1
  // res_conv of a ResnetBlock overlaps block1); join_before = main lane must wait for the side lane first
  int lane = 0;
  bool join_before = false;
  LnRowsParams lnr{};   // OP_LNROWS: second half of a sliced convolution
  AlgTcParams alg{};    // OP_ALG: one C x C product of the attention algebra on tcgen05 (maps.a[0..1] = A hi/lo, maps.b[0..1] = B hi/lo)
  struct { const __half *ah, *al, *bh, *bl; } algsrc{};
  UpSmallParams ups{};  // OP_UPSMALL: last Upsample of the eps context decoder (C_out <= 8, fp32 NCHW output)
  struct { __half *hi, *lo; int C, HW; } lat{};   // OP_LATENT: quantised latent fp32 NCHW -> NHWC hi + lo
  long long ctr_index = -1;   // K-split convolution with the fused finish: first of its 2 x tiles arrival counters
  // debug view of the op's fp16 NHWC output (if any)
  const __half* dbg = nullptr;
  int dC = 0, dH = 0, dW = 0;
  double flops = 0;
};

struct Arena {
  size_t top = 0, peak = 0;
  bool no_reuse = false;
  std::vector<std::pair<size_t, size_t>> free_list;  // (offset, size), sorted by offset
  size_t alloc(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    for (size_t i = 0; i < free_list.size(); ++i) {
      if (free_list[i].second >= bytes) {
        size_t off = free_list[i].first;
        if (free_list[i].second == bytes) free_list.erase(free_list.begin() + i);
        else { free_list[i].first += bytes; free_list[i].second -= bytes; }
        return off;
      }
    }
    size_t off = top;
    top += bytes;
    peak = std::max(peak, top);
    return off;
  }
  void release(size_t off, size_t bytes) {
    if (no_reuse) return;
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    auto it = std::lower_bound(free_list.begin(), free_list.end(), std::make_pair(off, size_t(0)));
    it = free_list.insert(it, {off, bytes});
    // coalesce with next / previous
    size_t i = it - free_list.begin();
    if (i + 1 < free_list.size() && free_list[i].first + free_list[i].second == free_list[i + 1].first) {
      free_list[i].second += free_list[i + 1].second;
      free_list.erase(free_list.begin() + i + 1);
    }
    if (i > 0 && free_list[i - 1].first + free_list[i - 1].second == free_list[i].first) {
      free_list[i - 1].second += free_list[i].second;
      free_list.erase(free_list.begin() + i);
      --i;
    }
    if (free_list[i].first + free_list[i].second == top) {
      top = free_list[i].first;
      free_list.erase(free_list.begin() + i);
    }
  }
};

struct Act {  // fp16 NHWC activation in the workspace (+ optional compensation tensor of the same shape)
  size_t off = 0, bytes = 0;
  size_t off_lo = 0;  // 0 = none.  value = hi + lo, lo = fp16(value - hi): kept for the "trunk" activations
  int C = 0, H = 0, W = 0;
  bool win8 = false;   // packed network input in window form (ConvSeg::win8)
};

struct Plan {
  int B = 0, H = 0, W = 0;
  uint8_t* ws = nullptr;
  size_t total_bytes = 0;
  std::vector<Op> ops;
  std::vector<Op> ctx_ops;         // context_fn.decode: latent -> the four context maps of the persistent region
  // persistent region
  std::vector<size_t> ctx_off;     // fp16 NHWC context per level (level 0 of the eps variant: fp32 NCHW copy)
  size_t shifts_off = 0;
  size_t xstate_off = 0;           // fp32 NCHW sampler state the captured graph works on
  size_t raw_off = 0, raw_bytes = 0;   // fp32 partial tiles of the sliced convolutions (after the arena)
  size_t raw2_off = 0, raw2_bytes = 0; // same, for side-lane ops (they overlap main-lane sliced convolutions)
  size_t ctr_off = 0;                  // per-tile arrival / departure counters of the K-split convolutions (fused finish)
  long long ctr_count = 0;
  int pack_op = -1, time_op = -1, final_op = -1;
  double flops = 0;
  // captured step graph
  cudaGraphExec_t graph = nullptr;
  int graph_pred = -1, graph_clip = -1;
  const float* graph_z = nullptr;   // noise buffer the graph was captured with (null: eta == 0 loop)
};

}  // namespace

struct RunArgs {
  const float* x = nullptr;     // network input (fp32 NCHW)
  const float* time = nullptr;  // [B] or null (use schedule table)
  float* out = nullptr;         // mode 0
  float* x_inout = nullptr;     // mode 1
  const float* z = nullptr;
  bool z_chunk = false;            // z holds the noise of consecutive steps (indexed from e->d_zfirst inside the kernels)
  const float* latent = nullptr;   // context decode: quantised latent, fp32 NCHW
  int mode = 0, pred = 0, clip = 0;
  bool advance = false;
};

struct cdc_engine {
  cdc_config cfg{};
  int device = 0;
  std::string err;
  std::map<std::string, HostTensor> weights;
  bool finalized = false;
  bool dry = false;  // created with device == -1: planning only
  bool debug_no_reuse = false;
  int mainloop = 1;   // 0 = mma.sync kernels, 1 = tcgen05/TMA kernels for stride-1 convolutions (default)
  int num_sms = 148;
  int vreuse = 1;      // CDC_VREUSE: 0 off, 1 when it fits the default occupancy, 2 also at one CTA per SM
  bool sliced = true;  // CDC_SLICED=0 disables the sliced low-resolution mode (A/B measurements)
  bool final_kx = true; // final conv with the horizontal taps folded into N; CDC_FINAL_KX=0: 49-tap form
  bool has_f_w2 = false;
  size_t f_w2 = 0, f_w3 = 0;
  size_t w_ident = 0;     // [64][64] fp16 identity "weights": identity residual of a 64-channel block through the MMA
  bool final_preln = true;  // last Upsample writes the LayerNorm-ed fp16 input of the final conv; CDC_FINAL_PRELN=0: final conv normalises
  bool final_tc = true; // tcgen05 form of the final conv (final_tc.cuh); CDC_FINAL_TC=0: mma.sync form
  bool final_persist = true;  // persistent, double-buffered tcgen05 final conv; CDC_FINAL_PERSIST=0: one tile per CTA
  bool fold_finish = true;   // attn_finish_kernel fused into the second C x C product; CDC_FOLD_FINISH=0: separate kernel
  bool dual_pass = true;   // W_hi / W_lo passes of the 3-pass convolutions share one activation load; CDC_DUAL_PASS=0: separate
  int slice_max_tiles = 100;   // layers with fewer 128-pixel output tiles (at the nominal batch of 8) run in sliced mode
  bool fuse_res = true;   // res_conv folded into block2 (second TMEM accumulator); CDC_FUSE_RES=0: separate launch
  bool attn_area = true;   // attention pixel chunks grow with the image area (CDC_ATTN_AREA=0: fixed count, round-1 behaviour)
  bool alg_tc = true;   // attention C x C products on tcgen05 (attn_alg_tc.cuh); CDC_ALG_TC=0: split-fp16 mma.sync kernel
  bool attn_tc = true;  // tcgen05 attention-context kernel (attn_tc.cuh); CDC_ATTN_TC=0: mma.sync kernel of attn.cuh
  bool nslice = true;  // fused column slices + cluster LayerNorm exchange; CDC_NSLICE=0: K-split fp32 partials + ln_rows_kernel
  bool fuse_lnrows = false;  // CDC_FUSE_LNROWS=1: K-split convolutions finish their rows in the same launch (per-tile arrival counters)
                             // instead of ln_rows_kernel.  Measured (B=8 256x256): 16 launches fewer but 2.653 vs 2.633 ms/step —
                             // fence + arrival atomics + the wait for the tile's slowest unit cost what the launch did; off.
  int slice_slots = 148, slice_kmax = 64;   // tuning knobs (CDC_SLICE_SLOTS / CDC_SLICE_KMAX)
  bool profiling = false;                   // cdc_engine_profile_ops in progress: launches are timed alone, without PDL overlap
  int pdl_mode = 1;                         // CDC_PDL: programmatic dependent launch policy (see pdl_for)
  // derived structure
  std::vector<int> dims, cdims;
  bool fold_ctx0 = false;  // small level-0 context folded into the packed input (eps demo)
  int R = 0;               // total timestep-shift rows
  // weights
  Blob blob;
  uint8_t* dblob = nullptr;
  size_t dblob_bytes = 0;
  std::vector<Level> downs, ups;
  ResW mid1, mid2;
  AttnW mid_attn;
  struct CtxStage {          // context_fn.dec.i = [ResnetBlock (no time embedding), (Identity), Upsample(C_mid -> C_out)]
    ResW rb;
    ConvW up;
    int cin = 0, cmid = 0, cout = 0;
    bool small_up = false;   // C_out <= 8: fp32 SIMT kernel, weights [ky][kx][c][o] at up_w32
    size_t up_w32 = 0, up_b32 = 0;
  };
  std::vector<CtxStage> ctxdec;
  size_t t_w1 = 0, t_b1 = 0, t_w2 = 0, t_b2 = 0, t_wcat = 0, t_bcat = 0;
  size_t f_g = 0, f_b = 0, f_w = 0, f_bias = 0;
  // plans
  std::map<std::tuple<int, int, int, uintptr_t>, std::unique_ptr<Plan>> plans;
  // schedule
  cdc_step_coef* d_table = nullptr;
  int table_cap = 0, S = 0;
  cdc_step_coef* h_table = nullptr;   // pinned staging copy of the schedule (cdc_set_schedule copies asynchronously)
  int h_table_cap = 0;
  cudaEvent_t ev_table = nullptr;
  int* d_step = nullptr;
  int* d_zfirst = nullptr;   // schedule index whose noise sits first in the caller's z buffer (eta != 0 loops)
  // stream the sampling loop runs on (the caller's stream may be the legacy default stream, which
  // cannot be captured); joined to the caller's stream with events on both sides
  cudaStream_t loop_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaStream_t side_stream = nullptr;   // second lane of the step (see Op::lane)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool two_lanes = true;                // CDC_TWO_LANES=0: everything on one stream
  bool lnrows_hoist = true;             // small-grid ln_rows_kernel loads its vectors before the partials (CDC_LNROWS_HOIST=0: compact form)
  bool time_lane = true;                // timestep MLP on the side lane (overlaps pack_input); CDC_TIME_LANE=0: main lane
  // context state
  bool ctx_set = false;
  int ctx_B = 0, ctx_H = 0, ctx_W = 0;
  uintptr_t ctx_ws = 0;
  Plan* last_plan = nullptr;
  RunArgs last_args;
};

namespace {

int fail(cdc_engine* e, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  else g_create_error = buf;
  return code;
}

#define CUDA_TRY(e, expr)                                                                         \
  do {                                                                                            \
    cudaError_t _err = (expr);                                                                    \
    if (_err != cudaSuccess)                                                                      \
      return fail(e, CDC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_err), __FILE__, __LINE__); \
  } while (0)

// Every C-ABI entry runs on the engine's device and leaves the calling thread's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (dev < 0) return;   // planning-only engine: no device work
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

template <typename T>
T* dptr(cdc_engine* e, size_t off) { return reinterpret_cast<T*>(e->dblob + off); }

// ---------------------------------------------------------------- weight lookup helpers
const HostTensor* find(cdc_engine* e, const std::string& key, std::initializer_list<int64_t> shape, int* rc) {
  auto it = e->weights.find(key);
  if (it == e->weights.end()) {
    *rc = fail(e, CDC_ERR_MISSING, "missing weight '%s'", key.c_str());
    return nullptr;
  }
  const HostTensor& t = it->second;
  std::vector<int64_t> want(shape);
  if (t.shape != want) {
    std::string got, exp;
    for (auto v : t.shape) got += std::to_string(v) + ",";
    for (auto v : want) exp += std::to_string(v) + ",";
    *rc = fail(e, CDC_ERR_MISSING, "weight '%s' has shape [%s] expected [%s]", key.c_str(), got.c_str(), exp.c_str());
    return nullptr;
  }
  return &t;
}

size_t put_f32(cdc_engine* e, const float* src, size_t n) {
  size_t off = e->blob.reserve(n * 4);
  memcpy(e->blob.at<float>(off), src, n * 4);
  return off;
}

struct SegSpec {  // how one K-segment of a conv maps onto the reference's OIHW input channels
  int C;          // padded channel count of the segment tensor (multiple of 64)
  int kh, kw;
  int mode;       // 0: channel k of chunk cc at tap (ty,tx) <- W[o][coff + cc*64+k][ty][tx]
                  // 1: packed input X0: k = kx*8 + c (c < creal) at vertical tap ty <- W[o][coff + c][ty][kx]  (kw==1)
                  // 2: packed input X0, 1x1 conv: k = 3*8 + c <- W[o][coff + c][0][0]
                  // 3: (unused since the packed input is in window form, where slot 7 is pixel x+4 and not a second copy of
                  //    the centre pixel) like 2, plus the weight remainder fp16(w - fp16(w)) at k = 7*8 + c
  int coff;       // first reference input channel of this segment
  int creal;      // real channels (mode 1/2)
  int part = 0;   // 0: fp16(w)   1: fp16(w - fp16(w))  (weight compensation term of the 3-pass trunk convolutions)
};

// the three passes  x_hi*W_hi + x_lo*W_hi + x_hi*W_lo  of a precision-critical ("trunk") convolution
std::vector<SegSpec> three_pass(const std::vector<SegSpec>& base, const std::vector<bool>& has_lo) {
  std::vector<SegSpec> out = base;
  for (size_t i = 0; i < base.size(); ++i)
    if (has_lo[i]) out.push_back(base[i]);
  for (size_t i = 0; i < base.size(); ++i) {
    if (base[i].mode == 3) continue;   // remainder already folded into the same chunk
    SegSpec s = base[i];
    s.part = 1;
    out.push_back(s);
  }
  return out;
}

// OIHW conv weight -> [chunk][N][64] fp16
int pack_conv(cdc_engine* e, const std::string& key, int N, int Cin_ref, int KH, int KW,
              const std::vector<SegSpec>& segs, bool with_bias, ConvW* out) {
  int rc = 0;
  const HostTensor* w = find(e, key + ".weight", {N, Cin_ref, KH, KW}, &rc);
  if (!w) return rc;
  int nchunks = 0;
  for (auto& s : segs) nchunks += s.kh * s.kw * (s.C / 64);
  out->nchunks = nchunks;
  out->N = N;
  out->macs_per_row = (double)N * Cin_ref * KH * KW;
  out->w = e->blob.reserve((size_t)nchunks * N * 64 * 2);
  std::vector<__half> tmp((size_t)nchunks * N * 64, __float2half(0.f));
  int q = 0;
  for (auto& s : segs) {
    for (int ty = 0; ty < s.kh; ++ty)
      for (int tx = 0; tx < s.kw; ++tx)
        for (int cc = 0; cc < s.C / 64; ++cc, ++q)
          for (int o = 0; o < N; ++o)
            for (int k = 0; k < 64; ++k) {
              float v = 0.f;
              if (s.mode == 0) {
                const int c = s.coff + cc * 64 + k;
                v = w->data[(((size_t)o * Cin_ref + c) * KH + ty) * KW + tx];
              } else if (s.mode == 1) {
                const int kx = k >> 3, c = k & 7;
                if (kx < KW && c < s.creal) v = w->data[(((size_t)o * Cin_ref + s.coff + c) * KH + ty) * KW + kx];
              } else {
                const int kx = k >> 3, c = k & 7;
                if ((kx == 3 || (s.mode == 3 && kx == 7)) && c < s.creal) v = w->data[((size_t)o * Cin_ref + s.coff + c)];
              }
              __half hv = __float2half_rn(v);
              if (s.part == 1 || (s.mode == 3 && (k >> 3) == 7)) hv = __float2half_rn(v - __half2float(hv));
              tmp[((size_t)q * N + o) * 64 + k] = hv;
            }
  }
  memcpy(e->blob.at<__half>(out->w), tmp.data(), tmp.size() * 2);
  if (with_bias) {
    const HostTensor* b = find(e, key + ".bias", {N}, &rc);
    if (!b) return rc;
    out->bias = put_f32(e, b->data.data(), N);
  }
  return 0;
}

// ConvTranspose2d(C,C,4,2,1) weight [Cin][Cout][4][4] -> 4 phases x [tap(2x2)*Cin/64][Cout][64]
int pack_convT(cdc_engine* e, const std::string& key, int Cin, int Cout, ConvW* out) {
  int rc = 0;
  const HostTensor* w = find(e, key + ".weight", {Cin, Cout, 4, 4}, &rc);
  if (!w) return rc;
  const HostTensor* b = find(e, key + ".bias", {Cout}, &rc);
  if (!b) return rc;
  const int cpt = Cin / 64;
  out->nchunks = 3 * 4 * cpt;  // three passes: (x_hi, W_hi), (x_lo, W_hi), (x_hi, W_lo)
  out->N = Cout;
  out->macs_per_row = (double)Cout * Cin * 4;  // per OUTPUT pixel: 2x2 taps
  std::vector<__half> tmp((size_t)4 * out->nchunks * Cout * 64);
  for (int z = 0; z < 4; ++z) {
    const int py = z >> 1, px = z & 1;
    int q = 0;
    for (int pass = 0; pass < 3; ++pass)
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx)
          for (int cc = 0; cc < cpt; ++cc, ++q) {
            const int ky = 3 - 2 * ty - py, kx = 3 - 2 * tx - px;
            for (int o = 0; o < Cout; ++o)
              for (int k = 0; k < 64; ++k) {
                const int c = cc * 64 + k;
                const float v = w->data[(((size_t)c * Cout + o) * 4 + ky) * 4 + kx];
                __half hv = __float2half_rn(v);
                if (pass == 2) hv = __float2half_rn(v - __half2float(hv));
                tmp[(((size_t)z * out->nchunks + q) * Cout + o) * 64 + k] = hv;
              }
          }
  }
  out->w = e->blob.reserve(tmp.size() * 2);
  memcpy(e->blob.at<__half>(out->w), tmp.data(), tmp.size() * 2);
  out->bias = put_f32(e, b->data.data(), Cout);
  return 0;
}

int pack_ln(cdc_engine* e, const std::string& key, int C, size_t* g, size_t* b) {
  int rc = 0;
  const HostTensor* tg = find(e, key + ".g", {1, C, 1, 1}, &rc);
  if (!tg) return rc;
  const HostTensor* tb = find(e, key + ".b", {1, C, 1, 1}, &rc);
  if (!tb) return rc;
  *g = put_f32(e, tg->data.data(), C);
  *b = put_f32(e, tb->data.data(), C);
  return 0;
}

int pack_resnet(cdc_engine* e, const std::string& p, int cin, int cout, int k1,
                const std::vector<SegSpec>& segs1, const std::vector<SegSpec>& segs_res_base,
                const std::vector<bool>& res_has_lo, ResW* out, std::vector<float>* wcat, std::vector<float>* bcat) {
  const std::vector<SegSpec> segs_res = three_pass(segs_res_base, res_has_lo);
  int rc;
  out->cin = cin;
  out->cout = cout;
  if ((rc = pack_conv(e, p + "block1.block.0", cout, cin, k1, k1, segs1, true, &out->b1.conv))) return rc;
  if ((rc = pack_ln(e, p + "block1.block.1", cout, &out->b1.g, &out->b1.b))) return rc;
  std::vector<SegSpec> s2 = {{cout, 3, 3, 0, 0, 0}};
  if ((rc = pack_conv(e, p + "block2.block.0", cout, cout, 3, 3, s2, true, &out->b2.conv))) return rc;
  if ((rc = pack_ln(e, p + "block2.block.1", cout, &out->b2.g, &out->b2.b))) return rc;
  out->has_res = cin != cout;
  if (out->has_res) {
    if ((rc = pack_conv(e, p + "res_conv", cout, cin, 1, 1, segs_res, true, &out->res))) return rc;
  }
  if (!wcat) {   // ResnetBlock built without a time embedding (context decoder): no shift
    out->shift_off = -1;
    return 0;
  }
  const int dim = e->cfg.dim;
  const HostTensor* mw = find(e, p + "mlp.1.weight", {cout, dim}, &rc);
  if (!mw) return rc;
  const HostTensor* mb = find(e, p + "mlp.1.bias", {cout}, &rc);
  if (!mb) return rc;
  out->shift_off = (int)bcat->size();
  wcat->insert(wcat->end(), mw->data.begin(), mw->data.end());
  bcat->insert(bcat->end(), mb->data.begin(), mb->data.end());
  return 0;
}

// context_fn.dec.* (BigCompressor / ResnetCompressor decoder, epsilonparam/modules/compress_modules.py:144-156,
// xparam/modules/compress_modules.py:142-151): optional — packed when the keys were registered.
int pack_context_decoder(cdc_engine* e) {
  e->ctxdec.clear();
  const std::string pre = "context_fn.dec.";
  int n = 0;
  while (e->weights.count(pre + std::to_string(n) + ".0.block1.block.0.weight")) ++n;
  if (n == 0) return 0;
  const cdc_config& cfg = e->cfg;
  if (n != cfg.n_context)
    return fail(e, CDC_ERR_UNSUPPORTED, "context decoder has %d stages, the Unet consumes %d context maps", n, cfg.n_context);
  for (int i = 0; i < n; ++i) {
    const std::string p = pre + std::to_string(i) + ".";
    if (e->weights.count(p + "1.scale.weight"))
      return fail(e, CDC_ERR_UNSUPPORTED, "context decoder with VBRCondition (vbr=True) is not implemented by the engine");
    cdc_engine::CtxStage st;
    const HostTensor& w1 = e->weights[p + "0.block1.block.0.weight"];
    if (w1.shape.size() != 4 || w1.shape[2] != 3) return fail(e, CDC_ERR_UNSUPPORTED, "unexpected shape of %s0.block1", p.c_str());
    st.cmid = (int)w1.shape[0];
    st.cin = (int)w1.shape[1];
    const std::string upk = e->weights.count(p + "2.conv.weight") ? p + "2.conv" : p + "1.conv";
    auto it = e->weights.find(upk + ".weight");
    if (it == e->weights.end()) return fail(e, CDC_ERR_MISSING, "missing weight '%s.weight'", upk.c_str());
    if (it->second.shape.size() != 4 || it->second.shape[0] != st.cmid || it->second.shape[2] != 4)
      return fail(e, CDC_ERR_MISSING, "weight '%s.weight' has an unexpected shape", upk.c_str());
    st.cout = (int)it->second.shape[1];
    const int level = n - 1 - i;   // stage 0 produces the coarsest map
    if (st.cout != e->cdims[level])
      return fail(e, CDC_ERR_UNSUPPORTED, "context decoder stage %d produces %d channels, the Unet expects %d at level %d", i,
                  st.cout, e->cdims[level], level);
    if (st.cin % 64 || st.cmid % 64 || st.cin > 384 || st.cmid > 384)
      return fail(e, CDC_ERR_UNSUPPORTED, "context decoder widths %d -> %d unsupported (multiples of 64, <= 384)", st.cin, st.cmid);
    int rc;
    std::vector<SegSpec> s1 = {{st.cin, 3, 3, 0, 0, 0}}, sr = {{st.cin, 1, 1, 0, 0, 0}};
    if ((rc = pack_resnet(e, p + "0.", st.cin, st.cmid, 3, s1, sr, {true}, &st.rb, nullptr, nullptr))) return rc;
    if (st.cout % 64 == 0) {
      if ((rc = pack_convT(e, upk, st.cmid, st.cout, &st.up))) return rc;
    } else if (st.cout <= 8 && level == 0 && e->fold_ctx0) {
      const HostTensor& w = it->second;
      const HostTensor* b = find(e, upk + ".bias", {st.cout}, &rc);
      if (!b) return rc;
      std::vector<float> w32((size_t)16 * st.cmid * st.cout);
      for (int c = 0; c < st.cmid; ++c)
        for (int o = 0; o < st.cout; ++o)
          for (int ky = 0; ky < 4; ++ky)
            for (int kx = 0; kx < 4; ++kx)
              w32[((size_t)(ky * 4 + kx) * st.cmid + c) * st.cout + o] = w.data[(((size_t)c * st.cout + o) * 4 + ky) * 4 + kx];
      st.small_up = true;
      st.up_w32 = put_f32(e, w32.data(), w32.size());
      st.up_b32 = put_f32(e, b->data.data(), b->data.size());
    } else {
      return fail(e, CDC_ERR_UNSUPPORTED, "context decoder output width %d at level %d unsupported", st.cout, level);
    }
    e->ctxdec.push_back(st);
  }
  return 0;
}

int pack_attn(cdc_engine* e, const std::string& p, int C, AttnW* out) {
  int rc = 0;
  const HostTensor* g = find(e, p + "fn.norm.g", {1, C, 1, 1}, &rc);
  if (!g) return rc;
  const HostTensor* bl = find(e, p + "fn.norm.b", {1, C, 1, 1}, &rc);
  if (!bl) return rc;
  const HostTensor* wqkv = find(e, p + "fn.fn.to_qkv.weight", {3 * C, C, 1, 1}, &rc);
  if (!wqkv) return rc;
  const HostTensor* wo = find(e, p + "fn.fn.to_out.weight", {C, C, 1, 1}, &rc);
  if (!wo) return rc;
  const HostTensor* bo = find(e, p + "fn.fn.to_out.bias", {C}, &rc);
  if (!bo) return rc;
  out->C = C;
  const int cb = C / 64;
  std::vector<__half> wkv((size_t)cb * 2 * C * 64);
  std::vector<float> u(2 * C, 0.f), cv(2 * C, 0.f);
  for (int r = 0; r < 2 * C; ++r) {
    const float* row = &wqkv->data[(size_t)(C + r) * C];  // rows C..3C-1 of to_qkv = [k; v]
    double su = 0, sc = 0;
    for (int c = 0; c < C; ++c) {
      const __half hv = __float2half_rn(row[c] * g->data[c]);
      wkv[((size_t)(c >> 6) * 2 * C + r) * 64 + (c & 63)] = hv;
      su += __half2float(hv);
      sc += (double)row[c] * bl->data[c];
    }
    u[r] = (float)su;
    cv[r] = (float)sc;
  }
  out->wkv = e->blob.reserve(wkv.size() * 2);
  memcpy(e->blob.at<__half>(out->wkv), wkv.data(), wkv.size() * 2);
  out->u = put_f32(e, u.data(), u.size());
  out->c = put_f32(e, cv.data(), cv.size());
  const float scale = 1.0f / sqrtf((float)C);
  std::vector<float> wq((size_t)C * C), woT((size_t)C * C);
  for (int d = 0; d < C; ++d)
    for (int c = 0; c < C; ++c) wq[(size_t)d * C + c] = wqkv->data[(size_t)d * C + c] * scale;
  for (int o = 0; o < C; ++o)
    for (int ee = 0; ee < C; ++ee) woT[(size_t)ee * C + o] = wo->data[(size_t)o * C + ee];
  out->wq = put_f32(e, wq.data(), wq.size());
  out->woT = put_f32(e, woT.data(), woT.size());
  {   // MN-blocked fp16 value / remainder copies: wq is [K = d][MN = c], woT is [K = e][MN = o]
    std::vector<__half> h((size_t)C * C), l((size_t)C * C);
    auto blocked = [&](const std::vector<float>& src, size_t* oh, size_t* ol) {
      for (int k = 0; k < C; ++k)
        for (int mn = 0; mn < C; ++mn) {
          const float v = src[(size_t)k * C + mn];
          const __half hv = __float2half_rn(v);
          const size_t o = ((size_t)(mn >> 6) * C + k) * 64 + (mn & 63);
          h[o] = hv;
          l[o] = __float2half_rn(v - __half2float(hv));
        }
      *oh = e->blob.reserve(h.size() * 2);
      memcpy(e->blob.at<__half>(*oh), h.data(), h.size() * 2);
      *ol = e->blob.reserve(l.size() * 2);
      memcpy(e->blob.at<__half>(*ol), l.data(), l.size() * 2);
    };
    blocked(wq, &out->wq16h, &out->wq16l);
    blocked(woT, &out->wo16h, &out->wo16l);
  }
  out->g = put_f32(e, g->data.data(), C);
  out->bln = put_f32(e, bl->data.data(), C);
  out->bout = put_f32(e, bo->data.data(), C);
  return 0;
}

// ---------------------------------------------------------------- kernel dispatch
bool g_pdl = false;  // set per launch by run_op (pdl_for)

template <typename... KArgs, typename... Args>
cudaError_t launch_kc(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                      Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (g_pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {   // thread-block cluster along x (sliced LayerNorm convolutions)
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_kc(kern, grid, block, smem, st, 0, std::forward<Args>(args)...);
}

template <int BM, int BN, int EPI>
cudaError_t launch_igemm_t(const Op& op, cudaStream_t st) {
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    cudaError_t err = cudaFuncSetAttribute(igemm_hmma_kernel<BM, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           IgemmSmem<BM, BN>::kBytes);
    if (err != cudaSuccess) return err;
    attr_set[dev] = true;
  }
  return launch_k(igemm_hmma_kernel<BM, BN, EPI>, op.grid, dim3(256), IgemmSmem<BM, BN>::kBytes, st, op.conv);
}

cudaError_t launch_igemm(const Op& op, cudaStream_t st) {
#define CASE(BM_, BN_, EPI_) \
  if (op.bm == BM_ && op.bn == BN_ && op.epi == EPI_) return launch_igemm_t<BM_, BN_, EPI_>(op, st);
  CASE(128, 64, EPI_BIAS) CASE(128, 128, EPI_BIAS) CASE(128, 64, EPI_AFFINE) CASE(128, 128, EPI_AFFINE)
  CASE(128, 64, EPI_LN_SHIFT) CASE(128, 128, EPI_LN_SHIFT) CASE(64, 192, EPI_LN_SHIFT) CASE(64, 256, EPI_LN_SHIFT)
  CASE(64, 320, EPI_LN_SHIFT) CASE(64, 384, EPI_LN_SHIFT)
  CASE(128, 64, EPI_LN_RES) CASE(128, 128, EPI_LN_RES) CASE(64, 192, EPI_LN_RES) CASE(64, 256, EPI_LN_RES)
  CASE(64, 320, EPI_LN_RES) CASE(64, 384, EPI_LN_RES)
#undef CASE
  return cudaErrorInvalidValue;
}

template <int EPI, int OCC, int N64, bool RT = false>
cudaError_t launch_tc_t(const Op& op, cudaStream_t st) {
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    // OCC only sizes the register budget; a launch may still ask for a whole SM's shared memory (one CTA per SM)
    cudaError_t err = cudaFuncSetAttribute(igemm_tc_kernel<EPI, OCC, N64, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024);
    if (err != cudaSuccess) return err;
    attr_set[dev] = true;
  }
  if constexpr (N64 == 2)
    return launch_kc(igemm_tc_kernel<EPI, OCC, N64, RT>, dim3(op.tc_grid), dim3(kTcThreads), (size_t)op.tc_smem, st,
                     op.tcp.cluster_n, op.maps, op.tcp, op.vecs64);
  else
    return launch_kc(igemm_tc_kernel<EPI, OCC, N64, RT>, dim3(op.tc_grid), dim3(kTcThreads), (size_t)op.tc_smem, st,
                     op.tcp.cluster_n, op.maps, op.tcp, TcNoVecs{0});
}

cudaError_t launch_tc(const Op& op, cudaStream_t st) {
  const int occ = op.tc_occ;
  const bool ln = op.tcp.epi == EPI_LN_SHIFT || op.tcp.epi == EPI_LN_RES;
  if (ln && op.tcp.Nc == 64) {   // 64-column CTAs: row-in-registers LayerNorm epilogue
    if (op.tcp.Ntot == 64) {     // C_out == 64: epilogue vectors travel as kernel parameters
      if (op.tcp.epi == EPI_LN_SHIFT) return launch_tc_t<EPI_LN_SHIFT, 2, 2>(op, st);
      if (op.tcp.res_acc) return launch_tc_t<EPI_LN_RES, 2, 2, true>(op, st);
      return launch_tc_t<EPI_LN_RES, 2, 2>(op, st);
    }
    if (op.tcp.epi == EPI_LN_SHIFT) return launch_tc_t<EPI_LN_SHIFT, 2, 1>(op, st);
    if (op.tcp.res_acc) return launch_tc_t<EPI_LN_RES, 2, 1, true>(op, st);
    return launch_tc_t<EPI_LN_RES, 2, 1>(op, st);
  }
  if (op.tcp.epi == EPI_AFFINE && op.tcp.res_acc) return launch_tc_t<EPI_AFFINE, 2, 0, true>(op, st);   // C == 64
#define CASE(E_)                                                \
  if (op.tcp.epi == E_) {                                       \
    return occ == 2 ? launch_tc_t<E_, 2, 0>(op, st) : launch_tc_t<E_, 1, 0>(op, st); \
  }
  CASE(EPI_BIAS) CASE(EPI_LN_SHIFT) CASE(EPI_LN_RES) CASE(EPI_AFFINE) CASE(EPI_RAW)
#undef CASE
  return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------- plan builder

struct Builder {
  cdc_engine* e;
  Plan* pl;
  Arena arena;
  size_t arena_base = 0;
  int B, H, W;
  bool pending_time_join = false;

  template <typename T>
  T* ws(size_t off) { return reinterpret_cast<T*>(pl->ws + off); }  // valid arithmetic even for ws == nullptr (dry run)

  Act new_act(int C, int h, int w, bool with_lo = false) {
    Act a;
    a.C = C; a.H = h; a.W = w;
    a.bytes = (size_t)B * h * w * C * 2;
    a.off = arena_base + arena.alloc(a.bytes);
    if (with_lo) a.off_lo = arena_base + arena.alloc(a.bytes);
    return a;
  }
  void drop(const Act& a) {
    arena.release(a.off - arena_base, a.bytes);
    if (a.off_lo) arena.release(a.off_lo - arena_base, a.bytes);
  }
  template <typename T>
  T* lo_ptr(const Act& a) { return a.off_lo ? ws<T>(a.off_lo) : nullptr; }
  size_t raw_alloc(size_t bytes) { return arena_base + arena.alloc(bytes); }
  void raw_free(size_t off, size_t bytes) { arena.release(off - arena_base, bytes); }

  struct SegIn {
    Act a;
    int kh, kw, dy0, dx0;
    bool lo = false;        // read the compensation tensor of `a`
    bool merged = false;    // hi and lo weight passes share this segment's chunks (packed input, SegSpec mode 3)
  };
  // mirror of three_pass() on the activation side: (x_hi, W_hi)..., (x_lo, W_hi) for sources that have lo, (x_hi, W_lo)...
  static std::vector<SegIn> three_pass_in(const std::vector<SegIn>& base) {
    std::vector<SegIn> out = base;
    for (const SegIn& s : base)
      if (s.a.off_lo) {
        SegIn t = s;
        t.lo = true;
        out.push_back(t);
      }
    for (const SegIn& s : base)
      if (!s.merged) out.push_back(s);
    return out;
  }

  // generic conv op. stride / phases describe resampling; epi selects the fused epilogue.
  Op& conv(const std::string& name, const std::vector<SegIn>& segs, const ConvW& w, int epi, const Act& out,
           int stride, int phases) {
    pl->ops.emplace_back();
    Op& op = pl->ops.back();
    op.kind = OP_CONV;
    op.name = name;
    ConvParams& p = op.conv;
    p.nseg = (int)segs.size();
    int total = 0;
    for (int i = 0; i < p.nseg; ++i) {
      p.seg[i].src = ws<__half>(segs[i].lo ? segs[i].a.off_lo : segs[i].a.off);
      p.seg[i].C = segs[i].a.C;
      p.seg[i].kh = segs[i].kh;
      p.seg[i].kw = segs[i].kw;
      p.seg[i].dy0 = segs[i].dy0;
      p.seg[i].dx0 = segs[i].dx0;
      p.seg[i].nchunk = segs[i].kh * segs[i].kw * (segs[i].a.C / 64);
      p.seg[i].win8 = segs[i].a.win8 ? 1 : 0;
      total += p.seg[i].nchunk;
    }
    p.Hs = segs[0].a.H;
    p.Ws = segs[0].a.W;
    if (phases) { p.Ho = p.Hs; p.Wo = p.Ws; }
    else { p.Ho = out.H; p.Wo = out.W; }
    p.stride = stride;
    p.rows_per_group = B * p.Ho * p.Wo;
    p.groups = 1;
    p.w_group_stride = 0;
    p.W = dptr<__half>(e, w.w);
    p.Ntot = w.N;
    p.total_chunks = total;
    p.phases = phases;
    p.w_phase_stride = (long long)total * w.N * 64;
    p.out = ws<__half>(out.off);
    p.out_lo = lo_ptr<__half>(out);
    p.out_H = out.H;
    p.out_W = out.W;
    p.out_sy = phases ? 2 : 1;
    p.out_sx = phases ? 2 : 1;
    p.bias = w.bias ? dptr<float>(e, w.bias) : nullptr;
    op.epi = epi;
    const bool ln = epi == EPI_LN_SHIFT || epi == EPI_LN_RES;
    if (ln) { op.bn = w.N; op.bm = w.N <= 128 ? 128 : 64; }
    else { op.bn = (w.N % 128 == 0) ? 128 : 64; op.bm = 128; }
    op.grid = dim3((p.rows_per_group + op.bm - 1) / op.bm, w.N / op.bn, phases ? 4 : 1);
    op.dbg = p.out;
    op.dC = w.N; op.dH = out.H; op.dW = out.W;
    op.flops = 2.0 * w.macs_per_row * (double)B * out.H * out.W;
    return op;
  }

  // ResnetBlock: block1 (+temb shift) -> block2 (+residual). Returns the output activation.
  Act resnet(const std::string& name, const std::vector<SegIn>& segs1, const std::vector<SegIn>& segs_res,
             const ResW& w, int h, int wd, float2* stats_out) {
    // res_conv first, on the side lane: it only depends on the block input and overlaps block1
    Act r;
    bool own_r = false;
    // tcgen05 path, 64-column CTAs (C_out == 64, or a layer that runs as fused column slices): the 1x1 res_conv is
    // folded into block2 as extra K segments with their own TMEM accumulator (igemm_tc.cuh, res_acc) — no res_conv
    // launch, no residual round trip through HBM.  (Wider CTAs would lose their TMEM double buffering.)
    bool fuse_res = false;
    if (w.has_res && e->mainloop == 1 && e->fuse_res) {
      const int tiles_nominal = (h * wd * 8 + 127) / 128, nsl = w.cout / 64;
      const bool fused_slices = e->sliced && e->nslice && w.cout >= 128 && nsl <= 8 && tiles_nominal < e->slice_max_tiles &&
                                tiles_nominal * nsl >= 64;
      fuse_res = (w.cout == 64 || fused_slices) && 1 + three_pass_in(segs_res).size() <= (size_t)kMaxSeg;
    }
    if (w.has_res && !fuse_res) {
      r = new_act(w.cout, h, wd, true);
      own_r = true;
      Op& rop = conv(name + "res_conv", three_pass_in(segs_res), w.res, EPI_BIAS, r, 1, 0);
      rop.lane = e->two_lanes ? 1 : 0;
    }
    Act h1 = new_act(w.cout, h, wd);
    {
      Op& op = conv(name + "block1", segs1, w.b1.conv, EPI_LN_SHIFT, h1, 1, 0);
      op.join_before = pending_time_join;   // the first block1 of the plan waits for the timestep MLP on the side lane
      pending_time_join = false;
      op.conv.ln_g = dptr<float>(e, w.b1.g);
      op.conv.ln_b = dptr<float>(e, w.b1.b);
      op.conv.shift = w.shift_off >= 0 ? ws<float>(pl->shifts_off) + w.shift_off : nullptr;
      op.conv.shift_stride = e->R;
    }
    Act out = new_act(w.cout, h, wd, true);
    {
      std::vector<SegIn> s2 = {{h1, 3, 3, -1, -1}};
      // identity residual of a 64-channel block: x_hi * I + x_lo * I into the second accumulator — the residual reaches
      // the epilogue through TMA + two trivial MMAs, in fp32, instead of strided global loads and transpositions
      const bool ident_res = !w.has_res && e->mainloop == 1 && e->fuse_res && w.cout == 64 && segs_res.size() == 1 &&
                             segs_res[0].a.C == 64;
      if (fuse_res)
        for (const SegIn& t : three_pass_in(segs_res)) s2.push_back(t);
      if (ident_res) {
        s2.push_back(segs_res[0]);
        if (segs_res[0].a.off_lo) {
          SegIn t = segs_res[0];
          t.lo = true;
          s2.push_back(t);
        }
      }
      Op& op = conv(name + "block2", s2, w.b2.conv, EPI_LN_RES, out, 1, 0);
      op.join_before = w.has_res && !fuse_res;
      op.conv.ln_g = dptr<float>(e, w.b2.g);
      op.conv.ln_b = dptr<float>(e, w.b2.b);
      if (fuse_res) {
        // segments 1.. are the res_conv's three passes: own weight chunks (w.res, in three_pass order), accumulator 1
        const __half* wr = dptr<__half>(e, w.res.w);
        size_t q = 0;
        for (int i = 1; i < op.conv.nseg; ++i) {
          op.conv.seg[i].acc = 1;
          op.conv.seg[i].W = wr + q * (size_t)w.cout * 64;
          q += (size_t)op.conv.seg[i].nchunk;
        }
        op.conv.res = nullptr;
        op.conv.res_acc = 1;
        op.conv.res_bias = w.res.bias ? dptr<float>(e, w.res.bias) : nullptr;
        op.flops += 2.0 * w.res.macs_per_row * (double)B * h * wd;
      } else if (ident_res) {
        for (int i = 1; i < op.conv.nseg; ++i) {
          op.conv.seg[i].acc = 1;
          op.conv.seg[i].W = dptr<__half>(e, e->w_ident);
        }
        op.conv.res = nullptr;
        op.conv.res_acc = 1;
        op.conv.res_bias = nullptr;
      } else if (w.has_res) {
        op.conv.res = ws<__half>(r.off);
        op.conv.res_lo = lo_ptr<__half>(r);
        op.conv.res_C0 = w.cout;
        op.conv.res2 = nullptr;
        op.conv.res2_lo = nullptr;
      } else {
        // identity residual: the (possibly concatenated) block input itself
        op.conv.res = ws<__half>(segs_res[0].a.off);
        op.conv.res_lo = lo_ptr<__half>(segs_res[0].a);
        op.conv.res_C0 = segs_res[0].a.C;
        op.conv.res2 = segs_res.size() > 1 ? ws<__half>(segs_res[1].a.off) : nullptr;
        op.conv.res2_lo = segs_res.size() > 1 ? lo_ptr<__half>(segs_res[1].a) : nullptr;
      }
      op.conv.stats_out = stats_out;
    }
    drop(h1);
    if (own_r) drop(r);
    return out;
  }

  Act attention(const std::string& name, const Act& x, float2* stats, size_t stats_off, size_t stats_bytes,
                const AttnW& w) {
    const int C = w.C, N = x.H * x.W, cb = C / 64;
    const int ntiles = (N + 63) / 64;
    // The split over pixels depends on (N, C) only — never on B — so every image goes through the same
    // sequence of fp32 additions whatever batch it is decoded in (batch-invariant, bit-reproducible shards).
    // tcgen05 context kernel (attn_tc.cuh): C in {64, 128, 192}, whole 64-pixel tiles; ~18 pixel chunks per image and
    // (K block, V block) pair keep 144 CTAs busy at the nominal batch of 8 and the partial buffers small.
    const bool ctx_tc = e->mainloop == 1 && e->attn_tc && (C == 64 || C == 128 || C == 192 || C == 256 || C == 320) && N % 64 == 0;
    const int pairs = C == 64 ? 1 : ((C + 127) / 128) * ((C + 127) / 128);
    // The chunk count grows with the image area (a function of H, W only — never of B): a single 512 x 768 image has as
    // many pixels as six 256 x 256 ones and must fill the GPU on its own (the demo scripts decode one image at a time).
    // Measured (round 2): proportional growth takes the single 512 x 768 image from 3.00 to 2.45 ms/step and costs a batch
    // of eight 512 x 512 images 3 % (their context kernels go from one wave of CTAs to several); sqrt growth costs the
    // batch the same 3 % and gives the single image only 2.55 ms.  Proportional it is.
    const int area = e->attn_area ? std::max(1, (pl->H * pl->W + 32768) / 65536) : 1;
    const int want_chunks = std::min(kCombineMaxChunks, (C == 64 ? 36 : std::max(1, 18 / pairs)) * area);   // C == 64 runs two CTAs per SM
    const int tpc = ctx_tc ? (ntiles + want_chunks - 1) / want_chunks
                           : std::max(1, std::max(8, (ntiles + 63) / 64) / area);
    const int nchunks = (ntiles + tpc - 1) / tpc;
    const size_t pc_b = (size_t)B * nchunks * C * C * 4, pv_b = (size_t)B * nchunks * C * 4;
    const size_t pc = raw_alloc(pc_b), pm = raw_alloc(pv_b), ps = raw_alloc(pv_b);
    const size_t cc_b = (size_t)B * C * C * 4;
    const size_t ctxn = raw_alloc(cc_b);
    const bool direct = nchunks == 1;   // one pixel chunk: the context kernel writes ctx / S itself
    // tcgen05 form of the two C x C products: ctx / S also leaves the producer as fp16 value + remainder (operand layout)
    // (C >= 256 only: at C <= 192 the products are a few microseconds of launch latency either way and the mma.sync kernel's
    // 64 x 64 tiles give more CTAs — measured 1-2 us faster per product)
    const bool alg = e->mainloop == 1 && e->alg_tc && C >= 256;
    const size_t c16_b = (size_t)B * C * C * 2;
    const size_t c16h = alg ? raw_alloc(c16_b) : 0, c16l = alg ? raw_alloc(c16_b) : 0;
    {
      pl->ops.emplace_back();
      Op& op = pl->ops.back();
      op.kind = OP_ATTN_CTX;
      op.name = name + "ctx";
      op.actx.x = ws<__half>(x.off);
      op.actx.stats = stats;
      op.actx.Wkv = dptr<__half>(e, w.wkv);
      op.actx.u = dptr<float>(e, w.u);
      op.actx.c = dptr<float>(e, w.c);
      op.actx.C = C;
      op.actx.N = N;
      op.actx.tiles_per_chunk = tpc;
      op.actx.nchunks = nchunks;
      op.actx.part_ctx = ws<float>(pc);
      op.actx.part_m = ws<float>(pm);
      op.actx.part_s = ws<float>(ps);
      op.actx.ctxn = direct ? ws<float>(ctxn) : nullptr;
      op.actx.ctx16_hi = (direct && alg) ? ws<__half>(c16h) : nullptr;
      op.actx.ctx16_lo = (direct && alg) ? ws<__half>(c16l) : nullptr;
      op.grid = dim3(nchunks, cb * cb, B);
      if (ctx_tc) {
        op.use_tc = true;
        AttnTcParams& q = op.atc;
        q.C = C; q.N = N;
        q.stacked = C == 64;
        q.cpt = cb;
        q.kbc = (C + 127) / 128;
        q.ntiles = ntiles;
        q.tiles_per_chunk = tpc;
        q.nchunks = nchunks;
        q.stats = stats;
        q.u = op.actx.u; q.c = op.actx.c;
        q.part_ctx = op.actx.part_ctx; q.part_m = op.actx.part_m; q.part_s = op.actx.part_s;
        q.ctxn = op.actx.ctxn;
        q.ctx16_hi = op.actx.ctx16_hi;
        q.ctx16_lo = op.actx.ctx16_lo;
        op.grid = dim3(nchunks, pairs, B);
      }
      // to_qkv 1x1 (3C x C per pixel) + the two einsums (2 * C*C per pixel) + to_out (C x C per pixel)
      op.flops = 2.0 * (double)B * N * C * (3.0 * C + 2.0 * C + C);
    }
    if (!direct) {
      pl->ops.emplace_back();
      Op& op = pl->ops.back();
      op.kind = OP_COMBINE;
      op.name = name + "combine";
      op.comb = {ws<float>(pc), ws<float>(pm), ws<float>(ps), C, nchunks, ws<float>(ctxn),
                 alg ? ws<__half>(c16h) : nullptr, alg ? ws<__half>(c16l) : nullptr};
      op.grid = dim3(C, B, 1);
    }
    raw_free(pc, pc_b); raw_free(pm, pv_b); raw_free(ps, pv_b);
    if (alg) return attention_tail_tc(name, x, stats, stats_off, stats_bytes, w, ctxn, cc_b, c16h, c16l, c16_b);
    const size_t T = raw_alloc(cc_b);
    {
      pl->ops.emplace_back();
      Op& op = pl->ops.back();
      op.kind = OP_SGEMM;
      op.name = name + "T";
      op.sg = {ws<float>(ctxn), dptr<float>(e, w.wq), ws<float>(T), C, C, C, (long long)C * C, 0, (long long)C * C};
      op.bm = 64;
      op.grid = dim3(C / 64, C / op.bm, B);
    }
    raw_free(ctxn, cc_b);
    // tcgen05 path: the second product writes M_b in fp16 weight layout and per-column-tile partial row sums itself
    // (fused finish); the mma.sync convolution keeps the separate finish kernel and whole-row sums.
    const bool fold = e->mainloop == 1 && e->fold_finish;
    const int parts = fold ? cb : 1;
    const size_t Mf = raw_alloc(cc_b);
    const size_t mg_b = (size_t)B * C * C * 2, v_b = (size_t)B * parts * C * 4;
    const size_t Mg = raw_alloc(mg_b), um = raw_alloc(v_b), cm = raw_alloc(v_b);
    {
      pl->ops.emplace_back();
      Op& op = pl->ops.back();
      op.kind = OP_SGEMM;
      op.name = name + "M";
      op.sg = {dptr<float>(e, w.woT), ws<float>(T), ws<float>(Mf), C, C, C, 0, (long long)C * C, (long long)C * C};
      op.bm = 64;
      op.grid = dim3(C / 64, C / op.bm, B);
      if (fold)
        op.gfin = {dptr<float>(e, w.g), dptr<float>(e, w.bln), dptr<float>(e, w.bout), ws<__half>(Mg), ws<float>(um),
                   ws<float>(cm)};
    }
    raw_free(T, cc_b);
    if (!fold) {
      pl->ops.emplace_back();
      Op& op = pl->ops.back();
      op.kind = OP_FINISH;
      op.name = name + "finish";
      op.fin = {ws<float>(Mf), dptr<float>(e, w.g), dptr<float>(e, w.bln), dptr<float>(e, w.bout), C,
                ws<__half>(Mg), ws<float>(um), ws<float>(cm)};
      op.grid = dim3(C, B, 1);
    }
    raw_free(Mf, cc_b);
    Act out = attention_out(name, x, stats, w, Mg, um, cm, parts);
    raw_free(Mg, mg_b); raw_free(um, v_b); raw_free(cm, v_b);
    raw_free(stats_off, stats_bytes);
    return out;
  }

  // Output GEMM of the attention block: out = M_b-folded per-image weights applied to raw x (EPI_AFFINE) + residual.
  Act attention_out(const std::string& name, const Act& x, float2* stats, const AttnW& w, size_t Mg, size_t um, size_t cm,
                    int parts) {
    const int C = w.C, N = x.H * x.W, cb = C / 64;
    Act out = new_act(C, x.H, x.W, true);
    {
      ConvW cw;
      cw.w = 0; cw.bias = 0; cw.N = C; cw.nchunks = cb; cw.macs_per_row = 0;
      std::vector<SegIn> s = {{x, 1, 1, 0, 0}};
      // C == 64: the residual x (hi + lo) reaches the epilogue through the MMA (identity weights, second accumulator)
      const bool ident_res = C == 64 && e->mainloop == 1 && e->fuse_res;
      if (ident_res) {
        s.push_back({x, 1, 1, 0, 0});
        if (x.off_lo) {
          SegIn t2 = {x, 1, 1, 0, 0};
          t2.lo = true;
          s.push_back(t2);
        }
      }
      Op& op = conv(name + "out", s, cw, EPI_AFFINE, out, 1, 0);
      ConvParams& p = op.conv;
      p.W = ws<__half>(Mg);
      p.groups = B;
      p.rows_per_group = N;
      p.w_group_stride = (long long)C * C;
      p.stats_in = stats;
      p.aff_u = ws<float>(um);
      p.aff_c = ws<float>(cm);
      p.aff_group_stride = C;
      p.aff_parts = parts;
      p.res = ws<__half>(x.off);
      p.res_lo = lo_ptr<__half>(x);
      p.res_C0 = C;
      p.res2 = nullptr;
      p.res2_lo = nullptr;
      p.bias = nullptr;
      if (ident_res) {
        for (int i = 1; i < p.nseg; ++i) {
          p.seg[i].acc = 1;
          p.seg[i].W = dptr<__half>(e, e->w_ident);
          p.seg[i].wshared = 1;
        }
        p.res = nullptr;
        p.res_lo = nullptr;
        p.res_acc = 1;
        p.res_bias = nullptr;
      }
      op.grid = dim3(B * ((N + op.bm - 1) / op.bm), C / op.bn, 1);
      op.flops = 0;
    }
    return out;
  }

  // tcgen05 form of the per-image algebra: T = ctx^T (C^-1/2 W_q), M_b = W_out T with the finish epilogue — two launches of
  // attn_alg_tc_kernel on MN-blocked fp16 value / remainder operands (attn_alg_tc.cuh).
  Act attention_tail_tc(const std::string& name, const Act& x, float2* stats, size_t stats_off, size_t stats_bytes,
                        const AttnW& w, size_t ctxn, size_t cc_b, size_t c16h, size_t c16l, size_t c16_b) {
    const int C = w.C, cb = C / 64;
    raw_free(ctxn, cc_b);
    const size_t t16h = raw_alloc(c16_b), t16l = raw_alloc(c16_b);
    auto alg_op = [&](const char* suffix, int mode) -> Op& {
      pl->ops.emplace_back();
      Op& op = pl->ops.back();
      op.kind = OP_ALG;
      op.name = name + suffix;
      AlgTcParams& q = op.alg;
      q.C = C; q.kchunks = cb; q.n_tiles = (C + 127) / 128; q.mode = mode;
      q.stages = std::min(3, cb);
      op.grid = dim3(q.n_tiles * q.n_tiles, 1, B);
      return op;
    };
    {
      Op& op = alg_op("T", 0);
      op.alg.a_img = 1; op.alg.b_img = 0;
      op.algsrc = {ws<__half>(c16h), ws<__half>(c16l), dptr<__half>(e, w.wq16h), dptr<__half>(e, w.wq16l)};
      op.alg.out_hi = ws<__half>(t16h);
      op.alg.out_lo = ws<__half>(t16l);
    }
    raw_free(c16h, c16_b); raw_free(c16l, c16_b);
    const int parts = cb;
    const size_t mg_b = (size_t)B * C * C * 2, v_b = (size_t)B * parts * C * 4;
    const size_t Mg = raw_alloc(mg_b), um = raw_alloc(v_b), cm = raw_alloc(v_b);
    {
      Op& op = alg_op("M", 1);
      op.alg.a_img = 0; op.alg.b_img = 1;
      op.algsrc = {dptr<__half>(e, w.wo16h), dptr<__half>(e, w.wo16l), ws<__half>(t16h), ws<__half>(t16l)};
      op.alg.g = dptr<float>(e, w.g); op.alg.bln = dptr<float>(e, w.bln); op.alg.bout = dptr<float>(e, w.bout);
      op.alg.Mg16 = ws<__half>(Mg); op.alg.um_part = ws<float>(um); op.alg.cm_part = ws<float>(cm);
    }
    raw_free(t16h, c16_b); raw_free(t16l, c16_b);
    Act out = attention_out(name, x, stats, w, Mg, um, cm, parts);
    raw_free(Mg, mg_b); raw_free(um, v_b); raw_free(cm, v_b);
    raw_free(stats_off, stats_bytes);
    return out;
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}


int pow2floor(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

// Route a stride-1 convolution op to the tcgen05/TMA kernel: tile geometry, pipeline depth, tensor maps.
int setup_tc(cdc_engine* e, Plan* pl, Op& op, bool allow_dual = true) {
  const ConvParams& c = op.conv;
  op.use_tc = false;
  if (e->mainloop != 1) return 0;
  // tile space = output pixels (for the transposed-conv phases: the input-resolution pixels of one phase)
  const int B = pl->B, h = c.Ho, w = c.Wo, N = c.Ntot;
  TcConvParams t{};
  // Tile = TB x TH x TW pixels, powers of two that never exceed the tensor extents; small images / batches give
  // tiles with fewer than 128 rows (the unused accumulator rows are masked).  Every choice below is independent
  // of what other images are in the batch, so results are bit-identical for any batch composition.
  t.TW = std::min(16, pow2floor(w));
  t.TH = std::min(128 / t.TW, pow2floor(h));
  t.TB = c.groups > 1 ? 1 : std::min(128 / (t.TW * t.TH), pow2floor(B));
  const int a_rows = t.TW * t.TH * t.TB;
  t.tiles_x = (w + t.TW - 1) / t.TW;
  t.tiles_y = (h + t.TH - 1) / t.TH;
  t.tiles_b = (B + t.TB - 1) / t.TB;
  t.B = B; t.H = h; t.W = w;
  t.stride = c.stride;
  // Segments of the launch.  The (x_hi, W_hi) and (x_hi, W_lo) passes of a 3-pass trunk convolution read the same
  // activations: outside vertical-reuse mode they are merged into one "dual" segment — one activation load per tap,
  // two weight tiles, two MMAs (Downsample: 27 -> 18 stage loads per tile, 54 -> 36 TMA operations).
  int khmax_all = 1;
  for (int i = 0; i < c.nseg; ++i) khmax_all = std::max(khmax_all, c.seg[i].kh);
  const bool dual_ok = allow_dual && e->dual_pass && c.groups == 1 && !c.phases && !(c.stride == 1 && khmax_all > 1);
  int cq0[kMaxSeg], cidx[kMaxSeg], dual_with[kMaxSeg];
  bool dropped[kMaxSeg];
  for (int i = 0, q0 = 0; i < c.nseg; ++i) {
    cq0[i] = q0;
    q0 += c.seg[i].nchunk;
    dropped[i] = false;
    dual_with[i] = -1;
  }
  // 1 x 1 segment pairs (the W_hi / W_lo passes of a fused res_conv) may share their activation load inside a vertical-reuse
  // launch too: a stage then carries two weight tiles and no tap offset, next to the convolution's kh-tile stages
  const bool dual_1x1 = allow_dual && e->dual_pass && c.groups == 1 && !c.phases && c.stride == 1;
  if (dual_ok || dual_1x1)
    for (int i = 0; i < c.nseg; ++i) {
      if (dropped[i] || dual_with[i] >= 0) continue;
      for (int j = i + 1; j < c.nseg; ++j) {
        const ConvSeg &a = c.seg[i], &b2 = c.seg[j];
        if (dropped[j] || a.src != b2.src || a.C != b2.C || a.kh != b2.kh || a.kw != b2.kw || a.dy0 != b2.dy0 ||
            a.dx0 != b2.dx0 || a.acc != b2.acc)
          continue;
        if (!dual_ok && !(a.kh == 1 && a.kw == 1)) continue;
        // segments with their own weight pointer: the W_lo set must sit where the chunk numbering says (same blob, same order)
        if ((a.W == nullptr) != (b2.W == nullptr)) continue;
        if (a.W && (b2.W - a.W) != (long long)(cq0[j] - cq0[i]) * N * 64) continue;
        dual_with[i] = j;
        dropped[j] = true;
        break;
      }
    }
  t.nseg = 0;
  for (int i = 0; i < c.nseg; ++i) {
    if (dropped[i]) continue;
    const int k = t.nseg++;
    cidx[k] = i;
    t.seg[k].cpt = c.seg[i].C / 64;
    t.seg[k].kh = c.seg[i].kh; t.seg[k].kw = c.seg[i].kw;
    t.seg[k].dy0 = c.seg[i].dy0; t.seg[k].dx0 = c.seg[i].dx0;
    t.seg[k].nchunk = c.seg[i].nchunk;
    t.seg[k].vr = 1;
    t.seg[k].a_bytes = 128 * a_rows;
    t.seg[k].q0 = cq0[i];
    t.seg[k].acc = c.seg[i].acc;
    t.seg[k].wshared = c.seg[i].wshared;
    t.seg[k].dual = dual_with[i] >= 0 ? 1 : 0;
    t.seg[k].nw = t.seg[k].dual ? 2 : 1;
    t.seg[k].avstep = 0;
  }
  int dual_set_chunks[kMaxSeg];   // chunks between the W_hi and W_lo tiles of a dual segment
  for (int k = 0; k < t.nseg; ++k) dual_set_chunks[k] = t.seg[k].dual ? cq0[dual_with[cidx[k]]] - cq0[cidx[k]] : 0;
  bool any_dual = false;
  for (int k = 0; k < t.nseg; ++k) any_dual |= t.seg[k].dual != 0;
  t.total_chunks = c.total_chunks;
  t.Ntot = N;
  // Sliced mode for layers that cannot fill the GPU with 128-pixel tiles: 64-column slices x K splits, fp32
  // partial tiles finished by ln_rows_kernel (see TcConvParams).
  const int tiles_total = t.tiles_x * t.tiles_y * t.tiles_b * (c.phases ? 4 : 1);
  // The slicing heuristic looks at a NOMINAL batch of 8 images, not the actual one: the K-split (hence the fp32
  // summation order) must not depend on the batch size.
  const int tiles_nominal = ((h * w * 8 + 127) / 128) * (c.phases ? 4 : 1);
  t.Nc = N; t.n_slices = 1; t.k_splits = 1;
  // attention output GEMM (per-image weights, EPI_AFFINE): column slices only (no LayerNorm, no K split needed)
  const bool affine_slices = e->sliced && e->nslice && op.epi == EPI_AFFINE && tiles_nominal < e->slice_max_tiles && N >= 128;
  const bool sliceable = affine_slices ||
                         (e->sliced && c.groups == 1 && op.epi != EPI_AFFINE && tiles_nominal < e->slice_max_tiles && N >= 128);
  // Fused column slices (default): no K split, the fused epilogue runs in the kernel; LayerNorm epilogues exchange
  // their row statistics inside a thread-block cluster of the n_slices CTAs of a tile (<= 8: portable cluster size).
  const bool ln_epi = op.epi == EPI_LN_SHIFT || op.epi == EPI_LN_RES;
  // Taken when the tile x slice grid alone gives enough CTAs (measured: >= 64 units, 128 for the 4-phase transposed
  // convolutions); the lowest levels keep the K-split + ln_rows_kernel form, and so do the stride-2 convolutions.
  const int units_nominal = tiles_nominal * (N / 64);
  const bool fused = affine_slices ||
                     (sliceable && e->nslice && N / 64 <= 8 && c.stride == 1 && units_nominal >= (c.phases ? 128 : 64));
  t.cluster_n = 0;
  t.xchg_stats = 0;
  if (fused) {
    t.Nc = 64;
    t.n_slices = N / 64;
    t.k_splits = 1;
    if (ln_epi) {
      t.cluster_n = t.n_slices;
      t.xchg_stats = c.stats_out != nullptr;
    }
  } else if (sliceable) {
    t.Nc = 64;
    t.n_slices = N / 64;
    const int slots = e->slice_slots;
    t.k_splits = std::max(1, std::min(std::min(c.total_chunks / 2, e->slice_kmax),
                                      (slots + tiles_nominal * t.n_slices - 1) / (tiles_nominal * t.n_slices)));
  }
  const int Nc = t.Nc;
  t.n_split = Nc > 256 ? 2 : 1;
  t.n_piece = Nc / t.n_split;
  if (c.res_acc && Nc != 64)
    return fail(e, CDC_ERR_STATE, "op %s: fused res_conv needs 64-column CTAs (builder / setup_tc slicing rules diverged)", op.name.c_str());
  t.res_acc = c.res_acc;
  t.res_bias = c.res_bias;
  t.nbuf = (2 * Nc * (c.res_acc ? 2 : 1) <= 512) ? 2 : 1;
  // Nc <= 128: two CTAs per SM (2 x 256 TMEM columns, half the shared memory each) double the epilogue warps
  int ctas_per_sm = Nc <= 128 ? 2 : 1;
  // few CTAs (lowest levels): one CTA per SM with the deepest pipeline — the K loop is TMA-latency bound there
  if (fused && tiles_total * t.n_slices <= e->num_sms) ctas_per_sm = 1;
  const int xsets = t.cluster_n > 1 ? (t.xchg_stats ? 2 : 1) : 0;
  auto smem_budget = [&](int ctas) {
    return (227 * 1024) / ctas - 1024 - tc_tail_bytes(Nc, t.cluster_n, xsets) - (ctas > 1 ? 1024 : 0);
  };
  int budget = smem_budget(ctas_per_sm);
  // two weight tiles per stage must still leave a two-stage pipeline (wide un-sliced layers): else plain 3-pass form
  if (any_dual && budget / tc_stage_bytes(16384, 2, Nc) < 2) return setup_tc(e, pl, op, false);
  // Vertical reuse: one activation box of TH+kh-1 tile rows serves all kh vertical taps of a (kx, channel chunk) —
  // tap ky reads it TW pixel rows further down, which is a whole number of 1024-byte swizzle atoms when TW is
  // 8 or 16.  Cuts the L2->SM activation traffic of a 3x3 conv from 9 to 3.75 tile loads (the top levels are
  // L2-bandwidth bound).  Needs one image per tile and >= 2 pipeline stages of (box + kh weight tiles).
  t.b_off = 16384;
  t.vr_max = any_dual ? 2 : 1;
  t.total_sc = 0;
  for (int k = 0; k < t.nseg; ++k) t.total_sc += t.seg[k].nchunk;
  const bool vr_ok = e->vreuse && c.stride == 1 && t.TB == 1 && (t.TW == 8 || t.TW == 16) && a_rows == 128;
  if (vr_ok) {
    int khmax = 1;
    for (int i = 0; i < t.nseg; ++i) khmax = std::max(khmax, t.seg[i].kh);
    const int b_off = ((128 * t.TW * (t.TH + khmax - 1)) + 1023) & ~1023;
    if (khmax > 1 && budget / tc_stage_bytes(b_off, khmax, Nc) < 2 && e->vreuse >= 2 && ctas_per_sm == 2 &&
        smem_budget(1) / tc_stage_bytes(b_off, khmax, Nc) >= 2) {
      ctas_per_sm = 1;   // trade the second CTA for the reuse pipeline (CDC_VREUSE=2)
      budget = smem_budget(1);
    }
    if (khmax > 1 && budget / tc_stage_bytes(b_off, khmax, Nc) >= 2) {
      t.b_off = b_off;
      t.vr_max = khmax;
      t.total_sc = 0;
      for (int i = 0; i < t.nseg; ++i) {
        if (t.seg[i].dual) {   // 1 x 1 dual segment: one box of the tile's own rows, two weight tiles, no tap offset
          t.seg[i].vr = 1;
          t.seg[i].nw = 2;
          t.seg[i].avstep = 0;
          t.seg[i].a_bytes = 128 * t.TW * t.TH;
          t.total_sc += t.seg[i].kw * t.seg[i].cpt;
          continue;
        }
        t.seg[i].vr = t.seg[i].kh;
        t.seg[i].nw = t.seg[i].kh;
        t.seg[i].avstep = (t.TW * 128) >> 4;
        t.seg[i].a_bytes = 128 * t.TW * (t.TH + t.seg[i].kh - 1);
        t.total_sc += t.seg[i].kw * t.seg[i].cpt;
      }
    }
  }
  if (t.k_splits > std::max(1, t.total_sc / 2)) t.k_splits = std::max(1, t.total_sc / 2);
  const int stage_bytes = tc_stage_bytes(t.b_off, t.vr_max, Nc);
  t.stages = std::max(2, std::min(kTcMaxStages, budget / stage_bytes));
  t.phases = c.phases ? 4 : 1;
  t.w_rows_per_phase = c.total_chunks * N;
  t.w_rows_per_image = c.groups > 1 ? c.total_chunks * N : 0;
  t.epi = (sliceable && !fused) ? EPI_RAW : op.epi;
  t.out = c.out; t.out_lo = c.out_lo; t.out_H = c.out_H; t.out_W = c.out_W; t.out_sy = c.out_sy; t.out_sx = c.out_sx;
  t.bias = c.bias; t.ln_g = c.ln_g; t.ln_b = c.ln_b; t.shift = c.shift; t.shift_stride = c.shift_stride;
  t.res = c.res; t.res_C0 = c.res_C0; t.res2 = c.res2; t.res_lo = c.res_lo; t.res2_lo = c.res2_lo;
  t.stats_in = c.stats_in; t.aff_u = c.aff_u; t.aff_c = c.aff_c; t.stats_out = c.stats_out;
  t.aff_parts = c.aff_parts;
  t.ln_out = c.ln_out;
  t.skip_out = c.skip_out;
  op.tc_smem = tc_smem_bytes(stage_bytes, t.stages, Nc, t.cluster_n, xsets);
  op.tc_occ = ctas_per_sm;
  op.tc_grid = std::min(tiles_total * t.n_slices * t.k_splits, e->num_sms * ctas_per_sm);
  if (fused)   // whole tiles per pass: the grid is a multiple of n_slices, so blockIdx % n_slices is the CTA's column slice
    op.tc_grid = std::min(tiles_total, std::max(1, (e->num_sms * ctas_per_sm) / t.n_slices)) * t.n_slices;
  if (sliceable && !fused) {
    // Fused finish (per-tile arrival counters, igemm_tc.cuh): main-lane launches only.  A CTA waits for its tile's other
    // units, so every CTA of the launch must become resident: the grid never exceeds the SM slots, and the side lane's
    // concurrent launch keeps the separate ln_rows_kernel (two spinning launches could starve each other of SM slots).
    if (e->fuse_lnrows && op.lane == 0) {
      t.fin_epi = op.epi;
      op.ctr_index = pl->ctr_count;
      pl->ctr_count += 2ll * tiles_total;
    }
    const long long out_pix = (long long)B * c.out_H * c.out_W;
    t.raw_split_stride = out_pix * N;
    const size_t need = (size_t)t.k_splits * (size_t)out_pix * (size_t)N * 4;
    if (op.lane == 1) pl->raw2_bytes = std::max(pl->raw2_bytes, need);   // raw pointer patched once raw_bytes is final
    else pl->raw_bytes = std::max(pl->raw_bytes, need);
    t.raw = nullptr;
  }
  t.dbg_skip_epi = (getenv("CDC_DBG_EPI") && atoi(getenv("CDC_DBG_EPI")) && t.cluster_n <= 1 && t.epi != EPI_RAW) ? 1 : 0;
  t.fd_upt = make_fastdiv((uint32_t)(t.n_slices * t.k_splits));
  t.fd_ks = make_fastdiv((uint32_t)t.k_splits);
  t.fd_tpp = make_fastdiv((uint32_t)(t.tiles_x * t.tiles_y * t.tiles_b));
  t.fd_txy = make_fastdiv((uint32_t)(t.tiles_x * t.tiles_y));
  t.fd_tx = make_fastdiv((uint32_t)t.tiles_x);
  op.tcp = t;
  op.use_tc = true;
  if (ln_epi && N == 64) {   // host copies of the epilogue vectors (the blob's host image mirrors the device blob)
    auto host_of = [&](const float* dev) {
      return reinterpret_cast<const float*>(e->blob.host.data() + (reinterpret_cast<const uint8_t*>(dev) - e->dblob));
    };
    for (int i = 0; i < 64; ++i) {
      op.vecs64.bias[i] = c.bias ? host_of(c.bias)[i] : 0.f;
      op.vecs64.g[i] = c.ln_g ? host_of(c.ln_g)[i] : 1.f;
      op.vecs64.b[i] = c.ln_b ? host_of(c.ln_b)[i] : 0.f;
      op.vecs64.rbias[i] = c.res_bias ? host_of(c.res_bias)[i] : 0.f;
    }
  }
  if (!pl->ws) return 0;  // dry run: geometry only
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  for (int i = 0; i < t.nseg; ++i) {
    const ConvSeg& cs = c.seg[cidx[i]];
    const cuuint64_t Cs = (cuuint64_t)cs.C;
    const cuuint64_t ws_ = (cuuint64_t)c.Ws, hs_ = (cuuint64_t)c.Hs;
    cuuint64_t gdim[4] = {Cs, ws_, hs_, (cuuint64_t)B};
    cuuint64_t gstr[3] = {Cs * 2, ws_ * Cs * 2, hs_ * ws_ * Cs * 2};
    if (cs.win8) {   // overlapping windows: pixel stride 16 B under a 128-B inner extent (tests/microbench/tma_window.cu)
      gstr[0] = 16;
      gstr[1] = (ws_ + 8) * 16;
      gstr[2] = hs_ * (ws_ + 8) * 16;
    }
    // strided convolution: the box spans stride*T source pixels, of which every stride-th is loaded (a dense 5-D
    // parity view of the source was measured: no faster)
    const cuuint32_t sx = (cuuint32_t)c.stride;
    cuuint32_t box[4] = {64, (cuuint32_t)t.TW * sx, (cuuint32_t)(t.TH + t.seg[i].vr - 1) * sx, (cuuint32_t)t.TB};
    cuuint32_t estr[4] = {1, sx, sx, 1};
    CUresult r = enc(&op.maps.a[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)cs.src, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled(A, op %s, seg %d) failed: %d", op.name.c_str(), i, (int)r);
  }
  for (int i = t.nseg; i < kMaxSeg; ++i) op.maps.a[i] = op.maps.a[0];
  // Weights [phase | image][chunk q][C_out][64]: per segment a 4-D view {64, kw*cpt*C_out, kh, phase | image} so that
  // one box {64, n_piece, vr, 1} brings the weight tiles of all vr vertical taps of a (kx, channel chunk).
  for (int i = 0; i < t.nseg; ++i) {
    const ConvSeg& cs = c.seg[cidx[i]];
    const cuuint64_t tap_rows = (cuuint64_t)cs.kw * t.seg[i].cpt * N;
    const bool dual = t.seg[i].dual != 0;
    cuuint64_t gdim[4] = {64, tap_rows, (cuuint64_t)cs.kh,
                          (cuuint64_t)(dual ? 2 : (cs.wshared ? 1 : (c.groups > 1 ? B : t.phases)))};
    // dim 3 = output phase (transposed conv: weight sets follow each other), image (per-image attention matrices), or
    // the (W_hi, W_lo) pair of a dual segment
    const cuuint64_t set_stride = dual ? (cuuint64_t)dual_set_chunks[i] * N * 128
                                       : (c.groups > 1 ? (cuuint64_t)c.w_group_stride * 2 : (cuuint64_t)c.total_chunks * N * 128);
    cuuint64_t gstr[3] = {128, tap_rows * 128, set_stride};
    cuuint32_t box[4] = {64, (cuuint32_t)t.n_piece, (cuuint32_t)t.seg[i].vr, (cuuint32_t)(dual ? 2 : 1)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    void* wbase = cs.W ? (void*)cs.W : (void*)((const __half*)c.W + (size_t)t.seg[i].q0 * N * 64);
    CUresult r = enc(&op.maps.b[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, wbase, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled(B, op %s, seg %d) failed: %d", op.name.c_str(), i, (int)r);
  }
  for (int i = t.nseg; i < kMaxSeg; ++i) op.maps.b[i] = op.maps.b[0];
  return 0;
}

// Tensor maps and pipeline depth of the tcgen05 attention-context kernel.
int setup_attn_tc(cdc_engine* e, Plan* pl, Op& op) {
  AttnTcParams& q = op.atc;
  const int nblk = q.stacked ? 1 : 2;
  q.pvbufs = q.C > 256 ? 1 : 2;   // C = 320: 160 KB of resident weights
  const int fixed = 1024 + nblk * q.cpt * 16384 + q.pvbufs * nblk * 16384 + kAttnStatSlots * 512 + 512;
  q.stages = std::max(2, std::min(8, (224 * 1024 - fixed) / 8192));
  if (q.stacked) q.stages = std::min(q.stages, 4);   // leaves room for two CTAs per SM
  op.atc_smem = attn_tc_smem_bytes(q.stacked, q.cpt, q.stages, q.pvbufs);
  if (!pl->ws) return 0;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t C = (cuuint64_t)q.C;
  {   // x: [B*N rows][C] fp16, box = 64 pixels x 64 channels
    cuuint64_t gdim[2] = {C, (cuuint64_t)pl->B * q.N};
    cuuint64_t gstr[1] = {C * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&op.maps.a[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)op.actx.x, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled(attention x, op %s) failed: %d", op.name.c_str(), (int)r);
  }
  {   // Wkv [C/64][2C][64]: view {64, rows, kv, chunk}; stacked: rows = 2C (K then V) and one kv plane
    const cuuint64_t rows = q.stacked ? 2 * C : C;
    cuuint64_t gdim[4] = {64, rows, (cuuint64_t)(q.stacked ? 1 : 2), (cuuint64_t)q.cpt};
    cuuint64_t gstr[3] = {128, C * 128, 2 * C * 128};
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&op.maps.b[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)op.actx.Wkv, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled(attention W, op %s) failed: %d", op.name.c_str(), (int)r);
  }
  for (int i = 1; i < kMaxSeg; ++i) { op.maps.a[i] = op.maps.a[0]; op.maps.b[i] = op.maps.b[0]; }
  return 0;
}

int build_plan(cdc_engine* e, Plan* pl, int B, int H, int W, uint8_t* wsp) {
  const cdc_config& cfg = e->cfg;
  const int L = cfg.n_levels;
  pl->B = B; pl->H = H; pl->W = W; pl->ws = wsp;
  pl->ctr_count = 0;
  Builder bd{e, pl};
  bd.B = B; bd.H = H; bd.W = W;
  bd.arena.no_reuse = e->debug_no_reuse;

  // ---- persistent region: context + shifts ----
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  pl->ctx_off.clear();
  for (int l = 0; l < cfg.n_context; ++l) {
    const int c = e->cdims[l], h = H >> l, w = W >> l;
    if (l == 0 && e->fold_ctx0) pl->ctx_off.push_back(take((size_t)B * c * h * w * 4));  // fp32 NCHW copy
    else pl->ctx_off.push_back(take((size_t)B * c * h * w * 2));
  }
  pl->shifts_off = take((size_t)B * e->R * 4);
  pl->xstate_off = take((size_t)B * cfg.channels * H * W * 4);
  bd.arena_base = off;

  // ---- context_fn.decode (optional): latent -> [ResnetBlock -> Upsample] x n, each Upsample output IS a context map of
  // the persistent region (fp16 NHWC, no conversion pass); built first so its activations share the arena with the U-Net's
  pl->ctx_ops.clear();
  if (!e->ctxdec.empty()) {
    const int n = (int)e->ctxdec.size();
    int h = H >> n, w = W >> n;
    Act x = bd.new_act(e->ctxdec[0].cin, h, w, true);
    bool x_in_arena = true;
    {
      pl->ops.emplace_back();
      Op& op = pl->ops.back();
      op.kind = OP_LATENT;
      op.name = "context_fn.latent";
      op.lat = {bd.ws<__half>(x.off), bd.ws<__half>(x.off_lo), x.C, h * w};
      op.dbg = bd.ws<__half>(x.off);
      op.dC = x.C; op.dH = h; op.dW = w;
    }
    auto drop_x = [&](const Act& a, bool whole) {
      if (whole) bd.drop(a);
      else if (a.off_lo) bd.arena.release(a.off_lo - bd.arena_base, a.bytes);   // hi lives in the persistent region
    };
    for (int i = 0; i < n; ++i) {
      const cdc_engine::CtxStage& stg = e->ctxdec[i];
      const int level = n - 1 - i;
      const std::string p = "context_fn.dec." + std::to_string(i) + ".";
      std::vector<Builder::SegIn> s1 = {{x, 3, 3, -1, -1}}, sr = {{x, 1, 1, 0, 0}};
      Act r = bd.resnet(p + "0.", s1, sr, stg.rb, h, w, nullptr);
      drop_x(x, x_in_arena);
      if (!stg.small_up) {
        Act u;
        u.C = stg.cout; u.H = 2 * h; u.W = 2 * w;
        u.bytes = (size_t)B * u.H * u.W * u.C * 2;
        u.off = pl->ctx_off[level];
        u.off_lo = i < n - 1 ? bd.arena_base + bd.arena.alloc(u.bytes) : 0;   // the next stage's residual path reads hi + lo
        std::vector<Builder::SegIn> su = {{r, 2, 2, 0, 0}};
        bd.conv(p + "up", Builder::three_pass_in(su), stg.up, EPI_BIAS, u, 1, 4);
        bd.drop(r);
        x = u;
        x_in_arena = false;
      } else {
        pl->ops.emplace_back();
        Op& op = pl->ops.back();
        op.kind = OP_UPSMALL;
        op.name = p + "up";
        op.ups.in = bd.ws<__half>(r.off);
        op.ups.in_lo = bd.lo_ptr<__half>(r);
        op.ups.wt = dptr<float>(e, stg.up_w32);
        op.ups.bias = dptr<float>(e, stg.up_b32);
        op.ups.out = bd.ws<float>(pl->ctx_off[level]);
        op.ups.B = B; op.ups.h = h; op.ups.w = w; op.ups.Cin = stg.cmid; op.ups.Cout = stg.cout;
        bd.drop(r);
        x = Act();
        x_in_arena = false;
      }
      h *= 2;
      w *= 2;
    }
    drop_x(x, x_in_arena);
    pl->ctx_ops.swap(pl->ops);
    pl->ops.clear();
  }

  // ---- ops ----
  {
    pl->ops.emplace_back();
    Op& op = pl->ops.back();
    op.kind = OP_TIME;
    op.name = "time_mlp";
    // side lane: the timestep MLP only feeds the shift vectors of the first block1 epilogue; it overlaps pack_input
    op.lane = (e->two_lanes && e->time_lane) ? 1 : 0;
    pl->time_op = (int)pl->ops.size() - 1;
    bd.pending_time_join = op.lane == 1;
  }
  Act x0 = bd.new_act(64, H, W);   // window form [B][H][W + 8][8] (1/8 of the allocation is used)
  x0.win8 = true;
  {
    pl->ops.emplace_back();
    Op& op = pl->ops.back();
    op.kind = OP_PACK;
    op.name = "pack_input";
    op.dbg = nullptr;   // window form: no [B][H][W][64] view to read back
    op.lat.hi = bd.ws<__half>(x0.off);
    pl->pack_op = (int)pl->ops.size() - 1;
  }
  auto ctx_act = [&](int l) {
    Act a;
    a.off = pl->ctx_off[l]; a.C = e->cdims[l]; a.H = H >> l; a.W = W >> l;
    a.bytes = 0;
    return a;
  };
  auto new_stats = [&](int h, int w, size_t* o, size_t* bytes) {
    *bytes = (size_t)B * h * w * 8;
    *o = bd.raw_alloc(*bytes);
    return bd.ws<float2>(*o);
  };

  std::vector<Act> skips;
  Act x = x0;
  for (int l = 0; l < L; ++l) {
    const Level& lv = e->downs[l];
    const int h = H >> l, w = W >> l;
    const std::string p = "downs." + std::to_string(l) + ".";
    std::vector<Builder::SegIn> s1, sr;
    const bool has_ctx = (l < L - 1) && (l < cfg.n_context);
    if (l == 0) {
      // the 7 vertical taps as three K segments of 3 + 2 + 2 rows (same weight chunk order): each fits the vertical-
      // reuse pipeline of the 3x3 convolutions (one activation box per segment and kx instead of one per tap)
      s1.push_back({x, 3, 1, -3, 0});
      s1.push_back({x, 2, 1, 0, 0});
      s1.push_back({x, 2, 1, 2, 0});
      sr.push_back({x, 1, 1, 0, 0});   // (hi weights, then the weight remainder as its own K chunk: window form has no spare centre copy)
      if (has_ctx && !e->fold_ctx0) {
        s1.push_back({ctx_act(0), 3, 7, -3, -3});
        s1.push_back({ctx_act(0), 2, 7, 0, -3});
        s1.push_back({ctx_act(0), 2, 7, 2, -3});
        sr.push_back({ctx_act(0), 1, 1, 0, 0});
      }
    } else {
      s1.push_back({x, 3, 3, -1, -1});
      sr.push_back({x, 1, 1, 0, 0});
      if (has_ctx) {
        s1.push_back({ctx_act(l), 3, 3, -1, -1});
        sr.push_back({ctx_act(l), 1, 1, 0, 0});
      }
    }
    Act a = bd.resnet(p + "0.", s1, sr, lv.rb0, h, w, nullptr);
    bd.drop(x);
    size_t so, sb;
    float2* st = new_stats(h, w, &so, &sb);
    std::vector<Builder::SegIn> s2 = {{a, 3, 3, -1, -1}}, sr2 = {{a, 1, 1, 0, 0}};
    Act bq = bd.resnet(p + "1.", s2, sr2, lv.rb1, h, w, st);
    bd.drop(a);
    Act c = bd.attention(p + "2.", bq, st, so, sb, lv.attn);
    bd.drop(bq);
    skips.push_back(c);
    if (lv.has_resample) {
      Act d = bd.new_act(lv.cout, h / 2, w / 2, true);
      std::vector<Builder::SegIn> sd = {{c, 3, 3, -1, -1}};
      bd.conv(p + "3.down", Builder::three_pass_in(sd), lv.resample, EPI_BIAS, d, 2, 0);
      x = d;
      if (l == 0) { bd.drop(c); }  // the level-0 skip is never consumed (unet.py:102 vs :113)
    } else {
      x = c;
    }
  }
  // After the loop: x is either the last level's attention output (== skips.back(), still owned by skips)
  // mid
  {
    const int h = H >> (L - 1), w = W >> (L - 1);
    std::vector<Builder::SegIn> s1 = {{x, 3, 3, -1, -1}}, sr = {{x, 1, 1, 0, 0}};
    size_t so, sb;
    float2* st = new_stats(h, w, &so, &sb);
    Act a = bd.resnet("mid_block1.", s1, sr, e->mid1, h, w, st);
    Act c = bd.attention("mid_attn.", a, st, so, sb, e->mid_attn);
    bd.drop(a);
    std::vector<Builder::SegIn> s2 = {{c, 3, 3, -1, -1}}, sr2 = {{c, 1, 1, 0, 0}};
    Act d = bd.resnet("mid_block2.", s2, sr2, e->mid2, h, w, nullptr);
    bd.drop(c);
    x = d;
  }
  Act xn;
  bool have_xn = false;
  for (int l = 0; l < L - 1; ++l) {
    const Level& lv = e->ups[l];
    const int lev = L - 1 - l;
    const int h = H >> lev, w = W >> lev;
    const std::string p = "ups." + std::to_string(l) + ".";
    Act skip = skips[lev];
    std::vector<Builder::SegIn> s1 = {{x, 3, 3, -1, -1}, {skip, 3, 3, -1, -1}};
    std::vector<Builder::SegIn> sr = {{x, 1, 1, 0, 0}, {skip, 1, 1, 0, 0}};
    Act a = bd.resnet(p + "0.", s1, sr, lv.rb0, h, w, nullptr);
    bd.drop(x);
    bd.drop(skip);
    size_t so, sb;
    float2* st = new_stats(h, w, &so, &sb);
    std::vector<Builder::SegIn> s2 = {{a, 3, 3, -1, -1}}, sr2 = {{a, 1, 1, 0, 0}};
    Act bq = bd.resnet(p + "1.", s2, sr2, lv.rb1, h, w, st);
    bd.drop(a);
    Act c = bd.attention(p + "2.", bq, st, so, sb, lv.attn);
    bd.drop(bq);
    Act u = bd.new_act(lv.cout, h * 2, w * 2, true);
    std::vector<Builder::SegIn> su = {{c, 2, 2, 0, 0}};
    Op& uop = bd.conv(p + "3.up", Builder::three_pass_in(su), lv.resample, EPI_BIAS, u, 1, 4);
    if (l == L - 2 && lv.cout == 64 && e->mainloop == 1 && e->final_tc && e->has_f_w2 && e->final_preln) {
      // the last Upsample feeds only the final LayerNorm + conv: its epilogue writes the normalised fp16 copy (and,
      // outside debug mode, nothing else), so final_conv_tc_kernel has no normalisation phase
      xn = bd.new_act(64, h * 2, w * 2);
      have_xn = true;
      uop.conv.ln_out = bd.ws<__half>(xn.off);
      uop.conv.ln_g = dptr<float>(e, e->f_g);
      uop.conv.ln_b = dptr<float>(e, e->f_b);
      uop.conv.skip_out = e->debug_no_reuse ? 0 : 1;
    }
    bd.drop(c);
    x = u;
  }
  {
    pl->ops.emplace_back();
    Op& op = pl->ops.back();
    op.kind = OP_FINAL;
    op.name = "final_conv";
    op.dbg = bd.ws<__half>(x.off);  // debug view: the final conv's INPUT activation
    op.dC = 64; op.dH = H; op.dW = W;
    op.flops = 2.0 * (double)B * H * W * cfg.channels * 64 * 49;
    pl->final_op = (int)pl->ops.size() - 1;
    // stash input pointer in conv.seg[0].src for the launcher
    op.conv.seg[0].src = bd.ws<__half>(x.off);
    op.conv.seg[1].src = bd.lo_ptr<__half>(x);
    op.conv.seg[2].src = have_xn ? bd.ws<__half>(xn.off) : nullptr;   // pre-normalised input (tcgen05 form)
  }
  bd.drop(x);
  if (have_xn) bd.drop(xn);
  pl->raw_off = (bd.arena_base + bd.arena.peak + 255) & ~size_t(255);
  pl->raw_bytes = 0;
  pl->raw2_bytes = 0;
  auto finish_ops = [&](std::vector<Op>& ops) -> int {
    std::vector<Op> ops2;
    ops2.reserve(ops.size() + 64);
    for (auto& op : ops) {
      if (op.kind == OP_ATTN_CTX && op.use_tc) {
        int rc = setup_attn_tc(e, pl, op);
        if (rc) return rc;
      }
      if (op.kind == OP_FINAL && op.conv.seg[2].src && pl->ws) {   // halo box of the pre-normalised input
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
        cuuint32_t box[4] = {64, (cuuint32_t)kFinalPitch, (cuuint32_t)kFinalHalo, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&op.maps.a[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)op.conv.seg[2].src, gdim, gstr, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled(final conv input) failed: %d", (int)r);
      }
      if (op.kind == OP_ALG && pl->ws) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        const cuuint64_t C = (cuuint64_t)op.alg.C;
        const __half* src[4] = {op.algsrc.ah, op.algsrc.al, op.algsrc.bh, op.algsrc.bl};
        for (int i = 0; i < 4; ++i) {
          const bool per_img = i < 2 ? op.alg.a_img != 0 : op.alg.b_img != 0;
          // MN-blocked operand [image][MN / 64][K][64]: view {64, K, MN blocks, image}
          cuuint64_t gdim[4] = {64, C, C / 64, (cuuint64_t)(per_img ? pl->B : 1)};
          cuuint64_t gstr[3] = {128, C * 128, C * C * 2};
          cuuint32_t box[4] = {64, 64, 2, 1};
          cuuint32_t estr[4] = {1, 1, 1, 1};
          CUtensorMap* m = i < 2 ? &op.maps.a[i] : &op.maps.b[i - 2];
          CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)src[i], gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS)
            return fail(e, CDC_ERR_CUDA, "cuTensorMapEncodeTiled(attention algebra operand %d, op %s) failed: %d", i, op.name.c_str(), (int)r);
        }
      }
      if (op.kind != OP_CONV) { ops2.push_back(op); continue; }
      int rc = setup_tc(e, pl, op);
      if (rc) return rc;
      if (!(op.use_tc && op.tcp.epi == EPI_RAW)) { ops2.push_back(op); continue; }
      if (op.ctr_index >= 0) { ops2.push_back(op); continue; }   // finishes its rows itself (TcConvParams::tile_ctr)
      // sliced convolution = raw partial tiles (this op) + row-wise epilogue (next op, keeps the op's name)
      const ConvParams& c = op.conv;
      Op fin;
      fin.kind = OP_LNROWS;
      fin.name = op.name;
      fin.dbg = op.dbg; fin.dC = op.dC; fin.dH = op.dH; fin.dW = op.dW;
      LnRowsParams& q = fin.lnr;
      q.raw = nullptr;   // patched below
      q.k_splits = op.tcp.k_splits;
      q.split_stride = op.tcp.raw_split_stride;
      q.N = c.Ntot; q.Ntot = c.Ntot;
      q.rows = (long long)pl->B * c.out_H * c.out_W;
      q.pix_per_image = c.out_H * c.out_W;
      q.epi = op.epi;
      q.bias = c.bias; q.ln_g = c.ln_g; q.ln_b = c.ln_b; q.shift = c.shift; q.shift_stride = c.shift_stride;
      q.res = c.res; q.res_C0 = c.res_C0; q.res2 = c.res2; q.res_lo = c.res_lo; q.res2_lo = c.res2_lo;
      q.out = c.out; q.out_lo = c.out_lo; q.stats_out = c.stats_out;
      fin.grid = dim3((unsigned)((q.rows + 7) / 8), 1, 1);
      fin.lane = op.lane;
      op.name += "#partials";
      op.dbg = nullptr;
      fin.flops = 0;
      ops2.push_back(op);
      ops2.push_back(fin);
    }
    ops.swap(ops2);
    return 0;
  };
  {
    int rc = finish_ops(pl->ctx_ops);
    if (rc) return rc;
    if ((rc = finish_ops(pl->ops))) return rc;
    for (size_t i = 0; i < pl->ops.size(); ++i) {
      if (pl->ops[i].kind == OP_TIME) pl->time_op = (int)i;
      if (pl->ops[i].kind == OP_PACK) pl->pack_op = (int)i;
      if (pl->ops[i].kind == OP_FINAL) pl->final_op = (int)i;
    }
  }
  pl->raw2_off = (pl->raw_off + pl->raw_bytes + 255) & ~size_t(255);
  pl->ctr_off = (pl->raw2_off + pl->raw2_bytes + 255) & ~size_t(255);
  for (std::vector<Op>* ops : {&pl->ctx_ops, &pl->ops})
    for (auto& op : *ops) {
      float* rawp = reinterpret_cast<float*>(pl->ws + (op.lane == 1 ? pl->raw2_off : pl->raw_off));
      if (op.kind == OP_CONV && op.use_tc && op.tcp.epi == EPI_RAW) op.tcp.raw = rawp;
      if (op.kind == OP_LNROWS) op.lnr.raw = rawp;
      if (op.kind == OP_CONV && op.ctr_index >= 0)
        op.tcp.tile_ctr = reinterpret_cast<unsigned int*>(pl->ws + pl->ctr_off) + op.ctr_index;
    }
  pl->total_bytes = pl->ctr_off + (size_t)pl->ctr_count * sizeof(unsigned int);
  pl->flops = 0;
  for (auto& op : pl->ops) pl->flops += op.flops;
  return 0;
}

Plan* get_plan(cdc_engine* e, int B, int H, int W, void* ws, int* rc) {
  *rc = 0;
  if (e->dry) { *rc = fail(e, CDC_ERR_STATE, "planning-only engine (device=-1) cannot compute: there is no CPU path"); return nullptr; }
  if (!e->finalized) { *rc = fail(e, CDC_ERR_STATE, "engine not finalized"); return nullptr; }
  if (B < 1 || H < 32 || W < 32 || (H % 32) || (W % 32)) {
    *rc = fail(e, CDC_ERR_INVALID, "unsupported shape B=%d H=%d W=%d (need B>=1, H,W multiples of 32)", B, H, W);
    return nullptr;
  }
  if ((H >> (e->cfg.n_levels - 1)) < 1) { *rc = fail(e, CDC_ERR_INVALID, "image too small"); return nullptr; }
  if ((long long)B * H * W * 64 >= (1ll << 31)) {
    *rc = fail(e, CDC_ERR_INVALID, "B*H*W too large for one engine call (%d x %d x %d): split the batch", B, H, W);
    return nullptr;
  }
  auto key = std::make_tuple(B, H, W, (uintptr_t)ws);
  auto it = e->plans.find(key);
  if (it != e->plans.end()) return it->second.get();
  if (e->plans.size() > 16) {
    for (auto& kv : e->plans) if (kv.second->graph) cudaGraphExecDestroy(kv.second->graph);
    e->plans.clear();
    e->last_plan = nullptr;
  }
  std::unique_ptr<Plan> pl(new Plan());
  *rc = build_plan(e, pl.get(), B, H, W, (uint8_t*)ws);
  if (*rc) return nullptr;
  Plan* raw = pl.get();
  e->plans[key] = std::move(pl);
  return raw;
}

// Programmatic dependent launch policy (CDC_PDL): 1 = every kernel of the step (default: every kernel issues
// griddepcontrol.launch_dependents first and griddepcontrol.wait after its prologue — barrier init, TMEM allocation,
// tensor-map prefetch, constant vectors — so that prologue overlaps the predecessor's tail), 0 = off, 2 = only the
// thread-block-cluster convolutions, 3 = mode 2 + every small-grid kernel.  Round 2, B=8 256x256: 2.631 (0) / 2.608 (2) /
// 2.590 (3) / 2.570 (1) ms per step.  (Round 1 measured mode 1 as a loss on a 4.7 ms step; launches were longer then.)
bool pdl_for(const cdc_engine* e, const Op& op) {
  if (e->pdl_mode <= 0 || e->profiling) return false;   // per-op profiling times every launch in plain stream order
  if (e->pdl_mode == 1) return true;
  const bool cluster_conv = op.kind == OP_CONV && op.use_tc && op.tcp.cluster_n > 1;
  if (e->pdl_mode == 2) return cluster_conv;
  const bool small = op.kind == OP_SGEMM || op.kind == OP_ALG || op.kind == OP_COMBINE || op.kind == OP_LNROWS || op.kind == OP_FINISH ||
                     (op.kind == OP_CONV && op.use_tc && op.tc_grid <= e->num_sms);
  return cluster_conv || small;
}

int run_op(cdc_engine* e, Plan* pl, const Op& op, size_t i, const RunArgs& a, cudaStream_t st) {
  const cdc_config& cfg = e->cfg;
  const int B = pl->B, H = pl->H, W = pl->W;
  {
    g_pdl = pdl_for(e, op);
    switch (op.kind) {
      case OP_TIME: {
        const size_t sm = (size_t)5 * cfg.dim * 4;
        launch_k(time_mlp_kernel, dim3(B, (e->R + 255) / 256), dim3(256), sm, st, a.time,
                 a.time ? (const cdc_step_coef*)nullptr : (const cdc_step_coef*)e->d_table, (const int*)e->d_step,
                 (const float*)dptr<float>(e, e->t_w1), (const float*)dptr<float>(e, e->t_b1),
                 (const float*)dptr<float>(e, e->t_w2), (const float*)dptr<float>(e, e->t_b2),
                 (const float*)dptr<float>(e, e->t_wcat), (const float*)dptr<float>(e, e->t_bcat), cfg.dim, e->R,
                 reinterpret_cast<float*>(pl->ws + pl->shifts_off));
        break;
      }
      case OP_PACK: {
        const long long total = (long long)B * H * (W + 8);
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 64);
        const float* c0 = e->fold_ctx0 ? reinterpret_cast<const float*>(pl->ws + pl->ctx_off[0]) : nullptr;
        launch_k(pack_input_kernel, dim3(blocks), dim3(256), 0, st, a.x, (int)cfg.channels, c0,
                 (int)(e->fold_ctx0 ? cfg.context_channels : 0), B, H, W, make_fastdiv((uint32_t)(W + 8)),
                 make_fastdiv((uint32_t)H), op.lat.hi);
        break;
      }
      case OP_CONV: {
        if (op.use_tc) {
          cudaError_t err = launch_tc(op, st);
          if (err != cudaSuccess)
            return fail(e, CDC_ERR_CUDA, "tcgen05 conv launch '%s': %s", op.name.c_str(), cudaGetErrorString(err));
          break;
        }
        cudaError_t err = launch_igemm(op, st);
        if (err != cudaSuccess)
          return fail(e, CDC_ERR_CUDA, "conv launch '%s' (bm=%d bn=%d epi=%d): %s", op.name.c_str(), op.bm, op.bn,
                      op.epi, cudaGetErrorString(err));
        break;
      }
      case OP_ATTN_CTX:
        if (op.use_tc)
          launch_k(attn_ctx_tc_kernel, op.grid, dim3(kAttnTcThreads), (size_t)op.atc_smem, st, op.maps, op.atc);
        else
          launch_k(attn_ctx_kernel, op.grid, dim3(256), (size_t)AttnCtxSmem::kBytes, st, op.actx);
        break;
      case OP_COMBINE:
        launch_k(attn_combine_kernel, op.grid, dim3(128), 0, st, op.comb.pc, op.comb.pm, op.comb.ps, op.comb.C,
                 op.comb.nchunks, op.comb.out, op.comb.o16h, op.comb.o16l);
        break;
      case OP_ALG:
        launch_k(attn_alg_tc_kernel, op.grid, dim3(kAlgThreads), (size_t)alg_tc_smem_bytes(op.alg.stages), st, op.maps, op.alg);
        break;
      case OP_SGEMM:
        launch_k(gemm3xf16_tn_kernel, op.grid, dim3(128), (size_t)Gemm3xSmem::kBytes, st, op.sg.At, op.sg.Bm,
                 op.sg.Cout, op.sg.M, op.sg.N, op.sg.K, op.sg.sA, op.sg.sB, op.sg.sC, op.gfin);
        break;
      case OP_LNROWS:
        if (e->lnrows_hoist && (int)op.grid.x <= e->num_sms) launch_k(ln_rows_kernel<true>, op.grid, dim3(256), 0, st, op.lnr);
        else launch_k(ln_rows_kernel<false>, op.grid, dim3(256), 0, st, op.lnr);
        break;
      case OP_FINISH:
        launch_k(attn_finish_kernel, op.grid, dim3(128), 0, st, op.fin.Mf, op.fin.g, op.fin.bln, op.fin.bout, op.fin.C,
                 op.fin.Mg, op.fin.um, op.fin.cm);
        break;
      case OP_FINAL: {
        FinalParams fp{};
        fp.in = op.conv.seg[0].src;
        fp.in_lo = op.conv.seg[1].src;
        fp.ln_g = dptr<float>(e, e->f_g);
        fp.ln_b = dptr<float>(e, e->f_b);
        fp.Wf = dptr<__half>(e, e->f_w);
        fp.bias = dptr<float>(e, e->f_bias);
        fp.B = B; fp.H = H; fp.W = W; fp.channels = cfg.channels;
        fp.mode = a.mode;
        fp.out = a.out;
        fp.x = a.x_inout;
        fp.z = a.z;
        fp.z_first = a.z_chunk ? e->d_zfirst : nullptr;
        fp.z_stride = (long long)B * cfg.channels * H * W;
        fp.table = e->d_table;
        fp.step_ptr = e->d_step;
        fp.variant = cfg.variant;
        fp.pred_mode = a.pred;
        fp.clip_mode = a.clip;
        if (e->has_f_w2 && e->final_tc && e->mainloop == 1) {
          fp.Wf = dptr<__half>(e, e->f_w3);
          if (op.conv.seg[2].src && e->final_persist) {
            const int total = (W / 16) * (H / 16) * B;
            launch_k(final_conv_tc_persist_kernel, dim3(std::min(total, e->num_sms)), dim3(kFinalPersistThreads),
                     (size_t)kFinalSmemBytesP, st, fp, op.maps.a[0], W / 16, H / 16, total);
          } else if (op.conv.seg[2].src)
            launch_k(final_conv_tc_kernel<true>, dim3(W / 16, H / 16, B), dim3(256), (size_t)kFinalSmemBytes3, st, fp,
                     op.maps.a[0]);
          else
            launch_k(final_conv_tc_kernel<false>, dim3(W / 16, H / 16, B), dim3(256), (size_t)kFinalSmemBytes3, st, fp,
                     op.maps.a[0]);
        } else if (e->has_f_w2 && e->final_kx) {
          fp.Wf = dptr<__half>(e, e->f_w2);
          launch_k(final_conv_kx_kernel, dim3(W / 16, H / 16, B), dim3(256), (size_t)kFinalSmemBytes2, st, fp);
        } else {
          launch_k(final_conv_kernel, dim3(W / 16, H / 16, B), dim3(256), (size_t)kFinalSmemBytes, st, fp);
        }
        break;
      }
      case OP_LATENT: {
        if (!a.latent) return fail(e, CDC_ERR_INVALID, "context decode without a latent");
        dim3 grid((op.lat.HW + 31) / 32, (op.lat.C + 31) / 32, B);
        launch_k(nchw_to_nhwc_hilo_kernel, grid, dim3(256), 0, st, a.latent, op.lat.C, op.lat.HW, op.lat.hi, op.lat.lo);
        break;
      }
      case OP_UPSMALL: {
        const long long total = (long long)op.ups.B * op.ups.h * op.ups.w * 4;
        const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)e->num_sms * 8);
        launch_k(upsample_small_kernel, dim3(blocks), dim3(256), (size_t)16 * op.ups.Cin * op.ups.Cout * 4, st, op.ups);
        break;
      }
      default:
        break;
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
      return fail(e, CDC_ERR_CUDA, "launch of op %zu '%s' failed: %s", i, op.name.c_str(), cudaGetErrorString(err));
  }
  return 0;
}

int run_ops(cdc_engine* e, Plan* pl, const std::vector<Op>& ops, const RunArgs& a, cudaStream_t st) {
  bool side_busy = false;
  for (size_t i = 0; i < ops.size(); ++i) {
    const Op& op = ops[i];
    cudaStream_t use = st;
    if (op.lane == 1) {
      if (!side_busy) {   // fork: the side lane starts after everything enqueued on the main lane so far
        CUDA_TRY(e, cudaEventRecord(e->ev_fork, st));
        CUDA_TRY(e, cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0));
        side_busy = true;
      }
      use = e->side_stream;
    } else if (op.join_before && side_busy) {
      CUDA_TRY(e, cudaEventRecord(e->ev_join, e->side_stream));
      CUDA_TRY(e, cudaStreamWaitEvent(st, e->ev_join, 0));
      side_busy = false;
    }
    int rc = run_op(e, pl, op, i, a, use);
    if (rc) return rc;
  }
  if (side_busy) {
    CUDA_TRY(e, cudaEventRecord(e->ev_join, e->side_stream));
    CUDA_TRY(e, cudaStreamWaitEvent(st, e->ev_join, 0));
  }
  return 0;
}

int run_plan(cdc_engine* e, Plan* pl, const RunArgs& a, cudaStream_t st) {
  int rc = run_ops(e, pl, pl->ops, a, st);
  if (rc) return rc;
  if (a.advance) {
    g_pdl = e->pdl_mode == 1;
    launch_k(advance_step_kernel, dim3(1), dim3(32), 0, st, e->d_step);
  }
  e->last_plan = pl;
  e->last_args = a;
  return 0;
}

__global__ void set_int_kernel(int* p, int v) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *p = v;
}

int reset_counters(cdc_engine* e, Plan* pl, cudaStream_t st) {
  if (pl->ctr_count > 0)   // arrival counters of the K-split convolutions: zero whenever the workspace is (re)initialised
    CUDA_TRY(e, cudaMemsetAsync(pl->ws + pl->ctr_off, 0, (size_t)pl->ctr_count * sizeof(unsigned int), st));
  return 0;
}

int convert_context(cdc_engine* e, Plan* pl, const float* const* ctx, int n_ctx, cudaStream_t st) {
  const cdc_config& cfg = e->cfg;
  if (n_ctx != cfg.n_context) return fail(e, CDC_ERR_INVALID, "expected %d context tensors, got %d", cfg.n_context, n_ctx);
  if (int rc = reset_counters(e, pl, st)) return rc;
  for (int l = 0; l < n_ctx; ++l) {
    const int c = e->cdims[l], h = pl->H >> l, w = pl->W >> l;
    if (!ctx[l]) return fail(e, CDC_ERR_INVALID, "context[%d] is null", l);
    if (l == 0 && e->fold_ctx0) {
      CUDA_TRY(e, cudaMemcpyAsync(pl->ws + pl->ctx_off[0], ctx[0], (size_t)pl->B * c * h * w * 4,
                                  cudaMemcpyDeviceToDevice, st));
    } else {
      dim3 grid((h * w + 31) / 32, (c + 31) / 32, pl->B);
      nchw_to_nhwc_half_kernel<<<grid, 256, 0, st>>>(ctx[l], c, h * w, reinterpret_cast<__half*>(pl->ws + pl->ctx_off[l]));
      CUDA_TRY(e, cudaGetLastError());
    }
  }
  return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int cdc_abi_version(void) { return CDC_ABI_VERSION; }

const char* cdc_last_error(const cdc_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int cdc_engine_create(const cdc_config* cfg, int device, cdc_engine** out) {
  if (!cfg || !out) return fail(nullptr, CDC_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != CDC_ABI_VERSION) return fail(nullptr, CDC_ERR_INVALID, "ABI version mismatch");
  if (cfg->variant != CDC_VARIANT_EPS && cfg->variant != CDC_VARIANT_X)
    return fail(nullptr, CDC_ERR_INVALID, "unknown variant %d", cfg->variant);
  if (cfg->dim != 64)
    return fail(nullptr, CDC_ERR_UNSUPPORTED, "dim=%d unsupported: the kernel family is specialised for dim=64", cfg->dim);
  if (cfg->n_levels < 2 || cfg->n_levels > CDC_MAX_LEVELS) return fail(nullptr, CDC_ERR_UNSUPPORTED, "n_levels=%d unsupported", cfg->n_levels);
  if (cfg->n_context < 0 || cfg->n_context > cfg->n_levels - 1)
    return fail(nullptr, CDC_ERR_UNSUPPORTED, "n_context=%d unsupported", cfg->n_context);
  if (cfg->channels < 1 || cfg->channels > 8) return fail(nullptr, CDC_ERR_UNSUPPORTED, "channels=%d unsupported", cfg->channels);
  for (int i = 0; i < cfg->n_levels; ++i)
    if (cfg->dim_mults[i] < 1 || cfg->dim_mults[i] > 6)
      return fail(nullptr, CDC_ERR_UNSUPPORTED, "dim_mults[%d]=%d unsupported (1..6)", i, cfg->dim_mults[i]);
  for (int i = 0; i < cfg->n_context; ++i)
    if (cfg->context_dim_mults[i] < 1 || cfg->context_dim_mults[i] > 6)
      return fail(nullptr, CDC_ERR_UNSUPPORTED, "context_dim_mults[%d] unsupported", i);
  std::unique_ptr<cdc_engine> e(new cdc_engine());
  e->cfg = *cfg;
  e->device = device;
  e->dims.push_back(cfg->channels);
  for (int i = 0; i < cfg->n_levels; ++i) e->dims.push_back(cfg->dim * cfg->dim_mults[i]);
  e->cdims.push_back(cfg->context_channels);
  for (int i = 0; i + 1 < cfg->n_context; ++i) e->cdims.push_back(cfg->dim * cfg->context_dim_mults[i]);
  if (cfg->n_context > 0) {
    if (cfg->context_channels % 64 == 0 && cfg->context_channels <= 384) e->fold_ctx0 = false;
    else if (cfg->channels + cfg->context_channels <= 8) e->fold_ctx0 = true;
    else return fail(nullptr, CDC_ERR_UNSUPPORTED, "context_channels=%d unsupported (multiple of 64, or channels+context_channels<=8)", cfg->context_channels);
  }
  if (device == -1) {
    // host-only planning engine: validates weights, sizes workspaces, counts launches/FLOPs; cannot compute.
    e->dry = true;
    *out = e.release();
    return CDC_OK;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, CDC_ERR_CUDA, "no CUDA device: the CDC engine has no CPU path");
  if (device < 0 || device >= ndev) return fail(nullptr, CDC_ERR_INVALID, "device %d out of range", device);
  DeviceGuard dg(device);
  {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device) return fail(nullptr, CDC_ERR_CUDA, "cudaSetDevice failed");
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) return fail(nullptr, CDC_ERR_UNSUPPORTED, "device sm_%d%d: this build targets sm_100a only", prop.major, prop.minor);
  if (cudaMalloc(&e->d_step, sizeof(int)) != cudaSuccess || cudaMalloc(&e->d_zfirst, sizeof(int)) != cudaSuccess)
    return fail(nullptr, CDC_ERR_CUDA, "cudaMalloc failed");
  if (cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess)
    return fail(nullptr, CDC_ERR_CUDA, "stream/event creation failed");
  if (cudaStreamCreateWithFlags(&e->loop_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_table, cudaEventDisableTiming) != cudaSuccess)
    return fail(nullptr, CDC_ERR_CUDA, "stream/event creation failed");
  e->num_sms = prop.multiProcessorCount;
  if (const char* v = getenv("CDC_SLICED")) e->sliced = atoi(v) != 0;
  if (const char* v = getenv("CDC_NSLICE")) e->nslice = atoi(v) != 0;
  if (const char* v = getenv("CDC_FUSE_LNROWS")) e->fuse_lnrows = atoi(v) != 0;
  if (const char* v = getenv("CDC_ATTN_TC")) e->attn_tc = atoi(v) != 0;
  if (const char* v = getenv("CDC_ALG_TC")) e->alg_tc = atoi(v) != 0;
  if (const char* v = getenv("CDC_ATTN_AREA")) e->attn_area = atoi(v) != 0;
  if (const char* v = getenv("CDC_FUSE_RES")) e->fuse_res = atoi(v) != 0;
  if (const char* v = getenv("CDC_DUAL_PASS")) e->dual_pass = atoi(v) != 0;
  if (const char* v = getenv("CDC_SLICE_MAXTILES")) e->slice_max_tiles = std::max(1, atoi(v));
  if (const char* v = getenv("CDC_FOLD_FINISH")) e->fold_finish = atoi(v) != 0;
  if (const char* v = getenv("CDC_TWO_LANES")) e->two_lanes = atoi(v) != 0;
  if (const char* v = getenv("CDC_TIME_LANE")) e->time_lane = atoi(v) != 0;
  if (const char* v = getenv("CDC_LNROWS_HOIST")) e->lnrows_hoist = atoi(v) != 0;
  if (const char* v = getenv("CDC_VREUSE")) e->vreuse = atoi(v);
  if (const char* v = getenv("CDC_PDL")) e->pdl_mode = atoi(v);
  if (const char* v = getenv("CDC_SLICE_SLOTS")) e->slice_slots = std::max(1, atoi(v));
  if (const char* v = getenv("CDC_SLICE_KMAX")) e->slice_kmax = std::max(1, atoi(v));
  cudaFuncSetAttribute(attn_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCtxSmem::kBytes);
  cudaFuncSetAttribute(attn_ctx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(attn_alg_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(gemm3xf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm3xSmem::kBytes);
  cudaFuncSetAttribute(final_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinalSmemBytes);
  cudaFuncSetAttribute(final_conv_kx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinalSmemBytes2);
  cudaFuncSetAttribute(final_conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinalSmemBytes3);
  cudaFuncSetAttribute(final_conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinalSmemBytes3);
  cudaFuncSetAttribute(final_conv_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinalSmemBytesP);
  if (const char* v = getenv("CDC_FINAL_PERSIST")) e->final_persist = atoi(v) != 0;
  if (const char* v = getenv("CDC_FINAL_PRELN")) e->final_preln = atoi(v) != 0;
  if (const char* v = getenv("CDC_FINAL_TC")) e->final_tc = atoi(v) != 0;
  if (const char* v = getenv("CDC_FINAL_KX")) e->final_kx = atoi(v) != 0;
  *out = e.release();
  return CDC_OK;
}

void cdc_engine_destroy(cdc_engine* e) {
  if (!e) return;
  if (e->dry) { delete e; return; }
  DeviceGuard dg(e->device);
  for (auto& kv : e->plans) if (kv.second->graph) cudaGraphExecDestroy(kv.second->graph);
  if (e->dblob) cudaFree(e->dblob);
  if (e->d_table) cudaFree(e->d_table);
  if (e->d_step) cudaFree(e->d_step);
  if (e->d_zfirst) cudaFree(e->d_zfirst);
  if (e->loop_stream) cudaStreamDestroy(e->loop_stream);
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  if (e->ev_table) { cudaEventSynchronize(e->ev_table); cudaEventDestroy(e->ev_table); }
  if (e->h_table) cudaFreeHost(e->h_table);
  delete e;
}

int cdc_engine_set_weight(cdc_engine* e, const char* key, const float* host_ptr, const int64_t* shape, int ndim) {
  if (!e || !key || !host_ptr || !shape || ndim < 1 || ndim > 8) return fail(e, CDC_ERR_INVALID, "bad set_weight argument");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.data.assign(host_ptr, host_ptr + n);
  e->weights[key] = std::move(t);
  e->finalized = false;
  return CDC_OK;
}

int cdc_engine_finalize(cdc_engine* e) {
  if (!e) return CDC_ERR_INVALID;
  DeviceGuard dg(e->dry ? -1 : e->device);
  const cdc_config& cfg = e->cfg;
  const int L = cfg.n_levels, dim = cfg.dim;
  e->blob.host.clear();
  e->downs.assign(L, Level());
  e->ups.assign(L - 1, Level());
  std::vector<float> wcat, bcat;
  int rc = 0;
  // time MLP
  {
    const HostTensor* w1 = find(e, "time_mlp.0.weight", {4 * dim, 1}, &rc); if (!w1) return rc;
    const HostTensor* b1 = find(e, "time_mlp.0.bias", {4 * dim}, &rc); if (!b1) return rc;
    const HostTensor* w2 = find(e, "time_mlp.2.weight", {dim, 4 * dim}, &rc); if (!w2) return rc;
    const HostTensor* b2 = find(e, "time_mlp.2.bias", {dim}, &rc); if (!b2) return rc;
    e->t_w1 = put_f32(e, w1->data.data(), w1->data.size());
    e->t_b1 = put_f32(e, b1->data.data(), b1->data.size());
    e->t_w2 = put_f32(e, w2->data.data(), w2->data.size());
    e->t_b2 = put_f32(e, b2->data.data(), b2->data.size());
  }
  for (int l = 0; l < L; ++l) {
    Level& lv = e->downs[l];
    const bool last = l >= L - 1;
    const bool has_ctx = !last && l < cfg.n_context;
    const int cx = e->dims[l], cc = has_ctx ? e->cdims[l] : 0;
    const int cin = cx + cc, cout = e->dims[l + 1];
    lv.cin = cin; lv.cout = cout;
    const std::string p = "downs." + std::to_string(l) + ".";
    std::vector<SegSpec> s1, sr;
    if (l == 0) {
      const int folded = (has_ctx && e->fold_ctx0) ? cc : 0;
      s1.push_back({64, 7, 1, 1, 0, cx + folded});
      sr.push_back({64, 1, 1, 2, 0, cx + folded});
      if (has_ctx && !e->fold_ctx0) {
        s1.push_back({cc, 7, 7, 0, cx, 0});
        sr.push_back({cc, 1, 1, 0, cx, 0});
      }
    } else {
      s1.push_back({cx, 3, 3, 0, 0, 0});
      sr.push_back({cx, 1, 1, 0, 0, 0});
      if (has_ctx) {
        s1.push_back({cc, 3, 3, 0, cx, 0});
        sr.push_back({cc, 1, 1, 0, cx, 0});
      }
    }
    // compensation ("lo") halves exist for trunk activations only: not for the packed input / context maps
    std::vector<bool> sr_lo(sr.size(), false);
    if (l > 0) sr_lo[0] = true;
    if ((rc = pack_resnet(e, p + "0.", cin, cout, l == 0 ? 7 : 3, s1, sr, sr_lo, &lv.rb0, &wcat, &bcat))) return rc;
    std::vector<SegSpec> s2 = {{cout, 3, 3, 0, 0, 0}}, sr2 = {{cout, 1, 1, 0, 0, 0}};
    if ((rc = pack_resnet(e, p + "1.", cout, cout, 3, s2, sr2, {true}, &lv.rb1, &wcat, &bcat))) return rc;
    if ((rc = pack_attn(e, p + "2.", cout, &lv.attn))) return rc;
    lv.has_resample = !last;
    if (!last) {
      std::vector<SegSpec> sd = three_pass({{cout, 3, 3, 0, 0, 0}}, {true});
      if ((rc = pack_conv(e, p + "3.conv", cout, cout, 3, 3, sd, true, &lv.resample))) return rc;
    }
  }
  {
    const int c = e->dims[L];
    std::vector<SegSpec> s = {{c, 3, 3, 0, 0, 0}}, sr = {{c, 1, 1, 0, 0, 0}};
    if ((rc = pack_resnet(e, "mid_block1.", c, c, 3, s, sr, {true}, &e->mid1, &wcat, &bcat))) return rc;
    if ((rc = pack_attn(e, "mid_attn.", c, &e->mid_attn))) return rc;
    if ((rc = pack_resnet(e, "mid_block2.", c, c, 3, s, sr, {true}, &e->mid2, &wcat, &bcat))) return rc;
  }
  for (int l = 0; l < L - 1; ++l) {
    Level& lv = e->ups[l];
    const int dim_in = e->dims[L - 1 - l], dim_out = e->dims[L - l];
    lv.cin = 2 * dim_out; lv.cout = dim_in;
    const std::string p = "ups." + std::to_string(l) + ".";
    std::vector<SegSpec> s1 = {{dim_out, 3, 3, 0, 0, 0}, {dim_out, 3, 3, 0, dim_out, 0}};
    std::vector<SegSpec> sr = {{dim_out, 1, 1, 0, 0, 0}, {dim_out, 1, 1, 0, dim_out, 0}};
    if ((rc = pack_resnet(e, p + "0.", 2 * dim_out, dim_in, 3, s1, sr, {true, true}, &lv.rb0, &wcat, &bcat))) return rc;
    std::vector<SegSpec> s2 = {{dim_in, 3, 3, 0, 0, 0}}, sr2 = {{dim_in, 1, 1, 0, 0, 0}};
    if ((rc = pack_resnet(e, p + "1.", dim_in, dim_in, 3, s2, sr2, {true}, &lv.rb1, &wcat, &bcat))) return rc;
    if ((rc = pack_attn(e, p + "2.", dim_in, &lv.attn))) return rc;
    lv.has_resample = true;
    if ((rc = pack_convT(e, p + "3.conv", dim_in, dim_in, &lv.resample))) return rc;
  }
  // final LayerNorm + 7x7 conv
  {
    if ((rc = pack_ln(e, "final_conv.0", dim, &e->f_g, &e->f_b))) return rc;
    const HostTensor* w = find(e, "final_conv.1.weight", {cfg.channels, dim, 7, 7}, &rc); if (!w) return rc;
    const HostTensor* b = find(e, "final_conv.1.bias", {cfg.channels}, &rc); if (!b) return rc;
    std::vector<__half> wf((size_t)8 * kFinalWStride, __float2half(0.f));
    for (int n = 0; n < cfg.channels; ++n)
      for (int c = 0; c < 64; ++c)
        for (int ky = 0; ky < 7; ++ky)
          for (int kx = 0; kx < 7; ++kx)
            wf[(size_t)n * kFinalWStride + (ky * 7 + kx) * 64 + c] =
                __float2half_rn(w->data[(((size_t)n * dim + c) * 7 + ky) * 7 + kx]);
    e->f_w = e->blob.reserve(wf.size() * 2);
    memcpy(e->blob.at<__half>(e->f_w), wf.data(), wf.size() * 2);
    // channels <= 3: horizontal taps folded into N (final_conv_kx_kernel): row n = kx*channels + ch, k = ky*64 + c
    e->f_w2 = 0;
    if (cfg.channels <= 3) {
      std::vector<__half> w2((size_t)24 * kFinalW2Stride, __float2half(0.f));
      for (int n = 0; n < cfg.channels; ++n)
        for (int c = 0; c < 64; ++c)
          for (int ky = 0; ky < 7; ++ky)
            for (int kx = 0; kx < 7; ++kx)
              w2[(size_t)(kx * cfg.channels + n) * kFinalW2Stride + ky * 64 + c] =
                  __float2half_rn(w->data[(((size_t)n * dim + c) * 7 + ky) * 7 + kx]);
      e->f_w2 = e->blob.reserve(w2.size() * 2);
      memcpy(e->blob.at<__half>(e->f_w2), w2.data(), w2.size() * 2);
      e->has_f_w2 = true;
      // tcgen05 form (final_conv_tc_kernel): [ky][32 rows n][64 c], stored as the SWIZZLE_128B shared-memory image
      std::vector<__half> w3((size_t)kFinalW3Bytes / 2, __float2half(0.f));
      for (int n = 0; n < cfg.channels; ++n)
        for (int c = 0; c < 64; ++c)
          for (int ky = 0; ky < 7; ++ky)
            for (int kx = 0; kx < 7; ++kx) {
              const int row = kx * cfg.channels + n;
              const size_t byte = (size_t)ky * 4096 + (size_t)row * 128 + ((((c >> 3) ^ (row & 7)) << 4)) + (c & 7) * 2;
              w3[byte / 2] = __float2half_rn(w->data[(((size_t)n * dim + c) * 7 + ky) * 7 + kx]);
            }
      e->f_w3 = e->blob.reserve(w3.size() * 2);
      memcpy(e->blob.at<__half>(e->f_w3), w3.data(), w3.size() * 2);
    }
    e->f_bias = put_f32(e, b->data.data(), cfg.channels);
    std::vector<__half> ident((size_t)64 * 64, __float2half(0.f));
    for (int i = 0; i < 64; ++i) ident[(size_t)i * 64 + i] = __float2half(1.f);
    e->w_ident = e->blob.reserve(ident.size() * 2);
    memcpy(e->blob.at<__half>(e->w_ident), ident.data(), ident.size() * 2);
  }
  if ((rc = pack_context_decoder(e))) return rc;
  e->R = (int)bcat.size();
  e->t_wcat = put_f32(e, wcat.data(), wcat.size());
  e->t_bcat = put_f32(e, bcat.data(), bcat.size());
  if (e->dry) {
    e->plans.clear();
    e->last_plan = nullptr;
    e->finalized = true;
    return CDC_OK;
  }
  // upload
  if (e->dblob && e->dblob_bytes < e->blob.host.size()) { cudaFree(e->dblob); e->dblob = nullptr; }
  if (!e->dblob) {
    CUDA_TRY(e, cudaMalloc(&e->dblob, e->blob.host.size()));
    e->dblob_bytes = e->blob.host.size();
  }
  CUDA_TRY(e, cudaMemcpy(e->dblob, e->blob.host.data(), e->blob.host.size(), cudaMemcpyHostToDevice));
  for (auto& kv : e->plans) if (kv.second->graph) cudaGraphExecDestroy(kv.second->graph);
  e->plans.clear();
  e->last_plan = nullptr;
  e->ctx_set = false;
  e->finalized = true;
  return CDC_OK;
}

int64_t cdc_engine_workspace_bytes(cdc_engine* e, int B, int H, int W) {
  if (!e) return CDC_ERR_INVALID;
  if (!e->finalized) return fail(e, CDC_ERR_STATE, "engine not finalized");
  if (B < 1 || H < 32 || W < 32 || (H % 32) || (W % 32))
    return fail(e, CDC_ERR_INVALID, "unsupported shape B=%d H=%d W=%d (need H,W multiples of 32)", B, H, W);
  Plan pl;
  int rc = build_plan(e, &pl, B, H, W, nullptr);
  if (rc) return rc;
  return (int64_t)pl.total_bytes;
}

static int check_ws(cdc_engine* e, Plan* pl, int64_t workspace_bytes) {
  if ((int64_t)pl->total_bytes > workspace_bytes)
    return fail(e, CDC_ERR_INVALID, "workspace too small: need %zu bytes, got %lld", pl->total_bytes, (long long)workspace_bytes);
  return 0;
}

int cdc_unet_forward(cdc_engine* e, const float* x, const float* time, const float* const* ctx, int n_ctx, float* out,
                     int B, int H, int W, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!e) return CDC_ERR_INVALID;
  if (!x || !time || !out || !workspace) return fail(e, CDC_ERR_INVALID, "null pointer argument");
  DeviceGuard dg(e->device);
  int rc;
  Plan* pl = get_plan(e, B, H, W, workspace, &rc);
  if (!pl) return rc;
  if ((rc = check_ws(e, pl, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = convert_context(e, pl, ctx, n_ctx, st))) return rc;
  e->ctx_set = false;  // forward() owns the persistent region for this call only
  RunArgs a;
  a.x = x; a.time = time; a.out = out; a.mode = 0;
  return run_plan(e, pl, a, st);
}

int cdc_set_context(cdc_engine* e, const float* const* ctx, int n_ctx, int B, int H, int W, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  if (!e) return CDC_ERR_INVALID;
  if (!workspace) return fail(e, CDC_ERR_INVALID, "null workspace");
  DeviceGuard dg(e->device);
  int rc;
  Plan* pl = get_plan(e, B, H, W, workspace, &rc);
  if (!pl) return rc;
  if ((rc = check_ws(e, pl, workspace_bytes))) return rc;
  if ((rc = convert_context(e, pl, ctx, n_ctx, (cudaStream_t)stream))) return rc;
  e->ctx_set = true;
  e->ctx_B = B; e->ctx_H = H; e->ctx_W = W; e->ctx_ws = (uintptr_t)workspace;
  return CDC_OK;
}

int cdc_engine_has_context_decoder(cdc_engine* e) {
  if (!e) return CDC_ERR_INVALID;
  if (!e->finalized) return fail(e, CDC_ERR_STATE, "engine not finalized");
  return e->ctxdec.empty() ? 0 : 1;
}

int cdc_context_decode(cdc_engine* e, const float* q_latent, int B, int H, int W, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  if (!e) return CDC_ERR_INVALID;
  if (!workspace || !q_latent) return fail(e, CDC_ERR_INVALID, "null pointer argument");
  DeviceGuard dg(e->device);
  int rc;
  Plan* pl = get_plan(e, B, H, W, workspace, &rc);
  if (!pl) return rc;
  if ((rc = check_ws(e, pl, workspace_bytes))) return rc;
  if (pl->ctx_ops.empty())
    return fail(e, CDC_ERR_STATE, "no context decoder weights registered (context_fn.dec.* keys before cdc_engine_finalize)");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = reset_counters(e, pl, st))) return rc;
  RunArgs a;
  a.latent = q_latent;
  if ((rc = run_ops(e, pl, pl->ctx_ops, a, st))) return rc;
  e->ctx_set = true;
  e->ctx_B = B; e->ctx_H = H; e->ctx_W = W; e->ctx_ws = (uintptr_t)workspace;
  return CDC_OK;
}

int cdc_engine_read_context(cdc_engine* e, int level, float* out, int B, int H, int W, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  if (!e) return CDC_ERR_INVALID;
  if (!workspace || !out) return fail(e, CDC_ERR_INVALID, "null pointer argument");
  DeviceGuard dg(e->device);
  if (!e->ctx_set || e->ctx_B != B || e->ctx_H != H || e->ctx_W != W || e->ctx_ws != (uintptr_t)workspace)
    return fail(e, CDC_ERR_STATE, "no context set for this shape and workspace");
  if (level < 0 || level >= e->cfg.n_context) return fail(e, CDC_ERR_INVALID, "context level %d out of range", level);
  int rc;
  Plan* pl = get_plan(e, B, H, W, workspace, &rc);
  if (!pl) return rc;
  if ((rc = check_ws(e, pl, workspace_bytes))) return rc;
  const int c = e->cdims[level], h = H >> level, w = W >> level;
  cudaStream_t st = (cudaStream_t)stream;
  if (level == 0 && e->fold_ctx0) {
    CUDA_TRY(e, cudaMemcpyAsync(out, pl->ws + pl->ctx_off[0], (size_t)B * c * h * w * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    dim3 grid((h * w + 31) / 32, (c + 31) / 32, B);
    nhwc_half_to_nchw_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(pl->ws + pl->ctx_off[level]), c, h * w, out);
    CUDA_TRY(e, cudaGetLastError());
  }
  return CDC_OK;
}

int cdc_set_schedule(cdc_engine* e, const cdc_step_coef* host_coefs, int S, void* stream) {
  if (!e || !host_coefs || S < 1) return fail(e, CDC_ERR_INVALID, "bad schedule");
  if (e->dry) return fail(e, CDC_ERR_STATE, "planning-only engine (device=-1) cannot compute: there is no CPU path");
  DeviceGuard dg(e->device);
  if (S > e->table_cap) {
    if (e->d_table) cudaFree(e->d_table);
    e->d_table = nullptr;
    CUDA_TRY(e, cudaMalloc(&e->d_table, (size_t)S * sizeof(cdc_step_coef)));
    e->table_cap = S;
    // graphs captured the old table pointer
    for (auto& kv : e->plans) if (kv.second->graph) { cudaGraphExecDestroy(kv.second->graph); kv.second->graph = nullptr; }
  }
  // The caller's buffer may be freed right after return: stage it in the engine's pinned buffer and copy on the
  // caller's stream (ordered after any decode already enqueued there).  Only a second set_schedule issued while the
  // previous copy is still in flight waits (for that copy alone) — the stream is never synchronised.
  if (S > e->h_table_cap) {
    if (e->h_table) { cudaEventSynchronize(e->ev_table); cudaFreeHost(e->h_table); e->h_table = nullptr; }
    CUDA_TRY(e, cudaHostAlloc((void**)&e->h_table, (size_t)S * sizeof(cdc_step_coef), cudaHostAllocDefault));
    e->h_table_cap = S;
  } else {
    CUDA_TRY(e, cudaEventSynchronize(e->ev_table));
  }
  memcpy(e->h_table, host_coefs, (size_t)S * sizeof(cdc_step_coef));
  CUDA_TRY(e, cudaMemcpyAsync(e->d_table, e->h_table, (size_t)S * sizeof(cdc_step_coef), cudaMemcpyHostToDevice,
                              (cudaStream_t)stream));
  CUDA_TRY(e, cudaEventRecord(e->ev_table, (cudaStream_t)stream));
  e->S = S;
  return CDC_OK;
}

static int sampler_prologue(cdc_engine* e, int B, int H, int W, void* workspace, int64_t workspace_bytes, int pred_mode,
                            int clip_mode, Plan** out) {
  if (!e->ctx_set || e->ctx_B != B || e->ctx_H != H || e->ctx_W != W || e->ctx_ws != (uintptr_t)workspace)
    return fail(e, CDC_ERR_STATE, "cdc_set_context must be called first with the same shape and workspace");
  if (e->S < 1) return fail(e, CDC_ERR_STATE, "cdc_set_schedule must be called first");
  if (e->cfg.variant == CDC_VARIANT_EPS && pred_mode != CDC_PRED_NOISE)
    return fail(e, CDC_ERR_UNSUPPORTED, "eps variant supports pred_mode 'noise' only");
  if (pred_mode < 0 || pred_mode > 2 || clip_mode < 0 || clip_mode > 2) return fail(e, CDC_ERR_INVALID, "bad pred/clip mode");
  int rc;
  Plan* pl = get_plan(e, B, H, W, workspace, &rc);
  if (!pl) return rc;
  if ((rc = check_ws(e, pl, workspace_bytes))) return rc;
  *out = pl;
  return 0;
}

int cdc_ddim_step(cdc_engine* e, float* x_inout, int i, const float* z, int pred_mode, int clip_mode, int B, int H,
                  int W, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!e) return CDC_ERR_INVALID;
  if (!x_inout) return fail(e, CDC_ERR_INVALID, "null x");
  DeviceGuard dg(e->device);
  Plan* pl;
  int rc = sampler_prologue(e, B, H, W, workspace, workspace_bytes, pred_mode, clip_mode, &pl);
  if (rc) return rc;
  if (i < 0 || i >= e->S) return fail(e, CDC_ERR_INVALID, "step index %d outside schedule of %d", i, e->S);
  cudaStream_t st = (cudaStream_t)stream;
  set_int_kernel<<<1, 32, 0, st>>>(e->d_step, i);
  RunArgs a;
  a.x = x_inout; a.x_inout = x_inout; a.z = z; a.mode = 1; a.pred = pred_mode; a.clip = clip_mode;
  return run_plan(e, pl, a, st);
}

static int sample_loop_impl(cdc_engine* e, float* x_inout, int i_first, int i_last, const float* z, int pred_mode,
                            int clip_mode, int B, int H, int W, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!e) return CDC_ERR_INVALID;
  if (!x_inout) return fail(e, CDC_ERR_INVALID, "null x");
  DeviceGuard dg(e->device);
  Plan* pl;
  int rc = sampler_prologue(e, B, H, W, workspace, workspace_bytes, pred_mode, clip_mode, &pl);
  if (rc) return rc;
  if (i_first >= e->S || i_last < 0 || i_last > i_first) return fail(e, CDC_ERR_INVALID, "bad step range %d..%d (S=%d)", i_first, i_last, e->S);
  cudaStream_t caller = (cudaStream_t)stream;
  cudaStream_t st = e->loop_stream;
  CUDA_TRY(e, cudaEventRecord(e->ev_in, caller));
  CUDA_TRY(e, cudaStreamWaitEvent(st, e->ev_in, 0));
  // the loop state lives in the workspace so the captured graph does not depend on the caller's buffer
  float* xs = reinterpret_cast<float*>(pl->ws + pl->xstate_off);
  const size_t xbytes = (size_t)B * e->cfg.channels * H * W * 4;
  auto body = [&]() -> int {
    int i_start = i_first;
    CUDA_TRY(e, cudaMemcpyAsync(xs, x_inout, xbytes, cudaMemcpyDeviceToDevice, st));
    if (z) set_int_kernel<<<1, 32, 0, st>>>(e->d_zfirst, i_first);
    if (!pl->graph || pl->graph_pred != pred_mode || pl->graph_clip != clip_mode || pl->graph_z != z) {
      if (pl->graph) { cudaGraphExecDestroy(pl->graph); pl->graph = nullptr; }
      RunArgs a;
      a.x = xs; a.x_inout = xs; a.z = z; a.z_chunk = z != nullptr; a.mode = 1; a.pred = pred_mode; a.clip = clip_mode;
      a.advance = true;
      // The first step runs eagerly: it is real work AND it forces every kernel's module load /
      // attribute setup to happen before stream capture starts.
      set_int_kernel<<<1, 32, 0, st>>>(e->d_step, i_first);
      int rc2 = run_plan(e, pl, a, st);
      if (rc2) return rc2;
      i_start = i_first - 1;
      cudaGraph_t g = nullptr;
      CUDA_TRY(e, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      rc2 = run_plan(e, pl, a, st);
      cudaError_t err = cudaStreamEndCapture(st, &g);
      if (rc2) { if (g) cudaGraphDestroy(g); return rc2; }
      if (err != cudaSuccess) return fail(e, CDC_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(err));
      err = cudaGraphInstantiate(&pl->graph, g, 0);
      cudaGraphDestroy(g);
      if (err != cudaSuccess) { pl->graph = nullptr; return fail(e, CDC_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(err)); }
      pl->graph_pred = pred_mode; pl->graph_clip = clip_mode; pl->graph_z = z;
    } else {
      set_int_kernel<<<1, 32, 0, st>>>(e->d_step, i_first);
    }
    for (int i = i_start; i >= i_last; --i) CUDA_TRY(e, cudaGraphLaunch(pl->graph, st));
    CUDA_TRY(e, cudaMemcpyAsync(x_inout, xs, xbytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  rc = body();
  // success or not, the caller's stream is re-joined with everything that was enqueued on the loop stream
  cudaError_t j1 = cudaEventRecord(e->ev_out, st);
  cudaError_t j2 = j1 == cudaSuccess ? cudaStreamWaitEvent(caller, e->ev_out, 0) : j1;
  if (rc) return rc;
  if (j2 != cudaSuccess) return fail(e, CDC_ERR_CUDA, "joining the loop stream failed: %s", cudaGetErrorString(j2));
  e->last_plan = pl;
  return CDC_OK;
}

int cdc_sample_loop(cdc_engine* e, float* x_inout, int i_first, int i_last, int pred_mode, int clip_mode, int B, int H,
                    int W, void* workspace, int64_t workspace_bytes, void* stream) {
  return sample_loop_impl(e, x_inout, i_first, i_last, nullptr, pred_mode, clip_mode, B, H, W, workspace, workspace_bytes, stream);
}

int cdc_sample_loop_noise(cdc_engine* e, float* x_inout, int i_first, int i_last, const float* z, int pred_mode,
                          int clip_mode, int B, int H, int W, void* workspace, int64_t workspace_bytes, void* stream) {
  if (e && !z) return fail(e, CDC_ERR_INVALID, "null noise buffer");
  return sample_loop_impl(e, x_inout, i_first, i_last, z, pred_mode, clip_mode, B, H, W, workspace, workspace_bytes, stream);
}

int cdc_engine_launches_per_forward(cdc_engine* e, int B, int H, int W) {
  if (!e) return CDC_ERR_INVALID;
  if (!e->finalized) return fail(e, CDC_ERR_STATE, "engine not finalized");
  Plan pl;
  int rc = build_plan(e, &pl, B, H, W, nullptr);
  if (rc) return rc;
  return (int)pl.ops.size();
}
int cdc_engine_launches_per_step(cdc_engine* e, int B, int H, int W) {
  int n = cdc_engine_launches_per_forward(e, B, H, W);
  return n < 0 ? n : n + 1;  // + advance_step
}
double cdc_engine_flops_per_forward(cdc_engine* e, int B, int H, int W) {
  if (!e || !e->finalized) return -1.0;
  Plan pl;
  if (build_plan(e, &pl, B, H, W, nullptr)) return -1.0;
  return pl.flops;
}
int cdc_engine_num_ops(cdc_engine* e, int B, int H, int W) { return cdc_engine_launches_per_forward(e, B, H, W); }

const char* cdc_engine_op_name(cdc_engine* e, int op_index) {
  if (!e || !e->last_plan || op_index < 0 || op_index >= (int)e->last_plan->ops.size()) return "";
  return e->last_plan->ops[op_index].name.c_str();
}

int64_t cdc_engine_debug_read(cdc_engine* e, int op_index, float* host_out, int64_t capacity, int* C, int* H, int* W) {
  if (!e || !e->last_plan) return fail(e, CDC_ERR_STATE, "no forward has run");
  Plan* pl = e->last_plan;
  if (op_index < 0 || op_index >= (int)pl->ops.size()) return fail(e, CDC_ERR_INVALID, "op index out of range");
  const Op& op = pl->ops[op_index];
  if (!op.dbg) return 0;
  const int64_t n = (int64_t)pl->B * op.dH * op.dW * op.dC;
  if (C) *C = op.dC;
  if (H) *H = op.dH;
  if (W) *W = op.dW;
  if (!host_out) return n;
  if (capacity < n) return fail(e, CDC_ERR_INVALID, "debug buffer too small");
  DeviceGuard dg(e->device);
  std::vector<__half> tmp((size_t)n);
  CUDA_TRY(e, cudaDeviceSynchronize());
  CUDA_TRY(e, cudaMemcpy(tmp.data(), op.dbg, (size_t)n * 2, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) host_out[i] = __half2float(tmp[(size_t)i]);
  return n;
}

int cdc_engine_profile_ops(cdc_engine* e, int iters, float* ms_out, double* flops_out, int capacity, void* stream) {
  if (!e || !e->last_plan) return fail(e, CDC_ERR_STATE, "no plan has run yet");
  if (iters < 1 || !ms_out || !flops_out) return fail(e, CDC_ERR_INVALID, "bad profile arguments");
  Plan* pl = e->last_plan;
  const int n = (int)pl->ops.size();
  if (capacity < n) return fail(e, CDC_ERR_INVALID, "profile buffers too small (%d ops)", n);
  DeviceGuard dg(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t a, b;
  CUDA_TRY(e, cudaEventCreate(&a));
  CUDA_TRY(e, cudaEventCreate(&b));
  RunArgs args = e->last_args;
  args.advance = false;
  long long* clkbuf = nullptr;
  if (getenv("CDC_DBG_CLK")) cudaMallocManaged(&clkbuf, 16 * sizeof(long long));
  struct ProfilingScope {
    cdc_engine* e;
    explicit ProfilingScope(cdc_engine* e_) : e(e_) { e->profiling = true; }
    ~ProfilingScope() { e->profiling = false; }
  } profiling_scope(e);
  for (int i = 0; i < n; ++i) {
    if (clkbuf) {
      cudaDeviceSynchronize();
      for (int k = 0; k < 16; ++k) clkbuf[k] = 0;
      pl->ops[i].tcp.dbg_clk = clkbuf;
    }
    int rc = run_op(e, pl, pl->ops[i], (size_t)i, args, st);  // warm
    if (rc) return rc;
    CUDA_TRY(e, cudaEventRecord(a, st));
    for (int k = 0; k < iters; ++k)
      if ((rc = run_op(e, pl, pl->ops[i], (size_t)i, args, st))) return rc;
    CUDA_TRY(e, cudaEventRecord(b, st));
    CUDA_TRY(e, cudaEventSynchronize(b));
    float ms = 0.f;
    CUDA_TRY(e, cudaEventElapsedTime(&ms, a, b));
    ms_out[i] = ms / iters;
    flops_out[i] = pl->ops[i].flops;
    if (clkbuf) {
      cudaDeviceSynchronize();
      pl->ops[i].tcp.dbg_clk = nullptr;
      if (pl->ops[i].use_tc && pl->ops[i].kind == OP_CONV) {
        const double d = iters + 1;
        fprintf(stderr, "CLK %-28s us=%7.1f prod: tot=%8.0f wait=%8.0f n=%5.1f pro=%6.0f | mma: tot=%8.0f wfull=%8.0f wtempty=%8.0f tiles=%5.1f | epi: tot=%8.0f wtfull=%8.0f pre=%8.0f p12=%8.0f p3=%8.0f\n",
                pl->ops[i].name.c_str(), ms_out[i] * 1000, clkbuf[0] / d, clkbuf[1] / d, clkbuf[2] / d, clkbuf[12] / d,
                clkbuf[3] / d, clkbuf[4] / d, clkbuf[5] / d, clkbuf[6] / d, clkbuf[7] / d, clkbuf[8] / d, clkbuf[9] / d, clkbuf[10] / d, clkbuf[11] / d);
      }
    }
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return n;
}

int cdc_engine_set_debug(cdc_engine* e, int no_reuse) {
  if (!e) return CDC_ERR_INVALID;
  e->debug_no_reuse = no_reuse != 0;
  for (auto& kv : e->plans) if (kv.second->graph) cudaGraphExecDestroy(kv.second->graph);
  e->plans.clear();
  e->last_plan = nullptr;
  e->ctx_set = false;
  return CDC_OK;
}

int cdc_engine_set_mainloop(cdc_engine* e, int kind) {
  if (!e) return CDC_ERR_INVALID;
  if (kind != 0 && kind != 1) return fail(e, CDC_ERR_UNSUPPORTED, "mainloop %d not built", kind);
  e->mainloop = kind;
  for (auto& kv : e->plans) if (kv.second->graph) cudaGraphExecDestroy(kv.second->graph);
  e->plans.clear();
  e->last_plan = nullptr;
  e->ctx_set = false;
  return CDC_OK;
}

int cdc_engine_tc_ops(cdc_engine* e, int B, int H, int W) {
  if (!e || !e->finalized) return fail(e, CDC_ERR_STATE, "engine not finalized");
  Plan pl;
  int rc = build_plan(e, &pl, B, H, W, nullptr);
  if (rc) return rc;
  int n = 0;
  for (auto& op : pl.ops) n += op.use_tc ? 1 : 0;
  return n;
}

}  // extern "C"
