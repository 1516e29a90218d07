// Patch assembly: dB features [n, n_bins, T] -> fp32 patch tensors [n, 3, OH, OW].
// Replaces GuitarTabDataset.__getitem__ + collate of /root/reference/ViT_dataloader.py:27-51 (bicubic, A=-0.75,
// align_corners=False, 3 identical channels) and the tensor contract of my_dataloader.py:17-21 (bilinear + ImageNet).
//
// This kernel is the binding roofline of the whole path: it reads 1 920 B and writes 602 112 B per segment, so it is a
// pure HBM store stream.  Design:
//   * work item = (segment, row block): OH is cut into `parts` row blocks so that even a training-size batch gives
//     every SM several items and the tail of a launch is short; items are handed out by an atomic ticket, so CTAs that
//     become resident late (e.g. when the CQT GEMM of the next chunk shares the GPU) simply take what is left;
//   * per item the (n_bins x T) source is normalised into shared memory (the next item's source is prefetched into
//     registers meanwhile), the vertical pass n_bins -> rows is done on the T-wide source (tiny), and the horizontal
//     pass T -> OW is fused with the store: every thread owns one float4 column group and keeps its 4 x T combined
//     interpolation coefficients in registers, so a row costs 2 LDS.128 + 4T FFMA and three 16-byte streaming
//     stores (one per channel); a warp writes 512 contiguous bytes per store instruction.
#include <stdlib.h>
#include <mutex>
#include "gtc_common.cuh"

namespace gtc {

struct PatchParams {
  const float* db;
  const int64_t* index;
  int64_t n;
  int h_in, t_in, oh, ow;
  int mode;
  int prenorm;          // input is already (x+120)/120 clipped (tablature-generator (1).py:334-335): skip normalise_db
  int parts;            // row blocks per segment
  int rows_per_part;
  unsigned int* ticket; // zeroed before the launch
  float* out;
};

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// 4 clamped taps of output index `d` along an axis of n_in -> n_out samples (ATen upsample_{bicubic,bilinear}2d,
// align_corners=False).  Bilinear uses taps 0,1 and zero weights on 2,3.
__device__ __forceinline__ void axis_taps(int d, int n_in, int n_out, int mode, int idx[4], float w[4]) {
  const float scale = (float)n_in / (float)n_out;
  float src = scale * ((float)d + 0.5f) - 0.5f;
  if (mode == GTC_PATCH_VIT) {
    const float fl = floorf(src);
    const float t = src - fl;
    const int b = (int)fl;
    const float A = -0.75f;
    w[0] = cubic2(t + 1.f, A);
    w[1] = cubic1(t, A);
    w[2] = cubic1(1.f - t, A);
    w[3] = cubic2((1.f - t) + 1.f, A);
#pragma unroll
    for (int i = 0; i < 4; ++i) idx[i] = min(max(b - 1 + i, 0), n_in - 1);
  } else {
    src = fmaxf(src, 0.f);
    const int i0 = min((int)floorf(src), n_in - 1);
    const float l1 = src - (float)i0;
    idx[0] = i0; idx[1] = min(i0 + 1, n_in - 1); idx[2] = 0; idx[3] = 0;
    w[0] = 1.f - l1; w[1] = l1; w[2] = 0.f; w[3] = 0.f;
  }
}

__device__ __forceinline__ float normalise_db(float x, int prenorm = 0) {           // ViT_dataloader.py:31-32
  return prenorm ? x : fminf(fmaxf((x + 120.f) / 120.f, 0.f), 1.f);
}

constexpr int kPatchThreads = 224;
constexpr int kMaxSrc = 4096;                          // n_bins * T floats staged per segment
constexpr int kPreFast = (96 * 16) / kPatchThreads + 1;   // prefetch registers of the fast path

// T_IN > 0: combined-coefficient fast path (OW % 4 == 0).  T_IN == 0: generic gather path, any size.
template <int T_IN>
// (capping T_IN == 5 at 64 registers so that a CTA fits beside the 160-register GEMM CTA was measured: the spills cost
// 10 % of the store bandwidth -- 0.94 instead of 1.05 x the copy peak -- and the co-resident overlap did not pay for it)
__global__ void __launch_bounds__(kPatchThreads, T_IN == 5 ? 4 : 2)
patch_kernel(const PatchParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned int s_ticket;
  const int tid = threadIdx.x;
  const int t_in = T_IN > 0 ? T_IN : p.t_in;
  const int tp = (t_in + 3) & ~3;                     // padded row of the vertically interpolated image
  const int h_in = p.h_in, oh = p.oh, ow = p.ow;
  const int n_src = h_in * t_in;
  const int rpp = p.rows_per_part;
  const unsigned int n_items = (unsigned int)(p.n * p.parts);

  float* s_src = smem;                                // [h_in][t_in]
  float* s_v = s_src + ((n_src + 3) & ~3);            // [rows_per_part][tp]
  float* s_wy = s_v + rpp * tp;                       // [oh][4]
  int* s_iy = reinterpret_cast<int*>(s_wy + oh * 4);  // [oh][4]

  for (int y = tid; y < oh; y += blockDim.x) {
    int idx[4]; float w[4];
    axis_taps(y, h_in, oh, p.mode, idx, w);
#pragma unroll
    for (int i = 0; i < 4; ++i) { s_wy[y * 4 + i] = w[i]; s_iy[y * 4 + i] = idx[i]; }
  }
  if (tid == 0) s_ticket = atomicAdd(p.ticket, 1u);

  const bool flip = p.mode == GTC_PATCH_CNN;          // picture orientation: highest bin on the top row
  float ch_scale[3] = {1.f, 1.f, 1.f}, ch_bias[3] = {0.f, 0.f, 0.f};
  if (p.mode == GTC_PATCH_CNN) {                      // my_dataloader.py:20  (x - mean) / std
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) { ch_scale[c] = 1.f / stdv[c]; ch_bias[c] = -mean[c] / stdv[c]; }
  }

  // per-thread horizontal coefficients: out[x] = sum_c coef[x][c] * v[c]
  constexpr int TC = T_IN > 0 ? T_IN : 1;
  float coef[4][TC];
  const int quads = ow >> 2;
  int q = 0, r0 = 0, rstep = 1;
  if (T_IN > 0) {
    q = tid % quads; r0 = tid / quads; rstep = blockDim.x / quads;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int idx[4]; float w[4];
      axis_taps(q * 4 + j, t_in, ow, p.mode, idx, w);
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a += (idx[i] == c) ? w[i] : 0.f;
        coef[j][c] = a;
      }
    }
  }

  const int n_pre = (n_src + blockDim.x - 1) / blockDim.x;
  const bool use_pre = T_IN > 0 && n_pre <= kPreFast;
  float pre[kPreFast];
  auto prefetch = [&](unsigned int item) {
    const int64_t seg = item / p.parts;
    const int64_t s = p.index ? __ldg(p.index + seg) : seg;
#pragma unroll
    for (int k = 0; k < kPreFast; ++k) {
      const int o = tid + k * blockDim.x;
      pre[k] = (k < n_pre && o < n_src) ? __ldg(p.db + s * n_src + o) : 0.f;
    }
  };

  __syncthreads();
  unsigned int item = s_ticket;
  if (use_pre && item < n_items) prefetch(item);

  while (item < n_items) {
    const int64_t seg = item / p.parts;
    const int part = (int)(item - seg * p.parts);
    const int y0 = part * rpp, y1 = min(oh, y0 + rpp);
    __syncthreads();                                   // previous item's readers of s_src / s_v / s_ticket are done
    if (use_pre) {
#pragma unroll
      for (int k = 0; k < kPreFast; ++k) {
        const int o = tid + k * blockDim.x;
        if (k < n_pre && o < n_src) {
          const int r = o / t_in, c = o - r * t_in;
          s_src[(flip ? (h_in - 1 - r) : r) * t_in + c] = normalise_db(pre[k], p.prenorm);
        }
      }
    } else {
      const int64_t s = p.index ? p.index[seg] : seg;
      for (int o = tid; o < n_src; o += blockDim.x) {
        const int r = o / t_in, c = o - r * t_in;
        s_src[(flip ? (h_in - 1 - r) : r) * t_in + c] = normalise_db(__ldg(p.db + s * n_src + o), p.prenorm);
      }
    }
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const unsigned int next = s_ticket;
    // vertical pass for this row block
    for (int o = tid; o < (y1 - y0) * t_in; o += blockDim.x) {
      const int yy = o / t_in, c = o - yy * t_in;
      const int y = y0 + yy;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) a += s_wy[y * 4 + i] * s_src[s_iy[y * 4 + i] * t_in + c];
      s_v[yy * tp + c] = a;
    }
    if (use_pre && next < n_items) prefetch(next);     // next item's source travels while this one is written out
    __syncthreads();

    const int64_t plane = (int64_t)oh * ow;
    float* obase = p.out + seg * 3 * plane;
    if (T_IN > 0) {
      for (int y = y0 + r0; y < y1; y += rstep) {
        float v[(TC + 3) & ~3];
        const float4* row = reinterpret_cast<const float4*>(s_v + (y - y0) * tp);
#pragma unroll
        for (int k = 0; k < (TC + 3) / 4; ++k) {
          const float4 f = row[k];
          v[4 * k + 0] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
        }
        float o4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = 0.f;
#pragma unroll
          for (int c = 0; c < TC; ++c) a = fmaf(coef[j][c], v[c], a);
          o4[j] = a;
        }
        float* dst = obase + (int64_t)y * ow + q * 4;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float4 st;
          st.x = fmaf(o4[0], ch_scale[ch], ch_bias[ch]);
          st.y = fmaf(o4[1], ch_scale[ch], ch_bias[ch]);
          st.z = fmaf(o4[2], ch_scale[ch], ch_bias[ch]);
          st.w = fmaf(o4[3], ch_scale[ch], ch_bias[ch]);
          __stcs(reinterpret_cast<float4*>(dst + ch * plane), st);
        }
      }
    } else {
      const int n_out = (y1 - y0) * ow;
      for (int o = tid; o < n_out; o += blockDim.x) {
        const int yy = o / ow, x = o - yy * ow;
        int idx[4]; float w[4];
        axis_taps(x, t_in, ow, p.mode, idx, w);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a += w[i] * s_v[yy * tp + idx[i]];
        float* dst = obase + (int64_t)(y0 + yy) * ow + x;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) __stcs(dst + ch * plane, fmaf(a, ch_scale[ch], ch_bias[ch]));
      }
    }
    item = next;
  }
}

// ToTensor + Normalize of decoded, resized RGB pictures (my_dataloader.py:19-20), gathered by the sampler's index.
// One thread = four pixels of one row: 12 input bytes (HWC) -> one float4 per channel plane (CHW).  x / 255, - mean, / std
// are the three separately rounded fp32 operations torchvision performs (tensor.div(255); sub_(mean); div_(std)).
__global__ void __launch_bounds__(256)
rgb8_normalize_kernel(const uint8_t* __restrict__ rgb, const int64_t* __restrict__ index, int64_t n, int h, int w,
                      float3 mean, float3 stdv, float* __restrict__ out) {
  const int wq = w >> 2;
  const int64_t quads_per_item = (int64_t)h * wq;
  const int64_t total = n * quads_per_item;
  const int64_t plane = (int64_t)h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t item = i / quads_per_item;
    const int64_t q = i - item * quads_per_item;                       // quad within the picture (row-major)
    const int64_t src = index ? __ldg(index + item) : item;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(rgb + (src * plane + q * 4) * 3);
    const uint32_t w0 = __ldcs(p), w1 = __ldcs(p + 1), w2 = __ldcs(p + 2);     // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
    const uint8_t b[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                           (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                           (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
    const float m[3] = {mean.x, mean.y, mean.z}, s[3] = {stdv.x, stdv.y, stdv.z};
    float* dst = out + item * 3 * plane + q * 4;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)b[3 * j + c], 255.f), m[c]), s[c]);
      __stcs(reinterpret_cast<float4*>(dst + c * plane), make_float4(v[0], v[1], v[2], v[3]));
    }
  }
}

// tickets: a ring of counters owned by the library, one ring per device.  A launch takes the next slot, zeroes it on its
// stream and leaves an event behind; the launch that reuses the slot kTicketSlots launches later first makes its stream
// wait for that event, so a counter is never reset under a kernel that is still drawing from it -- whatever streams the
// two launches are on (a training DeviceLoader beside FrontEnd's 9 ms patch launches).
constexpr int kTicketSlots = 1024;
__device__ unsigned int g_patch_tickets[kTicketSlots];

struct TicketRing {
  std::mutex mu;
  unsigned int* base = nullptr;
  unsigned int next = 0;
  cudaEvent_t done[kTicketSlots] = {};
};
static TicketRing g_ticket_rings[64];

}  // namespace gtc

using namespace gtc;

extern "C" int gtc_patches(const float* d_db, const int64_t* d_index, int64_t n, int n_bins, int n_frames, int out_h,
                           int out_w, int mode, float* d_out, gtc_stream_t stream) {
  GTC_REQUIRE(n >= 0, GTC_E_ARG, "gtc_patches: negative n");
  if (n == 0) return GTC_OK;
  GTC_REQUIRE(d_db && d_out, GTC_E_ARG, "gtc_patches: null pointer");
  GTC_REQUIRE(mode == GTC_PATCH_VIT || mode == GTC_PATCH_CNN || mode == GTC_PATCH_VIT_PRENORM, GTC_E_ARG, "gtc_patches: unknown mode %d", mode);
  const int prenorm = mode == GTC_PATCH_VIT_PRENORM;
  if (prenorm) mode = GTC_PATCH_VIT;
  GTC_REQUIRE(n_bins > 0 && n_frames > 0 && out_h > 0 && out_w > 0, GTC_E_ARG, "gtc_patches: non-positive size");
  GTC_REQUIRE((int64_t)n_bins * n_frames <= kMaxSrc, GTC_E_UNSUP, "gtc_patches: n_bins*n_frames > %d", kMaxSrc);
  GTC_REQUIRE(out_h <= 2048 && out_w <= 4096, GTC_E_UNSUP, "gtc_patches: output larger than 2048x4096");
  GTC_REQUIRE(n < (int64_t)1 << 26, GTC_E_UNSUP, "gtc_patches: more than 2^26 items in one call");
  int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;

  const int quads = out_w / 4;
  // the fast path stores float4: a sliced / offset output that is not 16-byte aligned takes the scalar path
  const bool fast = (out_w % 4 == 0) && quads <= kPatchThreads && (n_frames == 5 || n_frames == 9) &&
                    n_bins * n_frames <= 96 * 16 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0;
  GTC_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_db) & 3) == 0, GTC_E_ARG,
              "gtc_patches: d_db and d_out must be 4-byte aligned");
  int threads = kPatchThreads;
  if (fast) threads = (kPatchThreads / quads) * quads;
  // row blocks: ~56 rows each (14 store iterations per thread), at least 1
  int parts = out_h >= 112 ? (out_h + 55) / 56 : 1;
  const int rpp = (out_h + parts - 1) / parts;
  parts = (out_h + rpp - 1) / rpp;

  int dev = 0;
  GTC_CUDA_CHECK(cudaGetDevice(&dev));
  GTC_REQUIRE(dev >= 0 && dev < 64, GTC_E_UNSUP, "gtc_patches: device ordinal %d out of range", dev);
  TicketRing& ring = g_ticket_rings[dev];
  std::lock_guard<std::mutex> ring_lock(ring.mu);      // held until the launch and its event are enqueued
  if (!ring.base) GTC_CUDA_CHECK(cudaGetSymbolAddress((void**)&ring.base, g_patch_tickets));
  const unsigned int slot = ring.next++ % kTicketSlots;
  if (ring.done[slot]) GTC_CUDA_CHECK(cudaStreamWaitEvent(st, ring.done[slot], 0));      // the slot's previous kernel has finished
  else GTC_CUDA_CHECK(cudaEventCreateWithFlags(&ring.done[slot], cudaEventDisableTiming));
  unsigned int* ticket = ring.base + slot;
  GTC_CUDA_CHECK(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));

  PatchParams p{d_db, d_index, n, n_bins, n_frames, out_h, out_w, mode, prenorm, parts, rpp, ticket, d_out};
  const int tp = (n_frames + 3) & ~3;
  const size_t smem = sizeof(float) * (((size_t)n_bins * n_frames + 3) / 4 * 4 + (size_t)rpp * tp + (size_t)out_h * 8);
  GTC_REQUIRE(smem <= 200 * 1024, GTC_E_UNSUP, "gtc_patches: %zu bytes of shared memory needed", smem);
  // Residency is pinned through the shared-memory request.  The CTAs are persistent (ticket loop), so where they land at
  // launch is where they stay: with the small natural footprint (11 KB, 72 registers) four fit on an SM, and when another
  // kernel still holds some SMs at launch time (the label / framing kernels of the next chunk, a consumer's kernels) the
  // 2 x 148 CTAs pile up 3-4 deep on the free SMs and leave the others empty for the whole launch -- measured on B200:
  // 4.5 instead of 7.2 TB/s, per launch and depending on the chunk size (profiles/r01j_patch_residency.md).  Asking for
  // 228 KB / k per CTA makes k CTAs fill an SM, so the grid of k x 148 can only be placed evenly.
  const int want_per_sm = patch_ctas_per_sm();
  size_t smem_req = smem;
  {
    const size_t even = (size_t)(228 * 1024) / (size_t)want_per_sm - 1024 - 256;    // 1 KB per CTA is reserved by the system
    if (even > smem_req) smem_req = even;
  }
  auto launch = [&](auto kern) -> int {
    static const char* pin = getenv("GTC_PATCH_NO_PIN");                // experiment switch: natural footprint
    const size_t sm_bytes = pin ? smem : smem_req;
    if (sm_bytes > 48 * 1024) GTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bytes));
    int per_sm = 1;
    GTC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, sm_bytes));
    if (per_sm < 1) per_sm = 1;
    // measured on B200 (scripts/patch_cta_sweep.py): 2 CTAs/SM stream 7.08 TB/s, 4 CTAs/SM 6.88 TB/s, 1 CTA/SM 6.11 TB/s
    if (per_sm > want_per_sm) per_sm = want_per_sm;
    int64_t grid = (int64_t)sms * per_sm;
    const int64_t items = n * parts;
    if (grid > items) grid = items;
    if (patch_max_ctas() > 0 && grid > patch_max_ctas()) grid = patch_max_ctas();
    kern<<<(unsigned)grid, threads, sm_bytes, st>>>(p);
    GTC_CUDA_CHECK(cudaGetLastError());
    GTC_CUDA_CHECK(cudaEventRecord(ring.done[slot], st));
    return GTC_OK;
  };
  if (fast && n_frames == 5) return launch(patch_kernel<5>);
  if (fast && n_frames == 9) return launch(patch_kernel<9>);
  return launch(patch_kernel<0>);
}

extern "C" int gtc_patches_rgb8(const uint8_t* d_rgb, const int64_t* d_index, int64_t n, int h, int w, float mean_r,
                                float mean_g, float mean_b, float std_r, float std_g, float std_b, float* d_out,
                                gtc_stream_t stream) {
  GTC_REQUIRE(n >= 0, GTC_E_ARG, "gtc_patches_rgb8: negative n");
  if (n == 0) return GTC_OK;
  GTC_REQUIRE(d_rgb && d_out, GTC_E_ARG, "gtc_patches_rgb8: null pointer");
  GTC_REQUIRE(h > 0 && w > 0 && w % 4 == 0, GTC_E_ARG, "gtc_patches_rgb8: w must be a positive multiple of 4 (got %d x %d)", h, w);
  GTC_REQUIRE((reinterpret_cast<uintptr_t>(d_rgb) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0, GTC_E_ARG,
              "gtc_patches_rgb8: d_rgb must be 4-byte and d_out 16-byte aligned");
  GTC_REQUIRE(std_r != 0.f && std_g != 0.f && std_b != 0.f, GTC_E_ARG, "gtc_patches_rgb8: zero std");
  int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  const int64_t total = n * h * (int64_t)(w / 4);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  rgb8_normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_rgb, d_index, n, h, w,
                                                                            make_float3(mean_r, mean_g, mean_b),
                                                                            make_float3(std_r, std_g, std_b), d_out);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}
