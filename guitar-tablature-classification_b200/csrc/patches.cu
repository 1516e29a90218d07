// Patch assembly: dB features [n, n_bins, T] -> fp32 patch tensors [n, 3, OH, OW].
// Replaces GuitarTabDataset.__getitem__ + collate of /root/reference/ViT_dataloader.py:27-51 (bicubic, A=-0.75,
// align_corners=False, 3 identical channels) and the tensor contract of my_dataloader.py:17-21 (bilinear + ImageNet).
//
// This kernel is the binding roofline of the whole path: it reads 1 920 B and writes 602 112 B per segment, so it is a
// pure HBM store stream.  Design: one CTA works on one segment at a time (grid = SMs x resident CTAs, grid-stride).
//   1. the (n_bins x T) source is normalised into shared memory (next segment prefetched into registers meanwhile),
//   2. vertical pass n_bins -> OH on the T-wide source (OH x T values, tiny),
//   3. horizontal pass T -> OW fused with the store: every thread owns one float4 column group of the output row and
//      keeps the 4 x T combined interpolation coefficients in registers, so a row costs 2 LDS.128 + 4T FFMA and three
//      16-byte streaming stores (one per channel); a warp writes 512 contiguous bytes per store instruction.
#include "gtc_common.cuh"

namespace gtc {

struct PatchParams {
  const float* db;
  const int64_t* index;
  int64_t n;
  int h_in, t_in, oh, ow;
  int mode;
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

__device__ __forceinline__ float normalise_db(float x) {           // ViT_dataloader.py:31-32
  return fminf(fmaxf((x + 120.f) / 120.f, 0.f), 1.f);
}

constexpr int kPatchThreads = 224;
constexpr int kMaxSrc = 4096;      // n_bins * T floats staged per segment

// T_IN > 0: combined-coefficient fast path (OW % 4 == 0).  T_IN == 0: generic gather path, any size.
template <int T_IN>
__global__ void __launch_bounds__(kPatchThreads)
patch_kernel(const PatchParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int t_in = T_IN > 0 ? T_IN : p.t_in;
  const int tp = (t_in + 3) & ~3;                     // padded row of the vertically interpolated image
  const int h_in = p.h_in, oh = p.oh, ow = p.ow;
  const int n_src = h_in * t_in;

  float* s_src = smem;                                // [h_in][t_in]
  float* s_v = s_src + ((n_src + 3) & ~3);            // [oh][tp]
  float* s_wy = s_v + oh * tp;                        // [oh][4]
  int* s_iy = reinterpret_cast<int*>(s_wy + oh * 4);  // [oh][4]

  for (int y = tid; y < oh; y += blockDim.x) {
    int idx[4]; float w[4];
    axis_taps(y, h_in, oh, p.mode, idx, w);
#pragma unroll
    for (int i = 0; i < 4; ++i) { s_wy[y * 4 + i] = w[i]; s_iy[y * 4 + i] = idx[i]; }
  }

  const bool flip = p.mode == GTC_PATCH_CNN;          // picture orientation: highest bin on the top row
  float ch_scale[3] = {1.f, 1.f, 1.f}, ch_bias[3] = {0.f, 0.f, 0.f};
  if (p.mode == GTC_PATCH_CNN) {                      // my_dataloader.py:20  (x - mean) / std
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) { ch_scale[c] = 1.f / stdv[c]; ch_bias[c] = -mean[c] / stdv[c]; }
  }

  // per-thread horizontal coefficients
  constexpr int TC = T_IN > 0 ? T_IN : 1;
  float coef[4][TC];
  int gidx[4][4]; float gw[4][4];                     // generic path
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
  (void)gidx; (void)gw;

  constexpr int kPre = (kMaxSrc + kPatchThreads - 1) / kPatchThreads;   // prefetch registers (upper bound)
  const int n_pre = (n_src + blockDim.x - 1) / blockDim.x;
  float pre[T_IN > 0 ? ((96 * 16) / kPatchThreads + 1) : 1];
  constexpr int kPreFast = (96 * 16) / kPatchThreads + 1;
  (void)kPre;

  int64_t seg = blockIdx.x;
  const bool use_pre = T_IN > 0 && n_pre <= kPreFast;
  if (use_pre && seg < p.n) {
    const int64_t s = p.index ? p.index[seg] : seg;
#pragma unroll
    for (int k = 0; k < kPreFast; ++k) {
      const int o = tid + k * blockDim.x;
      pre[k] = (k < n_pre && o < n_src) ? __ldg(p.db + s * n_src + o) : 0.f;
    }
  }

  for (; seg < p.n; seg += gridDim.x) {
    __syncthreads();                                   // previous segment's readers of s_src / s_v are done
    if (use_pre) {
#pragma unroll
      for (int k = 0; k < kPreFast; ++k) {
        const int o = tid + k * blockDim.x;
        if (k < n_pre && o < n_src) {
          const int r = o / t_in, c = o - r * t_in;
          s_src[(flip ? (h_in - 1 - r) : r) * t_in + c] = normalise_db(pre[k]);
        }
      }
    } else {
      const int64_t s = p.index ? p.index[seg] : seg;
      for (int o = tid; o < n_src; o += blockDim.x) {
        const int r = o / t_in, c = o - r * t_in;
        s_src[(flip ? (h_in - 1 - r) : r) * t_in + c] = normalise_db(__ldg(p.db + s * n_src + o));
      }
    }
    __syncthreads();
    // vertical pass
    for (int o = tid; o < oh * t_in; o += blockDim.x) {
      const int y = o / t_in, c = o - y * t_in;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) a += s_wy[y * 4 + i] * s_src[s_iy[y * 4 + i] * t_in + c];
      s_v[y * tp + c] = a;
    }
    // prefetch the next segment's source while this one is being written out
    if (use_pre) {
      const int64_t nseg = seg + gridDim.x;
      if (nseg < p.n) {
        const int64_t s = p.index ? p.index[nseg] : nseg;
#pragma unroll
        for (int k = 0; k < kPreFast; ++k) {
          const int o = tid + k * blockDim.x;
          pre[k] = (k < n_pre && o < n_src) ? __ldg(p.db + s * n_src + o) : 0.f;
        }
      }
    }
    __syncthreads();

    float* obase = p.out + seg * 3 * (int64_t)oh * ow;
    const int64_t plane = (int64_t)oh * ow;
    if (T_IN > 0) {
      for (int y = r0; y < oh; y += rstep) {
        float v[TC > 4 ? ((TC + 3) & ~3) : 8];
        const float4* row = reinterpret_cast<const float4*>(s_v + y * tp);
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
      for (int64_t o = tid; o < plane; o += blockDim.x) {
        const int y = (int)(o / ow), x = (int)(o - (int64_t)y * ow);
        int idx[4]; float w[4];
        axis_taps(x, t_in, ow, p.mode, idx, w);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a += w[i] * s_v[y * tp + idx[i]];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) __stcs(obase + ch * plane + o, fmaf(a, ch_scale[ch], ch_bias[ch]));
      }
    }
  }
}

}  // namespace gtc

using namespace gtc;

extern "C" int gtc_patches(const float* d_db, const int64_t* d_index, int64_t n, int n_bins, int n_frames, int out_h,
                           int out_w, int mode, float* d_out, gtc_stream_t stream) {
  GTC_REQUIRE(n >= 0, GTC_E_ARG, "gtc_patches: negative n");
  if (n == 0) return GTC_OK;
  GTC_REQUIRE(d_db && d_out, GTC_E_ARG, "gtc_patches: null pointer");
  GTC_REQUIRE(mode == GTC_PATCH_VIT || mode == GTC_PATCH_CNN, GTC_E_ARG, "gtc_patches: unknown mode %d", mode);
  GTC_REQUIRE(n_bins > 0 && n_frames > 0 && out_h > 0 && out_w > 0, GTC_E_ARG, "gtc_patches: non-positive size");
  GTC_REQUIRE((int64_t)n_bins * n_frames <= kMaxSrc, GTC_E_UNSUP, "gtc_patches: n_bins*n_frames > %d", kMaxSrc);
  GTC_REQUIRE(out_h <= 2048 && out_w <= 4096, GTC_E_UNSUP, "gtc_patches: output larger than 2048x4096");
  PatchParams p{d_db, d_index, n, n_bins, n_frames, out_h, out_w, mode, d_out};
  const int tp = (n_frames + 3) & ~3;
  const size_t smem = sizeof(float) * (((size_t)n_bins * n_frames + 3) / 4 * 4 + (size_t)out_h * tp + (size_t)out_h * 8);
  GTC_REQUIRE(smem <= 200 * 1024, GTC_E_UNSUP, "gtc_patches: %zu bytes of shared memory needed", smem);
  int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  const int quads = out_w / 4;
  const bool fast = (out_w % 4 == 0) && quads <= kPatchThreads && (n_frames == 5 || n_frames == 9) &&
                    n_bins * n_frames <= 96 * 16;
  int threads = kPatchThreads;
  if (fast) threads = (kPatchThreads / quads) * quads;
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) GTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    GTC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > n) grid = n;
    kern<<<(unsigned)grid, threads, smem, st>>>(p);
    GTC_CUDA_CHECK(cudaGetLastError());
    return GTC_OK;
  };
  if (fast && n_frames == 5) return launch(patch_kernel<5>);
  if (fast && n_frames == 9) return launch(patch_kernel<9>);
  return launch(patch_kernel<0>);
}
