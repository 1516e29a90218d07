// Batch augmentation + dB normalisation of patch / feature batches, fused into ONE streaming pass.
// Replaces the torch op chains of /root/reference/ViT_engine.py:28-117 (time_shift, add_noise, frequency_mask,
// time_mask as composed by augment_batch, then db_normalize) and the (x+120)/120 clip of
// "tablature-generator (1).py":334-335.  Each reference op is a full read+write of the (B, C, H, W) batch (cat /
// zeros_like / randn_like / clamp temporaries); composed here, a batch is read once and written once (HBM-bound).
//
// The reference applies 1-3 DIFFERENT ops in a random order, batch-wide (one shift / one mask for the whole batch).
// The composition is evaluated per output element by walking the op list backwards: masks and an out-of-range shift
// terminate the walk with zero, a shift moves the row index, noise adds a sample keyed by the coordinate the value
// had when the noise op ran.  Noise is Philox4x32-10 + Box-Muller keyed by (seed, element index) -- statistically
// equivalent to torch.randn_like, not bit-identical to it (torch's generator stream is not reproducible outside torch).
#include "gtc_common.cuh"

namespace gtc {

struct AugParams {
  int n_ops;
  unsigned ops;     // GTC_AUG_* codes in application order, 4 bits each (a packed word: no dynamic indexing of params)
  int shift;        // time_shift: out[h] = in[h + shift] (zero fill), along dim 2
  int f0, fw;       // frequency_mask: [:, :, :, f0:f0+fw] = 0   (dim 3)
  int t0, tw;       // time_mask:      [:, :, t0:t0+tw, :] = 0   (dim 2)
  float noise_level;
  unsigned long long seed;
  int normalize;    // db_normalize after the ops
  float ref_db;
};

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

// four N(0,1) samples for counter `ctr`
__device__ __forceinline__ void normal4(unsigned long long seed, unsigned long long ctr, float (&z)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x9E3779B9u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float s = 2.3283064365386963e-10f;          // 2^-32
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float u1 = ((float)c[2 * j] + 1.f) * s;    // (0, 1]
    const float u2 = (float)c[2 * j + 1] * s;
    const float r = sqrtf(-2.f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    z[2 * j] = r * cs;
    z[2 * j + 1] = r * sn;
  }
}

// Where the 4 consecutive columns w0..w0+3 of output row (bc, h) come from: walks the op list backwards.
// Returns the source float4 index (or -1 when the whole quad is zero-filled), the per-column live mask and the noise sum.
struct QuadPlan {
  long long src;      // float4 index into the input, -1 = no load
  unsigned live;      // bit j: column j still carries the input value
  float add[4];
};

__device__ __forceinline__ QuadPlan aug_plan(const AugParams& a, long long bc, int h, int w0, int H, int W) {
  QuadPlan q;
  q.live = 0xfu;
  q.add[0] = q.add[1] = q.add[2] = q.add[3] = 0.f;
  int hc = h;
  bool row_dead = false;
  for (int k = a.n_ops - 1; k >= 0 && !row_dead; --k) {
    switch ((a.ops >> (4 * k)) & 15u) {
      case GTC_AUG_TIME_SHIFT:
        hc += a.shift;
        if (hc < 0 || hc >= H) row_dead = true;
        break;
      case GTC_AUG_TIME_MASK:
        if (hc >= a.t0 && hc < a.t0 + a.tw) row_dead = true;
        break;
      case GTC_AUG_FREQ_MASK:
#pragma unroll
        for (int j = 0; j < 4; ++j) if (w0 + j >= a.f0 && w0 + j < a.f0 + a.fw) q.live &= ~(1u << j);
        break;
      case GTC_AUG_NOISE: {
        float z[4];
        normal4(a.seed, (unsigned long long)((bc * H + hc) * W + w0) >> 2, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (q.live & (1u << j)) q.add[j] += z[j] * a.noise_level;
        break;
      }
      default: break;
    }
  }
  q.src = (row_dead || q.live == 0u) ? -1 : ((bc * H + hc) * W + w0) >> 2;
  if (row_dead) q.live = 0u;
  return q;
}

__device__ __forceinline__ float4 aug_finish(const AugParams& a, const QuadPlan& q, float4 x) {
  float o[4] = {(q.live & 1u) ? x.x : 0.f, (q.live & 2u) ? x.y : 0.f, (q.live & 4u) ? x.z : 0.f, (q.live & 8u) ? x.w : 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[j] += q.add[j];
    if (a.normalize) o[j] = fminf(fmaxf((o[j] - a.ref_db) / (0.f - a.ref_db), 0.f), 1.f);
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}

// Idx = unsigned (fast 32-bit div/mod) whenever the batch has < 2^31 quads.  Each thread owns kAugUnroll quads per
// step, all loads issued before the first store (enough bytes in flight per SM to cover HBM latency).
constexpr int kAugUnroll = 4;
template <typename Idx>
__global__ void __launch_bounds__(256)
augment_kernel(const float* __restrict__ in, float* __restrict__ out, long long BC, int H, int W, const AugParams a) {
  const Idx wq = (Idx)(W >> 2);
  const Idx total = (Idx)(BC * H) * wq;
  const Idx stride = (Idx)gridDim.x * blockDim.x;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (Idx i0 = (Idx)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * kAugUnroll) {
    QuadPlan q[kAugUnroll];
    float4 x[kAugUnroll];
#pragma unroll
    for (int u = 0; u < kAugUnroll; ++u) {
      const Idx i = i0 + (Idx)u * stride;
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      q[u].src = -1;
      if (i < total && i >= i0) {                       // i >= i0: no wrap-around of the 32-bit index
        const Idx row = i / wq;
        const int w0 = (int)(i - row * wq) << 2;
        const Idx bc = row / (Idx)H;
        const int h = (int)(row - bc * (Idx)H);
        q[u] = aug_plan(a, (long long)bc, h, w0, H, W);
        if (q[u].src >= 0) x[u] = __ldcs(in4 + q[u].src);
      }
    }
#pragma unroll
    for (int u = 0; u < kAugUnroll; ++u) {
      const Idx i = i0 + (Idx)u * stride;
      if (i < total && i >= i0) __stcs(out4 + i, aug_finish(a, q[u], x[u]));
    }
  }
}

__global__ void __launch_bounds__(256)
db_normalize_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float ref_db) {
  const int64_t n4 = n >> 2;
  const float inv = 0.f - ref_db;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = __ldcs(reinterpret_cast<const float4*>(in) + i);
    x.x = fminf(fmaxf((x.x - ref_db) / inv, 0.f), 1.f);
    x.y = fminf(fmaxf((x.y - ref_db) / inv, 0.f), 1.f);
    x.z = fminf(fmaxf((x.z - ref_db) / inv, 0.f), 1.f);
    x.w = fminf(fmaxf((x.w - ref_db) / inv, 0.f), 1.f);
    __stcs(reinterpret_cast<float4*>(out) + i, x);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = fminf(fmaxf((in[i] - ref_db) / inv, 0.f), 1.f);
  }
}

}  // namespace gtc

using namespace gtc;

extern "C" int gtc_db_normalize(const float* d_in, int64_t n, float ref_db, float* d_out, gtc_stream_t stream) {
  GTC_REQUIRE(n >= 0, GTC_E_ARG, "gtc_db_normalize: negative n");
  if (n == 0) return GTC_OK;
  GTC_REQUIRE(d_in && d_out, GTC_E_ARG, "gtc_db_normalize: null pointer");
  GTC_REQUIRE(ref_db < 0.f, GTC_E_ARG, "gtc_db_normalize: ref_db must be negative");
  GTC_REQUIRE(((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0, GTC_E_ARG,
              "gtc_db_normalize: buffers must be 16-byte aligned");
  const int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  int64_t blocks = ceil_div(n / 4 + 1, 256);
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  db_normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n, ref_db);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

extern "C" int gtc_augment_batch(const float* d_in, float* d_out, int64_t batch, int channels, int dim2, int dim3,
                                 const int* h_ops, int n_ops, int shift, int freq0, int freq_width, int time0,
                                 int time_width, float noise_level, uint64_t noise_seed, int normalize, float ref_db,
                                 gtc_stream_t stream) {
  GTC_REQUIRE(batch >= 0 && channels > 0 && dim2 > 0 && dim3 > 0, GTC_E_ARG, "gtc_augment_batch: bad shape");
  if (batch == 0) return GTC_OK;
  GTC_REQUIRE(d_in && d_out, GTC_E_ARG, "gtc_augment_batch: null pointer");
  GTC_REQUIRE(n_ops >= 0 && n_ops <= 4 && (n_ops == 0 || h_ops), GTC_E_ARG, "gtc_augment_batch: 0..4 ops");
  GTC_REQUIRE(dim3 % 4 == 0, GTC_E_UNSUP, "gtc_augment_batch: last dimension must be a multiple of 4 (got %d)", dim3);
  GTC_REQUIRE(((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0, GTC_E_ARG,
              "gtc_augment_batch: buffers must be 16-byte aligned");
  GTC_REQUIRE(!normalize || ref_db < 0.f, GTC_E_ARG, "gtc_augment_batch: ref_db must be negative");
  AugParams a;
  memset(&a, 0, sizeof(a));
  a.n_ops = n_ops;
  bool moves = false;
  for (int k = 0; k < n_ops; ++k) {
    GTC_REQUIRE(h_ops[k] >= GTC_AUG_TIME_SHIFT && h_ops[k] <= GTC_AUG_TIME_MASK, GTC_E_ARG, "gtc_augment_batch: unknown op %d", h_ops[k]);
    for (int j = 0; j < k; ++j) GTC_REQUIRE(h_ops[j] != h_ops[k], GTC_E_ARG, "gtc_augment_batch: op %d listed twice", h_ops[k]);
    a.ops |= (unsigned)h_ops[k] << (4 * k);
    if (h_ops[k] == GTC_AUG_TIME_SHIFT && shift != 0) moves = true;
  }
  GTC_REQUIRE(!(moves && d_in == d_out), GTC_E_ARG, "gtc_augment_batch: a time shift cannot run in place");
  GTC_REQUIRE(freq_width >= 0 && time_width >= 0, GTC_E_ARG, "gtc_augment_batch: negative mask width");
  a.shift = shift; a.f0 = freq0; a.fw = freq_width; a.t0 = time0; a.tw = time_width;
  a.noise_level = noise_level; a.seed = noise_seed; a.normalize = normalize ? 1 : 0; a.ref_db = ref_db;
  const int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  const int64_t total = batch * channels * dim2 * (dim3 / 4);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks > 1) blocks = ceil_div(ceil_div(total, kAugUnroll), 256) < blocks ? ceil_div(ceil_div(total, kAugUnroll), 256) : blocks;
  if (total < 0x7fffffffLL - (int64_t)kAugUnroll * blocks * 256)
    augment_kernel<unsigned><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, batch * channels, dim2, dim3, a);
  else
    augment_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, batch * channels, dim2, dim3, a);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}
