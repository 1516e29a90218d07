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

// The reference applies its ops in a drawn order; the host resolves that order ONCE into output-space facts, so the
// kernel is straight-line code per element:
//   value  = in[h + shift][w]        unless the shifted row leaves [0, H), or h is in the (output-space) time-mask rows
//                                    [t0, t1), or w is in the frequency-mask columns [f0, f1)
//   noise  = level * N(0,1) keyed by the coordinate the element had when add_noise ran (row h + noise_row_off); it
//            survives a mask / an out-of-range shift only if that op ran BEFORE add_noise
struct AugParams {
  int shift;            // 0 when no time_shift
  int t0, t1;           // output rows zeroed by time_mask (already moved by a later shift), empty when t1 <= t0
  int f0, f1;           // columns zeroed by frequency_mask
  int has_noise;
  int noise_row_off;    // shift if the shift runs after add_noise (noise travels with the rows), else 0
  int noise_dies_tmask, noise_dies_fmask, noise_dies_shift;   // the op runs after add_noise
  float noise_level;
  unsigned long long seed;
  int normalize;        // db_normalize after the ops
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

// One thread = one float4 (4 columns) of kAugRows rows spaced `row_step` apart; all loads are issued before the first
// store so every SM keeps enough bytes in flight to cover HBM latency.  blockDim = (quads per row | 256-cap, rows).
constexpr int kAugRows = 4;
__global__ void __launch_bounds__(256)
augment_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned n_rows, int H, int W, const AugParams a) {
  const int wq = W >> 2;
  const unsigned row_step = gridDim.x * blockDim.y;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (unsigned row0 = blockIdx.x * blockDim.y + threadIdx.y; row0 < n_rows; row0 += row_step * kAugRows) {
    for (int q = threadIdx.x; q < wq; q += blockDim.x) {
      const int w0 = q << 2;
      unsigned colmask = 0xfu;                                   // bit j: column w0+j keeps its value
#pragma unroll
      for (int j = 0; j < 4; ++j) if (w0 + j >= a.f0 && w0 + j < a.f1) colmask &= ~(1u << j);
      float4 x[kAugRows];
      bool keep[kAugRows], nz[kAugRows];
      unsigned bc[kAugRows];
      int h[kAugRows];
#pragma unroll
      for (int r = 0; r < kAugRows; ++r) {
        const unsigned row = row0 + r * row_step;
        x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        keep[r] = nz[r] = false;
        bc[r] = 0; h[r] = 0;
        if (row < n_rows) {
          bc[r] = row / (unsigned)H;
          h[r] = (int)(row - bc[r] * (unsigned)H);
          const int hs = h[r] + a.shift;
          const bool in_range = hs >= 0 && hs < H;
          const bool masked = h[r] >= a.t0 && h[r] < a.t1;
          keep[r] = in_range && !masked;
          nz[r] = a.has_noise && !(masked && a.noise_dies_tmask) && !(!in_range && a.noise_dies_shift);
          if (keep[r] && colmask) x[r] = __ldcs(in4 + ((size_t)bc[r] * H + hs) * wq + q);
        }
      }
#pragma unroll
      for (int r = 0; r < kAugRows; ++r) {
        const unsigned row = row0 + r * row_step;
        if (row >= n_rows) continue;
        float o[4] = {(colmask & 1u) ? x[r].x : 0.f, (colmask & 2u) ? x[r].y : 0.f, (colmask & 4u) ? x[r].z : 0.f, (colmask & 8u) ? x[r].w : 0.f};
        if (nz[r]) {
          float z[4];
          normal4(a.seed, ((unsigned long long)bc[r] * H + (h[r] + a.noise_row_off)) * wq + q, z);
          const unsigned nm = a.noise_dies_fmask ? colmask : 0xfu;
#pragma unroll
          for (int j = 0; j < 4; ++j) if (nm & (1u << j)) o[j] += z[j] * a.noise_level;
        }
        if (a.normalize) {
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fminf(fmaxf((o[j] - a.ref_db) / (0.f - a.ref_db), 0.f), 1.f);
        }
        __stcs(out4 + (size_t)row * wq + q, make_float4(o[0], o[1], o[2], o[3]));
      }
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
  GTC_REQUIRE(freq_width >= 0 && time_width >= 0, GTC_E_ARG, "gtc_augment_batch: negative mask width");
  GTC_REQUIRE(batch * channels * dim2 < 0xffffffffLL, GTC_E_UNSUP, "gtc_augment_batch: more than 2^32 rows; split the batch");
  // resolve the drawn order into output-space facts (see AugParams)
  int pos[5] = {-1, -1, -1, -1, -1};
  for (int k = 0; k < n_ops; ++k) {
    GTC_REQUIRE(h_ops[k] >= GTC_AUG_TIME_SHIFT && h_ops[k] <= GTC_AUG_TIME_MASK, GTC_E_ARG, "gtc_augment_batch: unknown op %d", h_ops[k]);
    GTC_REQUIRE(pos[h_ops[k]] < 0, GTC_E_ARG, "gtc_augment_batch: op %d listed twice", h_ops[k]);
    pos[h_ops[k]] = k;
  }
  AugParams a;
  memset(&a, 0, sizeof(a));
  const int p_shift = pos[GTC_AUG_TIME_SHIFT], p_noise = pos[GTC_AUG_NOISE], p_f = pos[GTC_AUG_FREQ_MASK], p_t = pos[GTC_AUG_TIME_MASK];
  a.shift = p_shift >= 0 ? shift : 0;
  GTC_REQUIRE(!(a.shift != 0 && d_in == d_out), GTC_E_ARG, "gtc_augment_batch: a time shift cannot run in place");
  if (p_t >= 0 && time_width > 0) {
    // a mask applied before the shift travels with the rows: source rows [t0, t0+tw) land on output rows [t0 - shift, ...)
    const int move = (p_shift > p_t) ? a.shift : 0;
    a.t0 = time0 - move;
    a.t1 = time0 + time_width - move;
  }
  if (p_f >= 0 && freq_width > 0) { a.f0 = freq0; a.f1 = freq0 + freq_width; }
  if (p_noise >= 0) {
    a.has_noise = 1;
    a.noise_row_off = (p_shift > p_noise) ? a.shift : 0;
    a.noise_dies_tmask = p_t > p_noise;
    a.noise_dies_fmask = p_f > p_noise;
    a.noise_dies_shift = p_shift > p_noise;
  }
  a.noise_level = noise_level; a.seed = noise_seed; a.normalize = normalize ? 1 : 0; a.ref_db = ref_db;
  const int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  const int wq = dim3 / 4;
  const unsigned n_rows = (unsigned)(batch * channels * dim2);
  ::dim3 block((unsigned)(wq < 256 ? wq : 256), 1, 1);
  block.y = 256 / block.x > 0 ? 256 / block.x : 1;
  int64_t blocks = ceil_div(ceil_div((int64_t)n_rows, kAugRows), block.y);
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  augment_kernel<<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(d_in, d_out, n_rows, dim2, dim3, a);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}
