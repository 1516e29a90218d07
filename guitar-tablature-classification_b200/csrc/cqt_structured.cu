// Structured (multirate) CQT of variable-length segments: librosa.cqt evaluated the way librosa itself does it --
// a chain of soxr-HQ 2:1 decimations and, per octave, the wavelet basis applied to centred, zero-padded frames --
// instead of the collapsed segment operator of cqt_api.cu.  It serves the recipes whose operator would not fit
// (/root/reference/tablature_generator.py:616-620: 3 s segments -> 21 840 x 66 150), whole-clip CQTs and arbitrary
// sample rates, and it is the on-device cross-check of the tensor-core operator path.
//
//   decimate2_kernel  : y_{i+1}[k] = sum_j h[j] * y_i[2k + c - j]      (taps in shared memory, even/odd input phases
//                       de-interleaved into two shared arrays so a warp reads unit-stride; zero-extended signal,
//                       group delay compensated, length ceil(n/2), sqrt(2) gain folded into h)
//                       == librosa.resample(orig_sr=2, target_sr=1, res_type='soxr_hq', scale=True)
//   response_kernel   : C[bin, t] = sum_n W_i[bin, n] * ypad_i[t*hop_i + n]   (the octave's time-domain filters
//                       W_i = fft_basis_i @ rfft-matrix, designed on the host; frames x basis contraction, fp32 FMA;
//                       32 frames x n_fft staged in shared memory per CTA, each thread owns 1 frame x 2 complex bins)
//                       == librosa.stft(window='ones', center=True, pad_mode='constant') followed by fft_basis.dot(D)
//   finish_kernel     : |C|^power -> amplitude_to_db(ref=max over the segment, amin, top_db) -> cut   (in place)
//
// Everything is HBM/L2-bound streaming work except the FIR (389 taps), which is shared-memory bound.
//
// Round 2: for dB features both contractions run on the tcgen05 engine of cqt_gemm_tc.cu (SlotArgs, gtc_common.cuh).  The
// signals of every octave live as fp16 hi/lo planes (x * 2^8 split into two halves, 22 mantissa bits), one zero-guarded
// slot per segment:
//   split_kernel        audio (fp32 / int16 PCM) -> octave-0 planes
//   decimator GEMM      rows = 128-output windows of 704 input samples (row step 256: overlapping TMA rows) x a fixed
//                       banded Toeplitz operator [128][704] of the taps, kept RESIDENT in shared memory as one master tile
//                       (k-block kb + 1 of a Toeplitz operator is k-block kb moved down 16 rows); only the rows that cover
//                       the longest segment's octave length are computed, the rest of a slot is zero-filled once per call;
//                       epilogue transposes 8 x 8 vectors by shuffles and writes the next octave's planes as whole lines
//   response GEMM       rows = frames (n_fft samples, row step hop_i) x the octave's filters [32][n_fft]; epilogue
//                       writes |C|^2 into [seg][bin][t] and the segment maximum
//   sfinish_kernel      in-place dB
// The fp32 SIMT kernels below stay for complex output, the stand-alone decimator API and recipes the tensor path does not
// take (more than 16 filters per octave, hops that leave the lowest octave's frames off 16-byte boundaries).
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>
#include "gtc_common.cuh"

namespace gtc {

constexpr int kMaxOctaves = 16;
constexpr int kDecTile = 1024;      // outputs per CTA of the decimator (256 threads x 4)
constexpr int kFramesPerCta = 32;

struct SPlanImpl {
  int device, sm_count;
  int use_tc;               // both contractions on the tensor cores (dB output)
  int dec_left, dec_k;      // Toeplitz window: starts dec_left samples before input 2k0, dec_k samples long
  void* dec_plan;           // gtc_plan* of the Toeplitz operator [128][dec_k]
  void* resp_plan[kMaxOctaves];   // gtc_plan* of octave i's filters [32][n_fft]
  int n_oct, n_fft, hop, n_bins, n_filters, n_taps;
  int groups;               // float4 groups of the 2*n_filters real outputs of one octave
  int bin_lo[kMaxOctaves];  // first output bin of octave i (octave 0 = top octave)
  int bin_cnt[kMaxOctaves];
  float* d_filters;         // [n_oct][n_fft][groups*4]: sample-major so a thread fetches its 4 outputs as one float4
  float* d_taps;            // [n_taps], sqrt(2) gain folded in
};

struct OctaveBufs {
  int64_t off[kMaxOctaves];     // float offset of octave i's signals inside the workspace (i >= 1)
  int64_t stride[kMaxOctaves];  // floats per segment
};

__device__ __forceinline__ float s_load(const float* p) { return __ldg(p); }
__device__ __forceinline__ float s_load(const int16_t* p) { return (float)__ldg(p) * (1.f / 32768.f); }

__device__ __forceinline__ int halved(int n, int times) { return halved_len(n, times); }
__device__ __host__ __forceinline__ int frames_of(int len, int hop, int n_oct) { return cqt_frames_of(len, hop, n_oct); }

// ---------------------------------------------------------------------------------------------------------------------
// 2:1 decimator.  blockIdx.x = segment * tiles + tile.
//
// With xe[m] = x[2 k0 - c + 2m], xo[m] = x[2 k0 - c + 2m + 1] (the two input phases, de-interleaved while staging) and
// the taps re-indexed he[j] = h[2(c-j)], ho[j] = h[2(c-1-j)+1], both phases are plain correlations
//     out[k0+u] = sum_j he[j] xe[u+j] + sum_j ho[j] xo[u+j].
// A thread owns 4 consecutive outputs and walks the taps 4 at a time with an 8-sample register window: per block ONE
// LDS.128 of samples (lanes 16 B apart: conflict-free) and ONE broadcast LDS.128 of taps feed 16 FFMA, so the loop is
// bound by FFMA issue, not by shared-memory bandwidth (the first version spent 9 shared wavefronts per 8 FFMA and sat
// at 99 % l1tex throughput, profiles/r01e).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void corr4(const float4* __restrict__ x4, const float4* __restrict__ t4, int n_blocks, float (&acc)[4]) {
  float4 w0 = x4[0];
#pragma unroll 2
  for (int jb = 0; jb < n_blocks; ++jb) {
    const float4 w1 = x4[jb + 1];
    const float4 t = t4[jb];
    acc[0] = fmaf(t.x, w0.x, acc[0]); acc[0] = fmaf(t.y, w0.y, acc[0]); acc[0] = fmaf(t.z, w0.z, acc[0]); acc[0] = fmaf(t.w, w0.w, acc[0]);
    acc[1] = fmaf(t.x, w0.y, acc[1]); acc[1] = fmaf(t.y, w0.z, acc[1]); acc[1] = fmaf(t.z, w0.w, acc[1]); acc[1] = fmaf(t.w, w1.x, acc[1]);
    acc[2] = fmaf(t.x, w0.z, acc[2]); acc[2] = fmaf(t.y, w0.w, acc[2]); acc[2] = fmaf(t.z, w1.x, acc[2]); acc[2] = fmaf(t.w, w1.y, acc[2]);
    acc[3] = fmaf(t.x, w0.w, acc[3]); acc[3] = fmaf(t.y, w1.x, acc[3]); acc[3] = fmaf(t.z, w1.y, acc[3]); acc[3] = fmaf(t.w, w1.z, acc[3]);
    w0 = w1;
  }
}

template <typename In>
__global__ void __launch_bounds__(256)
decimate2_kernel(const In* __restrict__ src, const int64_t* __restrict__ seg_start, const int32_t* __restrict__ seg_valid,
                 const int32_t* __restrict__ seg_len, int stage, int64_t src_stride, float* __restrict__ dst,
                 int64_t dst_stride, const float* __restrict__ taps, int n_taps, int tiles, float gain) {
  extern __shared__ __align__(16) float sm[];
  const int c = (n_taps - 1) >> 1;                 // even: n_taps == 1 (mod 4)
  const int nt = (c + 4) & ~3;                     // taps per phase, padded to a multiple of 4 (c+1 even-phase taps)
  const int span = kDecTile + nt + 4;              // staged samples per phase
  float* he = sm;                                  // [nt]
  float* ho = he + nt;                             // [nt]
  float* xe = ho + nt;                             // [span]
  float* xo = xe + span;                           // [span]
  const int64_t s = blockIdx.x / tiles;
  const int tile = blockIdx.x - (int)(s * tiles);
  const int len_in = halved(__ldg(seg_len + s), stage);
  const int len_out = (len_in + 1) >> 1;
  const int k0 = tile * kDecTile;
  if (k0 >= len_out) return;
  const int readable = stage == 0 ? min(len_in, __ldg(seg_valid + s)) : len_in;
  const In* x = src + (seg_start ? __ldg(seg_start + s) : s * src_stride);

  for (int j = threadIdx.x; j < nt; j += blockDim.x) {
    he[j] = j <= c ? __ldg(taps + 2 * (c - j)) : 0.f;
    ho[j] = j < c ? __ldg(taps + 2 * (c - 1 - j) + 1) : 0.f;
  }
  const int base = 2 * k0 - c;                     // input index of xe[0]
  const int n_here = min(kDecTile, len_out - k0);  // outputs this CTA really owes (short octaves fill a fraction of a tile)
  const int n_stage = ((n_here + 3) & ~3) + nt + 4;
  for (int m = threadIdx.x; m < 2 * n_stage; m += blockDim.x) {
    const int i = base + m;
    const float v = (i >= 0 && i < readable) ? s_load(x + i) : 0.f;
    if (m & 1) xo[m >> 1] = v; else xe[m >> 1] = v;
  }
  __syncthreads();
  const int u = threadIdx.x << 2;
  if (u >= n_here) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  corr4(reinterpret_cast<const float4*>(xe) + threadIdx.x, reinterpret_cast<const float4*>(he), nt >> 2, acc);
  corr4(reinterpret_cast<const float4*>(xo) + threadIdx.x, reinterpret_cast<const float4*>(ho), nt >> 2, acc);
  float* y = dst + s * dst_stride + k0 + u;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (u + r < n_here) y[r] = acc[r] * gain;
}

// ---------------------------------------------------------------------------------------------------------------------
// Octave response.  blockIdx.x covers 32 consecutive global frames q = segment * t_max + t.
// blockDim = 32 * groups: lane = frame slot, warp = output group (4 real outputs = 2 complex bins).
// ---------------------------------------------------------------------------------------------------------------------
template <typename In>
__global__ void __launch_bounds__(1024)
response_kernel(const In* __restrict__ src, const int64_t* __restrict__ seg_start, const int32_t* __restrict__ seg_valid,
                const int32_t* __restrict__ seg_len, int64_t n_seg, int octave, int n_oct, int hop0, int64_t src_stride,
                const float* __restrict__ filters, int n_fft, int groups, int bin_lo, int bin_cnt, int n_bins, int t_max,
                float* __restrict__ out, int complex_out, float* __restrict__ segmax) {
  extern __shared__ float sm[];
  const int fstride = n_fft + 1;
  float4* w4 = reinterpret_cast<float4*>(sm);                       // [n_fft][groups]
  float* fr = sm + (size_t)n_fft * groups * 4;                      // [32][n_fft + 1]
  int* meta = reinterpret_cast<int*>(fr + kFramesPerCta * fstride); // seg_lo, seg_hi, t, readable, valid, max  (x32)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int hop = hop0 >> octave;

  if (warp == 0) {
    const int64_t q = (int64_t)blockIdx.x * kFramesPerCta + lane;
    const int64_t s = q / t_max;
    const int t = (int)(q - s * t_max);
    int valid = 0, readable = 0;
    if (s < n_seg) {
      const int len0 = __ldg(seg_len + s);
      const int len = halved(len0, octave);
      readable = octave == 0 ? min(len, __ldg(seg_valid + s)) : len;
      valid = t < frames_of(len0, hop0, n_oct);
    }
    meta[lane] = (int)(s & 0xffffffff);
    meta[32 + lane] = (int)(s >> 32);
    meta[64 + lane] = t;
    meta[96 + lane] = readable;
    meta[128 + lane] = valid;
    meta[160 + lane] = 0;
  }
  const float4* gw = reinterpret_cast<const float4*>(filters) + (size_t)octave * n_fft * groups;
  for (int i = threadIdx.x; i < n_fft * groups; i += blockDim.x) w4[i] = __ldg(gw + i);
  __syncthreads();
  for (int f = warp; f < kFramesPerCta; f += n_warps) {
    float* row = fr + f * fstride;
    if (!meta[128 + f]) {
      for (int n = lane; n < n_fft; n += 32) row[n] = 0.f;
      continue;
    }
    const int64_t s = ((int64_t)meta[32 + f] << 32) | (uint32_t)meta[f];
    const In* x = src + (seg_start ? __ldg(seg_start + s) : s * src_stride);
    const int first = meta[64 + f] * hop - (n_fft >> 1);
    const int readable = meta[96 + f];
    for (int n = lane; n < n_fft; n += 32) {
      const int i = first + n;
      row[n] = (i >= 0 && i < readable) ? s_load(x + i) : 0.f;
    }
  }
  __syncthreads();

  const float* row = fr + lane * fstride;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int n = 0; n < n_fft; ++n) {
    const float xv = row[n];
    const float4 w = w4[n * groups + warp];
    acc.x = fmaf(xv, w.x, acc.x);
    acc.y = fmaf(xv, w.y, acc.y);
    acc.z = fmaf(xv, w.z, acc.z);
    acc.w = fmaf(xv, w.w, acc.w);
  }
  const int valid = meta[128 + lane];
  const int64_t s = ((int64_t)meta[32 + lane] << 32) | (uint32_t)meta[lane];
  const int t = meta[64 + lane];
  const int b0 = 2 * warp, b1 = 2 * warp + 1;
  if (s < n_seg) {
    if (complex_out) {
      float2* o = reinterpret_cast<float2*>(out) + (s * n_bins) * t_max + t;
      if (b0 < bin_cnt) o[(int64_t)(bin_lo + b0) * t_max] = valid ? make_float2(acc.x, acc.y) : make_float2(0.f, 0.f);
      if (b1 < bin_cnt) o[(int64_t)(bin_lo + b1) * t_max] = valid ? make_float2(acc.z, acc.w) : make_float2(0.f, 0.f);
    } else {
      const float m0 = valid ? fmaf(acc.x, acc.x, acc.y * acc.y) : 0.f;
      const float m1 = valid ? fmaf(acc.z, acc.z, acc.w * acc.w) : 0.f;
      float* o = out + (s * n_bins) * t_max + t;
      if (b0 < bin_cnt) o[(int64_t)(bin_lo + b0) * t_max] = m0;
      if (b1 < bin_cnt) o[(int64_t)(bin_lo + b1) * t_max] = m1;
      float m = 0.f;
      if (b0 < bin_cnt) m = m0;
      if (b1 < bin_cnt) m = fmaxf(m, m1);
      if (valid) atomicMax(meta + 160 + lane, __float_as_int(m));       // m >= 0: int order == float order
    }
  }
  if (!complex_out) {
    __syncthreads();
    if (warp == 0) {
      // one global atomic per run of equal segments inside the CTA's 32 consecutive frames
      const bool live = valid && s < n_seg;
      const int mine = live ? meta[160 + lane] : 0;
      const int64_t key = live ? s : (int64_t)(-1 - lane);      // frames past a segment's T separate the runs
      const int64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
      const bool head = lane == 0 || prev != key;
      int best = mine;
      for (int d = 1; d < 32; ++d) {
        const int64_t ko = __shfl_down_sync(0xffffffffu, key, d);
        const int vo = __shfl_down_sync(0xffffffffu, mine, d);
        if (lane + d < 32 && ko == key) best = max(best, vo);
      }
      if (head && live) atomicMax(reinterpret_cast<int*>(segmax) + s, best);
    }
  }
}

// in-place |C|^2 -> dB  (DbScale: the arithmetic of finish_db_kernel, cqt_frame_finish.cu), 16 bytes per thread and trip
__global__ void __launch_bounds__(256)
sfinish_kernel(float* __restrict__ io, const float* __restrict__ segmax, int64_t n_seg, int per_seg, float power,
               float amin, float top_db, float cut_db, float floor_db) {
  pdl_launch_dependents();
  pdl_wait();                              // tensor path: launched behind the last response GEMM (gtc_common.cuh)
  const int64_t total = n_seg * per_seg;
  if ((per_seg & 3) == 0) {
    float4* io4 = reinterpret_cast<float4*>(io);
    const int per4 = per_seg >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (total >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = io4[i];
      const DbScale scale(__ldg(segmax + i / per4), power, amin, top_db, cut_db, floor_db);
      io4[i] = make_float4(scale(v.x), scale(v.y), scale(v.z), scale(v.w));
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const DbScale scale(__ldg(segmax + i / per_seg), power, amin, top_db, cut_db, floor_db);
      io[i] = scale(io[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// tensor-core path: plane geometry, the split kernel and the launch sequence
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPlaneFront = 512;          // zero samples in front of segment 0 (windows start before their segment)
constexpr float kPlaneScale = 256.f;      // == the fp16x2 engine's x_scale (cqt_api.cu)

struct TcGeom {
  int64_t S[kMaxOctaves];                 // samples per segment slot of octave i (multiple of 256, >= len_i + guard) = slot stride
  int64_t R[kMaxOctaves];                 // samples of a slot the decimator computes: whole pairs of 128-output rows covering the
                                          // LONGEST segment's octave-i length; [R, S) is zero-filled by pad_zero_kernel instead
                                          // (the short low octaves of a 0.2 s window are 35-138 samples in 512-sample slots)
  int64_t plane_elems[kMaxOctaves];       // halves per plane
  size_t off_hi[kMaxOctaves], off_lo[kMaxOctaves];
  int t_pad;                              // frame rows per segment: t_max rounded up to a pair (the smallest TMA box is 64 segments x 2 rows)
};

static void tc_geometry(const SPlanImpl& p, int64_t n_seg, int64_t max_len, TcGeom& g) {
  const int t_max = frames_of((int)max_len, p.hop, p.n_oct);
  // Frame rows per segment: a TMA box takes 8, 4 or 2 consecutive rows of a segment (16 / 32 / 64 segments per 128-row block).
  // Fewer rows per segment and box mean shorter runs in the response epilogue's stores (a run = the box's frames of one bin) and
  // more segment maxima per tile.  Measured on B200 (cqt.py recipe, 9 frames: 16 rows in 8-row boxes 147 us, 12 rows in 4-row
  // boxes 149 us, 10 rows in 2-row boxes 135 us for the 8 response launches): a row costs ~1.3 x in a 4-row box and ~1.45 x in a
  // 2-row box, so 9 frames take 10 rows and the 130 frames of a 3 s segment stay at 136.
  {
    const double c8 = (double)round_up(t_max, 8), c4 = 1.3 * (double)round_up(t_max, 4), c2 = 1.45 * (double)round_up(t_max, 2);
    g.t_pad = (int)(c8 <= c4 && c8 <= c2 ? round_up(t_max, 8) : c4 <= c2 ? round_up(t_max, 4) : round_up(t_max, 2));
  }
  int64_t len = max_len;
  const int64_t guard = p.dec_left + 8 > p.n_fft / 2 ? p.dec_left + 8 : p.n_fft / 2;
  for (int i = 0; i < p.n_oct; ++i) {
    g.S[i] = round_up(len + guard, 256);                  // whole pairs of 128-output rows (the smallest TMA box is 64 segments x 2 rows)
    g.R[i] = round_up(len, 256) < g.S[i] ? round_up(len, 256) : g.S[i];
    len = (len + 1) / 2;
  }
  for (int i = 0; i < p.n_oct; ++i) {
    // furthest sample any window of the last segment touches, past that segment's slot
    int64_t reach = (int64_t)g.t_pad * (p.hop >> i) + p.n_fft;                       // response frames
    if (i + 1 < p.n_oct) reach = reach > 2 * g.R[i + 1] + p.dec_k ? reach : 2 * g.R[i + 1] + p.dec_k;   // decimator windows
    const int64_t tail = reach > g.S[i] ? reach - g.S[i] : 0;
    g.plane_elems[i] = round_up(kPlaneFront + n_seg * g.S[i] + tail + 64, 512);
  }
}

// The zero front and the tail of every plane (everything between them is written by the split kernel / the decimator
// epilogues).  The tail matters: a window of the last segment reaches past its slot, and although the operator is zero
// there, 0 x (uninitialised NaN bit pattern) would poison the row.
struct PadList {
  __half* plane[2 * kMaxOctaves];
  int64_t tail_at[2 * kMaxOctaves], tail_len[2 * kMaxOctaves];
  int64_t slot_stride[2 * kMaxOctaves];   // per slot: samples [gap_at, slot_stride) are not written by any decimator row
  int gap_at[2 * kMaxOctaves];
  int n;
  const int32_t* seg_len;       // also: frames kept per segment, for the response epilogues
  int32_t* seg_frames;
  int64_t n_seg;
  int hop, n_oct;
};
__global__ void __launch_bounds__(256) pad_zero_kernel(const PadList pl) {
  pdl_launch_dependents();                 // launched normally (its predecessor is a memset); the split kernel may overlap its tail
  if (blockIdx.y == 0)
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < pl.n_seg; s += (int64_t)gridDim.x * blockDim.x)
      pl.seg_frames[s] = frames_of(__ldg(pl.seg_len + s), pl.hop, pl.n_oct);
  for (int i = blockIdx.y; i < pl.n; i += gridDim.y) {
    __half* p = pl.plane[i];
    const int64_t total = kPlaneFront + pl.tail_len[i];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x)
      p[k < kPlaneFront ? k : pl.tail_at[i] + (k - kPlaneFront)] = __float2half(0.f);
    const int64_t gap = pl.slot_stride[i] - pl.gap_at[i];            // multiple of 256 samples: 16-byte vectors
    if (gap > 0) {
      const int64_t per = gap >> 3, vecs = pl.n_seg * per;
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < vecs; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = k / per;
        *reinterpret_cast<uint4*>(p + kPlaneFront + s * pl.slot_stride[i] + pl.gap_at[i] + ((k - s * per) << 3)) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

// octave-0 planes: every sample of every slot is written (zeros beyond the segment), 8 samples per thread
template <typename In>
__global__ void __launch_bounds__(256)
split_kernel(const In* __restrict__ audio, const int64_t* __restrict__ seg_start, const int32_t* __restrict__ seg_valid,
             const int32_t* __restrict__ seg_len, int64_t n_seg, int64_t S, __half* __restrict__ hi, __half* __restrict__ lo) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t per_seg = S >> 3;
  const int64_t total = n_seg * per_seg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / per_seg;
    const int64_t k0 = (i - s * per_seg) << 3;
    const int readable = min(__ldg(seg_len + s), __ldg(seg_valid + s));
    const In* x = audio + __ldg(seg_start + s);
    __align__(16) __half h8[8], l8[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float v = (k0 + u < readable) ? s_load(x + k0 + u) * kPlaneScale : 0.f;
      h8[u] = __float2half_rn(v);
      l8[u] = __float2half_rn(v - __half2float(h8[u]));
    }
    const int64_t at = kPlaneFront + s * S + k0;
    *reinterpret_cast<uint4*>(hi + at) = *reinterpret_cast<const uint4*>(h8);
    *reinterpret_cast<uint4*>(lo + at) = *reinterpret_cast<const uint4*>(l8);
  }
}

static inline size_t dec_smem_bytes(int n_taps) {
  const int c = (n_taps - 1) / 2, nt = (c + 4) & ~3;
  return (size_t)(2 * nt + 2 * (kDecTile + nt + 4)) * sizeof(float);
}

struct SWorkspace {
  OctaveBufs bufs;
  TcGeom tc;
  size_t off_segmax, off_segframes, total, total_tc;
};

static SWorkspace s_layout(const SPlanImpl& p, int64_t n_seg, int64_t max_len) {
  SWorkspace w;
  memset(&w, 0, sizeof(w));
  size_t o = 0;
  auto take = [&](size_t b) { size_t at = o; o += (b + 1023) & ~(size_t)1023; return at; };
  w.off_segmax = take((size_t)(n_seg > 0 ? n_seg : 1) * sizeof(float));
  w.off_segframes = take((size_t)(n_seg > 0 ? n_seg : 1) * sizeof(int32_t));
  if (p.use_tc && n_seg > 0 && max_len > 0) {             // tensor path: hi/lo planes of every octave, after the maxima
    size_t o_tc = o;
    auto take_tc = [&](size_t b) { size_t at = o_tc; o_tc += (b + 1023) & ~(size_t)1023; return at; };
    tc_geometry(p, n_seg, max_len, w.tc);
    for (int i = 0; i < p.n_oct; ++i) {
      w.tc.off_hi[i] = take_tc((size_t)w.tc.plane_elems[i] * sizeof(__half));
      w.tc.off_lo[i] = take_tc((size_t)w.tc.plane_elems[i] * sizeof(__half));
    }
    w.total_tc = o_tc;
  }
  int64_t len = max_len;
  for (int i = 1; i < p.n_oct; ++i) {
    len = (len + 1) / 2;
    w.bufs.stride[i] = round_up(len, 4);
    w.bufs.off[i] = (int64_t)(take((size_t)n_seg * w.bufs.stride[i] * sizeof(float)) / sizeof(float));
  }
  w.total = o > w.total_tc ? o : w.total_tc;              // one workspace serves both paths (complex output is SIMT only)
  return w;
}

// dB features with both contractions on the tensor cores
template <typename In>
static int run_structured_tc(const SPlanImpl& p, const In* d_audio, const int64_t* d_seg_start, const int32_t* d_seg_valid,
                             const int32_t* d_seg_len, int64_t n_seg, int64_t max_len, float* d_out, char* ws, const SWorkspace& w,
                             float power, float amin, float top_db, float cut_db, float floor_db, cudaStream_t st) {
  const TcGeom& g = w.tc;
  float* segmax = reinterpret_cast<float*>(ws + w.off_segmax);
  int32_t* seg_frames = reinterpret_cast<int32_t*>(ws + w.off_segframes);
  const int t_max = frames_of((int)max_len, p.hop, p.n_oct);
  GTC_CUDA_CHECK(cudaMemsetAsync(segmax, 0, (size_t)n_seg * sizeof(float), st));
  auto hi = [&](int i) { return reinterpret_cast<__half*>(ws + g.off_hi[i]); };
  auto lo = [&](int i) { return reinterpret_cast<__half*>(ws + g.off_lo[i]); };
  {
    PadList pl;
    pl.n = 2 * p.n_oct;
    for (int i = 0; i < p.n_oct; ++i)
      for (int h = 0; h < 2; ++h) {
        pl.plane[2 * i + h] = h ? lo(i) : hi(i);
        pl.tail_at[2 * i + h] = kPlaneFront + n_seg * g.S[i];
        pl.tail_len[2 * i + h] = g.plane_elems[i] - (kPlaneFront + n_seg * g.S[i]);
        pl.slot_stride[2 * i + h] = g.S[i];
        pl.gap_at[2 * i + h] = i == 0 ? (int)g.S[i] : (int)g.R[i];    // octave 0 is written whole by the split kernel
      }
    pl.seg_len = d_seg_len; pl.seg_frames = seg_frames; pl.n_seg = n_seg; pl.hop = p.hop; pl.n_oct = p.n_oct;
    pad_zero_kernel<<<dim3(128, (unsigned)pl.n), 256, 0, st>>>(pl);
    GTC_CUDA_CHECK(cudaGetLastError());
  }
  {
    int64_t blocks = ceil_div(n_seg * (g.S[0] >> 3), 256);
    const int64_t cap = (int64_t)p.sm_count * 32;
    if (blocks > cap) blocks = cap;
    static const bool pdl = getenv("GTC_SCQT_NO_PDL") == nullptr;
    GTC_CUDA_CHECK(launch_pdl(split_kernel<In>, dim3((unsigned)blocks), dim3(256), 0, st, pdl, d_audio, d_seg_start, d_seg_valid, d_seg_len, n_seg,
                              g.S[0], hi(0), lo(0)));
    GTC_CUDA_CHECK(cudaGetLastError());
  }
  SlotArgs sl;
  memset(&sl, 0, sizeof(sl));
  sl.n_slots = n_seg;
  sl.seg_len = d_seg_len;
  sl.seg_frames = seg_frames;
  sl.plane_scale = kPlaneScale;
  const PlanImpl& dec = *reinterpret_cast<const PlanImpl*>(p.dec_plan);
  // Launch order: the response of octave i right after the decimation that produced it (and before the one that consumes it), so
  // that the planes of the small octaves are still in L2 for both readers.  GTC_SCQT_ORDER=1: all decimations, then all responses.
  static const bool grouped = getenv("GTC_SCQT_ORDER") != nullptr && atoi(getenv("GTC_SCQT_ORDER")) == 1;
  auto decimate = [&](int i) {                            // octave i+1 = 2:1 decimation of octave i
    SlotArgs d = sl;
    d.slot_mode = 1;
    d.out_hi = hi(i + 1); d.out_lo = lo(i + 1);
    d.out_stride = g.S[i + 1]; d.out_base = kPlaneFront;
    d.stage_out = i + 1;
    return launch_gemm_tc_slots(dec, hi(i), lo(i), kPlaneFront - p.dec_left, g.S[i], 256, (int)(g.R[i + 1] / 128), d, st);
  };
  auto respond = [&](int i) {
    SlotArgs r = sl;
    r.slot_mode = 2;
    r.out = d_out; r.segmax = segmax;
    r.n_bins = p.n_bins; r.t_max = t_max; r.hop0 = p.hop; r.n_oct = p.n_oct;
    r.bin_lo = p.bin_lo[i]; r.bin_cnt = p.bin_cnt[i];
    const PlanImpl& rp = *reinterpret_cast<const PlanImpl*>(p.resp_plan[i]);
    return launch_gemm_tc_slots(rp, hi(i), lo(i), kPlaneFront - p.n_fft / 2, g.S[i], p.hop >> i, g.t_pad, r, st);
  };
  if (grouped) {
    for (int i = 0; i + 1 < p.n_oct; ++i) { int rc = decimate(i); if (rc != GTC_OK) return rc; }
    for (int i = 0; i < p.n_oct; ++i) { int rc = respond(i); if (rc != GTC_OK) return rc; }
  } else {
    for (int i = 0; i < p.n_oct; ++i) {
      int rc = respond(i);
      if (rc == GTC_OK && i + 1 < p.n_oct) rc = decimate(i);
      if (rc != GTC_OK) return rc;
    }
  }
  const int per_seg = p.n_bins * t_max;
  int64_t fblocks = ceil_div(n_seg * per_seg, 256 * 4);
  const int64_t cap = (int64_t)p.sm_count * 16;
  if (fblocks > cap) fblocks = cap;
  if (fblocks < 1) fblocks = 1;
  static const bool pdl = getenv("GTC_SCQT_NO_PDL") == nullptr;
  GTC_CUDA_CHECK(launch_pdl(sfinish_kernel, dim3((unsigned)fblocks), dim3(256), 0, st, pdl, d_out, (const float*)segmax, n_seg, per_seg, power, amin,
                            top_db, cut_db, floor_db));
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

template <typename In>
static int run_structured(const SPlanImpl& p, const In* d_audio, const int64_t* d_seg_start, const int32_t* d_seg_valid,
                          const int32_t* d_seg_len, int64_t n_seg, int64_t max_len, float* d_out, bool complex_out,
                          char* ws, const SWorkspace& w, float power, float amin, float top_db, float cut_db,
                          float floor_db, cudaStream_t st) {
  float* wsf = reinterpret_cast<float*>(ws);
  float* segmax = reinterpret_cast<float*>(ws + w.off_segmax);
  const int t_max = frames_of((int)max_len, p.hop, p.n_oct);
  if (!complex_out) GTC_CUDA_CHECK(cudaMemsetAsync(segmax, 0, (size_t)n_seg * sizeof(float), st));

  // decimation chain: octave i+1 from octave i
  const size_t dec_smem = dec_smem_bytes(p.n_taps);
  int64_t len = max_len;
  for (int i = 0; i + 1 < p.n_oct; ++i) {
    const int64_t out_len = (len + 1) / 2;
    const int tiles = (int)ceil_div(out_len, kDecTile);
    const int64_t blocks = n_seg * tiles;
    GTC_REQUIRE(blocks < 0x7fffffffLL, GTC_E_ARG, "gtc_scqt: too many decimator tiles (%lld); split the batch", (long long)blocks);
    float* dst = wsf + w.bufs.off[i + 1];
    if (i == 0)
      decimate2_kernel<In><<<(unsigned)blocks, 256, dec_smem, st>>>(d_audio, d_seg_start, d_seg_valid, d_seg_len, 0, 0, dst,
                                                                     w.bufs.stride[1], p.d_taps, p.n_taps, tiles, 1.f);
    else
      decimate2_kernel<float><<<(unsigned)blocks, 256, dec_smem, st>>>(wsf + w.bufs.off[i], nullptr, d_seg_valid, d_seg_len, i,
                                                                        w.bufs.stride[i], dst, w.bufs.stride[i + 1], p.d_taps,
                                                                        p.n_taps, tiles, 1.f);
    GTC_CUDA_CHECK(cudaGetLastError());
    len = out_len;
  }

  const size_t resp_smem = ((size_t)p.n_fft * p.groups * 4 + (size_t)kFramesPerCta * (p.n_fft + 1) + 192) * sizeof(float);
  const int64_t rblocks = ceil_div(n_seg * t_max, kFramesPerCta);
  GTC_REQUIRE(rblocks < 0x7fffffffLL, GTC_E_ARG, "gtc_scqt: too many frames (%lld blocks); split the batch", (long long)rblocks);
  for (int i = 0; i < p.n_oct; ++i) {
    if (i == 0)
      response_kernel<In><<<(unsigned)rblocks, 32 * p.groups, resp_smem, st>>>(
          d_audio, d_seg_start, d_seg_valid, d_seg_len, n_seg, 0, p.n_oct, p.hop, 0, p.d_filters, p.n_fft, p.groups, p.bin_lo[0],
          p.bin_cnt[0], p.n_bins, t_max, d_out, complex_out ? 1 : 0, segmax);
    else
      response_kernel<float><<<(unsigned)rblocks, 32 * p.groups, resp_smem, st>>>(
          wsf + w.bufs.off[i], nullptr, d_seg_valid, d_seg_len, n_seg, i, p.n_oct, p.hop, w.bufs.stride[i], p.d_filters, p.n_fft,
          p.groups, p.bin_lo[i], p.bin_cnt[i], p.n_bins, t_max, d_out, complex_out ? 1 : 0, segmax);
    GTC_CUDA_CHECK(cudaGetLastError());
  }
  if (complex_out) return GTC_OK;
  const int per_seg = p.n_bins * t_max;
  int64_t fblocks = ceil_div(n_seg * per_seg, 256 * 4);
  const int64_t cap = (int64_t)p.sm_count * 16;
  if (fblocks > cap) fblocks = cap;
  if (fblocks < 1) fblocks = 1;
  sfinish_kernel<<<(unsigned)fblocks, 256, 0, st>>>(d_out, segmax, n_seg, per_seg, power, amin, top_db, cut_db, floor_db);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

}  // namespace gtc

using namespace gtc;

struct gtc_splan {
  SPlanImpl impl;
};

extern "C" int gtc_scqt_frames(int64_t seg_len, int hop_length, int n_octaves) {
  if (seg_len < 0 || seg_len > 0x7fffffffLL || hop_length <= 0 || n_octaves <= 0 || n_octaves > kMaxOctaves) return GTC_E_ARG;
  return frames_of((int)seg_len, hop_length, n_octaves);
}

extern "C" int gtc_scqt_plan_create(gtc_splan** out, int device, int n_octaves, int n_fft, int hop_length, int n_bins,
                                    int filters_per_octave, const float* h_filters, const float* h_taps, int n_taps) {
  GTC_REQUIRE(out != nullptr, GTC_E_ARG, "gtc_scqt_plan_create: out is NULL");
  *out = nullptr;
  GTC_REQUIRE(h_filters && h_taps, GTC_E_ARG, "gtc_scqt_plan_create: NULL filter/tap table (design them with gtc_b200.cqt_design)");
  GTC_REQUIRE(n_octaves > 0 && n_octaves <= kMaxOctaves, GTC_E_UNSUP, "gtc_scqt_plan_create: 1..%d octaves supported", kMaxOctaves);
  GTC_REQUIRE(n_fft >= 32 && (n_fft & (n_fft - 1)) == 0 && n_fft <= 4096, GTC_E_UNSUP, "gtc_scqt_plan_create: n_fft must be a power of two in [32, 4096]");
  GTC_REQUIRE(n_bins > 0 && filters_per_octave > 0 && filters_per_octave <= 64, GTC_E_UNSUP, "gtc_scqt_plan_create: 1..64 filters per octave supported");
  GTC_REQUIRE(n_bins <= filters_per_octave * n_octaves && n_bins > filters_per_octave * (n_octaves - 1), GTC_E_ARG,
              "gtc_scqt_plan_create: n_bins inconsistent with octaves x filters");
  GTC_REQUIRE(hop_length > 0 && hop_length % (1 << (n_octaves - 1)) == 0, GTC_E_ARG,
              "gtc_scqt_plan_create: hop_length must be a multiple of 2^(n_octaves-1) (librosa's own requirement)");
  GTC_REQUIRE(n_taps >= 5 && n_taps % 4 == 1 && n_taps <= 4097, GTC_E_UNSUP, "gtc_scqt_plan_create: n_taps must be 1 (mod 4), <= 4097");
  GTC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GTC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  GTC_REQUIRE(prop.major == 10, GTC_E_UNSUP, "libgtc is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);

  gtc_splan* plan = new (std::nothrow) gtc_splan();
  GTC_REQUIRE(plan != nullptr, GTC_E_NOMEM, "gtc_scqt_plan_create: out of host memory");
  SPlanImpl& p = plan->impl;
  memset(&p, 0, sizeof(p));
  p.device = device; p.sm_count = prop.multiProcessorCount;
  p.n_oct = n_octaves; p.n_fft = n_fft; p.hop = hop_length; p.n_bins = n_bins; p.n_filters = filters_per_octave; p.n_taps = n_taps;
  p.groups = (2 * filters_per_octave + 3) / 4;
  for (int i = 0; i < n_octaves; ++i) {            // librosa.vqt: octave i owns the top-most remaining filters
    const int hi = n_bins - filters_per_octave * i;
    const int lo = hi - filters_per_octave > 0 ? hi - filters_per_octave : 0;
    p.bin_lo[i] = lo; p.bin_cnt[i] = hi - lo;
  }
  const size_t resp_smem = ((size_t)p.n_fft * p.groups * 4 + (size_t)kFramesPerCta * (p.n_fft + 1) + 192) * sizeof(float);
  const size_t dec_smem = dec_smem_bytes(n_taps);
  if (resp_smem > 227 * 1024 || dec_smem > 227 * 1024) {
    delete plan;
    set_error("gtc_scqt_plan_create: n_fft %d x %d filters needs %zu B of shared memory (> 227 KB)", n_fft, filters_per_octave, resp_smem);
    return GTC_E_UNSUP;
  }
  // h_filters [n_oct][2*fpo][n_fft] (row = filter*2 + {re,im}) -> device [n_oct][n_fft][groups*4]
  const int n_real = 2 * filters_per_octave, g4 = p.groups * 4;
  std::vector<float> tr((size_t)n_octaves * n_fft * g4, 0.f);
  for (int i = 0; i < n_octaves; ++i)
    for (int r = 0; r < n_real; ++r)
      for (int n = 0; n < n_fft; ++n)
        tr[((size_t)i * n_fft + n) * g4 + r] = h_filters[((size_t)i * n_real + r) * n_fft + n];
  cudaError_t e = cudaMalloc((void**)&p.d_filters, tr.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(p.d_filters, tr.data(), tr.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p.d_taps, (size_t)n_taps * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(p.d_taps, h_taps, (size_t)n_taps * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(response_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resp_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(response_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resp_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(decimate2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(decimate2_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem);
  if (e != cudaSuccess) {
    set_error("gtc_scqt_plan_create: %s", cudaGetErrorString(e));
    if (p.d_filters) cudaFree(p.d_filters);
    if (p.d_taps) cudaFree(p.d_taps);
    delete plan;
    return GTC_E_CUDA;
  }
  // ---- tensor-core path (dB output): the decimator as a banded Toeplitz operator, the octave filters as [32][n_fft]
  //      operators, each an fp16x2 plan of the segment-operator engine (cqt_api.cu).  Window geometry of the decimator:
  //      row j of a segment produces outputs k = 128 j + n, n < 128, from the inputs 256 j - left .. 256 j - left + K - 1,
  //      left = c rounded up to 32 (64-byte window starts), out[k] = sum_m h[m] x[2k + c - m]  =>  Op[n][i] = h[2n + c + left - i].
  p.use_tc = 0;
  bool eligible = filters_per_octave <= 16 && n_fft % 32 == 0 && getenv("GTC_SCQT_SIMT") == nullptr;
  for (int i = 0; i < n_octaves; ++i) eligible = eligible && ((hop_length >> i) % 8 == 0);
  if (eligible) {
    const int c = (n_taps - 1) / 2;
    // windows start on 64-byte boundaries (left % 32 == 0 with kPlaneFront % 32 == 0): every 64-byte k-block row of the TMA
    // box is then two whole 32-byte sectors.  With left = c rounded to 8 (16-byte aligned only) each row straddled three
    // sectors and the kernel, which is bound by L2 -> SM bandwidth, moved 1.5 x the bytes (ncu: 4.54 GB per 86 M outputs).
    p.dec_left = (int)round_up(c, 32);
    p.dec_k = (int)round_up(p.dec_left + 254 + c + 1, 32);
    eligible = p.dec_left + 8 <= kPlaneFront && n_fft / 2 <= kPlaneFront;
  }
  if (eligible) {
    std::vector<float> op((size_t)128 * p.dec_k, 0.f);
    const int c = (n_taps - 1) / 2;
    for (int n = 0; n < 128; ++n)
      for (int i = 0; i < p.dec_k; ++i) {
        const int m = 2 * n + c + p.dec_left - i;
        if (m >= 0 && m < n_taps) op[(size_t)n * p.dec_k + i] = h_taps[m];
      }
    gtc_plan* sub = nullptr;
    int rc = tc_slots_init();
    if (rc == GTC_OK) rc = gtc_cqt_plan_create(&sub, device, p.dec_k, p.dec_k, 64, 1, op.data(), GTC_GEMM_TCGEN05_FP16X2);
    p.dec_plan = sub;
    if (rc == GTC_OK) {
      // The operator is banded Toeplitz: k-block kb + 1 is k-block kb moved down by (elements per k-block) / 2 rows.  Verified on
      // the values just built (the hi/lo split is elementwise, so the planes inherit it); the slotted GEMM then keeps ONE master
      // tile of it resident in shared memory instead of streaming 16 KB of operator with every k-block (cqt_gemm_tc.cu: RES).
      PlanImpl& dp = *reinterpret_cast<PlanImpl*>(sub);
      const int e = dp.kb_elems, sh = e / 2;
      bool toeplitz = p.dec_k % e == 0;
      for (int n = sh; n < 128 && toeplitz; ++n)
        for (int i = 0; i + e < p.dec_k; ++i)
          if (op[(size_t)n * p.dec_k + i + e] != op[(size_t)(n - sh) * p.dec_k + i]) { toeplitz = false; break; }
      dp.res_shift = (toeplitz && getenv("GTC_SCQT_STREAM_OP") == nullptr) ? sh : 0;    // the switch is for A/B runs and the bit-identity test
      // The band: 389 taps against windows of 704 samples leave 45 % of the operator zero -- k-block kb only reaches the output rows
      // n with 0 <= 2n + c + left - i < taps for some i of the block.  Per k-block the contiguous range of 16-row groups that hold
      // a non-zero (tcgen05.mma N is a multiple of 16), a full k-block first because the first MMA initialises the accumulator.
      const int nkb = p.dec_k / e;
      std::vector<int> g0(nkb, -1), g1(nkb, -1);
      for (int kb = 0; kb < nkb; ++kb)
        for (int n = 0; n < 128; ++n)
          for (int i = kb * e; i < (kb + 1) * e; ++i)
            if (op[(size_t)n * p.dec_k + i] != 0.f) { if (g0[kb] < 0) g0[kb] = n / 16; g1[kb] = n / 16; break; }
      int full = -1;
      for (int kb = 0; kb < nkb && full < 0; ++kb)
        if (g0[kb] == 0 && g1[kb] == 7) full = kb;
      dp.band_n = 0;
      if (full >= 0 && nkb <= kMaxBand) {
        auto push = [&](int kb) {
          dp.band_kb[dp.band_n] = (uint8_t)kb; dp.band_g0[dp.band_n] = (uint8_t)g0[kb]; dp.band_ng[dp.band_n] = (uint8_t)(g1[kb] - g0[kb] + 1);
          ++dp.band_n;
        };
        push(full);
        for (int kb = 0; kb < nkb; ++kb)
          if (kb != full && g0[kb] >= 0) push(kb);
      }
    }
    for (int i = 0; i < n_octaves && rc == GTC_OK; ++i) {
      std::vector<float> f((size_t)32 * n_fft, 0.f);
      for (int r = 0; r < n_real && r < 32; ++r)
        memcpy(&f[(size_t)r * n_fft], &h_filters[((size_t)i * n_real + r) * n_fft], (size_t)n_fft * sizeof(float));
      sub = nullptr;
      rc = gtc_cqt_plan_create(&sub, device, n_fft, n_fft, 16, 1, f.data(), GTC_GEMM_TCGEN05_FP16X2);
      p.resp_plan[i] = sub;
    }
    if (rc != GTC_OK) {
      gtc_scqt_plan_destroy(plan);
      return rc;
    }
    p.use_tc = 1;
  }
  *out = plan;
  return GTC_OK;
}

extern "C" int gtc_scqt_plan_destroy(gtc_splan* plan) {
  if (!plan) return GTC_OK;
  if (plan->impl.d_filters) cudaFree(plan->impl.d_filters);
  if (plan->impl.d_taps) cudaFree(plan->impl.d_taps);
  if (plan->impl.dec_plan) gtc_cqt_plan_destroy(reinterpret_cast<gtc_plan*>(plan->impl.dec_plan));
  for (int i = 0; i < kMaxOctaves; ++i)
    if (plan->impl.resp_plan[i]) gtc_cqt_plan_destroy(reinterpret_cast<gtc_plan*>(plan->impl.resp_plan[i]));
  delete plan;
  return GTC_OK;
}

extern "C" int gtc_scqt_workspace_bytes(const gtc_splan* plan, int64_t n_seg, int64_t max_len, size_t* bytes) {
  GTC_REQUIRE(plan && bytes && n_seg >= 0 && max_len >= 0 && max_len < 0x7fffffffLL, GTC_E_ARG, "gtc_scqt_workspace_bytes: bad argument");
  *bytes = s_layout(plan->impl, n_seg, max_len).total;
  return GTC_OK;
}

static int scqt_run(const gtc_splan* plan, const void* d_audio, int sample_format, const int64_t* d_seg_start,
                    const int32_t* d_seg_valid, const int32_t* d_seg_len, int64_t n_seg, int64_t max_len, float* d_out,
                    bool complex_out, void* d_workspace, size_t workspace_bytes, float power, float amin, float top_db,
                    float cut_db, float floor_db, cudaStream_t st) {
  GTC_REQUIRE(plan != nullptr, GTC_E_ARG, "gtc_scqt: plan is NULL");
  GTC_REQUIRE(n_seg >= 0 && max_len >= 0 && max_len < 0x7fffffffLL, GTC_E_ARG, "gtc_scqt: sizes out of range");
  GTC_REQUIRE(sample_format == GTC_SAMPLES_F32 || sample_format == GTC_SAMPLES_PCM16, GTC_E_ARG, "gtc_scqt: unknown sample format %d", sample_format);
  if (n_seg == 0) return GTC_OK;
  GTC_REQUIRE(d_audio && d_seg_start && d_seg_valid && d_seg_len && d_out && d_workspace, GTC_E_ARG, "gtc_scqt: null pointer");
  const SPlanImpl& p = plan->impl;
  int dev = -1;
  GTC_CUDA_CHECK(cudaGetDevice(&dev));
  GTC_REQUIRE(dev == p.device, GTC_E_ARG, "gtc_scqt: plan belongs to device %d, current device is %d", p.device, dev);
  const SWorkspace w = s_layout(p, n_seg, max_len);
  GTC_REQUIRE(workspace_bytes >= w.total, GTC_E_NOMEM, "gtc_scqt: workspace of %zu bytes, %zu needed", workspace_bytes, w.total);
  GTC_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 15) == 0, GTC_E_ARG, "gtc_scqt: workspace must be 16-byte aligned");
  char* ws = static_cast<char*>(d_workspace);
  if (p.use_tc && !complex_out && max_len > 0) {
    GTC_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, GTC_E_ARG, "gtc_scqt: workspace must be 256-byte aligned");
    if (sample_format == GTC_SAMPLES_PCM16)
      return run_structured_tc(p, (const int16_t*)d_audio, d_seg_start, d_seg_valid, d_seg_len, n_seg, max_len, d_out, ws, w, power, amin,
                               top_db, cut_db, floor_db, st);
    return run_structured_tc(p, (const float*)d_audio, d_seg_start, d_seg_valid, d_seg_len, n_seg, max_len, d_out, ws, w, power, amin,
                             top_db, cut_db, floor_db, st);
  }
  if (sample_format == GTC_SAMPLES_PCM16)
    return run_structured(p, (const int16_t*)d_audio, d_seg_start, d_seg_valid, d_seg_len, n_seg, max_len, d_out, complex_out, ws, w,
                          power, amin, top_db, cut_db, floor_db, st);
  return run_structured(p, (const float*)d_audio, d_seg_start, d_seg_valid, d_seg_len, n_seg, max_len, d_out, complex_out, ws, w,
                        power, amin, top_db, cut_db, floor_db, st);
}

extern "C" int gtc_scqt_segments_db(const gtc_splan* plan, const void* d_audio, int sample_format, const int64_t* d_seg_start,
                                    const int32_t* d_seg_valid, const int32_t* d_seg_len, int64_t n_seg, int64_t max_len,
                                    float* d_out_db, void* d_workspace, size_t workspace_bytes, float power, float amin,
                                    float top_db, float cut_db, float floor_db, gtc_stream_t stream) {
  GTC_REQUIRE(power > 0.f && amin > 0.f, GTC_E_ARG, "gtc_scqt_segments_db: power and amin must be positive");
  return scqt_run(plan, d_audio, sample_format, d_seg_start, d_seg_valid, d_seg_len, n_seg, max_len, d_out_db, false, d_workspace,
                  workspace_bytes, power, amin, top_db, cut_db, floor_db, (cudaStream_t)stream);
}

extern "C" int gtc_scqt_segments_complex(const gtc_splan* plan, const void* d_audio, int sample_format, const int64_t* d_seg_start,
                                         const int32_t* d_seg_valid, const int32_t* d_seg_len, int64_t n_seg, int64_t max_len,
                                         float* d_out_c, void* d_workspace, size_t workspace_bytes, gtc_stream_t stream) {
  return scqt_run(plan, d_audio, sample_format, d_seg_start, d_seg_valid, d_seg_len, n_seg, max_len, d_out_c, true, d_workspace,
                  workspace_bytes, 1.f, 1e-5f, 80.f, -60.f, -120.f, (cudaStream_t)stream);
}

// One 2:1 soxr-HQ stage on its own: librosa.load(path, sr=native/2) / librosa.resample(orig_sr=2k, target_sr=k,
// res_type='soxr_hq') as called at /root/reference/tablature_generator.py:613,650 for 44.1 kHz files (scale=False there,
// so `gain` = 1/sqrt(2) undoes the sqrt(2) folded into the plan's taps; gain = 1 gives the CQT's own scale=True stage).
extern "C" int gtc_scqt_decimate(const gtc_splan* plan, const void* d_audio, int sample_format, const int64_t* d_seg_start,
                                 const int32_t* d_seg_valid, const int32_t* d_seg_len, int64_t n_seg, int64_t max_len,
                                 float* d_out, int64_t out_stride, float gain, gtc_stream_t stream) {
  GTC_REQUIRE(plan != nullptr, GTC_E_ARG, "gtc_scqt_decimate: plan is NULL");
  GTC_REQUIRE(n_seg >= 0 && max_len >= 0 && max_len < 0x7fffffffLL, GTC_E_ARG, "gtc_scqt_decimate: sizes out of range");
  GTC_REQUIRE(sample_format == GTC_SAMPLES_F32 || sample_format == GTC_SAMPLES_PCM16, GTC_E_ARG, "gtc_scqt_decimate: unknown sample format %d", sample_format);
  if (n_seg == 0 || max_len == 0) return GTC_OK;
  GTC_REQUIRE(d_audio && d_seg_start && d_seg_valid && d_seg_len && d_out, GTC_E_ARG, "gtc_scqt_decimate: null pointer");
  GTC_REQUIRE(out_stride >= (max_len + 1) / 2, GTC_E_ARG, "gtc_scqt_decimate: out_stride smaller than ceil(max_len/2)");
  const SPlanImpl& p = plan->impl;
  const size_t dec_smem = dec_smem_bytes(p.n_taps);
  const int tiles = (int)ceil_div((max_len + 1) / 2, kDecTile);
  const int64_t blocks = n_seg * tiles;
  GTC_REQUIRE(blocks < 0x7fffffffLL, GTC_E_ARG, "gtc_scqt_decimate: too many tiles; split the batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (sample_format == GTC_SAMPLES_PCM16)
    decimate2_kernel<int16_t><<<(unsigned)blocks, 256, dec_smem, st>>>((const int16_t*)d_audio, d_seg_start, d_seg_valid, d_seg_len, 0, 0,
                                                                        d_out, out_stride, p.d_taps, p.n_taps, tiles, gain);
  else
    decimate2_kernel<float><<<(unsigned)blocks, 256, dec_smem, st>>>((const float*)d_audio, d_seg_start, d_seg_valid, d_seg_len, 0, 0,
                                                                      d_out, out_stride, p.d_taps, p.n_taps, tiles, gain);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}
