// Framing (audio -> operand row matrix, hi/lo split) and the dB finish of the segment CQT.
//
// Framing: a segment of seg_len samples with hop seg_hop = seg_len / P is P consecutive "audio rows" of seg_hop
// samples, so the GEMM's M operand is the NON-overlapping row matrix and every audio sample is read once
// (/root/reference/cqt.py:26-45 re-slices each sample into two windows).  Clip c owns rows
// [seg_off[c] + c*(P-1), seg_off[c+1] + (c+1)*(P-1)); row r of the clip starts at sample clip_off[c] + r*seg_hop.
// Rows are padded to kp elements (a whole number of TC_KB_BYTES k-blocks = TMA/UMMA swizzle rows) and written twice:
//   fp16x2 engine (default): hi = fp16(x * 2^8), lo = fp16(x * 2^8 - hi)          -> 22 mantissa bits in two fp16 operands
//   3xTF32 engine          : hi = x rounded to tf32 (RN), lo = x - hi (exact fp32) -> 3xTF32 operands
// Input samples are fp32 (librosa.load's array) or the WAV file's int16 PCM (x / 32768, exact).
//
// Finish: per segment, mag2 = |C|^2 of all n_bins*T outputs plus the row maximum -> |C|^power ->
// librosa.amplitude_to_db(ref=np.amax, amin, top_db) -> cqt_lim (cqt.py:56-58), written as [n_seg, n_bins, T].
#include <cuda_fp16.h>
#include "gtc_common.cuh"

namespace gtc {

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u = __float_as_uint(x);
  u = (u + 0x00000fffu + ((u >> 13) & 1u)) & 0xffffe000u;     // round to nearest even on the 13 dropped bits
  return __uint_as_float(u);
}

// fp16x2 operands: the (power-of-two scaled) sample is split into hi = fp16(x) and lo = fp16(x - hi): 22 mantissa bits
__device__ __forceinline__ void split_store(__half* hi, __half* lo, int64_t i, const float (&v)[4], float scale) {
  __half h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x = v[j] * scale;
    h[j] = __float2half_rn(x);
    l[j] = __float2half_rn(x - __half2float(h[j]));
  }
  reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
  reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
}
__device__ __forceinline__ void split_store(float* hi, float* lo, int64_t i, const float (&v)[4], float) {
  float4 h, l;
  h.x = tf32_rn(v[0]); h.y = tf32_rn(v[1]); h.z = tf32_rn(v[2]); h.w = tf32_rn(v[3]);
  l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// input sample types: fp32 (what librosa.load returns) or the WAV file's own 16-bit PCM, converted exactly as
// librosa/soundfile do (x / 32768 in fp32, exact) -- half the host->device bytes of the end-to-end path
__device__ __forceinline__ float load_sample(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_sample(const int16_t* p) { return (float)__ldg(p) * (1.f / 32768.f); }

// one warp per audio row: lane 0 finds the owning clip once (the first version searched per float4: 552 binary searches
// per row), then the warp streams the row's kp/4 vectors
template <typename T, typename In>
__global__ void __launch_bounds__(256)
frame_kernel(const In* __restrict__ audio, const int64_t* __restrict__ clip_off, const int64_t* __restrict__ seg_off,
             int n_clips, int parts, int row_len, int seg_hop, int kp, int64_t n_rows, int64_t n_rows_alloc,
             T* __restrict__ xhi, T* __restrict__ xlo, float scale, float* __restrict__ rowmax, int64_t n_rowmax,
             int* __restrict__ tile_done, int* __restrict__ seg_of_row) {
  const int vec_per_row = kp >> 2;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  if (blockIdx.x == 0)                                   // row-block completion counters of the fused GEMM epilogue
    for (int64_t i = threadIdx.x; i < (n_rowmax >> 7); i += blockDim.x) tile_done[i] = 0;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows_alloc; row += warps_total) {
    int64_t base = 0, clip_end = 0;                 // clip_end == 0: a padding row, all zeros
    int seg = -1;                                   // segment that starts at this row (lane 0)
    if (row < n_rows) {
      if (lane == 0) {
        // clip of this row: rows of clip c start at seg_off[c] + c*(P-1)
        int lo = 0, hi = n_clips;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(seg_off + mid) + (int64_t)mid * (parts - 1) <= row) lo = mid; else hi = mid;
        }
        const int64_t r = row - (__ldg(seg_off + lo) + (int64_t)lo * (parts - 1));
        clip_end = __ldg(clip_off + lo + 1);
        base = __ldg(clip_off + lo) + r * seg_hop;
        const int64_t g = __ldg(seg_off + lo) + r;  // the clip's last P-1 rows start no segment
        if (g < __ldg(seg_off + lo + 1)) seg = (int)g;
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      clip_end = __shfl_sync(0xffffffffu, clip_end, 0);
    }
    const int64_t out0 = row * vec_per_row;
    for (int q = lane; q < vec_per_row; q += 32) {
      const int k0 = q << 2;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + j < row_len && base + k0 + j < clip_end) v[j] = load_sample(audio + base + k0 + j);
      split_store(xhi, xlo, out0 + q, v, scale);
    }
    if (lane == 0 && row < n_rowmax) { rowmax[row] = 0.f; seg_of_row[row] = seg; }
  }
}

template <typename In>
static int launch_frame_t(const PlanImpl& p, const In* d_audio, const int64_t* d_clip_off, const int64_t* d_seg_off,
                          int n_clips, int64_t n_rows, int64_t n_rows_alloc, void* d_xhi, void* d_xlo, float* d_rowmax,
                          int* d_tile_done, int* d_seg_of_row, cudaStream_t st) {
  int64_t blocks = ceil_div(n_rows_alloc, 8);
  const int64_t cap = (int64_t)p.sm_count * 16;
  if (blocks > cap) blocks = cap;
  const int64_t n_rowmax = round_up(n_rows, 128);
  if (p.elem_bytes == 2)
    frame_kernel<__half, In><<<(unsigned)blocks, 256, 0, st>>>(d_audio, d_clip_off, d_seg_off, n_clips, p.parts, p.row_len,
                                                          p.seg_hop, p.kp, n_rows, n_rows_alloc, (__half*)d_xhi, (__half*)d_xlo,
                                                          p.x_scale, d_rowmax, n_rowmax, d_tile_done, d_seg_of_row);
  else
    frame_kernel<float, In><<<(unsigned)blocks, 256, 0, st>>>(d_audio, d_clip_off, d_seg_off, n_clips, p.parts, p.row_len,
                                                         p.seg_hop, p.kp, n_rows, n_rows_alloc, (float*)d_xhi, (float*)d_xlo,
                                                         1.f, d_rowmax, n_rowmax, d_tile_done, d_seg_of_row);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

int launch_frame(const PlanImpl& p, const void* d_audio, int pcm16, const int64_t* d_clip_off, const int64_t* d_seg_off,
                 int n_clips, int64_t n_rows, int64_t n_rows_alloc, void* d_xhi, void* d_xlo, float* d_rowmax,
                 int* d_tile_done, int* d_seg_of_row, cudaStream_t st) {
  if (pcm16)
    return launch_frame_t(p, (const int16_t*)d_audio, d_clip_off, d_seg_off, n_clips, n_rows, n_rows_alloc, d_xhi, d_xlo, d_rowmax, d_tile_done, d_seg_of_row, st);
  return launch_frame_t(p, (const float*)d_audio, d_clip_off, d_seg_off, n_clips, n_rows, n_rows_alloc, d_xhi, d_xlo, d_rowmax, d_tile_done, d_seg_of_row, st);
}

// A warp converts one operand row per trip: row -> segment comes from frame_kernel's table (no search), the row maximum and
// the row's 4 vectors per lane are in flight before the first logarithm.  Measured on B200, same box, 28 200-row chunk
// (profiles/r02e_gemm_finish_ab.md): the r01 kernel (clip search per segment, transposing gather, log10f) 54 us; this
// one with log10f and 4 rows per warp (115 registers, half the occupancy) 62 us; with MUFU.LG2 44 us; with one row per
// warp (44 registers) 19.5 us -- the pass is latency-bound, so occupancy beats loads in flight per thread.
#ifndef GTC_FINISH_ROWS
#define GTC_FINISH_ROWS 1
#endif
constexpr int kFinishRows = GTC_FINISH_ROWS;
__global__ void __launch_bounds__(256)
finish_db_kernel(const float* __restrict__ mag2, const float* __restrict__ rowmax, const int* __restrict__ seg_of_row,
                 int64_t n_rows, int n_bins, int n_frames, float* __restrict__ out,
                 float power, float amin, float top_db, float cut_db, float floor_db) {
  const int lane = threadIdx.x & 31;
  const int n_mag = n_bins * n_frames;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t r0 = warp0 * kFinishRows; r0 < n_rows; r0 += warps_total * kFinishRows) {
    int seg[kFinishRows];
    float ref[kFinishRows];
#pragma unroll
    for (int k = 0; k < kFinishRows; ++k) {
      seg[k] = r0 + k < n_rows ? __ldg(seg_of_row + r0 + k) : -1;
      ref[k] = r0 + k < n_rows ? __ldcg(rowmax + r0 + k) : 0.f;
    }
    if ((n_mag & 3) == 0) {
      const int nv = n_mag >> 2;
      for (int ob = 0; ob < nv; ob += 128) {
        float4 v[kFinishRows][4];
#pragma unroll
        for (int k = 0; k < kFinishRows; ++k) {
          const float4* s4 = reinterpret_cast<const float4*>(mag2 + (r0 + k) * n_mag);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int o = ob + lane + 32 * u;
            v[k][u] = (seg[k] >= 0 && o < nv) ? __ldcg(s4 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int k = 0; k < kFinishRows; ++k) {
          if (seg[k] < 0) continue;
          const DbScale scale(ref[k], power, amin, top_db, cut_db, floor_db);
          float4* d4 = reinterpret_cast<float4*>(out + (int64_t)seg[k] * n_mag);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int o = ob + lane + 32 * u;
            if (o < nv) d4[o] = make_float4(scale(v[k][u].x), scale(v[k][u].y), scale(v[k][u].z), scale(v[k][u].w));
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < kFinishRows; ++k)
        if (seg[k] >= 0)
          finish_row_db(mag2 + (r0 + k) * n_mag, ref[k], out + (int64_t)seg[k] * n_mag, lane, n_bins, n_frames, power, amin,
                        top_db, cut_db, floor_db);
    }
  }
}

int launch_finish_db(const PlanImpl& p, const float* d_mag2, const float* d_rowmax, const int* d_seg_of_row,
                     int64_t n_rows, float* d_out_db, float power, float amin, float top_db, float cut_db,
                     float floor_db, cudaStream_t st) {
  int64_t blocks = ceil_div(n_rows, 8 * kFinishRows);
  const int64_t cap = (int64_t)p.sm_count * 16;
  if (blocks > cap) blocks = cap;
  finish_db_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_mag2, d_rowmax, d_seg_of_row, n_rows, p.n_bins, p.n_frames, d_out_db,
                                                      power, amin, top_db, cut_db, floor_db);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

__global__ void __launch_bounds__(256)
finish_complex_kernel(const float* __restrict__ cplx, const int64_t* __restrict__ seg_off, int n_clips, int parts,
                      int64_t n_seg, int n_bins, int n_frames, const OpLayout lay, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int n_mag = n_bins * n_frames;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < n_seg; g += warps_total) {
    const int c = find_clip(seg_off, n_clips, g);
    const int64_t row = g + (int64_t)c * (parts - 1);
    const float2* src = reinterpret_cast<const float2*>(cplx + row * 2 * n_mag);
    float2* dst = reinterpret_cast<float2*>(out + g * 2 * n_mag);
    for (int o = lane; o < n_mag; o += 32) {
      const int bin = o / n_frames, t = o - bin * n_frames;
      dst[o] = src[lay.gemm_row(bin, t, 0) >> 1];        // the GEMM's row order (OpLayout) -> [bin][t]
    }
  }
}

int launch_finish_complex(const PlanImpl& p, const float* d_cplx, const int64_t* d_seg_off, int n_clips, int64_t n_seg,
                          float* d_out, cudaStream_t st) {
  int64_t blocks = ceil_div(n_seg, 8);
  const int64_t cap = (int64_t)p.sm_count * 8;
  if (blocks > cap) blocks = cap;
  finish_complex_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_cplx, d_seg_off, n_clips, p.parts, n_seg, p.n_bins,
                                                           p.n_frames, op_layout(p), d_out);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

}  // namespace gtc
