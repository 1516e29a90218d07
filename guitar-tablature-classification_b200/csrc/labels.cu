// Label rasterisation: JAMS note events -> (n_seg, 6, 19) int8 multi-hot tablature tensors, bit-exact.
// Replaces the Python loops of /root/reference/jam_to_tablature.py:55-178 and the stats of :327-331.
//
// One warp per segment.  Lanes stride over the clip's note events, test `onset <= t < onset + dur` in fp64
// (an add and two compares -- nothing the compiler may contract into an FMA), map the pitch to (string, fret)
// with rint() (= Python's round-half-even) and OR a 114-bit mask together with warp reductions.  If no note is
// active the same warp scans the clip's pitch-contour observations (|time - t| < 0.05, conf >= 0.5).
#include "gtc_common.cuh"

namespace gtc {

constexpr int kStrings = 6;
constexpr int kFrets = 19;
constexpr int kTabBytes = kStrings * kFrets;   // 114

// jam_to_tablature.py:95-107 : lowest valid fret over the six strings (stable on ties), or -1
__device__ __forceinline__ int pitch_to_bit(double pitch) {
  const double open[kStrings] = {40.0, 45.0, 50.0, 55.0, 59.0, 64.0};
  if (!isfinite(pitch)) return -1;      // float('nan'|'inf') -> round() raises -> every string skipped
  int best_bit = -1;
  double best_fret = 1e300;
#pragma unroll
  for (int s = 0; s < kStrings; ++s) {
    double fret = rint(pitch - open[s]);                 // int(round(x)), half-to-even
    if (fret >= 0.0 && fret < (double)kFrets && fret < best_fret) {
      best_fret = fret;
      best_bit = s * kFrets + (int)fret;
    }
  }
  return best_bit;
}

__global__ void __launch_bounds__(256)
rasterize_kernel(const double* __restrict__ onset, const double* __restrict__ dur, const double* __restrict__ pitch,
                 const int64_t* __restrict__ evt_off,
                 const double* __restrict__ con_time, const double* __restrict__ con_midi,
                 const double* __restrict__ con_conf, const int8_t* __restrict__ con_kind,
                 const int64_t* __restrict__ con_off,
                 const double* __restrict__ seg_time, const int64_t* __restrict__ seg_off, int n_clips, int64_t n_seg,
                 int8_t* __restrict__ out, unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  unsigned long long n_total = 0, n_notes = 0, n_first = 0;     // lane 0 only

  for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < n_seg; g += warps_total) {
    const int c = find_clip(seg_off, n_clips, g);
    const double t = seg_time[g];
    unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;

    const int64_t e0 = evt_off[c], e1 = evt_off[c + 1];
    for (int64_t e = e0 + lane; e < e1; e += 32) {
      const double start = onset[e];
      const double end = start + dur[e];
      if (start <= t && t < end) {
        const int b = pitch_to_bit(pitch[e]);
        if (b >= 0) {
          const unsigned bit = 1u << (b & 31);
          switch (b >> 5) { case 0: m0 |= bit; break; case 1: m1 |= bit; break; case 2: m2 |= bit; break; default: m3 |= bit; }
        }
      }
    }
    m0 = __reduce_or_sync(0xffffffffu, m0);
    m1 = __reduce_or_sync(0xffffffffu, m1);
    m2 = __reduce_or_sync(0xffffffffu, m2);
    m3 = __reduce_or_sync(0xffffffffu, m3);

    if ((m0 | m1 | m2 | m3) == 0u && con_off != nullptr) {
      // extract_tablature_from_pitch_contour (jam_to_tablature.py:145-178)
      unsigned poison = 0;
      const int64_t o0 = con_off[c], o1 = con_off[c + 1];
      for (int64_t o = o0 + lane; o < o1; o += 32) {
        if (fabs(con_time[o] - t) < 0.05) {
          if (con_kind[o] == 1) { poison = 1; continue; }        // confidence None -> `conf < 0.5` raises TypeError
          const double conf = con_conf[o];
          if (conf < 0.5) continue;                               // NaN confidence is NOT skipped (nan < 0.5 is False)
          const int b = pitch_to_bit(con_midi[o]);
          if (b >= 0) {
            const unsigned bit = 1u << (b & 31);
            switch (b >> 5) { case 0: m0 |= bit; break; case 1: m1 |= bit; break; case 2: m2 |= bit; break; default: m3 |= bit; }
          }
        }
      }
      poison = __reduce_or_sync(0xffffffffu, poison);
      m0 = __reduce_or_sync(0xffffffffu, m0);
      m1 = __reduce_or_sync(0xffffffffu, m1);
      m2 = __reduce_or_sync(0xffffffffu, m2);
      m3 = __reduce_or_sync(0xffffffffu, m3);
      if (poison) { m0 = m1 = m2 = m3 = 0; }                      // process_file:319-320 swallows it, zeros stay
    }

    int8_t* dst = out + g * kTabBytes;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = lane + 32 * i;
      if (b < kTabBytes) {
        const unsigned w = i == 0 ? m0 : i == 1 ? m1 : i == 2 ? m2 : m3;
        dst[b] = (int8_t)((w >> lane) & 1u);
      }
    }
    if (lane == 0) {
      n_total += 1;
      n_notes += (m0 | m1 | m2 | m3) != 0u;
      n_first += (m0 & ((1u << kFrets) - 1u)) != 0u;             // string 0 = bits 0..18
    }
  }
  if (lane == 0 && n_total) {
    atomicAdd(stats + 0, n_total);
    atomicAdd(stats + 1, n_notes);
    atomicAdd(stats + 2, n_first);
  }
}

// my_dataloader.py:40-41 : np.argmax(annotation, axis=1)
__global__ void argmax_kernel(const int8_t* __restrict__ tabs, const int64_t* __restrict__ index, int64_t n_rows /* n*6 */,
                              int64_t* __restrict__ out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t item = r / kStrings, s = r - item * kStrings;
    const int8_t* row = tabs + ((index ? index[item] : item) * kStrings + s) * kFrets;
    int best = 0;
    int8_t bv = row[0];
#pragma unroll
    for (int f = 1; f < kFrets; ++f) {
      const int8_t v = row[f];
      if (v > bv) { bv = v; best = f; }
    }
    out[r] = best;
  }
}

// ViT_dataloader.py:28,54 + collate : heads[s][i][f] = (int64) tabs[i][s][f]
__global__ void vit_heads_kernel(const int8_t* __restrict__ tabs, const int64_t* __restrict__ index, int64_t n,
                                 int64_t* __restrict__ out) {
  const int64_t total = n * kTabBytes;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = j / (n * kFrets);
    const int64_t rem = j - s * n * kFrets;
    const int64_t i = rem / kFrets;
    const int64_t f = rem - i * kFrets;
    out[j] = (int64_t)tabs[((index ? index[i] : i) * kStrings + s) * kFrets + f];
  }
}

}  // namespace gtc

using namespace gtc;

extern "C" int gtc_rasterize_tabs(const double* d_onset, const double* d_dur, const double* d_pitch,
                                  const int64_t* d_evt_off, const double* d_con_time, const double* d_con_midi,
                                  const double* d_con_conf, const int8_t* d_con_kind, const int64_t* d_con_off,
                                  const double* d_seg_time, const int64_t* d_seg_off, int64_t n_clips, int64_t n_seg,
                                  int8_t* d_out, int64_t* d_stats, gtc_stream_t stream) {
  GTC_REQUIRE(n_clips >= 0 && n_seg >= 0, GTC_E_ARG, "gtc_rasterize_tabs: negative sizes");
  if (n_seg == 0) return GTC_OK;
  GTC_REQUIRE(n_clips > 0 && n_clips < (1 << 30), GTC_E_ARG, "gtc_rasterize_tabs: n_clips out of range");
  GTC_REQUIRE(d_evt_off && d_seg_time && d_seg_off && d_out && d_stats, GTC_E_ARG, "gtc_rasterize_tabs: null pointer");
  // contour arrays may be NULL when every clip's contour range is empty (nothing is dereferenced then)
  const int threads = 256, wpb = threads / 32;
  int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  int64_t blocks = ceil_div(n_seg, wpb);
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  rasterize_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      d_onset, d_dur, d_pitch, d_evt_off, d_con_time, d_con_midi, d_con_conf, d_con_kind, d_con_off, d_seg_time,
      d_seg_off, (int)n_clips, n_seg, d_out, reinterpret_cast<unsigned long long*>(d_stats));
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

extern "C" int gtc_labels_argmax(const int8_t* d_tabs, const int64_t* d_index, int64_t n, int64_t* d_out,
                                 gtc_stream_t stream) {
  GTC_REQUIRE(n >= 0, GTC_E_ARG, "gtc_labels_argmax: negative n");
  if (n == 0) return GTC_OK;
  GTC_REQUIRE(d_tabs && d_out, GTC_E_ARG, "gtc_labels_argmax: null pointer");
  int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  int64_t blocks = ceil_div(n * 6, 256);
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  argmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_tabs, d_index, n * 6, d_out);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

extern "C" int gtc_labels_vit_heads(const int8_t* d_tabs, const int64_t* d_index, int64_t n, int64_t* d_out,
                                    gtc_stream_t stream) {
  GTC_REQUIRE(n >= 0, GTC_E_ARG, "gtc_labels_vit_heads: negative n");
  if (n == 0) return GTC_OK;
  GTC_REQUIRE(d_tabs && d_out, GTC_E_ARG, "gtc_labels_vit_heads: null pointer");
  int sms = sm_count_of_current_device();
  if (sms <= 0) return GTC_E_CUDA;
  int64_t blocks = ceil_div(n * 114, 256);
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  vit_heads_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_tabs, d_index, n, d_out);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}
