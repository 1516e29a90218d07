// Shared helpers for libgtc.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "gtc.h"

// tcgen05 engines: bytes of K per operand row and shared-memory stage (one swizzle row), and ring depth.  128 B x 2 stages
// and 64 B x 4 stages hold the same 184 KB in flight; the finer ring refills a slot after half as many MMAs, so a TMA
// round trip has three stages of MMA work to hide under instead of one.  Measured on B200 (scripts/gemm_bench.py, same
// box, GTC_LIB_PATH A/B): GEMM + finish of an 18 900-row chunk 0.351 -> 0.319 ms, step 11.59 -> 11.41 ms.  64 is the default;
// 32 B x 8 stages (one MMA k-step per stage) is correct but TMA-request-bound: 0.515 ms.
#ifndef TC_KB_BYTES
#define TC_KB_BYTES 64
#endif
#ifndef TC_STAGES
#define TC_STAGES (256 / TC_KB_BYTES)
#endif

namespace gtc {

void set_error(const char* fmt, ...);

#define GTC_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::gtc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GTC_E_CUDA;                                                                       \
    }                                                                                          \
  } while (0)

#define GTC_REQUIRE(cond, code, ...)                                                           \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      ::gtc::set_error(__VA_ARGS__);                                                           \
      return (code);                                                                           \
    }                                                                                          \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int sm_count_of_current_device();
int patch_max_ctas();
int patch_ctas_per_sm();

// ---- programmatic dependent launch (structured CQT chain: 17-19 short kernels back to back).  A kernel launched with the
// attribute may start while its predecessor in the stream is still draining: it runs its prologue (barrier init, TMEM
// allocation, the resident operator tile) and then waits in pdl_wait() until the predecessor grid has completed and its
// writes are visible.  EVERY kernel of such a chain calls pdl_wait() before its first dependent access (completion is then
// transitive); in a kernel launched normally both calls are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// index of the clip owning global item `g`:  largest c with off[c] <= g   (off has n+1 monotone entries)
__device__ __forceinline__ int find_clip(const int64_t* __restrict__ off, int n, int64_t g) {
  int lo = 0, hi = n;              // invariant: off[lo] <= g < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= g) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __host__ __forceinline__ int halved_len(int n, int times) {
  for (int i = 0; i < times; ++i) n = (n + 1) >> 1;
  return n;
}
// frames librosa keeps (__trim_stack): min over octaves of 1 + len_i // hop_i
__device__ __host__ __forceinline__ int cqt_frames_of(int len, int hop, int n_oct) {
  int t = 0x7fffffff;
  for (int i = 0; i < n_oct; ++i) {
    const int f = 1 + len / hop;
    t = f < t ? f : t;
    if ((hop & 1) == 0) { hop >>= 1; len = (len + 1) >> 1; }
  }
  return t;
}

// ---- dB finish of one segment (cqt.py:56-58), shared by finish_db_kernel and the fused tcgen05 epilogue so that both
//      produce the same bits: |C|^2 -> |C|^power -> librosa.amplitude_to_db(ref=np.amax, amin, top_db) -> cqt_lim.
struct FinishArgs {
  float* out_db;             // [n_seg][n_bins][n_frames]; null = do not fuse (the caller runs finish_db_kernel)
  const int64_t* seg_off;    // [n_clips + 1]
  const int* seg_of_row;     // [n_rows_pad] segment started by each operand row, -1 for a clip's last P-1 rows and padding (frame_kernel)
  int* tile_done;            // [n_rows_pad / 128] counters, zero before the launch (frame_kernel) and after it
  int n_clips, parts, n_bins, n_frames;
  int64_t n_rows, n_seg;
  float power, amin, top_db, cut_db, floor_db;
};

__device__ __forceinline__ float mag_power(float m2, float power) {
  // |C|^power from |C|^2 : power 4 -> m2*m2 ; power 2 -> m2 ; power 1 -> sqrt(m2) ; else powf
  if (power == 4.f) return m2 * m2;
  if (power == 2.f) return m2;
  if (power == 1.f) return sqrtf(m2);
  return powf(m2, 0.5f * power);
}

// 10 * log10(x) through MUFU.LG2 (__log2f: abs. error 2^-22 of the logarithm, i.e. < 2e-6 dB against the 0.01 dB gate).
// The dB passes were bound by the ~30-instruction log10f subroutine: finish_db_kernel 44 -> 19 us per 28 200-row chunk on
// B200 (same box, profiles/r02e_gemm_finish_ab.md).  -DGTC_EXACT_LOG restores log10f.
__device__ __forceinline__ float ten_log10(float x) {
#ifdef GTC_EXACT_LOG
  return 10.f * log10f(x);
#else
  // __fmul_rn: the product must be rounded on its own -- contracted into `fma(c, log2 x, -ref_db)` the segment's peak element
  // would come out as the rounding error of c * log2(ref) instead of exactly 0 dB
  return __fmul_rn(3.010299956639812f, __log2f(x));
#endif
}

// dB value of one element given the segment's reference (both as |C|^2)
struct DbScale {
  float power, amin2, ref_db, lo_clamp, cut_db, floor_db;
  __device__ __forceinline__ DbScale(float m2max, float power_, float amin, float top_db, float cut_db_, float floor_db_)
      : power(power_), amin2(amin * amin), cut_db(cut_db_), floor_db(floor_db_) {
    // S = |C|^power ; amplitude_to_db squares it again: 10*log10(max(amin^2, S^2)) - 10*log10(max(amin^2, ref^2))
    const float ref = mag_power(m2max, power);
    ref_db = ten_log10(fmaxf(amin2, ref * ref));
    lo_clamp = 0.f - top_db;         // log_spec.max() is the peak element's own value -> exactly 0 dB
  }
  __device__ __forceinline__ float operator()(float m2) const {
    const float s = mag_power(m2, power);
    float db = ten_log10(fmaxf(amin2, s * s)) - ref_db;
    db = fmaxf(db, lo_clamp);
    return db < cut_db ? floor_db : db;
  }
};

// one warp: src = the segment's |C|^2, dst = its dB features, both [bin][t] (the GEMM's operator rows are ordered so that
// its |C|^2 output already has the final layout: OpLayout below), 16-byte vectors, all loads issued before the first use.
template <int U = 4>
__device__ __forceinline__ void finish_row_db(const float* src, float m2max, float* dst, int lane, int n_bins, int n_frames,
                                              float power, float amin, float top_db, float cut_db, float floor_db) {
  const int n_mag = n_bins * n_frames;
  const DbScale scale(m2max, power, amin, top_db, cut_db, floor_db);
  if ((n_mag & 3) == 0) {
    const int nv = n_mag >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int o0 = lane; o0 < nv; o0 += 32 * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (o0 + 32 * u < nv) ? __ldcg(s4 + o0 + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (o0 + 32 * u < nv) d4[o0 + 32 * u] = make_float4(scale(v[u].x), scale(v[u].y), scale(v[u].z), scale(v[u].w));
    }
  } else {
    for (int o = lane; o < n_mag; o += 32) dst[o] = scale(__ldcg(src + o));
  }
}

// ---- operator row order.  The host hands the operator over with rows (t * n_bins + bin) * 2 + {re, im} (gtc.h); inside
// the library the rows are re-ordered so that (a) the GEMM's |C|^2 output column of (bin, t) is bin * T + t, the final
// [bin][t] feature layout (cqt.py:58 saves (n_bins, T)), and (b) where the tile width allows it, an N tile holds
// `bins_per_tile` whole bins ordered FRAME-major inside the tile:
//     gemm row = tile * nc + t * (2 * bins_per_tile) + bin_in_tile * 2 + c,     tile = bin / bins_per_tile
// so that the rows of one frame are one contiguous column range of the accumulator -- a frame's operator rows are non-zero
// only on a window of samples around the frame, and the tensor-core engine multiplies a k-block only with the frames
// whose window reaches it (cqt_gemm_tc.cu).  bins_per_tile == 0: plain order, gemm row = (bin * T + t) * 2 + c.
struct OpLayout {
  int nc, bins_per_tile, n_bins, n_frames;
  __host__ __device__ __forceinline__ int gemm_row(int bin, int t, int c) const {
    if (bins_per_tile == 0) return (bin * n_frames + t) * 2 + c;
    const int tile = bin / bins_per_tile, bl = bin - tile * bins_per_tile;
    return tile * nc + t * 2 * bins_per_tile + bl * 2 + c;
  }
  // final |C|^2 column (bin * T + t) of the pair of gemm rows (n, n + 1), n even
  __host__ __device__ __forceinline__ int mag_col(int n) const {
    if (bins_per_tile == 0) return n >> 1;
    const int tile = n / nc, j = n - tile * nc;
    const int t = j / (2 * bins_per_tile), bl = (j - t * 2 * bins_per_tile) >> 1;
    return (tile * bins_per_tile + bl) * n_frames + t;
  }
};

// ---- "slotted" M operand of the tensor-core engine (structured CQT, cqt_structured.cu).  The rows of the GEMM are windows
// of per-segment signals stored as fp16 hi/lo planes: segment s owns `stride` samples starting at base + s * stride, row j
// of a segment is the window starting at j * row_step samples (windows overlap: the TMA tensor map simply has a row stride
// smaller than the row length).  A 128-row block is 16 segments x 8 consecutive rows (3-D TMA box) -- or 32 x 4 / 64 x 2 for
// the short low octaves -- so short and ragged segments waste at most 7 (3, 1) rows each.  slot_mode: 1 = 2:1 decimator (output: the next octave's hi/lo planes, samples
// beyond the segment's length zeroed), 2 = octave response (output: |C|^2 into [seg][bin][t] + the segment maximum).
struct SlotArgs {
  int slot_mode;            // 0 = off
  int jgroups;              // row groups per segment
  int box_log2;             // a 128-row block is (128 >> box_log2) segments x (1 << box_log2) consecutive rows: 3 (16 x 8), 2 (32 x 4) or 1 (64 x 2)
  int64_t n_slots;          // segments
  // mode 1
  __half* out_hi;
  __half* out_lo;
  int64_t out_stride;       // samples per segment of the output planes
  int64_t out_base;         // first sample of segment 0 in the output planes
  const int32_t* seg_len;   // [n_slots] full-rate lengths
  const int32_t* seg_frames;   // [n_slots] frames librosa keeps for each segment (mode 2)
  int stage_out;            // the output is octave `stage_out`: its valid length is seg_len halved stage_out times
  float plane_scale;        // power-of-two scale of the planes (the fp16x2 engine's x_scale)
  // mode 2
  float* out;               // [n_slots][n_bins][t_max] |C|^2
  float* segmax;            // [n_slots]
  int n_bins, bin_lo, bin_cnt, t_max, hop0, n_oct;
};

constexpr int kMaxBand = 48;   // k-blocks of a decimator operator that can carry a band schedule (PlanImpl, TcParams)

// ---- segment-operator plan (cqt_api.cu) -------------------------------------------------------------------
struct PlanImpl {
  int device;
  int seg_len, seg_hop, n_bins, n_frames;
  int parts;        // P: audio rows per segment
  int row_len;      // samples per audio row (seg_hop when P>1 or seg_len when P==1)
  int elem_bytes;   // operand element size: 4 (fp32 containers of the tf32 / SIMT engines) or 2 (fp16x2 engine)
  int kb_elems;     // elements per 128-byte k-block: 32 (fp32) or 64 (fp16)
  int kp;           // row_len rounded up to kb_elems (one 128-byte swizzle atom)
  int k_total;      // P * kp
  int n_out;        // 2 * n_bins * n_frames  (real operator rows)
  int n_pad;        // n_out rounded up to 128
  int nc;           // N tile width of the tensor-core engines (operator rows per tile)
  int bins_per_tile;   // > 0: tiles hold whole bins, frame-major inside the tile (OpLayout); 0: plain bin-major rows
  int engine;
  int sm_count;
  int tc_max_ctas;      // persistent GEMM grid limit (0 = sm_count)
  int tc_kb_per_split;  // K blocks per tensor-core accumulation split (0 = default), env GTC_TC_KSPLIT
  int tc_fuse_finish;   // 1: the tcgen05 epilogue also does the dB finish (GTC_OPT_FUSE_FINISH; default 1)
  float x_scale;    // fp16x2 engine: audio is multiplied by this power of two before the hi/lo split (1 otherwise)
  float out_scale;  // epilogue factor undoing x_scale and the operator's power-of-two scale (1 otherwise)
  float* d_op;      // [n_pad][k_total]  operator, fp32, K padded per part (SIMT engine only)
  void* d_op_hi;    // [n_pad][k_total]  high part: tf32-representable fp32 (RN) or fp16
  void* d_op_lo;    // [n_pad][k_total]  residual of the high part, same element type
  void* tmap_op_hi; // CUtensorMap storage (128 B each), tcgen05 engines only
  void* tmap_op_lo;
  // ---- zero-skipping schedule of the tensor-core engines (cqt_gemm_tc.cu): per N tile the list of k-blocks that hold any
  //      non-zero operator entry, each with the contiguous range of row groups (frames of a frame-major tile) it touches
  int grp_rows;            // operator rows per group (2 * bins_per_tile, or 16 for plain tiles)
  int sched_pitch;         // entries per N tile in d_sched
  int sched_ksplit;        // k-blocks per accumulation split the schedule was built for (its first entry per split is dense)
  uint32_t* d_sched;       // [n_chunks][sched_pitch]  kb | g0 << 16 | ng << 24 ; null = dense
  int* d_sched_len;        // [n_chunks]
  void* d_op_maps;         // CUtensorMap[n_groups][2] in device memory: operator boxes of (g + 1) * grp_rows rows, hi / lo
  float sched_fill;        // scheduled tensor work / dense tensor work (1 = nothing to skip)
  int sched_rotate;        // 1: the N tile index is rotated by the CTA's iteration (balances tile types); 0: experiments
  // ---- resident operator (slotted decimator GEMM): > 0 = the operator is banded Toeplitz, Op[n][i + kb_elems] == Op[n - res_shift][i],
  //      so one master tile in shared memory serves every k-block (cqt_gemm_tc.cu, RingRes); set by cqt_structured.cu after it has
  //      verified the property on the operator it built
  int res_shift;
  // ---- band schedule of the same operator: entry e = k-block band_kb[e] x the band_ng[e] groups of 16 operator rows from group
  //      band_g0[e] on that hold a non-zero on it; entry 0 is a full k-block (zero-initialises the accumulator).  0 = dense.
  int band_n;
  uint8_t band_kb[kMaxBand], band_g0[kMaxBand], band_ng[kMaxBand];
};

// launchers (each enqueues on `st`, returns a GTC_* code); xhi/xlo element type follows PlanImpl::elem_bytes
int launch_frame(const PlanImpl& p, const void* d_audio, int pcm16, const int64_t* d_clip_off, const int64_t* d_seg_off,
                 int n_clips, int64_t n_rows, int64_t n_rows_alloc, void* d_xhi, void* d_xlo, float* d_rowmax,
                 int* d_tile_done, int* d_seg_of_row, cudaStream_t st);
inline OpLayout op_layout(const PlanImpl& p) { return OpLayout{p.nc, p.bins_per_tile, p.n_bins, p.n_frames}; }
int launch_gemm_simt(const PlanImpl& p, const float* d_xhi, const float* d_xlo, int64_t n_rows_pad,
                     float* d_mag2, float* d_cplx, float* d_rowmax, cudaStream_t st);
int launch_gemm_tc(const PlanImpl& p, const void* d_xhi, const void* d_xlo, int64_t n_rows_pad, int64_t n_rows_alloc,
                   float* d_mag2, float* d_cplx, float* d_rowmax, const FinishArgs& fin, cudaStream_t st);
// structured CQT: GEMM of slotted rows (3-D tensor maps over the hi/lo planes) against plan `p`'s operator
int launch_gemm_tc_slots(const PlanImpl& p, const __half* x_hi, const __half* x_lo, int64_t x_first, int64_t x_stride,
                         int64_t row_step, int rows_per_slot, const SlotArgs& slots, cudaStream_t st);
int tc_slots_init();
int tc_plan_init(PlanImpl& p, const uint8_t* h_nz /* [n_pad / grp_rows][k_blocks] non-zero flags, or null */, int kb_per_split);
int tc_pick_plain_width(int n_out);                       // tile width of the plain row order
bool tc_has_frame_major_kernel(int nc, int n_frames);     // is gemm_tc_kernel instantiated for this frame-major tile?
void tc_plan_free(PlanImpl& p);
int launch_finish_db(const PlanImpl& p, const float* d_mag2, const float* d_rowmax, const int* d_seg_of_row,
                     int64_t n_rows, float* d_out_db, float power, float amin, float top_db, float cut_db,
                     float floor_db, cudaStream_t st);
int launch_finish_complex(const PlanImpl& p, const float* d_cplx, const int64_t* d_seg_off, int n_clips, int64_t n_seg,
                          float* d_out, cudaStream_t st);

}  // namespace gtc
