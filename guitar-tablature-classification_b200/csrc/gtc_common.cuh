// Shared helpers for libgtc.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "gtc.h"

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

// index of the clip owning global item `g`:  largest c with off[c] <= g   (off has n+1 monotone entries)
__device__ __forceinline__ int find_clip(const int64_t* __restrict__ off, int n, int64_t g) {
  int lo = 0, hi = n;              // invariant: off[lo] <= g < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= g) lo = mid; else hi = mid;
  }
  return lo;
}

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
  int engine;
  int sm_count;
  int tc_max_ctas;      // persistent GEMM grid limit (0 = sm_count)
  int tc_kb_per_split;  // K blocks per tensor-core accumulation split (0 = default), env GTC_TC_KSPLIT
  float x_scale;    // fp16x2 engine: audio is multiplied by this power of two before the hi/lo split (1 otherwise)
  float out_scale;  // epilogue factor undoing x_scale and the operator's power-of-two scale (1 otherwise)
  float* d_op;      // [n_pad][k_total]  operator, fp32, K padded per part (SIMT engine only)
  void* d_op_hi;    // [n_pad][k_total]  high part: tf32-representable fp32 (RN) or fp16
  void* d_op_lo;    // [n_pad][k_total]  residual of the high part, same element type
  void* tmap_op_hi; // CUtensorMap storage (128 B each), tcgen05 engines only
  void* tmap_op_lo;
};

// launchers (each enqueues on `st`, returns a GTC_* code); xhi/xlo element type follows PlanImpl::elem_bytes
int launch_frame(const PlanImpl& p, const void* d_audio, int pcm16, const int64_t* d_clip_off, const int64_t* d_seg_off,
                 int n_clips, int64_t n_rows, int64_t n_rows_alloc, void* d_xhi, void* d_xlo, float* d_rowmax,
                 cudaStream_t st);
int launch_gemm_simt(const PlanImpl& p, const float* d_xhi, const float* d_xlo, int64_t n_rows_pad,
                     float* d_mag2, float* d_cplx, float* d_rowmax, cudaStream_t st);
int launch_gemm_tc(const PlanImpl& p, const void* d_xhi, const void* d_xlo, int64_t n_rows_pad, int64_t n_rows_alloc,
                   float* d_mag2, float* d_cplx, float* d_rowmax, cudaStream_t st);
int tc_plan_init(PlanImpl& p);
void tc_plan_free(PlanImpl& p);
int launch_finish_db(const PlanImpl& p, const float* d_mag2, const float* d_rowmax, const int64_t* d_seg_off,
                     int n_clips, int64_t n_seg, float* d_out_db, float power, float amin, float top_db, float cut_db,
                     float floor_db, cudaStream_t st);
int launch_finish_complex(const PlanImpl& p, const float* d_cplx, const int64_t* d_seg_off, int n_clips, int64_t n_seg,
                          float* d_out, cudaStream_t st);

}  // namespace gtc
