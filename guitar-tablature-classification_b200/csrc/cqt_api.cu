// C-ABI entry points of the segment CQT (plan, workspace, frame -> GEMM -> finish) and library-wide helpers.
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>
#include <new>
#include <vector>
#include <cuda_fp16.h>
#include "gtc_common.cuh"

namespace gtc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count_of_current_device() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("no usable CUDA device (libgtc has no CPU fallback): %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return sms;
}

static inline float tf32_rn_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x00000fffu + ((u >> 13) & 1u)) & 0xffffe000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

struct Workspace {
  int64_t n_rows, n_rows_pad, n_rows_alloc;
  size_t off_xhi, off_xlo, off_out, off_rowmax, off_done, off_segrow, total;
};

static Workspace workspace_layout(const PlanImpl& p, int64_t n_seg, int64_t n_clips, bool complex_out) {
  Workspace w;
  w.n_rows = n_seg + n_clips * (p.parts - 1);
  w.n_rows_pad = round_up(w.n_rows > 0 ? w.n_rows : 1, 128);
  w.n_rows_alloc = w.n_rows_pad + 8;                    // rows read with the +p offset of the last tile
  const size_t xbytes = (size_t)w.n_rows_alloc * p.kp * (size_t)p.elem_bytes;
  const size_t obytes = (size_t)w.n_rows_pad * (complex_out ? p.n_out : p.n_out / 2) * sizeof(float);
  size_t o = 0;
  auto take = [&](size_t b) { size_t at = o; o += (b + 1023) & ~(size_t)1023; return at; };
  w.off_xhi = take(xbytes);
  w.off_xlo = take(xbytes);
  w.off_out = take(obytes);
  w.off_rowmax = take((size_t)w.n_rows_pad * sizeof(float));
  w.off_done = take((size_t)(w.n_rows_pad / 128) * sizeof(int));      // row-block counters of the fused dB finish
  w.off_segrow = take((size_t)w.n_rows_pad * sizeof(int));            // operand row -> segment it starts (or -1)
  w.total = o;
  return w;
}

}  // namespace gtc

using namespace gtc;

struct gtc_plan {
  PlanImpl impl;
};

extern "C" int gtc_version(void) { return GTC_VERSION; }
extern "C" const char* gtc_last_error(void) { return g_err; }

extern "C" int gtc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  cudaDeviceProp prop;
  GTC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return GTC_OK;
}

extern "C" int gtc_cqt_plan_create(gtc_plan** out, int device, int seg_len, int seg_hop, int n_bins, int n_frames,
                                   const float* h_operator, int gemm_engine) {
  GTC_REQUIRE(out != nullptr, GTC_E_ARG, "gtc_cqt_plan_create: out is NULL");
  *out = nullptr;
  GTC_REQUIRE(h_operator != nullptr, GTC_E_ARG, "gtc_cqt_plan_create: h_operator is NULL (design it with gtc_b200.cqt_design)");
  GTC_REQUIRE(seg_len > 0 && seg_hop > 0 && n_bins > 0 && n_frames > 0, GTC_E_ARG, "gtc_cqt_plan_create: non-positive size");
  GTC_REQUIRE(gemm_engine == GTC_GEMM_TCGEN05_3XTF32 || gemm_engine == GTC_GEMM_SIMT_FP32 || gemm_engine == GTC_GEMM_TCGEN05_FP16X2, GTC_E_ARG,
              "gtc_cqt_plan_create: unknown engine %d", gemm_engine);
  GTC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GTC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  GTC_REQUIRE(prop.major == 10, GTC_E_UNSUP, "libgtc is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
              prop.minor);

  gtc_plan* plan = new (std::nothrow) gtc_plan();
  GTC_REQUIRE(plan != nullptr, GTC_E_NOMEM, "gtc_cqt_plan_create: out of host memory");
  PlanImpl& p = plan->impl;
  memset(&p, 0, sizeof(p));
  p.device = device;
  p.seg_len = seg_len; p.seg_hop = seg_hop; p.n_bins = n_bins; p.n_frames = n_frames;
  const bool divides = (seg_len % seg_hop == 0) && (seg_len / seg_hop <= 8);
  p.parts = divides ? seg_len / seg_hop : 1;
  p.row_len = divides ? seg_hop : seg_len;
  const bool half = gemm_engine == GTC_GEMM_TCGEN05_FP16X2;
  const bool tensor = gemm_engine != GTC_GEMM_SIMT_FP32;
  p.elem_bytes = half ? 2 : 4;
  p.kb_elems = (tensor ? TC_KB_BYTES : 128) / p.elem_bytes;
  p.x_scale = half ? 256.f : 1.f;                       // |x| < 256 stays finite in fp16; lo part normal down to |x| ~ 1e-3
  p.out_scale = 1.f;
  p.kp = (int)round_up(p.row_len, p.kb_elems);
  p.k_total = p.parts * p.kp;
  p.n_out = 2 * n_bins * n_frames;
  p.engine = gemm_engine;
  // N tile width and operator row order (OpLayout, gtc_common.cuh): the widest tile of whole bins x all frames that the
  // tensor core takes (N <= 256, N % 16 == 0) and that tiles the bins exactly; else plain bin-major rows.
  p.bins_per_tile = 0;
  p.nc = tc_pick_plain_width(p.n_out);
  for (int b = 256 / (2 * n_frames); b >= 1 && tensor; --b)
    if ((2 * n_frames * b) % 16 == 0 && n_bins % b == 0 && tc_has_frame_major_kernel(2 * n_frames * b, n_frames)) {
      p.bins_per_tile = b;
      p.nc = 2 * n_frames * b;
      break;
    }
  p.n_pad = (int)round_up(round_up(p.n_out, p.nc), 128);      // whole tiles, and the SIMT engine's 128-row blocks
  p.sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("GTC_TC_KSPLIT")) p.tc_kb_per_split = atoi(e);
  p.tc_fuse_finish = 0;                                 // 1 = dB finish inside the GEMM epilogue (measured slower, cqt_gemm_tc.cu)
  if (const char* e = getenv("GTC_FUSE_FINISH")) p.tc_fuse_finish = atoi(e) != 0;
  if (tensor && (p.n_out % 16 != 0)) {
    delete plan;
    set_error("gtc_cqt_plan_create: tcgen05 engine needs 2*n_bins*n_frames %% 16 == 0 (got %d)", p.n_out);
    return GTC_E_UNSUP;
  }

  const size_t elems = (size_t)p.n_pad * p.k_total;
  int rc = GTC_OK;
  auto up = [&](void** d, const void* h, size_t bytes) -> int {
    GTC_CUDA_CHECK(cudaMalloc(d, bytes));
    GTC_CUDA_CHECK(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
    return GTC_OK;
  };
  auto at_of = [&](int j) { const int part = j / p.row_len; return (size_t)part * p.kp + (j - part * p.row_len); };
  // host row (t * n_bins + bin) * 2 + c  ->  library row (OpLayout)
  const OpLayout lay = op_layout(p);
  std::vector<int> lib_row(p.n_out);
  for (int t = 0; t < n_frames; ++t)
    for (int b = 0; b < n_bins; ++b)
      for (int c = 0; c < 2; ++c) lib_row[(t * n_bins + b) * 2 + c] = lay.gemm_row(b, t, c);
  // non-zero map of the operator at the tensor engine's granularity (row group x k-block), for its zero-skipping schedule
  p.grp_rows = p.bins_per_tile > 0 ? 2 * p.bins_per_tile : 16;
  const int n_kblocks = p.k_total / p.kb_elems;
  std::vector<uint8_t> nz;
  if (tensor && p.nc % p.grp_rows == 0) {
    nz.assign((size_t)(p.n_pad / p.grp_rows) * n_kblocks, 0);
    for (int n = 0; n < p.n_out; ++n) {
      uint8_t* row = &nz[(size_t)(lib_row[n] / p.grp_rows) * n_kblocks];
      const float* src = h_operator + (size_t)n * seg_len;
      for (int j = 0; j < seg_len; ++j)
        if (src[j] != 0.f) row[at_of(j) / p.kb_elems] = 1;
    }
  }
  if (half) {
    // fp16x2: A * 2^s = hi + lo with s such that max|A| * 2^s <= 2^13; absolute quantisation error <= 2^-25 (fp16
    // subnormal spacing / 2) against row maxima of ~2^13, i.e. < 2^-36 relative -- see DESIGN.md 3.1
    float amax = 0.f;
    for (size_t i = 0; i < (size_t)p.n_out * seg_len; ++i) amax = fmaxf(amax, fabsf(h_operator[i]));
    int s = amax > 0.f ? 13 - (int)ceilf(log2f(amax)) : 0;
    if (s > 24) s = 24;
    if (s < -24) s = -24;
    const float a_scale = ldexpf(1.f, s);
    p.out_scale = 1.f / (a_scale * p.x_scale);
    std::vector<__half> hi(elems, __float2half(0.f)), lo(elems, __float2half(0.f));
    for (int n = 0; n < p.n_out; ++n)
      for (int j = 0; j < seg_len; ++j) {
        const float v = h_operator[(size_t)n * seg_len + j] * a_scale;
        const size_t at = (size_t)lib_row[n] * p.k_total + at_of(j);
        hi[at] = __float2half_rn(v);
        lo[at] = __float2half_rn(v - __half2float(hi[at]));
      }
    if ((rc = up(&p.d_op_hi, hi.data(), elems * 2)) != GTC_OK || (rc = up(&p.d_op_lo, lo.data(), elems * 2)) != GTC_OK) {
      gtc_cqt_plan_destroy(plan);
      return rc;
    }
  } else {
    std::vector<float> raw(elems, 0.f), hi(elems, 0.f), lo(elems, 0.f);
    for (int n = 0; n < p.n_out; ++n)
      for (int j = 0; j < seg_len; ++j) {
        const float v = h_operator[(size_t)n * seg_len + j];
        const size_t at = (size_t)lib_row[n] * p.k_total + at_of(j);
        raw[at] = v;
        hi[at] = tf32_rn_host(v);
        lo[at] = v - hi[at];
      }
    if ((!tensor && (rc = up((void**)&p.d_op, raw.data(), elems * 4)) != GTC_OK) ||
        (tensor && ((rc = up(&p.d_op_hi, hi.data(), elems * 4)) != GTC_OK || (rc = up(&p.d_op_lo, lo.data(), elems * 4)) != GTC_OK))) {
      gtc_cqt_plan_destroy(plan);
      return rc;
    }
  }
  const int ksplit = (p.tc_kb_per_split > 0 ? p.tc_kb_per_split : 8) * (128 / TC_KB_BYTES);
  if (tensor && (rc = tc_plan_init(p, nz.empty() || getenv("GTC_TC_DENSE") ? nullptr : nz.data(), ksplit)) != GTC_OK) {
    gtc_cqt_plan_destroy(plan);
    return rc;
  }
  *out = plan;
  return GTC_OK;
}

extern "C" int gtc_cqt_plan_destroy(gtc_plan* plan) {
  if (!plan) return GTC_OK;
  PlanImpl& p = plan->impl;
  tc_plan_free(p);
  if (p.d_op) cudaFree(p.d_op);
  if (p.d_op_hi) cudaFree(p.d_op_hi);
  if (p.d_op_lo) cudaFree(p.d_op_lo);
  delete plan;
  return GTC_OK;
}

extern "C" int gtc_cqt_plan_configure(gtc_plan* plan, int option, int value) {
  GTC_REQUIRE(plan != nullptr && value >= 0, GTC_E_ARG, "gtc_cqt_plan_configure: bad argument");
  switch (option) {
    case GTC_OPT_TC_KSPLIT: plan->impl.tc_kb_per_split = value; return GTC_OK;
    case GTC_OPT_GEMM_MAX_CTAS: plan->impl.tc_max_ctas = value; return GTC_OK;
    case GTC_OPT_FUSE_FINISH: plan->impl.tc_fuse_finish = value != 0; return GTC_OK;
    default: set_error("gtc_cqt_plan_configure: unknown option %d", option); return GTC_E_ARG;
  }
}

extern "C" int gtc_cqt_plan_parts(const gtc_plan* plan) { return plan ? plan->impl.parts : GTC_E_ARG; }

extern "C" int gtc_cqt_workspace_bytes(const gtc_plan* plan, int64_t n_seg, int64_t n_clips, size_t* bytes) {
  GTC_REQUIRE(plan && bytes && n_seg >= 0 && n_clips >= 0, GTC_E_ARG, "gtc_cqt_workspace_bytes: bad argument");
  *bytes = workspace_layout(plan->impl, n_seg, n_clips, true).total;     // complex layout is the larger one
  return GTC_OK;
}

static int run_segments(const gtc_plan* plan, const void* d_audio, const int64_t* d_clip_off, const int64_t* d_seg_off,
                        int64_t n_clips, int64_t n_seg, float* d_out, bool complex_out, void* d_workspace,
                        size_t workspace_bytes, float power, float amin, float top_db, float cut_db, float floor_db,
                        cudaStream_t st, int stages = 3 /* bit 0: framing, bit 1: contraction + finish */, int pcm16 = 0) {
  GTC_REQUIRE(plan != nullptr, GTC_E_ARG, "gtc_cqt_segments: plan is NULL");
  GTC_REQUIRE(n_clips >= 0 && n_seg >= 0, GTC_E_ARG, "gtc_cqt_segments: negative sizes");
  if (n_seg == 0) return GTC_OK;
  GTC_REQUIRE(n_clips > 0 && n_clips < (1 << 30), GTC_E_ARG, "gtc_cqt_segments: n_clips out of range");
  GTC_REQUIRE(d_clip_off && d_seg_off && d_workspace && ((stages & 1) == 0 || d_audio) && ((stages & 2) == 0 || d_out),
              GTC_E_ARG, "gtc_cqt_segments: null pointer");
  const PlanImpl& p = plan->impl;
  const Workspace w = workspace_layout(p, n_seg, n_clips, complex_out);
  GTC_REQUIRE(workspace_bytes >= w.total, GTC_E_NOMEM, "gtc_cqt_segments: workspace of %zu bytes, %zu needed",
              workspace_bytes, w.total);
  // TMA needs 16-byte aligned global tiles; 256 is what cudaMalloc / torch's caching allocator (512) guarantee
  GTC_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, GTC_E_ARG, "gtc_cqt_segments: workspace must be 256-byte aligned");
  int dev = -1;
  GTC_CUDA_CHECK(cudaGetDevice(&dev));
  GTC_REQUIRE(dev == p.device, GTC_E_ARG, "gtc_cqt_segments: plan belongs to device %d, current device is %d", p.device, dev);
  char* ws = static_cast<char*>(d_workspace);
  void* xhi = ws + w.off_xhi;
  void* xlo = ws + w.off_xlo;
  float* gout = reinterpret_cast<float*>(ws + w.off_out);
  float* rowmax = reinterpret_cast<float*>(ws + w.off_rowmax);
  int* tile_done = reinterpret_cast<int*>(ws + w.off_done);
  int* seg_of_row = reinterpret_cast<int*>(ws + w.off_segrow);
  GTC_REQUIRE(n_seg < ((int64_t)1 << 31), GTC_E_UNSUP, "gtc_cqt_segments: more than 2^31 segments in one call");
  int rc = GTC_OK;
  if (stages & 1) rc = launch_frame(p, d_audio, pcm16, d_clip_off, d_seg_off, (int)n_clips, w.n_rows, w.n_rows_alloc, xhi, xlo, rowmax, tile_done, seg_of_row, st);
  if (rc != GTC_OK || (stages & 2) == 0) return rc;
  float* mag2 = complex_out ? nullptr : gout;
  float* cplx = complex_out ? gout : nullptr;
  FinishArgs fin;
  memset(&fin, 0, sizeof(fin));
  const bool fused = p.engine != GTC_GEMM_SIMT_FP32 && !complex_out && p.tc_fuse_finish;
  if (fused) {
    fin.out_db = d_out; fin.seg_off = d_seg_off; fin.tile_done = tile_done; fin.seg_of_row = seg_of_row;
    fin.n_clips = (int)n_clips; fin.parts = p.parts; fin.n_bins = p.n_bins; fin.n_frames = p.n_frames;
    fin.n_rows = w.n_rows; fin.n_seg = n_seg;
    fin.power = power; fin.amin = amin; fin.top_db = top_db; fin.cut_db = cut_db; fin.floor_db = floor_db;
  }
  if (p.engine != GTC_GEMM_SIMT_FP32)
    rc = launch_gemm_tc(p, xhi, xlo, w.n_rows_pad, w.n_rows_alloc, mag2, cplx, rowmax, fin, st);
  else
    rc = launch_gemm_simt(p, (const float*)xhi, (const float*)xlo, w.n_rows_pad, mag2, cplx, rowmax, st);
  if (rc != GTC_OK || fused) return rc;
  if (complex_out) return launch_finish_complex(p, cplx, d_seg_off, (int)n_clips, n_seg, d_out, st);
  return launch_finish_db(p, mag2, rowmax, seg_of_row, w.n_rows, d_out, power, amin, top_db, cut_db, floor_db, st);
}

extern "C" int gtc_cqt_segments_db(const gtc_plan* plan, const float* d_audio, const int64_t* d_clip_off,
                                   const int64_t* d_seg_off, int64_t n_clips, int64_t n_seg, float* d_out_db,
                                   void* d_workspace, size_t workspace_bytes, float power, float amin, float top_db,
                                   float cut_db, float floor_db, gtc_stream_t stream) {
  GTC_REQUIRE(power > 0.f && amin > 0.f, GTC_E_ARG, "gtc_cqt_segments_db: power and amin must be positive");
  return run_segments(plan, d_audio, d_clip_off, d_seg_off, n_clips, n_seg, d_out_db, false, d_workspace, workspace_bytes,
                      power, amin, top_db, cut_db, floor_db, (cudaStream_t)stream);
}

extern "C" int gtc_cqt_frame(const gtc_plan* plan, const float* d_audio, const int64_t* d_clip_off, const int64_t* d_seg_off,
                             int64_t n_clips, int64_t n_seg, void* d_workspace, size_t workspace_bytes, gtc_stream_t stream) {
  return run_segments(plan, d_audio, d_clip_off, d_seg_off, n_clips, n_seg, nullptr, false, d_workspace, workspace_bytes,
                      4.f, 1e-5f, 80.f, -60.f, -120.f, (cudaStream_t)stream, 1);
}

extern "C" int gtc_cqt_frame_pcm16(const gtc_plan* plan, const int16_t* d_pcm, const int64_t* d_clip_off, const int64_t* d_seg_off,
                                   int64_t n_clips, int64_t n_seg, void* d_workspace, size_t workspace_bytes, gtc_stream_t stream) {
  return run_segments(plan, d_pcm, d_clip_off, d_seg_off, n_clips, n_seg, nullptr, false, d_workspace, workspace_bytes,
                      4.f, 1e-5f, 80.f, -60.f, -120.f, (cudaStream_t)stream, 1, 1);
}

extern "C" int gtc_cqt_contract_db(const gtc_plan* plan, const int64_t* d_clip_off, const int64_t* d_seg_off, int64_t n_clips,
                                   int64_t n_seg, float* d_out_db, void* d_workspace, size_t workspace_bytes, float power,
                                   float amin, float top_db, float cut_db, float floor_db, gtc_stream_t stream) {
  GTC_REQUIRE(power > 0.f && amin > 0.f, GTC_E_ARG, "gtc_cqt_contract_db: power and amin must be positive");
  return run_segments(plan, nullptr, d_clip_off, d_seg_off, n_clips, n_seg, d_out_db, false, d_workspace, workspace_bytes,
                      power, amin, top_db, cut_db, floor_db, (cudaStream_t)stream, 2);
}

static int g_patch_max_ctas = 0;
static int g_patch_ctas_per_sm = 2;
namespace gtc { int patch_max_ctas() { return g_patch_max_ctas; } int patch_ctas_per_sm() { return g_patch_ctas_per_sm; } }

extern "C" int gtc_set_option(int option, int value) {
  GTC_REQUIRE(value >= 0, GTC_E_ARG, "gtc_set_option: negative value");
  switch (option) {
    case GTC_OPT_PATCH_MAX_CTAS: g_patch_max_ctas = value; return GTC_OK;
    case GTC_OPT_PATCH_CTAS_PER_SM: g_patch_ctas_per_sm = value > 0 ? value : 2; return GTC_OK;
    default: set_error("gtc_set_option: unknown option %d", option); return GTC_E_ARG;
  }
}

extern "C" int gtc_cqt_segments_complex(const gtc_plan* plan, const float* d_audio, const int64_t* d_clip_off,
                                        const int64_t* d_seg_off, int64_t n_clips, int64_t n_seg, float* d_out_c,
                                        void* d_workspace, size_t workspace_bytes, gtc_stream_t stream) {
  return run_segments(plan, d_audio, d_clip_off, d_seg_off, n_clips, n_seg, d_out_c, true, d_workspace, workspace_bytes,
                      4.f, 1e-5f, 80.f, -60.f, -120.f, (cudaStream_t)stream);
}
