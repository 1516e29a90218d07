// Tensor-core engines for the segment-operator contraction, sm_100a only:
//   GTC_GEMM_TCGEN05_FP16X2 (default)  operands split into fp16 hi + fp16 lo, tcgen05.mma.kind::f16, K = 16 per MMA
//   GTC_GEMM_TCGEN05_3XTF32            operands split into tf32 hi + fp32 lo, tcgen05.mma.kind::tf32, K = 8 per MMA
//
//     C[row][n] = sum_p sum_k X[row + p][k] * Op[n][p*kp + k]          (row = audio row, n = operator row)
//
// Split product: X = Xhi + Xlo, Op = Ohi + Olo and
//     C ~= Xhi*Ohi + Xlo*Ohi + Xhi*Olo        (the dropped Xlo*Olo term is ~2^-22 relative)
// accumulated in fp32 in TMEM.  Per CTA tile: TBM = 128 rows x NC operator rows (NC <= 256 = one tcgen05.mma N; 240 for
// the 960-row operator of the cqt.py recipe).
//
// Warp roles (TC_THREADS = 64 + 32 * TC_EPI_WARPS = 320 threads, persistent CTAs, one per SM):
//   warp 0      TMA producer : cp.async.bulk.tensor.2d tiles of Xhi/Xlo/Ohi/Olo into a TC_STAGES-deep smem ring; a stage
//                              holds one k-block = TC_KB_BYTES bytes of K per row (default 64 B = one 64-byte swizzle row,
//                              4 stages; gtc_common.cuh)
//   warp 1      MMA issuer   : one elected lane issues the three split MMAs per k-step (TKB_BYTES / 32 k-steps per
//                              k-block) and commits to mbarriers; also owns the TMEM allocation (512 columns = 2
//                              accumulator stages of up to 256 columns)
//   warps 2..9  epilogue     : tcgen05.ld 32x32b (thread = one row, warp = one TMEM lane quarter x one column half), K-split
//                              partial sums added in fp32 RN registers, then re^2+im^2 + row max (or the raw complex
//                              values) stored straight to global memory
// The second audio row of a segment (hop = window/2, cqt.py:26-27) is just the TMA row coordinate + p.
//
// Slotted variants (SLOT = 1 decimator, 2 octave response) serve the structured CQT (cqt_structured.cu): rows are overlapping
// windows of per-segment fp16 hi/lo planes (3-D tensor maps).  The decimator's banded Toeplitz operator is resident in shared
// memory (RES, RingRes), its tile is one K split read straight from TMEM, and its epilogue transposes vectors across lanes so
// that every store instruction writes whole 128-byte lines.  What was measured on the way (ring depth, L2 prefetch of the next
// tile, TMA box shape, slot stride, no-MMA / no-load / no-store builds) is in profiles/r02t_decimator.md.
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include <vector>
#include "gtc_common.cuh"

namespace gtc {

constexpr int TBM = 128;             // rows per tile (UMMA M)
constexpr int TKB_BYTES = TC_KB_BYTES;   // bytes of K per k-block = one swizzle row (128 or 64)
static_assert(TKB_BYTES == 128 || TKB_BYTES == 64 || TKB_BYTES == 32, "k-block rows are one 128-, 64- or 32-byte swizzle row");
constexpr int TBK = TKB_BYTES / 4;   // fp32 per k-block
constexpr int TMAXN = 256;           // max operator rows per tile (UMMA N)
constexpr int TSTAGES = TC_STAGES;
constexpr int TUMMA_K = 8;           // tf32 per MMA = 32 bytes of K (16 fp16)
constexpr uint32_t X_TILE_BYTES = TBM * TBK * 4;        // 128 rows x TKB_BYTES   ( 8 KB at 64 B)
constexpr uint32_t OP_TILE_BYTES = TMAXN * TBK * 4;     // 256 rows x TKB_BYTES   (16 KB at 64 B; NC rows used)
constexpr uint32_t STAGE_BYTES = 2 * X_TILE_BYTES + 2 * OP_TILE_BYTES;   // hi + lo of both operands (48 KB at 64 B)
constexpr uint32_t TC_SMEM_BYTES = TSTAGES * STAGE_BYTES + 1024 /*align slack*/;
constexpr int TMAXSTAGES = 8;
// Ring geometry per tile width: a stage holds hi + lo of 128 X rows and of NC operator rows, and the ring is as deep as the
// shared memory allows (at most 8 stages).  NC = 240: 46 KB x 4; NC = 128 (decimator): 32 KB x 6; NC = 32 (octave response):
// 20 KB x 8 -- the narrow kernels have K = 4..22 k-blocks per tile and live on load latency, so depth is what they need.
// The decimator (SLOT = 1) multiplies MH = 2 row blocks per operator tile: 48 KB x 4.
template <int NC, int MH = 1>
struct Ring {
  static constexpr uint32_t op_bytes = ((uint32_t)NC * TBK * 4 + 1023u) & ~1023u;
  static constexpr uint32_t stage_bytes = 2 * MH * X_TILE_BYTES + 2 * op_bytes;     // MH row blocks of 128 share one operator tile
  static constexpr int fit = (int)((TC_SMEM_BYTES - 1024) / stage_bytes);
  static constexpr int stages = fit > TMAXSTAGES ? TMAXSTAGES : fit;
};
// RES kernels (decimator): the operator never travels again after the first load.  The banded Toeplitz operator of a 2:1 FIR,
// Op[n][i] = h[2n + c - i], repeats itself from one k-block to the next shifted by `res_shift` = (elements per k-block) / 2 rows:
//     Op[n][kb * EPK + e] = Master[n + res_shift * (nkb - 1 - kb)][e],      Master[r][e] = Op[r][(nkb - 1) * EPK + e] extended upwards
// so ONE [kResRows][k-block] tile per hi/lo plane (2 x 32 KB at 64-byte k-blocks), loaded once per CTA, serves every k-block of
// every tile: the B descriptor of k-block kb just starts res_shift * (nkb - 1 - kb) rows further down (a multiple of the 8-row
// swizzle group, so the swizzle phase is unchanged).  The ring then carries X rows only.
constexpr int kResRows = 512;
constexpr uint32_t RES_MASTER_BYTES = 2u * kResRows * TKB_BYTES;              // hi + lo
template <int MH>
struct RingRes {
  static constexpr uint32_t stage_bytes = 2 * MH * X_TILE_BYTES;
  static constexpr uint32_t budget = 227u * 1024u - 1024u /*static smem + slack*/ - 1024u /*align*/ - RES_MASTER_BYTES;
  static constexpr int fit = (int)(budget / stage_bytes);
#ifdef TC_RES_STAGES                      // experiments: ring depth of the resident-operator kernels
  static constexpr int stages = TC_RES_STAGES;
#else
  static constexpr int stages = fit > TMAXSTAGES ? TMAXSTAGES : fit;
#endif
  static constexpr uint32_t smem_bytes = RES_MASTER_BYTES + stages * stage_bytes + 1024;
};
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
#ifndef TC_MAXNREG
#define TC_MAXNREG 168     // 10 warps = 3 on the fullest SM sub-partition (16 384 registers): 3 x 32 x 168 is the most that launches
#endif

struct TcParams {
  int nc;                // operator rows per tile
  float out_scale;       // 2^-(x_scale + operator scale) of the fp16x2 engine, 1 otherwise
  int kb_per_split;      // k-blocks accumulated inside the tensor core before an fp32 RN add in registers
  int n_chunks;          // tiles along N
  int n_out;             // valid operator rows
  int kb_per_part;       // kp / 32
  int parts;
  int64_t m_tiles;
  float* mag2;           // [rows][n_out/2] or null
  float* cplx;           // [rows][n_out]   or null
  float* rowmax;         // [rows]
  FinishArgs fin;        // fin.out_db != null: the CTA that completes the last N tile of a 128-row block also does its dB finish
  SlotArgs slots;        // SLOT > 0 kernels: slotted rows of the structured CQT
  // zero-skipping schedule (null = every k-block against the whole N tile).  Entry = kb | g0 << 16 | ng << 24: multiply k-block
  // kb with the ng row groups starting at group g0 of the tile; the first entry of every K split covers the whole tile
  // (it zero-initialises the accumulator stage).  The N tile index is then rotated by the CTA's iteration (tile_of).
  const uint32_t* sched;
  const int* sched_len;
  const CUtensorMap* op_maps;   // [n_groups][2] operator boxes of (g + 1) * grp_rows rows (hi, lo), device memory
  int sched_pitch, grp_rows, rotate;
  int res_shift;         // RES kernels: operator rows by which consecutive k-blocks of the Toeplitz operator are shifted
  int reverse;           // slotted kernels: walk the tiles from the last segment to the first
  // SLOT == 1 (decimator): the band of the Toeplitz operator.  Entry e multiplies k-block band_kb[e] with the band_ng[e] groups of
  // 16 operator rows starting at group band_g0[e] (tcgen05.mma with N = 16 * ng at accumulator column 16 * g0); entry 0 covers
  // the whole tile (it zero-initialises the accumulator).  band_n == 0: every k-block against all rows.
  int band_n;
  uint8_t band_kb[kMaxBand], band_g0[kMaxBand], band_ng[kMaxBand];
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile whose rows are ONE swizzle row of TKB_BYTES (128, 64 or 32) bytes; 8-row groups 8 * TKB_BYTES apart
// (cute::UMMA::SmemDescriptor; layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B)
__device__ __forceinline__ uint64_t make_swizzle_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);          // start address  [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)((8 * TKB_BYTES) >> 4) << 32;          // stride byte offset = 8 rows                      [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)                      [46,48)
  d |= (uint64_t)(TKB_BYTES == 128 ? 2 : TKB_BYTES == 64 ? 4 : 6) << 61;   // layout type                   [61,64)
  return d;
}
// cute::UMMA::InstrDescriptor : c=F32, a=b=fmt (0 = F16, 2 = TF32), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(uint32_t fmt, int m, int n) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// tcgen05.ld of 8 columns (32x32b.x8)
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr) : "memory");
}

// One round of the 8 x 8 transpose of 16-byte vectors among the 8 lanes of a group (xor butterfly): lane i, slot c  <->  lane c,
// slot i after the rounds B = 4, 2, 1.  All register indices are compile-time.
template <int B>
__device__ __forceinline__ void xpose8_round(uint4 (&ch)[8], int i8) {
  const bool up = (i8 & B) != 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c & B) continue;
    const uint4 send = up ? ch[c] : ch[c | B];
    uint4 recv;
    recv.x = __shfl_xor_sync(0xffffffffu, send.x, B);
    recv.y = __shfl_xor_sync(0xffffffffu, send.y, B);
    recv.z = __shfl_xor_sync(0xffffffffu, send.z, B);
    recv.w = __shfl_xor_sync(0xffffffffu, send.w, B);
    if (up) ch[c] = recv; else ch[c | B] = recv;
  }
}

// dB finish of one completed 128-row block by the 8 epilogue warps of the CTA that completed it (see the call site).
// Two rows x 4 vectors in flight per lane: the accumulation loop above it runs at the 168-register cap (10 warps = 3 on the
// fullest SM sub-partition), more rows in flight -- or a non-inlined call -- push ptxas into spilling inside that loop.
__device__ __forceinline__ void fused_finish_block(const TcParams& prm, int64_t m_tile, int warp, int lane) {
  const FinishArgs& f = prm.fin;
  __threadfence();
  const int n_mag_f = f.n_bins * f.n_frames;
  constexpr int RG = 2;                                    // rows in flight per warp (x 4 vectors per lane)
  for (int i0 = warp - 2; i0 < TBM; i0 += TC_EPI_WARPS * RG) {
    int seg[RG];
#pragma unroll
    for (int k = 0; k < RG; ++k) seg[k] = __ldg(f.seg_of_row + m_tile * TBM + i0 + TC_EPI_WARPS * k);   // -1: starts no segment
    if ((n_mag_f & 3) == 0) {
      const int nv = n_mag_f >> 2;
      for (int ob = 0; ob < nv; ob += 128) {
        float4 v[RG][4];
        float ref[RG];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
          const int64_t r = m_tile * TBM + i0 + TC_EPI_WARPS * k;
          ref[k] = seg[k] >= 0 ? __ldcg(prm.rowmax + r) : 0.f;
          const float4* s4 = reinterpret_cast<const float4*>(prm.mag2 + r * n_mag_f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int o = ob + lane + 32 * u;
            v[k][u] = (seg[k] >= 0 && o < nv) ? __ldcg(s4 + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int k = 0; k < RG; ++k) {
          if (seg[k] < 0) continue;
          const DbScale scale(ref[k], f.power, f.amin, f.top_db, f.cut_db, f.floor_db);
          float4* d4 = reinterpret_cast<float4*>(f.out_db + (int64_t)seg[k] * n_mag_f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int o = ob + lane + 32 * u;
            if (o < nv) d4[o] = make_float4(scale(v[k][u].x), scale(v[k][u].y), scale(v[k][u].z), scale(v[k][u].w));
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < RG; ++k) {
        const int64_t r = m_tile * TBM + i0 + TC_EPI_WARPS * k;
        if (seg[k] >= 0)
          finish_row_db(prm.mag2 + r * n_mag_f, __ldcg(prm.rowmax + r), f.out_db + (int64_t)seg[k] * n_mag_f, lane, f.n_bins,
                        f.n_frames, f.power, f.amin, f.top_db, f.cut_db, f.floor_db);
      }
    }
  }
}

// NC = operator rows per tile (tcgen05.mma N).  Each of the 8 epilogue warps owns 32 rows x NC/2 columns and keeps
// their running sums in registers: the tensor core adds into its TMEM accumulator with truncation (RZ), which over
// the ~1650 accumulator updates of a full K pass biases the result by ~3e-5 relative -- measured on B200 -- so K is
// cut into splits of `kb_per_split` k-blocks, each accumulated from zero in one of two TMEM stages and then summed
// in fp32 (round-to-nearest) by the epilogue warps while the next split is already being multiplied.
// kHalf: operands are fp16 hi/lo pairs (kind::f16, 64 elements per 128-byte k-block, 16 per MMA) instead of tf32 pairs
// (kind::tf32, 32 per k-block, 8 per MMA); in bytes the tiles, the swizzle and the +32 B k-step are identical.
//
// Register cap TC_MAXNREG: the CTA owns the SM (184 KB of shared memory), so the cap is simply what 320 threads can have.
// (r01 experiment: -DTC_MAXNREG=152 lets one 224-thread x 72-register patch CTA (patches.cu) fit beside this one; co-running
// the two kernels measured SLOWER than back to back -- both live on L2 bandwidth -- profiles/r01j_coresident.md.)
// TFM > 0: frame-major tiles (OpLayout, gtc_common.cuh) of a TFM-frame recipe: NC = 2 * TFM * bins_per_tile; 0: plain rows.
// SLOT > 0: the M operand is slotted (SlotArgs, gtc_common.cuh): 1 = decimator epilogue, 2 = response epilogue.
// SCHED: multiply every k-block only with the row groups (frames) of the tile that are non-zero on it (TcParams::sched).
// RES: the (Toeplitz) operator is resident in shared memory as one master tile (see RingRes); SLOT == 1 only.
template <int NC, bool kComplex, bool kHalf, int TFM, int SLOT = 0, bool SCHED = false, bool RES = false>
__global__ void __maxnreg__(TC_MAXNREG)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
               const __grid_constant__ CUtensorMap tm_ohi, const __grid_constant__ CUtensorMap tm_olo,
               const __grid_constant__ TcParams prm) {
  constexpr int H = NC / 2;                  // columns per epilogue warp
  constexpr int EPK = TKB_BYTES / (kHalf ? 2 : 4);   // operand elements per k-block
  static_assert(NC % 16 == 0 && NC <= TMAXN && H % 8 == 0, "unsupported tile width");
  static_assert(TFM == 0 || NC % (TFM > 0 ? 2 * TFM : 1) == 0, "a frame-major tile holds whole bins x TFM frames");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // MH: 128-row blocks that share one operator tile.  The decimator is bound by L2 -> SM operand delivery and every tile uses
  // the SAME Toeplitz operator: two row blocks (32 segments x 8 rows) per operator stage cut its traffic by a quarter.
  constexpr int MH = SLOT > 0 ? 2 : 1;
  static_assert(MH * NC <= TMAXN, "the row blocks of a tile share one accumulator stage");
  static_assert(!RES || (SLOT == 1 && !SCHED && kHalf && NC == TBM), "resident operator: the decimator's square fp16x2 tiles only");
  constexpr int NSTAGES = RES ? RingRes<MH>::stages : Ring<NC, MH>::stages;
  constexpr uint32_t STAGE_B = RES ? RingRes<MH>::stage_bytes : Ring<NC, MH>::stage_bytes, OP_B = Ring<NC, MH>::op_bytes;
  __shared__ __align__(8) uint64_t s_bars[2 * TMAXSTAGES + 5];
  __shared__ uint32_t s_tmem_slot;
  __shared__ int s_block_done;
  // the schedule lives in shared memory: the producer and the MMA issuer read one entry per k-block on their critical path,
  // and with 193 KB of the SM carved out as shared memory a global load is an L2 round trip (~300 cycles x 138 entries)
  constexpr int kSchedMax = 4096;
  __shared__ uint32_t s_sched[SCHED ? kSchedMax : 1];
  __shared__ int s_sched_len[SCHED ? 16 : 1];
  const uint32_t smem_al = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // RES: [master hi][master lo] in front of the ring
  const uint32_t res_hi = smem_al, res_lo = smem_al + RES_MASTER_BYTES / 2;
  const uint32_t smem_base = smem_al + (RES ? RES_MASTER_BYTES : 0u);
  // stage s : [Xhi][Xlo][Ohi][Olo]   (X_TILE_BYTES, X_TILE_BYTES, OP_TILE_BYTES, OP_TILE_BYTES); RES: [Xhi][Xlo] only
  auto st_xhi = [&](int s) { return smem_base + s * STAGE_B; };
  auto st_xlo = [&](int s) { return smem_base + s * STAGE_B + MH * X_TILE_BYTES; };
  auto st_ohi = [&](int s) { return smem_base + s * STAGE_B + 2 * MH * X_TILE_BYTES; };
  auto st_olo = [&](int s) { return smem_base + s * STAGE_B + 2 * MH * X_TILE_BYTES + OP_B; };
  const uint32_t bar_base = smem_u32(s_bars);
  auto bar_full = [&](int s) { return bar_base + 8 * s; };
  auto bar_empty = [&](int s) { return bar_base + 8 * (NSTAGES + s); };
  auto bar_tfull = [&](int a) { return bar_base + 8 * (2 * NSTAGES + a); };
  auto bar_tempty = [&](int a) { return bar_base + 8 * (2 * NSTAGES + 2 + a); };
  const uint32_t bar_res = bar_base + 8 * (2 * NSTAGES + 4);       // RES: the master tile has landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), TC_EPI_WARPS); }
    mbar_init(bar_res, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem_slot), 512);
  if (SCHED) {
    for (int i = threadIdx.x; i < prm.n_chunks * prm.sched_pitch; i += blockDim.x) s_sched[i] = __ldg(prm.sched + i);
    if (threadIdx.x < prm.n_chunks) s_sched_len[threadIdx.x] = __ldg(prm.sched_len + threadIdx.x);
  }
  if (SLOT > 0) pdl_launch_dependents();          // structured CQT chain: the next kernel may start its prologue
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_slot;
  if (SLOT > 0) pdl_wait();                       // ... and this one touches its operands only after its predecessor has completed

  const int nkb = prm.parts * prm.kb_per_part;
  const int n_splits = (nkb + prm.kb_per_split - 1) / prm.kb_per_split;
  const int64_t n_tiles = prm.m_tiles * prm.n_chunks;
  // tile -> (row block, N tile): row-block major, so that the N tiles of a row block run side by side on adjacent CTAs and
  // share their X rows in L2.  With a schedule the N tiles cost differently (top octaves ~0.2, low octaves 1), so the N tile
  // index is rotated by the CTA's iteration: every CTA sees every tile type in turn instead of always the same one.
  // (Passes over the N tiles in order of cost balance better on paper and measured 14 % SLOWER: each pass re-reads all X
  // rows from HBM -- 4.7 TB/s in the dense passes.)
  auto tile_of = [&](int64_t tile, int iter, int64_t& m_tile, int& chunk) {
    if (SLOT > 0 && prm.reverse) tile = n_tiles - 1 - tile;      // last segments first: the ones the previous kernel wrote last
    m_tile = tile / prm.n_chunks;
    chunk = (int)(tile - m_tile * prm.n_chunks);
    if (SCHED && prm.rotate) chunk = (chunk + iter) % prm.n_chunks;
  };
#ifdef TC_EXP_SKIP_OLO   // timing experiment only (wrong results): is the kernel bound by operand delivery from L2?
  constexpr uint32_t stage_tx = 2 * X_TILE_BYTES + 1 * (uint32_t)NC * TBK * 4;
#else
  constexpr uint32_t stage_tx = 2 * MH * X_TILE_BYTES + (RES ? 0u : 2 * (uint32_t)NC * TBK * 4);
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_xhi)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_xlo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_ohi)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_olo)) : "memory");
      if (RES) {
        // master tile: 128-row boxes of the operator's LAST k-block, of the k-block 128 / shift further left, ... and, once the
        // k-blocks run out, of k-block 0 from row r - shift * (nkb - 1) on (rows past the operator are zero-filled by the TMA)
        const int need = TBM + prm.res_shift * (nkb - 1);
        mbar_expect_tx(bar_res, (uint32_t)((need + TBM - 1) / TBM) * 2u * TBM * TKB_BYTES);
        for (int r = 0; r < need; r += TBM) {
          int kb = nkb - 1 - r / prm.res_shift, n_start = 0;
          if (kb < 0) { n_start = r - prm.res_shift * (nkb - 1); kb = 0; }
          tma_load_2d(res_hi + (uint32_t)r * TKB_BYTES, &tm_ohi, bar_res, kb * EPK, n_start);
          tma_load_2d(res_lo + (uint32_t)r * TKB_BYTES, &tm_olo, bar_res, kb * EPK, n_start);
        }
      }
      int stage = 0; uint32_t phase = 0;
      int iter = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
        int64_t m_tile; int chunk;
        tile_of(tile, iter, m_tile, chunk);
        const int row0 = (int)(m_tile * TBM);
        const int n0 = chunk * NC;
        const int n_ent = SCHED ? s_sched_len[chunk] : (SLOT == 1 && prm.band_n > 0) ? prm.band_n : nkb;
        for (int e = 0; e < n_ent; ++e) {
          int kb = (SLOT == 1 && prm.band_n > 0) ? (int)prm.band_kb[e] : e, g0 = 0, rows = NC;
          if (SCHED) {
            const uint32_t w = s_sched[chunk * prm.sched_pitch + e];
            kb = (int)(w & 0xffffu); g0 = (int)((w >> 16) & 0xffu); rows = (int)(w >> 24) * prm.grp_rows;
          }
          mbar_wait(bar_empty(stage), phase ^ 1);
#ifdef TC_EXP_NO_XLOAD      // timing experiment only: no operand traffic at all in the slotted kernels
          if (SLOT > 0) { mbar_arrive(bar_full(stage)); if (++stage == NSTAGES) { stage = 0; phase ^= 1; } continue; }
#endif
          mbar_expect_tx(bar_full(stage), rows == NC ? stage_tx : 2 * X_TILE_BYTES + 2 * (uint32_t)rows * TBK * 4);
          const int p = kb / prm.kb_per_part;
          const int kx = (kb - p * prm.kb_per_part) * EPK;
          if (SLOT == 0) {
            tma_load_2d(st_xhi(stage), &tm_xhi, bar_full(stage), kx, row0 + p);
            tma_load_2d(st_xlo(stage), &tm_xlo, bar_full(stage), kx, row0 + p);
          } else {                                           // per row block: 16 segments x 8 consecutive rows (windows) of each
            const int sg = (int)(m_tile / prm.slots.jgroups), jg = (int)(m_tile - (int64_t)sg * prm.slots.jgroups);
#pragma unroll
            for (int h = 0; h < MH; ++h) {
              const int bl = prm.slots.box_log2;
              tma_load_3d(st_xhi(stage) + h * X_TILE_BYTES, &tm_xhi, bar_full(stage), kx, jg << bl, (MH * sg + h) << (7 - bl));
              tma_load_3d(st_xlo(stage) + h * X_TILE_BYTES, &tm_xlo, bar_full(stage), kx, jg << bl, (MH * sg + h) << (7 - bl));
            }
          }
          if (RES) {
            // nothing: the operator is resident
          } else if (!SCHED || rows == NC) {
            tma_load_2d(st_ohi(stage), &tm_ohi, bar_full(stage), kb * EPK, n0);
#ifndef TC_EXP_SKIP_OLO
            tma_load_2d(st_olo(stage), &tm_olo, bar_full(stage), kb * EPK, n0);
#endif
          } else {                                           // only the row groups whose support reaches this k-block
            const CUtensorMap* om = prm.op_maps + 2 * (rows / prm.grp_rows - 1);
            tma_load_2d(st_ohi(stage), om, bar_full(stage), kb * EPK, n0 + g0 * prm.grp_rows);
            tma_load_2d(st_olo(stage), om + 1, bar_full(stage), kb * EPK, n0 + g0 * prm.grp_rows);
          }
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kHalf ? 0u : 2u, TBM, NC);
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;                                       // accumulator-stage use counter (one per K split)
      int iter = 0;
      if (RES) { mbar_wait(bar_res, 0); tc_fence_after(); }
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
        int64_t m_tile; int chunk;
        tile_of(tile, iter, m_tile, chunk);
        const int n_ent = SCHED ? s_sched_len[chunk] : (SLOT == 1 && prm.band_n > 0) ? prm.band_n : nkb;
        const int n_sp = SLOT == 1 ? 1 : SCHED ? (n_ent + prm.kb_per_split - 1) / prm.kb_per_split : n_splits;   // decimator: see its epilogue
        int kb = 0;                                          // entry index
        for (int sp = 0; sp < n_sp; ++sp, ++it) {
          const int acc = (int)(it & 1u);
          mbar_wait(bar_tempty(acc), ((it >> 1) & 1u) ^ 1u); // epilogue drained this accumulator stage
          tc_fence_after();
          uint32_t tmem_d = tmem_base + (uint32_t)acc * TMAXN;
          const int kb_end = SLOT == 1 ? n_ent : min(n_ent, kb + prm.kb_per_split);
          for (int first = 1; kb < kb_end; ++kb, first = 0) {
            uint32_t idesc_e = idesc;
            tmem_d = tmem_base + (uint32_t)acc * TMAXN;
            if (SCHED) {                                     // N = the row groups this k-block touches, at their accumulator columns
              const uint32_t w = s_sched[chunk * prm.sched_pitch + kb];
              const int rows = (int)(w >> 24) * prm.grp_rows;
              if (rows != NC) {
                idesc_e = make_idesc(kHalf ? 0u : 2u, TBM, rows);
                tmem_d += (uint32_t)(((w >> 16) & 0xffu) * prm.grp_rows);
              }
            }
            int kbk = kb;                                   // the operator's k-block of this entry
            uint32_t band_at = 0;                           // byte offset of the entry's first operator row inside the k-block tile
            if (SLOT == 1 && prm.band_n > 0) {              // only the 16-row groups of the Toeplitz band that reach this k-block
              kbk = (int)prm.band_kb[kb];
              const uint32_t g0 = prm.band_g0[kb], ng = prm.band_ng[kb];
              idesc_e = make_idesc(kHalf ? 0u : 2u, TBM, (int)(16u * ng));
              tmem_d += 16u * g0;
              band_at = 16u * g0 * TKB_BYTES;               // whole 8-row swizzle groups: the swizzle phase is unchanged
            }
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            const uint64_t dxh = make_swizzle_desc(st_xhi(stage)), dxl = make_swizzle_desc(st_xlo(stage));
            const uint32_t res_at = RES ? (uint32_t)(prm.res_shift * (nkb - 1 - kbk)) * TKB_BYTES : 0u;
            const uint64_t doh = make_swizzle_desc((RES ? res_hi + res_at : st_ohi(stage)) + band_at);
            const uint64_t dol = make_swizzle_desc((RES ? res_lo + res_at : st_olo(stage)) + band_at);
#ifdef TC_EXP_NO_MMA        // timing experiment only
            if (SLOT == 0)
#endif
#pragma unroll
            for (int k = 0; k < TBK / TUMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * TUMMA_K * 4) >> 4);   // +32 B per k-step inside the swizzle row
#pragma unroll
              for (int h = 0; h < MH; ++h) {                              // row blocks sharing this operator tile
                const uint64_t xa = adv + (uint64_t)((h * X_TILE_BYTES) >> 4);
                const uint32_t td = tmem_d + (uint32_t)(h * NC);
                if (kHalf) {
                  umma_f16(td, dxh + xa, doh + adv, idesc_e, (first && k == 0) ? 0u : 1u);
                  umma_f16(td, dxl + xa, doh + adv, idesc_e, 1u);
                  umma_f16(td, dxh + xa, dol + adv, idesc_e, 1u);
                } else {
                  umma_tf32(td, dxh + xa, doh + adv, idesc_e, (first && k == 0) ? 0u : 1u);
                  umma_tf32(td, dxl + xa, doh + adv, idesc_e, 1u);
                  umma_tf32(td, dxh + xa, dol + adv, idesc_e, 1u);
                }
              }
            }
            umma_commit(bar_empty(stage));                  // frees the smem slot when these MMAs retire
            if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_tfull(acc));                      // this split's partial sums are complete
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                        // which NC/2 column half
    const int n_mag = prm.n_out >> 1;
    uint32_t it = 0;
    int iter = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
      int64_t m_tile; int chunk;
      tile_of(tile, iter, m_tile, chunk);
      const int64_t row = m_tile * TBM + q * 32 + lane;
      const int n0 = chunk * NC + half * H;
      if constexpr (SLOT == 1) {
        // ---- decimator (structured CQT): ONE K split per tile (K = 22 k-blocks = 132 accumulator updates: the truncation of the
        //      tensor core's accumulation stays at the 4e-6 level the K splits of the long contraction are there to reach), so the
        //      epilogue takes each row block straight from TMEM, 64 values at a time -- no running sums in registers.
        //      tile row rho = (segment within the block) << bl | (row within the segment's group)
        const SlotArgs& sl = prm.slots;
        const int acc = (int)(it & 1u);
        mbar_wait(bar_tfull(acc), (it >> 1) & 1u);
        tc_fence_after();
        ++it;
        const int rho = q * 32 + lane;
        const int sg = (int)(m_tile / sl.jgroups), jg = (int)(m_tile - (int64_t)sg * sl.jgroups);
        const int bl = sl.box_log2, spb = 128 >> bl;                   // segments per row block
        const int64_t slot = (int64_t)sg * MH * spb + (rho >> bl);     // of row block 0; block h: + spb * h
        const int j = (jg << bl) + (rho & ((1 << bl) - 1));
        const float scale = prm.out_scale * sl.plane_scale;            // both powers of two
        // Row j holds outputs k = j * NC + n of the next octave; beyond the octave's length -> zeros (librosa fixes the length to
        // ceil(n / 2), and the next stage must see a zero-extended signal).
        // A thread owns 64 consecutive outputs of ONE row (128 bytes per plane), its neighbours rows 256 bytes further on: stored
        // as they lie, every 16-byte vector is a memory request of its own -- 8 192 per tile through the SM's one request port,
        // 4.2 us of a 17 us tile, and the TMA loads queue behind them (measured on B200: stage 1 of the inference recipe 330 us,
        // 259 us without the stores).  So the 8 lanes of a group first transpose their 8 x 8 vectors (3 xor-butterfly rounds of
        // shuffles): lane i then holds vector i of each of the group's 8 rows, and a store instruction writes 4 whole 128-byte
        // lines.
        static_assert(SLOT != 1 || H == 64, "decimator epilogue: 8 vectors of 8 halves per thread and plane");
        const int i8 = lane & 7, g8 = lane & ~7;
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * TMAXN + (uint32_t)(h * NC + half * H);
          uint32_t v[64];
#pragma unroll
          for (int c = 0; c < 64; c += 16) tmem_ld16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
          tmem_ld_wait();
          if (h == MH - 1) {                                           // the accumulator stage is free for the tile after next
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(acc));
          }
          const int64_t slot_h = slot + spb * h;
          const int valid = slot_h < sl.n_slots ? halved_len(__ldg(sl.seg_len + slot_h), sl.stage_out) : 0;
          const int k0 = j * NC + half * H;
          uint4 chh[8], chl[8];                                        // hi and lo planes of the row's 64 values: 2 x 8 vectors
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t wh[4], wl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = 8 * c + 2 * u;
              const float v0 = (k0 + e < valid) ? __uint_as_float(v[e]) * scale : 0.f;
              const float v1 = (k0 + e + 1 < valid) ? __uint_as_float(v[e + 1]) * scale : 0.f;
              const __half2 hh = __floats2half2_rn(v0, v1);
              const float2 hf = __half22float2(hh);
              const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
              wh[u] = *reinterpret_cast<const uint32_t*>(&hh);
              wl[u] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            chh[c] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            chl[c] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
          }
          xpose8_round<4>(chh, i8); xpose8_round<2>(chh, i8); xpose8_round<1>(chh, i8);
          xpose8_round<4>(chl, i8); xpose8_round<2>(chl, i8); xpose8_round<1>(chl, i8);
          // ch?[r] = vector i8 of the row held by lane g8 + r
          const int64_t col = sl.out_base + half * H + 8 * i8;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int rho_r = q * 32 + g8 + r;
            const int64_t slot_r = (int64_t)sg * MH * spb + spb * h + (rho_r >> bl);
            const int j_r = (jg << bl) + (rho_r & ((1 << bl) - 1));
            if (slot_r < sl.n_slots) {
              const int64_t at = col + slot_r * sl.out_stride + (int64_t)j_r * NC;
              *reinterpret_cast<uint4*>(sl.out_hi + at) = chh[r];
              *reinterpret_cast<uint4*>(sl.out_lo + at) = chl[r];
            }
          }
        }
        continue;
      }
      const int n_sp = SCHED ? (s_sched_len[chunk] + prm.kb_per_split - 1) / prm.kb_per_split : n_splits;
      float sum[MH * H];
      for (int sp = 0; sp < n_sp; ++sp, ++it) {
        const int acc = (int)(it & 1u);
        mbar_wait(bar_tfull(acc), (it >> 1) & 1u);
        tc_fence_after();
        if (TFM == 0) {
#pragma unroll
          for (int h = 0; h < MH; ++h) {                         // row blocks of the tile: accumulator columns h * NC ..
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * TMAXN + (uint32_t)(h * NC + half * H);
#pragma unroll
            for (int c = 0; c + 16 <= H; c += 16) {
              uint32_t r[16];
              tmem_ld16(taddr + c, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                sum[h * H + c + j] = sp == 0 ? __uint_as_float(r[j]) : sum[h * H + c + j] + __uint_as_float(r[j]);
            }
            if (H % 16) {
              constexpr int c = H - 8;
              uint32_t r[8];
              tmem_ld8(taddr + c, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sum[h * H + c + j] = sp == 0 ? __uint_as_float(r[j]) : sum[h * H + c + j] + __uint_as_float(r[j]);
            }
          }
        } else {
          // frame-major tile: accumulator column = t * 2B + bin_in_tile * 2 + {re, im}.  This warp takes HALF THE BINS of every
          // frame (B columns of each of the TFM frame groups), so that a thread ends up with all frames of its bins, i.e.
          // with a contiguous run of the final [bin][t] layout: sum[t * B + i] <- column t * 2B + half * B + i
          constexpr int B = TFM > 0 ? NC / (2 * TFM) : 16;
          static_assert(TFM == 0 || B % 8 == 0, "half a frame group must be whole 8-column TMEM loads");
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * TMAXN + (uint32_t)(half * B);
#pragma unroll
          for (int t = 0; t < (TFM > 0 ? TFM : 1); ++t) {
#pragma unroll
            for (int c = 0; c + 16 <= B; c += 16) {
              uint32_t r[16];
              tmem_ld16(taddr + t * 2 * B + c, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                sum[t * B + c + j] = sp == 0 ? __uint_as_float(r[j]) : sum[t * B + c + j] + __uint_as_float(r[j]);
            }
            if (B % 16) {
              constexpr int c = B - 8;
              uint32_t r[8];
              tmem_ld8(taddr + t * 2 * B + c, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sum[t * B + c + j] = sp == 0 ? __uint_as_float(r[j]) : sum[t * B + c + j] + __uint_as_float(r[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(acc));
      }
      // tile finished: undo the power-of-two operand scaling (exact), then |.|^2 + row max, or the raw complex values
      if (kHalf) {
#pragma unroll
        for (int c = 0; c < MH * H; ++c) sum[c] *= prm.out_scale;
      }
      if (SLOT > 0) {
        // ---- structured CQT: tile row rho = 8 * (segment within the group of 16) + (row within the group of 8)
        const SlotArgs& sl = prm.slots;
        const int rho = q * 32 + lane;
        const int sg = (int)(m_tile / sl.jgroups), jg = (int)(m_tile - (int64_t)sg * sl.jgroups);
        const int bl = sl.box_log2, spb = 128 >> bl;                   // segments per row block
        const int64_t slot = (int64_t)sg * MH * spb + (rho >> bl);     // of row block 0; block h: + spb * h
        const int j = (jg << bl) + (rho & ((1 << bl) - 1));
        if constexpr (SLOT == 2) {                                     // (the decimator, SLOT == 1, has its own epilogue above)
          // response: row j = frame t of the segment, columns = (bin, {re, im}) of the octave's filters
          const int t = j;
#pragma unroll
          for (int h = 0; h < MH; ++h) {
            const int64_t slot_h = slot + spb * h;
            const bool live_h = slot_h < sl.n_slots;
            const bool valid = live_h && t < __ldg(sl.seg_frames + slot_h);      // frames librosa keeps (precomputed: 7 divisions)
            float mx = 0.f;
#pragma unroll
            for (int c = 0; c < H; c += 2) {
              const int b = (half * H + c) >> 1;
              const float m = valid ? sum[h * H + c] * sum[h * H + c] + sum[h * H + c + 1] * sum[h * H + c + 1] : 0.f;
#ifdef TC_EXP_NO_STORE2     // timing experiment only: the response epilogue without its stores
              if (m == 12345.f)
#endif
              if (live_h && t < sl.t_max && b < sl.bin_cnt) {
                sl.out[(slot_h * sl.n_bins + sl.bin_lo + b) * sl.t_max + t] = m;
                mx = fmaxf(mx, m);
              }
            }
            // the lanes of a segment's row group share one atomic
            for (int d = 1; d < (1 << bl); d <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            if ((rho & ((1 << bl) - 1)) == 0 && live_h && mx > 0.f) atomicMax(reinterpret_cast<int*>(sl.segmax + slot_h), __float_as_int(mx));
          }
        }
      } else if (kComplex) {
        if (TFM == 0) {
#pragma unroll
          for (int c = 0; c < H; c += 4)
            if (n0 + c < prm.n_out)
              *reinterpret_cast<float4*>(prm.cplx + row * prm.n_out + n0 + c) = make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
        } else {
          constexpr int B = TFM > 0 ? NC / (2 * TFM) : 16;                 // raw values stay in the GEMM's column order
          float* dst = prm.cplx + row * prm.n_out + chunk * NC + half * B;
#pragma unroll
          for (int t = 0; t < (TFM > 0 ? TFM : 1); ++t)
#pragma unroll
            for (int c = 0; c < B; c += 4)
              *reinterpret_cast<float4*>(dst + t * 2 * B + c) = make_float4(sum[t * B + c], sum[t * B + c + 1], sum[t * B + c + 2], sum[t * B + c + 3]);
        }
      } else {
        float rmax = 0.f;
        if (TFM == 0) {
#pragma unroll
          for (int c = 0; c < H; c += 8) {
            float m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = sum[c + 2 * j] * sum[c + 2 * j] + sum[c + 2 * j + 1] * sum[c + 2 * j + 1];
            if (n0 + c < prm.n_out) {
              rmax = fmaxf(fmaxf(rmax, fmaxf(m[0], m[1])), fmaxf(m[2], m[3]));
              *reinterpret_cast<float4*>(prm.mag2 + row * n_mag + ((n0 + c) >> 1)) = make_float4(m[0], m[1], m[2], m[3]);
            }
          }
        } else {
          // frame-major tile: this thread holds all TFM frames of B/2 bins (see the loads above) = a contiguous run of
          // (B/2) * TFM values of the final [bin][t] layout, stored as 16-byte vectors
          constexpr int B = TFM > 0 ? NC / (2 * TFM) : 16;
          constexpr int NV = (B / 2) * (TFM > 0 ? TFM : 1);
          static_assert(TFM == 0 || NV % 4 == 0, "a thread's run of outputs must be whole float4s");
          float* dst = prm.mag2 + row * n_mag + (chunk * B + half * (B / 2)) * TFM;
#pragma unroll
          for (int k = 0; k < NV; k += 4) {
            float m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int bl = (k + j) / (TFM > 0 ? TFM : 1), t = (k + j) - bl * (TFM > 0 ? TFM : 1);
              const float re = sum[t * B + 2 * bl], im = sum[t * B + 2 * bl + 1];
              m[j] = re * re + im * im;
            }
            rmax = fmaxf(fmaxf(rmax, fmaxf(m[0], m[1])), fmaxf(m[2], m[3]));
            *reinterpret_cast<float4*>(dst + k) = make_float4(m[0], m[1], m[2], m[3]);
          }
        }
        atomicMax(reinterpret_cast<int*>(prm.rowmax + row), __float_as_int(rmax));
        // ---- optional fused dB finish (cqt.py:56-58; GTC_OPT_FUSE_FINISH, OFF by default).  The reference level of a segment
        //      is its maximum over ALL its outputs, i.e. over the n_chunks N tiles of this 128-row block, which different CTAs
        //      compute.  Every tile publishes its |C|^2 and row maxima (fence), then bumps the block's counter; the CTA
        //      that brings it to n_chunks owns the finished block and converts it while its MMA warp runs ahead into
        //      the next tile (two TMEM stages = two K splits of slack).  Nobody waits for anybody, so co-scheduling of
        //      the CTAs is not assumed.  |C|^2 is already in the final [bin][t] order (OpLayout), so the conversion is a
        //      straight 16-byte read of the L2-resident block.
        //      Measured on B200, same box (profiles/r02e_gemm_finish_ab.md): GEMM + separate finish 0.437 ms per
        //      28 200-row chunk, fused 0.77 ms.  Two reasons, both structural: (1) the block's re-read queues behind the
        //      TMA operand stream that saturates this SM's L2 port (~40 us per block instead of the ~16 us of slack), and
        //      (2) the CTA that arrives last at a block is the one that is already behind, converting puts it further
        //      behind, so the SAME CTA converts in every wave and the kernel ends when it does (+6 conversions, not +1.5).
        //      A 4-CTA cluster exchanging row maxima through DSMEM would avoid the re-read but strands 16 of 148 SMs
        //      (cluster size 4 packs 132); the stand-alone pass costs 19.5 us with nothing else in flight.
        if (prm.fin.out_db != nullptr) {
          const FinishArgs& f = prm.fin;
          __threadfence();
          asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
          if (warp == 2 && lane == 0) s_block_done = atomicAdd(f.tile_done + m_tile, 1) == prm.n_chunks - 1;
          asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
          if (s_block_done) {
            fused_finish_block(prm, m_tile, warp, lane);
            if (warp == 2 && lane == 0) f.tile_done[m_tile] = 0;     // ready for the next contraction over this workspace
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2D row-major [rows][cols] tensor of fp32 or fp16, box = one TKB_BYTES swizzle row of K x box_rows, OOB -> zeros
static int encode_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return GTC_E_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(TKB_BYTES / elem_bytes), box_rows};    // one swizzle row of K
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, TKB_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : TKB_BYTES == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GTC_REQUIRE(r == CUDA_SUCCESS, GTC_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GTC_OK;
}

// Kernel instantiations: plain row order for a few tile widths, frame-major tiles (NC = 2 * T * bins_per_tile) for the
// recipes in use: T = 5 (22.05 kHz, cqt.py recipe: 24 bins x 5 frames x 2 = 240) and T = 9 (44.1 kHz: 8 x 9 x 2 = 144).
static const int kPlainWidths[] = {256, 192, 128, 64, 32};

int tc_pick_plain_width(int n_out) {
  for (int nc : kPlainWidths)
    if (n_out % nc == 0) return nc;
  return 256;                       // ragged: TMA zero-fills operator rows beyond n_pad, stores are guarded
}

bool tc_has_frame_major_kernel(int nc, int n_frames) { return (nc == 240 && n_frames == 5) || (nc == 144 && n_frames == 9); }

// calls f(kernel) for the four (complex, half) variants of one (NC, TFM), or for the selected one
template <int NC, int TFM, typename F>
static int for_variant(int cplx /* -1 = all */, int half, F&& f) {
  int rc = GTC_OK;
  if ((cplx < 0 || cplx == 0) && (half < 0 || half == 0) && rc == GTC_OK) rc = f(gemm_tc_kernel<NC, false, false, TFM>);
  if ((cplx < 0 || cplx == 1) && (half < 0 || half == 0) && rc == GTC_OK) rc = f(gemm_tc_kernel<NC, true, false, TFM>);
  if ((cplx < 0 || cplx == 0) && (half < 0 || half == 1) && rc == GTC_OK) rc = f(gemm_tc_kernel<NC, false, true, TFM>);
  if ((cplx < 0 || cplx == 1) && (half < 0 || half == 1) && rc == GTC_OK) rc = f(gemm_tc_kernel<NC, true, true, TFM>);
  return rc;
}

template <typename F>
static int for_plan_kernels(const PlanImpl& p, int cplx, int half, F&& f) {
  if (p.bins_per_tile > 0) {
    if (p.nc == 240 && p.n_frames == 5) return for_variant<240, 5>(cplx, half, f);
    if (p.nc == 144 && p.n_frames == 9) return for_variant<144, 9>(cplx, half, f);
    set_error("no frame-major tensor-core kernel for tile %d x %d frames", p.nc, p.n_frames);
    return GTC_E_UNSUP;
  }
  switch (p.nc) {
    case 256: return for_variant<256, 0>(cplx, half, f);
    case 192: return for_variant<192, 0>(cplx, half, f);
    case 128: return for_variant<128, 0>(cplx, half, f);
    case 64: return for_variant<64, 0>(cplx, half, f);
    case 32: return for_variant<32, 0>(cplx, half, f);
    default: set_error("no tensor-core kernel for tile width %d", p.nc); return GTC_E_UNSUP;
  }
}

int tc_plan_init(PlanImpl& p, const uint8_t* h_nz, int kb_per_split) {
  CUtensorMap* maps = new CUtensorMap[2];
  p.tmap_op_hi = &maps[0];
  p.tmap_op_lo = &maps[1];
  int rc = encode_2d(&maps[0], p.d_op_hi, (uint64_t)p.n_pad, (uint64_t)p.k_total, (uint32_t)p.nc, p.elem_bytes);
  if (rc != GTC_OK) return rc;
  rc = encode_2d(&maps[1], p.d_op_lo, (uint64_t)p.n_pad, (uint64_t)p.k_total, (uint32_t)p.nc, p.elem_bytes);
  if (rc != GTC_OK) return rc;
  rc = for_plan_kernels(p, -1, -1, [](auto kern) -> int {
    GTC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    return GTC_OK;
  });
  if (rc != GTC_OK || h_nz == nullptr || p.bins_per_tile == 0 || p.elem_bytes != 2) return rc;   // scheduled kernels: frame-major fp16x2 plans

  // ---- zero-skipping schedule.  Per N tile: the k-blocks with any non-zero entry and, for each, the contiguous range of row
  //      groups it touches (a group = one frame of a frame-major tile: an operator row is non-zero on a window of samples
  //      around its frame, so the range is 1-2 frames wide in the top octaves and all of them in the low ones).
  const int n_chunks = (int)ceil_div(p.n_out, p.nc), n_grp = p.nc / p.grp_rows, nkb = p.k_total / p.kb_elems;
  const int exp_mode = getenv("GTC_TC_SCHED_MODE") ? atoi(getenv("GTC_TC_SCHED_MODE")) : 0;   // experiments: 1 = whole-tile entries, 2 = no rotation of the N tiles
  p.sched_rotate = (exp_mode & 2) ? 0 : 1;
  if (n_grp < 2 || n_grp > 255 || nkb > 65535) return GTC_OK;
  std::vector<uint32_t> sched((size_t)n_chunks * nkb, 0);
  std::vector<int> len(n_chunks, 0);
  long long work = 0;
  for (int c = 0; c < n_chunks; ++c) {
    int n = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      int g0 = -1, g1 = -1;
      for (int g = 0; g < n_grp; ++g)
        if (h_nz[(size_t)(c * n_grp + g) * nkb + kb]) { if (g0 < 0) g0 = g; g1 = g; }
      if (g0 < 0) continue;                                         // nothing of this tile lives on this k-block
      if (n % kb_per_split == 0 || (exp_mode & 1)) { g0 = 0; g1 = n_grp - 1; }   // first entry of a K split: whole tile, zero-initialises
      sched[(size_t)c * nkb + n++] = (uint32_t)kb | ((uint32_t)g0 << 16) | ((uint32_t)(g1 - g0 + 1) << 24);
      work += g1 - g0 + 1;
    }
    len[c] = n;
  }
  p.sched_fill = (float)work / (float)((long long)n_chunks * nkb * n_grp);
  if (p.sched_fill > 0.92f && !(exp_mode & 1)) return GTC_OK;                          // nothing worth skipping: keep the dense loops
  std::vector<CUtensorMap> gm((size_t)2 * n_grp);
  for (int g = 0; g < n_grp; ++g) {
    rc = encode_2d(&gm[2 * g], p.d_op_hi, (uint64_t)p.n_pad, (uint64_t)p.k_total, (uint32_t)((g + 1) * p.grp_rows), p.elem_bytes);
    if (rc == GTC_OK) rc = encode_2d(&gm[2 * g + 1], p.d_op_lo, (uint64_t)p.n_pad, (uint64_t)p.k_total, (uint32_t)((g + 1) * p.grp_rows), p.elem_bytes);
    if (rc != GTC_OK) return rc;
  }
  GTC_CUDA_CHECK(cudaMalloc((void**)&p.d_sched, sched.size() * sizeof(uint32_t)));
  GTC_CUDA_CHECK(cudaMemcpy(p.d_sched, sched.data(), sched.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  GTC_CUDA_CHECK(cudaMalloc((void**)&p.d_sched_len, n_chunks * sizeof(int)));
  GTC_CUDA_CHECK(cudaMemcpy(p.d_sched_len, len.data(), n_chunks * sizeof(int), cudaMemcpyHostToDevice));
  GTC_CUDA_CHECK(cudaMalloc(&p.d_op_maps, gm.size() * sizeof(CUtensorMap)));
  GTC_CUDA_CHECK(cudaMemcpy(p.d_op_maps, gm.data(), gm.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  p.sched_pitch = nkb;
  p.sched_ksplit = kb_per_split;
  if (p.bins_per_tile > 0 && p.elem_bytes == 2) {
    if (p.nc == 240 && p.n_frames == 5) {
      GTC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<240, false, true, 5, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
      GTC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<240, true, true, 5, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    } else {
      GTC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<144, false, true, 9, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
      GTC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<144, true, true, 9, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    }
  }
  return GTC_OK;
}

void tc_plan_free(PlanImpl& p) {
  if (p.tmap_op_hi) delete[] reinterpret_cast<CUtensorMap*>(p.tmap_op_hi);
  p.tmap_op_hi = p.tmap_op_lo = nullptr;
  if (p.d_sched) cudaFree(p.d_sched);
  if (p.d_sched_len) cudaFree(p.d_sched_len);
  if (p.d_op_maps) cudaFree(p.d_op_maps);
  p.d_sched = nullptr; p.d_sched_len = nullptr; p.d_op_maps = nullptr;
}

int launch_gemm_tc(const PlanImpl& p, const void* d_xhi, const void* d_xlo, int64_t n_rows_pad, int64_t n_rows_alloc,
                   float* d_mag2, float* d_cplx, float* d_rowmax, const FinishArgs& fin, cudaStream_t st) {
  GTC_REQUIRE(p.tmap_op_hi != nullptr, GTC_E_ARG, "plan was not created with the tcgen05 engine");
  CUtensorMap tm_xhi, tm_xlo;
  int rc = encode_2d(&tm_xhi, d_xhi, (uint64_t)n_rows_alloc, (uint64_t)p.kp, TBM, p.elem_bytes);
  if (rc != GTC_OK) return rc;
  rc = encode_2d(&tm_xlo, d_xlo, (uint64_t)n_rows_alloc, (uint64_t)p.kp, TBM, p.elem_bytes);
  if (rc != GTC_OK) return rc;
  TcParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.nc = p.nc;
  prm.kb_per_split = (p.tc_kb_per_split > 0 ? p.tc_kb_per_split : 8) * (128 / TKB_BYTES);   // option counts 128-byte blocks
  prm.n_chunks = (int)ceil_div(p.n_out, prm.nc);
  prm.n_out = p.n_out;
  prm.kb_per_part = p.kp / p.kb_elems;
  prm.out_scale = p.out_scale;
  prm.parts = p.parts;
  prm.m_tiles = n_rows_pad / TBM;
  prm.mag2 = d_mag2; prm.cplx = d_cplx; prm.rowmax = d_rowmax;
  prm.fin = fin;
  if (d_cplx != nullptr) prm.fin.out_db = nullptr;
  memset(&prm.slots, 0, sizeof(prm.slots));
  const bool use_sched = p.d_sched != nullptr && p.sched_ksplit == prm.kb_per_split && p.bins_per_tile > 0 && p.elem_bytes == 2 &&
                         prm.n_chunks <= 16 && prm.n_chunks * p.sched_pitch <= 4096;
  prm.sched = use_sched ? p.d_sched : nullptr;
  prm.sched_len = p.d_sched_len;
  prm.op_maps = reinterpret_cast<const CUtensorMap*>(p.d_op_maps);
  prm.sched_pitch = p.sched_pitch;
  prm.grp_rows = p.grp_rows;
  const int64_t n_tiles = prm.m_tiles * prm.n_chunks;
  const int64_t max_ctas = p.tc_max_ctas > 0 && p.tc_max_ctas < p.sm_count ? p.tc_max_ctas : p.sm_count;
  const unsigned grid = (unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas);
  const CUtensorMap& tm_ohi = *reinterpret_cast<const CUtensorMap*>(p.tmap_op_hi);
  const CUtensorMap& tm_olo = *reinterpret_cast<const CUtensorMap*>(p.tmap_op_lo);
  prm.rotate = use_sched && p.sched_rotate && grid % (unsigned)prm.n_chunks == 0;
  if (use_sched) {
    const bool cplx = d_cplx != nullptr;
    if (p.nc == 240 && p.n_frames == 5) {
      if (cplx) gemm_tc_kernel<240, true, true, 5, 0, true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_xhi, tm_xlo, tm_ohi, tm_olo, prm);
      else      gemm_tc_kernel<240, false, true, 5, 0, true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_xhi, tm_xlo, tm_ohi, tm_olo, prm);
    } else {
      if (cplx) gemm_tc_kernel<144, true, true, 9, 0, true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_xhi, tm_xlo, tm_ohi, tm_olo, prm);
      else      gemm_tc_kernel<144, false, true, 9, 0, true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_xhi, tm_xlo, tm_ohi, tm_olo, prm);
    }
    GTC_CUDA_CHECK(cudaGetLastError());
    return GTC_OK;
  }
  rc = for_plan_kernels(p, d_cplx != nullptr ? 1 : 0, p.elem_bytes == 2 ? 1 : 0, [&](auto kern) -> int {
    kern<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tm_xhi, tm_xlo, tm_ohi, tm_olo, prm);
    return GTC_OK;
  });
  if (rc != GTC_OK) return rc;
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

// ---- structured CQT: slotted rows (SlotArgs).  3-D map over one fp16 plane: (sample within the window, row within the
// segment, segment); rows overlap (row stride < window length), box = one k-block x 8 rows x 16 segments = a 128-row tile.
static int encode_3d(CUtensorMap* tm, const void* base, uint64_t k_extent, uint64_t rows, uint64_t slots, uint64_t row_step,
                     uint64_t slot_stride, int box_log2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return GTC_E_CUDA;
  cuuint64_t gdim[3] = {k_extent, rows, slots};
  cuuint64_t gstr[2] = {row_step * 2, slot_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)(TKB_BYTES / 2), (cuuint32_t)(1 << box_log2), (cuuint32_t)(128 >> box_log2)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  TKB_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : TKB_BYTES == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GTC_REQUIRE(r == CUDA_SUCCESS, GTC_E_CUDA, "cuTensorMapEncodeTiled (3-D, overlapping rows) failed with CUresult %d", (int)r);
  return GTC_OK;
}

// function attributes are per device: called by gtc_scqt_plan_create on the plan's device
int tc_slots_init() {
  GTC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<128, false, true, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
  GTC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<32, false, true, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
  GTC_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<128, false, true, 0, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)RingRes<2>::smem_bytes));
  return GTC_OK;
}

int launch_gemm_tc_slots(const PlanImpl& p, const __half* x_hi, const __half* x_lo, int64_t x_first, int64_t x_stride,
                         int64_t row_step, int rows_per_slot, const SlotArgs& slots, cudaStream_t st) {
  GTC_REQUIRE(p.tmap_op_hi != nullptr && p.elem_bytes == 2 && p.parts == 1, GTC_E_ARG, "slotted GEMM needs an fp16x2 plan with one part");
  GTC_REQUIRE(p.bins_per_tile == 0 && p.n_out <= p.nc, GTC_E_ARG, "slotted GEMM: the operator must be one plain N tile");
  GTC_REQUIRE((slots.slot_mode == 1 && p.nc == 128) || (slots.slot_mode == 2 && p.nc == 32), GTC_E_UNSUP,
              "slotted GEMM: no kernel for mode %d with tile width %d", slots.slot_mode, p.nc);
  GTC_REQUIRE(rows_per_slot > 0 && rows_per_slot % 2 == 0 && slots.n_slots > 0, GTC_E_ARG, "slotted GEMM: bad geometry");
  const int box_log2 = rows_per_slot % 8 == 0 ? 3 : rows_per_slot % 4 == 0 ? 2 : 1;     // rows of a segment per TMA box
  GTC_REQUIRE(((x_first * 2) & 15) == 0 && ((row_step * 2) & 15) == 0 && ((x_stride * 2) & 15) == 0, GTC_E_ARG,
              "slotted GEMM: windows must start on 16-byte boundaries");
  CUtensorMap tm_xhi, tm_xlo;
  int rc = encode_3d(&tm_xhi, x_hi + x_first, (uint64_t)p.kp, (uint64_t)rows_per_slot, (uint64_t)slots.n_slots, (uint64_t)row_step, (uint64_t)x_stride, box_log2);
  if (rc != GTC_OK) return rc;
  rc = encode_3d(&tm_xlo, x_lo + x_first, (uint64_t)p.kp, (uint64_t)rows_per_slot, (uint64_t)slots.n_slots, (uint64_t)row_step, (uint64_t)x_stride, box_log2);
  if (rc != GTC_OK) return rc;
  TcParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.nc = p.nc;
  prm.kb_per_split = (p.tc_kb_per_split > 0 ? p.tc_kb_per_split : 8) * (128 / TKB_BYTES);
  prm.n_chunks = 1;
  prm.n_out = p.n_out;
  prm.kb_per_part = p.kp / p.kb_elems;
  prm.out_scale = p.out_scale;
  prm.parts = 1;
  prm.slots = slots;
  prm.slots.box_log2 = box_log2;
  prm.slots.jgroups = rows_per_slot >> box_log2;
  prm.m_tiles = ceil_div(slots.n_slots, 2 * (128 >> box_log2)) * prm.slots.jgroups;   // a slotted tile is two 128-row blocks
  const unsigned grid = (unsigned)(prm.m_tiles < p.sm_count ? prm.m_tiles : p.sm_count);
  const CUtensorMap& tm_ohi = *reinterpret_cast<const CUtensorMap*>(p.tmap_op_hi);
  const CUtensorMap& tm_olo = *reinterpret_cast<const CUtensorMap*>(p.tmap_op_lo);
  // decimator: resident master tile when the operator is Toeplitz with a whole-swizzle-group shift per k-block and fits it
  const int nkb = p.kp / p.kb_elems;
  prm.res_shift = p.res_shift;
  // The producer of a plane (split kernel, decimator) writes it from the first segment to the last, and only the last ~100 MB are
  // still in L2 when it ends.  The response reads its plane backwards (hot end first) and so leaves the HEAD hot for the decimator
  // that reads the same plane next, forwards (cqt_structured.cu launches them in that order).  GTC_SCQT_FORWARD=1: both forwards.
  static const bool fwd_env = getenv("GTC_SCQT_FORWARD") != nullptr;
  prm.reverse = slots.slot_mode == 2 && !fwd_env;
  static const int band_env = getenv("GTC_SCQT_DENSE_BAND") ? atoi(getenv("GTC_SCQT_DENSE_BAND")) : 0;   // A/B: 1 = every k-block against all rows
  if (slots.slot_mode == 1 && p.band_n > 0 && p.band_n <= kMaxBand && !band_env) {
    prm.band_n = p.band_n;
    memcpy(prm.band_kb, p.band_kb, (size_t)p.band_n);
    memcpy(prm.band_g0, p.band_g0, (size_t)p.band_n);
    memcpy(prm.band_ng, p.band_ng, (size_t)p.band_n);
  }
  const bool resident = slots.slot_mode == 1 && p.res_shift > 0 && p.res_shift % 8 == 0 && TBM % p.res_shift == 0 &&
                        TBM + p.res_shift * (nkb - 1) <= kResRows && p.n_pad == TBM;
  static const bool pdl = getenv("GTC_SCQT_NO_PDL") == nullptr;                             // programmatic dependent launch (gtc_common.cuh)
  if (resident)
    GTC_CUDA_CHECK(launch_pdl(gemm_tc_kernel<128, false, true, 0, 1, false, true>, dim3(grid), dim3(TC_THREADS), RingRes<2>::smem_bytes, st, pdl,
                              tm_xhi, tm_xlo, tm_ohi, tm_olo, prm));
  else if (slots.slot_mode == 1)
    GTC_CUDA_CHECK(launch_pdl(gemm_tc_kernel<128, false, true, 0, 1>, dim3(grid), dim3(TC_THREADS), TC_SMEM_BYTES, st, pdl, tm_xhi, tm_xlo, tm_ohi, tm_olo, prm));
  else
    GTC_CUDA_CHECK(launch_pdl(gemm_tc_kernel<32, false, true, 0, 2>, dim3(grid), dim3(TC_THREADS), TC_SMEM_BYTES, st, pdl, tm_xhi, tm_xlo, tm_ohi, tm_olo, prm));
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

}  // namespace gtc
