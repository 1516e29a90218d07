// CUDA-core fp32 engine for the segment-operator contraction (GTC_GEMM_SIMT_FP32).
// It is the validation engine for the tcgen05 path (same operands, plain fp32 FMA, no tensor cores):
//     C[row][n] = sum_p sum_k X[row + p][k] * Op[n][p*kp + k],    X = xhi + xlo (exact fp32 audio)
// 128x128 tile per CTA, 8x8 per thread, BK = 16, register-prefetch double buffering.
// Epilogue: mag2[row][n/2] = re^2 + im^2 and atomicMax of the row maximum, or the raw complex values.
#include "gtc_common.cuh"

namespace gtc {

constexpr int SBM = 128, SBN = 128, SBK = 16, SPAD = 4;

template <bool kComplex>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ xhi, const float* __restrict__ xlo, const float* __restrict__ op,
                 int kp, int parts, int k_total, int n_out, float* __restrict__ mag2, float* __restrict__ cplx,
                 float* __restrict__ rowmax) {
  __shared__ __align__(16) float As[2][SBK][SBM + SPAD];
  __shared__ __align__(16) float Bs[2][SBK][SBN + SPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;
  const int kb_per_part = kp / SBK;
  const int nkb = parts * kb_per_part;

  // global-load mapping: 128 rows x 4 float4 per operand tile -> 2 float4 per thread
  const int lrow = tid >> 2, lq = tid & 3;
  float4 ra[2], rb[2];
  auto gload = [&](int kb) {
    const int p = kb / kb_per_part;
    const int kk = (kb - p * kb_per_part) * SBK + lq * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t row = m0 + lrow + h * 64 + p;
      const float4 a = __ldg(reinterpret_cast<const float4*>(xhi + row * kp + kk));
      const float4 b = __ldg(reinterpret_cast<const float4*>(xlo + row * kp + kk));
      ra[h] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      rb[h] = __ldg(reinterpret_cast<const float4*>(op + (int64_t)(n0 + lrow + h * 64) * k_total + kb * SBK + lq * 4));
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      As[buf][lq * 4 + 0][r] = ra[h].x; As[buf][lq * 4 + 1][r] = ra[h].y;
      As[buf][lq * 4 + 2][r] = ra[h].z; As[buf][lq * 4 + 3][r] = ra[h].w;
      Bs[buf][lq * 4 + 0][r] = rb[h].x; Bs[buf][lq * 4 + 1][r] = rb[h].y;
      Bs[buf][lq * 4 + 2][r] = rb[h].z; Bs[buf][lq * 4 + 3][r] = rb[h].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int cur = kb & 1;
    if (kb + 1 < nkb) gload(kb + 1);
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nkb) {
      sstore(cur ^ 1);
      __syncthreads();
    }
  }

  const int n_mag = n_out >> 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (kComplex) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + h * 64 + tx * 4;
        if (n < n_out)
          *reinterpret_cast<float4*>(cplx + row * n_out + n) =
              make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      }
    } else {
      float rmax = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + h * 64 + tx * 4;
        const float ma = acc[i][h * 4 + 0] * acc[i][h * 4 + 0] + acc[i][h * 4 + 1] * acc[i][h * 4 + 1];
        const float mb = acc[i][h * 4 + 2] * acc[i][h * 4 + 2] + acc[i][h * 4 + 3] * acc[i][h * 4 + 3];
        if (n < n_out) {
          *reinterpret_cast<float2*>(mag2 + row * n_mag + (n >> 1)) = make_float2(ma, mb);
          rmax = fmaxf(rmax, fmaxf(ma, mb));
        }
      }
#pragma unroll
      for (int s = 1; s < 16; s <<= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, s));
      if (tx == 0) atomicMax(reinterpret_cast<int*>(rowmax + row), __float_as_int(rmax));   // rmax >= 0
    }
  }
}

int launch_gemm_simt(const PlanImpl& p, const float* d_xhi, const float* d_xlo, int64_t n_rows_pad, float* d_mag2,
                     float* d_cplx, float* d_rowmax, cudaStream_t st) {
  dim3 grid((unsigned)(n_rows_pad / SBM), (unsigned)(p.n_pad / SBN));
  if (d_cplx)
    gemm_simt_kernel<true><<<grid, 256, 0, st>>>(d_xhi, d_xlo, p.d_op, p.kp, p.parts, p.k_total, p.n_out, nullptr,
                                                 d_cplx, nullptr);
  else
    gemm_simt_kernel<false><<<grid, 256, 0, st>>>(d_xhi, d_xlo, p.d_op, p.kp, p.parts, p.k_total, p.n_out, d_mag2,
                                                  nullptr, d_rowmax);
  GTC_CUDA_CHECK(cudaGetLastError());
  return GTC_OK;
}

}  // namespace gtc
