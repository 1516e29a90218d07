// Experiment: does cuTensorMapEncodeTiled accept OVERLAPPING rows (dim-1 stride smaller than the dim-0 extent) and does
// cp.async.bulk.tensor.3d deliver them?  (sliding windows over a 1-D signal as the M operand of a Toeplitz GEMM)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_overlap_test tma_overlap_test.cu -lcuda && ./tma_overlap_test
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap tm, __half* out, int c0, int c1, int c2) {
  __shared__ __align__(1024) __half tile[16 * 8 * 32];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(tile);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16 * 8 * 32 * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(d), "l"((uint64_t)&tm), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  __syncthreads();
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < 16 * 8 * 32; i += blockDim.x) out[i] = tile[i];
}
int main() {
  const int PF = 512, S = 4096, NS = 20, KT = 672, R = 8, STR = 256;
  size_t n = PF + (size_t)NS * S + 4096;
  std::vector<__half> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2half((float)(i % 2048));
  __half *d, *o;
  cudaMalloc(&d, n * 2); cudaMalloc(&o, 16 * 8 * 32 * 2);
  cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*E)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  for (int sw = 0; sw < 2; ++sw) {
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)KT, (cuuint64_t)R, (cuuint64_t)NS};
    cuuint64_t gstr[2] = {(cuuint64_t)STR * 2, (cuuint64_t)S * 2};
    cuuint32_t box[3] = {32, 8, 16}, es[3] = {1, 1, 1};
    CUresult r = ((E)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d + PF - 200, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (swizzle %d): CUresult %d\n", sw, (int)r);
    if (r != CUDA_SUCCESS) continue;
    const int c0 = 64, c1 = 0, c2 = 8;     // k offset 64, rows 0..7, slots 8..23 (20..23 out of bounds -> zeros)
    k<<<1, 128>>>(tm, o, c0, c1, c2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<__half> res(16 * 8 * 32);
    cudaMemcpy(res.data(), o, res.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int s = 0; s < 16; ++s) for (int j = 0; j < 8; ++j) for (int kk = 0; kk < 32; ++kk) {
      int row = s * 8 + j;
      int phys = sw ? (kk ^ (((row >> 1) & 3) << 3)) : kk;          // 64-byte swizzle: 16-byte chunk index ^= (row/2)%4
      float got = __half2float(res[row * 32 + phys]);
      size_t src = (size_t)PF - 200 + (size_t)(c2 + s) * S + (size_t)(c1 + j) * STR + c0 + kk;
      float want = (c2 + s) < NS ? (float)(src % 2048) : 0.f;
      if (got != want && bad++ < 5) printf("  mismatch s=%d j=%d k=%d got %g want %g\n", s, j, kk, got, want);
    }
    printf("swizzle %d: %d mismatches\n", sw, bad);
  }
  return 0;
}
