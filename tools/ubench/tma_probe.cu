// Probe: how does a TMA tensor-map load with SWIZZLE_128B lay a box out in shared memory when the box's inner row is only
// 64 bytes?  Hypothesis H1: rows are packed densely (row-major box order) and the 16-byte chunk index (address bits 4..6)
// is XOR-ed with address bits 7..9.  The lean sweep kernel's "layout A" depends on the answer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ubench/tma_probe tools/ubench/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint64_t* out, int n_elems, int c1, int c4) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned sm = (unsigned)__cvta_generic_to_shared(smem), mb = (unsigned)__cvta_generic_to_shared(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(n_elems * 8));
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(sm),
                 "l"(&tm), "r"(mb), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(c4)
                 : "memory");
  }
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(mb) : "memory");
  }
  for (int i = threadIdx.x; i < n_elems; i += blockDim.x) out[i] = reinterpret_cast<uint64_t*>(smem)[i];
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  // complex128 state of n = 20 qubits, batch 2, as 8-byte elements: tile bits {0,1} + {11..19}  (lean sweep of cfg2)
  const int n = 20, batch = 2;
  const size_t total = ((size_t)batch << n) * 2;
  std::vector<uint64_t> h(total);
  for (size_t i = 0; i < total; ++i) h[i] = i;
  uint64_t *d, *dout;
  cudaMalloc(&d, total * 8);
  cudaMemcpy(d, h.data(), total * 8, cudaMemcpyHostToDevice);
  const int n_elems = 4096;  // 2^11 amplitudes x 2
  cudaMalloc(&dout, n_elems * 8);
  for (int mode = 0; mode < 2; ++mode) {
    cuuint64_t size[5] = {8, 512, 256, 2, (cuuint64_t)batch};
    cuuint64_t stride[4] = {16ull << 2, 16ull << 11, 16ull << 19, 16ull << 20};
    cuuint32_t box[5] = {8, 1, 256, 2, 1}, es[5] = {1, 1, 1, 1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mode ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %s: encode rc=%d\n", mode ? "SWIZZLE_128B" : "SWIZZLE_NONE", (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int c1 = 37, c4 = 1;
    cudaMemset(dout, 0xff, n_elems * 8);
    probe<<<1, 128, n_elems * 8 + 1024, 0>>>(tm, dout, n_elems, c1, c4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint64_t> o(n_elems);
    cudaMemcpy(o.data(), dout, n_elems * 8, cudaMemcpyDeviceToHost);
    // dense box order: k = e0 + 8 * (w0 + 256 * w1); global element index of that box element:
    long bad_dense = 0, bad_h1 = 0;
    for (int k = 0; k < n_elems; ++k) {
      const int e0 = k & 7, w0 = (k >> 3) & 255, w1 = k >> 11;
      const uint64_t want = (uint64_t)e0 + ((uint64_t)c1 << 3) + ((uint64_t)w0 << 12) + ((uint64_t)w1 << 20) + ((uint64_t)c4 << 21);
      const unsigned off = k * 8, sw = off ^ (((off >> 7) & 7u) << 4);
      if (o[off / 8] != want) ++bad_dense;
      if (o[sw / 8] != want) ++bad_h1;
    }
    printf("  mismatches vs dense layout: %ld   vs dense + XOR(bits 4..6 ^= bits 7..9): %ld   (of %d)\n", bad_dense, bad_h1, n_elems);
    printf("  first 20 smem elements:");
    for (int i = 0; i < 20; ++i) printf(" %llu", (unsigned long long)o[i]);
    printf("\n");
  }
  return 0;
}
