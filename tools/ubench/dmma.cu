// Microbenchmark: FP64 FMA pipe vs FP64 mma.sync (DMMA) throughput and issue-slot cost on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu ; run: ./dmma
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
    double acc[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = seed * (threadIdx.x + i);
    double a = seed + threadIdx.x, b = seed * 0.5;
    double av[8], bv[4];
    #pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i;
    #pragma unroll
    for (int i = 0; i < 4; ++i) bv[i] = b + i;
    __shared__ double sm[256 * 2];
    int ialu = threadIdx.x;
    sm[threadIdx.x] = seed; sm[threadIdx.x + 256] = seed;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // 16 independent DFMA
            #pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        } else if (MODE == 1) {  // 8 independent m8n8k4
            #pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
        } else if (MODE == 2) {  // 8 m8n8k4 + 16 integer ALU ops interleaved
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                dmma884(acc[2 * i], acc[2 * i + 1], a, b);
                ialu = ialu * 3 + i;
                ialu ^= (ialu >> 3);
            }
        } else if (MODE == 3) {  // 4 m16n8k8
            #pragma unroll
            for (int i = 0; i < 4; ++i) dmma1688(acc + 4 * i, av, bv);
        } else if (MODE == 4) {  // 4 m16n8k16
            #pragma unroll
            for (int i = 0; i < 4; ++i) dmma16816(acc + 4 * i, av, bv);
        } else if (MODE == 5) {  // 8 m8n8k4 + 8 smem round trips (LDS.128 + STS.128)
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                dmma884(acc[2 * i], acc[2 * i + 1], a, b);
                double2 v = *reinterpret_cast<double2*>(&sm[2 * ((threadIdx.x + i) & 255)]);
                v.x += 1.0;
                *reinterpret_cast<double2*>(&sm[2 * threadIdx.x]) = v;
            }
        } else if (MODE == 6) {  // 16 DFMA + 8 smem round trips
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[2 * i] = fma(acc[2 * i], a, b);
                acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
                double2 v = *reinterpret_cast<double2*>(&sm[2 * ((threadIdx.x + i) & 255)]);
                v.x += 1.0;
                *reinterpret_cast<double2*>(&sm[2 * threadIdx.x]) = v;
            }
        } else if (MODE == 7) {  // 16 DFMA + 16 integer ALU ops
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[2 * i] = fma(acc[2 * i], a, b);
                acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
                ialu = ialu * 3 + i;
                ialu ^= (ialu >> 3);
            }
        }
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + ialu + sm[threadIdx.x];
}

template <int MODE>
int run(const char* name, double fma_per_warp_iter, int ctas_per_sm) {
    int sms = 148, iters = 20000;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 256));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * ctas_per_sm, 256>>>(out, 100, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas_per_sm, 256>>>(out, iters, 1e-9);
    cudaEventRecord(e1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = double(sms) * ctas_per_sm * 8;
    double fma = warps * iters * fma_per_warp_iter;
    double cyc_per_iter_per_smsp = ms * 1e-3 * 1.965e9 / (iters * (ctas_per_sm * 8 / 4.0));
    printf("%-44s ctas/sm=%d  %8.3f ms  %7.2f TFLOP/s  %7.1f clk per warp-iteration per SMSP\n", name, ctas_per_sm, ms, 2 * fma / ms * 1e-9, cyc_per_iter_per_smsp);
    cudaFree(out);
    return 0;
}

int main() {
    for (int c : {1, 2}) {
        run<0>("16x DFMA", 16 * 32, c);
        run<1>("8x DMMA m8n8k4", 8 * 256, c);
        run<2>("8x DMMA m8n8k4 + 24 int ops", 8 * 256, c);
        run<7>("16x DFMA + 24 int ops", 16 * 32, c);
        run<3>("4x DMMA m16n8k8", 4 * 1024, c);
        run<4>("4x DMMA m16n8k16", 4 * 2048, c);
        run<5>("8x DMMA m8n8k4 + 8x(LDS.128+DADD+STS.128)", 8 * 256, c);
        run<6>("16x DFMA + 8x(LDS.128+DADD+STS.128)", 16 * 32, c);
    }
    return 0;
}
