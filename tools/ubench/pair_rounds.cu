// Microbenchmark: "pair rounds" — a shared-memory tile of complex128 amplitudes, each round loads every amplitude with a
// lane mapping that puts a qubit pair on lane bits 0,1, applies a 4x4 complex gate with two DMMA.8x8x4 and stores it back.
// Compared with the register-bit scheme (4 merged 2x2 gates per round trip, DFMA).  No global traffic: measures the
// on-chip part only.  clk/amp/gate is per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int M, int R>  // tile 2^M amps, 2^R per thread
__global__ void __launch_bounds__(1 << (M - R)) pair_rounds(double2* out, const double* gates, int rounds) {
    extern __shared__ double2 tile[];
    constexpr int T = 1 << (M - R);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = 0; e < (1 << R); ++e) tile[tid + e * T] = make_double2(1.0 / (1 + tid + e), 0.5);
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
        // pair position p = rd % (M-1): lane bits 0,1 <- tile bits p, p+1; the other thread bits fill the remaining low bits
        const int p = rd % (M - 1);
        const double b0 = gates[(rd & 15) * 64 + lane], b1 = gates[(rd & 15) * 64 + 32 + lane];
        // spread tid bits (above bit 1) and register bits around the pair position
        const unsigned lowmask = (1u << p) - 1;
        double2 v[1 << R];
        int addr[1 << R];
        #pragma unroll
        for (int e = 0; e < (1 << R); ++e) {
            unsigned rest = (tid >> 2) | (e << (M - R - 2));           // M-2 bits
            unsigned idx = (rest & lowmask) | ((tid & 3) << p) | ((rest & ~lowmask) << 2);
            if (p >= 2) idx ^= ((idx >> p) & 3) << 1;                 // per-round conflict-free placement (timing only)
            addr[e] = idx;
            v[e] = tile[idx];
        }
        #pragma unroll
        for (int e = 0; e < (1 << R); ++e) {
            double d0, d1;
            dmma(d0, d1, v[e].x, b0, 0.0, 0.0);
            dmma(d0, d1, v[e].y, b1, d0, d1);
            tile[addr[e]] = make_double2(d0, d1);
        }
        __syncthreads();
    }
    double2 s = make_double2(0, 0);
    for (int e = 0; e < (1 << R); ++e) { s.x += tile[tid + e * T].x; s.y += tile[tid + e * T].y; }
    out[blockIdx.x * T + tid] = s;
}

template <int M, int R>  // register-bit scheme: per round R merged 2x2 gates (one per register bit) with DFMA
__global__ void __launch_bounds__(1 << (M - R)) reg_rounds(double2* out, const double* gates, int rounds) {
    extern __shared__ double2 tile[];
    constexpr int T = 1 << (M - R);
    const int tid = threadIdx.x;
    for (int e = 0; e < (1 << R); ++e) tile[tid + e * T] = make_double2(1.0 / (1 + tid + e), 0.5);
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
        const int p = 3 + rd % (M - R - 2);
        const unsigned lowmask = (1u << p) - 1;
        double2 v[1 << R];
        int addr[1 << R];
        #pragma unroll
        for (int e = 0; e < (1 << R); ++e) {
            unsigned idx = (tid & lowmask) | (e << p) | ((tid & ~lowmask) << R);
            addr[e] = idx;
            v[e] = tile[idx];
        }
        #pragma unroll
        for (int j = 0; j < R; ++j) {
            const double* g = gates + ((rd + j) & 15) * 64;
            const double ar = g[0], ai = g[1], br = g[2], bi = g[3], cr = g[4], ci = g[5], dr = g[6], di = g[7];
            #pragma unroll
            for (int e = 0; e < (1 << R); ++e) {
                if (e & (1 << j)) continue;
                double2 x = v[e], y = v[e | (1 << j)];
                v[e].x = ar * x.x - ai * x.y + br * y.x - bi * y.y;
                v[e].y = ar * x.y + ai * x.x + br * y.y + bi * y.x;
                v[e | (1 << j)].x = cr * x.x - ci * x.y + dr * y.x - di * y.y;
                v[e | (1 << j)].y = cr * x.y + ci * x.x + dr * y.y + di * y.x;
            }
        }
        #pragma unroll
        for (int e = 0; e < (1 << R); ++e) tile[addr[e]] = v[e];
        __syncthreads();
    }
    double2 s = make_double2(0, 0);
    for (int e = 0; e < (1 << R); ++e) { s.x += tile[tid + e * T].x; s.y += tile[tid + e * T].y; }
    out[blockIdx.x * T + tid] = s;
}

template <typename K>
int run(const char* name, K kern, int M, int R, int gates_per_round, int ctas_per_sm) {
    int sms = 148, rounds = 2000, T = 1 << (M - R);
    size_t smem = sizeof(double2) << M;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    if (occ < ctas_per_sm) { printf("%-40s M=%d R=%d: occupancy %d < %d, skipped\n", name, M, R, occ, ctas_per_sm); return 0; }
    double2* out; double* gates;
    CK(cudaMalloc(&out, sizeof(double2) * sms * ctas_per_sm * T));
    CK(cudaMalloc(&gates, sizeof(double) * 16 * 64));
    double h[16 * 64];
    for (int i = 0; i < 16 * 64; ++i) h[i] = 0.35 * ((i * 37) % 11 - 5) / 5.0;
    CK(cudaMemcpy(gates, h, sizeof(h), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<sms * ctas_per_sm, T, smem>>>(out, gates, 10);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    kern<<<sms * ctas_per_sm, T, smem>>>(out, gates, rounds);
    cudaEventRecord(e1);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double clk = ms * 1e-3 * 1.965e9;
    double amps_per_sm = double(ctas_per_sm) * (1 << M);
    printf("%-40s M=%d R=%d ctas/sm=%d (max %d): %8.3f ms  %6.3f clk/amp/round  %6.3f clk/amp/gate (per SM)\n", name, M, R, ctas_per_sm, occ, ms,
           clk / rounds / amps_per_sm, clk / rounds / amps_per_sm / gates_per_round);
    cudaFree(out); cudaFree(gates);
    return 0;
}

int main() {
    for (int c : {1, 2, 3, 4}) {
        run("pair rounds (2 gates/round, DMMA)", pair_rounds<12, 4>, 12, 4, 2, c);
        run("pair rounds (2 gates/round, DMMA)", pair_rounds<11, 3>, 11, 3, 2, c);
        run("pair rounds (2 gates/round, DMMA)", pair_rounds<11, 4>, 11, 4, 2, c);
        run("register rounds (4 gates/round, DFMA)", reg_rounds<11, 4>, 11, 4, 4, c);
        run("register rounds (3 gates/round, DFMA)", reg_rounds<10, 3>, 10, 3, 3, c);
    }
    return 0;
}
