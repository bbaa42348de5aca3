// Micro-benchmark: does the NUMBER OF CONCURRENT ADDRESS STREAMS bound a per-cell kernel on B200?
// Reads NI arrays and writes NO arrays of n doubles, one element per thread, 128 threads per CTA,
// (a) structure of arrays (the product's layout), (b) array of tiles (all arrays of one tile of 128
// cells contiguous), at full and at register-limited occupancy (dynamic shared memory pads).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NI = 60, NO = 50, T = 128;
template <bool TILED, int CHAIN, int PF = 0>
__global__ void __launch_bounds__(T) k(const double* __restrict__ in, double* __restrict__ out, long n) {
    long i = (long)blockIdx.x * T + threadIdx.x;
    if (i >= n) return;
    long ns = n;
    if (PF > 0 && !TILED) {
        // L2 prefetch of the inputs of the tile PF CTAs ahead: one bulk prefetch of 1 KB per array
        long j = ((long)blockIdx.x + PF) * T;
        if (j + T <= n && threadIdx.x < NI) {
            const double* p = in + (long)threadIdx.x * ns + j;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(T * 8) : "memory");
        }
    }
    double acc = 0.0;
    double v[NI];
#pragma unroll
    for (int a = 0; a < NI; ++a)
        v[a] = TILED ? in[((long)blockIdx.x * NI + a) * T + threadIdx.x] : in[(long)a * ns + i];
    if (CHAIN == 0) {
#pragma unroll
        for (int a = 0; a < NI; ++a) acc += v[a];
    } else {
        // dependent FP64 chain between uses, CHAIN operations per input
#pragma unroll
        for (int a = 0; a < NI; ++a) {
            acc += v[a];
#pragma unroll
            for (int c = 0; c < CHAIN; ++c) acc = fma(acc, 1.0000001, 1e-9);
        }
    }
#pragma unroll
    for (int a = 0; a < NO; ++a) {
        double r = acc + a;
        if (TILED) out[((long)blockIdx.x * NO + a) * T + threadIdx.x] = r;
        else out[(long)a * ns + i] = r;
    }
}
template <bool TILED, int CHAIN, int PF = 0>
void run(const char* name, const double* in, double* out, long n, int smem) {
    cudaFuncSetAttribute(k<TILED, CHAIN, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = (int)((n + T - 1) / T);
    for (int w = 0; w < 3; ++w) k<TILED, CHAIN, PF><<<grid, T, smem>>>(in, out, n);
    cudaEventRecord(e0);
    const int R = 10;
    for (int r = 0; r < R; ++r) k<TILED, CHAIN, PF><<<grid, T, smem>>>(in, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= R;
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<TILED, CHAIN, PF>, T, smem);
    printf("%-28s smem %6d  CTAs/SM %2d  %.1f us  %.0f GB/s  (%s)\n", name, smem, nb, ms * 1e3,
           (double)(NI + NO) * 8 * n / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long n = 1000064;
    double *in, *out;
    cudaMalloc(&in, sizeof(double) * NI * n); cudaMalloc(&out, sizeof(double) * NO * n);
    cudaMemset(in, 0, sizeof(double) * NI * n);
    int pads[] = {0, 44 * 1024, 56 * 1024};  // -> all / 5 / 4 CTAs per SM
    for (int s : pads) {
        run<false, 0>("SoA   loads first", in, out, n, s);
        run<true, 0>("tiled loads first", in, out, n, s);
        run<false, 8>("SoA   chain 8/input", in, out, n, s);
        run<true, 8>("tiled chain 8/input", in, out, n, s);
        run<false, 32>("SoA   chain 32/input", in, out, n, s);
        run<false, 32, 148>("SoA   chain 32 pf+148", in, out, n, s);
        run<false, 32, 370>("SoA   chain 32 pf+370", in, out, n, s);
        run<false, 32, 740>("SoA   chain 32 pf+740", in, out, n, s);
        run<false, 8, 370>("SoA   chain 8 pf+370", in, out, n, s);
        run<true, 32>("tiled chain 32/input", in, out, n, s);
    }
    return 0;
}
