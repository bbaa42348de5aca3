// microbench.cu -- dependent-chain latencies (cycles) of the FP64 operations the routing
// kernels sit on, measured on the device with clock64. Developer aid, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -shared -Xcompiler -fPIC -o scripts/microbench.so scripts/microbench.cu
//   python -c "import ctypes; ctypes.CDLL('scripts/microbench.so').run_microbench()"
#include <cstdio>
#include <cuda_runtime.h>

template <class F>
__device__ double chain(F f, double x, int n, long long& cyc) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = f(x);
  long long t1 = clock64();
  cyc = t1 - t0;
  return x;
}

__global__ void lat(double* out, long long* cyc, double seed, double* gmem) {
  __shared__ double sm[256];
  sm[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  const int n = 512;
  int k = 0;
  double x = seed;
  long long c;
  x = chain([](double v) { return v * 1.0000001; }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([](double v) { return v + 1.0000001; }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([](double v) { return fma(v, 0.999999, 0.5); }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([](double v) { return 3.0 / v + 1.0; }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([](double v) { return exp(v * 1e-3); }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([](double v) { return log(v + 2.0); }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([](double v) { return exp(0.2 * log(v + 2.0)); }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([](double v) { return cbrt(v + 2.0); }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  x = chain([&](double v) { return sm[((int)v) & 255]; }, x, n, c); if (!threadIdx.x) cyc[k] = c; k++;
  {  // __syncthreads round
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t1 = clock64();
    if (!threadIdx.x) cyc[k] = t1 - t0; k++;
  }
  {  // L2 round trip: ld.relaxed.gpu chain
    unsigned long long idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      unsigned long long v;
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"((unsigned long long*)gmem + idx) : "memory");
      idx = v & 1023;
    }
    long long t1 = clock64();
    if (!threadIdx.x) cyc[k] = t1 - t0; k++;
    x += (double)idx;
  }
  {  // one Newton iteration of kinematic_wave (dependent)
    double u = 0.5 + 1e-3 * seed, dt_dx = 3.6, alpha = 2.0, C = 1.7;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const double u2 = u * u, u3 = u2 * u;
      const double f_u = u3 * (dt_dx * u2 + alpha) - C;
      const double df_u = u2 * (5.0 * dt_dx * u2 + 3.0 * alpha);
      u -= f_u / df_u;
      u += 1e-9;  // keep it moving
    }
    long long t1 = clock64();
    if (!threadIdx.x) cyc[k] = t1 - t0; k++;
    x += u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (!threadIdx.x) cyc[31] = n;
}

// throughput: W warps per SM, each running a dependent chain (mode 0: dfma, 1: Newton iteration
// of kinematic_wave, 2: pow) with `lanes` active lanes per warp
__global__ void thr(double* out, double seed, int n, int mode, int lanes) {
  double x = seed + threadIdx.x * 1e-3;
  if ((threadIdx.x & 31) < lanes) {
    if (mode == 0) {
      for (int i = 0; i < n; ++i) x = fma(x, 0.999999, 0.5);
    } else if (mode == 1) {
      double u = 0.5 + 1e-3 * x, dt_dx = 3.6, alpha = 2.0, C = 1.7;
      for (int i = 0; i < n; ++i) {
        const double u2 = u * u, u3 = u2 * u;
        const double f_u = u3 * (dt_dx * u2 + alpha) - C;
        const double df_u = u2 * (5.0 * dt_dx * u2 + 3.0 * alpha);
        u -= f_u / df_u;
        u += 1e-9;
      }
      x = u;
    } else {
      for (int i = 0; i < n; ++i) x = exp(0.2 * log(x + 2.0));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

extern "C" int run_microbench() {
  double *out, *gmem; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&gmem, 8192); cudaMemset(gmem, 0, 8192);
  cudaMalloc(&cyc, 32 * sizeof(long long)); long long hc[32];
  const char* names[] = {"dmul", "dadd", "dfma", "ddiv+dadd", "exp", "log(+dadd)", "pow0.2(+dadd)", "cbrt(+dadd)",
                         "smem ld (+cvt)", "syncthreads", "L2 ld.relaxed.gpu", "newton iter"};
  for (int threads : {32, 256}) {
    lat<<<1, threads>>>(out, cyc, 1.5, gmem);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    printf("threads=%d (%s)\n", threads, cudaGetErrorString(cudaGetLastError()));
    for (int k = 0; k < 12; ++k) printf("  %-20s %8.1f cycles\n", names[k], (double)hc[k] / hc[31]);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* mn[] = {"dfma chain", "newton iter chain", "pow chain"};
  for (int mode = 0; mode < 3; ++mode)
    for (int lanes : {32, 4})
      for (int wps : {4, 8, 16, 24, 32}) {
        const int n = mode == 0 ? 20000 : 2000;
        thr<<<148, wps * 32>>>(out, 1.5, 10, mode, lanes);
        cudaEventRecord(e0);
        thr<<<148, wps * 32>>>(out, 1.5, n, mode, lanes);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-18s lanes=%2d warps/SM=%2d: %.1f cycles per chain step (at 1.965 GHz)\n", mn[mode], lanes, wps,
               ms * 1e-3 * 1.965e9 / n);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
