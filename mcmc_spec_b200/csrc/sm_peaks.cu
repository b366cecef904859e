// sm_peaks.cu -- measured ceilings of the units that bound the streaming kernels on this GPU (sm_100a):
//   * fp64 DFMA throughput (all SMs, many independent chains per thread) and dependent-issue latency,
//   * L2 -> SM read bandwidth with 256-bit loads from an L2-resident buffer (the shape of the grid-row gathers),
//   * shared-memory read bandwidth (128-bit conflict-free loads),
// each timed with CUDA events after warm-up, best of several repetitions.  Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sm_peaks sm_peaks.cu   (scripts/measure_sm_peaks.py)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__)); exit(1); } } while (0)

// ILP independent FMA chains per thread, ITER steps each
template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one warp, one dependent chain: cycles per DFMA
__global__ void dfma_latency_kernel(long long* cycles, double* out, int iters, double a, double b) {
  double x = 1.0 + threadIdx.x * 1e-6;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
    x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  if (x == 12345.678) *out = x;
}

struct __align__(32) d4 { double x, y, z, w; };
__global__ void __launch_bounds__(256) l2_read_kernel(const double* __restrict__ buf, size_t n_vec, int reps, double* out) {
  // every CTA sweeps the whole (L2-resident) buffer `reps` times with 32-byte loads, starting at its own offset
  double s = 0.0;
  const size_t stride = (size_t)blockDim.x;
  for (int r = 0; r < reps; ++r) {
    size_t i = ((size_t)blockIdx.x * 9973 * stride + threadIdx.x) % n_vec;
    for (size_t k = 0; k < n_vec / stride; k += 4) {
      d4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double* p = buf + 4 * ((i + (size_t)u * stride) % n_vec);
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[u].x), "=d"(v[u].y), "=d"(v[u].z), "=d"(v[u].w) : "l"(p));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) s += v[u].x + v[u].y + v[u].z + v[u].w;
      i = (i + 4 * stride) % n_vec;
    }
  }
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) smem_read_kernel(int iters, double* out) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  double s = 0.0;
  const double2* p = reinterpret_cast<const double2*>(sm);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) { const double2 v = p[(idx + u * 256) & 4095]; s += v.x + v.y; }
    idx = (idx + 32) & 4095;
  }
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static float best_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clock_khz = 0; CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  double* out; CK(cudaMalloc(&out, 1 << 24));
  long long* cyc; CK(cudaMalloc(&cyc, 8));

  // ---- DFMA throughput
  const int iters = 4096;
  auto run_dfma = [&](int ilp, int ctas_per_sm) {
    const int grid = sms * ctas_per_sm;
    float ms = 0;
    if (ilp == 1) ms = best_ms([&] { dfma_kernel<1><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    if (ilp == 2) ms = best_ms([&] { dfma_kernel<2><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    if (ilp == 4) ms = best_ms([&] { dfma_kernel<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    if (ilp == 8) ms = best_ms([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    return 2.0 * (double)grid * 256 * ilp * iters / (ms * 1e-3) / 1e12;       // TFLOP/s
  };
  double dfma_best = 0.0;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_nominal\": %d,\n \"dfma_tflops\": {", prop.name, sms, clock_khz / 1000);
  bool first = true;
  for (int cps : {1, 2, 4, 8}) for (int ilp : {1, 2, 4, 8}) {
    const double t = run_dfma(ilp, cps);
    dfma_best = std::max(dfma_best, t);
    printf("%s\"warps_per_sm=%d,ilp=%d\": %.2f", first ? "" : ", ", cps * 8, ilp, t);
    first = false;
  }
  printf("},\n \"dfma_tflops_peak\": %.2f,\n", dfma_best);

  // ---- DFMA latency
  dfma_latency_kernel<<<1, 32>>>(cyc, out, 2048, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
  dfma_latency_kernel<<<1, 32>>>(cyc, out, 2048, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
  long long h_cyc; CK(cudaMemcpy(&h_cyc, cyc, 8, cudaMemcpyDeviceToHost));
  printf(" \"dfma_dependent_latency_cycles\": %.2f,\n", (double)h_cyc / (2048.0 * 8));

  // ---- L2 -> SM read bandwidth: 32 MiB buffer (L2 resident), 256-bit loads
  {
    const size_t bytes = 32ull << 20;
    double* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    const size_t n_vec = bytes / 32;
    double best = 0.0;
    for (int cps : {2, 4, 8}) {
      const int grid = sms * cps, reps = 2;
      const float ms = best_ms([&] { l2_read_kernel<<<grid, 256>>>(buf, n_vec, reps, out); });
      const double per_cta = (double)(n_vec / 256 / 4) * 4 * 256 * 32 * reps;
      best = std::max(best, per_cta * grid / (ms * 1e-3) / 1e9);
    }
    printf(" \"l2_read_gbs\": %.0f,\n", best);
    CK(cudaFree(buf));
  }
  // ---- shared-memory read bandwidth
  {
    const int it = 4096;
    CK(cudaFuncSetAttribute(smem_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    double best = 0.0;
    for (int cps : {1, 2, 3}) {
      const int grid = sms * cps;
      const float ms = best_ms([&] { smem_read_kernel<<<grid, 256, 65536>>>(it, out); });
      best = std::max(best, (double)grid * 256 * it * 8 * 16 / (ms * 1e-3) / 1e9);
    }
    printf(" \"smem_read_gbs\": %.0f}\n", best);
  }
  return 0;
}
