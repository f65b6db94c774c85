// Micro-benchmark (not part of the product): a 7-point gather  y[i] = sum_k c_k x[i + o_k]  with 64-bit loads (one row per
// lane) vs 128-bit loads (two rows per lane): does the L1/LSU pipe handle fp64 warp gathers at half rate?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_width gather_width.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Offs { int o[7]; double c[7]; };

template <int NOFF>
__global__ void __launch_bounds__(256, 4) k64(const double* __restrict__ x, double* __restrict__ y, int n, int lo, Offs of) {
  const int ngroups = n / 2048;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
#pragma unroll 4
    for (int j = 0; j < 8; ++j) {
      const int i = g * 2048 + j * 256 + threadIdx.x;
      if (i < lo || i >= n - lo) continue;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < NOFF; ++k) s = fma(of.c[k], __ldg(x + i + of.o[k]), s);
      y[i] = s;
    }
  }
}

template <int NOFF>
__global__ void __launch_bounds__(256, 4) k128(const double* __restrict__ x, double* __restrict__ y, int n, int lo, Offs of) {
  const int ngroups = n / 2048;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
#pragma unroll 4
    for (int j = 0; j < 4; ++j) {
      const int i = g * 2048 + j * 512 + threadIdx.x * 2;
      if (i < lo || i >= n - lo) continue;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < NOFF; ++k) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(x + i + of.o[k]));
        s0 = fma(of.c[k], v.x, s0);
        s1 = fma(of.c[k], v.y, s1);
      }
      *reinterpret_cast<double2*>(y + i) = make_double2(s0, s1);
    }
  }
}

int main() {
  const int n = 1 << 24;
  double *x, *y;
  CHECK(cudaMalloc(&x, sizeof(double) * (size_t)n));
  CHECK(cudaMalloc(&y, sizeof(double) * (size_t)n));
  CHECK(cudaMemset(x, 0, sizeof(double) * (size_t)n));
  CHECK(cudaMemset(y, 0, sizeof(double) * (size_t)n));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int sets[4][7] = {{0, -2, 2, -256, 256, -65536, 65536},   // aligned +-2 instead of +-1
                          {0, -1, 1, -256, 256, -65536, 65536},   // the real 7-point pattern (64-bit only)
                          {0, -2, 2, -256, 256, 0, 0},            // near only (5 distinct + 2 repeats of the centre)
                          {0, 0, 0, 0, 0, 0, 0}};                 // 7 x the same line
  const char* names[4] = {"aligned(+-2,+-256,+-65536)", "real(+-1,+-256,+-65536)", "near(+-2,+-256,0,0)", "centre x7"};
  for (int ctas = 2; ctas <= 6; ctas += 2) {
    const int grid = 148 * ctas;
    for (int sidx = 0; sidx < 4; ++sidx) {
      Offs of;
      for (int k = 0; k < 7; ++k) { of.o[k] = sets[sidx][k]; of.c[k] = 1.0 + k; }
      for (int width = 64; width <= 128; width += 64) {
        if (width == 128 && sidx == 1) continue;
        for (int noff = 1; noff <= 7; noff += 6) {
          float best = 1e30f;
          for (int rep = 0; rep < 8; ++rep) {
            cudaEventRecord(e0);
            if (width == 64) { if (noff == 7) k64<7><<<grid, 256>>>(x, y, n, 65536, of); else k64<1><<<grid, 256>>>(x, y, n, 65536, of); }
            else { if (noff == 7) k128<7><<<grid, 256>>>(x, y, n, 65536, of); else k128<1><<<grid, 256>>>(x, y, n, 65536, of); }
            cudaEventRecord(e1);
            CHECK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
          }
          printf("{\"what\": \"gather_width\", \"ctas_per_sm\": %d, \"pattern\": \"%s\", \"load_bits\": %d, \"gathers\": %d, \"us\": %.2f}\n",
                 ctas, names[sidx], width, noff, best * 1e3f);
        }
      }
    }
  }
  return 0;
}
