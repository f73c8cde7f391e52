// solve_bench.cu — micro-benchmark (diagnostic, not product): cycles of ONE plane solve (plane_from_moments: the 3x3 closed-form
// eigen-solve of pwpp_math.cuh) when a single warp runs it alone on an SM, with 1 / 32 lanes active, and how the time grows
// when 2..8 warps of one SM partition solve concurrently. nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../patchwork-plusplus_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>

#include "pwpp_math.cuh"
using namespace pwpp;

__global__ void k_solve(int reps, int lanes, double* out, long long* cyc) {
  const int lane = threadIdx.x & 31;
  Moments m;
  m.n = 500 + lane;
  // moments of a roughly planar patch (shifted sums)
  m.s1[0] = 12.5 + lane * 0.01; m.s1[1] = -7.25; m.s1[2] = 0.75;
  m.s2[0] = 910.0 + lane; m.s2[1] = 33.0; m.s2[2] = 2.5; m.s2[3] = 640.0; m.s2[4] = -1.25; m.s2[5] = 0.35;
  double c[3] = {10.0, 3.0, -1.7};
  Plane pl;
  double acc = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  if (lane < lanes) {
    for (int r = 0; r < reps; ++r) {
      plane_from_moments(m, c, pl);
      acc += pl.normal[2] + pl.sv[2];
      m.s2[5] += 1e-3 * pl.normal[0];   // dependency between repetitions
    }
  }
  __syncwarp();
  const long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1 << 16);
  const int reps = 200;
  for (int lanes : {1, 32}) {
    for (int warps : {1, 2, 4, 8, 16, 32}) {
      k_solve<<<1, warps * 32>>>(reps, lanes, out, cyc);
      cudaDeviceSynchronize();
      k_solve<<<1, warps * 32>>>(reps, lanes, out, cyc);
      cudaDeviceSynchronize();
      long long h[64];
      cudaMemcpy(h, cyc, warps * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
      std::printf("lanes %2d warps/SM %2d: %.0f cycles per solve (slowest warp)\n", lanes, warps, (double) mx / reps);
    }
  }
  std::printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
