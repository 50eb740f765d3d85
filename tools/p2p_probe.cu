// Microbenchmark: NVLink peer-memory access patterns from a kernel on GPU0 touching memory of GPU1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/p2p_probe tools/p2p_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <typename T, int U>
__global__ void rd(const T* __restrict__ src, T* __restrict__ dst, size_t n) {   // remote read -> local write
  size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
  T v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) if (i + (size_t)u * blockDim.x < n) v[u] = src[i + (size_t)u * blockDim.x];
#pragma unroll
  for (int u = 0; u < U; ++u) if (i + (size_t)u * blockDim.x < n) dst[i + (size_t)u * blockDim.x] = v[u];
}
template <typename T, int U>
__global__ void rmw(T* __restrict__ rem, const T* __restrict__ loc, size_t n) {  // remote read + remote write
  size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
  T v[U], w[U];
#pragma unroll
  for (int u = 0; u < U; ++u) if (i + (size_t)u * blockDim.x < n) { v[u] = rem[i + (size_t)u * blockDim.x]; w[u] = loc[i + (size_t)u * blockDim.x]; }
#pragma unroll
  for (int u = 0; u < U; ++u) if (i + (size_t)u * blockDim.x < n) rem[i + (size_t)u * blockDim.x] = v[u] + w[u];
}
__device__ inline double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

template <typename F>
float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  const size_t n = (size_t)1 << 27;   // doubles: 1 GiB
  double *l0, *r1, *l0b;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&r1, n * 8)); CK(cudaMemset(r1, 0, n * 8));
  CK(cudaSetDevice(0)); CK(cudaMalloc(&l0, n * 8)); CK(cudaMalloc(&l0b, n * 8)); CK(cudaMemset(l0, 0, n * 8)); CK(cudaMemset(l0b, 0, n * 8));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  const double gb = n * 8 / 1e9;
  float ms;
  ms = timeit([&] { rd<double, 1><<<(unsigned)(n / 256), 256>>>(l0b, l0, n); });
  printf("local  read->local write  8B x1 : %7.1f GB/s (read)\n", gb / ms * 1e3);
  ms = timeit([&] { rd<double, 1><<<(unsigned)(n / 256), 256>>>(r1, l0, n); });
  printf("remote read->local write  8B x1 : %7.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { rd<double, 4><<<(unsigned)(n / 1024), 256>>>(r1, l0, n); });
  printf("remote read->local write  8B x4 : %7.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { rd<double, 8><<<(unsigned)(n / 2048), 256>>>(r1, l0, n); });
  printf("remote read->local write  8B x8 : %7.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { rd<double2, 1><<<(unsigned)(n / 512), 256>>>((double2*)r1, (double2*)l0, n / 2); });
  printf("remote read->local write 16B x1 : %7.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { rd<double2, 4><<<(unsigned)(n / 2048), 256>>>((double2*)r1, (double2*)l0, n / 2); });
  printf("remote read->local write 16B x4 : %7.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { rd<double, 4><<<(unsigned)(n / 1024), 256>>>(l0, r1, n); });
  printf("local read->remote write  8B x4 : %7.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { rd<double2, 4><<<(unsigned)(n / 2048), 256>>>((double2*)l0, (double2*)r1, n / 2); });
  printf("local read->remote write 16B x4 : %7.1f GB/s\n", gb / ms * 1e3);
  ms = timeit([&] { rmw<double, 4><<<(unsigned)(n / 1024), 256>>>(r1, l0, n); });
  printf("remote read+remote write  8B x4 : %7.1f GB/s each way\n", gb / ms * 1e3);
  ms = timeit([&] { rmw<double2, 4><<<(unsigned)(n / 2048), 256>>>((double2*)r1, (double2*)l0, n / 2); });
  printf("remote read+remote write 16B x4 : %7.1f GB/s each way\n", gb / ms * 1e3);
  ms = timeit([&] { cudaMemcpyPeerAsync(l0, 0, r1, 1, n * 8, 0); });
  printf("cudaMemcpyPeer (pull)           : %7.1f GB/s\n", gb / ms * 1e3);
  return 0;
}
