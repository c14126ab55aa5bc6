// Microbenchmark: how fast can B200 serve random 8/16/32-byte gathers out of L2, alone and mixed with a
// streaming read?  Sets the practical ceiling for gather-per-nonzero SpMV.  nvcc -arch=sm_100a -O3.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <typename V, int U>
__global__ void gather_kernel(const V* __restrict__ table, uint32_t nelem, const int* __restrict__ idx, size_t n, double* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    int c[U]; V v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) c[j] = idx[i + j * stride];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = table[c[j]];
#pragma unroll
    for (int j = 0; j < U; ++j) { const double* d = (const double*)&v[j]; for (int t = 0; t < (int)(sizeof(V) / 8); ++t) acc += d[t]; }
  }
  if (acc == 12345.678) out[0] = acc;
}

template <int U>
__global__ void stream_kernel(const double2* __restrict__ a, size_t n, double* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    double2 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = a[i + j * stride];
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j].x + v[j].y;
  }
  if (acc == 12345.678) out[0] = acc;
}

struct alignas(32) Q4 { double v[4]; };

template <typename V>
void run(const char* name, size_t table_bytes, int threads, int blocks_per_sm) {
  const size_t n = 40u << 20;   // gathers per launch
  const uint32_t nelem = table_bytes / sizeof(V);
  V* table; int* idx; double* out;
  cudaMalloc(&table, table_bytes); cudaMemset(table, 0, table_bytes);
  cudaMalloc(&idx, n * sizeof(int)); cudaMalloc(&out, 8);
  std::vector<int> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = hash32((uint32_t)i * 2654435761u + 17) % nelem;
  cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * blocks_per_sm;
  for (int rep = 0; rep < 3; ++rep) gather_kernel<V, 8><<<grid, threads>>>(table, nelem, idx, n, out);
  cudaEventRecord(e0);
  for (int rep = 0; rep < 5; ++rep) gather_kernel<V, 8><<<grid, threads>>>(table, nelem, idx, n, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-28s table %7.2f MB  %d thr x %d/SM: %7.2f Ggather/s  (idx stream %.0f GB/s, sectors %.0f GB/s)  err=%s\n", name,
         table_bytes / 1e6, threads, blocks_per_sm, n / ms / 1e6, n * 4 / ms / 1e6, n * 32.0 / ms / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(table); cudaFree(idx); cudaFree(out);
}

int main() {
  // pure streaming reference
  {
    const size_t n = (size_t)1 << 27;  // 2 GiB of double2
    double2* a; double* out; cudaMalloc(&a, n * 16); cudaMemset(a, 0, n * 16); cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    stream_kernel<8><<<148 * 8, 256>>>(a, n, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) stream_kernel<8><<<148 * 8, 256>>>(a, n, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("stream read 2 GiB: %.0f GB/s\n", n * 16 / ms / 1e6);
    // L2-resident stream (64 MB)
    const size_t n2 = (size_t)1 << 22;
    for (int r = 0; r < 3; ++r) stream_kernel<8><<<148 * 8, 256>>>(a, n2, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 50; ++r) stream_kernel<8><<<148 * 8, 256>>>(a, n2, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); ms /= 50;
    printf("stream read 64 MiB (L2 resident): %.0f GB/s\n", n2 * 16 / ms / 1e6);
    cudaFree(a); cudaFree(out);
  }
  for (int bps : {4, 8}) {
    run<double>("gather 8B", 800e3, 256, bps);
    run<double>("gather 8B", 1.6e6, 256, bps);
    run<double2>("gather 16B", 3.2e6, 256, bps);
    run<Q4>("gather 32B", 6.4e6, 256, bps);
    run<double>("gather 8B", 32e6, 256, bps);
    run<double>("gather 8B (HBM)", 1024e6, 256, bps);
  }
  return 0;
}
