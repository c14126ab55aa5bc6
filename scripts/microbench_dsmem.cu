// Microbenchmark: random 8-byte gathers from a table distributed over the shared memory of a thread-block
// cluster (DSMEM, ld.shared::cluster) vs the same gathers from local shared memory.  nvcc -arch=sm_100a -O3.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int CLUSTER, int U, bool LOCAL>
__global__ void dsmem_gather(int part_elems, int iters, double* out) {
  extern __shared__ double part[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = cluster.block_rank();
  for (int i = threadIdx.x; i < part_elems; i += blockDim.x) part[i] = (double)(rank * part_elems + i);
  cluster.sync();
  const uint32_t total = (uint32_t)part_elems * CLUSTER;
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
    double v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      seed = hash32(seed + j);
      const uint32_t idx = seed % total;
      const int r = LOCAL ? rank : (int)(idx / part_elems);
      const int o = idx % part_elems;
      const double* p = cluster.map_shared_rank(part, r);
      v[j] = p[o];
    }
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j];
  }
  cluster.sync();
  if (acc == 1.2345) out[0] = acc;
}

template <int CLUSTER, bool LOCAL>
void run(int threads, int part_kb) {
  const int part_elems = part_kb * 1024 / 8;
  const int iters = 2000;
  constexpr int U = 8;
  double* out; cudaMalloc(&out, 8);
  auto kern = dsmem_gather<CLUSTER, U, LOCAL>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, part_kb * 1024);
  if (CLUSTER > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int nclusters = 148 / CLUSTER;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CLUSTER); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = part_kb * 1024;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaError_t err = cudaLaunchKernelEx(&cfg, kern, part_elems, 10, out);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  err = cudaLaunchKernelEx(&cfg, kern, part_elems, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double gathers = (double)nclusters * CLUSTER * threads * iters * U;
  printf("cluster %2d %s part %3d KB (table %5d KB) %4d thr/CTA: %8.1f Ggather/s = %.2f gathers/ns/SM  (%s)\n", CLUSTER,
         LOCAL ? "LOCAL smem" : "DSMEM     ", part_kb, part_kb * CLUSTER, threads, gathers / ms / 1e6,
         gathers / ms / 1e6 / (nclusters * CLUSTER), cudaGetErrorString(err == cudaSuccess ? cudaGetLastError() : err));
  cudaFree(out);
}

int main() {
  run<8, true>(512, 100);
  run<8, false>(512, 100);
  run<8, false>(1024, 100);
  run<4, false>(512, 200);
  run<16, false>(512, 50);
  run<16, false>(1024, 50);
  run<2, false>(512, 100);
  return 0;
}
