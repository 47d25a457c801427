// Developer tool: L2 -> SM read bandwidth of this GPU for an L2-resident working set (coalesced 128-bit loads,
// 8 independent loads in flight per thread, the block -> data mapping rotated every sweep so that L1 cannot serve
// re-reads).  Buffer sizes are powers of two.  Gives the denominator for the "l2" roofline bench.py reports beside
// the HBM one.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -cudart shared
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(256) k_l2_read(const float4* __restrict__ buf, uint32_t mask, int sweeps, float* sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t n_vec = mask + 1;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int s = 0; s < sweeps; ++s) {
    const uint32_t rot = (uint32_t)s * 2654435761u;
    for (uint32_t i0 = t; i0 < n_vec; i0 += stride * 8) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(buf + ((i0 + k * stride + rot) & mask));
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123456.f) *sink = acc.x;
}

extern "C" __attribute__((visibility("default"))) double l2_read_gbs(int64_t bytes, int sweeps, int blocks_per_sm) {
  float4* buf;
  float* sink;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, bytes);
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t n_vec = (uint32_t)(bytes / 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_l2_read<<<sm * blocks_per_sm, 256>>>(buf, n_vec - 1, 2, sink);  // warm the L2
  cudaEventRecord(e0);
  k_l2_read<<<sm * blocks_per_sm, 256>>>(buf, n_vec - 1, sweeps, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(buf); cudaFree(sink);
  return (double)bytes * sweeps / (ms * 1e-3) / 1e9;
}
