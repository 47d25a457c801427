// Developer tool: random row gather from an L2-resident slab, two ways.
//   A  register path of the product kernels: 8 lanes per row, 128-bit LDG, 4 steps (12 loads) in flight per lane
//   B  TMA-style bulk copies: one `cp.async.bulk` (UBLKCP) per row into a per-warp shared-memory ring, completion on an
//      mbarrier, rows consumed with LDS.128
// Shape of the proteins layer: rows of `row_floats` (= 80) floats inside table rows of `ld` (= 480) floats, N rows,
// `n_idx` random indices.  Every gathered float is summed (one FADD per float) so that the loads cannot be elided.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -cudart shared
//   C  TMA tile::gather4: one `cp.async.bulk.tensor.2d...tile::gather4` (UTMALDG) per FOUR rows through a tensor map
//      over the (N x ld) table, box = row_floats columns, into the same per-warp ring (variant 5..7)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kWarps = 4;

__global__ void __launch_bounds__(kWarps * 32, 4)
k_gather_ldg(const float* __restrict__ table, int ld, int n_vec, const int* __restrict__ idx, int64_t n_idx, float* sink) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, l8 = lane & 7;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kWarps;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = warp * 32; base < n_idx; base += n_warps * 32) {
    const int mine = base + lane < n_idx ? __ldg(idx + base + lane) : 0;
#pragma unroll
    for (int e = 0; e < 32; e += 16) {
      float4 x[4][3];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int r = __shfl_sync(0xffffffffu, mine, e + s * 4 + grp);
        const float4* p = reinterpret_cast<const float4*>(table + (int64_t)r * ld);
#pragma unroll
        for (int i = 0; i < 3; ++i) x[s][i] = __ldg(p + min(l8 + 8 * i, n_vec - 1));
      }
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < 3; ++i) { acc.x += x[s][i].x; acc.y += x[s][i].y; acc.z += x[s][i].z; acc.w += x[s][i].w; }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123456.f) *sink = acc.x;
}

// 256-bit loads (LDG.E.ENL2.256): 4 lanes per 128-byte line, 3 slots cover the 320-byte row, 8 rows per step, 2 steps in flight
__global__ void __launch_bounds__(kWarps * 32, 4)
k_gather_ldg256(const float* __restrict__ table, int ld, int n_vec8, const int* __restrict__ idx, int64_t n_idx, float* sink) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, l4 = lane & 3;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kWarps;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t base = warp * 32; base < n_idx; base += n_warps * 32) {
    const int mine = base + lane < n_idx ? __ldg(idx + base + lane) : 0;
#pragma unroll
    for (int e = 0; e < 32; e += 16) {
      float x[2][3][8];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int r = __shfl_sync(0xffffffffu, mine, e + s * 8 + grp);
        const float* p = table + (int64_t)r * ld;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float* q = p + 8 * min(l4 + 4 * i, n_vec8 - 1);
          asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=f"(x[s][i][0]), "=f"(x[s][i][1]), "=f"(x[s][i][2]), "=f"(x[s][i][3]), "=f"(x[s][i][4]),
                         "=f"(x[s][i][5]), "=f"(x[s][i][6]), "=f"(x[s][i][7])
                       : "l"(q));
        }
      }
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] += x[s][i][c];
    }
  }
  float t = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) t += acc[c];
  if (t == 123456.f) *sink = t;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ROWS rows per stage, STAGES stages per warp
template <int ROWS, int STAGES>
__global__ void __launch_bounds__(kWarps * 32)
k_gather_bulk(const float* __restrict__ table, int ld, int row_floats, const int* __restrict__ idx, int64_t n_idx, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row_bytes = row_floats * 4;
  unsigned char* ring = smem + (size_t)w * STAGES * ROWS * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWarps * STAGES * ROWS * row_bytes) + w * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t warp = (int64_t)blockIdx.x * kWarps + w, n_warps = (int64_t)gridDim.x * kWarps;
  const int64_t n_chunks = (n_idx + ROWS - 1) / ROWS;
  const int64_t my_chunks = warp < n_chunks ? (n_chunks - warp + n_warps - 1) / n_warps : 0;
  auto issue = [&](int64_t c) {  // chunk number c of this warp -> stage c % STAGES
    const int st = (int)(c % STAGES);
    const int64_t base = (warp + c * n_warps) * ROWS;
    const int cnt = (int)min((int64_t)ROWS, n_idx - base);
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + st)), "r"(cnt * row_bytes) : "memory");
    __syncwarp();
    for (int j = lane; j < cnt; j += 32) {
      const int r = __ldg(idx + base + j);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                       "r"(smem_u32(ring + ((size_t)st * ROWS + j) * row_bytes)), "l"(table + (int64_t)r * ld), "r"(row_bytes),
                   "r"(smem_u32(bars + st))
                   : "memory");
    }
  };
  for (int64_t c = 0; c < STAGES - 1 && c < my_chunks; ++c) issue(c);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int vec_per_row = row_bytes / 16;
  for (int64_t c = 0; c < my_chunks; ++c) {
    if (c + STAGES - 1 < my_chunks) issue(c + STAGES - 1);
    const int st = (int)(c % STAGES);
    const uint32_t parity = (uint32_t)((c / STAGES) & 1);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bars + st)), "r"(parity) : "memory");
    const int64_t base = (warp + c * n_warps) * ROWS;
    const int cnt = (int)min((int64_t)ROWS, n_idx - base);
    const float4* rows = reinterpret_cast<const float4*>(ring + (size_t)st * ROWS * row_bytes);
    for (int v = lane; v < cnt * vec_per_row; v += 32) {  // consecutive lanes read consecutive 16-byte vectors
      const float4 x = rows[v];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    __syncwarp();  // every lane is done with the stage before it is refilled
  }
  if (acc.x + acc.y + acc.z + acc.w == 123456.f) *sink = acc.x;
}

template <int ROWS, int STAGES>
static float run_bulk(const float* table, int ld, int row_floats, const int* idx, int64_t n_idx, float* sink, int blocks) {
  const size_t bytes = (size_t)kWarps * STAGES * ROWS * row_floats * 4 + kWarps * STAGES * 8;
  cudaFuncSetAttribute(k_gather_bulk<ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_gather_bulk<ROWS, STAGES><<<blocks, kWarps * 32, bytes>>>(table, ld, row_floats, idx, n_idx, sink);
  cudaEventRecord(e0);
  k_gather_bulk<ROWS, STAGES><<<blocks, kWarps * 32, bytes>>>(table, ld, row_floats, idx, n_idx, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (cudaGetLastError() != cudaSuccess) return -1.f;
  return ms;
}

// ---- tile::gather4 ------------------------------------------------------------------------------------------------
// ROWS (multiple of 4) rows per stage, STAGES stages per warp.  Row j of a stage lands at ring + j * row_bytes (a
// gather4 writes its four rows back to back).  col0 = first column of the slab inside the table row.
template <int ROWS, int STAGES>
__global__ void __launch_bounds__(kWarps * 32)
k_gather_g4(const __grid_constant__ CUtensorMap tmap, int col0, int row_floats, const int* __restrict__ idx, int64_t n_idx,
            float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row_bytes = row_floats * 4;
  unsigned char* ring = smem + (size_t)w * STAGES * ROWS * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWarps * STAGES * ROWS * row_bytes) + w * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t warp = (int64_t)blockIdx.x * kWarps + w, n_warps = (int64_t)gridDim.x * kWarps;
  const int64_t n_chunks = n_idx / ROWS;  // whole chunks only (microbenchmark)
  const int64_t my_chunks = warp < n_chunks ? (n_chunks - warp + n_warps - 1) / n_warps : 0;
  auto issue = [&](int64_t c) {
    const int st = (int)(c % STAGES);
    const int64_t base = (warp + c * n_warps) * ROWS;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + st)), "r"(ROWS * row_bytes) : "memory");
    __syncwarp();
    if (lane < ROWS / 4) {
      const int4 r = __ldg(reinterpret_cast<const int4*>(idx + base) + lane);
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
              "r"(smem_u32(ring + ((size_t)st * ROWS + lane * 4) * row_bytes)), "l"(&tmap), "r"(col0), "r"(r.x), "r"(r.y), "r"(r.z),
          "r"(r.w), "r"(smem_u32(bars + st))
          : "memory");
    }
  };
  for (int64_t c = 0; c < STAGES - 1 && c < my_chunks; ++c) issue(c);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int vec_per_row = row_bytes / 16;
  for (int64_t c = 0; c < my_chunks; ++c) {
    if (c + STAGES - 1 < my_chunks) issue(c + STAGES - 1);
    const int st = (int)(c % STAGES);
    const uint32_t parity = (uint32_t)((c / STAGES) & 1);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bars + st)), "r"(parity) : "memory");
    const float4* rows = reinterpret_cast<const float4*>(ring + (size_t)st * ROWS * row_bytes);
    for (int v = lane; v < ROWS * vec_per_row; v += 32) {
      const float4 x = rows[v];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    __syncwarp();
  }
  // the sum leaves the kernel so that the host can check it against the LDG variant's
  acc.x += acc.y + acc.z + acc.w;
  for (int o = 16; o > 0; o >>= 1) acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
  if (lane == 0) atomicAdd(sink, acc.x);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS, int STAGES>
static float run_g4(const float* table, int ld, int row_floats, int64_t n_rows, const int* idx, int64_t n_idx, float* sink,
                    int blocks, int box_rows) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -2.f;
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)n_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)row_floats, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)table, gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -3.f - (float)r;
  const size_t bytes = (size_t)kWarps * STAGES * ROWS * row_floats * 4 + kWarps * STAGES * 8;
  cudaFuncSetAttribute(k_gather_g4<ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_gather_g4<ROWS, STAGES><<<blocks, kWarps * 32, bytes>>>(tmap, 0, row_floats, idx, n_idx, sink);
  cudaMemset(sink, 0, 4);
  cudaEventRecord(e0);
  k_gather_g4<ROWS, STAGES><<<blocks, kWarps * 32, bytes>>>(tmap, 0, row_floats, idx, n_idx, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { fprintf(stderr, "gather4: %s\n", cudaGetErrorString(err)); return -1.f; }
  return ms;
}

// sum of the gathered floats as the gather4 kernel accumulates it (validation against a host/torch sum)
extern "C" __attribute__((visibility("default"))) double gather4_checksum(const float* table, int ld, int row_floats, int64_t n_rows,
                                                                          const int* idx, int64_t n_idx, int blocks, int box_rows) {
  float* sink;
  cudaMalloc(&sink, 4);
  cudaMemset(sink, 0, 4);
  float ms = run_g4<16, 3>(table, ld, row_floats, n_rows, idx, n_idx, sink, blocks, box_rows);
  float h = 0.f;
  cudaMemcpy(&h, sink, 4, cudaMemcpyDeviceToHost);
  cudaFree(sink);
  return ms < 0 ? (double)ms * 1e30 : (double)h;
}

extern "C" __attribute__((visibility("default"))) double gather4_ms(const float* table, int ld, int row_floats, int64_t n_rows,
                                                                    const int* idx, int64_t n_idx, int variant, int blocks, int box_rows) {
  float* sink;
  cudaMalloc(&sink, 4);
  float ms = -1.f;
  if (variant == 5) ms = run_g4<16, 3>(table, ld, row_floats, n_rows, idx, n_idx, sink, blocks, box_rows);
  else if (variant == 6) ms = run_g4<32, 2>(table, ld, row_floats, n_rows, idx, n_idx, sink, blocks, box_rows);
  else if (variant == 7) ms = run_g4<8, 4>(table, ld, row_floats, n_rows, idx, n_idx, sink, blocks, box_rows);
  else if (variant == 8) ms = run_g4<16, 2>(table, ld, row_floats, n_rows, idx, n_idx, sink, blocks, box_rows);
  cudaFree(sink);
  return (double)ms;
}

// variant: 0 = LDG; 1 = bulk 16 rows x 3 stages; 2 = bulk 32 rows x 2 stages; 3 = bulk 8 rows x 4 stages
extern "C" __attribute__((visibility("default"))) double gather_ms(const float* table, int ld, int row_floats, const int* idx,
                                                                   int64_t n_idx, int variant, int blocks) {
  float* sink;
  cudaMalloc(&sink, 4);
  float ms = -1.f;
  if (variant == 0) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather_ldg<<<blocks, kWarps * 32>>>(table, ld, row_floats / 4, idx, n_idx, sink);
    cudaEventRecord(e0);
    k_gather_ldg<<<blocks, kWarps * 32>>>(table, ld, row_floats / 4, idx, n_idx, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  } else if (variant == 4) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather_ldg256<<<blocks, kWarps * 32>>>(table, ld, row_floats / 8, idx, n_idx, sink);
    cudaEventRecord(e0);
    k_gather_ldg256<<<blocks, kWarps * 32>>>(table, ld, row_floats / 8, idx, n_idx, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  } else if (variant == 1) ms = run_bulk<16, 3>(table, ld, row_floats, idx, n_idx, sink, blocks);
  else if (variant == 2) ms = run_bulk<32, 2>(table, ld, row_floats, idx, n_idx, sink, blocks);
  else if (variant == 3) ms = run_bulk<8, 4>(table, ld, row_floats, idx, n_idx, sink, blocks);
  cudaFree(sink);
  return (double)ms;
}
