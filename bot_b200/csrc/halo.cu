// Halo exchange of the 1-D destination-row partition through NVLink peer memory (sm_100a).
//
// The reference is single-GPU; the partitioned layer needs, per layer and direction, ONE
// exchange of the row-sharded source table [ft | el].  Instead of an NCCL all-gather /
// reduce-scatter of whole tables these two kernels move a COLUMN RANGE of the table straight between the ranks'
// buffers with peer loads (every rank maps the others' buffers: CUDA VMM / torch symmetric memory):
//
//   pull        table[r * rows + i, c0:c0+w] = shard_r[i, c0:c0+w]          for every rank r     (forward)
//   pull_reduce out[i, c0:c0+w] = sum_r table_r[me * rows + i, c0:c0+w]     fixed order r = 0..P-1 (backward)
//
// A column range = a range of heads, so the exchange of head range k+1 overlaps the gather kernel of head range k
// (the kernels work head-major anyway) without any strided repacking pass, and the reduction order is fixed: the
// partitioned gradients are run-to-run deterministic, which a ring reduce-scatter does not promise.
//
// Bandwidth: NVLink 5 moves 900 GB/s per direction and GPU at ~2-3 us latency, i.e. ~2.5 MB must be in flight:
// every thread keeps kUnroll 128-bit peer loads outstanding; the grid is a parameter (the exchange shares the GPU
// with the gather kernel it overlaps).
#include "common.cuh"

namespace botgat {

constexpr int kMaxWorld = 16;
constexpr int kHaloThreads = 512;
constexpr int kUnroll = 8;

struct PeerPtrs {
  const float* p[kMaxWorld];
};

__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 r;
  // peer memory is written by another GPU in this very step: a plain (coherent at system scope after the
  // inter-GPU barrier) load, no non-coherent / read-only path, no L1 allocation of stale lines
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

// VEC = 4: width, column offset and every leading dimension are multiples of 4 floats and all bases 16-byte aligned
template <int VEC>
__global__ void __launch_bounds__(kHaloThreads)
k_halo_pull(int world, PeerPtrs peers, int64_t rows, int64_t ld_shard, int64_t c0, int width, float* __restrict__ table,
            int64_t ld_table) {
  const int wv = width / VEC;                      // vectors per row piece
  const int64_t per_rank = rows * wv;
  const int64_t total = per_rank * world;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if constexpr (VEC == 4) {
    for (; i + (kUnroll - 1) * stride < total; i += kUnroll * stride) {
      float4 v[kUnroll];
      int64_t dsto[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t j = i + u * stride;
        const int r = (int)(j / per_rank);
        const int64_t k = j - r * per_rank, row = k / wv;
        const int col = (int)(k - row * wv) * 4;
        v[u] = ld_peer(reinterpret_cast<const float4*>(peers.p[r] + row * ld_shard + c0 + col));
        dsto[u] = (r * rows + row) * ld_table + c0 + col;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) *reinterpret_cast<float4*>(table + dsto[u]) = v[u];
    }
  }
  for (; i < total; i += stride) {
    const int r = (int)(i / per_rank);
    const int64_t k = i - r * per_rank, row = k / wv;
    const int col = (int)(k - row * wv) * VEC;
    const float* s = peers.p[r] + row * ld_shard + c0 + col;
    float* d = table + (r * rows + row) * ld_table + c0 + col;
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(d) = ld_peer(reinterpret_cast<const float4*>(s));
    else {
      float x;
      asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(x) : "l"(s) : "memory");
      *d = x;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kHaloThreads)
k_halo_pull_reduce(int world, int me, PeerPtrs peers, int64_t rows, int64_t ld_table, int64_t c0, int width,
                   float* __restrict__ out, int64_t ld_out) {
  const int wv = width / VEC;
  const int64_t total = rows * wv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / wv;
    const int col = (int)(i - row * wv) * VEC;
    const int64_t off = (me * rows + row) * ld_table + c0 + col;
    if constexpr (VEC == 4) {
      float4 v[kMaxWorld];
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r)
        if (r < world) v[r] = ld_peer(reinterpret_cast<const float4*>(peers.p[r] + off));   // all peers in flight together
      float4 a = v[0];
#pragma unroll
      for (int r = 1; r < kMaxWorld; ++r)
        if (r < world) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }      // fixed order: deterministic
      *reinterpret_cast<float4*>(out + row * ld_out + c0 + col) = a;
    } else {
      float a = 0.f;
      for (int r = 0; r < world; ++r) {
        float x;
        asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(x) : "l"(peers.p[r] + off) : "memory");
        a += x;
      }
      out[row * ld_out + c0 + col] = a;
    }
  }
}

static bool aligned4(const void* p, int64_t a, int64_t b, int64_t c, int64_t d) {
  return ((uintptr_t)p % 16) == 0 && a % 4 == 0 && b % 4 == 0 && c % 4 == 0 && d % 4 == 0;
}

}  // namespace botgat

using namespace botgat;

extern "C" int botgat_halo_pull(int32_t world, const float* const* peer_shards, int64_t rows_per_rank, int64_t ld_shard,
                                int64_t col0, int64_t width, float* table, int64_t ld_table, int32_t n_blocks, void* stream) {
  BG_REQUIRE(world >= 1 && world <= kMaxWorld, "halo_pull: world must be in [1, %d]", kMaxWorld);
  BG_REQUIRE(peer_shards && table, "halo_pull: null pointer");
  BG_REQUIRE(rows_per_rank >= 0 && width >= 0 && col0 >= 0 && col0 + width <= ld_shard && col0 + width <= ld_table,
             "halo_pull: column range [%lld, +%lld) outside the rows", (long long)col0, (long long)width);
  if (rows_per_rank == 0 || width == 0) return 0;
  BG_REQUIRE(width < (1 << 30), "halo_pull: width too large");
  PeerPtrs pp;
  bool vec = aligned4(table, ld_shard, ld_table, col0, width);
  for (int r = 0; r < kMaxWorld; ++r) {
    pp.p[r] = r < world ? peer_shards[r] : nullptr;
    if (r < world) {
      BG_REQUIRE(pp.p[r], "halo_pull: null peer pointer for rank %d", r);
      vec = vec && ((uintptr_t)pp.p[r] % 16) == 0;
    }
  }
  const int grid = n_blocks > 0 ? n_blocks : 64;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) k_halo_pull<4><<<grid, kHaloThreads, 0, st>>>(world, pp, rows_per_rank, ld_shard, col0, (int)width, table, ld_table);
  else k_halo_pull<1><<<grid, kHaloThreads, 0, st>>>(world, pp, rows_per_rank, ld_shard, col0, (int)width, table, ld_table);
  BG_LAUNCHED(1);
  BG_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int botgat_halo_pull_reduce(int32_t world, int32_t rank, const float* const* peer_tables, int64_t rows_per_rank,
                                       int64_t ld_table, int64_t col0, int64_t width, float* out, int64_t ld_out,
                                       int32_t n_blocks, void* stream) {
  BG_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "halo_pull_reduce: bad world / rank");
  BG_REQUIRE(peer_tables && out, "halo_pull_reduce: null pointer");
  BG_REQUIRE(rows_per_rank >= 0 && width >= 0 && col0 >= 0 && col0 + width <= ld_table && col0 + width <= ld_out,
             "halo_pull_reduce: column range [%lld, +%lld) outside the rows", (long long)col0, (long long)width);
  if (rows_per_rank == 0 || width == 0) return 0;
  BG_REQUIRE(width < (1 << 30), "halo_pull_reduce: width too large");
  PeerPtrs pp;
  bool vec = aligned4(out, ld_table, ld_out, col0, width);
  for (int r = 0; r < kMaxWorld; ++r) {
    pp.p[r] = r < world ? peer_tables[r] : nullptr;
    if (r < world) {
      BG_REQUIRE(pp.p[r], "halo_pull_reduce: null peer pointer for rank %d", r);
      vec = vec && ((uintptr_t)pp.p[r] % 16) == 0;
    }
  }
  const int grid = n_blocks > 0 ? n_blocks : 64;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) k_halo_pull_reduce<4><<<grid, kHaloThreads, 0, st>>>(world, rank, pp, rows_per_rank, ld_table, col0, (int)width, out, ld_out);
  else k_halo_pull_reduce<1><<<grid, kHaloThreads, 0, st>>>(world, rank, pp, rows_per_rank, ld_table, col0, (int)width, out, ld_out);
  BG_LAUNCHED(1);
  BG_CHECK(cudaGetLastError());
  return 0;
}
