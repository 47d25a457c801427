// Halo exchange of the 1-D destination-row partition through NVLink peer memory (sm_100a).
//
// The reference is single-GPU; the partitioned layer needs, per layer and direction, ONE
// exchange of the row-sharded source table [ft | el].  Instead of an NCCL all-gather /
// reduce-scatter of whole tables these two kernels move a COLUMN RANGE of the table straight between the ranks'
// buffers with peer loads (every rank maps the others' buffers: CUDA VMM / torch symmetric memory):
//
//   pull        table[r * rows + i, c0:c0+w] = table_r[r * rows + i, c0:c0+w]   for every rank r != me  (forward)
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
// Small blocks: the exchange runs BESIDE a gather kernel that owns (nearly) the whole register file; a 128-thread block
// fits into the space one retiring gather block frees, a 512-thread block would wait for the gather grid to drain.
constexpr int kHaloThreads = 128;

struct PeerPtrs {
  const float* p[kMaxWorld];
};

__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 r;
  // peer memory is written by another GPU in this very step: a system-scope load (no read-only / L1 path that could
  // serve a stale line); ordering against the writer comes from the inter-GPU barrier the caller places before the launch
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float x;
  asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(x) : "l"(p) : "memory");
  return x;
}

// table[r * rows + i, c0:c0+w] = peer_r[r * rows + i, c0:c0+w] for every rank r != me.  Every rank's table has the
// same layout and rank r's OWN slice of its table is the authoritative copy of its rows: nothing is staged.
// A thread reads the same (row, column) piece from EVERY peer, starting with the peer after its own rank: all NVLink
// ports of every GPU carry traffic all the time (walking the peers one after the other would have all ranks pull from
// rank 0 first — one egress port saturated, the rest idle: 3x slower at 8 GPUs).
// VEC = 4: width, column offset and the leading dimension are multiples of 4 floats and all bases 16-byte aligned.
template <int VEC, int WORLD>
__global__ void __launch_bounds__(kHaloThreads)
k_halo_pull(int world_rt, int me, PeerPtrs peers, int64_t rows, int64_t ld, int64_t c0, int width, float* __restrict__ table) {
  const int world = WORLD > 0 ? WORLD : world_rt;
  constexpr int NP = WORLD > 0 ? WORLD - 1 : 1;                      // peers read per piece
  constexpr int U = VEC == 4 && WORLD > 0 ? (NP >= 8 ? 1 : NP >= 4 ? 2 : NP >= 2 ? 4 : 8) : 1;   // ~8 loads in flight
  const int wv = width / VEC;
  const int64_t total = rows * wv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if constexpr (VEC == 4 && WORLD > 0) {
    for (; i + (U - 1) * stride < total; i += U * stride) {
      float4 v[U][NP];
      int64_t off[U][NP];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = i + u * stride, row = j / wv;
        const int64_t inner = row * ld + c0 + (j - row * wv) * 4;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          int r = me + 1 + q;
          r -= r >= WORLD ? WORLD : 0;
          off[u][q] = r * rows * ld + inner;
          v[u][q] = ld_peer(reinterpret_cast<const float4*>(peers.p[r] + off[u][q]));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < NP; ++q) *reinterpret_cast<float4*>(table + off[u][q]) = v[u][q];
    }
  }
  for (; i < total; i += stride) {
    const int64_t row = i / wv;
    const int64_t inner = row * ld + c0 + (i - row * wv) * VEC;
    for (int q = 0; q < world - 1; ++q) {
      const int r = (me + 1 + q) % world;
      const int64_t off = r * rows * ld + inner;
      if constexpr (VEC == 4) *reinterpret_cast<float4*>(table + off) = ld_peer(reinterpret_cast<const float4*>(peers.p[r] + off));
      else table[off] = ld_peer1(peers.p[r] + off);
    }
  }
}

// out[i, c0:c0+w] = sum_r table_r[me * rows + i, c0:c0+w], summed in the fixed order r = 0..world-1
template <int VEC, int WORLD>
__global__ void __launch_bounds__(kHaloThreads)
k_halo_pull_reduce(int world_rt, int me, PeerPtrs peers, int64_t rows, int64_t ld_table, int64_t c0, int width,
                   float* __restrict__ out, int64_t ld_out) {
  const int world = WORLD > 0 ? WORLD : world_rt;
  constexpr int U = VEC == 4 ? (WORLD == 2 ? 4 : (WORLD > 0 && WORLD <= 4) ? 2 : 1) : 1;   // ~8 peer loads in flight per thread
  const int wv = width / VEC;
  const int64_t total = rows * wv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if constexpr (VEC == 4 && WORLD > 0) {
    for (; i + (U - 1) * stride < total; i += U * stride) {
      float4 v[U][WORLD];
      int64_t oo[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = i + u * stride, row = j / wv;
        const int col = (int)(j - row * wv) * 4;
        const int64_t off = (me * rows + row) * ld_table + c0 + col;
        oo[u] = row * ld_out + c0 + col;
#pragma unroll
        for (int r = 0; r < WORLD; ++r) v[u][r] = ld_peer(reinterpret_cast<const float4*>(peers.p[r] + off));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 a = v[u][0];
#pragma unroll
        for (int r = 1; r < WORLD; ++r) { a.x += v[u][r].x; a.y += v[u][r].y; a.z += v[u][r].z; a.w += v[u][r].w; }
        *reinterpret_cast<float4*>(out + oo[u]) = a;
      }
    }
  }
  for (; i < total; i += stride) {
    const int64_t row = i / wv;
    const int col = (int)(i - row * wv) * VEC;
    const int64_t off = (me * rows + row) * ld_table + c0 + col;
    if constexpr (VEC == 4) {
      float4 a = ld_peer(reinterpret_cast<const float4*>(peers.p[0] + off));
      for (int r = 1; r < world; ++r) {
        const float4 x = ld_peer(reinterpret_cast<const float4*>(peers.p[r] + off));
        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      }
      *reinterpret_cast<float4*>(out + row * ld_out + c0 + col) = a;
    } else {
      float a = ld_peer1(peers.p[0] + off);
      for (int r = 1; r < world; ++r) a += ld_peer1(peers.p[r] + off);
      out[row * ld_out + c0 + col] = a;
    }
  }
}

static bool aligned4(const void* p, int64_t a, int64_t b, int64_t c, int64_t d) {
  return ((uintptr_t)p % 16) == 0 && a % 4 == 0 && b % 4 == 0 && c % 4 == 0 && d % 4 == 0;
}

static int fill_peers(PeerPtrs& pp, int world, const float* const* ptrs, bool& vec) {
  for (int r = 0; r < kMaxWorld; ++r) {
    pp.p[r] = r < world ? ptrs[r] : nullptr;
    if (r < world) {
      if (!pp.p[r]) { set_error("halo: null peer pointer for rank %d", r); return -1; }
      vec = vec && ((uintptr_t)pp.p[r] % 16) == 0;
    }
  }
  return 0;
}

}  // namespace botgat

using namespace botgat;

extern "C" int botgat_halo_pull(int32_t world, int32_t rank, const float* const* peer_tables, int64_t rows_per_rank, int64_t ld,
                                int64_t col0, int64_t width, int32_t n_blocks, void* stream) {
  BG_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "halo_pull: world must be in [1, %d], rank in [0, world)", kMaxWorld);
  BG_REQUIRE(peer_tables, "halo_pull: null pointer");
  BG_REQUIRE(rows_per_rank >= 0 && width >= 0 && col0 >= 0 && col0 + width <= ld,
             "halo_pull: column range [%lld, +%lld) outside the rows", (long long)col0, (long long)width);
  if (rows_per_rank == 0 || width == 0 || world == 1) return 0;
  BG_REQUIRE(width < (1 << 30), "halo_pull: width too large");
  PeerPtrs pp;
  bool vec = aligned4(peer_tables[rank], ld, ld, col0, width);
  if (fill_peers(pp, world, peer_tables, vec)) return -1;
  float* table = const_cast<float*>(peer_tables[rank]);
  const int grid = n_blocks > 0 ? n_blocks : 592;
  cudaStream_t st = (cudaStream_t)stream;
#define BG_PL(VEC, W) k_halo_pull<VEC, W><<<grid, kHaloThreads, 0, st>>>(world, rank, pp, rows_per_rank, ld, col0, (int)width, table)
  if (!vec) BG_PL(1, 0);
  else if (world == 2) BG_PL(4, 2);
  else if (world == 4) BG_PL(4, 4);
  else if (world == 8) BG_PL(4, 8);
  else BG_PL(4, 0);
#undef BG_PL
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
  if (fill_peers(pp, world, peer_tables, vec)) return -1;
  const int grid = n_blocks > 0 ? n_blocks : 592;
  cudaStream_t st = (cudaStream_t)stream;
#define BG_PR(VEC, W) k_halo_pull_reduce<VEC, W><<<grid, kHaloThreads, 0, st>>>(world, rank, pp, rows_per_rank, ld_table, col0, (int)width, out, ld_out)
  if (!vec) BG_PR(1, 0);
  else if (world == 2) BG_PR(4, 2);
  else if (world == 4) BG_PR(4, 4);
  else if (world == 8) BG_PR(4, 8);
  else BG_PR(4, 0);
#undef BG_PR
  BG_LAUNCHED(1);
  BG_CHECK(cudaGetLastError());
  return 0;
}
