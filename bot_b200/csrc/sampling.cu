// Neighbour sampling and block (message-flow-graph) construction on the device.
// Replaces, for the sampled training / inference loops, dgl.dataloading.MultiLayerNeighborSampler + NodeDataLoader
// and their CPU worker processes (src/ogbn-proteins/gat.py:177-201, src/ogbn-products/gat.py:202-233):
//   frontier = sample_neighbors(g, seeds, fanout)   -> botgat_sample_count + botgat_sample_neighbors
//   block    = to_block(frontier, seeds)            -> botgat_block_compact (+ botgat_graph_create)
// Sampling is uniform without replacement over the in-edges of each seed (all of them when in-degree <= fanout).
// One warp per seed; the draw is a stream of Philox candidates accepted when new (= sequential sampling without
// replacement), processed 32 at a time: duplicates inside a batch by match.any, against earlier picks through a
// per-warp list in shared memory.  When more than half of the row is wanted the complement is drawn instead.
// Deterministic in (graph, seeds, fanout, seed); integer work, bit-exact against tests/util.py's restatement.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace botgat {

constexpr int kMaxFanout = 256;
constexpr int kSampleWarps = 8;

__device__ __forceinline__ void philox4c(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&o)[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// n-th candidate position in [0, d) for the row of node v: 64 random bits, multiply-shift
__device__ __forceinline__ uint32_t candidate(uint64_t seed, uint32_t v, uint32_t n, uint32_t d) {
  uint32_t o[4];
  philox4c(seed, n >> 1, 0x40000000u, v, 0x5BD1E995u, o);
  const uint64_t x = (n & 1) ? (((uint64_t)o[2] << 32) | o[3]) : (((uint64_t)o[0] << 32) | o[1]);
  return (uint32_t)__umul64hi(x, (uint64_t)d);
}

__global__ void k_sample_count(const int32_t* __restrict__ indptr, int64_t n_rows, int64_t n_seeds,
                               const int64_t* __restrict__ seeds, int fanout, int64_t* __restrict__ counts,
                               int* __restrict__ bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  const int64_t v = seeds[i];
  if (v < 0 || v >= n_rows) {  // reported by botgat_sample_count after its synchronisation; never used as an index
    *bad = 1;
    counts[i] = 0;
    return;
  }
  const int d = indptr[v + 1] - indptr[v];
  counts[i] = (fanout <= 0 || d <= fanout) ? d : fanout;
}

__global__ void __launch_bounds__(kSampleWarps * 32)
k_sample(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const int32_t* __restrict__ eids,
         int64_t n_rows, int64_t n_seeds, const int64_t* __restrict__ seeds, int fanout, uint64_t seed, const int64_t* __restrict__ offsets,
         int64_t* __restrict__ out_src, int64_t* __restrict__ out_dst, int64_t* __restrict__ out_eid) {
  __shared__ uint32_t s_picks[kSampleWarps][kMaxFanout];
  __shared__ uint32_t s_bits[kSampleWarps][2 * kMaxFanout / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* picks = s_picks[warp];
  uint32_t* bits = s_bits[warp];
  for (int64_t i = (int64_t)blockIdx.x * kSampleWarps + warp; i < n_seeds; i += (int64_t)gridDim.x * kSampleWarps) {
    const int64_t v = seeds[i];
    if (v < 0 || v >= n_rows) continue;  // botgat_sample_count already rejected such a seed list (count 0 here)
    const int beg = indptr[v], d = indptr[v + 1] - beg;
    const int64_t base = offsets[i];
    if (fanout <= 0 || d <= fanout) {
      for (int j = lane; j < d; j += 32) {
        out_src[base + j] = indices[beg + j];
        out_dst[base + j] = i;
        out_eid[base + j] = eids[beg + j];
      }
      continue;
    }
    const bool complement = fanout > d / 2;  // then d < 2 * fanout <= 2 * kMaxFanout
    const int m = complement ? d - fanout : fanout;
    int t = 0;       // accepted so far
    uint32_t n0 = 0; // candidates consumed
    while (t < m) {
      const uint32_t c = candidate(seed, (uint32_t)v, n0 + lane, (uint32_t)d);
      n0 += 32;
      // first occurrence inside the batch ...
      const unsigned same = __match_any_sync(kFull, c);
      bool fresh = (__ffs(same) - 1) == lane;
      // ... and not picked before
      for (int j = 0; j < t; ++j) fresh = fresh && picks[j] != c;
      const unsigned acc = __ballot_sync(kFull, fresh);
      const int pos = t + __popc(acc & ((1u << lane) - 1u));
      if (fresh && pos < m) picks[pos] = c;
      t = min(m, t + __popc(acc));
      __syncwarp();
    }
    if (!complement) {
      for (int j = lane; j < m; j += 32) {
        const int p = beg + (int)picks[j];
        out_src[base + j] = indices[p];
        out_dst[base + j] = i;
        out_eid[base + j] = eids[p];
      }
    } else {
      // the picks are the EXCLUDED positions: emit the rest in row order
      const int words = (d + 31) >> 5;
      for (int w = lane; w < words; w += 32) bits[w] = 0;
      __syncwarp();
      for (int j = lane; j < m; j += 32) atomicOr(&bits[picks[j] >> 5], 1u << (picks[j] & 31));
      __syncwarp();
      int written = 0;
      for (int j0 = 0; j0 < d; j0 += 32) {
        const int j = j0 + lane;
        const bool take = j < d && !((bits[j >> 5] >> (j & 31)) & 1u);
        const unsigned mk = __ballot_sync(kFull, take);
        if (take) {
          const int64_t o = base + written + __popc(mk & ((1u << lane) - 1u));
          out_src[o] = indices[beg + j];
          out_dst[o] = i;
          out_eid[o] = eids[beg + j];
        }
        written += __popc(mk);
      }
    }
    __syncwarp();
  }
}

// ---- block construction ------------------------------------------------------------------------------------------
// node ids outside [0, n_parent) set *bad (reported by botgat_block_compact after its synchronisation) and are never
// used as an index
__global__ void k_block_mark_seeds(int64_t n_parent, int64_t n_seeds, const int64_t* __restrict__ seeds,
                                   int32_t* __restrict__ seed_pos, int* __restrict__ bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  const int64_t v = seeds[i];
  if (v < 0 || v >= n_parent) { *bad = 1; return; }
  seed_pos[v] = (int32_t)i;
}
__global__ void k_block_mark_hits(int64_t n_parent, int64_t n_edges, const int64_t* __restrict__ src,
                                  const int32_t* __restrict__ seed_pos, int32_t* __restrict__ hit, int* __restrict__ bad) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n_edges) {
    const int64_t g = src[e];
    if (g < 0 || g >= n_parent) { *bad = 1; return; }
    if (seed_pos[g] < 0) hit[g] = 1;
  }
}
__global__ void k_block_nodes(int64_t n_parent, int64_t n_seeds, const int64_t* __restrict__ seeds,
                              const int32_t* __restrict__ hit, const int32_t* __restrict__ rank, int64_t* __restrict__ src_nodes) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g < n_seeds) src_nodes[g] = seeds[g];
  if (g < n_parent && hit[g]) src_nodes[n_seeds + rank[g]] = g;
}
__global__ void k_block_relabel(int64_t n_parent, int64_t n_edges, int64_t n_seeds, const int64_t* __restrict__ src,
                                const int32_t* __restrict__ seed_pos, const int32_t* __restrict__ rank,
                                int64_t* __restrict__ src_local) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n_edges) {
    const int64_t g = src[e];
    if (g < 0 || g >= n_parent) { src_local[e] = 0; return; }
    const int32_t sp = seed_pos[g];
    src_local[e] = sp >= 0 ? (int64_t)sp : n_seeds + rank[g];
  }
}

static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
struct BlockLayout { size_t seed_pos, hit, rank, cub, bad, total, cub_bytes; };
static BlockLayout block_layout(int64_t n_parent) {
  BlockLayout L;
  size_t o = 0;
  L.seed_pos = o; o += up256(sizeof(int32_t) * (size_t)n_parent);
  L.hit = o; o += up256(sizeof(int32_t) * (size_t)(n_parent + 1));
  L.rank = o; o += up256(sizeof(int32_t) * (size_t)(n_parent + 1));
  L.cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, L.cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n_parent + 1));
  L.cub = o; o += up256(L.cub_bytes);
  L.bad = o; o += 256;
  L.total = o;
  return L;
}

}  // namespace botgat

using namespace botgat;

extern "C" int64_t botgat_sample_workspace_bytes(int64_t n_seeds) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(n_seeds + 1));
  return (int64_t)(up256(bytes) + up256(sizeof(int64_t) * (size_t)(n_seeds + 1)) + 256);  // scan temp | counts | bad flag
}

extern "C" int botgat_sample_count(const botgat_graph* g, int64_t n_seeds, const int64_t* seeds, int32_t fanout,
                                   int64_t* offsets, int64_t* n_out, void* workspace, void* stream) {
  BG_REQUIRE(g && n_seeds >= 0 && n_out, "sample_count: bad arguments");
  BG_REQUIRE(fanout <= kMaxFanout, "sample_count: fanout must be <= %d (or <= 0 for every neighbour)", kMaxFanout);
  *n_out = 0;
  if (n_seeds == 0) return 0;
  BG_REQUIRE(seeds && offsets && workspace, "sample_count: NULL pointer");
  DeviceGuard guard(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(n_seeds + 1));
  int64_t* counts = (int64_t*)((char*)workspace + up256(cub_bytes));
  int* bad = (int*)((char*)workspace + up256(cub_bytes) + up256(sizeof(int64_t) * (size_t)(n_seeds + 1)));
  BG_CHECK(cudaMemsetAsync(counts + n_seeds, 0, sizeof(int64_t), st));
  BG_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
  k_sample_count<<<(unsigned)((n_seeds + 255) / 256), 256, 0, st>>>(g->in_indptr, g->n_dst, n_seeds, seeds, fanout, counts, bad);
  BG_CHECK(cub::DeviceScan::ExclusiveSum(workspace, cub_bytes, counts, offsets, (int)(n_seeds + 1), st));
  BG_LAUNCHED(2);
  int hbad = 0;
  BG_CHECK(cudaMemcpyAsync(n_out, offsets + n_seeds, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  BG_CHECK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  BG_CHECK(cudaStreamSynchronize(st));
  if (hbad) *n_out = 0;
  BG_REQUIRE(!hbad, "sample_count: seed node id out of range [0, %lld)", (long long)g->n_dst);
  return 0;
}

extern "C" int botgat_sample_neighbors(const botgat_graph* g, int64_t n_seeds, const int64_t* seeds, int32_t fanout,
                                       uint64_t seed, const int64_t* offsets, int64_t* out_src, int64_t* out_dst,
                                       int64_t* out_eid, void* stream) {
  BG_REQUIRE(g && n_seeds >= 0, "sample_neighbors: bad arguments");
  BG_REQUIRE(fanout <= kMaxFanout, "sample_neighbors: fanout must be <= %d (or <= 0 for every neighbour)", kMaxFanout);
  if (n_seeds == 0) return 0;
  BG_REQUIRE(seeds && offsets && out_src && out_dst && out_eid, "sample_neighbors: NULL pointer");
  DeviceGuard guard(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t work = (n_seeds + kSampleWarps - 1) / kSampleWarps;
  k_sample<<<resident_grid(k_sample, kSampleWarps * 32, work), kSampleWarps * 32, 0, st>>>(
      g->in_indptr, g->in_indices, g->in_eid, g->n_dst, n_seeds, seeds, fanout, seed, offsets, out_src, out_dst, out_eid);
  BG_LAUNCHED(1);
  BG_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int64_t botgat_block_workspace_bytes(int64_t n_parent) { return (int64_t)block_layout(n_parent).total; }

extern "C" int botgat_block_compact(int64_t n_parent, int64_t n_seeds, const int64_t* seeds, int64_t n_edges,
                                    const int64_t* src_global, int64_t* src_local, int64_t* src_nodes, int64_t* n_src,
                                    void* workspace, int device, void* stream) {
  BG_REQUIRE(n_parent >= 0 && n_parent < ((int64_t)1 << 31) - 1 && n_seeds >= 0 && n_edges >= 0 && n_src, "block_compact: bad sizes");
  *n_src = n_seeds;
  if (n_parent == 0) return 0;
  BG_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0, "block_compact: 256-byte aligned workspace required");
  BG_REQUIRE((n_seeds == 0 || (seeds && src_nodes)) && (n_edges == 0 || (src_global && src_local && src_nodes)), "block_compact: NULL pointer");
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  const BlockLayout L = block_layout(n_parent);
  char* ws = (char*)workspace;
  int32_t* seed_pos = (int32_t*)(ws + L.seed_pos);
  int32_t* hit = (int32_t*)(ws + L.hit);
  int32_t* rank = (int32_t*)(ws + L.rank);
  int* bad = (int*)(ws + L.bad);
  BG_CHECK(cudaMemsetAsync(seed_pos, 0xFF, sizeof(int32_t) * (size_t)n_parent, st));
  BG_CHECK(cudaMemsetAsync(hit, 0, sizeof(int32_t) * (size_t)(n_parent + 1), st));
  BG_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
  if (n_seeds) k_block_mark_seeds<<<(unsigned)((n_seeds + 255) / 256), 256, 0, st>>>(n_parent, n_seeds, seeds, seed_pos, bad);
  if (n_edges) k_block_mark_hits<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(n_parent, n_edges, src_global, seed_pos, hit, bad);
  size_t cub_bytes = L.cub_bytes;
  BG_CHECK(cub::DeviceScan::ExclusiveSum(ws + L.cub, cub_bytes, hit, rank, (int)(n_parent + 1), st));
  const int64_t span = n_parent > n_seeds ? n_parent : n_seeds;
  k_block_nodes<<<(unsigned)((span + 255) / 256), 256, 0, st>>>(n_parent, n_seeds, seeds, hit, rank, src_nodes);
  if (n_edges) k_block_relabel<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(n_parent, n_edges, n_seeds, src_global, seed_pos, rank, src_local);
  BG_LAUNCHED(5);
  int32_t n_new = 0;
  int hbad = 0;
  BG_CHECK(cudaMemcpyAsync(&n_new, rank + n_parent, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  BG_CHECK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  BG_CHECK(cudaStreamSynchronize(st));
  BG_REQUIRE(!hbad, "block_compact: node id out of range [0, %lld)", (long long)n_parent);
  *n_src = n_seeds + n_new;
  return 0;
}
