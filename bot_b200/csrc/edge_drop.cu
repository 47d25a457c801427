// Edge-drop keep mask with the reference's semantics, without the sort.
// Reference (src/no-sampling/models.py:528-532, src/ogbn-proteins/models.py:136-139):
//     perm = torch.randperm(E); bound = int(E * p); eids = perm[bound:]        # kept; perm[:bound] dropped
// i.e. a uniformly random subset of EXACTLY `bound` edges is dropped.  randperm is a device sort of E random keys
// (2.7 ms at E = 39.6 M); only the set perm[:bound] matters, and that is "the `bound` smallest keys": a selection.
// Every edge gets a 64-bit Philox4x32-10 key (recomputed on the fly, never stored); one histogram pass over the
// top 12 bits finds the digit that contains the bound-th smallest key, one marking pass writes the mask and
// collects the ~E/4096 undecided edges of that digit, which are sorted (cub, a few 10k items) to pick the rest.
// Same seed -> same mask; exactly `n_drop` zeros.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace botgat {

constexpr int kDropBits = 12;
constexpr int kDropBins = 1 << kDropBits;

struct DropState {
  uint32_t bin;       // first digit whose inclusive count reaches n_drop
  uint32_t need;      // how many edges of that digit are dropped
  uint32_t n_cand;    // edges of that digit seen by the marking pass
  uint32_t overflow;  // candidate buffer too small (cannot happen for keys that are uniform)
};

__device__ __forceinline__ void philox4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t (&o)[4]) {
  uint32_t c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
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

// keys of the edges 2*pair and 2*pair + 1: one Philox call, counter (pair, 0x80000000 | pair >> 32) — the high bit
// keeps the stream apart from the attention-dropout draws (philox_u32: second counter word head >> 2)
__device__ __forceinline__ void edge_pair_keys(uint64_t seed, int64_t pair, uint64_t& even, uint64_t& odd) {
  uint32_t o[4];
  philox4(seed, (uint32_t)pair, 0x80000000u | (uint32_t)((uint64_t)pair >> 32), o);
  even = ((uint64_t)o[0] << 32) | o[1];
  odd = ((uint64_t)o[2] << 32) | o[3];
}

__global__ void __launch_bounds__(256)
k_drop_hist(int64_t n_edges, uint64_t seed, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kDropBins];
  for (int i = threadIdx.x; i < kDropBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int64_t n_pairs = (n_edges + 1) >> 1;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
    uint64_t a, b;
    edge_pair_keys(seed, p, a, b);
    atomicAdd(&sh[a >> (64 - kDropBits)], 1u);
    if (2 * p + 1 < n_edges) atomicAdd(&sh[b >> (64 - kDropBits)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kDropBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one block of 1024 threads, four bins each
__global__ void __launch_bounds__(1024)
k_drop_pick(const uint32_t* __restrict__ hist, int64_t n_drop, DropState* __restrict__ st) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;
  unsigned long long mine = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) mine += hist[4 * t + i];
  part[t] = mine;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // inclusive scan
    const unsigned long long v = t >= off ? part[t - off] : 0ull;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  unsigned long long before = part[t] - mine;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned long long c = hist[4 * t + i];
    // the first bin whose inclusive count reaches n_drop (the host guarantees 1 <= n_drop < n_edges)
    if (before < (unsigned long long)n_drop && before + c >= (unsigned long long)n_drop) {
      st->bin = 4 * t + i;
      st->need = (uint32_t)((unsigned long long)n_drop - before);
    }
    before += c;
  }
}

__global__ void __launch_bounds__(256)
k_drop_mark(int64_t n_edges, uint64_t seed, DropState* __restrict__ st, uint8_t* __restrict__ keep,
            uint64_t* __restrict__ cand_key, uint32_t* __restrict__ cand_eid, uint32_t cap) {
  const uint32_t bin = st->bin;
  const int64_t n_pairs = (n_edges + 1) >> 1;
  const bool pair_store = ((uintptr_t)keep & 1) == 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
    uint64_t key[2];
    edge_pair_keys(seed, p, key[0], key[1]);
    uint8_t k[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t e = 2 * p + j;
      const uint32_t d = (uint32_t)(key[j] >> (64 - kDropBits));
      k[j] = d < bin ? 0 : 1;
      if (d == bin && e < n_edges) {
        const uint32_t i = atomicAdd(&st->n_cand, 1u);
        if (i < cap) { cand_key[i] = key[j]; cand_eid[i] = (uint32_t)e; }
        else st->overflow = 1;
      }
    }
    if (2 * p + 1 < n_edges && pair_store) *reinterpret_cast<uchar2*>(keep + 2 * p) = make_uchar2(k[0], k[1]);
    else {
      keep[2 * p] = k[0];
      if (2 * p + 1 < n_edges) keep[2 * p + 1] = k[1];
    }
  }
}

__global__ void k_drop_apply(const DropState* __restrict__ st, const uint32_t* __restrict__ sorted_eid, uint8_t* __restrict__ keep) {
  const uint32_t need = st->need;  // <= n_cand <= cap unless st->overflow (keys are uniform: never)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < need; i += gridDim.x * blockDim.x) keep[sorted_eid[i]] = 0;
}

// Small graphs (a partition's local edge set): the undecided candidates fit one block — bitonic sort in shared memory
// and apply in the same kernel instead of a device-wide radix sort of 64-bit keys (8 passes over a padded buffer).
constexpr int kSmallCap = 4096;
__global__ void __launch_bounds__(1024)
k_drop_sort_apply_small(const DropState* __restrict__ st, const uint64_t* __restrict__ cand_key,
                        const uint32_t* __restrict__ cand_eid, uint8_t* __restrict__ keep) {
  __shared__ uint64_t sk[kSmallCap];
  __shared__ uint32_t sv[kSmallCap];
  const uint32_t n = min(st->n_cand, (uint32_t)kSmallCap), need = min(st->need, n);
  int m = 2;  // sort the next power of two >= n only
  while (m < (int)n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    sk[i] = i < (int)n ? cand_key[i] : ~0ull;
    sv[i] = i < (int)n ? cand_eid[i] : 0u;
  }
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          // ties (equal 64-bit keys: practically never) are broken by edge id so that the result is deterministic
          const bool gt = sk[i] > sk[l] || (sk[i] == sk[l] && sv[i] > sv[l]);
          if (gt == up) {
            const uint64_t tk = sk[i]; sk[i] = sk[l]; sk[l] = tk;
            const uint32_t tv = sv[i]; sv[i] = sv[l]; sv[l] = tv;
          }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = threadIdx.x; i < need; i += blockDim.x) keep[sv[i]] = 0;
}

// the one-block path is taken when twice the expected population of a digit (mean E/4096, standard deviation
// sqrt(mean): 2x is > 40 sigma away) still fits it, i.e. up to 8.4 M edges
static inline bool drop_small(int64_t n_edges) { return (n_edges >> (kDropBits - 1)) <= kSmallCap; }

static inline uint32_t drop_cap(int64_t n_edges) {  // 8x the expected population of one digit, at least 64k
  if (drop_small(n_edges)) return kSmallCap;
  return (uint32_t)std::max<int64_t>(65536, n_edges >> (kDropBits - 3));
}
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct DropLayout {
  size_t hist, state, key_in, key_out, eid_in, eid_out, cub, total, cub_bytes;
  uint32_t cap;
};
static DropLayout drop_layout(int64_t n_edges) {
  DropLayout L;
  L.cap = drop_cap(n_edges);
  size_t o = 0;
  L.hist = o; o += align256(sizeof(uint32_t) * kDropBins);
  L.state = o; o += align256(sizeof(DropState));
  L.key_in = o; o += align256(sizeof(uint64_t) * L.cap);
  L.key_out = o; o += align256(sizeof(uint64_t) * L.cap);
  L.eid_in = o; o += align256(sizeof(uint32_t) * L.cap);
  L.eid_out = o; o += align256(sizeof(uint32_t) * L.cap);
  L.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, L.cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)L.cap);
  L.cub = o; o += align256(L.cub_bytes);
  L.total = o;
  return L;
}

}  // namespace botgat

using namespace botgat;

extern "C" int64_t botgat_edge_drop_workspace_bytes(int64_t n_edges) {
  return n_edges <= 0 ? 256 : (int64_t)drop_layout(n_edges).total;
}

extern "C" int botgat_edge_drop_draw(int64_t n_edges, int64_t n_drop, uint64_t seed, uint8_t* keep, void* workspace,
                                     int device, void* stream) {
  BG_REQUIRE(n_edges >= 0 && n_edges < ((int64_t)1 << 32), "edge_drop_draw: n_edges out of range");
  if (n_edges == 0) return 0;
  BG_REQUIRE(keep, "edge_drop_draw: keep is NULL");
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_drop <= 0) { BG_CHECK(cudaMemsetAsync(keep, 1, (size_t)n_edges, st)); return 0; }
  if (n_drop >= n_edges) { BG_CHECK(cudaMemsetAsync(keep, 0, (size_t)n_edges, st)); return 0; }
  BG_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0, "edge_drop_draw: workspace of botgat_edge_drop_workspace_bytes() bytes, 256-byte aligned, required");
  const DropLayout L = drop_layout(n_edges);
  char* ws = (char*)workspace;
  uint32_t* hist = (uint32_t*)(ws + L.hist);
  DropState* state = (DropState*)(ws + L.state);
  uint64_t* key_in = (uint64_t*)(ws + L.key_in);
  uint64_t* key_out = (uint64_t*)(ws + L.key_out);
  uint32_t* eid_in = (uint32_t*)(ws + L.eid_in);
  uint32_t* eid_out = (uint32_t*)(ws + L.eid_out);
  const bool small = drop_small(n_edges);
  BG_CHECK(cudaMemsetAsync(ws, 0, L.key_in, st));                                    // histogram + state
  if (!small) BG_CHECK(cudaMemsetAsync(key_in, 0xFF, sizeof(uint64_t) * L.cap, st));  // unused slots sort last
  const int64_t n_pairs = (n_edges + 1) >> 1;
  // at least 16 pairs per thread: every block merges its 4096-bin histogram into the global one
  const int64_t work = (n_pairs + 256 * 16 - 1) / (256 * 16);
  k_drop_hist<<<resident_grid(k_drop_hist, 256, work), 256, 0, st>>>(n_edges, seed, hist);
  k_drop_pick<<<1, 1024, 0, st>>>(hist, n_drop, state);
  k_drop_mark<<<resident_grid(k_drop_mark, 256, work), 256, 0, st>>>(n_edges, seed, state, keep, key_in, eid_in, L.cap);
  if (small) {
    k_drop_sort_apply_small<<<1, 1024, 0, st>>>(state, key_in, eid_in, keep);
  } else {
    size_t cub_bytes = L.cub_bytes;
    BG_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.cub, cub_bytes, key_in, key_out, eid_in, eid_out, (int)L.cap, 0, 64, st));
    k_drop_apply<<<64, 256, 0, st>>>(state, eid_out, keep);
  }
  BG_LAUNCHED(4);
  BG_CHECK(cudaGetLastError());
  return 0;
}
