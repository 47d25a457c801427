// Fused per-edge encoder + logit projection:   ee = relu(x W1^T + b1) W2^T
// x: raw edge features (E, C), W1: (M, C) + b1 (M) = the model's per-layer `edge_encoder[i]`
// (src/ogbn-proteins/models.py:201,245-247: Linear(8, 16) then ReLU), W2: (H, M) = the layer's `attn_edge_fc`
// (models.py:57-60,131).  The reference materialises the (E, M) embedding (2.5 GB at E = 39.6 M, saved for backward
// in every layer) and runs three skinny GEMMs around it; here the hidden units live in registers:
//   forward   reads 4C B + writes 32 B per edge
//   backward  reads 4C + 32 B per edge, recomputes the hidden units, accumulates grad W1 / b1 / W2 in registers,
//             fixed-order reduction (quads -> warp shuffles -> shared memory -> per-block partials -> one block):
//             deterministic, no atomics.
// Four lanes per edge row: lane q owns the hidden units 4q..4q+3 (M <= 16); its slices of W1, b1, W2 stay in
// registers for the whole kernel, so the row loop has no shared-memory traffic.  SURVEY.md section 8f rank 2.
#include "common.cuh"

namespace botgat {

constexpr int kMlpH = 8;       // padded head count = width of an edge record
constexpr int kMlpM = 16;      // hidden units covered (4 per lane)
constexpr int kMlpBlocks = 148 * 3;

template <int CP>
struct LaneWeights {
  float w1[4][CP];   // W1[4q + i][c]
  float b1[4];
  float w2[kMlpH][4];  // W2[h][4q + i]
  __device__ __forceinline__ void load(int q, int C, int M, int H, const float* __restrict__ W1,
                                       const float* __restrict__ B1, const float* __restrict__ W2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = 4 * q + i;
#pragma unroll
      for (int c = 0; c < CP; ++c) w1[i][c] = (m < M && c < C) ? __ldg(W1 + (int64_t)m * C + c) : 0.f;
      b1[i] = (m < M && B1 != nullptr) ? __ldg(B1 + m) : 0.f;
#pragma unroll
      for (int h = 0; h < kMlpH; ++h) w2[h][i] = (m < M && h < H) ? __ldg(W2 + (int64_t)h * M + m) : 0.f;
    }
  }
  // pre-activation of this lane's hidden units; the same expression in forward and backward, so the ReLU mask
  // of the recomputation is the forward's bit for bit
  __device__ __forceinline__ void pre(const float (&x)[CP], float (&p)[4]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = b1[i];
#pragma unroll
      for (int c = 0; c < CP; ++c) a = fmaf(x[c], w1[i][c], a);
      p[i] = a;
    }
  }
};

template <int CP, bool VEC>
__device__ __forceinline__ void load_row(float (&v)[CP], const float* __restrict__ p, int C, bool ok) {
  if constexpr (VEC) {
#pragma unroll
    for (int c = 0; c < CP; c += 4) {
      const float4 t = ok ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CP; ++c) v[c] = (ok && c < C) ? __ldg(p + c) : 0.f;
  }
}

constexpr int kMlpRows = 4;  // rows per thread and loop iteration (independent loads in flight)

template <int CP, bool VECX>
__global__ void __launch_bounds__(256, 2)
k_edge_mlp_fwd(int64_t n, int C, int M, int H, int Hw, const float* __restrict__ x, int64_t ld_x,
               const float* __restrict__ W1, const float* __restrict__ B1, const float* __restrict__ W2,
               float* __restrict__ y, int64_t ld_y, int vec2_store) {
  const int q = threadIdx.x & 3;
  LaneWeights<CP> w;
  w.load(q, C, M, H, W1, B1, W2);
  const int hb = (q & 1) * 4 + (q >> 1) * 2;  // the two heads this lane stores
  const int64_t rows_per_pass = (int64_t)gridDim.x * 64 * kMlpRows;
  const int64_t r_first = (int64_t)blockIdx.x * 64 * kMlpRows + (threadIdx.x >> 2);
  for (int64_t base = 0; base < n; base += rows_per_pass) {  // uniform trip count (whole-warp shuffles below)
    float xr[kMlpRows][CP];
#pragma unroll
    for (int u = 0; u < kMlpRows; ++u) {
      const int64_t r = base + r_first + 64 * u;
      load_row<CP, VECX>(xr[u], x + r * ld_x, C, r < n);
    }
#pragma unroll
    for (int u = 0; u < kMlpRows; ++u) {
      float p[4];
      w.pre(xr[u], p);
      float acc[kMlpH];
#pragma unroll
      for (int h = 0; h < kMlpH; ++h) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a = fmaf(fmaxf(p[i], 0.f), w.w2[h][i], a);
        acc[h] = a;
      }
      float k4[4], k2[2];  // packed butterfly over the quad: 8 partial sums -> 4 -> 2 per lane
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = (q & 1) ? acc[i] : acc[4 + i];
        const float keep = (q & 1) ? acc[4 + i] : acc[i];
        k4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = (q & 2) ? k4[i] : k4[2 + i];
        const float keep = (q & 2) ? k4[2 + i] : k4[i];
        k2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      const int64_t r = base + r_first + 64 * u;
      if (r < n) {
        float* yr = y + r * ld_y + hb;
        if (vec2_store) {
          if (hb < Hw) *reinterpret_cast<float2*>(yr) = make_float2(k2[0], k2[1]);
        } else {
          if (hb < Hw) yr[0] = k2[0];
          if (hb + 1 < Hw) yr[1] = k2[1];
        }
      }
    }
  }
}

// partial layout per block: [ gW1 (M*C) | gb1 (M) | gW2 (H*M) ]
template <int CP, bool VECX, bool VECG>
__global__ void __launch_bounds__(128, 3)
k_edge_mlp_bwd(int64_t n, int C, int M, int H, const float* __restrict__ x, int64_t ld_x,
               const float* __restrict__ W1, const float* __restrict__ B1, const float* __restrict__ W2,
               const float* __restrict__ gy, int64_t ld_gy, float* __restrict__ partials) {
  constexpr int kWarps = 4;
  constexpr int kPer = 4 * CP + 4 + kMlpH * 4;  // accumulators per lane
  __shared__ float sred[kWarps][4][kPer];
  const int q = threadIdx.x & 3, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LaneWeights<CP> w;
  w.load(q, C, M, H, W1, B1, W2);
  float a1[4][CP], ab[4], a2[kMlpH][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ab[i] = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) a1[i][c] = 0.f;
#pragma unroll
    for (int h = 0; h < kMlpH; ++h) a2[h][i] = 0.f;
  }
  constexpr int U = 2;
  const int64_t stride = (int64_t)gridDim.x * 32;  // 32 quads per block
  for (int64_t r0 = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 2); r0 < n; r0 += stride * U) {
    float xr[U][CP], g[U][kMlpH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * stride;
      load_row<CP, VECX>(xr[u], x + r * ld_x, C, r < n);
      load_row<kMlpH, VECG>(g[u], gy + r * ld_gy, H, r < n);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)  // after every load of the iteration has been issued
#pragma unroll
      for (int h = 0; h < kMlpH; ++h) g[u][h] = h < H ? g[u][h] : 0.f;  // whatever the record padding holds
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float p[4];
      w.pre(xr[u], p);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float hid = fmaxf(p[i], 0.f);
        float gh = 0.f;
#pragma unroll
        for (int h = 0; h < kMlpH; ++h) {  // padding heads: w2 = 0
          gh = fmaf(g[u][h], w.w2[h][i], gh);
          a2[h][i] = fmaf(g[u][h], hid, a2[h][i]);
        }
        gh = p[i] > 0.f ? gh : 0.f;
        ab[i] += gh;
#pragma unroll
        for (int c = 0; c < CP; ++c) a1[i][c] = fmaf(gh, xr[u][c], a1[i][c]);
      }
    }
  }
  // quads of a warp -> lanes 0..3, warps through shared memory in fixed order
  auto fold = [&](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
  };
  float* mine = sred[warp][q];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      const float v = fold(a1[i][c]);
      if (lane < 4) mine[i * CP + c] = v;
    }
    const float vb = fold(ab[i]);
    if (lane < 4) mine[4 * CP + i] = vb;
#pragma unroll
    for (int h = 0; h < kMlpH; ++h) {
      const float v = fold(a2[h][i]);
      if (lane < 4) mine[4 * CP + 4 + h * 4 + i] = v;
    }
  }
  __syncthreads();
  const int n1 = M * C, n2 = M, n3 = H * M;
  float* out = partials + (int64_t)blockIdx.x * (n1 + n2 + n3);
  for (int o = threadIdx.x; o < n1 + n2 + n3; o += blockDim.x) {
    int qq, idx;
    if (o < n1) {
      const int m = o / C, c = o - m * C;
      qq = m >> 2; idx = (m & 3) * CP + c;
    } else if (o < n1 + n2) {
      const int m = o - n1;
      qq = m >> 2; idx = 4 * CP + (m & 3);
    } else {
      const int h = (o - n1 - n2) / M, m = (o - n1 - n2) - h * M;
      qq = m >> 2; idx = 4 * CP + 4 + h * 4 + (m & 3);
    }
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kWarps; ++wv) s += sred[wv][qq][idx];
    out[o] = s;
  }
}

__global__ void k_edge_mlp_final(int n_blocks, int total, const float* __restrict__ partials, float* __restrict__ gW1, int n1,
                                 float* __restrict__ gb1, int n2, float* __restrict__ gW2) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partials[(int64_t)b * total + o];
  if (o < n1) { if (gW1) gW1[o] = s; }
  else if (o < n1 + n2) { if (gb1) gb1[o - n1] = s; }
  else if (gW2) gW2[o - n1 - n2] = s;
}

static inline bool mlp_supported(int C, int M, int H) { return C >= 1 && C <= 8 && M >= 1 && M <= kMlpM && H >= 1 && H <= kMlpH; }

}  // namespace botgat

using namespace botgat;

extern "C" int botgat_edge_mlp_supported(int32_t C, int32_t M, int32_t H) { return mlp_supported(C, M, H) ? 1 : 0; }

extern "C" int64_t botgat_edge_mlp_workspace_floats(int32_t C, int32_t M, int32_t H) {
  return (int64_t)kMlpBlocks * ((int64_t)M * C + M + (int64_t)H * M);
}

extern "C" int botgat_edge_mlp_forward(int64_t n, int32_t C, int32_t M, int32_t H, const float* x, int64_t ld_x,
                                       const float* W1, const float* b1, const float* W2, float* y, int64_t ld_y,
                                       int device, void* stream) {
  BG_REQUIRE(n >= 0 && mlp_supported(C, M, H), "edge_mlp: needs C <= 8 input features, M <= %d hidden units, H <= %d heads", kMlpM, kMlpH);
  if (n == 0) return 0;
  BG_REQUIRE(x && W1 && W2 && y && ld_x >= C && ld_y >= H, "edge_mlp_forward: bad pointers / strides");
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  const int Hw = (int)std::min<int64_t>(ld_y, kMlpH);
  const int v2 = ld_y % 2 == 0 && (uintptr_t)y % 8 == 0 && Hw % 2 == 0;
  const bool vx = C % 4 == 0 && ld_x % 4 == 0 && (uintptr_t)x % 16 == 0;
  const int64_t per_block = 64 * kMlpRows, work = (n + per_block - 1) / per_block;
#define BG_MLP_FWD(CP, VX) \
  k_edge_mlp_fwd<CP, VX><<<resident_grid(k_edge_mlp_fwd<CP, VX>, 256, work), 256, 0, st>>>(n, C, M, H, Hw, x, ld_x, W1, b1, W2, y, ld_y, v2)
  if (C <= 4) {
    if (vx) BG_MLP_FWD(4, true); else BG_MLP_FWD(4, false);
  } else {
    if (vx) BG_MLP_FWD(8, true); else BG_MLP_FWD(8, false);
  }
#undef BG_MLP_FWD
  BG_LAUNCHED(1);
  BG_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int botgat_edge_mlp_backward(int64_t n, int32_t C, int32_t M, int32_t H, const float* x, int64_t ld_x,
                                        const float* W1, const float* b1, const float* W2, const float* gy, int64_t ld_gy,
                                        float* gW1, float* gb1, float* gW2, float* partials, int device, void* stream) {
  BG_REQUIRE(n >= 0 && mlp_supported(C, M, H), "edge_mlp: needs C <= 8 input features, M <= %d hidden units, H <= %d heads", kMlpM, kMlpH);
  BG_REQUIRE(partials, "edge_mlp_backward: workspace of botgat_edge_mlp_workspace_floats() floats required");
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  const int n1 = M * C, n2 = M, n3 = H * M, total = n1 + n2 + n3;
  if (n > 0) {
    BG_REQUIRE(x && W1 && W2 && gy && ld_x >= C && ld_gy >= H, "edge_mlp_backward: bad pointers / strides");
    const bool vx = C % 4 == 0 && ld_x % 4 == 0 && (uintptr_t)x % 16 == 0;
    const bool vg = ld_gy % 4 == 0 && ld_gy >= kMlpH && (uintptr_t)gy % 16 == 0;  // padded 32-byte records
#define BG_MLP_BWD(CP, VX, VG) \
  k_edge_mlp_bwd<CP, VX, VG><<<kMlpBlocks, 128, 0, st>>>(n, C, M, H, x, ld_x, W1, b1, W2, gy, ld_gy, partials)
    if (C <= 4) {
      if (vx && vg) BG_MLP_BWD(4, true, true); else if (vx) BG_MLP_BWD(4, true, false);
      else if (vg) BG_MLP_BWD(4, false, true); else BG_MLP_BWD(4, false, false);
    } else {
      if (vx && vg) BG_MLP_BWD(8, true, true); else if (vx) BG_MLP_BWD(8, true, false);
      else if (vg) BG_MLP_BWD(8, false, true); else BG_MLP_BWD(8, false, false);
    }
#undef BG_MLP_BWD
  } else {
    BG_CHECK(cudaMemsetAsync(partials, 0, sizeof(float) * (size_t)kMlpBlocks * total, st));
  }
  k_edge_mlp_final<<<(total + 127) / 128, 128, 0, st>>>(kMlpBlocks, total, partials, gW1, n1, gb1, n2, gW2);
  BG_LAUNCHED(2);
  BG_CHECK(cudaGetLastError());
  return 0;
}
