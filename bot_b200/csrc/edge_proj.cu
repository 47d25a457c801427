// Per-edge logit projection  ee = feat_edge @ W^T  (W: (H, C), H <= 8 heads, C <= 64 edge features) and its
// backward, as streaming kernels.  This is `attn_edge_fc(feat_edge)` of src/ogbn-proteins/models.py:131 (K2 in
// SURVEY.md section 2b).  At E = 39.6 M, C = 16, H = 6 the three cuBLAS SGEMMs this replaces (skinny N = 8,
// "largek" K = E for the weight gradient) take 3.7 + 3.3 + 5.6 ms; they are pure streaming problems:
//   forward   reads 64 B + writes 32 B per edge
//   grad_x    reads 32 B + writes 64 B per edge
//   grad_W    reads 96 B per edge, tree-free fixed-order reduction (per-block partials, then one block) —
//             deterministic, no atomics.
// The output rows are written with the padded width the gather path wants (functional.pad_heads: one aligned
// 32-byte record per edge); padding columns are zero.
#include "common.cuh"

namespace botgat {

constexpr int kProjMaxH = 8;
constexpr int kProjMaxC = 64;

template <bool VEC4>
__global__ void __launch_bounds__(256)
k_edge_proj_fwd(int64_t n, int C, int H, int Hw, const float* __restrict__ x, int64_t ld_x, const float* __restrict__ W,
                float* __restrict__ y, int64_t ld_y) {
  __shared__ float sW[kProjMaxH * kProjMaxC];
  for (int i = threadIdx.x; i < kProjMaxH * C; i += blockDim.x) sW[i] = (i / C) < H ? W[i] : 0.f;
  __syncthreads();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const float* xr = x + r * ld_x;
    float acc[kProjMaxH];
#pragma unroll
    for (int h = 0; h < kProjMaxH; ++h) acc[h] = 0.f;
    if constexpr (VEC4) {
      for (int c = 0; c < C; c += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
#pragma unroll
        for (int h = 0; h < kProjMaxH; ++h) {
          const float* w = sW + h * C + c;
          acc[h] = fmaf(v.x, w[0], fmaf(v.y, w[1], fmaf(v.z, w[2], fmaf(v.w, w[3], acc[h]))));
        }
      }
    } else {
      for (int c = 0; c < C; ++c) {
        const float v = __ldg(xr + c);
#pragma unroll
        for (int h = 0; h < kProjMaxH; ++h) acc[h] = fmaf(v, sW[h * C + c], acc[h]);
      }
    }
    float* yr = y + r * ld_y;  // heads >= H have zero weights, so acc is already 0 in the padding columns
    if (VEC4 && (Hw & 3) == 0) {
#pragma unroll
      for (int q = 0; q < kProjMaxH / 4; ++q)
        if (q * 4 < Hw) *reinterpret_cast<float4*>(yr + q * 4) = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
    } else {
#pragma unroll
      for (int h = 0; h < kProjMaxH; ++h)
        if (h < Hw) yr[h] = acc[h];
    }
  }
}

template <bool VEC4>
__global__ void __launch_bounds__(256)
k_edge_proj_gx(int64_t n, int C, int H, const float* __restrict__ gy, int64_t ld_gy, const float* __restrict__ W,
               float* __restrict__ gx, int64_t ld_gx) {
  __shared__ float sW[kProjMaxH * kProjMaxC];
  for (int i = threadIdx.x; i < kProjMaxH * C; i += blockDim.x) sW[i] = (i / C) < H ? W[i] : 0.f;
  __syncthreads();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const float* g = gy + r * ld_gy;
    float gv[kProjMaxH];
#pragma unroll
    for (int h = 0; h < kProjMaxH; ++h) gv[h] = h < H ? __ldg(g + h) : 0.f;
    float* o = gx + r * ld_gx;
    if constexpr (VEC4) {
      for (int c = 0; c < C; c += 4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < kProjMaxH; ++h) {
          const float* w = sW + h * C + c;
          a.x = fmaf(gv[h], w[0], a.x); a.y = fmaf(gv[h], w[1], a.y);
          a.z = fmaf(gv[h], w[2], a.z); a.w = fmaf(gv[h], w[3], a.w);
        }
        *reinterpret_cast<float4*>(o + c) = a;
      }
    } else {
      for (int c = 0; c < C; ++c) {
        float a = 0.f;
#pragma unroll
        for (int h = 0; h < kProjMaxH; ++h) a = fmaf(gv[h], sW[h * C + c], a);
        o[c] = a;
      }
    }
  }
}

// grad_W[h][c] = sum_r gy[r][h] * x[r][c].  Block: 256 threads, tiles of kTile rows staged in shared memory; thread t
// owns output (t % OP) for the rows r = (t / OP) mod groups of every tile (OP = H*C rounded up to a power of two).
constexpr int kTile = 128;
__global__ void __launch_bounds__(256)
k_edge_proj_gw(int64_t n, int C, int H, const float* __restrict__ x, int64_t ld_x, const float* __restrict__ gy,
               int64_t ld_gy, int OP, float* __restrict__ partials) {
  __shared__ float sx[kTile * kProjMaxC];
  __shared__ float sg[kTile * kProjMaxH];
  __shared__ float sred[256];
  const int O = H * C;
  const int groups = 256 / OP;  // >= 1 (OP <= 256 enforced by the host)
  const int o = threadIdx.x % OP, gi = threadIdx.x / OP;
  const int oh = o / C, oc = o - oh * C;
  const bool live = o < O && gi < groups;
  float acc = 0.f;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t r0 = t * kTile;
    const int rows = (int)((n - r0) < (int64_t)kTile ? (n - r0) : (int64_t)kTile);
    for (int i = threadIdx.x; i < rows * C; i += 256) {
      const int r = i / C, c = i - r * C;
      sx[r * C + c] = __ldg(x + (r0 + r) * ld_x + c);
    }
    for (int i = threadIdx.x; i < rows * H; i += 256) {
      const int r = i / H, h = i - r * H;
      sg[r * kProjMaxH + h] = __ldg(gy + (r0 + r) * ld_gy + h);
    }
    __syncthreads();
    if (live)
      for (int r = gi; r < rows; r += groups) acc = fmaf(sg[r * kProjMaxH + oh], sx[r * C + oc], acc);
    __syncthreads();
  }
  // fixed-order combine of the row groups inside the block, then one partial per block
  sred[threadIdx.x] = live ? acc : 0.f;
  __syncthreads();
  if (threadIdx.x < O) {
    float s = 0.f;
    for (int q = 0; q < groups; ++q) s += sred[q * OP + threadIdx.x];
    partials[(int64_t)blockIdx.x * O + threadIdx.x] = s;
  }
}

__global__ void k_edge_proj_gw_final(int n_blocks, int O, const float* __restrict__ partials, float* __restrict__ gW) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partials[(int64_t)b * O + o];
  gW[o] = s;
}

// ---- aligned fast paths: four lanes per edge row ----------------------------------------------------------------
// Lane q of a row's quad owns the columns {4q + 16j}: a warp instruction reads 8 rows x 64 B = 512 contiguous bytes
// (C = 16) instead of 32 rows at a 64-byte stride, kRowsPerIter independent row loads are in flight per thread.
constexpr int kRowsPerIter = 4;

template <int JT>
__global__ void __launch_bounds__(256, JT == 1 ? 3 : (JT == 2 ? 2 : 1))
k_edge_proj_fwd_quad(int64_t n, int C, int H, int Hw, const float* __restrict__ x, int64_t ld_x,
                     const float* __restrict__ W, float* __restrict__ y, int64_t ld_y, int vec2_store) {
  const int q = threadIdx.x & 3;
  float4 wq[JT][kProjMaxH];  // this lane's slice of W stays in registers: no shared-memory traffic in the row loop
#pragma unroll
  for (int j = 0; j < JT; ++j)
#pragma unroll
    for (int h = 0; h < kProjMaxH; ++h) {
      const int c = 4 * q + 16 * j;
      wq[j][h] = (h < H && c < C) ? __ldg(reinterpret_cast<const float4*>(W + h * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  const int64_t rows_per_pass = (int64_t)gridDim.x * 64 * kRowsPerIter;
  const int64_t r_first = (int64_t)blockIdx.x * 64 * kRowsPerIter + (threadIdx.x >> 2);  // + 64 u: 8 adjacent rows per warp load
  const int hb = (q & 1) * 4 + (q >> 1) * 2;  // the two heads this lane ends up owning
  for (int64_t base = 0; base < n; base += rows_per_pass) {  // uniform trip count: the shuffles need whole warps
    float4 v[kRowsPerIter][JT];  // every row load of the iteration is issued before the first use
#pragma unroll
    for (int u = 0; u < kRowsPerIter; ++u) {
      const int64_t r = base + r_first + 64 * u;
#pragma unroll
      for (int j = 0; j < JT; ++j) {
        const int c = 4 * q + 16 * j;
        v[u][j] = (r < n && c < C) ? __ldg(reinterpret_cast<const float4*>(x + r * ld_x + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsPerIter; ++u) {
      float acc[kProjMaxH];
#pragma unroll
      for (int h = 0; h < kProjMaxH; ++h) acc[h] = 0.f;
#pragma unroll
      for (int j = 0; j < JT; ++j) {
#pragma unroll
        for (int h = 0; h < kProjMaxH; ++h) {
          const float4 w = wq[j][h];
          acc[h] = fmaf(v[u][j].x, w.x, fmaf(v[u][j].y, w.y, fmaf(v[u][j].z, w.z, fmaf(v[u][j].w, w.w, acc[h]))));
        }
      }
      // packed butterfly over the quad: 8 partial sums -> 4 -> 2 per lane
      float k4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = (q & 1) ? acc[i] : acc[4 + i];
        const float keep = (q & 1) ? acc[4 + i] : acc[i];
        k4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      float k2[2];
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

// one grad record (8 floats, heads >= H read as 0): two float4 when the rows are the padded 32-byte records
template <bool VECG>
__device__ __forceinline__ void load_grad_record(float (&g)[kProjMaxH], const float* __restrict__ gr, int H, bool ok) {
  if constexpr (VECG) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a = ok ? __ldg(reinterpret_cast<const float4*>(gr)) : z;
    const float4 b = ok ? __ldg(reinterpret_cast<const float4*>(gr + 4)) : z;
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w;
    g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
  } else {
#pragma unroll
    for (int h = 0; h < kProjMaxH; ++h) g[h] = (ok && h < H) ? __ldg(gr + h) : 0.f;
  }
}

template <bool VECG>
__global__ void __launch_bounds__(256, 3)
k_edge_proj_gx_quad(int64_t n, int C, int H, const float* __restrict__ gy, int64_t ld_gy,
                    const float* __restrict__ W, float* __restrict__ gx, int64_t ld_gx) {
  __shared__ __align__(16) float sW[kProjMaxH * kProjMaxC];
  for (int i = threadIdx.x; i < kProjMaxH * C; i += blockDim.x) sW[i] = (i / C) < H ? W[i] : 0.f;
  __syncthreads();
  const int q = threadIdx.x & 3;
  const int64_t stride = (int64_t)gridDim.x * 64;
  for (int64_t r0 = (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2); r0 < n; r0 += stride * kRowsPerIter) {
    float g[kRowsPerIter][kProjMaxH];
#pragma unroll
    for (int u = 0; u < kRowsPerIter; ++u) {
      const int64_t r = r0 + u * stride;
      load_grad_record<VECG>(g[u], gy + r * ld_gy, H, r < n);
    }
#pragma unroll
    for (int u = 0; u < kRowsPerIter; ++u)  // after every load of the iteration has been issued
#pragma unroll
      for (int h = 0; h < kProjMaxH; ++h) g[u][h] = h < H ? g[u][h] : 0.f;  // whatever the record padding holds
    for (int c = 4 * q; c < C; c += 16) {
      float4 a[kRowsPerIter];
#pragma unroll
      for (int u = 0; u < kRowsPerIter; ++u) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int h = 0; h < kProjMaxH; ++h) {  // padding heads have zero weights
        const float4 w = *reinterpret_cast<const float4*>(sW + h * C + c);
#pragma unroll
        for (int u = 0; u < kRowsPerIter; ++u) {
          a[u].x = fmaf(g[u][h], w.x, a[u].x); a[u].y = fmaf(g[u][h], w.y, a[u].y);
          a[u].z = fmaf(g[u][h], w.z, a[u].z); a[u].w = fmaf(g[u][h], w.w, a[u].w);
        }
      }
#pragma unroll
      for (int u = 0; u < kRowsPerIter; ++u) {
        const int64_t r = r0 + u * stride;
        if (r < n) *reinterpret_cast<float4*>(gx + r * ld_gx + c) = a[u];
      }
    }
  }
}

// grad_W with register accumulators: lane q of a quad owns grad_W[0..7][4q + 16j .. +3], j < JT, for the rows its quad
// streams; quads of a warp are merged by shuffles, warps of a block in fixed order through shared memory.
template <int JT, bool VECG>
__global__ void __launch_bounds__(256)
k_edge_proj_gw_quad(int64_t n, int C, int H, const float* __restrict__ x, int64_t ld_x, const float* __restrict__ gy,
                    int64_t ld_gy, float* __restrict__ partials) {
  __shared__ float sred[8][kProjMaxH * 16 * JT];
  const int q = threadIdx.x & 3, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[JT][kProjMaxH][4];
#pragma unroll
  for (int j = 0; j < JT; ++j)
#pragma unroll
    for (int h = 0; h < kProjMaxH; ++h)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][h][i] = 0.f;
  const int64_t stride = (int64_t)gridDim.x * 64;
  constexpr int U = JT == 1 ? 4 : 2;
  for (int64_t r0 = (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2); r0 < n; r0 += stride * U) {
    float g[U][kProjMaxH];
    float4 v[U][JT];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * stride;
      load_grad_record<VECG>(g[u], gy + r * ld_gy, H, r < n);
#pragma unroll
      for (int j = 0; j < JT; ++j) {
        const int c = 4 * q + 16 * j;
        v[u][j] = (r < n && c < C) ? __ldg(reinterpret_cast<const float4*>(x + r * ld_x + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < JT; ++j)
#pragma unroll
        for (int h = 0; h < kProjMaxH; ++h) {
          acc[j][h][0] = fmaf(g[u][h], v[u][j].x, acc[j][h][0]);
          acc[j][h][1] = fmaf(g[u][h], v[u][j].y, acc[j][h][1]);
          acc[j][h][2] = fmaf(g[u][h], v[u][j].z, acc[j][h][2]);
          acc[j][h][3] = fmaf(g[u][h], v[u][j].w, acc[j][h][3]);
        }
  }
#pragma unroll
  for (int j = 0; j < JT; ++j)
#pragma unroll
    for (int h = 0; h < kProjMaxH; ++h)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = acc[j][h][i];
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 16);
        if (lane < 4) sred[warp][h * (16 * JT) + 16 * j + 4 * q + i] = a;
      }
  __syncthreads();
  for (int o = threadIdx.x; o < H * C; o += 256) {
    const int h = o / C, c = o - h * C;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sred[w][h * (16 * JT) + c];
    partials[(int64_t)blockIdx.x * (H * C) + o] = s;
  }
}

static inline int stream_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 16));
}

// blocks of work of the quad kernels: 64 rows per block and pass, `per_thread` passes per loop iteration
static inline int64_t quad_blocks(int64_t n, int per_thread) { return (n + 64 * per_thread - 1) / (64 * per_thread); }

}  // namespace botgat

using namespace botgat;

extern "C" int botgat_edge_proj_gw_blocks(void) { return 148 * 4; }

extern "C" int botgat_edge_proj_forward(int64_t n, int32_t C, int32_t H, const float* x, int64_t ld_x, const float* W,
                                        float* y, int64_t ld_y, int device, void* stream) {
  BG_REQUIRE(n >= 0 && C > 0 && C <= kProjMaxC && H > 0 && H <= kProjMaxH, "edge_proj: needs C <= %d and H <= %d", kProjMaxC, kProjMaxH);
  if (n == 0) return 0;
  BG_REQUIRE(x && W && y && ld_x >= C && ld_y >= H, "edge_proj_forward: bad pointers / strides");
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  const int Hw = (int)std::min<int64_t>(ld_y, kProjMaxH);  // columns written (heads + zero padding)
  const bool v4 = C % 4 == 0 && ld_x % 4 == 0 && (uintptr_t)x % 16 == 0 && ld_y % 4 == 0 && (uintptr_t)y % 16 == 0;
  const bool vx = C % 4 == 0 && ld_x % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)W % 16 == 0;
  if (vx) {
    const int v2 = ld_y % 2 == 0 && (uintptr_t)y % 8 == 0 && Hw % 2 == 0;
    const int JT = (C + 15) / 16;
    const int64_t work = quad_blocks(n, kRowsPerIter);
#define BG_PROJ_FWD(J) \
  k_edge_proj_fwd_quad<J><<<resident_grid(k_edge_proj_fwd_quad<J>, 256, work), 256, 0, st>>>(n, C, H, Hw, x, ld_x, W, y, ld_y, v2)
    if (JT == 1) BG_PROJ_FWD(1);
    else if (JT == 2) BG_PROJ_FWD(2);
    else if (JT == 3) BG_PROJ_FWD(3);
    else BG_PROJ_FWD(4);
#undef BG_PROJ_FWD
  } else if (v4) k_edge_proj_fwd<true><<<stream_grid(n), 256, 0, st>>>(n, C, H, Hw, x, ld_x, W, y, ld_y);
  else k_edge_proj_fwd<false><<<stream_grid(n), 256, 0, st>>>(n, C, H, Hw, x, ld_x, W, y, ld_y);
  BG_LAUNCHED(1);
  BG_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int botgat_edge_proj_backward(int64_t n, int32_t C, int32_t H, const float* x, int64_t ld_x, const float* W,
                                         const float* gy, int64_t ld_gy, float* gx, int64_t ld_gx, float* gW,
                                         float* partials, int device, void* stream) {
  BG_REQUIRE(n >= 0 && C > 0 && C <= kProjMaxC && H > 0 && H <= kProjMaxH, "edge_proj: needs C <= %d and H <= %d", kProjMaxC, kProjMaxH);
  BG_REQUIRE(H * C <= 256, "edge_proj_backward: H*C must be <= 256");
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  // the 32-byte padded records of functional.pad_heads can be read as two float4
  const int vec_gy = gy && ld_gy % 4 == 0 && ld_gy >= kProjMaxH && (uintptr_t)gy % 16 == 0;
  if (gx && n > 0) {
    BG_REQUIRE(gy && W && ld_gy >= H && ld_gx >= C, "edge_proj_backward: bad pointers / strides");
    const bool v4 = C % 4 == 0 && ld_gx % 4 == 0 && (uintptr_t)gx % 16 == 0;
    if (v4 && vec_gy)
      k_edge_proj_gx_quad<true><<<resident_grid(k_edge_proj_gx_quad<true>, 256, quad_blocks(n, kRowsPerIter)), 256, 0, st>>>(
          n, C, H, gy, ld_gy, W, gx, ld_gx);
    else if (v4)
      k_edge_proj_gx_quad<false><<<resident_grid(k_edge_proj_gx_quad<false>, 256, quad_blocks(n, kRowsPerIter)), 256, 0, st>>>(
          n, C, H, gy, ld_gy, W, gx, ld_gx);
    else k_edge_proj_gx<false><<<stream_grid(n), 256, 0, st>>>(n, C, H, gy, ld_gy, W, gx, ld_gx);
    BG_LAUNCHED(1);
    BG_CHECK(cudaGetLastError());
  }
  if (gW) {
    BG_REQUIRE(partials, "edge_proj_backward: partials workspace (botgat_edge_proj_gw_blocks()*H*C floats) required");
    const int O = H * C;
    int OP = 1;
    while (OP < O) OP <<= 1;
    const int nb = botgat_edge_proj_gw_blocks();
    if (n > 0) {
      BG_REQUIRE(x && gy && ld_x >= C && ld_gy >= H, "edge_proj_backward: bad pointers / strides");
      const bool vx = C % 4 == 0 && ld_x % 4 == 0 && (uintptr_t)x % 16 == 0;
      const int JT = (C + 15) / 16;
      if (!vx) k_edge_proj_gw<<<nb, 256, 0, st>>>(n, C, H, x, ld_x, gy, ld_gy, OP, partials);
#define BG_PROJ_GW(J) \
  do { if (vec_gy) k_edge_proj_gw_quad<J, true><<<nb, 256, 0, st>>>(n, C, H, x, ld_x, gy, ld_gy, partials); \
       else k_edge_proj_gw_quad<J, false><<<nb, 256, 0, st>>>(n, C, H, x, ld_x, gy, ld_gy, partials); } while (0)
      else if (JT == 1) BG_PROJ_GW(1);
      else if (JT == 2) BG_PROJ_GW(2);
      else if (JT == 3) BG_PROJ_GW(3);
      else BG_PROJ_GW(4);
#undef BG_PROJ_GW
    } else {
      BG_CHECK(cudaMemsetAsync(partials, 0, sizeof(float) * (size_t)nb * O, st));
    }
    k_edge_proj_gw_final<<<(O + 127) / 128, 128, 0, st>>>(nb, O, partials, gW);
    BG_LAUNCHED(2);
    BG_CHECK(cudaGetLastError());
  }
  return 0;
}
