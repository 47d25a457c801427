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

static inline int stream_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 16));
}

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
  if (v4) k_edge_proj_fwd<true><<<stream_grid(n), 256, 0, st>>>(n, C, H, Hw, x, ld_x, W, y, ld_y);
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
  if (gx && n > 0) {
    BG_REQUIRE(gy && W && ld_gy >= H && ld_gx >= C, "edge_proj_backward: bad pointers / strides");
    const bool v4 = C % 4 == 0 && ld_gx % 4 == 0 && (uintptr_t)gx % 16 == 0;
    if (v4) k_edge_proj_gx<true><<<stream_grid(n), 256, 0, st>>>(n, C, H, gy, ld_gy, W, gx, ld_gx);
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
      k_edge_proj_gw<<<nb, 256, 0, st>>>(n, C, H, x, ld_x, gy, ld_gy, OP, partials);
    } else {
      BG_CHECK(cudaMemsetAsync(partials, 0, sizeof(float) * (size_t)nb * O, st));
    }
    k_edge_proj_gw_final<<<(O + 127) / 128, 128, 0, st>>>(nb, O, partials, gW);
    BG_LAUNCHED(2);
    BG_CHECK(cudaGetLastError());
  }
  return 0;
}
