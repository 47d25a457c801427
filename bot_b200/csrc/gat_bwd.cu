// GAT backward for sm_100a (adjoint of gat_fwd.cu).
// Replaces the autograd replay of DGL's GSpMM / GSDDMM / EdgeSoftmax backward
// kernels behind src/no-sampling/models.py:523-555 and
// src/ogbn-proteins/models.py:125-156.
//
//   node pass  : t[v,h] = <out[v,h,:], gout[v,h,:]>; packs drec[h][v] =
//                {er, row_max, 1/row_sum, t}; g' = gout * dst_scale
//   src pass   : out-CSR (src-major)  -> grad_ft, grad_el            (always)
//   dst pass   : in-CSR  (dst-major)  -> grad_er, gz (= grad of the edge logit
//                term, in-CSR order)                                 (on request)
//
// Attention weights are recomputed from (el, er, eb, row_max, row_sum); nothing
// E-sized is saved by the forward.  No atomics: every output element is produced
// by exactly one warp in a fixed order, so results are run-to-run deterministic.
#include "common.cuh"

namespace botgat {

struct BwdParams {
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eid;
  int n_rows;     // rows of the CSR being walked
  int n_dst;
  int64_t n_edges;
  int H, D;
  int64_t ld_ft, ld_g, ld_gft;
  const float *ft, *el, *eb, *am, *cs;
  const float* g;       // g' (n_dst, ld_g)
  const float4* drec;   // (H, n_dst)
  int Hb;
  float slope, attn_p, inv_keep;
  uint64_t seed;
  float *grad_ft, *grad_el, *grad_er, *gz;
  int gshift;
  int blocks_per_slab;
};

// ---------------------------------------------------------------------------
// node pass: one warp per destination row
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gat_bwd_node_kernel(int n_dst, int H, int D, int64_t ld, const float* __restrict__ out,
                    const float* __restrict__ gout, const float* __restrict__ er,
                    const float* __restrict__ row_max, const float* __restrict__ row_sum,
                    const float* __restrict__ ds, float4* __restrict__ drec, float* __restrict__ gprime) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x * kWarpsPerBlock + warp;
  if (v >= n_dst) return;
  const float* o = out + (int64_t)v * ld;
  const float* g = gout + (int64_t)v * ld;
  const float dsv = ds ? ds[v] : 1.f;
  for (int h = 0; h < H; ++h) {
    float t = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float gv = g[h * D + d];
      t = fmaf(o[h * D + d], gv, t);
      if (gprime) gprime[(int64_t)v * ld + h * D + d] = gv * dsv;
    }
    t = warp_sum(t);
    if (lane == 0) {
      const float l = row_sum[(int64_t)v * H + h];
      drec[(int64_t)h * n_dst + v] =
          make_float4(er ? er[(int64_t)v * H + h] : 0.f, row_max[(int64_t)v * H + h], l > 0.f ? 1.f / l : 0.f, t);
    }
  }
}

// deliver a per-group value (held by every lane of group q) to lane e+q
__device__ __forceinline__ void deliver(float part, int e, int EPS, int gshift, int lane, float& d_lane) {
  const float got = __shfl_sync(kFull, part, ((lane - e) << gshift) & 31);
  if (lane >= e && lane < e + EPS) d_lane = got;
}

// ---------------------------------------------------------------------------
// src pass: one warp per (head, source row u) over the out-CSR
// ---------------------------------------------------------------------------
template <int VW, int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) gat_bwd_src_kernel(const BwdParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x / p.blocks_per_slab;
  const int row = (blockIdx.x - h * p.blocks_per_slab) * kWarpsPerBlock + warp;
  if (row >= p.n_rows) return;
  const int G = 1 << p.gshift;
  const int j = lane & (G - 1);
  const int grp = lane >> p.gshift;
  const int EPS = 32 >> p.gshift;
  const int gstride = G * VW;

  bool act[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) act[i] = (i * G + j) * VW < p.D;

  const int beg = p.indptr[row], end = p.indptr[row + 1];
  const float csu = p.cs ? p.cs[row] : 1.f;
  const float el_u = p.el[(int64_t)row * p.H + h];
  Vec<VW> fu[VPL], acc[VPL];
  {
    const float* f = p.ft + (int64_t)row * p.ld_ft + h * p.D + j * VW;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (act[i]) { fu[i].load(f + i * gstride); fu[i].scale(csu); } else fu[i].zero();
      acc[i].zero();
    }
  }
  const float* __restrict__ g_h = p.g + h * p.D + j * VW;
  const float4* __restrict__ drec_h = p.drec + (int64_t)h * p.n_dst;
  const float* __restrict__ eb_h = p.eb ? p.eb + (int64_t)(p.Hb == 1 ? 0 : h) * p.n_edges : nullptr;
  const float* __restrict__ am_h = p.am ? p.am + (int64_t)h * p.n_edges : nullptr;
  const bool philox = (p.am == nullptr) && p.attn_p > 0.f;
  float gel_lane = 0.f;

  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int v = 0;
    float alpha = 0.f, amul = 1.f, dz = 0.f, t = 0.f;
    if (lane < cnt) {
      const int pos = base + lane;
      v = __ldg(p.indices + pos);
      const float4 rec = __ldg(drec_h + v);
      float z = el_u + rec.x;
      if (eb_h) z += __ldg(eb_h + pos);
      const float s = leaky_relu(z, p.slope);
      alpha = (s == -INFINITY) ? 0.f : expf(s - rec.y) * rec.z;
      dz = z > 0.f ? 1.f : p.slope;
      t = rec.w;
      if (am_h) amul = __ldg(am_h + pos);
      else if (philox) amul = philox_dropout_mul(p.seed, (uint32_t)__ldg(p.eid + pos), (uint32_t)h, p.attn_p, p.inv_keep);
    }
    const float w_lane = alpha * amul;
    float d_lane = 0.f;
    for (int e = 0; e < cnt; e += EPS) {
      const int my = e + grp;
      const int vv = __shfl_sync(kFull, v, my & 31);
      const float w = __shfl_sync(kFull, w_lane, my & 31);
      float part = 0.f;
      if (my < cnt) {
        const float* r = g_h + (int64_t)vv * p.ld_g;
        Vec<VW> x[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          if (act[i]) x[i].load(r + i * gstride); else x[i].zero();
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          acc[i].fma(w, x[i]);
          part = x[i].dot(fu[i], part);
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
      deliver(part, e, EPS, p.gshift, lane, d_lane);
    }
    // d_lane = <src_scale*ft[u], g'[v]>; softmax + leaky_relu adjoint
    gel_lane += alpha * (d_lane * amul - t) * dz;
  }

  for (int o = G; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i].add_shfl_xor(o);
  }
  if (grp == 0) {
    float* o = p.grad_ft + (int64_t)row * p.ld_gft + h * p.D + j * VW;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (act[i]) {
        acc[i].scale(csu);
        acc[i].store(o + i * gstride);
      }
    }
  }
  const float gel = warp_sum(gel_lane);
  if (lane == 0) p.grad_el[(int64_t)row * p.H + h] = gel;
}

// ---------------------------------------------------------------------------
// dst pass: one warp per (head, destination row v) over the in-CSR
// ---------------------------------------------------------------------------
template <int VW, int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) gat_bwd_dst_kernel(const BwdParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x / p.blocks_per_slab;
  const int row = (blockIdx.x - h * p.blocks_per_slab) * kWarpsPerBlock + warp;
  if (row >= p.n_rows) return;
  const int G = 1 << p.gshift;
  const int j = lane & (G - 1);
  const int grp = lane >> p.gshift;
  const int EPS = 32 >> p.gshift;
  const int gstride = G * VW;

  bool act[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) act[i] = (i * G + j) * VW < p.D;

  const int beg = p.indptr[row], end = p.indptr[row + 1];
  const float4 rec = p.drec[(int64_t)h * p.n_dst + row];  // {er, max, 1/sum, t}
  Vec<VW> gv[VPL];
  {
    const float* g = p.g + (int64_t)row * p.ld_g + h * p.D + j * VW;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (act[i]) gv[i].load(g + i * gstride); else gv[i].zero();
    }
  }
  const float* __restrict__ ft_h = p.ft + h * p.D + j * VW;
  const float* __restrict__ eb_h = p.eb ? p.eb + (int64_t)(p.Hb == 1 ? 0 : h) * p.n_edges : nullptr;
  const float* __restrict__ am_h = p.am ? p.am + (int64_t)h * p.n_edges : nullptr;
  const bool philox = (p.am == nullptr) && p.attn_p > 0.f;
  float ger_lane = 0.f;

  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int u = 0;
    float alpha = 0.f, mul = 1.f, dz = 0.f;
    if (lane < cnt) {
      const int pos = base + lane;
      u = __ldg(p.indices + pos);
      float z = __ldg(p.el + (int64_t)u * p.H + h) + rec.x;
      if (eb_h) z += __ldg(eb_h + pos);
      const float s = leaky_relu(z, p.slope);
      alpha = (s == -INFINITY) ? 0.f : expf(s - rec.y) * rec.z;
      dz = z > 0.f ? 1.f : p.slope;
      if (p.cs) mul = __ldg(p.cs + u);
      if (am_h) mul *= __ldg(am_h + pos);
      else if (philox) mul *= philox_dropout_mul(p.seed, (uint32_t)__ldg(p.eid + pos), (uint32_t)h, p.attn_p, p.inv_keep);
    }
    float d_lane = 0.f;
    for (int e = 0; e < cnt; e += EPS) {
      const int my = e + grp;
      const int uu = __shfl_sync(kFull, u, my & 31);
      float part = 0.f;
      if (my < cnt) {
        const float* r = ft_h + (int64_t)uu * p.ld_ft;
        Vec<VW> x[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          if (act[i]) x[i].load(r + i * gstride); else x[i].zero();
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) part = x[i].dot(gv[i], part);
      }
      for (int o = G >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
      deliver(part, e, EPS, p.gshift, lane, d_lane);
    }
    // d_lane = <ft[u], g'[v]>; mul = src_scale[u] * dropout multiplier
    const float gz = alpha * (d_lane * mul - rec.w) * dz;
    if (p.gz && lane < cnt) p.gz[(int64_t)h * p.n_edges + base + lane] = gz;
    ger_lane += gz;
  }
  const float ger = warp_sum(ger_lane);
  if (p.grad_er && lane == 0) p.grad_er[(int64_t)row * p.H + h] = ger;
}

#define BG_VPL_SWITCH(KERNEL, VW)                                                        \
  switch (vpl) {                                                                         \
    case 1: KERNEL<VW, 1><<<grid, block, 0, st>>>(p); BG_LAUNCHED(1); break;                             \
    case 2: KERNEL<VW, 2><<<grid, block, 0, st>>>(p); BG_LAUNCHED(1); break;                             \
    case 3: KERNEL<VW, 3><<<grid, block, 0, st>>>(p); BG_LAUNCHED(1); break;                             \
    case 4: KERNEL<VW, 4><<<grid, block, 0, st>>>(p); BG_LAUNCHED(1); break;                             \
    case 5: KERNEL<VW, 5><<<grid, block, 0, st>>>(p); BG_LAUNCHED(1); break;                             \
    case 6: KERNEL<VW, 6><<<grid, block, 0, st>>>(p); BG_LAUNCHED(1); break;                             \
    case 8: KERNEL<VW, 8><<<grid, block, 0, st>>>(p); BG_LAUNCHED(1); break;                             \
    default: set_error("backward: unsupported vectors-per-lane %d", vpl); return -1;     \
  }

static int launch_src(const BwdParams& p, int vw, int vpl, dim3 grid, cudaStream_t st) {
  dim3 block(kWarpsPerBlock * 32);
  if (vw == 4) { BG_VPL_SWITCH(gat_bwd_src_kernel, 4) }
  else if (vw == 2) { BG_VPL_SWITCH(gat_bwd_src_kernel, 2) }
  else { BG_VPL_SWITCH(gat_bwd_src_kernel, 1) }
  return 0;
}
static int launch_dst(const BwdParams& p, int vw, int vpl, dim3 grid, cudaStream_t st) {
  dim3 block(kWarpsPerBlock * 32);
  if (vw == 4) { BG_VPL_SWITCH(gat_bwd_dst_kernel, 4) }
  else if (vw == 2) { BG_VPL_SWITCH(gat_bwd_dst_kernel, 2) }
  else { BG_VPL_SWITCH(gat_bwd_dst_kernel, 1) }
  return 0;
}

}  // namespace botgat

using namespace botgat;

extern "C" int botgat_gat_backward(const botgat_graph* g, const botgat_bwd_args* a, void* stream) {
  BG_REQUIRE(g && a, "backward: null graph/args");
  BG_REQUIRE(a->H > 0 && a->D > 0, "backward: bad H=%d D=%d", a->H, a->D);
  BG_REQUIRE(a->ft && a->el && a->out && a->row_max && a->row_sum && a->gout, "backward: null input");
  BG_REQUIRE(a->drec && a->grad_ft && a->grad_el, "backward: null drec/grad_ft/grad_el");
  BG_REQUIRE(!a->dst_scale || a->gprime, "backward: gprime workspace required with dst_scale");
  const int64_t HD = (int64_t)a->H * a->D;
  BG_REQUIRE(a->ld_ft >= HD && a->ld_out >= HD && a->ld_gft >= HD, "backward: leading dimension < H*D");
  BG_REQUIRE(a->eb_out ? (a->Hb == 1 || a->Hb == a->H) : true, "backward: Hb must be 1 or H");
  BG_REQUIRE((a->eb_in == nullptr) == (a->eb_out == nullptr) || !(a->grad_er || a->gz),
             "backward: eb_in and eb_out must both be given");
  if (g->n_dst == 0 || g->n_src == 0) return 0;
  DeviceGuard guard(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 block(kWarpsPerBlock * 32);

  const int phases = a->phases ? a->phases : 7;
  // node pass
  if (phases & 1) {
    dim3 grid((unsigned)((g->n_dst + kWarpsPerBlock - 1) / kWarpsPerBlock));
    gat_bwd_node_kernel<<<grid, block, 0, st>>>((int)g->n_dst, a->H, a->D, a->ld_out, a->out, a->gout, a->er,
                                                a->row_max, a->row_sum, a->dst_scale, (float4*)a->drec,
                                                a->dst_scale ? a->gprime : nullptr); BG_LAUNCHED(1);
    BG_CHECK(cudaGetLastError());
  }
  const float* gp = a->dst_scale ? a->gprime : a->gout;

  BwdParams p;
  p.n_dst = (int)g->n_dst; p.n_edges = g->n_edges;
  p.H = a->H; p.D = a->D; p.ld_ft = a->ld_ft; p.ld_g = a->ld_out; p.ld_gft = a->ld_gft;
  p.ft = a->ft; p.el = a->el; p.cs = a->src_scale; p.g = gp; p.drec = (const float4*)a->drec;
  p.Hb = a->Hb; p.slope = a->slope; p.attn_p = a->attn_p; p.inv_keep = 1.f / (1.f - a->attn_p); p.seed = a->seed;
  p.grad_ft = a->grad_ft; p.grad_el = a->grad_el; p.grad_er = a->grad_er; p.gz = a->gz;

  // src pass (out-CSR): the gathered table is g' (ld_out), the row-local one is ft
  if (phases & 2) {
    Tiling t = choose_tiling(a->D, a->ld_out, a->ld_ft, gp, a->ft, 1, g->n_dst, 8);
    // grad_ft stores use the same vector width
    if ((a->ld_gft % t.vw) != 0 || ((uintptr_t)a->grad_ft % (t.vw * 4)) != 0)
      t = choose_tiling(a->D, 1, 1, gp, a->ft, 1, g->n_dst, 8);  // forces vw = 1
    BG_REQUIRE(t.col_parts == 1, "backward: D=%d too wide for one pass (max %d)", a->D, 32 * 8 * t.vw);
    p.indptr = g->out_indptr; p.indices = g->out_indices; p.eid = g->out_eid;
    p.n_rows = (int)g->n_src; p.eb = a->eb_out; p.am = a->am_out; p.gshift = t.gshift;
    p.blocks_per_slab = (p.n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t nblocks = (int64_t)p.blocks_per_slab * a->H;
    BG_REQUIRE(nblocks < (1ll << 31), "backward: grid too large");
    int rc = launch_src(p, t.vw, t.vpl, dim3((unsigned)nblocks), st);
    if (rc) return rc;
    BG_CHECK(cudaGetLastError());
  }
  // dst pass (in-CSR), only when something needs it
  if ((phases & 4) && (a->grad_er || a->gz)) {
    Tiling t = choose_tiling(a->D, a->ld_ft, a->ld_out, a->ft, gp, 1, g->n_src, 8);
    BG_REQUIRE(t.col_parts == 1, "backward: D=%d too wide for one pass (max %d)", a->D, 32 * 8 * t.vw);
    p.indptr = g->in_indptr; p.indices = g->in_indices; p.eid = g->in_eid;
    p.n_rows = (int)g->n_dst; p.eb = a->eb_in; p.am = a->am_in; p.gshift = t.gshift;
    p.blocks_per_slab = (p.n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t nblocks = (int64_t)p.blocks_per_slab * a->H;
    BG_REQUIRE(nblocks < (1ll << 31), "backward: grid too large");
    int rc = launch_dst(p, t.vw, t.vpl, dim3((unsigned)nblocks), st);
    if (rc) return rc;
    BG_CHECK(cudaGetLastError());
  }
  return 0;
}
