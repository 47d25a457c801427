// Fused GAT forward for sm_100a:
//   per-edge logits (el[src] + er[dst] + eb[edge]) -> leaky_relu -> online
//   per-destination softmax -> attention dropout -> weighted neighbour sum ->
//   degree scaling, in ONE pass over the in-CSR.  No E-sized intermediate is
//   written.  Replaces src/no-sampling/models.py:500-505,523-555 and
//   src/ogbn-proteins/models.py:125-156 of the reference (DGL apply_edges,
//   edge_softmax, update_all and the torch elementwise ops between them).
//
// L2/HBM-bound gather: no tensor cores (the only dense contraction, fc, stays a
// torch matmul).  See common.cuh "Work decomposition" for the item/lane layout.
//
// Latency hiding (the kernel is load-latency bound, profiles/r01_a_*):
//   * the per-neighbour scalars run in a 3-stage software pipeline — index of
//     chunk c+2, logit operands of chunk c+1 and the row gathers of chunk c are in
//     flight together, so no load waits on a load issued in the same iteration;
//   * NS steps of row gathers (NS * VPL 128-bit loads per lane) are issued before
//     the first FMA consumes them.
#include "common.cuh"

namespace botgat {

struct FwdParams {
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eid;
  int n_rows;
  int64_t n_edges;
  int H, D;
  int64_t ld_ft, ld_out;
  const float *ft, *el, *er, *eb, *am, *cs, *ds;
  int Hb;
  float slope, attn_p, inv_keep;
  uint64_t seed;
  float *out, *row_max, *row_sum;
  int col_parts, part_cols, gshift, omask;
  int blocks_per_slab;
};

template <int VW, int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) gat_fwd_kernel(const FwdParams p) {
  constexpr int NS = steps_in_flight(VPL);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x / p.blocks_per_slab;
  const int row = (blockIdx.x - slab * p.blocks_per_slab) * kWarpsPerBlock + warp;
  if (row >= p.n_rows) return;
  const int h = slab / p.col_parts;
  const int cp = slab - h * p.col_parts;
  const int c0 = cp * p.part_cols;
  const int ncols = min(p.D - c0, p.part_cols);

  const int G = 1 << p.gshift;
  const int grp = lane >> p.gshift;
  const int EPS = 32 >> p.gshift;  // neighbours per warp step
  const int gstride = G * VW;      // floats between a lane's consecutive vector slots
  // first vector of this lane, shifted so that every slot is 128-byte-line aligned
  const int v0 = (lane & (G - 1)) - (((h * p.D + c0) / VW) & p.omask);

  bool act[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = v0 + i * G;
    act[i] = v >= 0 && v * VW < ncols;
  }

  const int beg = p.indptr[row], end = p.indptr[row + 1];
  const float er_v = p.er ? p.er[(int64_t)row * p.H + h] : 0.f;
  const float* __restrict__ ft_h = p.ft + h * p.D + c0 + v0 * VW;
  const float* __restrict__ el_h = p.el + h;
  const float* __restrict__ eb_h = p.eb ? p.eb + (int64_t)(p.Hb == 1 ? 0 : h) * p.n_edges : nullptr;
  const float* __restrict__ am_h = p.am ? p.am + (int64_t)h * p.n_edges : nullptr;
  const bool philox = (p.am == nullptr) && p.attn_p > 0.f;

  Vec<VW> acc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) acc[i].zero();
  float m = -INFINITY;  // running row max (warp-uniform)
  float l_lane = 0.f;   // this lane's share of sum exp(s - m)

  // ---- software pipeline over 32-neighbour chunks ----
  // stage 0: neighbour index; stage 1: logit operands (need the index); stage 2: row gathers
  int u1 = 0, u2 = 0;                       // indices of chunk c+1 (u1) and c+2 (u2)
  float z0 = -INFINITY, mul0 = 1.f;         // chunk c operands, complete
  int u0 = 0;
  auto load_index = [&](int base) -> int {
    const int pos = base + lane;
    return pos < end ? __ldg(p.indices + pos) : 0;
  };
  auto load_operands = [&](int base, int u, float& z, float& mul) {
    const int pos = base + lane;
    z = -INFINITY;
    mul = 1.f;
    if (pos < end) {
      z = __ldg(el_h + (int64_t)u * p.H) + er_v;
      if (eb_h) z += __ldg(eb_h + pos);
      if (p.cs) mul = __ldg(p.cs + u);
      if (am_h) mul *= __ldg(am_h + pos);
      else if (philox) mul *= philox_dropout_mul(p.seed, (uint32_t)__ldg(p.eid + pos), (uint32_t)h, p.attn_p, p.inv_keep);
    }
  };
  u0 = load_index(beg);
  u1 = load_index(beg + 32);
  load_operands(beg, u0, z0, mul0);

  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    // issue next stages' loads first; they are consumed one iteration later
    u2 = load_index(base + 64);
    float z1, mul1;
    load_operands(base + 32, u1, z1, mul1);

    // ---- online softmax on chunk c ----
    const float s = leaky_relu(z0, p.slope);  // -inf stays -inf (dropped edge / lane past the row end)
    const float m_new = fmaxf(m, warp_max(s));
    if (m_new > m) {
      const float f = expf(m - m_new);  // m == -inf -> 0, and everything accumulated so far is 0
      l_lane *= f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) acc[i].scale(f);
      m = m_new;
    }
    const float pexp = (s == -INFINITY) ? 0.f : expf(s - m);
    l_lane += pexp;
    const float w_lane = pexp * mul0;

    // ---- group = neighbour: gather slab rows, acc += w * row; NS steps in flight ----
    for (int e = 0; e < cnt; e += NS * EPS) {
      Vec<VW> x[NS][VPL];
      float w[NS];
#pragma unroll
      for (int s_ = 0; s_ < NS; ++s_) {
        const int my = e + s_ * EPS + grp;
        const int uu = __shfl_sync(kFull, u0, my & 31);
        const float ww = __shfl_sync(kFull, w_lane, my & 31);
        const bool ok = my < cnt;
        w[s_] = ok ? ww : 0.f;
        const float* r = ft_h + (int64_t)uu * p.ld_ft;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          if (ok && act[i]) x[s_][i].load_stream(r + i * gstride); else x[s_][i].zero();
        }
      }
#pragma unroll
      for (int s_ = 0; s_ < NS; ++s_) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) acc[i].fma(w[s_], x[s_][i]);
      }
    }
    u0 = u1; u1 = u2; z0 = z1; mul0 = mul1;
  }

  // ---- epilogue: combine the groups, normalise, degree-scale, store ----
  const float l = warp_sum(l_lane);
  for (int o = G; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) acc[i].add_shfl_xor(o);
  }
  float scale = l > 0.f ? 1.f / l : 0.f;
  if (p.ds) scale *= p.ds[row];
  if (grp == 0) {
    float* o = p.out + (int64_t)row * p.ld_out + h * p.D + c0 + v0 * VW;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (act[i]) {
        acc[i].scale(scale);
        acc[i].store(o + i * gstride);
      }
    }
  }
  if (cp == 0 && lane == 0) {
    p.row_max[(int64_t)row * p.H + h] = m;
    p.row_sum[(int64_t)row * p.H + h] = l;
  }
}

template <int VW>
static int launch_fwd_vw(const FwdParams& p, int vpl, dim3 grid, cudaStream_t st) {
  dim3 block(kWarpsPerBlock * 32);
  switch (vpl) {
    case 1: gat_fwd_kernel<VW, 1><<<grid, block, 0, st>>>(p); break;
    case 2: gat_fwd_kernel<VW, 2><<<grid, block, 0, st>>>(p); break;
    case 3: gat_fwd_kernel<VW, 3><<<grid, block, 0, st>>>(p); break;
    case 4: gat_fwd_kernel<VW, 4><<<grid, block, 0, st>>>(p); break;
    case 5: gat_fwd_kernel<VW, 5><<<grid, block, 0, st>>>(p); break;
    case 6: gat_fwd_kernel<VW, 6><<<grid, block, 0, st>>>(p); break;
    case 8: gat_fwd_kernel<VW, 8><<<grid, block, 0, st>>>(p); break;
    default: set_error("forward: unsupported vector slots per lane %d", vpl); return -1;
  }
  BG_LAUNCHED(1);
  return 0;
}

}  // namespace botgat

using namespace botgat;

extern "C" int botgat_gat_forward(const botgat_graph* g, const botgat_fwd_args* a, void* stream) {
  BG_REQUIRE(g && a, "forward: null graph/args");
  BG_REQUIRE(a->H > 0 && a->D > 0, "forward: bad H=%d D=%d", a->H, a->D);
  BG_REQUIRE(a->ft && a->el && a->out && a->row_max && a->row_sum, "forward: null ft/el/out/row_max/row_sum");
  BG_REQUIRE(a->ld_ft >= (int64_t)a->H * a->D && a->ld_out >= (int64_t)a->H * a->D, "forward: leading dimension < H*D");
  BG_REQUIRE(a->eb ? (a->Hb == 1 || a->Hb == a->H) : true, "forward: Hb must be 1 or H when eb is given (got %d)", a->Hb);
  BG_REQUIRE(a->attn_p >= 0.f && a->attn_p < 1.f, "forward: attn_p must be in [0,1)");
  if (g->n_dst == 0) return 0;
  DeviceGuard guard(g->device);
  cudaStream_t st = (cudaStream_t)stream;

  Tiling t = choose_tiling(a->H, a->D, a->ld_ft, a->ft, a->ld_out, a->out, a->col_parts, g->n_src);
  FwdParams p;
  p.indptr = g->in_indptr; p.indices = g->in_indices; p.eid = g->in_eid;
  p.n_rows = (int)g->n_dst; p.n_edges = g->n_edges;
  p.H = a->H; p.D = a->D; p.ld_ft = a->ld_ft; p.ld_out = a->ld_out;
  p.ft = a->ft; p.el = a->el; p.er = a->er; p.eb = a->eb; p.am = a->am;
  p.cs = a->src_scale; p.ds = a->dst_scale; p.Hb = a->Hb;
  p.slope = a->slope; p.attn_p = a->attn_p; p.inv_keep = 1.f / (1.f - a->attn_p); p.seed = a->seed;
  p.out = a->out; p.row_max = a->row_max; p.row_sum = a->row_sum;
  p.col_parts = t.col_parts; p.part_cols = t.part_cols; p.gshift = t.gshift; p.omask = t.omask;
  p.blocks_per_slab = (p.n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t nblocks = (int64_t)p.blocks_per_slab * a->H * t.col_parts;
  BG_REQUIRE(nblocks < (1ll << 31), "forward: grid too large");
  dim3 grid((unsigned)nblocks);
  int rc;
  if (t.vw == 4) rc = launch_fwd_vw<4>(p, t.vpl, grid, st);
  else if (t.vw == 2) rc = launch_fwd_vw<2>(p, t.vpl, grid, st);
  else rc = launch_fwd_vw<1>(p, t.vpl, grid, st);
  if (rc) return rc;
  BG_CHECK(cudaGetLastError());
  return 0;
}
