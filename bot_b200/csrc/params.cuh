// Kernel parameter blocks shared by the warp-per-row and the group-per-row gather kernels.
#pragma once
#include "common.cuh"

namespace botgat {

// Fused layer epilogue of the forward (the elementwise tail of a reference layer in inference: residual adds, the
// eval-mode norm as a per-column scale / shift, ReLU; src/no-sampling/models.py:720-731,
// src/ogbn-proteins/models.py:159-160,253-260), applied where an output vector is still in registers:
//   out[v, c] = dst_scale[v] * agg[v, c] + res[v, c] + res2[v, c]       y[v, c] = act(out[v, c] * scale[c] + shift[c])
struct Epilogue {
  const float *res, *res2;
  int64_t ld_res, ld_res2;
  const float *scale, *shift;  // (H*D) or null (= 1 / 0)
  int relu;
  float* y;
  int64_t ld_y;
  __host__ __device__ bool any() const { return res || res2 || y; }
  // `a` holds dst_scale * agg of columns [col, col + VW) of row `row`; on return it holds `out`
  template <int VW>
  __device__ __forceinline__ void apply(Vec<VW>& a, int64_t row, int64_t col) const {
#ifdef BG_NO_EPILOGUE  // developer A/B builds: what the fused tail costs the training-path forward
    return;
#endif
    Vec<VW> t;
    if (res) { t.load(res + row * ld_res + col); a.add(t); }
    if (res2) { t.load(res2 + row * ld_res2 + col); a.add(t); }
    if (y) {
      Vec<VW> r = a;
      if (scale) { t.load(scale + col); r.mul(t); }
      if (shift) { t.load(shift + col); r.add(t); }
      if (relu) r.relu();
      r.store(y + row * ld_y + col);
    }
  }
};

struct FwdParams {
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eid;
  int n_rows;
  int64_t n_src_table;  // rows of the gathered table ft
  int64_t n_edges;
  int H, D;
  int64_t ld_ft, ld_out;
  const float *ft, *el, *er, *eb, *am, *cs, *ds;
  const float *ee, *amul_e;  // edge-id-ordered operands (direct mode)
  const uint8_t* keep;
  int Hb;
  float slope, attn_p, inv_keep;
  uint64_t seed;
  float *out, *row_max, *row_sum;
  Epilogue ep;
  int col_parts, part_cols, omask;
  int blocks_per_slab;
  int h_begin, h_count;  // head range of this launch
  // row splitting (segments.cu): work items are segments when seg_row != nullptr
  const int32_t *seg_row, *seg_beg, *seg_end, *seg_slot;
  int n_items;
  float* scratch;
};

struct BwdParams {
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eid;
  int n_rows;  // rows of the out-CSR (= n_src)
  int n_dst;
  int64_t n_edges;
  int H, D;
  int64_t ld_ft, ld_g, ld_gft;
  const float *ft, *el, *eb, *am, *cs;
  const float *ee, *amul_e;  // edge-id-ordered operands (direct mode)
  const uint8_t* keep;
  float* gz_e;         // (n_edges, H) edge-id order (direct mode), written by the src pass
  const float* g;      // g' (n_dst, ld_g)
  const float4* drec;  // per-destination records.  The head-major kernels read them head-major, drec[h * n_dst + v]; the
                       // all-heads-per-row kernels (gat_rowwise.cu) through the strides below, which are node-major
                       // (hs = 1, vs = H: one contiguous 16*H-byte read per edge) whenever drec_node_major() says so
  int drec_hs, drec_vs;  // 32-bit: n_dst * H < 2^31 is checked at launch
  int Hb;
  float slope, attn_p, inv_keep;
  uint64_t seed;
  float *grad_ft, *grad_el, *gz;
  int omask;
  int blocks_per_slab;
  int h_begin, h_count;  // head range of this launch
  const int32_t *seg_row, *seg_beg, *seg_end, *seg_slot;
  int n_items;
  float* scratch;
};

struct SrcOps {
  float4 rec;  // {er[v], row_max[v], 1/row_sum[v], t[v]}
  // raw loads, combined one pipeline stage later by logit_term(): consuming a load where it is issued would
  // stall the warp there (even a predicated-off FADD waits for its source register)
  // every field has one producer (a default written before the loads + at most one predicated load): an
  // if / else-if chain would leave a default MOV behind the loads that waits on their scoreboard slot
  float eb, ee, amul, ame, amp;
  int kp;
  __device__ __forceinline__ float logit_term() const { return kp ? eb + ee : -INFINITY; }
  __device__ __forceinline__ float multiplier() const { return amul * ame * amp; }
};

// Low-degree graphs (average row shorter than one 32-neighbour chunk and a half) use the group-per-row kernels
// (gat_lowdeg.cu): a G-lane group owns a row, so a warp works on 32/G rows at once and the per-row latency
// chain (index -> logit operands -> row gathers) and epilogue are shared.  BOTGAT_LOWDEG overrides the threshold.
bool use_lowdeg_kernels(int64_t n_edges, int64_t n_rows, bool backward);
int launch_fwd_lowdeg(const FwdParams& p, const Tiling& t, cudaStream_t st);
int launch_src_lowdeg(const BwdParams& p, const Tiling& t, cudaStream_t st);
// gat_bwd_tma.cu: the warp-per-row src pass with TMA-staged rows; returns 1 when the shape is not covered (caller
// falls back to gat_bwd_src_kernel), 0 when launched, < 0 on error
int launch_src_tma(const BwdParams& p, const Tiling& t, cudaStream_t st);
// gat_rowwise.cu: one warp per row for ALL heads, for gathered tables far beyond the L2 (no head slab can be resident);
// same return convention as launch_src_tma
// layout of the per-destination records: a function of the shape alone, so that the node phase (writer) and every src
// kernel (readers) agree even when the phases are launched by separate calls
bool drec_node_major(int H, int D, int64_t n_dst, int64_t n_edges, int64_t n_src);
int launch_fwd_rowwise(const FwdParams& p, const Tiling& t, cudaStream_t st);
int launch_src_rowwise(const BwdParams& p, const Tiling& t, cudaStream_t st);
int segment_length();
// floats per scratch slot (rounded to 4 so that float4 stores into a slot stay aligned)
__host__ __device__ inline int64_t fwd_slot_floats(int H, int D) { return ((int64_t)H * (D + 2) + 3) / 4 * 4; }
__host__ __device__ inline int64_t bwd_slot_floats(int H, int D) { return ((int64_t)H * (D + 1) + 3) / 4 * 4; }
int launch_fwd_combine(const botgat_graph::SegTable& t, int H, int D, int64_t ld_out, const float* scratch,
                       const float* ds, float* out, float* row_max, float* row_sum, const Epilogue& ep, cudaStream_t st);
int launch_bwd_combine(const botgat_graph::SegTable& t, int H, int D, int64_t ld_gft, const float* scratch,
                       const float* cs, float* grad_ft, float* grad_el, cudaStream_t st);

}  // namespace botgat
