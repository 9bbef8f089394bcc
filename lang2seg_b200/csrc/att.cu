// att2in2 caption-step kernels for sm_100a.
//
//  * att_step fwd/bwd  -- Attention.forward (lib/caption_models/AttModel.py:406-423): score,
//    softmax over the A attention locations and weighted feature sum fused in ONE kernel per
//    decode step.  A thread-block CLUSTER of kCluster CTAs shares one sample: each CTA streams a
//    slice of the locations (p_att and att_feats rows, coalesced float4), keeps flash-style
//    partial (max, sum, weighted accumulator) state, and the partials are merged through
//    distributed shared memory -- so even a batch of 16 samples spreads over 64 SMs.  The
//    reference needs 7 launches and 5 intermediate (B,A,512) tensors for the same step.
//  * att2in2 gates fwd/bwd -- Att2in2Core.forward AttModel.py:450-464 after the three Linears.
//  * logsoftmax + masked NLL fwd/bwd -- AttModel.py:98 + lib/misc/utils.py:43-53.
//  * caption feature prep fwd/bwd -- network_cycle_response.py:428-438.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace l2s {
namespace {

constexpr int kCluster = 4;
constexpr int kAttThreads = 256;
constexpr int kMaxLoc = 256;   // locations per CTA slice held in smem

struct AttGeom {
  int B, A, D, Dh;
  int ldh;     // row stride of att_h (>= Dh): the decode loop reads it in place from the [att_h | sums] buffer
  int ld_dah;  // row stride of the datt_h output
};

__device__ __forceinline__ void slice(int A, int rank, int* a0, int* a1) {
  const int per = (A + kCluster - 1) / kCluster;
  *a0 = min(A, rank * per);
  *a1 = min(A, *a0 + per);
}

// ------------------------------------------------------------------------------- forward
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kAttThreads)
att_step_fwd_kernel(const float* __restrict__ att_h, const float* __restrict__ att_feats,
                    const float* __restrict__ p_att, const float* __restrict__ alpha_w,
                    const float* __restrict__ alpha_b, float* __restrict__ weight, float* __restrict__ att_res,
                    AttGeom g) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / kCluster;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  extern __shared__ __align__(16) float smem[];
  float* s_ah = smem;                 // [Dh] att_h row
  float* s_aw = s_ah + g.Dh;          // [Dh] alpha_w
  float* s_acc = s_aw + g.Dh;         // [D]  partial weighted sum (exchanged through DSMEM)
  __shared__ float s_e[kMaxLoc];
  __shared__ float s_ml[2];           // partial max, partial sum (exchanged through DSMEM)
  __shared__ float s_red[8];

  int a0, a1;
  slice(g.A, rank, &a0, &a1);
  const int na = a1 - a0;
  pdl_launch_dependents();
  pdl_wait();                          // att_h / datt_res come from the previous kernel of the decode chain
  for (int d = t; d < g.Dh; d += kAttThreads) {
    s_ah[d] = __ldg(att_h + (size_t)b * g.ldh + d);
    s_aw[d] = __ldg(alpha_w + d);
  }
  __syncthreads();
  const float ab = __ldg(alpha_b);

  // scores: one warp per location, lanes over Dh (float4)
  for (int a = wid; a < na; a += kAttThreads / 32) {
    const float4* row = reinterpret_cast<const float4*>(p_att + ((size_t)b * g.A + a0 + a) * g.Dh);
    float s = 0.f;
    for (int q = lane; q < g.Dh / 4; q += 32) {
      const float4 p = __ldg(row + q);
      const float4 h = reinterpret_cast<const float4*>(s_ah)[q];
      const float4 w = reinterpret_cast<const float4*>(s_aw)[q];
      s = fmaf(w.x, tanhf_fast_acc(p.x + h.x), s);
      s = fmaf(w.y, tanhf_fast_acc(p.y + h.y), s);
      s = fmaf(w.z, tanhf_fast_acc(p.z + h.z), s);
      s = fmaf(w.w, tanhf_fast_acc(p.w + h.w), s);
    }
    s = warp_sum(s);
    if (lane == 0) s_e[a] = s + ab;
  }
  __syncthreads();
  // partial softmax statistics of this slice
  float m = -INFINITY;
  for (int a = t; a < na; a += kAttThreads) m = fmaxf(m, s_e[a]);
  m = warp_max(m);
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float l = 0.f;
  for (int a = t; a < na; a += kAttThreads) {
    const float ex = expf(s_e[a] - m);
    s_e[a] = ex;
    l += ex;
  }
  l = warp_sum(l);
  if (lane == 0) s_red[wid] = l;
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_red[w];
    s_ml[0] = m;
    s_ml[1] = tot;
  }
  // partial weighted sum over this slice: thread = float4 column group, loop over locations
  for (int q = t; q < g.D / 4; q += kAttThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* col = reinterpret_cast<const float4*>(att_feats + ((size_t)b * g.A + a0) * g.D) + q;
#pragma unroll 4
    for (int a = 0; a < na; ++a) {
      const float4 v = __ldg(col + (size_t)a * (g.D / 4));
      const float w = s_e[a];
      acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
    reinterpret_cast<float4*>(s_acc)[q] = acc;
  }
  cluster.sync();
  // merge the kCluster partial states
  float ms[kCluster], ls[kCluster], M = -INFINITY;
#pragma unroll
  for (int r = 0; r < kCluster; ++r) {
    const float* rm = cluster.map_shared_rank(s_ml, r);
    ms[r] = rm[0];
    ls[r] = rm[1];
    M = fmaxf(M, ms[r]);
  }
  float L = 0.f, sc[kCluster];
#pragma unroll
  for (int r = 0; r < kCluster; ++r) {
    sc[r] = (ls[r] > 0.f) ? expf(ms[r] - M) : 0.f;
    L = fmaf(sc[r], ls[r], L);
  }
  const float invL = 1.f / L;
  // this CTA finalises columns [rank*D/kCluster, ...) of att_res
  const int dper = (g.D + kCluster - 1) / kCluster;
  for (int d = rank * dper + t; d < min(g.D, (rank + 1) * dper); d += kAttThreads) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < kCluster; ++r) v = fmaf(sc[r], cluster.map_shared_rank(s_acc, r)[d], v);
    att_res[(size_t)b * g.D + d] = v * invL;
  }
  const float myscale = sc[rank] * invL;
  for (int a = t; a < na; a += kAttThreads) weight[(size_t)b * g.A + a0 + a] = s_e[a] * myscale;
  cluster.sync();   // nobody leaves while its shared memory may still be read remotely
}

// ------------------------------------------------------------------------------- backward
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kAttThreads)
att_step_bwd_kernel(const float* __restrict__ datt_res, const float* __restrict__ att_h,
                    const float* __restrict__ att_feats, const float* __restrict__ p_att,
                    const float* __restrict__ alpha_w, const float* __restrict__ weight,
                    float* __restrict__ datt_h, float* __restrict__ de_out, float* __restrict__ dp_att,
                    float* __restrict__ datt_feats, float* __restrict__ dalpha_w, AttGeom g) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / kCluster;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  extern __shared__ __align__(16) float smem[];
  float* s_ah = smem;                 // [Dh]
  float* s_aw = s_ah + g.Dh;          // [Dh]
  float* s_dah = s_aw + g.Dh;         // [Dh] partial d att_h (DSMEM exchanged)
  float* s_do = s_dah + g.Dh;         // [D]  datt_res row
  __shared__ float s_w[kMaxLoc], s_dpi[kMaxLoc];
  __shared__ float s_part[1];
  __shared__ float s_red[8];

  int a0, a1;
  slice(g.A, rank, &a0, &a1);
  const int na = a1 - a0;
  pdl_launch_dependents();
  pdl_wait();                          // att_h / datt_res come from the previous kernel of the decode chain
  for (int d = t; d < g.Dh; d += kAttThreads) {
    s_ah[d] = __ldg(att_h + (size_t)b * g.ldh + d);
    s_aw[d] = __ldg(alpha_w + d);
  }
  for (int d = t; d < g.D; d += kAttThreads) s_do[d] = __ldg(datt_res + (size_t)b * g.D + d);
  for (int a = t; a < na; a += kAttThreads) s_w[a] = __ldg(weight + (size_t)b * g.A + a0 + a);
  __syncthreads();

  // d pi_a = <datt_res, att_feats[a]> ; optional datt_feats[a] += pi_a * datt_res
  for (int a = wid; a < na; a += kAttThreads / 32) {
    const size_t off = ((size_t)b * g.A + a0 + a) * g.D;
    const float4* row = reinterpret_cast<const float4*>(att_feats + off);
    const float wa = s_w[a];
    float s = 0.f;
    for (int q = lane; q < g.D / 4; q += 32) {
      const float4 v = __ldg(row + q);
      const float4 d = reinterpret_cast<const float4*>(s_do)[q];
      s = fmaf(v.x, d.x, s); s = fmaf(v.y, d.y, s); s = fmaf(v.z, d.z, s); s = fmaf(v.w, d.w, s);
      if (datt_feats) {
        float4* gp = reinterpret_cast<float4*>(datt_feats + off) + q;
        float4 o = *gp;
        o.x = fmaf(wa, d.x, o.x); o.y = fmaf(wa, d.y, o.y); o.z = fmaf(wa, d.z, o.z); o.w = fmaf(wa, d.w, o.w);
        *gp = o;
      }
    }
    s = warp_sum(s);
    if (lane == 0) s_dpi[a] = s;
  }
  __syncthreads();
  // S = sum_a pi_a dpi_a over the whole sample (cluster-wide)
  float part = 0.f;
  for (int a = t; a < na; a += kAttThreads) part = fmaf(s_w[a], s_dpi[a], part);
  part = warp_sum(part);
  if (lane == 0) s_red[wid] = part;
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_red[w];
    s_part[0] = tot;
  }
  cluster.sync();
  float S = 0.f;
#pragma unroll
  for (int r = 0; r < kCluster; ++r) S += cluster.map_shared_rank(s_part, r)[0];
  // de_a = pi_a (dpi_a - S)   (reuse s_dpi)
  for (int a = t; a < na; a += kAttThreads) {
    const float de = s_w[a] * (s_dpi[a] - S);
    s_dpi[a] = de;
    de_out[(size_t)b * g.A + a0 + a] = de;
  }
  __syncthreads();
  // g[a,d] = de_a * alpha_d * (1 - th^2): column sums -> d att_h ; de_a*th -> d alpha_w ; += d p_att
  // thread = (float4 column group q, location phase ph): nph phases split the locations of the slice so
  // that all 256 threads work when Dh/4 < 256; the phases' column sums meet in shared memory.
  const int nq = g.Dh / 4;
  const int nph = (nq <= kAttThreads / 2) ? 2 : 1;
  const int items = nq * nph;
  for (int base = 0; base < items; base += kAttThreads) {      // uniform trip count: barriers inside
    const int qq = base + t;
    const bool on = qq < items;
    const int q = on ? qq % nq : 0, ph = on ? qq / nq : 0;
    const float4 h = reinterpret_cast<const float4*>(s_ah)[q];
    const float4 w = reinterpret_cast<const float4*>(s_aw)[q];
    float4 dah = make_float4(0.f, 0.f, 0.f, 0.f), daw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int a = on ? ph : na; a < na; a += nph) {
      const size_t off = ((size_t)b * g.A + a0 + a) * g.Dh;
      const float4 p = __ldg(reinterpret_cast<const float4*>(p_att + off) + q);
      const float de = s_dpi[a];
      float4 th, gg;
      th.x = tanhf_fast_acc(p.x + h.x); th.y = tanhf_fast_acc(p.y + h.y);
      th.z = tanhf_fast_acc(p.z + h.z); th.w = tanhf_fast_acc(p.w + h.w);
      gg.x = de * w.x * (1.f - th.x * th.x); gg.y = de * w.y * (1.f - th.y * th.y);
      gg.z = de * w.z * (1.f - th.z * th.z); gg.w = de * w.w * (1.f - th.w * th.w);
      dah.x += gg.x; dah.y += gg.y; dah.z += gg.z; dah.w += gg.w;
      daw.x = fmaf(de, th.x, daw.x); daw.y = fmaf(de, th.y, daw.y);
      daw.z = fmaf(de, th.z, daw.z); daw.w = fmaf(de, th.w, daw.w);
      if (dp_att) {
        float4* gp = reinterpret_cast<float4*>(dp_att + off) + q;
        float4 o = *gp;
        o.x += gg.x; o.y += gg.y; o.z += gg.z; o.w += gg.w;
        *gp = o;
      }
    }
    // two phases at most: phase 1 adds after phase 0 stored (ordered by the barrier below)
    if (on && ph == 0) reinterpret_cast<float4*>(s_dah)[q] = dah;
    __syncthreads();
    if (on && ph == 1) {
      float4 o = reinterpret_cast<float4*>(s_dah)[q];
      o.x += dah.x; o.y += dah.y; o.z += dah.z; o.w += dah.w;
      reinterpret_cast<float4*>(s_dah)[q] = o;
    }
    if (on && dalpha_w) {
      atomicAdd(dalpha_w + 4 * q + 0, daw.x);
      atomicAdd(dalpha_w + 4 * q + 1, daw.y);
      atomicAdd(dalpha_w + 4 * q + 2, daw.z);
      atomicAdd(dalpha_w + 4 * q + 3, daw.w);
    }
  }
  cluster.sync();
  const int dper = (g.Dh + kCluster - 1) / kCluster;
  for (int d = rank * dper + t; d < min(g.Dh, (rank + 1) * dper); d += kAttThreads) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < kCluster; ++r) v += cluster.map_shared_rank(s_dah, r)[d];
    datt_h[(size_t)b * g.ld_dah + d] = v;
  }
  cluster.sync();
}

// ------------------------------------------------------------------------------- gates
// sums has row stride lds (the decode loop keeps [att_h | sums] rows in one buffer); c_prev may be
// NULL (zero state at t = 0); dh = dh_a + dh_b (either may be NULL); dsums has row stride ld_ds.
__global__ void gates_fwd_kernel(const float* __restrict__ sums, int lds, const float* __restrict__ a2c,
                                 const float* __restrict__ c_prev, float* __restrict__ h, float* __restrict__ c,
                                 int B, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx - b * D;
  const float* s = sums + (size_t)b * lds;
  const float ig = sigmoidf_acc(s[d]), fg = sigmoidf_acc(s[D + d]), og = sigmoidf_acc(s[2 * D + d]);
  const float g1 = s[3 * D + d] + a2c[(size_t)b * 2 * D + d];
  const float g2 = s[4 * D + d] + a2c[(size_t)b * 2 * D + D + d];
  const float gg = fmaxf(g1, g2);
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const float cn = fmaf(fg, cp, ig * gg);
  c[idx] = cn;
  h[idx] = og * tanhf(cn);
}

__global__ void gates_bwd_kernel(const float* __restrict__ sums, int lds, const float* __restrict__ a2c,
                                 const float* __restrict__ c_prev, const float* __restrict__ c,
                                 const float* __restrict__ dh_a, const float* __restrict__ dh_b,
                                 const float* __restrict__ dc, float* __restrict__ dsums, int ld_ds,
                                 float* __restrict__ da2c, float* __restrict__ dc_prev, int B, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx - b * D;
  const float* s = sums + (size_t)b * lds;
  float* ds = dsums + (size_t)b * ld_ds;
  const float ig = sigmoidf_acc(s[d]), fg = sigmoidf_acc(s[D + d]), og = sigmoidf_acc(s[2 * D + d]);
  const float g1 = s[3 * D + d] + a2c[(size_t)b * 2 * D + d];
  const float g2 = s[4 * D + d] + a2c[(size_t)b * 2 * D + D + d];
  const bool first = g1 >= g2;
  const float gg = first ? g1 : g2;
  const float tc = tanhf(c[idx]);
  const float gh = (dh_a ? dh_a[idx] : 0.f) + (dh_b ? dh_b[idx] : 0.f);
  const float dct = (dc ? dc[idx] : 0.f) + gh * og * (1.f - tc * tc);
  ds[d] = dct * gg * ig * (1.f - ig);
  ds[D + d] = dct * (c_prev ? c_prev[idx] : 0.f) * fg * (1.f - fg);
  ds[2 * D + d] = gh * tc * og * (1.f - og);
  const float dg = dct * ig;
  ds[3 * D + d] = first ? dg : 0.f;
  ds[4 * D + d] = first ? 0.f : dg;
  da2c[(size_t)b * 2 * D + d] = first ? dg : 0.f;
  da2c[(size_t)b * 2 * D + D + d] = first ? 0.f : dg;
  dc_prev[idx] = dct * fg;
}

// ------------------------------------------------------------------------------- log-softmax + NLL
__global__ void __launch_bounds__(256)
lsm_nll_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                   const float* __restrict__ mask, float* __restrict__ logp, float* __restrict__ nll, int V) {
  const int r = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const float* row = logits + (size_t)r * V;
  __shared__ float s_red[8];
  float m = -INFINITY;
  for (int v = t; v < V; v += 256) m = fmaxf(m, __ldg(row + v));
  m = warp_max(m);
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float l = 0.f;
  for (int v = t; v < V; v += 256) l += expf(__ldg(row + v) - m);
  l = warp_sum(l);
  if (lane == 0) s_red[wid] = l;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += s_red[w];
  const float lse = m + logf(tot);
  if (logp)
    for (int v = t; v < V; v += 256) logp[(size_t)r * V + v] = __ldg(row + v) - lse;
  if (t == 0 && nll) {
    const int64_t tg = target[r];
    const float mk = mask ? mask[r] : 1.f;
    nll[r] = (tg >= 0 && tg < V) ? -(__ldg(row + tg) - lse) * mk : 0.f;
  }
}

__global__ void __launch_bounds__(256)
lsm_nll_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                   const float* __restrict__ mask, const float* __restrict__ gscale, float* __restrict__ dlogits,
                   int V) {
  const int r = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const float* row = logits + (size_t)r * V;
  __shared__ float s_red[8];
  float m = -INFINITY;
  for (int v = t; v < V; v += 256) m = fmaxf(m, __ldg(row + v));
  m = warp_max(m);
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float l = 0.f;
  for (int v = t; v < V; v += 256) l += expf(__ldg(row + v) - m);
  l = warp_sum(l);
  if (lane == 0) s_red[wid] = l;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += s_red[w];
  const float inv = 1.f / tot;
  const float gs = __ldg(gscale) * (mask ? mask[r] : 1.f);
  const int64_t tg = target[r];
  for (int v = t; v < V; v += 256) {
    const float sm = expf(__ldg(row + v) - m) * inv;
    dlogits[(size_t)r * V + v] = gs * (sm - ((int64_t)v == tg ? 1.f : 0.f));
  }
}

// ------------------------------------------------------------------------------- caption feature prep
// adaptive bin [floor(i*n/S), ceil((i+1)*n/S))
__device__ __forceinline__ int bin_lo(int i, int n, int S) { return (i * n) / S; }
__device__ __forceinline__ int bin_hi(int i, int n, int S) { return ((i + 1) * n + S - 1) / S; }

template <int CH>
__global__ void __launch_bounds__(256)
caption_feats_fwd_kernel(const float* __restrict__ feats, float* __restrict__ fc, float* __restrict__ att, int C,
                         int H, int W, int S, int ldc, int col_off) {
  extern __shared__ __align__(16) float xs[];   // [CH][HW+1]
  const int HW = H * W, ld = HW + 1;
  const int b = blockIdx.y, c0 = blockIdx.x * CH, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  for (int idx = t; idx < CH * HW; idx += 256) {
    const int c = idx / HW, p = idx - c * HW;
    xs[c * ld + p] = (c0 + c < C) ? __ldg(feats + ((size_t)b * C + c0 + c) * HW + p) : 0.f;
  }
  __syncthreads();
  // fc: reference takes mean over W then mean over H (:428) -- same grouping here
  for (int c = wid; c < CH; c += 8) {
    float tot = 0.f;
    for (int y = lane; y < H; y += 32) {
      float rs = 0.f;
      for (int x = 0; x < W; ++x) rs += xs[c * ld + y * W + x];
      tot += rs / (float)W;
    }
    tot = warp_sum(tot);
    if (lane == 0 && c0 + c < C) fc[(size_t)b * ldc + col_off + c0 + c] = tot / (float)H;
  }
  // att: thread = (bin, channel) with channel fastest => coalesced NHWC stores
  for (int idx = t; idx < S * S * CH; idx += 256) {
    const int c = idx % CH, bin = idx / CH;
    const int i = bin / S, j = bin - i * S;
    const int y0 = bin_lo(i, H, S), y1 = bin_hi(i, H, S), x0 = bin_lo(j, W, S), x1 = bin_hi(j, W, S);
    float s = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) s += xs[c * ld + y * W + x];
    if (c0 + c < C) att[((size_t)b * S * S + bin) * ldc + col_off + c0 + c] = s / (float)((y1 - y0) * (x1 - x0));
  }
}

template <int CH>
__global__ void __launch_bounds__(256)
caption_feats_bwd_kernel(const float* __restrict__ dfc, const float* __restrict__ datt, float* __restrict__ dfeats,
                         int C, int H, int W, int S, int ldc, int col_off) {
  extern __shared__ __align__(16) float ds[];   // [S*S][CH+1]
  const int HW = H * W, ld = CH + 1;
  const int b = blockIdx.y, c0 = blockIdx.x * CH, t = threadIdx.x;
  for (int idx = t; idx < S * S * CH; idx += 256) {
    const int c = idx % CH, bin = idx / CH;
    ds[bin * ld + c] = (datt && c0 + c < C) ? __ldg(datt + ((size_t)b * S * S + bin) * ldc + col_off + c0 + c) : 0.f;
  }
  __syncthreads();
  for (int idx = t; idx < CH * HW; idx += 256) {
    const int c = idx / HW, p = idx - c * HW;
    if (c0 + c >= C) continue;
    const int y = p / W, x = p - y * W;
    float g = dfc ? __ldg(dfc + (size_t)b * ldc + col_off + c0 + c) / (float)HW : 0.f;
    // bins containing y: i in [ilo, ihi]
    int ilo = (y * S) / H;
    while (ilo > 0 && bin_hi(ilo - 1, H, S) > y) --ilo;
    int jlo = (x * S) / W;
    while (jlo > 0 && bin_hi(jlo - 1, W, S) > x) --jlo;
    for (int i = ilo; i < S && bin_lo(i, H, S) <= y; ++i) {
      const int hy = bin_hi(i, H, S) - bin_lo(i, H, S);
      if (bin_hi(i, H, S) <= y) continue;
      for (int j = jlo; j < S && bin_lo(j, W, S) <= x; ++j) {
        if (bin_hi(j, W, S) <= x) continue;
        const int wx = bin_hi(j, W, S) - bin_lo(j, W, S);
        g += ds[(i * S + j) * ld + c] / (float)(hy * wx);
      }
    }
    dfeats[((size_t)b * C + c0 + c) * HW + p] = g;
  }
}

int launch_cluster(const void* kern, dim3 grid, size_t smem, cudaStream_t st, void** args) {
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kAttThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];      // cluster dims are compiled in (__cluster_dims__)
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  L2S_CUDA_OK(cudaLaunchKernelExC(&cfg, kern, args));
  count_launch();
  return L2S_OK;
}

int check_att(int B, int A, int D, int Dh) {
  L2S_REQUIRE(B > 0 && A > 0 && D > 0 && Dh > 0, L2S_ERR_SHAPE, "att_step: bad shape");
  L2S_REQUIRE(D % 4 == 0 && Dh % 4 == 0 && D <= 4096 && Dh <= 4096, L2S_ERR_SHAPE,
              "att_step: D and Dh must be multiples of 4 and <= 4096 (got %d, %d)", D, Dh);
  L2S_REQUIRE((A + kCluster - 1) / kCluster <= kMaxLoc, L2S_ERR_SHAPE, "att_step: A=%d too large (max %d)", A,
              kMaxLoc * kCluster);
  return L2S_OK;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

// ---- internal launchers (also used by the decode loop in decode.cu; declared in common.cuh) ----
namespace l2s {

int launch_att_step_fwd(const float* att_h, int ldh, const float* att_feats, const float* p_att, const float* alpha_w,
                        const float* alpha_b, float* weight, float* att_res, int B, int A, int D, int Dh,
                        cudaStream_t st) {
  L2S_REQUIRE(att_h && att_feats && p_att && alpha_w && alpha_b && weight && att_res, L2S_ERR_ARG, "att_step_fwd: null pointer");
  int rc = check_att(B, A, D, Dh);
  if (rc) return rc;
  L2S_REQUIRE(ldh >= Dh, L2S_ERR_SHAPE, "att_step_fwd: att_h row stride %d < Dh %d", ldh, Dh);
  L2S_REQUIRE(aligned16(att_feats) && aligned16(p_att), L2S_ERR_ALIGN, "att_step_fwd: att_feats / p_att must be 16-byte aligned");
  AttGeom g{B, A, D, Dh, ldh, Dh};
  void* args[] = {&att_h, &att_feats, &p_att, &alpha_w, &alpha_b, &weight, &att_res, &g};
  return launch_cluster((const void*)att_step_fwd_kernel, dim3(B * kCluster), (size_t)(2 * Dh + D) * 4, st, args);
}

int launch_att_step_bwd(const float* datt_res, const float* att_h, int ldh, const float* att_feats, const float* p_att,
                        const float* alpha_w, const float* weight, float* datt_h, int ld_dah, float* de, float* dp_att,
                        float* datt_feats, float* dalpha_w, int B, int A, int D, int Dh, cudaStream_t st) {
  L2S_REQUIRE(datt_res && att_h && att_feats && p_att && alpha_w && weight && datt_h && de, L2S_ERR_ARG,
              "att_step_bwd: null pointer");
  int rc = check_att(B, A, D, Dh);
  if (rc) return rc;
  L2S_REQUIRE(ldh >= Dh && ld_dah >= Dh, L2S_ERR_SHAPE, "att_step_bwd: row strides must be >= Dh");
  L2S_REQUIRE(aligned16(att_feats) && aligned16(p_att) && aligned16(dp_att) && aligned16(datt_feats), L2S_ERR_ALIGN,
              "att_step_bwd: feature pointers must be 16-byte aligned");
  AttGeom g{B, A, D, Dh, ldh, ld_dah};
  void* args[] = {&datt_res, &att_h, &att_feats, &p_att, &alpha_w, &weight, &datt_h, &de, &dp_att, &datt_feats,
                  &dalpha_w, &g};
  return launch_cluster((const void*)att_step_bwd_kernel, dim3(B * kCluster), (size_t)(3 * Dh + D) * 4, st, args);
}

int launch_gates_fwd(const float* sums, int lds, const float* a2c_out, const float* c_prev, float* h, float* c, int B,
                     int D, cudaStream_t st) {
  L2S_REQUIRE(sums && a2c_out && h && c, L2S_ERR_ARG, "gates_fwd: null pointer");
  L2S_REQUIRE(B > 0 && D > 0 && lds >= 5 * D, L2S_ERR_SHAPE, "gates_fwd: bad shape");
  L2S_CUDA_OK(launch_chain(gates_fwd_kernel, dim3((B * D + 255) / 256), dim3(256), 0, st, sums, lds, a2c_out, c_prev, h, c, B, D));
  L2S_LAUNCH_OK("gates_fwd_kernel");
  count_launch();
  return L2S_OK;
}

int launch_gates_bwd(const float* sums, int lds, const float* a2c_out, const float* c_prev, const float* c,
                     const float* dh_a, const float* dh_b, const float* dc, float* dsums, int ld_ds, float* da2c,
                     float* dc_prev, int B, int D, cudaStream_t st) {
  L2S_REQUIRE(sums && a2c_out && c && (dh_a || dh_b) && dsums && da2c && dc_prev, L2S_ERR_ARG, "gates_bwd: null pointer");
  L2S_REQUIRE(B > 0 && D > 0 && lds >= 5 * D && ld_ds >= 5 * D, L2S_ERR_SHAPE, "gates_bwd: bad shape");
  L2S_CUDA_OK(launch_chain(gates_bwd_kernel, dim3((B * D + 255) / 256), dim3(256), 0, st, sums, lds, a2c_out, c_prev, c, dh_a,
                           dh_b, dc, dsums, ld_ds, da2c, dc_prev, B, D));
  L2S_LAUNCH_OK("gates_bwd_kernel");
  count_launch();
  return L2S_OK;
}

}  // namespace l2s

extern "C" int l2s_att_step_fwd(const float* att_h, const float* att_feats, const float* p_att, const float* alpha_w,
                                const float* alpha_b, float* weight, float* att_res, int B, int A, int D, int Dh,
                                l2s_stream_t stream) {
  return launch_att_step_fwd(att_h, Dh, att_feats, p_att, alpha_w, alpha_b, weight, att_res, B, A, D, Dh,
                             (cudaStream_t)stream);
}

extern "C" int l2s_att_step_bwd(const float* datt_res, const float* att_h, const float* att_feats, const float* p_att,
                                const float* alpha_w, const float* weight, float* datt_h, float* de, float* dp_att,
                                float* datt_feats, float* dalpha_w, int B, int A, int D, int Dh, l2s_stream_t stream) {
  return launch_att_step_bwd(datt_res, att_h, Dh, att_feats, p_att, alpha_w, weight, datt_h, Dh, de, dp_att,
                             datt_feats, dalpha_w, B, A, D, Dh, (cudaStream_t)stream);
}

extern "C" int l2s_att2in2_gates_fwd(const float* sums, const float* a2c_out, const float* c_prev, float* h, float* c,
                                     int B, int D, l2s_stream_t stream) {
  L2S_REQUIRE(c_prev, L2S_ERR_ARG, "gates_fwd: null pointer");
  return launch_gates_fwd(sums, 5 * D, a2c_out, c_prev, h, c, B, D, (cudaStream_t)stream);
}

extern "C" int l2s_att2in2_gates_bwd(const float* sums, const float* a2c_out, const float* c_prev, const float* c,
                                     const float* dh, const float* dc, float* dsums, float* da2c, float* dc_prev,
                                     int B, int D, l2s_stream_t stream) {
  L2S_REQUIRE(c_prev && dh, L2S_ERR_ARG, "gates_bwd: null pointer");
  return launch_gates_bwd(sums, 5 * D, a2c_out, c_prev, c, dh, nullptr, dc, dsums, 5 * D, da2c, dc_prev, B, D,
                          (cudaStream_t)stream);
}

extern "C" int l2s_logsoftmax_nll_fwd(const float* logits, const int64_t* target, const float* mask, float* logp,
                                      float* nll, int R, int V, l2s_stream_t stream) {
  L2S_REQUIRE(logits && (logp || nll) && (!nll || target), L2S_ERR_ARG, "logsoftmax_nll_fwd: null pointer");
  L2S_REQUIRE(R >= 0 && V > 0, L2S_ERR_SHAPE, "logsoftmax_nll_fwd: bad shape");
  if (R == 0) return L2S_OK;
  lsm_nll_fwd_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(logits, target, mask, logp, nll, V);
  L2S_LAUNCH_OK("lsm_nll_fwd_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_logsoftmax_nll_bwd(const float* logits, const int64_t* target, const float* mask,
                                      const float* gscale, float* dlogits, int R, int V, l2s_stream_t stream) {
  L2S_REQUIRE(logits && target && gscale && dlogits, L2S_ERR_ARG, "logsoftmax_nll_bwd: null pointer");
  L2S_REQUIRE(R >= 0 && V > 0, L2S_ERR_SHAPE, "logsoftmax_nll_bwd: bad shape");
  if (R == 0) return L2S_OK;
  lsm_nll_bwd_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(logits, target, mask, gscale, dlogits, V);
  L2S_LAUNCH_OK("lsm_nll_bwd_kernel");
  count_launch();
  return L2S_OK;
}

template <int CH>
static int caption_fwd_launch(const float* feats, float* fc, float* att, int B, int C, int H, int W, int S, int ldc,
                              int col_off, cudaStream_t st) {
  const size_t smem = (size_t)CH * (H * W + 1) * 4;
  auto kern = caption_feats_fwd_kernel<CH>;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((C + CH - 1) / CH, B), 256, smem, st>>>(feats, fc, att, C, H, W, S, ldc, col_off);
  L2S_LAUNCH_OK("caption_feats_fwd_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_caption_feats_fwd(const float* feats, float* fc, float* att, int B, int C, int H, int W, int S,
                                     int ldc, int col_off, l2s_stream_t stream) {
  L2S_REQUIRE(feats && fc && att, L2S_ERR_ARG, "caption_feats_fwd: null pointer");
  L2S_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && S > 0 && ldc >= col_off + C, L2S_ERR_SHAPE, "caption_feats_fwd: bad shape");
  const size_t cap = (size_t)max_smem_optin();
  if ((size_t)32 * (H * W + 1) * 4 <= cap) return caption_fwd_launch<32>(feats, fc, att, B, C, H, W, S, ldc, col_off, (cudaStream_t)stream);
  if ((size_t)8 * (H * W + 1) * 4 <= cap) return caption_fwd_launch<8>(feats, fc, att, B, C, H, W, S, ldc, col_off, (cudaStream_t)stream);
  return fail(L2S_ERR_SHAPE, "caption_feats_fwd: H*W=%d too large", H * W);
}

extern "C" int l2s_caption_feats_bwd(const float* dfc, const float* datt, float* dfeats, int B, int C, int H, int W,
                                     int S, int ldc, int col_off, l2s_stream_t stream) {
  L2S_REQUIRE(dfeats && (dfc || datt), L2S_ERR_ARG, "caption_feats_bwd: null pointer");
  L2S_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && S > 0 && ldc >= col_off + C, L2S_ERR_SHAPE, "caption_feats_bwd: bad shape");
  constexpr int CH = 32;
  const size_t smem = (size_t)S * S * (CH + 1) * 4;
  L2S_REQUIRE(smem <= (size_t)max_smem_optin(), L2S_ERR_SHAPE, "caption_feats_bwd: S=%d too large", S);
  auto kern = caption_feats_bwd_kernel<CH>;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((C + CH - 1) / CH, B), 256, smem, (cudaStream_t)stream>>>(dfc, datt, dfeats, C, H, W, S, ldc, col_off);
  L2S_LAUNCH_OK("caption_feats_bwd_kernel");
  count_launch();
  return L2S_OK;
}
