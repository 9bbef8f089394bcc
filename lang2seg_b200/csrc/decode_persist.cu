// att2in2 decode recurrence (AttModel.py:75-99 around Att2in2Core :446-466 and Attention :406-423) as ONE persistent,
// weight-stationary kernel per direction.
//
// decode.cu runs a token as four (fwd) / four (bwd) dependent launches of 5-17 us each -- ~9 launches per token, 23 % of
// the cfg-2 step for 0.14 ms worth of traffic.  Here the whole T-step loop is one cooperative launch of 128 CTAs:
//
//   * every CTA keeps its slice of [W_h2att ; W_h2h] and W_a2c (64 KB fp32) in shared memory for all T steps; the only
//     per-step operand traffic of the skinny GEMMs is the (B x 512) state row block (h_{t-1}, att_res_t, ...), staged
//     in shared memory once per phase;
//   * a step is a sequence of phases separated by GRID barriers (one global counter, split arrive / wait so that
//     independent work -- the h2h part of the pre-activations -- runs while the barrier completes) instead of kernel
//     boundaries;
//   * the attention phase keeps the design of att_step_{fwd,bwd}_kernel: a thread-block cluster shares one sample,
//     flash-style partial (max, sum, weighted accumulator) per CTA, merged through distributed shared memory;
//   * operand row blocks reach shared memory through the TMA engine (1-D bulk copies on an mbarrier), not LDG loops;
//   * the attention phase processes ALL samples of a cluster as one batch (every load of the batch in flight
//     together, one pair of cluster barriers per step) -- a per-sample loop was a chain of exposed L2 latencies;
//   * skinny GEMM tiles: lanes split K (one float4 per k-step), R rows x NC columns register blocked against shared
//     memory (4 (R + NC) LDS.128 per 16 R NC FMA), one 31-shuffle transpose-reduce per 32 outputs; exact fp32 FFMA
//     with a fixed summation order (deterministic; same arithmetic class as linear_small_kernel).
//
// forward  (CTA j owns hidden units 4j..4j+3 = columns of att_h, the five gate pre-activations and a2c):
//     A1  att_h_t  += h_{t-1} W_h2att^T             -> cat_all[t][:, :512]           | grid barrier 1 (arrive)
//     A2  sums_t    = h_{t-1} W_h2h^T (own columns; kept in shared memory)            | grid barrier 1 (wait)
//     B   pi_t, att_res_t = attention(att_h_t, p_att, att)   (cluster per sample)     | grid barrier 2
//     C   a2c_t = att_res_t W_a2c^T + b ; gates -> h_t, c_t                           | grid barrier 3
// backward (cluster c of 4 CTAs owns 16 output columns, rank r owns a quarter of K; partial sums meet in DSMEM):
//     S1  gates backward (elementwise over the CTA's share of (b, unit))              | grid barrier 1
//     S2  datt_res_t = da2c_t W_a2c              ;  S4a  dh_{t-1} partial over the dsums part of K
//                                                                                    | grid barrier 2
//     S3  attention backward (light: datt_h_t, de_t)  (cluster per sample)            | grid barrier 3
//     S4b dh_{t-1} += datt_h_t W_h2att  -> cluster reduce -> dh carry                 | grid barrier 4
// Restrictions: rnn_size == att_hid_size == 512, B <= 64 (backward: 48), A <= 1024; anything else takes the launch
// chain of decode.cu.
#include <cooperative_groups.h>

#include "persist.cuh"

namespace l2s {
namespace {

constexpr int NIMAX = 2;          // samples a cluster handles per step: ceil(PMAXB / clusters) with >= 32 clusters
constexpr int LC = 6 * PD;        // [att_h | 5 gate pre-activations]

struct DecFwdArgs {
  float* cat_all;          // (T,B,LC) in: [b_h2att | i2h(x_t) + b_i2h + b_h2h]  out: [att_h_t | sums_t]
  const float* att;        // (B,A,D)
  const float* p_att;      // (B,A,D)
  const float* w_cat;      // (LC, D) = [W_h2att ; W_h2h]
  const float* w_a2c;      // (2D, D)
  const float* b_a2c;      // (2D)
  const float* alpha_w;    // (D)
  const float* alpha_b;    // (1)
  float* h_all;            // (T,B,D)
  float* c_all;            // (T,B,D)
  float* a2c_all;          // (T,B,2D)
  float* pi_all;           // (T,B,A)
  float* res_all;          // (T,B,D)
  unsigned* bar;
  unsigned long long* prof;   // diagnostics or null
  int T, B, A;
};

// attention scratch (floats), aliased onto the operand staging area (never live at the same time)
struct AttScratch {
  static constexpr int AH = 0;                          // [NIMAX][PD]       att_h rows
  static constexpr int DO = AH + NIMAX * PD;            // [NIMAX][PD]       (bwd) datt_res rows
  static constexpr int ACC = DO + NIMAX * PD;           // [NIMAX][4][PD]    per-phase partial column sums
  static constexpr int ACCF = ACC + NIMAX * 4 * PD;     // [NIMAX][PD]       their sum (read remotely)
  static constexpr int E = ACCF + NIMAX * PD;           // [NIMAX][PMAXLOC]  scores / exp / de
  static constexpr int W = E + NIMAX * PMAXLOC;         // [NIMAX][PMAXLOC]  (bwd) pi
  static constexpr int ML = W + NIMAX * PMAXLOC;        // [NIMAX][2]        partial (max, sum) | (bwd) partial S
  static constexpr int TOTAL = ML + NIMAX * 2 + 4;
};

// Attention.forward (AttModel.py:411-421) for the `ni` samples b = b0 + i * bstride of this cluster, as one batch
template <int CS>
__device__ __forceinline__ void attention_fwd_batch(cg::cluster_group& cluster, int rank, int b0, int bstride, int ni,
                                                    const float* __restrict__ cat_t, const float* __restrict__ att,
                                                    const float* __restrict__ p_att, float alpha_b,
                                                    float* __restrict__ pi_t, float* __restrict__ res_t, int A,
                                                    float* sc_, const float* s_aw) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  float* s_ah = sc_ + AttScratch::AH;
  float* s_acc = sc_ + AttScratch::ACC;
  float* s_accf = sc_ + AttScratch::ACCF;
  float* s_e = sc_ + AttScratch::E;
  float* s_ml = sc_ + AttScratch::ML;
  const int per = (A + CS - 1) / CS;
  const int a0 = min(A, rank * per), a1 = min(A, a0 + per);
  const int na = a1 - a0;
  for (int i = t; i < ni * PQ; i += PT) {
    const int it = i / PQ, q = i - it * PQ;
    reinterpret_cast<float4*>(s_ah)[i] = __ldcg(reinterpret_cast<const float4*>(cat_t + (size_t)(b0 + it * bstride) * LC) + q);
  }
  __syncthreads();
  // ---- scores: one warp per (sample, location), two pairs per batch: 8 row loads per lane, 128 per SM in flight
  const int npairs = ni * na;
  for (int p0 = wid; p0 < npairs; p0 += 2 * (PT / 32)) {
    float4 pr[2][4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = p0 + k * (PT / 32);
      if (p < npairs) {
        const int it = p / na, a = p - it * na;
        const float4* row = reinterpret_cast<const float4*>(p_att + ((size_t)(b0 + it * bstride) * A + a0 + a) * PD);
#pragma unroll
        for (int j = 0; j < 4; ++j) pr[k][j] = __ldg(row + lane + 32 * j);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = p0 + k * (PT / 32);
      if (p < npairs) {                          // warp uniform
        const int it = p / na, a = p - it * na;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = lane + 32 * j;
          const float4 h = reinterpret_cast<const float4*>(s_ah + it * PD)[q];
          const float4 w = reinterpret_cast<const float4*>(s_aw)[q];
          s = fmaf(w.x, tanhf_fast_acc(pr[k][j].x + h.x), s);
          s = fmaf(w.y, tanhf_fast_acc(pr[k][j].y + h.y), s);
          s = fmaf(w.z, tanhf_fast_acc(pr[k][j].z + h.z), s);
          s = fmaf(w.w, tanhf_fast_acc(pr[k][j].w + h.w), s);
        }
        s = warp_sum(s);
        if (lane == 0) s_e[it * PMAXLOC + a] = s + alpha_b;
      }
    }
  }
  __syncthreads();
  // ---- partial softmax statistics of this slice: warp `it` takes sample `it`
  if (wid < ni) {
    float* e = s_e + wid * PMAXLOC;
    float m = -INFINITY;
    for (int a = lane; a < na; a += 32) m = fmaxf(m, e[a]);
    m = warp_max(m);
    float l = 0.f;
    for (int a = lane; a < na; a += 32) {
      const float ex = expf(e[a] - m);
      e[a] = ex;
      l += ex;
    }
    l = warp_sum(l);
    if (lane == 0) {
      s_ml[wid * 2] = m;
      s_ml[wid * 2 + 1] = l;
    }
  }
  __syncthreads();
  // ---- partial weighted sums: thread = (float4 column group q, location phase ph), 6 row loads in flight
  {
    const int q = t & (PQ - 1), ph = t / PQ;        // PT / PQ == 4 phases
    for (int it = 0; it < ni; ++it) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* col = reinterpret_cast<const float4*>(att + ((size_t)(b0 + it * bstride) * A + a0) * PD) + q;
      const float* e = s_e + it * PMAXLOC;
      for (int ab = ph; ab < na; ab += 6 * (PT / PQ)) {
        float4 v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int a = ab + k * (PT / PQ);
          if (a < na) v[k] = __ldg(col + (size_t)a * PQ);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int a = ab + k * (PT / PQ);
          if (a < na) {
            const float w = e[a];
            acc.x = fmaf(w, v[k].x, acc.x); acc.y = fmaf(w, v[k].y, acc.y);
            acc.z = fmaf(w, v[k].z, acc.z); acc.w = fmaf(w, v[k].w, acc.w);
          }
        }
      }
      reinterpret_cast<float4*>(s_acc + (it * 4 + ph) * PD)[q] = acc;
    }
  }
  __syncthreads();
  for (int i = t; i < ni * PQ; i += PT) {
    const int it = i / PQ, q = i - it * PQ;
    const float4* src = reinterpret_cast<const float4*>(s_acc + it * 4 * PD) + q;
    const float4 x0 = src[0], x1 = src[PQ], x2 = src[2 * PQ], x3 = src[3 * PQ];
    reinterpret_cast<float4*>(s_accf)[i] = make_float4((x0.x + x1.x) + (x2.x + x3.x), (x0.y + x1.y) + (x2.y + x3.y),
                                                       (x0.z + x1.z) + (x2.z + x3.z), (x0.w + x1.w) + (x2.w + x3.w));
  }
  cluster.sync();
  // ---- merge the CS partial states; this CTA finalises columns [rank*DPER, ...) of att_res and its slice of pi
  constexpr int DPER = PD / CS;
  for (int it = 0; it < ni; ++it) {
    float ms[CS], ls[CS], M = -INFINITY;
#pragma unroll
    for (int r = 0; r < CS; ++r) {
      const float* rm = cluster.map_shared_rank(s_ml, r);
      ms[r] = rm[it * 2];
      ls[r] = rm[it * 2 + 1];
      M = fmaxf(M, ms[r]);
    }
    float L = 0.f, scl[CS], mine = 0.f;
#pragma unroll
    for (int r = 0; r < CS; ++r) {
      scl[r] = (ls[r] > 0.f) ? expf(ms[r] - M) : 0.f;
      L = fmaf(scl[r], ls[r], L);
      if (r == rank) mine = scl[r];
    }
    const float invL = 1.f / L;
    const int b = b0 + it * bstride;
    for (int d = rank * DPER + t; d < (rank + 1) * DPER; d += PT) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < CS; ++r) v = fmaf(scl[r], cluster.map_shared_rank(s_accf, r)[it * PD + d], v);
      res_t[(size_t)b * PD + d] = v * invL;
    }
    const float myscale = mine * invL;
    for (int a = t; a < na; a += PT) pi_t[(size_t)b * A + a0 + a] = s_e[it * PMAXLOC + a] * myscale;
  }
  cluster.sync();   // nobody moves on while its shared memory may still be read remotely
}

// shared-memory plan of the forward kernel (floats)
struct DecSmem {
  static constexpr int W = 0;                        // [32][PD] stationary weight rows
  static constexpr int A_OFF = W + 32 * PD;          // [PMAXB][PD] staged operand rows | attention scratch
  static constexpr int SUMS = A_OFF + PMAXB * PD;    // [PMAXB][20]
  static constexpr int A2C = SUMS + PMAXB * 20;      // [PMAXB][8]
  static constexpr int C = A2C + PMAXB * 8;          // [PMAXB][4]
  static constexpr int AW = C + PMAXB * 4;           // [PD]
  static constexpr int BAR = AW + PD;                // mbarrier (2 floats)
  static constexpr int TOTAL = BAR + 4;
};
static_assert(AttScratch::TOTAL <= PMAXB * PD, "attention scratch must fit the staging area");

template <int CS>
__global__ void __launch_bounds__(PT, 1) decode_fwd_persist_kernel(const DecFwdArgs p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int j = blockIdx.x;                // hidden units 4j .. 4j+3
  const int cid = j / CS;
  constexpr int NCL = PG / CS;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int B = p.B, A = p.A, T = p.T;

  extern __shared__ __align__(16) float smem[];
  float* sW = smem + DecSmem::W;                  // row 0-3 h2att | 4 + 5u + g: h2h gate g of unit u | 24 + 4h + u: a2c
  float* sA = smem + DecSmem::A_OFF;
  float* s_sums = smem + DecSmem::SUMS;           // h_{t-1} W_h2h^T for the CTA's units (col = 5u + g)
  float* s_a2c = smem + DecSmem::A2C;             // col = 4 half + u
  float* s_c = smem + DecSmem::C;                 // cell state of the CTA's units
  float* s_aw = smem + DecSmem::AW;
  Stager stg{reinterpret_cast<uint64_t*>(smem + DecSmem::BAR), 0u};
  stg.init();

  for (int i = t; i < 32 * PQ; i += PT) {
    const int r = i / PQ, q = i - r * PQ;
    const float* src;
    if (r < 4) src = p.w_cat + (size_t)(4 * j + r) * PD;
    else if (r < 24) {
      const int u = (r - 4) / 5, g = (r - 4) % 5;
      src = p.w_cat + (size_t)(PD + g * PD + 4 * j + u) * PD;
    } else {
      const int h = (r - 24) >> 2, u = (r - 24) & 3;
      src = p.w_a2c + (size_t)(h * PD + 4 * j + u) * PD;
    }
    reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(src) + q);
  }
  for (int i = t; i < PD; i += PT) s_aw[i] = __ldg(p.alpha_w + i);
  for (int i = t; i < PMAXB * 4; i += PT) s_c[i] = 0.f;
  for (int i = t; i < PMAXB * 20; i += PT) s_sums[i] = 0.f;
  const float alpha_b = __ldg(p.alpha_b);
  __syncthreads();

  GridBar gb{p.bar, 0u, (unsigned)PG};
  const PhaseProf prof{p.prof};
  const float4* sA4 = reinterpret_cast<const float4*>(sA);
  const float4* sW4 = reinterpret_cast<const float4*>(sW);
  const int ni = cid < B ? (B - cid + NCL - 1) / NCL : 0;     // samples of this cluster: cid, cid + NCL, ...

  for (int step = 0; step < T; ++step) {
    float* cat_t = p.cat_all + (size_t)step * B * LC;
    prof.mark(step, 0);
    if (step > 0) {
      stg.load_contig(sA, p.h_all + (size_t)(step - 1) * B * PD, (uint32_t)B * PD * 4u);
      stg.wait();
      // ---- A1: att_h columns 4j..4j+3 for rows b = wid + 16 r
      if (wid < B) {                        // warp uniform
        int arow[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) arow[r] = (wid + 16 * r < B) ? wid + 16 * r : 0;
        const float tot = gemv_tile<4, 4, 4>(sA4, PQ, arow, sW4, PQ, 0, lane);
        const int b = wid + 16 * (lane >> 2), c = lane & 3;
        if (lane < 16 && b < B) {
          float* dp = cat_t + (size_t)b * LC + 4 * j + c;
          *dp = __ldcg(dp) + tot;
        }
      }
      prof.mark(step, 1);
      gb.arrive();
      // ---- A2: the five gate columns of unit u = wid & 3 for rows b = rg + 4 i (runs while barrier 1 completes)
      {
        const int u = wid & 3, rg = wid >> 2;
        for (int i0 = 0; rg + 4 * i0 < B; i0 += 6) {
          int arow[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) arow[r] = (rg + 4 * (i0 + r) < B) ? rg + 4 * (i0 + r) : 0;
          const float tot = gemv_tile<6, 5, 4>(sA4, PQ, arow, sW4, PQ, 4 + 5 * u, lane);
          const int r = lane / 5, g = lane - 5 * r;
          const int b = rg + 4 * (i0 + r);
          if (lane < 30 && b < B) s_sums[b * 20 + 5 * u + g] = tot;
        }
      }
      prof.mark(step, 2);
      gb.wait();
    }
    prof.mark(step, 3);
    // ---- B: attention, all samples of this cluster in one batch
    if (ni > 0)
      attention_fwd_batch<CS>(cluster, rank, cid, NCL, ni, cat_t, p.att, p.p_att, alpha_b, p.pi_all + (size_t)step * B * A,
                              p.res_all + (size_t)step * B * PD, A, sA, s_aw);
    prof.mark(step, 4);
    gb.arrive();
    gb.wait();
    prof.mark(step, 5);
    // ---- C: a2c columns + gates
    stg.load_contig(sA, p.res_all + (size_t)step * B * PD, (uint32_t)B * PD * 4u);
    stg.wait();
    {
      const int half = wid & 1, rg = wid >> 1;
      if (rg < B) {
        int arow[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) arow[r] = (rg + 8 * r < B) ? rg + 8 * r : 0;
        const float tot = gemv_tile<8, 4, 4>(sA4, PQ, arow, sW4, PQ, 24 + 4 * half, lane);
        const int b = rg + 8 * (lane >> 2), u = lane & 3;
        if (b < B) s_a2c[b * 8 + 4 * half + u] = tot;
      }
    }
    __syncthreads();
    if (t < 4 * B) {
      const int b = t >> 2, u = t & 3, d = 4 * j + u;
      float* crow = cat_t + (size_t)b * LC + PD;            // sums part of the row
      float s[5];
#pragma unroll
      for (int g = 0; g < 5; ++g) {
        s[g] = crow[g * PD + d] + s_sums[b * 20 + 5 * u + g];
        crow[g * PD + d] = s[g];
      }
      const float a0 = s_a2c[b * 8 + u] + __ldg(p.b_a2c + d), a1 = s_a2c[b * 8 + 4 + u] + __ldg(p.b_a2c + PD + d);
      float* a2c_t = p.a2c_all + ((size_t)step * B + b) * 2 * PD;
      a2c_t[d] = a0;
      a2c_t[PD + d] = a1;
      const float ig = sigmoidf_acc(s[0]), fg = sigmoidf_acc(s[1]), og = sigmoidf_acc(s[2]);
      const float gg = fmaxf(s[3] + a0, s[4] + a1);
      const float cn = fmaf(fg, s_c[b * 4 + u], ig * gg);
      s_c[b * 4 + u] = cn;
      p.c_all[((size_t)step * B + b) * PD + d] = cn;
      p.h_all[((size_t)step * B + b) * PD + d] = og * tanhf(cn);
    }
    prof.mark(step, 6);
    if (step + 1 < T) {
      gb.arrive();
      gb.wait();
    }
    prof.mark(step, 7);
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int BCS = 4;            // cluster size of the backward kernel: 32 clusters x 16 output columns, rank = K quarter
constexpr int BMAXB = 48;         // samples (the staged (B x 640) operand block must fit next to the weights)
constexpr int BNC = PD / (PG / BCS);   // 16 output columns per cluster

struct DecBwdArgs {
  const float* dh_all;     // (T,B,D) upstream gradient on h_t
  const float* cat_all;    // (T,B,LC) [att_h | sums] kept by the forward
  const float* att;
  const float* p_att;
  const float* w_cat_t;    // (D, LC) = [W_h2att ; W_h2h]^T
  const float* w_a2c_t;    // (D, 2D) = W_a2c^T
  const float* alpha_w;
  const float* c_all;      // (T,B,D)
  const float* a2c_all;    // (T,B,2D)
  const float* pi_all;     // (T,B,A)
  float* dcat_all;         // (T,B,LC) [datt_h | dsums]
  float* da2c_all;         // (T,B,2D)
  float* dres_all;         // (T,B,D)
  float* de_all;           // (T,B,A)
  float* dh_carry;         // (B,D) scratch
  unsigned* bar;
  unsigned long long* prof;   // diagnostics or null
  int T, B, A;
};

// out[b][4 cg + c] (partial over this CTA's K slice) for rows b = rg + 4 i : warp = (column group cg, row group rg)
template <int KS>
__device__ __forceinline__ void partial_gemm16(const float4* sA4, int lda4, const float4* sW4, int B, float* s_part,
                                               bool accumulate, int lane, int wid) {
  const int cgp = wid & 3, rg = wid >> 2;
  for (int i0 = 0; rg + 4 * i0 < B; i0 += 8) {
    int arow[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) arow[r] = (rg + 4 * (i0 + r) < B) ? rg + 4 * (i0 + r) : 0;
    const float tot = gemv_tile<8, 4, KS>(sA4, lda4, arow, sW4, lda4, 4 * cgp, lane);
    const int b = rg + 4 * (i0 + (lane >> 2)), c = 4 * cgp + (lane & 3);
    if (b < B) s_part[b * BNC + c] = accumulate ? s_part[b * BNC + c] + tot : tot;
  }
}

// sum of the four ranks' partials for this rank's quarter of the cluster's 16 columns -> dst[b][col0 + ...]
__device__ __forceinline__ void cluster_reduce16(cg::cluster_group& cluster, int rank, const float* s_part, int B,
                                                 float* __restrict__ dst, int ld, int col0) {
  const int t = threadIdx.x;
  if (t < 4 * B) {
    const int b = t >> 2, c = 4 * rank + (t & 3);
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < BCS; ++q) v += cluster.map_shared_rank(s_part, q)[b * BNC + c];
    dst[(size_t)b * ld + col0 + c] = v;
  }
}

// light attention backward (datt_h -> dcat rows, de) for the `ni` samples of this cluster as one batch
__device__ __forceinline__ void attention_bwd_batch(cg::cluster_group& cluster, int rank, int b0, int bstride, int ni,
                                                    const float* __restrict__ dres_t, const float* __restrict__ cat_t,
                                                    const float* __restrict__ att, const float* __restrict__ p_att,
                                                    const float* __restrict__ pi_t, float* __restrict__ dcat_t,
                                                    float* __restrict__ de_t, int A, float* sc_, const float* s_aw) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  float* s_ah = sc_ + AttScratch::AH;
  float* s_do = sc_ + AttScratch::DO;
  float* s_dah = sc_ + AttScratch::ACC;
  float* s_dahf = sc_ + AttScratch::ACCF;
  float* s_dpi = sc_ + AttScratch::E;
  float* s_w = sc_ + AttScratch::W;
  float* s_ps = sc_ + AttScratch::ML;
  const int per = (A + BCS - 1) / BCS;
  const int a0 = min(A, rank * per), a1 = min(A, a0 + per);
  const int na = a1 - a0;
  for (int i = t; i < ni * PQ; i += PT) {
    const int it = i / PQ, q = i - it * PQ;
    const int b = b0 + it * bstride;
    reinterpret_cast<float4*>(s_ah)[i] = __ldg(reinterpret_cast<const float4*>(cat_t + (size_t)b * LC) + q);
    reinterpret_cast<float4*>(s_do)[i] = __ldcg(reinterpret_cast<const float4*>(dres_t + (size_t)b * PD) + q);
  }
  for (int i = t; i < ni * na; i += PT) {
    const int it = i / na, a = i - it * na;
    s_w[it * PMAXLOC + a] = __ldg(pi_t + (size_t)(b0 + it * bstride) * A + a0 + a);
  }
  __syncthreads();
  // ---- d pi_a = <datt_res, att_feats[a]>: warp per (sample, location), two pairs per batch
  const int npairs = ni * na;
  for (int p0 = wid; p0 < npairs; p0 += 2 * (PT / 32)) {
    float4 pr[2][4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = p0 + k * (PT / 32);
      if (p < npairs) {
        const int it = p / na, a = p - it * na;
        const float4* row = reinterpret_cast<const float4*>(att + ((size_t)(b0 + it * bstride) * A + a0 + a) * PD);
#pragma unroll
        for (int j = 0; j < 4; ++j) pr[k][j] = __ldg(row + lane + 32 * j);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = p0 + k * (PT / 32);
      if (p < npairs) {
        const int it = p / na, a = p - it * na;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) s = dot4(pr[k][j], reinterpret_cast<const float4*>(s_do + it * PD)[lane + 32 * j], s);
        s = warp_sum(s);
        if (lane == 0) s_dpi[it * PMAXLOC + a] = s;
      }
    }
  }
  __syncthreads();
  if (wid < ni) {
    float part = 0.f;
    for (int a = lane; a < na; a += 32) part = fmaf(s_w[wid * PMAXLOC + a], s_dpi[wid * PMAXLOC + a], part);
    part = warp_sum(part);
    if (lane == 0) s_ps[wid] = part;
  }
  cluster.sync();
  for (int i = t; i < ni * na; i += PT) {
    const int it = i / na, a = i - it * na;
    float S = 0.f;
#pragma unroll
    for (int r = 0; r < BCS; ++r) S += cluster.map_shared_rank(s_ps, r)[it];
    const float de = s_w[it * PMAXLOC + a] * (s_dpi[it * PMAXLOC + a] - S);
    s_dpi[it * PMAXLOC + a] = de;
    de_t[(size_t)(b0 + it * bstride) * A + a0 + a] = de;
  }
  __syncthreads();
  // ---- datt_h[d] = alpha_d sum_a de_a (1 - tanh^2(p_att[a,d] + att_h[d])): thread = (float4 column group, phase)
  {
    const int q = t & (PQ - 1), ph = t / PQ;
    const float4 w = reinterpret_cast<const float4*>(s_aw)[q];
    for (int it = 0; it < ni; ++it) {
      const float4 h = reinterpret_cast<const float4*>(s_ah + it * PD)[q];
      const float4* col = reinterpret_cast<const float4*>(p_att + ((size_t)(b0 + it * bstride) * A + a0) * PD) + q;
      const float* de = s_dpi + it * PMAXLOC;
      float4 dah = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ab = ph; ab < na; ab += 6 * (PT / PQ)) {
        float4 v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int a = ab + k * (PT / PQ);
          if (a < na) v[k] = __ldg(col + (size_t)a * PQ);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int a = ab + k * (PT / PQ);
          if (a < na) {
            const float d = de[a];
            const float tx = tanhf_fast_acc(v[k].x + h.x), ty = tanhf_fast_acc(v[k].y + h.y);
            const float tz = tanhf_fast_acc(v[k].z + h.z), tw = tanhf_fast_acc(v[k].w + h.w);
            dah.x += d * w.x * (1.f - tx * tx); dah.y += d * w.y * (1.f - ty * ty);
            dah.z += d * w.z * (1.f - tz * tz); dah.w += d * w.w * (1.f - tw * tw);
          }
        }
      }
      reinterpret_cast<float4*>(s_dah + (it * 4 + ph) * PD)[q] = dah;
    }
  }
  __syncthreads();
  for (int i = t; i < ni * PQ; i += PT) {
    const int it = i / PQ, q = i - it * PQ;
    const float4* src = reinterpret_cast<const float4*>(s_dah + it * 4 * PD) + q;
    const float4 x0 = src[0], x1 = src[PQ], x2 = src[2 * PQ], x3 = src[3 * PQ];
    reinterpret_cast<float4*>(s_dahf)[i] = make_float4((x0.x + x1.x) + (x2.x + x3.x), (x0.y + x1.y) + (x2.y + x3.y),
                                                       (x0.z + x1.z) + (x2.z + x3.z), (x0.w + x1.w) + (x2.w + x3.w));
  }
  cluster.sync();
  constexpr int DPER = PD / BCS;
  for (int i = t; i < ni * DPER; i += PT) {
    const int it = i / DPER, d = rank * DPER + (i - it * DPER);
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < BCS; ++r) v += cluster.map_shared_rank(s_dahf, r)[it * PD + d];
    dcat_t[(size_t)(b0 + it * bstride) * LC + d] = v;
  }
  cluster.sync();
}

struct DecBwdSmem {
  static constexpr int W2 = 0;                        // [16][256]  W_a2c^T rows of the cluster, this rank's K quarter
  static constexpr int W4A = W2 + BNC * 256;          // [16][640]  W_cat^T, dsums part of K
  static constexpr int W4B = W4A + BNC * 640;         // [16][128]  W_cat^T, datt_h part of K
  static constexpr int A_OFF = W4B + BNC * 128;       // [BMAXB][640] staged operand block | attention scratch
  static constexpr int PART2 = A_OFF + BMAXB * 640;   // [BMAXB][16]
  static constexpr int PART4 = PART2 + BMAXB * BNC;   // [BMAXB][16]
  static constexpr int DC = PART4 + BMAXB * BNC;      // [BMAXB][4]
  static constexpr int AW = DC + BMAXB * 4;           // [PD]
  static constexpr int BAR = AW + PD;
  static constexpr int TOTAL = BAR + 4;
};
static_assert(AttScratch::TOTAL <= BMAXB * 640, "attention scratch must fit the staging area");

__global__ void __launch_bounds__(PT, 1) decode_bwd_persist_kernel(const DecBwdArgs p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int j = blockIdx.x;
  const int cid = j / BCS;
  constexpr int NCL = PG / BCS;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int B = p.B, A = p.A, T = p.T;

  extern __shared__ __align__(16) float smem[];
  float* sW2 = smem + DecBwdSmem::W2;
  float* sW4a = smem + DecBwdSmem::W4A;
  float* sW4b = smem + DecBwdSmem::W4B;
  float* sA = smem + DecBwdSmem::A_OFF;
  float* s_part2 = smem + DecBwdSmem::PART2;
  float* s_part4 = smem + DecBwdSmem::PART4;
  float* s_dc = smem + DecBwdSmem::DC;            // dc carry of the CTA's units
  float* s_aw = smem + DecBwdSmem::AW;
  Stager stg{reinterpret_cast<uint64_t*>(smem + DecBwdSmem::BAR), 0u};
  stg.init();

  // ---- stationary weights: rows n = 16 cid + i of the transposed matrices, this rank's K slices
  for (int i = t; i < BNC * 64; i += PT) {
    const int r = i >> 6, q = i & 63;
    reinterpret_cast<float4*>(sW2)[i] =
        __ldg(reinterpret_cast<const float4*>(p.w_a2c_t + (size_t)(BNC * cid + r) * 2 * PD + 256 * rank) + q);
  }
  for (int i = t; i < BNC * 160; i += PT) {
    const int r = i / 160, q = i - r * 160;
    reinterpret_cast<float4*>(sW4a)[i] =
        __ldg(reinterpret_cast<const float4*>(p.w_cat_t + (size_t)(BNC * cid + r) * LC + PD + 640 * rank) + q);
  }
  for (int i = t; i < BNC * 32; i += PT) {
    const int r = i >> 5, q = i & 31;
    reinterpret_cast<float4*>(sW4b)[i] =
        __ldg(reinterpret_cast<const float4*>(p.w_cat_t + (size_t)(BNC * cid + r) * LC + 128 * rank) + q);
  }
  for (int i = t; i < PD; i += PT) s_aw[i] = __ldg(p.alpha_w + i);
  for (int i = t; i < BMAXB * 4; i += PT) s_dc[i] = 0.f;
  __syncthreads();

  GridBar gb{p.bar, 0u, (unsigned)PG};
  const PhaseProf prof{p.prof};
  const float4* sA4 = reinterpret_cast<const float4*>(sA);
  const int ni = cid < B ? (B - cid + NCL - 1) / NCL : 0;

  for (int step = T - 1; step >= 0; --step) {
    prof.mark(step, 0);
    const float* cat_t = p.cat_all + (size_t)step * B * LC;
    float* dcat_t = p.dcat_all + (size_t)step * B * LC;
    float* da2c_t = p.da2c_all + (size_t)step * B * 2 * PD;
    float* dres_t = p.dres_all + (size_t)step * B * PD;
    const bool last = (step == T - 1);
    // ---- S1: gates backward for (b, unit 4j + u)
    if (t < 4 * B) {
      const int b = t >> 2, u = t & 3, d = 4 * j + u;
      const float* srow = cat_t + (size_t)b * LC + PD;
      const float* a2c = p.a2c_all + ((size_t)step * B + b) * 2 * PD;
      const size_t idx = ((size_t)step * B + b) * PD + d;
      const float ig = sigmoidf_acc(__ldg(srow + d)), fg = sigmoidf_acc(__ldg(srow + PD + d));
      const float og = sigmoidf_acc(__ldg(srow + 2 * PD + d));
      const float g1 = __ldg(srow + 3 * PD + d) + __ldg(a2c + d), g2 = __ldg(srow + 4 * PD + d) + __ldg(a2c + PD + d);
      const bool first = g1 >= g2;
      const float gg = first ? g1 : g2;
      const float tc = tanhf(__ldg(p.c_all + idx));
      const float cp = step > 0 ? __ldg(p.c_all + idx - (size_t)B * PD) : 0.f;
      const float gh = __ldg(p.dh_all + idx) + (last ? 0.f : __ldcg(p.dh_carry + (size_t)b * PD + d));
      const float dct = (last ? 0.f : s_dc[b * 4 + u]) + gh * og * (1.f - tc * tc);
      float* ds = dcat_t + (size_t)b * LC + PD;
      ds[d] = dct * gg * ig * (1.f - ig);
      ds[PD + d] = dct * cp * fg * (1.f - fg);
      ds[2 * PD + d] = gh * tc * og * (1.f - og);
      const float dg = dct * ig;
      ds[3 * PD + d] = first ? dg : 0.f;
      ds[4 * PD + d] = first ? 0.f : dg;
      da2c_t[(size_t)b * 2 * PD + d] = first ? dg : 0.f;
      da2c_t[(size_t)b * 2 * PD + PD + d] = first ? 0.f : dg;
      s_dc[b * 4 + u] = dct * fg;
    }
    prof.mark(step, 1);
    gb.arrive();
    gb.wait();
    prof.mark(step, 2);
    // ---- S2: datt_res_t[:, 16 cid ..] = da2c_t W_a2c  (K quarter per rank, summed through DSMEM)
    stg.load_rows(sA, da2c_t + 256 * rank, B, 256 * 4, (size_t)2 * PD * 4);
    stg.wait();
    partial_gemm16<2>(sA4, 64, reinterpret_cast<const float4*>(sW2), B, s_part2, false, lane, wid);
    cluster.sync();
    cluster_reduce16(cluster, rank, s_part2, B, dres_t, PD, BNC * cid);
    prof.mark(step, 3);
    gb.arrive();
    // ---- S4a (runs while barrier 2 completes): dh_{t-1} partial over the dsums part of K
    if (step > 0) {
      __syncthreads();                   // every warp is done with the S2 operand block
      stg.load_rows(sA, dcat_t + PD + 640 * rank, B, 640 * 4, (size_t)LC * 4);
      stg.wait();
      partial_gemm16<5>(sA4, 160, reinterpret_cast<const float4*>(sW4a), B, s_part4, false, lane, wid);
    }
    prof.mark(step, 4);
    gb.wait();
    prof.mark(step, 5);
    // ---- S3: attention backward, all samples of this cluster in one batch
    if (ni > 0)
      attention_bwd_batch(cluster, rank, cid, NCL, ni, dres_t, cat_t, p.att, p.p_att, p.pi_all + (size_t)step * B * A,
                          dcat_t, p.de_all + (size_t)step * B * A, A, sA, s_aw);
    prof.mark(step, 6);
    if (step > 0) {
      gb.arrive();
      gb.wait();
      prof.mark(step, 7);
      // ---- S4b: + datt_h part of K, cluster sum -> dh carry
      stg.load_rows(sA, dcat_t + 128 * rank, B, 128 * 4, (size_t)LC * 4);
      stg.wait();
      partial_gemm16<1>(sA4, 32, reinterpret_cast<const float4*>(sW4b), B, s_part4, true, lane, wid);
      cluster.sync();
      cluster_reduce16(cluster, rank, s_part4, B, p.dh_carry, PD, BNC * cid);
      gb.arrive();
      gb.wait();
    }
  }
}

size_t dec_fwd_smem() { return (size_t)DecSmem::TOTAL * sizeof(float) + 64; }
size_t dec_bwd_smem() { return (size_t)DecBwdSmem::TOTAL * sizeof(float) + 64; }

}  // namespace

// 0 = not applicable (shape outside the persistent kernel's range, or the grid cannot be co-resident), else the cluster
// size to use.  Cached per (device): the occupancy query is not free.
int decode_persist_cluster(int B, int A, int D, int Dh) {
  static const bool off = env_flag("L2S_DECODE_CHAIN");       // diagnostics: force the launch chain of decode.cu
  if (off || D != PD || Dh != PD || B < 1 || B > PMAXB || A < 1 || A > 4 * PMAXLOC) return 0;
  static thread_local int cached_dev = -1, cached_cs = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    cached_dev = dev;
    cached_cs = 0;
    if (sm_count() >= PG && (size_t)max_smem_optin() >= dec_fwd_smem()) {
      // 8-CTA clusters would halve the per-CTA attention slice, but 16 of them (one CTA per SM) do not fit the GPCs of a
      // B200 (cudaOccupancyMaxActiveClusters < 16 on the pool's parts): 32 clusters of 4
      if (max_clusters(decode_fwd_persist_kernel<4>, 4, dec_fwd_smem()) >= PG / 4) cached_cs = 4;
    }
  }
  int cs = cached_cs;
  while (cs && (A + cs - 1) / cs > PMAXLOC) cs = 0;
  return cs;
}

int launch_decode_fwd_persist(int cs, float* cat_all, const float* att, const float* p_att, const float* w_cat,
                              const float* w_a2c, const float* b_a2c, const float* alpha_w, const float* alpha_b,
                              float* h_all, float* c_all, float* a2c_all, float* pi_all, float* res_all, int T, int B,
                              int A, unsigned* bar, cudaStream_t st) {
  L2S_REQUIRE(aligned16(cat_all) && aligned16(att) && aligned16(p_att) && aligned16(w_cat) && aligned16(w_a2c) &&
                  aligned16(h_all) && aligned16(res_all), L2S_ERR_ALIGN, "att2in2_decode_fwd: pointers must be 16-byte aligned");
  L2S_CUDA_OK(cudaMemsetAsync(bar, 0, 64, st));
  DecFwdArgs a{cat_all, att, p_att, w_cat, w_a2c, b_a2c, alpha_w, alpha_b, h_all, c_all, a2c_all, pi_all, res_all, bar,
               T <= 64 ? debug_buffer(0) : nullptr, T, B, A};
  static const bool coop = !env_flag("L2S_DECODE_NOCOOP");
  L2S_REQUIRE(cs == 4, L2S_ERR_ARG, "att2in2_decode_fwd: unsupported cluster size %d", cs);
  return launch_persistent(decode_fwd_persist_kernel<4>, 4, dec_fwd_smem(), st, a, coop);
}


// backward eligibility: forward eligibility + B <= 48 + the 4-CTA-cluster grid must be co-resident
bool decode_bwd_persist_ok(int B, int A, int D, int Dh) {
  if (!decode_persist_cluster(B, A, D, Dh) || B > BMAXB || (A + BCS - 1) / BCS > PMAXLOC) return false;
  static thread_local int cached_dev = -1;
  static thread_local bool cached_ok = false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev != cached_dev) {
    cached_dev = dev;
    cached_ok = (size_t)max_smem_optin() >= dec_bwd_smem() &&
                max_clusters(decode_bwd_persist_kernel, BCS, dec_bwd_smem()) >= PG / BCS;
  }
  return cached_ok;
}

int launch_decode_bwd_persist(const float* dh_all, const float* cat_all, const float* att, const float* p_att,
                              const float* w_cat_t, const float* w_a2c_t, const float* alpha_w, const float* c_all,
                              const float* a2c_all, const float* pi_all, float* dcat_all, float* da2c_all,
                              float* dres_all, float* de_all, float* dh_carry, int T, int B, int A, unsigned* bar,
                              cudaStream_t st) {
  L2S_REQUIRE(aligned16(dh_all) && aligned16(cat_all) && aligned16(att) && aligned16(p_att) && aligned16(w_cat_t) &&
                  aligned16(w_a2c_t) && aligned16(dcat_all) && aligned16(da2c_all) && aligned16(dres_all) &&
                  aligned16(dh_carry), L2S_ERR_ALIGN, "att2in2_decode_bwd: pointers must be 16-byte aligned");
  L2S_CUDA_OK(cudaMemsetAsync(bar, 0, 64, st));
  DecBwdArgs a{dh_all, cat_all, att, p_att, w_cat_t, w_a2c_t, alpha_w, c_all, a2c_all, pi_all, dcat_all, da2c_all,
               dres_all, de_all, dh_carry, bar, T <= 64 ? debug_buffer(1) : nullptr, T, B, A};
  static const bool coop = !env_flag("L2S_DECODE_NOCOOP");
  return launch_persistent(decode_bwd_persist_kernel, BCS, dec_bwd_smem(), st, a, coop);
}

}  // namespace l2s
