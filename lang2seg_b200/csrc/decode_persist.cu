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
// Restrictions: rnn_size == att_hid_size == 512, B <= 64, A <= 1024; anything else takes the launch chain of decode.cu.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace l2s {
namespace {

constexpr int PD = 512;           // rnn_size == att_hid_size
constexpr int PG = 128;           // CTAs of the persistent grid
constexpr int PT = 256;           // threads per CTA
constexpr int PMAXB = 64;         // samples
constexpr int PQ = PD / 4;        // float4 per 512-float row
constexpr int PMAXLOC = 256;      // attention locations per CTA slice

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// Grid barrier on one monotonically increasing counter (zeroed by the host before the launch).  arrive() and wait()
// are separate so that work which does not depend on the other CTAs can run in between.
struct GridBar {
  unsigned* ctr;
  unsigned epoch;
  __device__ __forceinline__ void arrive() {
    __syncthreads();                     // every thread's global writes of this phase are ordered before the release
    if (threadIdx.x == 0) {
      __threadfence();
      red_release_gpu(ctr);
    }
    ++epoch;
  }
  __device__ __forceinline__ void wait() const {
    if (threadIdx.x == 0) {
      const unsigned target = epoch * (unsigned)PG;
      while (ld_acquire_gpu(ctr) < target) {
      }
      __threadfence();
    }
    __syncthreads();
  }
};

// transpose-reduce: lane l ends with the sum over all lanes of v[l]
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float a = v[i], b = v[i + s];
      const float keep = up ? b : a, send = up ? a : b;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__device__ __forceinline__ float dot4(const float4& a, const float4& w, float acc) {
  acc = fmaf(a.x, w.x, acc);
  acc = fmaf(a.y, w.y, acc);
  acc = fmaf(a.z, w.z, acc);
  return fmaf(a.w, w.w, acc);
}

// One register-blocked tile of a skinny GEMM: out[r*NC + c] = sum_k A[row_r][k] * W[col_c][k] for R rows and NC columns
// held in shared memory (row stride lda4 / ldw4 float4), K = 128 * KS (lane l owns float4 l + 32 ks of every row).
// On return lane (r*NC + c) holds the total of output (r, c).
template <int R, int NC, int KS>
__device__ __forceinline__ float gemv_tile(const float4* __restrict__ sA4, int lda4, const int (&arow)[R],
                                           const float4* __restrict__ sW4, int ldw4, int wrow0, int lane) {
  static_assert(R * NC <= 32, "tile too large for one transpose-reduce");
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    float4 a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = sA4[arow[r] * lda4 + ks * 32 + lane];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 w = sW4[(wrow0 + c) * ldw4 + ks * 32 + lane];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r * NC + c] = dot4(a[r], w, acc[r * NC + c]);
    }
  }
  return transpose_reduce32(acc, lane);
}

// rows [0,B) x 512 floats from global (produced by other CTAs of this launch: L2 loads, never L1) into shared memory
__device__ __forceinline__ void stage_rows(float4* __restrict__ sA4, const float* __restrict__ src, int B, int ld) {
  for (int i = threadIdx.x; i < B * PQ; i += PT) {
    const int b = i / PQ, q = i - b * PQ;
    sA4[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)b * ld) + q);
  }
}

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
  int T, B, A;
};

// attention of one sample by a cluster of CS CTAs (Attention.forward, AttModel.py:411-421): scores, softmax, weighted sum
template <int CS>
__device__ __forceinline__ void attention_fwd_item(cg::cluster_group& cluster, int rank, int b, const float* __restrict__ att_h_row,
                                                   const float* __restrict__ att, const float* __restrict__ p_att,
                                                   float alpha_b, float* __restrict__ pi_out, float* __restrict__ res_out,
                                                   int A, float* s_ah, const float* s_aw, float* s_acc /*[2][PD]*/,
                                                   float* s_e, float* s_ml, float* s_red) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (A + CS - 1) / CS;
  const int a0 = min(A, rank * per), a1 = min(A, a0 + per);
  const int na = a1 - a0;
  for (int q = t; q < PQ; q += PT) reinterpret_cast<float4*>(s_ah)[q] = __ldcg(reinterpret_cast<const float4*>(att_h_row) + q);
  __syncthreads();
  // scores: one warp per location, lanes over the hidden dimension
  for (int a = wid; a < na; a += PT / 32) {
    const float4* row = reinterpret_cast<const float4*>(p_att + ((size_t)b * A + a0 + a) * PD);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PQ / 32; ++k) {
      const int q = lane + 32 * k;
      const float4 p = __ldg(row + q);
      const float4 h = reinterpret_cast<const float4*>(s_ah)[q];
      const float4 w = reinterpret_cast<const float4*>(s_aw)[q];
      s = fmaf(w.x, tanhf_fast_acc(p.x + h.x), s);
      s = fmaf(w.y, tanhf_fast_acc(p.y + h.y), s);
      s = fmaf(w.z, tanhf_fast_acc(p.z + h.z), s);
      s = fmaf(w.w, tanhf_fast_acc(p.w + h.w), s);
    }
    s = warp_sum(s);
    if (lane == 0) s_e[a] = s + alpha_b;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int a = t; a < na; a += PT) m = fmaxf(m, s_e[a]);
  m = warp_max(m);
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < PT / 32; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float l = 0.f;
  for (int a = t; a < na; a += PT) {
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
    for (int w = 0; w < PT / 32; ++w) tot += s_red[w];
    s_ml[0] = m;
    s_ml[1] = tot;
  }
  // partial weighted sum over this slice: thread = (float4 column group, location phase)
  {
    const int q = t & (PQ - 1), ph = t / PQ;      // PT / PQ == 2 phases
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* col = reinterpret_cast<const float4*>(att + ((size_t)b * A + a0) * PD) + q;
#pragma unroll 4
    for (int a = ph; a < na; a += PT / PQ) {
      const float4 v = __ldg(col + (size_t)a * PQ);
      const float w = s_e[a];
      acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
    reinterpret_cast<float4*>(s_acc + ph * PD)[q] = acc;
  }
  cluster.sync();
  // merge the CS partial states
  float ms[CS], ls[CS], M = -INFINITY;
#pragma unroll
  for (int r = 0; r < CS; ++r) {
    const float* rm = cluster.map_shared_rank(s_ml, r);
    ms[r] = rm[0];
    ls[r] = rm[1];
    M = fmaxf(M, ms[r]);
  }
  float L = 0.f, sc[CS];
#pragma unroll
  for (int r = 0; r < CS; ++r) {
    sc[r] = (ls[r] > 0.f) ? expf(ms[r] - M) : 0.f;
    L = fmaf(sc[r], ls[r], L);
  }
  const float invL = 1.f / L;
  constexpr int DPER = PD / CS;                  // this CTA finalises columns [rank*DPER, ...) of att_res
  for (int d = rank * DPER + t; d < (rank + 1) * DPER; d += PT) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < CS; ++r) {
      const float* ra = cluster.map_shared_rank(s_acc, r);
      v = fmaf(sc[r], ra[d] + ra[PD + d], v);
    }
    res_out[d] = v * invL;
  }
  float myscale = 0.f;
#pragma unroll
  for (int r = 0; r < CS; ++r)
    if (r == rank) myscale = sc[r] * invL;
  for (int a = t; a < na; a += PT) pi_out[a0 + a] = s_e[a] * myscale;
  cluster.sync();   // nobody moves on while its shared memory may still be read remotely
}

// shared-memory plan of both kernels (floats)
struct DecSmem {
  static constexpr int W = 32 * PD;                 // stationary weight rows
  static constexpr int A_OFF = W;                   // [PMAXB][PD] staged operand rows
  static constexpr int MISC_OFF = A_OFF + PMAXB * PD;
};

template <int CS>
__global__ void __launch_bounds__(PT, 1) decode_fwd_persist_kernel(const DecFwdArgs p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int j = blockIdx.x;                // hidden units 4j .. 4j+3
  const int cid = j / CS;
  constexpr int NCL = PG / CS;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int B = p.B, A = p.A, T = p.T;
  constexpr int LC = 6 * PD;

  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                               // [32][PD]: 0-3 h2att | 4 + 5u + g: h2h gate g of unit u | 24 + 4h + u: a2c
  float* sA = smem + DecSmem::A_OFF;              // [B][PD]
  float* s_sums = smem + DecSmem::MISC_OFF;       // [PMAXB][20]  h_{t-1} W_h2h^T for the CTA's units (col = 5u + g)
  float* s_a2c = s_sums + PMAXB * 20;             // [PMAXB][8]   col = 4 half + u
  float* s_c = s_a2c + PMAXB * 8;                 // [PMAXB][4]   cell state of the CTA's units
  float* s_ah = s_c + PMAXB * 4;                  // [PD]
  float* s_aw = s_ah + PD;                        // [PD]
  float* s_acc = s_aw + PD;                       // [2][PD]
  float* s_e = s_acc + 2 * PD;                    // [PMAXLOC]
  float* s_ml = s_e + PMAXLOC;                    // [2] (+2 pad)
  float* s_red = s_ml + 4;                        // [8]

  // ---- stationary weights
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

  GridBar gb{p.bar, 0u};
  const float4* sA4 = reinterpret_cast<const float4*>(sA);
  const float4* sW4 = reinterpret_cast<const float4*>(sW);

  for (int step = 0; step < T; ++step) {
    float* cat_t = p.cat_all + (size_t)step * B * LC;
    if (step > 0) {
      stage_rows(reinterpret_cast<float4*>(sA), p.h_all + (size_t)(step - 1) * B * PD, B, PD);
      __syncthreads();
      // ---- A1: att_h columns 4j..4j+3 for rows b = wid + 8 r
      {
        int arow[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) arow[r] = (wid + 8 * r < B) ? wid + 8 * r : 0;
        if (wid < B) {                      // warp uniform
          const float tot = gemv_tile<8, 4, 4>(sA4, PQ, arow, sW4, PQ, 0, lane);
          const int b = wid + 8 * (lane >> 2), c = lane & 3;
          if (b < B) {
            float* dp = cat_t + (size_t)b * LC + 4 * j + c;
            *dp = __ldcg(dp) + tot;
          }
        }
      }
      gb.arrive();
      // ---- A2: the five gate columns of unit u = wid & 3 for rows b = rg + 2 i (runs while barrier 1 completes)
      {
        const int u = wid & 3, rg = wid >> 2;
        for (int i0 = 0; rg + 2 * i0 < B; i0 += 6) {
          int arow[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) arow[r] = (rg + 2 * (i0 + r) < B) ? rg + 2 * (i0 + r) : 0;
          const float tot = gemv_tile<6, 5, 4>(sA4, PQ, arow, sW4, PQ, 4 + 5 * u, lane);
          const int r = lane / 5, g = lane - 5 * r;
          const int b = rg + 2 * (i0 + r);
          if (lane < 30 && b < B) s_sums[b * 20 + 5 * u + g] = tot;
        }
      }
      gb.wait();
    }
    // ---- B: attention, one sample per cluster round
    for (int b = cid; b < B; b += NCL)
      attention_fwd_item<CS>(cluster, rank, b, cat_t + (size_t)b * LC, p.att, p.p_att, alpha_b,
                             p.pi_all + ((size_t)step * B + b) * A, p.res_all + ((size_t)step * B + b) * PD, A, s_ah, s_aw,
                             s_acc, s_e, s_ml, s_red);
    gb.arrive();
    gb.wait();
    // ---- C: a2c columns + gates
    stage_rows(reinterpret_cast<float4*>(sA), p.res_all + (size_t)step * B * PD, B, PD);
    __syncthreads();
    {
      const int half = wid & 1, rg = wid >> 1;
      for (int i0 = 0; rg + 4 * i0 < B; i0 += 8) {
        int arow[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) arow[r] = (rg + 4 * (i0 + r) < B) ? rg + 4 * (i0 + r) : 0;
        const float tot = gemv_tile<8, 4, 4>(sA4, PQ, arow, sW4, PQ, 24 + 4 * half, lane);
        const int b = rg + 4 * (i0 + (lane >> 2)), u = lane & 3;
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
    if (step + 1 < T) {
      gb.arrive();
      gb.wait();
    }
  }
}

size_t dec_fwd_smem() { return (size_t)(DecSmem::MISC_OFF + PMAXB * 32 + 4 * PD + PMAXLOC + 4 + 8) * sizeof(float) + 64; }

template <class Kern, class Args>
int launch_persistent(Kern kern, int cs, size_t smem, cudaStream_t st, const Args& args, bool coop) {
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PG);
  cfg.blockDim = dim3(PT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = coop ? 2 : 1;
  L2S_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, args));
  count_launch();
  return L2S_OK;
}

// how many clusters of `cs` CTAs of this kernel can be resident at once (0 on error)
template <class Kern>
int max_clusters(Kern kern, int cs, size_t smem) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PG);
  cfg.blockDim = dim3(PT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

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
      if (max_clusters(decode_fwd_persist_kernel<8>, 8, dec_fwd_smem()) >= PG / 8) cached_cs = 8;
      else if (max_clusters(decode_fwd_persist_kernel<4>, 4, dec_fwd_smem()) >= PG / 4) cached_cs = 4;
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
  DecFwdArgs a{cat_all, att, p_att, w_cat, w_a2c, b_a2c, alpha_w, alpha_b, h_all, c_all, a2c_all, pi_all, res_all, bar, T, B, A};
  static const bool coop = !env_flag("L2S_DECODE_NOCOOP");
  if (cs == 8) return launch_persistent(decode_fwd_persist_kernel<8>, 8, dec_fwd_smem(), st, a, coop);
  return launch_persistent(decode_fwd_persist_kernel<4>, 4, dec_fwd_smem(), st, a, coop);
}

}  // namespace l2s
