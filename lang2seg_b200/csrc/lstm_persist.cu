// Bidirectional LSTM recurrence of the expression encoder (RNNEncoder.forward, lib/layers/lang_encoder.py:38-80) as ONE
// persistent, weight-stationary kernel per direction of autograd (hidden size 512, B <= 64; lstm.cu keeps the launch
// chain for every other shape).  Both LSTM directions run in the same launch on disjoint halves of the grid: 64 CTAs
// each, with their own grid-barrier counter.
//
// forward   CTA (dir, jj) owns hidden units 8jj..8jj+7 of direction dir: the 32 rows (unit, gate) of W_hh stay in
//           shared memory for all L steps.  Step: stage h_prev (B x 512) -> skinny GEMM (warp = unit: its four gate
//           columns, 8 rows per tile) -> the four gates of (b, unit) meet in one lane by shuffles -> cell -> h_t, c_t,
//           out -> one barrier over the 64 CTAs of the direction.
// backward  cluster of 4 CTAs owns 32 columns of dh_prev = dG_t W_hh, rank = gate = K quarter of 4H; the CTA that
//           reduces a column (DSMEM) also runs the cell backward of that unit in the next step, so the recurrent
//           gradient never leaves shared memory: ONE grid barrier per step (after the cell backward wrote dG_t).
// Masking instead of packing exactly as lstm.cu: sequence b runs for len[b] steps in each direction.
#include "persist.cuh"

namespace l2s {
namespace {

constexpr int LH = 512;
constexpr int LQ = LH / 4;
constexpr int LMAXB = 64;
constexpr int LDIR = PG / 2;      // CTAs per direction

struct LstmFwdArgs {
  float* G;                // (B,L,2,4H) in: x W_ih^T + b ; out: full pre-activations (+ h W_hh^T)
  const float* w_hh;       // (2,4H,H)
  const int* lens;
  float* c_all;            // (L,2,B,H)
  float* h_all;            // (L,2,B,H)
  float* out;              // (B,L,2H)
  float* hidden;           // (B,2H)
  unsigned* bar;           // 2 counters, 128 B apart
  int L, B;
};

__global__ void __launch_bounds__(PT, 1) bilstm_fwd_persist_kernel(const LstmFwdArgs p) {
  const int dir = blockIdx.x / LDIR, jj = blockIdx.x % LDIR;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int B = p.B, L = p.L;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                         // [32][LH]  row = 4 u + gate
  float* sA = sW + 32 * LH;                 // [LMAXB][LH]
  float* s_c = sA + LMAXB * LH;             // [LMAXB][8]
  Stager stg{reinterpret_cast<uint64_t*>(s_c + LMAXB * 8), 0u};
  stg.init();
  for (int i = t; i < 32 * LQ; i += PT) {
    const int r = i / LQ, q = i - r * LQ;
    const int u = r >> 2, g = r & 3;
    reinterpret_cast<float4*>(sW)[i] =
        __ldg(reinterpret_cast<const float4*>(p.w_hh + ((size_t)dir * 4 * LH + (size_t)g * LH + 8 * jj + u) * LH) + q);
  }
  for (int i = t; i < LMAXB * 8; i += PT) s_c[i] = 0.f;
  __syncthreads();
  GridBar gb{p.bar + 32 * dir, 0u, (unsigned)LDIR};
  const float4* sA4 = reinterpret_cast<const float4*>(sA);
  const float4* sW4 = reinterpret_cast<const float4*>(sW);
  const int uw = wid & 7, rg = wid >> 3;     // warp = (unit of the CTA, row group): rows b = rg + 2 i
  const int unit = 8 * jj + uw;

  for (int s = 0; s < L; ++s) {
    const int tt = dir ? L - 1 - s : s;
    const int tp = dir ? tt + 1 : tt - 1;
    if (s > 0) {
      stg.load_contig(sA, p.h_all + ((size_t)(tp * 2 + dir) * B) * LH, (uint32_t)B * LH * 4u);
      stg.wait();
    }
    for (int i0 = 0; rg + 2 * i0 < B; i0 += 8) {
      float tot = 0.f;
      if (s > 0) {
        int arow[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) arow[r] = (rg + 2 * (i0 + r) < B) ? rg + 2 * (i0 + r) : 0;
        tot = gemv_tile<8, 4, 4>(sA4, LQ, arow, sW4, LQ, 4 * uw, lane);
      }
      const int b = rg + 2 * (i0 + (lane >> 2)), g = lane & 3;
      float pre = 0.f;
      if (b < B) {
        float* gp = p.G + (((size_t)b * L + tt) * 2 + dir) * 4 * LH + (size_t)g * LH + unit;
        pre = *gp + tot;
        if (s > 0) *gp = pre;
      }
      const int base = lane & ~3;
      const float pf = __shfl_sync(0xffffffffu, pre, base + 1), pg = __shfl_sync(0xffffffffu, pre, base + 2);
      const float po = __shfl_sync(0xffffffffu, pre, base + 3);
      if (g == 0 && b < B) {
        const float cp = s > 0 ? s_c[b * 8 + uw] : 0.f;
        const float hp = s > 0 ? sA[(size_t)b * LH + unit] : 0.f;
        float c = cp, h = hp, o_out = 0.f;
        if (tt < __ldg(p.lens + b)) {
          const float ig = sigmoidf_acc(pre), fg = sigmoidf_acc(pf), gg = tanhf(pg), og = sigmoidf_acc(po);
          c = fmaf(fg, cp, ig * gg);
          h = og * tanhf(c);
          o_out = h;
        }
        s_c[b * 8 + uw] = c;
        const size_t st = ((size_t)(tt * 2 + dir) * B + b) * LH + unit;
        p.c_all[st] = c;
        p.h_all[st] = h;
        p.out[((size_t)b * L + tt) * 2 * LH + dir * LH + unit] = o_out;
        if (s == L - 1) p.hidden[(size_t)b * 2 * LH + dir * LH + unit] = h;
      }
    }
    if (s + 1 < L) {
      gb.arrive();
      gb.wait();
    }
  }
}

struct LstmBwdArgs {
  const float* dout;       // (B,L,2H) or null
  const float* dhidden;    // (B,2H) or null
  const float* G;          // (B,L,2,4H) full pre-activations
  const float* w_hh_t;     // (2,H,4H) = W_hh^T per direction
  const int* lens;
  const float* c_all;      // (L,2,B,H)
  float* dG;               // (B,L,2,4H)
  unsigned* bar;
  int L, B;
};

constexpr int LBCS = 4;           // backward cluster: rank = gate = K quarter

__global__ void __launch_bounds__(PT, 1) bilstm_bwd_persist_kernel(const LstmBwdArgs p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int dir = blockIdx.x / LDIR;
  const int cid = (blockIdx.x % LDIR) / LBCS;          // 16 clusters per direction x 32 columns
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int B = p.B, L = p.L;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                         // [32][LH]: W_hh^T rows 32 cid + i, K quarter `rank`
  float* sA = sW + 32 * LH;                 // [LMAXB][LH]
  float* s_part = sA + LMAXB * LH;          // [LMAXB][32]
  float* s_dh = s_part + LMAXB * 32;        // [LMAXB][8] recurrent gradient of the CTA's units
  float* s_dc = s_dh + LMAXB * 8;           // [LMAXB][8]
  float* s_dhp = s_dc + LMAXB * 8;          // [LMAXB][8] gradient passed through masked steps
  Stager stg{reinterpret_cast<uint64_t*>(s_dhp + LMAXB * 8), 0u};
  stg.init();
  const int cgp = wid & 7, rg = wid >> 3;   // warp = (4-column group, row group): rows b = rg + 2 i
  for (int i = t; i < 32 * LQ; i += PT) {
    const int r = i / LQ, q = i - r * LQ;
    reinterpret_cast<float4*>(sW)[i] = __ldg(
        reinterpret_cast<const float4*>(p.w_hh_t + ((size_t)dir * LH + 32 * cid + r) * 4 * LH + (size_t)LH * rank) + q);
  }
  // elementwise phases: item e = (sample e / 8, unit e % 8 of this CTA's eight units); e = t, t + 256
  const int ubase = 32 * cid + 8 * rank;
  for (int e = t; e < 8 * B; e += PT) {
    s_dh[e] = 0.f;
    s_dc[e] = 0.f;
    s_dhp[e] = p.dhidden ? __ldg(p.dhidden + (size_t)(e >> 3) * 2 * LH + dir * LH + ubase + (e & 7)) : 0.f;
  }
  __syncthreads();
  GridBar gb{p.bar + 32 * dir, 0u, (unsigned)LDIR};
  const float4* sA4 = reinterpret_cast<const float4*>(sA);
  const float4* sW4 = reinterpret_cast<const float4*>(sW);
  const size_t ldg = (size_t)L * 8 * LH;

  for (int s = L - 1; s >= 0; --s) {
    const int tt = dir ? L - 1 - s : s;
    const int tp = dir ? tt + 1 : tt - 1;
    // ---- cell backward of (ub, unit)
    for (int e = t; e < 8 * B; e += PT) {
      const int ub = e >> 3, unit = ubase + (e & 7);
      const float dh_state = s_dh[e] + s_dhp[e];
      const float dc_state = s_dc[e];
      const size_t go = (((size_t)ub * L + tt) * 2 + dir) * 4 * LH + unit;
      float* dgp = p.dG + go;
      if (tt < __ldg(p.lens + ub)) {
        const float* gp = p.G + go;
        const float ig = sigmoidf_acc(__ldg(gp)), fg = sigmoidf_acc(__ldg(gp + LH));
        const float gg = tanhf(__ldg(gp + 2 * LH)), og = sigmoidf_acc(__ldg(gp + 3 * LH));
        const size_t st = ((size_t)(tt * 2 + dir) * B + ub) * LH + unit;
        const float cp = s == 0 ? 0.f : __ldg(p.c_all + ((size_t)(tp * 2 + dir) * B + ub) * LH + unit);
        const float tc = tanhf(__ldg(p.c_all + st));
        const float dh = dh_state + (p.dout ? __ldg(p.dout + ((size_t)ub * L + tt) * 2 * LH + dir * LH + unit) : 0.f);
        const float dc = dc_state + dh * og * (1.f - tc * tc);
        dgp[0] = dc * gg * ig * (1.f - ig);
        dgp[LH] = dc * cp * fg * (1.f - fg);
        dgp[2 * LH] = dc * ig * (1.f - gg * gg);
        dgp[3 * LH] = dh * tc * og * (1.f - og);
        s_dc[e] = dc * fg;
        s_dhp[e] = 0.f;                  // the previous state receives its dh through dG . W_hh
      } else {
        dgp[0] = 0.f; dgp[LH] = 0.f; dgp[2 * LH] = 0.f; dgp[3 * LH] = 0.f;
        s_dc[e] = dc_state;              // untouched state: gradients pass straight through
        s_dhp[e] = dh_state;
      }
    }
    if (s == 0) break;
    gb.arrive();
    gb.wait();
    // ---- dh_prev[:, 32 cid ..] = dG_t[:, gate `rank`] . W_hh^T slice, summed over the four gates through DSMEM
    stg.load_rows(sA, p.dG + ((size_t)tt * 2 + dir) * 4 * LH + (size_t)LH * rank, B, LH * 4, ldg * 4);
    stg.wait();
    for (int i0 = 0; rg + 2 * i0 < B; i0 += 8) {
      int arow[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) arow[r] = (rg + 2 * (i0 + r) < B) ? rg + 2 * (i0 + r) : 0;
      const float tot = gemv_tile<8, 4, 4>(sA4, LQ, arow, sW4, LQ, 4 * cgp, lane);
      const int b = rg + 2 * (i0 + (lane >> 2));
      if (b < B) s_part[b * 32 + 4 * cgp + (lane & 3)] = tot;
    }
    cluster.sync();
    for (int e = t; e < 8 * B; e += PT) {
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < LBCS; ++q) v += cluster.map_shared_rank(s_part, q)[(e >> 3) * 32 + 8 * rank + (e & 7)];
      s_dh[e] = v;
    }
    // s_part is next written after the next grid barrier: every peer has finished reading it by then
  }
}

size_t lstm_fwd_smem() { return (size_t)(32 * LH + LMAXB * LH + LMAXB * 8 + 4) * sizeof(float) + 64; }
size_t lstm_bwd_smem() { return (size_t)(32 * LH + LMAXB * LH + LMAXB * 32 + 3 * LMAXB * 8 + 4) * sizeof(float) + 64; }

}  // namespace

bool bilstm_persist_ok(int L, int B, int H) {
  static const bool off = env_flag("L2S_LSTM_CHAIN");         // diagnostics: force the launch chain of lstm.cu
  if (off || H != LH || B < 1 || B > LMAXB || L < 1) return false;
  static thread_local int cached_dev = -1;
  static thread_local bool cached_ok = false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev != cached_dev) {
    cached_dev = dev;
    cached_ok = sm_count() >= PG && (size_t)max_smem_optin() >= lstm_bwd_smem() &&
                max_clusters(bilstm_fwd_persist_kernel, 1, lstm_fwd_smem()) >= PG &&
                max_clusters(bilstm_bwd_persist_kernel, LBCS, lstm_bwd_smem()) >= PG / LBCS;
  }
  return cached_ok;
}

int launch_bilstm_fwd_persist(float* G, const float* w_hh, const int* lens, float* c_all, float* h_all, float* out,
                              float* hidden, int L, int B, unsigned* bar, cudaStream_t st) {
  L2S_REQUIRE(aligned16(G) && aligned16(w_hh) && aligned16(h_all), L2S_ERR_ALIGN, "bilstm_fwd: pointers must be 16-byte aligned");
  L2S_CUDA_OK(cudaMemsetAsync(bar, 0, 256, st));
  LstmFwdArgs a{G, w_hh, lens, c_all, h_all, out, hidden, bar, L, B};
  static const bool coop = !env_flag("L2S_DECODE_NOCOOP");
  return launch_persistent(bilstm_fwd_persist_kernel, 1, lstm_fwd_smem(), st, a, coop);
}

int launch_bilstm_bwd_persist(const float* dout, const float* dhidden, const float* G, const float* w_hh_t,
                              const int* lens, const float* c_all, float* dG, int L, int B, unsigned* bar,
                              cudaStream_t st) {
  L2S_REQUIRE(aligned16(dG) && aligned16(w_hh_t), L2S_ERR_ALIGN, "bilstm_bwd: pointers must be 16-byte aligned");
  L2S_CUDA_OK(cudaMemsetAsync(bar, 0, 256, st));
  LstmBwdArgs a{dout, dhidden, G, w_hh_t, lens, c_all, dG, bar, L, B};
  static const bool coop = !env_flag("L2S_DECODE_NOCOOP");
  return launch_persistent(bilstm_bwd_persist_kernel, LBCS, lstm_bwd_smem(), st, a, coop);
}

}  // namespace l2s
