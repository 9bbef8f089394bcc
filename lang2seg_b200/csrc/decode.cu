// att2in2 decode loop (AttModel.forward's T-step recurrence, lib/caption_models/AttModel.py:75-99 with
// Att2in2Core.forward :446-466 and Attention.forward :406-423) as ONE library call per direction.
//
// The reference runs ~17 framework launches per token.  Here everything that does not depend on the
// recurrent state is hoisted out of the loop by the caller (i2h(xt) for all t as one GEMM, logit +
// log-softmax + NLL for all t as one GEMM + one kernel, every weight gradient as one GEMM over the
// stacked T*B rows) and a step is four launches, issued back to back from this translation unit:
//
//   fwd  t:  [att_h | sums]_t += h_{t-1} . [W_h2att ; W_h2h]^T      linear_small_kernel (exact fp32 FFMA)
//            att_res_t, pi_t   = attention(att_h_t, p_att, att)     att_step_fwd_kernel (cluster of 4 CTAs)
//            a2c_t             = att_res_t . W_a2c^T + b            linear_small_kernel
//            h_t, c_t          = gates(sums_t, a2c_t, c_{t-1})      gates_fwd_kernel
//   bwd  t:  gates_bwd -> dsums_t, da2c_t, dc ; datt_res_t = da2c_t . W_a2c (split-K) ;
//            att_step_bwd (light: datt_h_t and de_t only) ; dh_{t-1} = [datt_h|dsums]_t . [W_h2att;W_h2h]
//   after the loop: att_accum_kernel turns the stored de_t / pi_t / datt_res_t into dp_att and datt_feats
//   in ONE pass over p_att (the per-step read-modify-write of two (B,A,512) tensors is gone: the
//   algorithmic traffic of SURVEY 8d, (2T+2) * 2*A*D*4 bytes per expression).
//
// linear_small_kernel: D[M,N] (+)= A[M,K] . W[N,K]^T (+ bias) for M <= 64 rows per CTA ("skinny" GEMM,
// batch of expressions x hidden size).  A K-chunk of A sits in shared memory; a warp owns CPW columns,
// lanes split K (coalesced float4 weight rows straight from L2), per-(row,column) partial sums are
// combined with a 31-shuffle transpose-reduce per 32 rows.  K > 512 is split across CTAs; the last CTA
// to arrive sums the partials in a fixed order (deterministic, no float atomics).
#include "common.cuh"

namespace l2s {
namespace {

constexpr int kLinThreads = 256;
constexpr int kLinKC = 512;      // K chunk held in shared memory (lane owns 4 float4 of it)
constexpr int kLinMaxRows = 64;  // rows of A per CTA

// transpose-reduce: lane l ends with the sum over all lanes of v[l % NR]  (NR = 32 or 16)
template <int NR>
__device__ __forceinline__ float transpose_reduce(float (&v)[NR], int lane) {
#pragma unroll
  for (int s = NR / 2; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float a = v[i], b = v[i + s];
      const float keep = up ? b : a, send = up ? a : b;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  float r = v[0];
  if (NR == 16) r += __shfl_xor_sync(0xffffffffu, r, 16);   // the two half-warps hold the two halves of the K sum
  return r;
}

struct LinGeom {
  int M, N, K;
  int lda, ldw, ldd;
  int ks;          // number of K splits (gridDim.y)
  int accumulate;  // D += result
  int rows_pad;    // rows of A staged per CTA, padded to a multiple of 16 (<= 64)
  int rows_blk;    // rows of A per CTA (blockIdx.z selects the row block)
};

// NR rows of the staged A chunk against this warp's CPW weight rows; lane (l % NR) ends with row mg + l % NR
template <int CPW, int NR, class Store>
__device__ __forceinline__ void row_group(const float4* __restrict__ As4, const float4 (&w)[CPW][4], int mg, int lane,
                                          int n0, Store& store) {
  float p[CPW][NR];
#pragma unroll
  for (int mi = 0; mi < NR; ++mi) {
    const float4* row = As4 + (size_t)(mg + mi) * (kLinKC / 4) + lane;
    const float4 a0 = row[0], a1 = row[32], a2 = row[64], a3 = row[96];
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      float acc = a0.x * w[c][0].x;
      acc = fmaf(a0.y, w[c][0].y, acc); acc = fmaf(a0.z, w[c][0].z, acc); acc = fmaf(a0.w, w[c][0].w, acc);
      acc = fmaf(a1.x, w[c][1].x, acc); acc = fmaf(a1.y, w[c][1].y, acc);
      acc = fmaf(a1.z, w[c][1].z, acc); acc = fmaf(a1.w, w[c][1].w, acc);
      acc = fmaf(a2.x, w[c][2].x, acc); acc = fmaf(a2.y, w[c][2].y, acc);
      acc = fmaf(a2.z, w[c][2].z, acc); acc = fmaf(a2.w, w[c][2].w, acc);
      acc = fmaf(a3.x, w[c][3].x, acc); acc = fmaf(a3.y, w[c][3].y, acc);
      acc = fmaf(a3.z, w[c][3].z, acc); acc = fmaf(a3.w, w[c][3].w, acc);
      p[c][mi] = acc;
    }
  }
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const float tot = transpose_reduce<NR>(p[c], lane);
    if (NR == 32 || lane < 16) store(mg + (lane % NR), n0 + c, tot);
  }
}

template <int CPW>
__global__ void __launch_bounds__(kLinThreads, 2)
linear_small_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ D, float* __restrict__ partials, unsigned int* __restrict__ counters,
                    LinGeom g) {
  extern __shared__ __align__(16) float As[];      // [rows_pad][kLinKC]
  __shared__ int s_last;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cb = blockIdx.x, s = blockIdx.y, mb = blockIdx.z;
  const int m0 = mb * g.rows_blk;
  const int rows = min(g.rows_blk, g.M - m0);      // valid rows in this block
  const int rp = min(g.rows_pad, (rows + 15) & ~15);   // rows this block stages and multiplies (the last block may be short)
  const int k0 = s * kLinKC;
  const int kc = min(kLinKC, g.K - k0);            // valid k in this chunk (multiple of 4)

  // ---- this warp's weight rows: lane holds float4 #(lane + 32 j), j = 0..3 of the chunk.  The weights do not depend
  // on the previous kernel of the chain, so they are requested BEFORE the dependency wait (programmatic dependent
  // launch): their L2 latency overlaps the tail of that kernel.
  pdl_launch_dependents();
  const int n0 = (cb * (kLinThreads / 32) + wid) * CPW;
  float4 w[CPW][4];
#pragma unroll
  for (int c = 0; c < CPW; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = lane + 32 * j;
      w[c][j] = (n0 + c < g.N && 4 * q < kc)
                    ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + c) * g.ldw + k0) + q)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  pdl_wait();

  // ---- stage the A chunk (zero fill outside M / K)
  {
    const int nf4 = rp * (kLinKC / 4);
    float4* As4 = reinterpret_cast<float4*>(As);
#pragma unroll 4
    for (int i = t; i < nf4; i += kLinThreads) {
      const int r = i / (kLinKC / 4), q = i % (kLinKC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows && 4 * q < kc) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * g.lda + k0) + q);
      As4[i] = v;
    }
  }
  __syncthreads();

  const float4* As4 = reinterpret_cast<const float4*>(As);
  auto store = [&](int mrow, int n, float tot) {
    const int m = m0 + mrow;
    if (mrow < rows && n < g.N) {
      if (g.ks == 1) {
        float* dp = D + (size_t)m * g.ldd + n;
        float v = tot + (bias ? __ldg(bias + n) : 0.f);
        if (g.accumulate) v += *dp;
        *dp = v;
      } else {
        partials[((size_t)s * g.M + m) * g.N + n] = tot;
      }
    }
  };
  int mg = 0;
  for (; mg + 32 <= rp; mg += 32) row_group<CPW, 32>(As4, w, mg, lane, n0, store);
  if (mg < rp) row_group<CPW, 16>(As4, w, mg, lane, n0, store);
  if (g.ks == 1) return;

  // ---- split-K: the last CTA of this (column block, row block) sums the partials in split order
  __threadfence();
  __syncthreads();
  const int slot = mb * gridDim.x + cb;
  if (t == 0) {
    const unsigned int ticket = atomicAdd(counters + slot, 1u);
    s_last = (ticket == (unsigned int)g.ks - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nb0 = cb * (kLinThreads / 32) * CPW;
  const int ncols = min((kLinThreads / 32) * CPW, g.N - nb0);
  const int mrows = rows;
  for (int i = t; i < mrows * ncols; i += kLinThreads) {
    const int m = m0 + i / ncols, n = nb0 + i % ncols;
    float v = bias ? __ldg(bias + n) : 0.f;
    for (int ss = 0; ss < g.ks; ++ss) v += __ldcg(partials + ((size_t)ss * g.M + m) * g.N + n);
    float* dp = D + (size_t)m * g.ldd + n;
    if (g.accumulate) v += *dp;
    *dp = v;
  }
  if (t == 0) counters[slot] = 0u;   // ready for the next launch on this stream
}

// split-K arrival counters: one per (column block, row block); sized for the narrowest column blocks (8 columns)
size_t lin_counter_bytes(int M, int N) {
  const int mblocks = M > 32 ? (M + 31) / 32 : 1;
  return (((size_t)((N + 7) / 8) * mblocks * sizeof(unsigned int)) + 255) & ~(size_t)255;
}

int pick_cpw(int N, int ks, int mblocks) {
  // largest column count per warp that still gives every SM a CTA; more columns per warp reuse the
  // shared-memory A rows more often
  const int sms = sm_count();
  for (int cpw = 3; cpw >= 2; --cpw) {
    const int cbs = (N + 8 * cpw - 1) / (8 * cpw);
    if (cbs * ks * mblocks >= (sms * 3) / 4) return cpw;
  }
  return 1;
}

}  // namespace

size_t linear_small_counter_bytes(int M, int N) { return lin_counter_bytes(M > 0 ? M : 1, N > 0 ? N : 1); }

size_t linear_small_workspace_bytes(int M, int N, int K) {
  const int ks = (K + kLinKC - 1) / kLinKC;
  return linear_small_counter_bytes(M, N) + (ks > 1 ? (size_t)ks * M * N * sizeof(float) : 0) + 256;
}

// workspace: [linear_small_counter_bytes(M,N) of counters (zero before first use; the kernel leaves them zero) | partials]
int launch_linear_small(const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M,
                        int N, int K, int accumulate, void* workspace, size_t ws_bytes, cudaStream_t st) {
  L2S_REQUIRE(A && W && D, L2S_ERR_ARG, "linear_small: null pointer");
  L2S_REQUIRE(M > 0 && N > 0 && K > 0, L2S_ERR_SHAPE, "linear_small: bad shape M=%d N=%d K=%d", M, N, K);
  L2S_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && aligned16(A) && aligned16(W), L2S_ERR_ALIGN,
              "linear_small: K and the row strides must be multiples of 4 floats and A, W 16-byte aligned");
  L2S_REQUIRE(workspace && ws_bytes >= linear_small_workspace_bytes(M, N, K), L2S_ERR_WORKSPACE,
              "linear_small: workspace too small");
  LinGeom g;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.ldd = ldd;
  g.ks = (K + kLinKC - 1) / kLinKC;
  g.accumulate = accumulate;
  // Row blocks of 32: at M = 48 (expressions per GPU) one 48-row CTA per SM leaves 8 warps per SM, too few to hide the
  // smem / FFMA latencies of the inner loop; a 32-row and a 16-row CTA (64 + 32 KB of shared memory) co-reside and
  // double the warps in flight for the same staging traffic.
  g.rows_blk = M > 32 ? 32 : M;
  const int mblocks = (M + g.rows_blk - 1) / g.rows_blk;
  g.rows_pad = (g.rows_blk + 15) & ~15;   // staged rows per CTA: 16 or 32
  const int cpw = pick_cpw(N, g.ks, mblocks);
  const int cbs = (N + 8 * cpw - 1) / (8 * cpw);
  const size_t cbytes = linear_small_counter_bytes(M, N);
  L2S_REQUIRE((size_t)cbs * mblocks * sizeof(unsigned int) <= cbytes, L2S_ERR_SHAPE,
              "linear_small: counter area too small for M=%d N=%d", M, N);
  unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + cbytes);
  const size_t smem = (size_t)g.rows_pad * kLinKC * sizeof(float);
  dim3 grid(cbs, g.ks, mblocks);
#define L2S_LIN_LAUNCH(CPW)                                                                              \
  do {                                                                                                   \
    auto kern = linear_small_kernel<CPW>;                                                                \
    L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    L2S_CUDA_OK(launch_chain(kern, grid, dim3(kLinThreads), smem, st, A, W, bias, D, partials, counters, g)); \
  } while (0)
  if (cpw == 3) L2S_LIN_LAUNCH(3);
  else if (cpw == 2) L2S_LIN_LAUNCH(2);
  else L2S_LIN_LAUNCH(1);
#undef L2S_LIN_LAUNCH
  L2S_LAUNCH_OK("linear_small_kernel");
  count_launch();
  return L2S_OK;
}

namespace {

// ------------------------------------------------------------------ deferred attention gradients
// dp_att[b,a,:]  = alpha * sum_t de[t,b,a] * (1 - tanh^2(p_att[b,a,:] + att_h[t,b,:]))
// datt[b,a,:]    = sum_t pi[t,b,a] * datt_res[t,b,:]
// dalpha_part[cta,:] = sum_{a in chunk} sum_t de[t,b,a] * tanh(...)        (summed by colsum_kernel)
// CTA = (sample b, chunk of kAccLoc locations); thread = (float4 column group, location phase).
constexpr int kAccThreads = 256;
constexpr int kAccLoc = 14;

__global__ void __launch_bounds__(kAccThreads)
att_accum_kernel(const float* __restrict__ p_att, const float* __restrict__ cat_all /* att_h rows, stride ldc */,
                 int ldc, const float* __restrict__ dres_all, const float* __restrict__ de_all,
                 const float* __restrict__ pi_all, const float* __restrict__ alpha_w, float* __restrict__ dp_att,
                 float* __restrict__ datt, float* __restrict__ dalpha_part, int T, int B, int A, int D, int Dh) {
  extern __shared__ __align__(16) float sm[];
  float* s_ah = sm;                          // [T][Dh]
  float* s_dr = s_ah + (size_t)T * Dh;       // [T][D]
  float* s_de = s_dr + (size_t)T * D;        // [T][kAccLoc]
  float* s_pi = s_de + T * kAccLoc;          // [T][kAccLoc]
  float* s_da = s_pi + T * kAccLoc;          // [Dh] alpha-gradient partial (phase 1 adds onto phase 0)
  const int b = blockIdx.y, a0 = blockIdx.x * kAccLoc, na = min(kAccLoc, A - a0);
  const int t = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = t; i < T * Dh; i += kAccThreads) {
    const int tt = i / Dh, d = i - tt * Dh;
    s_ah[i] = __ldg(cat_all + ((size_t)tt * B + b) * ldc + d);
  }
  for (int i = t; i < T * D; i += kAccThreads) {
    const int tt = i / D, d = i - tt * D;
    s_dr[i] = __ldg(dres_all + ((size_t)tt * B + b) * D + d);
  }
  for (int i = t; i < T * kAccLoc; i += kAccThreads) {
    const int tt = i / kAccLoc, a = i - tt * kAccLoc;
    const bool ok = a < na;
    s_de[i] = ok ? __ldg(de_all + ((size_t)tt * B + b) * A + a0 + a) : 0.f;
    s_pi[i] = ok ? __ldg(pi_all + ((size_t)tt * B + b) * A + a0 + a) : 0.f;
  }
  __syncthreads();
  // D == Dh is required by the host (both 512 in att2in2), so one column quad serves both outputs
  const int nq = Dh / 4;
  const int nph = max(1, kAccThreads / nq);
  const int items = nq * nph;
  for (int base = 0; base < items; base += kAccThreads) {
    const int qq = base + t;
    const bool on = qq < items;
    const int q = on ? qq % nq : 0, ph = on ? qq / nq : 0;
    const float4 al = __ldg(reinterpret_cast<const float4*>(alpha_w) + q);
    float4 daw = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = on ? ph : na; a < na; a += nph) {
      const size_t off = ((size_t)b * A + a0 + a) * Dh;
      const float4 p = __ldg(reinterpret_cast<const float4*>(p_att + off) + q);
      float4 dp = make_float4(0.f, 0.f, 0.f, 0.f), da = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int tt = 0; tt < T; ++tt) {
        const float de = s_de[tt * kAccLoc + a], pi = s_pi[tt * kAccLoc + a];
        const float4 h = reinterpret_cast<const float4*>(s_ah + (size_t)tt * Dh)[q];
        const float4 dr = reinterpret_cast<const float4*>(s_dr + (size_t)tt * D)[q];
        float4 th;
        th.x = tanhf_fast_acc(p.x + h.x); th.y = tanhf_fast_acc(p.y + h.y);
        th.z = tanhf_fast_acc(p.z + h.z); th.w = tanhf_fast_acc(p.w + h.w);
        dp.x = fmaf(de, 1.f - th.x * th.x, dp.x); dp.y = fmaf(de, 1.f - th.y * th.y, dp.y);
        dp.z = fmaf(de, 1.f - th.z * th.z, dp.z); dp.w = fmaf(de, 1.f - th.w * th.w, dp.w);
        daw.x = fmaf(de, th.x, daw.x); daw.y = fmaf(de, th.y, daw.y);
        daw.z = fmaf(de, th.z, daw.z); daw.w = fmaf(de, th.w, daw.w);
        da.x = fmaf(pi, dr.x, da.x); da.y = fmaf(pi, dr.y, da.y);
        da.z = fmaf(pi, dr.z, da.z); da.w = fmaf(pi, dr.w, da.w);
      }
      dp.x *= al.x; dp.y *= al.y; dp.z *= al.z; dp.w *= al.w;
      reinterpret_cast<float4*>(dp_att + off)[q] = dp;
      reinterpret_cast<float4*>(datt + off)[q] = da;
    }
    // combine the phases' alpha partials in a fixed order
    for (int r = 0; r < nph; ++r) {
      if (on && ph == r) {
        float4* dst = reinterpret_cast<float4*>(s_da) + q;
        float4 o = (r == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : *dst;
        o.x += daw.x; o.y += daw.y; o.z += daw.z; o.w += daw.w;
        *dst = o;
      }
      __syncthreads();
    }
    if (on && ph == 0) {
      const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      reinterpret_cast<float4*>(dalpha_part + cta * Dh)[q] = reinterpret_cast<const float4*>(s_da)[q];
    }
    __syncthreads();
  }
}

// out[c] = sum_r in[r][c]   (fixed order)
__global__ void colsum_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (c >= C) return;
  float s = 0.f;
  int r = 0;
  for (; r < R; ++r) s += in[(size_t)r * C + c];
  out[c] = s;
}

// One-stage variant for up to ~1000 rows (the T*B = 528 rows of the decode chain, the 672 dalpha partials): CTA = 32
// columns x 32 warps, warp w adds rows w, w + 32, ... (lane = column: one 128-byte line per row), the 32 partial sums are
// added in warp order.  The two-stage pair below costs 32 dependent round trips + a second launch (10-14 us) for these
// sizes; ptxas keeps only a couple of a thread's loads in flight, so the parallelism has to come from warps.
constexpr int kColsumSmallWarps = 32;
__global__ void __launch_bounds__(kColsumSmallWarps * 32)
colsum_small_kernel(const float* __restrict__ in, int64_t ld, float* __restrict__ out, int R, int C) {
  __shared__ float s[kColsumSmallWarps][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    int r = wid;
    for (; r + kColsumSmallWarps < R; r += 2 * kColsumSmallWarps) {
      a0 += __ldg(in + (size_t)r * ld + c);
      a1 += __ldg(in + (size_t)(r + kColsumSmallWarps) * ld + c);
    }
    if (r < R) a0 += __ldg(in + (size_t)r * ld + c);
  }
  s[wid][lane] = a0 + a1;
  __syncthreads();
  if (wid == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kColsumSmallWarps; ++w) t += s[w][lane];
    out[c] = t;
  }
}

// partial[chunk][c] = sum of rows [chunk*RPC, (chunk+1)*RPC) of in (R x C, row stride ld): thread = column, coalesced rows
constexpr int kColsumRows = 128;
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const float* __restrict__ in, int64_t ld, float* __restrict__ partial, int R, int C) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * kColsumRows, r1 = min(R, r0 + kColsumRows);
  if (c >= C) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    s0 += __ldg(in + (size_t)r * ld + c);
    s1 += __ldg(in + (size_t)(r + 1) * ld + c);
    s2 += __ldg(in + (size_t)(r + 2) * ld + c);
    s3 += __ldg(in + (size_t)(r + 3) * ld + c);
  }
  for (; r < r1; ++r) s0 += __ldg(in + (size_t)r * ld + c);
  partial[(size_t)blockIdx.y * C + c] = (s0 + s1) + (s2 + s3);
}

struct DecodeWs {
  unsigned* bar;      // grid-barrier counter of the persistent kernels (256 B)
  char* lin;          // linear_small workspace (shared by every GEMM of the loop: they are stream ordered)
  size_t lin_bytes;
  float* dalpha_part; // [B * ceil(A/kAccLoc)][Dh]
  float* carry;       // dh carry (B,D), dc ping-pong 2 x (B,D)
};

size_t decode_ws_layout(int B, int A, int D, int Dh, DecodeWs* w, char* base) {
  const int LC = Dh + 5 * D;
  size_t lin = linear_small_workspace_bytes(B, LC, D);
  lin = lin > linear_small_workspace_bytes(B, 2 * D, D) ? lin : linear_small_workspace_bytes(B, 2 * D, D);
  lin = lin > linear_small_workspace_bytes(B, D, 2 * D) ? lin : linear_small_workspace_bytes(B, D, 2 * D);
  lin = lin > linear_small_workspace_bytes(B, D, LC) ? lin : linear_small_workspace_bytes(B, D, LC);
  lin = (lin + 255) & ~(size_t)255;
  const size_t nchunk = (size_t)B * ((A + kAccLoc - 1) / kAccLoc);
  size_t off = 0;
  if (w) w->bar = reinterpret_cast<unsigned*>(base + off);
  off += 256;
  if (w) { w->lin = base + off; w->lin_bytes = lin; }
  off += lin;
  if (w) w->dalpha_part = reinterpret_cast<float*>(base + off);
  off += ((nchunk * Dh * sizeof(float)) + 255) & ~(size_t)255;
  if (w) w->carry = reinterpret_cast<float*>(base + off);
  off += (size_t)3 * B * D * sizeof(float);
  return off + 256;
}

int check_decode(int T, int B, int A, int D, int Dh) {
  L2S_REQUIRE(T > 0 && B > 0 && A > 0, L2S_ERR_SHAPE, "att2in2_decode: bad shape T=%d B=%d A=%d", T, B, A);
  L2S_REQUIRE(D == Dh && D % 4 == 0 && D <= 1024, L2S_ERR_SHAPE,
              "att2in2_decode: rnn_size must equal att_hid_size, be a multiple of 4 and <= 1024 (got %d, %d)", D, Dh);
  return L2S_OK;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" size_t l2s_linear_small_workspace_bytes(int M, int N, int K) { return linear_small_workspace_bytes(M, N, K); }

extern "C" int l2s_linear_small(const float* A, const float* W, const float* bias, float* D, int M, int N, int K,
                                int lda, int ldw, int ldd, int accumulate, void* workspace, size_t workspace_bytes,
                                l2s_stream_t stream) {
  L2S_REQUIRE(workspace, L2S_ERR_WORKSPACE, "linear_small: workspace missing");
  cudaStream_t st = (cudaStream_t)stream;
  L2S_REQUIRE(M > 0 && N > 0, L2S_ERR_SHAPE, "linear_small: bad shape M=%d N=%d", M, N);
  L2S_REQUIRE(workspace_bytes >= linear_small_workspace_bytes(M, N, K), L2S_ERR_WORKSPACE, "linear_small: workspace too small");
  L2S_CUDA_OK(cudaMemsetAsync(workspace, 0, linear_small_counter_bytes(M, N), st));
  return launch_linear_small(A, lda, W, ldw, bias, D, ldd, M, N, K, accumulate, workspace, workspace_bytes, st);
}

// dW[v,:] = sum_{r: idx[r] == v} dy[r,:] in a FIXED order (bit-reproducible).  CTA = vocabulary row v, 16 warps.  Warp w
// takes the 32-row blocks w, w + 16, ... of the index list (lane = index, one ballot per block) and adds the matching
// rows, four float4 columns per lane; the warps' partial sums are then added in warp order by warp 0.
// Zero-padded captions make token 0 match half of the ~500 rows of a step, and ptxas keeps only two of a lane's loads in
// flight however the loop is written -- versions with 4 warps that walked the matches of that one CTA took 25-75 us per
// launch (torch's sort + segmented reduce: 24-30 us); with 16 warps the matches of the hot row are spread over 16
// independent load streams.  The zero fill of unmatched rows is part of the store.
constexpr int kEmbWarps = 16;
__global__ void __launch_bounds__(kEmbWarps * 32)
embedding_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dy, float* __restrict__ dW, int R, int D) {
  __shared__ float4 s_acc[kEmbWarps - 1][128];
  __shared__ int s_any;
  const int v = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nq = D >> 2;
  for (int q0 = 0; q0 < nq; q0 += 128) {                 // column block of 128 float4: lane owns q0 + lane + 32 c
    __syncthreads();                                      // s_acc / s_any of the previous column block read out
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    float4 acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool any = false;
    for (int r0 = 32 * wid; r0 < R; r0 += 32 * kEmbWarps) {
      const int r = r0 + lane;
      const int my = r < R ? (int)__ldg(idx + r) : -1;
      unsigned m = __ballot_sync(0xffffffffu, my == v);
      any |= m != 0;
      while (m) {
        const float4* row = reinterpret_cast<const float4*>(dy + (size_t)(r0 + __ffs(m) - 1) * D);
        m &= m - 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int q = q0 + lane + 32 * c;
          if (q < nq) {
            const float4 x = __ldg(row + q);
            acc[c].x += x.x; acc[c].y += x.y; acc[c].z += x.z; acc[c].w += x.w;
          }
        }
      }
    }
    if (any && lane == 0) s_any = 1;                      // benign race: every writer stores 1
    if (wid > 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s_acc[wid - 1][lane + 32 * c] = acc[c];
    }
    __syncthreads();
    if (wid == 0) {
      const bool some = s_any != 0;                       // rows nobody matched (most of the vocabulary): plain zero fill
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int q = q0 + lane + 32 * c;
        if (some)
          for (int w = 0; w < kEmbWarps - 1; ++w) {
            const float4 p = s_acc[w][lane + 32 * c];
            acc[c].x += p.x; acc[c].y += p.y; acc[c].z += p.z; acc[c].w += p.w;
          }
        if (q < nq) reinterpret_cast<float4*>(dW + (size_t)v * D)[q] = acc[c];
      }
    }
  }
}

extern "C" int l2s_embedding_bwd(const int64_t* idx, const float* dy, float* dW, int R, int V, int D, l2s_stream_t stream) {
  L2S_REQUIRE(R >= 0 && V > 0 && D > 0 && D % 4 == 0, L2S_ERR_SHAPE, "embedding_bwd: bad shape R=%d V=%d D=%d", R, V, D);
  L2S_REQUIRE(dW && (R == 0 || (idx && dy)), L2S_ERR_ARG, "embedding_bwd: null pointer");
  L2S_REQUIRE(aligned16(dW) && (R == 0 || aligned16(dy)), L2S_ERR_ARG, "embedding_bwd: dy / dW must be 16-byte aligned");
  embedding_bwd_kernel<<<V, kEmbWarps * 32, 0, (cudaStream_t)stream>>>(idx, dy, dW, R, D);
  L2S_LAUNCH_OK("embedding_bwd_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" size_t l2s_colsum_workspace_bytes(int R, int C) {
  const size_t chunks = (size_t)((R > 0 ? R : 0) + kColsumRows - 1) / kColsumRows;
  return chunks * (size_t)(C > 0 ? C : 0) * sizeof(float) + 256;
}

extern "C" int l2s_colsum(const float* in, int64_t ld, float* out, int R, int C, void* workspace, size_t workspace_bytes,
                          l2s_stream_t stream) {
  L2S_REQUIRE(R >= 0 && C > 0 && ld >= C, L2S_ERR_SHAPE, "colsum: bad shape R=%d C=%d ld=%lld", R, C, (long long)ld);
  L2S_REQUIRE(out, L2S_ERR_ARG, "colsum: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0) {
    L2S_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, st));
    return L2S_OK;
  }
  L2S_REQUIRE(in, L2S_ERR_ARG, "colsum: null input");
  L2S_REQUIRE(workspace && workspace_bytes >= l2s_colsum_workspace_bytes(R, C), L2S_ERR_WORKSPACE, "colsum: workspace too small");
  if (R <= 1024) {
    colsum_small_kernel<<<(C + 31) / 32, kColsumSmallWarps * 32, 0, st>>>(in, ld, out, R, C);
    L2S_LAUNCH_OK("colsum_small_kernel");
    count_launch();
    return L2S_OK;
  }
  const int chunks = (R + kColsumRows - 1) / kColsumRows;
  float* partial = reinterpret_cast<float*>(workspace);
  colsum_partial_kernel<<<dim3((C + 127) / 128, chunks), 128, 0, st>>>(in, ld, partial, R, C);
  L2S_LAUNCH_OK("colsum_partial_kernel");
  colsum_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, out, chunks, C);
  L2S_LAUNCH_OK("colsum_kernel");
  count_launch(2);
  return L2S_OK;
}

extern "C" size_t l2s_att2in2_decode_workspace_bytes(int T, int B, int A, int D, int Dh) {
  (void)T;
  return decode_ws_layout(B, A, D, Dh, nullptr, nullptr);
}

extern "C" int l2s_att2in2_decode_fwd(float* cat_all, const float* att_feats, const float* p_att, const float* w_cat,
                                      const float* w_a2c, const float* b_a2c, const float* alpha_w,
                                      const float* alpha_b, float* h_all, float* c_all, float* a2c_all,
                                      float* pi_all, float* att_res_all, int T, int B, int A, int D, int Dh,
                                      void* workspace, size_t workspace_bytes, l2s_stream_t stream) {
  L2S_REQUIRE(cat_all && att_feats && p_att && w_cat && w_a2c && b_a2c && alpha_w && alpha_b && h_all && c_all &&
                  a2c_all && pi_all && att_res_all, L2S_ERR_ARG, "att2in2_decode_fwd: null pointer");
  int rc = check_decode(T, B, A, D, Dh);
  if (rc) return rc;
  L2S_REQUIRE(workspace && workspace_bytes >= decode_ws_layout(B, A, D, Dh, nullptr, nullptr), L2S_ERR_WORKSPACE,
              "att2in2_decode_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  DecodeWs w;
  decode_ws_layout(B, A, D, Dh, &w, reinterpret_cast<char*>(workspace));
  // one persistent weight-stationary kernel for all T steps where the shape allows it (decode_persist.cu)
  if (const int cs = decode_persist_cluster(B, A, D, Dh))
    return launch_decode_fwd_persist(cs, cat_all, att_feats, p_att, w_cat, w_a2c, b_a2c, alpha_w, alpha_b, h_all, c_all,
                                     a2c_all, pi_all, att_res_all, T, B, A, w.bar, st);
  L2S_CUDA_OK(cudaMemsetAsync(w.lin, 0, linear_small_counter_bytes(B, Dh + 5 * D), st));   // the widest GEMM of the loop
  const int LC = Dh + 5 * D;
  for (int t = 0; t < T; ++t) {
    float* cat_t = cat_all + (size_t)t * B * LC;
    float* h_t = h_all + (size_t)t * B * D;
    float* c_t = c_all + (size_t)t * B * D;
    float* a2c_t = a2c_all + (size_t)t * B * 2 * D;
    float* pi_t = pi_all + (size_t)t * B * A;
    float* res_t = att_res_all + (size_t)t * B * D;
    if (t > 0) {   // h_{-1} = 0 (AttModel.py:62): the rows keep their initial [b_h2att | i2h(x_0)+b]
      rc = launch_linear_small(h_t - (size_t)B * D, D, w_cat, D, nullptr, cat_t, LC, B, LC, D, 1, w.lin, w.lin_bytes, st);
      if (rc) return rc;
    }
    rc = launch_att_step_fwd(cat_t, LC, att_feats, p_att, alpha_w, alpha_b, pi_t, res_t, B, A, D, Dh, st);
    if (rc) return rc;
    rc = launch_linear_small(res_t, D, w_a2c, D, b_a2c, a2c_t, 2 * D, B, 2 * D, D, 0, w.lin, w.lin_bytes, st);
    if (rc) return rc;
    rc = launch_gates_fwd(cat_t + Dh, LC, a2c_t, t > 0 ? c_t - (size_t)B * D : nullptr, h_t, c_t, B, D, st);
    if (rc) return rc;
  }
  return L2S_OK;
}

extern "C" int l2s_att2in2_decode_bwd(const float* dh_all, const float* cat_all, const float* att_feats,
                                      const float* p_att, const float* w_cat_t, const float* w_a2c_t,
                                      const float* alpha_w, const float* c_all, const float* a2c_all,
                                      const float* pi_all, float* dcat_all, float* da2c_all, float* dres_all,
                                      float* de_all, float* dp_att, float* datt_feats, float* dalpha_w, int T, int B,
                                      int A, int D, int Dh, void* workspace, size_t workspace_bytes,
                                      l2s_stream_t stream) {
  L2S_REQUIRE(dh_all && cat_all && att_feats && p_att && w_cat_t && w_a2c_t && alpha_w && c_all && a2c_all && pi_all &&
                  dcat_all && da2c_all && dres_all && de_all && dp_att && datt_feats && dalpha_w, L2S_ERR_ARG,
              "att2in2_decode_bwd: null pointer");
  int rc = check_decode(T, B, A, D, Dh);
  if (rc) return rc;
  L2S_REQUIRE(workspace && workspace_bytes >= decode_ws_layout(B, A, D, Dh, nullptr, nullptr), L2S_ERR_WORKSPACE,
              "att2in2_decode_bwd: workspace too small");
  L2S_REQUIRE(aligned16(p_att) && aligned16(dp_att) && aligned16(datt_feats) && aligned16(alpha_w), L2S_ERR_ALIGN,
              "att2in2_decode_bwd: feature pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  DecodeWs w;
  decode_ws_layout(B, A, D, Dh, &w, reinterpret_cast<char*>(workspace));
  L2S_CUDA_OK(cudaMemsetAsync(w.lin, 0, linear_small_counter_bytes(B, Dh + 5 * D), st));
  const int LC = Dh + 5 * D;
  float* dh_carry = w.carry;
  float* dc_buf[2] = {w.carry + (size_t)B * D, w.carry + (size_t)2 * B * D};
  const bool persistent = decode_bwd_persist_ok(B, A, D, Dh);
  if (persistent) {      // the whole reverse loop as one persistent weight-stationary kernel (decode_persist.cu)
    rc = launch_decode_bwd_persist(dh_all, cat_all, att_feats, p_att, w_cat_t, w_a2c_t, alpha_w, c_all, a2c_all, pi_all,
                                   dcat_all, da2c_all, dres_all, de_all, dh_carry, T, B, A, w.bar, st);
    if (rc) return rc;
  }
  for (int t = persistent ? -1 : T - 1; t >= 0; --t) {
    const float* cat_t = cat_all + (size_t)t * B * LC;
    const float* c_t = c_all + (size_t)t * B * D;
    const float* a2c_t = a2c_all + (size_t)t * B * 2 * D;
    float* dcat_t = dcat_all + (size_t)t * B * LC;
    float* da2c_t = da2c_all + (size_t)t * B * 2 * D;
    float* dres_t = dres_all + (size_t)t * B * D;
    const bool last = (t == T - 1);
    float* dc_in = dc_buf[t & 1];
    float* dc_out = dc_buf[(t & 1) ^ 1];
    rc = launch_gates_bwd(cat_t + Dh, LC, a2c_t, t > 0 ? c_t - (size_t)B * D : nullptr, c_t,
                          dh_all + (size_t)t * B * D, last ? nullptr : dh_carry, last ? nullptr : dc_in,
                          dcat_t + Dh, LC, da2c_t, dc_out, B, D, st);
    if (rc) return rc;
    // datt_res_t = da2c_t . W_a2c   (w_a2c_t is W_a2c^T stored (D, 2D))
    rc = launch_linear_small(da2c_t, 2 * D, w_a2c_t, 2 * D, nullptr, dres_t, D, B, D, 2 * D, 0, w.lin, w.lin_bytes, st);
    if (rc) return rc;
    rc = launch_att_step_bwd(dres_t, cat_t, LC, att_feats, p_att, alpha_w, pi_all + (size_t)t * B * A, dcat_t, LC,
                             de_all + (size_t)t * B * A, nullptr, nullptr, nullptr, B, A, D, Dh, st);
    if (rc) return rc;
    if (t > 0) {   // dh_{t-1} = [datt_h | dsums]_t . [W_h2att ; W_h2h]   (w_cat_t is its transpose, (D, LC))
      rc = launch_linear_small(dcat_t, LC, w_cat_t, LC, nullptr, dh_carry, D, B, D, LC, 0, w.lin, w.lin_bytes, st);
      if (rc) return rc;
    }
  }
  const int nchunk = (A + kAccLoc - 1) / kAccLoc;
  const size_t smem = ((size_t)T * (Dh + D) + (size_t)2 * T * kAccLoc + Dh) * sizeof(float);
  L2S_REQUIRE(smem <= (size_t)max_smem_optin(), L2S_ERR_SHAPE, "att2in2_decode_bwd: T=%d too long for one pass", T);
  L2S_CUDA_OK(cudaFuncSetAttribute(att_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  L2S_CUDA_OK(launch_chain(att_accum_kernel, dim3(nchunk, B), dim3(kAccThreads), smem, st, p_att, cat_all, LC, dres_all,
                           de_all, pi_all, alpha_w, dp_att, datt_feats, w.dalpha_part, T, B, A, D, Dh));
  if (nchunk * B <= 4096) {     // 672 partial rows at cfg-2: warps in parallel over the rows (the one-thread-per-column chain took 19 us)
    colsum_small_kernel<<<(Dh + 31) / 32, kColsumSmallWarps * 32, 0, st>>>((const float*)w.dalpha_part, (int64_t)Dh, dalpha_w,
                                                                         nchunk * B, Dh);
    L2S_LAUNCH_OK("colsum_small_kernel");
  } else {
    L2S_CUDA_OK(launch_chain(colsum_kernel, dim3((Dh + 127) / 128), dim3(128), 0, st, (const float*)w.dalpha_part, dalpha_w,
                             nchunk * B, Dh));
  }
  count_launch(2);
  return L2S_OK;
}
