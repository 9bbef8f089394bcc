// Spatial dynamic-filter response layer, forward and backward, for sm_100a.
//
// Semantics (reference pyutils/mask-faster-rcnn/lib/nets/
// network_cycle_response.py:534-570 and the response loss :415-422):
//   r_k[p] = M_k[p] * sum_c f_k[c] X[c,p]      k = 0..6, M_k = spatial partition masks
//   r[p]   = sum_k w_k r_k[p]                   (pre-sigmoid response, kept)
//   Y[c,p] = X[c,p] * sigmoid(r[p])             (linear gate variant: Y = X * r)
// The reference materialises 6 masked copies of X and runs 8 cuDNN 1x1 convs per expression;
// here one CTA keeps a [C x TP] pixel tile of X resident in shared memory, evaluates all
// expressions of that image against it and writes each gated tile once: X is read once per
// image and Y written once per expression (algorithmic bytes 4*C*HW*(I+E)).
//
// The contraction has M = 7 rows (x expressions of the image): ~3.5 flop/byte, far below the
// tensor-pipe ridge, and fp32 parity at 1e-4 rules out single-pass TF32, so the dot products run
// on the FMA pipe at < 20% utilisation while the kernel stays HBM-bound.
//
// Backward (given dY, optional explicit dr and the response-loss term):
//   ds[p] = sum_c dY[c,p] X[c,p] ; dr = ds*sig'(r) + dr_ext
//   dX[c,p] = sum_e ( dY_e[c,p]*gate_e[p] + dr_e[p] * sum_k w_k M_k[p] f_k[c] )
//   df_k[c] = w_k sum_p M_k[p] dr[p] X[c,p] ; dw_k = sum_p dr[p] r_k[p]
// dY is streamed exactly once; df needs a reduction over ALL pixels of an image and is done by
// a second kernel that re-reads X (deterministic, no atomics).
#include "common.cuh"

namespace l2s {
namespace {

constexpr int kThreads = 256;
constexpr int NF = L2S_NUM_FILTERS;

struct DfGeom {
  int I, E, C, H, W, HW;
  int h2, h4, h34, w2, w4, w34;   // int(H/2), int(H/4), int(H*3/4), ...
  int linear;
};

__device__ __forceinline__ float mask_k(const DfGeom& g, int k, int p) {
  const int y = p / g.W, x = p - y * g.W;
  bool m;
  switch (k) {
    case 0: m = true; break;
    case 1: m = y < g.h2; break;
    case 2: m = y >= g.h2; break;
    case 3: m = x < g.w2; break;
    case 4: m = x >= g.w2; break;
    case 5: m = y >= g.h4 && y < g.h34; break;
    default: m = x >= g.w4 && x < g.w34; break;
  }
  return m ? 1.f : 0.f;
}

// floats of the [7][C] filter block in shared memory, rounded so that the arrays carved out behind it stay 16-byte aligned
// for any C (they are read as float4 on the vector paths)
__host__ __device__ __forceinline__ size_t fs_floats(int C) { return ((size_t)NF * C + 3) & ~(size_t)3; }

// expressions of image i: expr2img is non-decreasing
__device__ __forceinline__ void expr_range(const int* __restrict__ e2i, int E, int i, int* e0, int* e1) {
  int lo = 0, hi = E;
  while (lo < hi) { int m = (lo + hi) >> 1; if (__ldg(e2i + m) < i) lo = m + 1; else hi = m; }
  *e0 = lo;
  hi = E;
  while (lo < hi) { int m = (lo + hi) >> 1; if (__ldg(e2i + m) <= i) lo = m + 1; else hi = m; }
  *e1 = lo;
}

__device__ __forceinline__ float bce_logits(float x, float t) {
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}

// Four consecutive pixels of a map row at an 8-byte aligned address.  H*W only has to be EVEN for the vector paths (600 x
// 1000 inputs give 38x63 / 37x62 maps: H*W % 4 == 2, so a row starts 16- or 8-byte aligned alternately): one 16-byte
// access where the address allows it, two 8-byte accesses otherwise; `n` = pixels of the quad inside the row (4, 2 or <= 0).
template <bool STREAM>
__device__ __forceinline__ float4 ld_quad(const float* p, int n) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n >= 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    v = STREAM ? __ldcs(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p));
  } else if (n >= 2) {
    const float2 a = STREAM ? __ldcs(reinterpret_cast<const float2*>(p)) : __ldg(reinterpret_cast<const float2*>(p));
    v.x = a.x; v.y = a.y;
    if (n >= 4) {
      const float2 b = STREAM ? __ldcs(reinterpret_cast<const float2*>(p) + 1) : __ldg(reinterpret_cast<const float2*>(p) + 1);
      v.z = b.x; v.w = b.y;
    }
  }
  return v;
}
template <bool STREAM>
__device__ __forceinline__ void st_quad(float* p, float4 v, int n) {
  if (n >= 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    if (STREAM) __stcs(reinterpret_cast<float4*>(p), v); else *reinterpret_cast<float4*>(p) = v;
  } else if (n >= 2) {
    if (STREAM) __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); else *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
    if (n >= 4) {
      if (STREAM) __stcs(reinterpret_cast<float2*>(p) + 1, make_float2(v.z, v.w));
      else *(reinterpret_cast<float2*>(p) + 1) = make_float2(v.z, v.w);
    }
  }
}

// load a [C x TP] tile of a (C,HW) map into smem (row stride TP), zero padded
template <int TP, bool VEC>
__device__ __forceinline__ void load_tile(float* __restrict__ xs, const float* __restrict__ src, int C, int HW,
                                          int p0) {
  const int t = threadIdx.x;
  if (VEC) {
    constexpr int Q = TP / 4;
    const int pq = t % Q, cs = t / Q;
    constexpr int CSTEP = kThreads / Q;
    const int p = p0 + 4 * pq;
    for (int c = cs; c < C; c += CSTEP) {
      reinterpret_cast<float4*>(xs)[c * Q + pq] = ld_quad<false>(src + (size_t)c * HW + p, HW - p);
    }
  } else {
    const int pp = t % TP, cs = t / TP;
    constexpr int CSTEP = kThreads / TP;
    const int p = p0 + pp;
    for (int c = cs; c < C; c += CSTEP) xs[c * TP + pp] = (p < HW) ? __ldg(src + (size_t)c * HW + p) : 0.f;
  }
}

// block reduction of per-thread partial sums acc[NV][4] over the channel-split threads.
// thread layout: pq = t % Q (pixel quad), cs = t / Q.  result: out[v][TP] in smem.
template <int TP, int NV>
__device__ __forceinline__ void reduce_quads(float (&acc)[NV][4], float* __restrict__ red /*[8][NV][TP]*/,
                                             float* __restrict__ out /*[NV][TP]*/) {
  constexpr int Q = TP / 4;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = acc[v][j];
#pragma unroll
      for (int o = Q; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      acc[v][j] = s;
    }
  if (lane < Q) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(wid * NV + v) * TP + 4 * lane + j] = acc[v][j];
  }
  __syncthreads();
  for (int i = t; i < NV * TP; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w * NV * TP + i];
    out[i] = s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------- forward
template <int TP, bool VEC>
__global__ void __launch_bounds__(kThreads)
dynfilter_fwd_kernel(const float* __restrict__ X, const float* __restrict__ filt, const float* __restrict__ fuse,
                     const int* __restrict__ e2i, float* __restrict__ response, float* __restrict__ rk_saved,
                     float* __restrict__ Y, const float* __restrict__ target, float* __restrict__ loss, DfGeom g) {
  constexpr int Q = TP / 4;
  constexpr int CSTEP = kThreads / Q;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                        // [C][TP]
  float* fs = xs + (size_t)g.C * TP;       // [7][C]
  float* red = fs + fs_floats(g.C);      // [8][7][TP]
  float* dk = red + 8 * NF * TP;           // [7][TP]
  float* gate = dk + NF * TP;              // [TP]

  const int i = blockIdx.y, p0 = blockIdx.x * TP, t = threadIdx.x;
  const int pq = t % Q, cs = t / Q;
  int e0, e1;
  expr_range(e2i, g.E, i, &e0, &e1);
  if (e0 == e1) return;

  load_tile<TP, VEC>(xs, X + (size_t)i * g.C * g.HW, g.C, g.HW, p0);
  __syncthreads();

  for (int e = e0; e < e1; ++e) {
    for (int idx = t; idx < NF * g.C; idx += kThreads) fs[idx] = __ldg(filt + (size_t)e * NF * g.C + idx);
    __syncthreads();
    float acc[NF][4];
#pragma unroll
    for (int k = 0; k < NF; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
    for (int c = cs; c < g.C; c += CSTEP) {
      const float4 x = reinterpret_cast<const float4*>(xs)[c * Q + pq];
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const float f = fs[k * g.C + c];
        acc[k][0] = fmaf(f, x.x, acc[k][0]);
        acc[k][1] = fmaf(f, x.y, acc[k][1]);
        acc[k][2] = fmaf(f, x.z, acc[k][2]);
        acc[k][3] = fmaf(f, x.w, acc[k][3]);
      }
    }
    reduce_quads<TP, NF>(acc, red, dk);
    if (t < TP) {
      const int p = p0 + t;
      float r = 0.f;
      if (p < g.HW) {
#pragma unroll
        for (int k = 0; k < NF; ++k) {
          const float rk = mask_k(g, k, p) * dk[k * TP + t];
          if (rk_saved) rk_saved[((size_t)e * NF + k) * g.HW + p] = rk;
          r = fmaf(__ldg(fuse + e * NF + k), rk, r);
        }
        response[(size_t)e * g.HW + p] = r;
      }
      gate[t] = g.linear ? r : sigmoidf_acc(r);
      if (loss != nullptr && target != nullptr) {
        float l = (p < g.HW) ? bce_logits(r, __ldg(target + (size_t)e * g.HW + p)) : 0.f;
        constexpr unsigned kTileMask = TP >= 32 ? 0xffffffffu : ((1u << TP) - 1u);
#pragma unroll
        for (int o = TP / 2; o > 0; o >>= 1) l += __shfl_xor_sync(kTileMask, l, o, TP);
        if (t == 0) atomicAdd(loss + e, l / (float)g.HW);
      }
    }
    __syncthreads();
    float* Ye = Y + (size_t)e * g.C * g.HW;
    if (VEC) {
      const float4 gq = reinterpret_cast<const float4*>(gate)[pq];
      const int p = p0 + 4 * pq;
      if (p < g.HW)
        for (int c = cs; c < g.C; c += CSTEP) {
          float4 x = reinterpret_cast<const float4*>(xs)[c * Q + pq];
          x.x *= gq.x; x.y *= gq.y; x.z *= gq.z; x.w *= gq.w;
          st_quad<true>(Ye + (size_t)c * g.HW + p, x, g.HW - p);
        }
    } else {
      const int pp = t % TP, c1 = t / TP;
      const int p = p0 + pp;
      if (p < g.HW)
        for (int c = c1; c < g.C; c += kThreads / TP) Ye[(size_t)c * g.HW + p] = xs[c * TP + pp] * gate[pp];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------- backward (A)
template <int TP, bool VEC>
__global__ void __launch_bounds__(kThreads)
dynfilter_bwd_kernel(const float* __restrict__ X, const float* __restrict__ filt, const float* __restrict__ fuse,
                     const int* __restrict__ e2i, const float* __restrict__ response,
                     const float* __restrict__ dY, const float* __restrict__ dresp,
                     const float* __restrict__ target, const float* __restrict__ gscale,
                     float* __restrict__ dX, float* __restrict__ drbuf, DfGeom g) {
  constexpr int Q = TP / 4;
  constexpr int CSTEP = kThreads / Q;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                         // [C][TP]
  float* dxs = xs + (size_t)g.C * TP;       // [C][TP]
  float* fs = dxs + (size_t)g.C * TP;       // [7][C]
  float* red = fs + fs_floats(g.C);       // [8][1][TP]
  float* ds = red + 8 * TP;                 // [TP]
  float* gate = ds + TP;                    // [TP]
  float* mwdr = gate + TP;                  // [7][TP]  w_k * M_k[p] * dr[p]

  const int i = blockIdx.y, p0 = blockIdx.x * TP, t = threadIdx.x;
  const int pq = t % Q, cs = t / Q;
  int e0, e1;
  expr_range(e2i, g.E, i, &e0, &e1);

  load_tile<TP, VEC>(xs, X + (size_t)i * g.C * g.HW, g.C, g.HW, p0);
  for (int idx = t; idx < g.C * TP; idx += kThreads) dxs[idx] = 0.f;
  __syncthreads();

  for (int e = e0; e < e1; ++e) {
    for (int idx = t; idx < NF * g.C; idx += kThreads) fs[idx] = __ldg(filt + (size_t)e * NF * g.C + idx);
    if (t < TP) {
      const int p = p0 + t;
      const float r = (p < g.HW) ? __ldg(response + (size_t)e * g.HW + p) : 0.f;
      gate[t] = g.linear ? r : sigmoidf_acc(r);
    }
    __syncthreads();
    // pass 1: stream dY once; ds partials and the gate term of dX
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    const float* dYe = dY + (size_t)e * g.C * g.HW;
    if (VEC) {
      const float4 gq = reinterpret_cast<const float4*>(gate)[pq];
      const int p = p0 + 4 * pq;
      if (p < g.HW)
        for (int c = cs; c < g.C; c += CSTEP) {
          const float4 v = ld_quad<true>(dYe + (size_t)c * g.HW + p, g.HW - p);
          const float4 x = reinterpret_cast<const float4*>(xs)[c * Q + pq];
          acc[0][0] = fmaf(v.x, x.x, acc[0][0]);
          acc[0][1] = fmaf(v.y, x.y, acc[0][1]);
          acc[0][2] = fmaf(v.z, x.z, acc[0][2]);
          acc[0][3] = fmaf(v.w, x.w, acc[0][3]);
          float4 d = reinterpret_cast<float4*>(dxs)[c * Q + pq];
          d.x = fmaf(v.x, gq.x, d.x); d.y = fmaf(v.y, gq.y, d.y);
          d.z = fmaf(v.z, gq.z, d.z); d.w = fmaf(v.w, gq.w, d.w);
          reinterpret_cast<float4*>(dxs)[c * Q + pq] = d;
        }
    } else {
      // scalar path: thread (cs, pq) still owns pixels 4pq..4pq+3 of its channels
      for (int c = cs; c < g.C; c += CSTEP)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pp = 4 * pq + j, p = p0 + pp;
          if (p < g.HW) {
            const float v = __ldg(dYe + (size_t)c * g.HW + p);
            acc[0][j] = fmaf(v, xs[c * TP + pp], acc[0][j]);
            dxs[c * TP + pp] = fmaf(v, gate[pp], dxs[c * TP + pp]);
          }
        }
    }
    reduce_quads<TP, 1>(acc, red, ds);
    if (t < TP) {
      const int p = p0 + t;
      float dr = 0.f;
      if (p < g.HW) {
        const float r = __ldg(response + (size_t)e * g.HW + p);
        const float sg = sigmoidf_acc(r);
        dr = g.linear ? ds[t] : ds[t] * sg * (1.f - sg);
        if (dresp) dr += __ldg(dresp + (size_t)e * g.HW + p);
        if (target && gscale) dr += __ldg(gscale + e) * (sg - __ldg(target + (size_t)e * g.HW + p)) / (float)g.HW;
        drbuf[(size_t)e * g.HW + p] = dr;
      }
#pragma unroll
      for (int k = 0; k < NF; ++k)
        mwdr[k * TP + t] = (p < g.HW) ? __ldg(fuse + e * NF + k) * mask_k(g, k, p) * dr : 0.f;
    }
    __syncthreads();
    // pass 2: dX += dr[p] * sum_k w_k M_k[p] f_k[c]
    for (int c = cs; c < g.C; c += CSTEP) {
      float4 d = reinterpret_cast<float4*>(dxs)[c * Q + pq];
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const float f = fs[k * g.C + c];
        const float4 m = reinterpret_cast<const float4*>(mwdr)[k * Q + pq];
        d.x = fmaf(f, m.x, d.x); d.y = fmaf(f, m.y, d.y); d.z = fmaf(f, m.z, d.z); d.w = fmaf(f, m.w, d.w);
      }
      reinterpret_cast<float4*>(dxs)[c * Q + pq] = d;
    }
    __syncthreads();
  }
  float* dXi = dX + (size_t)i * g.C * g.HW;
  if (VEC) {
    const int p = p0 + 4 * pq;
    if (p < g.HW)
      for (int c = cs; c < g.C; c += CSTEP)
        st_quad<false>(dXi + (size_t)c * g.HW + p, reinterpret_cast<const float4*>(dxs)[c * Q + pq], g.HW - p);
  } else {
    const int pp = t % TP, c1 = t / TP;
    const int p = p0 + pp;
    if (p < g.HW)
      for (int c = c1; c < g.C; c += kThreads / TP) dXi[(size_t)c * g.HW + p] = dxs[c * TP + pp];
  }
}

// ------------------------------------------------------------------------------- backward (A'), register accumulators
// Same arithmetic as dynfilter_bwd_kernel for the common case C <= 16 * (256 / (TP/4)) and float4-aligned maps:
// the dX tile lives in registers (thread (pixel quad, channel slot) owns <= 16 channels x 4 pixels) instead of a
// second [C x TP] shared-memory tile, which halves the shared-memory footprint (2 CTAs per SM instead of 1) and
// removes the read-modify-write traffic of the accumulation.
constexpr int kRegCh = 16;
template <int TP>
__global__ void __launch_bounds__(kThreads, 2)
dynfilter_bwd_reg_kernel(const float* __restrict__ X, const float* __restrict__ filt, const float* __restrict__ fuse,
                         const int* __restrict__ e2i, const float* __restrict__ response,
                         const float* __restrict__ dY, const float* __restrict__ dresp,
                         const float* __restrict__ target, const float* __restrict__ gscale,
                         float* __restrict__ dX, float* __restrict__ drbuf, DfGeom g) {
  constexpr int Q = TP / 4;
  constexpr int CSTEP = kThreads / Q;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                         // [C][TP]
  float* fs = xs + (size_t)g.C * TP;        // [7][C]
  float* red = fs + fs_floats(g.C);       // [8][1][TP]
  float* ds = red + 8 * TP;                 // [TP]
  float* gate = ds + TP;                    // [TP]
  float* mwdr = gate + TP;                  // [7][TP]  w_k * M_k[p] * dr[p]

  const int i = blockIdx.y, p0 = blockIdx.x * TP, t = threadIdx.x;
  const int pq = t % Q, cs = t / Q;
  int e0, e1;
  expr_range(e2i, g.E, i, &e0, &e1);

  load_tile<TP, true>(xs, X + (size_t)i * g.C * g.HW, g.C, g.HW, p0);
  float4 dacc[kRegCh];
#pragma unroll
  for (int j = 0; j < kRegCh; ++j) dacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int p = p0 + 4 * pq;
  const bool pin = p < g.HW;
  __syncthreads();

  for (int e = e0; e < e1; ++e) {
    for (int idx = t; idx < NF * g.C; idx += kThreads) fs[idx] = __ldg(filt + (size_t)e * NF * g.C + idx);
    if (t < TP) {
      const int pp = p0 + t;
      const float r = (pp < g.HW) ? __ldg(response + (size_t)e * g.HW + pp) : 0.f;
      gate[t] = g.linear ? r : sigmoidf_acc(r);
    }
    __syncthreads();
    // pass 1: stream dY once; ds partials and the gate term of dX
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    const float* dYe = dY + (size_t)e * g.C * g.HW;
    const float4 gq = reinterpret_cast<const float4*>(gate)[pq];
    if (pin) {
#pragma unroll
      for (int h = 0; h < kRegCh; h += 8) {      // 8 streaming loads of the thread in flight at once
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = cs + (h + j) * CSTEP;
          v[j] = (c < g.C) ? ld_quad<true>(dYe + (size_t)c * g.HW + p, g.HW - p) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = cs + (h + j) * CSTEP;
          if (c < g.C) {
            const float4 x = reinterpret_cast<const float4*>(xs)[c * Q + pq];
            acc[0][0] = fmaf(v[j].x, x.x, acc[0][0]);
            acc[0][1] = fmaf(v[j].y, x.y, acc[0][1]);
            acc[0][2] = fmaf(v[j].z, x.z, acc[0][2]);
            acc[0][3] = fmaf(v[j].w, x.w, acc[0][3]);
            float4& d = dacc[h + j];
            d.x = fmaf(v[j].x, gq.x, d.x); d.y = fmaf(v[j].y, gq.y, d.y);
            d.z = fmaf(v[j].z, gq.z, d.z); d.w = fmaf(v[j].w, gq.w, d.w);
          }
        }
      }
    }
    reduce_quads<TP, 1>(acc, red, ds);
    if (t < TP) {
      const int pp = p0 + t;
      float dr = 0.f;
      if (pp < g.HW) {
        const float r = __ldg(response + (size_t)e * g.HW + pp);
        const float sg = sigmoidf_acc(r);
        dr = g.linear ? ds[t] : ds[t] * sg * (1.f - sg);
        if (dresp) dr += __ldg(dresp + (size_t)e * g.HW + pp);
        if (target && gscale) dr += __ldg(gscale + e) * (sg - __ldg(target + (size_t)e * g.HW + pp)) / (float)g.HW;
        drbuf[(size_t)e * g.HW + pp] = dr;
      }
#pragma unroll
      for (int k = 0; k < NF; ++k)
        mwdr[k * TP + t] = (pp < g.HW) ? __ldg(fuse + e * NF + k) * mask_k(g, k, pp) * dr : 0.f;
    }
    __syncthreads();
    // pass 2: dX += dr[p] * sum_k w_k M_k[p] f_k[c]
    float4 m[NF];
#pragma unroll
    for (int k = 0; k < NF; ++k) m[k] = reinterpret_cast<const float4*>(mwdr)[k * Q + pq];
#pragma unroll
    for (int j = 0; j < kRegCh; ++j) {
      const int c = cs + j * CSTEP;
      if (c < g.C) {
#pragma unroll
        for (int k = 0; k < NF; ++k) {
          const float f = fs[k * g.C + c];
          dacc[j].x = fmaf(f, m[k].x, dacc[j].x); dacc[j].y = fmaf(f, m[k].y, dacc[j].y);
          dacc[j].z = fmaf(f, m[k].z, dacc[j].z); dacc[j].w = fmaf(f, m[k].w, dacc[j].w);
        }
      }
    }
    __syncthreads();
  }
  float* dXi = dX + (size_t)i * g.C * g.HW;
  if (pin) {
#pragma unroll
    for (int j = 0; j < kRegCh; ++j) {
      const int c = cs + j * CSTEP;
      if (c < g.C) st_quad<false>(dXi + (size_t)c * g.HW + p, dacc[j], g.HW - p);
    }
  }
}

// ------------------------------------------------------------------------------- backward (B)
// df[e,k,c] = w_k * sum_p M_k[p] dr[e,p] X[img(e),c,p].  CTA = (image, 8 channels); warp = channel;
// lanes stride over pixels; expressions of the image in chunks of EB.
constexpr int EB = 4;
__global__ void __launch_bounds__(256)
dynfilter_dfilt_kernel(const float* __restrict__ X, const float* __restrict__ fuse, const int* __restrict__ e2i,
                       const float* __restrict__ drbuf, float* __restrict__ dfilt, DfGeom g) {
  extern __shared__ __align__(16) float smem[];   // [EB][HW] dr
  const int i = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + wid;
  int e0, e1;
  expr_range(e2i, g.E, i, &e0, &e1);
  const float* Xc = X + ((size_t)i * g.C + min(c, g.C - 1)) * g.HW;
  for (int eb = e0; eb < e1; eb += EB) {
    const int ne = min(EB, e1 - eb);
    __syncthreads();
    for (int idx = threadIdx.x; idx < ne * g.HW; idx += blockDim.x) smem[idx] = __ldg(drbuf + (size_t)eb * g.HW + idx);
    __syncthreads();
    float acc[EB][NF];
#pragma unroll
    for (int a = 0; a < EB; ++a)
#pragma unroll
      for (int k = 0; k < NF; ++k) acc[a][k] = 0.f;
    for (int p = lane; p < g.HW; p += 32) {
      const float x = __ldg(Xc + p);
      float m[NF];
#pragma unroll
      for (int k = 0; k < NF; ++k) m[k] = mask_k(g, k, p);
#pragma unroll
      for (int a = 0; a < EB; ++a) {
        if (a < ne) {
          const float v = smem[a * g.HW + p] * x;
#pragma unroll
          for (int k = 0; k < NF; ++k) acc[a][k] = fmaf(m[k], v, acc[a][k]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < EB; ++a)
#pragma unroll
      for (int k = 0; k < NF; ++k) acc[a][k] = warp_sum(acc[a][k]);
    if (lane == 0 && c < g.C) {
#pragma unroll
      for (int a = 0; a < EB; ++a)
        if (a < ne) {
#pragma unroll
          for (int k = 0; k < NF; ++k)
            dfilt[((size_t)(eb + a) * NF + k) * g.C + c] = __ldg(fuse + (eb + a) * NF + k) * acc[a][k];
        }
    }
  }
}

// Vectorised variant of the above: warp = channel row, every lane issues its eight 16-byte loads of X up front (the
// first version walked the row with one dependent 4-byte load per iteration and took as long as the whole forward), the
// 7 partition masks of a pixel come from a one-byte bit table in shared memory (built once per CTA: the only place with
// a division) and are shared by the expressions of the chunk.  A16 = rows are 16-byte aligned (H*W % 4 == 0).  Otherwise
// H*W is even (600 x 1000 inputs: 38x63 and 37x62 maps) and a row starts 16- or only 8-byte aligned depending on the
// parity of its index: the float4 body then starts at pixel `head` (0 or 2) of the row, the two pixels left over at the
// other end are taken by two lanes, and the dr rows / mask bits -- whose alignment relative to the body differs from
// warp to warp -- are read in 8-byte / 2-byte halves.
constexpr int EBV = 3;

__device__ __forceinline__ void dfilt_accum(float (&acc)[NF], float v, unsigned bits) {
  acc[0] += v;
#pragma unroll
  for (int k = 1; k < NF; ++k) acc[k] = fmaf(((bits >> k) & 1u) ? 1.f : 0.f, v, acc[k]);
}

template <bool A16>
__global__ void __launch_bounds__(256)
dynfilter_dfilt_vec_kernel(const float* __restrict__ X, const float* __restrict__ fuse, const int* __restrict__ e2i,
                           const float* __restrict__ drbuf, float* __restrict__ dfilt, DfGeom g) {
  extern __shared__ __align__(16) float smem[];   // [EBV][HW] dr, then [HW] mask bits (u8)
  uint8_t* mbits = reinterpret_cast<uint8_t*>(smem + (size_t)EBV * g.HW);
  const int i = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + wid;
  const size_t rowoff = ((size_t)i * g.C + min(c, g.C - 1)) * g.HW;
  const int head = A16 ? 0 : (int)((4 - (rowoff & 3)) & 3);            // 0 or 2: pixels before the 16-byte aligned body
  const float* Xrow = X + rowoff;
  const float4* Xc = reinterpret_cast<const float4*>(Xrow + head);
  const int nq = (g.HW - head) >> 2;                                    // float4 of the body
  const int rest = head ? 0 : nq * 4;                                   // first of the (HW - 4 nq) left-over pixels
  const int nrest = g.HW - 4 * nq;                                      // 0 (A16) or 2
  // The first eight float4 of the row are requested before anything else: the expression-range search, the mask table
  // and the staging of dr (three dependent L2 round trips on a CTA that lives for a few microseconds) then overlap the
  // DRAM latency of X instead of preceding it.  X does not depend on the expression chunk, so the batch is reused.
  float4 xpre[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int q = u * 32 + lane;
    xpre[u] = q < nq ? __ldg(Xc + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int e0, e1;
  expr_range(e2i, g.E, i, &e0, &e1);
  if (e0 == e1) return;
  for (int p = threadIdx.x; p < g.HW; p += blockDim.x) {
    unsigned b = 0;
#pragma unroll
    for (int k = 0; k < NF; ++k) b |= (mask_k(g, k, p) != 0.f ? 1u : 0u) << k;
    mbits[p] = (uint8_t)b;
  }
  for (int eb = e0; eb < e1; eb += EBV) {
    const int ne = min(EBV, e1 - eb);
    __syncthreads();
    if (A16) {
      for (int idx = threadIdx.x; idx < ne * (g.HW >> 2); idx += blockDim.x)
        reinterpret_cast<float4*>(smem)[idx] = __ldg(reinterpret_cast<const float4*>(drbuf + (size_t)eb * g.HW) + idx);
    } else {
      for (int idx = threadIdx.x; idx < ne * (g.HW >> 1); idx += blockDim.x)
        reinterpret_cast<float2*>(smem)[idx] = __ldg(reinterpret_cast<const float2*>(drbuf + (size_t)eb * g.HW) + idx);
    }
    __syncthreads();
    float acc[EBV][NF];
#pragma unroll
    for (int a = 0; a < EBV; ++a)
#pragma unroll
      for (int k = 0; k < NF; ++k) acc[a][k] = 0.f;
    if (!A16 && lane < nrest) {                                         // the two pixels outside the aligned body
      const int p = rest + lane;
      const float x = __ldg(Xrow + p);
      const unsigned bits = mbits[p];
#pragma unroll
      for (int a = 0; a < EBV; ++a)
        if (a < ne) dfilt_accum(acc[a], smem[(size_t)a * g.HW + p] * x, bits);
    }
    for (int q0 = 0; q0 < nq; q0 += 32 * 8) {
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int q = q0 + u * 32 + lane;
        x[u] = q0 == 0 ? xpre[u] : (q < nq ? __ldg(Xc + q) : make_float4(0.f, 0.f, 0.f, 0.f));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int q = q0 + u * 32 + lane;
        if (q < nq) {
          const int p = head + 4 * q;
          unsigned bits;
          if (A16) bits = reinterpret_cast<const unsigned*>(mbits)[q];
          else bits = (unsigned)*reinterpret_cast<const uint16_t*>(mbits + p) |
                      ((unsigned)*reinterpret_cast<const uint16_t*>(mbits + p + 2) << 16);
          const float xv[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
          // the masks of the four pixels as 0/1 floats, once for all expressions of the chunk (k = 0 is all ones)
          float mf[4][NF - 1];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 1; k < NF; ++k) mf[j][k - 1] = ((bits >> (8 * j + k)) & 1u) ? 1.f : 0.f;
#pragma unroll
          for (int a = 0; a < EBV; ++a) {
            if (a < ne) {
              float4 d;
              if (A16) {
                d = reinterpret_cast<const float4*>(smem + (size_t)a * g.HW)[q];
              } else {
                const float2 d0 = *reinterpret_cast<const float2*>(smem + (size_t)a * g.HW + p);
                const float2 d1 = *reinterpret_cast<const float2*>(smem + (size_t)a * g.HW + p + 2);
                d = make_float4(d0.x, d0.y, d1.x, d1.y);
              }
              const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float v = dv[j] * xv[j];
                acc[a][0] += v;
#pragma unroll
                for (int k = 1; k < NF; ++k) acc[a][k] = fmaf(mf[j][k - 1], v, acc[a][k]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < EBV; ++a)
#pragma unroll
      for (int k = 0; k < NF; ++k) acc[a][k] = warp_sum(acc[a][k]);
    if (lane == 0 && c < g.C) {
#pragma unroll
      for (int a = 0; a < EBV; ++a)
        if (a < ne) {
#pragma unroll
          for (int k = 0; k < NF; ++k)
            dfilt[((size_t)(eb + a) * NF + k) * g.C + c] = __ldg(fuse + (eb + a) * NF + k) * acc[a][k];
        }
    }
  }
}

// dw[e,k] = sum_p dr[e,p] * r_k[e,p]
__global__ void dynfilter_dfuse_kernel(const float* __restrict__ drbuf, const float* __restrict__ rk,
                                       float* __restrict__ dfuse, int HW) {
  const int e = blockIdx.x;
  float acc[NF];
#pragma unroll
  for (int k = 0; k < NF; ++k) acc[k] = 0.f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const float d = __ldg(drbuf + (size_t)e * HW + p);
#pragma unroll
    for (int k = 0; k < NF; ++k) acc[k] = fmaf(d, __ldg(rk + ((size_t)e * NF + k) * HW + p), acc[k]);
  }
  __shared__ float s[8][NF];
#pragma unroll
  for (int k = 0; k < NF; ++k) acc[k] = warp_sum(acc[k]);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < NF; ++k) s[threadIdx.x >> 5][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < NF) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += s[w][threadIdx.x];
    dfuse[e * NF + threadIdx.x] = v;
  }
}

// r_k recomputation for the backward when the forward did not save it: r_k = M_k * f_k . X
__global__ void __launch_bounds__(256)
dynfilter_rk_kernel(const float* __restrict__ X, const float* __restrict__ filt, const int* __restrict__ e2i,
                    float* __restrict__ rk, DfGeom g) {
  const int e = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.HW) return;
  const float* Xi = X + (size_t)__ldg(e2i + e) * g.C * g.HW;
  float acc[NF];
#pragma unroll
  for (int k = 0; k < NF; ++k) acc[k] = 0.f;
  for (int c = 0; c < g.C; ++c) {
    const float x = __ldg(Xi + (size_t)c * g.HW + p);
#pragma unroll
    for (int k = 0; k < NF; ++k) acc[k] = fmaf(__ldg(filt + ((size_t)e * NF + k) * g.C + c), x, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < NF; ++k) rk[((size_t)e * NF + k) * g.HW + p] = mask_k(g, k, p) * acc[k];
}

DfGeom make_geom(int I, int E, int C, int H, int W, int flags) {
  DfGeom g;
  g.I = I; g.E = E; g.C = C; g.H = H; g.W = W; g.HW = H * W;
  g.h2 = H / 2; g.h4 = H / 4; g.h34 = (H * 3) / 4;
  g.w2 = W / 2; g.w4 = W / 4; g.w34 = (W * 3) / 4;
  g.linear = (flags & L2S_GATE_LINEAR) ? 1 : 0;
  return g;
}

size_t fwd_smem(int C, int TP) { return ((size_t)C * TP + fs_floats(C) + 8 * NF * TP + NF * TP + TP) * 4; }
size_t bwd_reg_smem(int C, int TP) { return ((size_t)C * TP + fs_floats(C) + 8 * TP + 2 * TP + NF * TP) * 4; }
size_t bwd_smem(int C, int TP) { return ((size_t)2 * C * TP + fs_floats(C) + 8 * TP + 2 * TP + NF * TP) * 4; }

template <int TP, bool VEC>
int launch_fwd(const float* X, const float* filt, const float* fuse, const int* e2i, float* response, float* rk,
               float* Y, const float* target, float* loss, const DfGeom& g, cudaStream_t st) {
  auto kern = dynfilter_fwd_kernel<TP, VEC>;
  const size_t smem = fwd_smem(g.C, TP);
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.HW + TP - 1) / TP, g.I);
  kern<<<grid, kThreads, smem, st>>>(X, filt, fuse, e2i, response, rk, Y, target, loss, g);
  L2S_LAUNCH_OK("dynfilter_fwd_kernel");
  count_launch();
  return L2S_OK;
}

template <int TP, bool VEC>
int launch_bwd(const float* X, const float* filt, const float* fuse, const int* e2i, const float* response,
               const float* dY, const float* dresp, const float* target, const float* gscale, float* dX,
               float* drbuf, const DfGeom& g, cudaStream_t st) {
  auto kern = dynfilter_bwd_kernel<TP, VEC>;
  const size_t smem = bwd_smem(g.C, TP);
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.HW + TP - 1) / TP, g.I);
  kern<<<grid, kThreads, smem, st>>>(X, filt, fuse, e2i, response, dY, dresp, target, gscale, dX, drbuf, g);
  L2S_LAUNCH_OK("dynfilter_bwd_kernel");
  count_launch();
  return L2S_OK;
}

int check(const void* X, const void* f, const void* w, const void* e2i, int I, int E, int C, int H, int W, int flags) {
  L2S_REQUIRE(X && f && w && e2i, L2S_ERR_ARG, "dynfilter: null pointer");
  L2S_REQUIRE(I > 0 && E >= 0 && C > 0 && H > 0 && W > 0, L2S_ERR_SHAPE, "dynfilter: bad shape I=%d E=%d C=%d H=%d W=%d", I, E, C, H, W);
  L2S_REQUIRE((flags & ~L2S_GATE_LINEAR) == 0, L2S_ERR_ARG, "dynfilter: unknown flags %d", flags);
  return L2S_OK;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" size_t l2s_dynfilter_fwd_workspace_bytes(int I, int E, int C, int H, int W) {
  (void)I;
  return dynfilter_tc_workspace_bytes(E > 0 ? E : 0, C > 0 ? C : 0, (H > 0 && W > 0) ? H * W : 0);
}

extern "C" int l2s_dynfilter_fwd(const float* X, const float* filt, const float* fuse, const int32_t* expr2img,
                                 float* response, float* rk_saved, float* Y, const float* resp_target,
                                 float* resp_loss, int I, int E, int C, int H, int W, int flags, void* workspace,
                                 size_t workspace_bytes, l2s_stream_t stream) {
  int rc = check(X, filt, fuse, expr2img, I, E, C, H, W, flags);
  if (rc) return rc;
  L2S_REQUIRE(response && Y, L2S_ERR_ARG, "dynfilter_fwd: null output");
  if (E == 0) return L2S_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const DfGeom g = make_geom(I, E, C, H, W, flags);
  // tensor-core kernel (dynfilter_tc.cu) where the shape allows it: 0 = ran, 1 = not applicable, < 0 = error
  rc = launch_dynfilter_tc_fwd(X, filt, fuse, expr2img, response, rk_saved, Y, resp_target, resp_loss, I, E, C, H, W, flags,
                               workspace, workspace_bytes, st);
  if (rc <= 0) return rc;
  if (resp_loss) L2S_CUDA_OK(cudaMemsetAsync(resp_loss, 0, sizeof(float) * E, st));   // the FFMA kernel accumulates into it
  const bool vec = (g.HW % 2 == 0) && aligned16(X) && aligned16(Y);      // rows 8-byte aligned at least (ld_quad / st_quad)
  const size_t cap = (size_t)max_smem_optin();
  if (fwd_smem(C, 16) <= cap && vec)
    return launch_fwd<16, true>(X, filt, fuse, expr2img, response, rk_saved, Y, resp_target, resp_loss, g, st);
  if (fwd_smem(C, 16) <= cap)
    return launch_fwd<16, false>(X, filt, fuse, expr2img, response, rk_saved, Y, resp_target, resp_loss, g, st);
  if (fwd_smem(C, 4) <= cap)
    return launch_fwd<4, false>(X, filt, fuse, expr2img, response, rk_saved, Y, resp_target, resp_loss, g, st);
  return fail(L2S_ERR_SHAPE, "dynfilter_fwd: C=%d too large for the shared-memory tile", C);
}

extern "C" size_t l2s_dynfilter_bwd_workspace_bytes(int I, int E, int C, int H, int W) {
  (void)C;
  // dr + recomputed r_k + the per-image expression ranges of the TMA-streamed kernel
  return (size_t)E * H * W * sizeof(float) * (1 + NF) + 256 + dynfilter_bwd_tma_workspace_bytes(I > 0 ? I : 0);
}

extern "C" int l2s_dynfilter_bwd(const float* X, const float* filt, const float* fuse, const int32_t* expr2img,
                                 const float* response, const float* rk_saved, const float* dY,
                                 const float* dresponse, const float* resp_target, const float* resp_gscale,
                                 float* dX, float* dfilt,
                                 float* dfuse, int I, int E, int C, int H, int W, int flags, void* workspace,
                                 size_t workspace_bytes, l2s_stream_t stream) {
  int rc = check(X, filt, fuse, expr2img, I, E, C, H, W, flags);
  if (rc) return rc;
  L2S_REQUIRE(response && dY && dX && dfilt && dfuse, L2S_ERR_ARG, "dynfilter_bwd: null pointer");
  L2S_REQUIRE(workspace && workspace_bytes >= l2s_dynfilter_bwd_workspace_bytes(I, E, C, H, W), L2S_ERR_WORKSPACE,
              "dynfilter_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const DfGeom g = make_geom(I, E, C, H, W, flags);
  if (E == 0) {
    L2S_CUDA_OK(cudaMemsetAsync(dX, 0, sizeof(float) * (size_t)I * C * g.HW, st));
    return L2S_OK;
  }
  float* drbuf = reinterpret_cast<float*>(workspace);
  float* rk_ws = drbuf + (size_t)E * g.HW;
  const bool vec = (g.HW % 2 == 0) && aligned16(X) && aligned16(dY) && aligned16(dX);
  const size_t cap = (size_t)max_smem_optin();
  // TMA-streamed kernel (dynfilter_bwd_tma.cu) where the shape allows it: 0 = ran, 1 = not applicable, < 0 = error
  void* seg_ws = reinterpret_cast<char*>(workspace) + (((size_t)E * g.HW * sizeof(float) * (1 + NF) + 255) & ~(size_t)255);
  rc = launch_dynfilter_bwd_tma(X, filt, fuse, expr2img, response, dY, dresponse, resp_target, resp_gscale, dX, drbuf, I, E, C,
                                H, W, flags, seg_ws, dynfilter_bwd_tma_workspace_bytes(I), st);
  if (rc < 0) return rc;
  if (rc == 0) {
    // ran
  } else if (vec && C <= kRegCh * (kThreads / 4) && bwd_reg_smem(C, 16) <= cap) {
    rc = 0;
    auto kern = dynfilter_bwd_reg_kernel<16>;
    const size_t smem = bwd_reg_smem(C, 16);
    L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.HW + 15) / 16, I);
    kern<<<grid, kThreads, smem, st>>>(X, filt, fuse, expr2img, response, dY, dresponse, resp_target, resp_gscale, dX,
                                       drbuf, g);
    L2S_LAUNCH_OK("dynfilter_bwd_reg_kernel");
    count_launch();
  } else if (bwd_smem(C, 16) <= cap && vec)
    rc = launch_bwd<16, true>(X, filt, fuse, expr2img, response, dY, dresponse, resp_target, resp_gscale, dX, drbuf, g, st);
  else if (bwd_smem(C, 16) <= cap)
    rc = launch_bwd<16, false>(X, filt, fuse, expr2img, response, dY, dresponse, resp_target, resp_gscale, dX, drbuf, g, st);
  else if (bwd_smem(C, 4) <= cap)
    rc = launch_bwd<4, false>(X, filt, fuse, expr2img, response, dY, dresponse, resp_target, resp_gscale, dX, drbuf, g, st);
  else
    return fail(L2S_ERR_SHAPE, "dynfilter_bwd: C=%d too large for the shared-memory tile", C);
  if (rc) return rc;
  {
    const size_t smem_v = ((size_t)EBV * g.HW * sizeof(float) + g.HW + 15) & ~(size_t)15;
    dim3 grid((C + 7) / 8, I);
    static const bool scalar_dfilt = env_flag("L2S_DFILT_SCALAR");
    if (vec && aligned16(drbuf) && smem_v <= cap && !scalar_dfilt) {
      auto kern = (g.HW % 4 == 0) ? dynfilter_dfilt_vec_kernel<true> : dynfilter_dfilt_vec_kernel<false>;
      L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
      kern<<<grid, 256, smem_v, st>>>(X, fuse, expr2img, drbuf, dfilt, g);
      L2S_LAUNCH_OK("dynfilter_dfilt_vec_kernel");
    } else {
      const size_t smem = (size_t)EB * g.HW * sizeof(float);
      L2S_REQUIRE(smem <= cap, L2S_ERR_SHAPE, "dynfilter_bwd: H*W=%d too large", g.HW);
      L2S_CUDA_OK(cudaFuncSetAttribute(dynfilter_dfilt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dynfilter_dfilt_kernel<<<grid, 256, smem, st>>>(X, fuse, expr2img, drbuf, dfilt, g);
      L2S_LAUNCH_OK("dynfilter_dfilt_kernel");
    }
  }
  {
    const float* rk = rk_saved;
    if (!rk) {   // the forward did not keep r_k: recompute it (slow path)
      dim3 grid((g.HW + 255) / 256, E);
      dynfilter_rk_kernel<<<grid, 256, 0, st>>>(X, filt, expr2img, rk_ws, g);
      L2S_LAUNCH_OK("dynfilter_rk_kernel");
      count_launch();
      rk = rk_ws;
    }
    dynfilter_dfuse_kernel<<<E, 256, 0, st>>>(drbuf, rk, dfuse, g.HW);
    L2S_LAUNCH_OK("dynfilter_dfuse_kernel");
  }
  count_launch(2);
  return L2S_OK;
}
