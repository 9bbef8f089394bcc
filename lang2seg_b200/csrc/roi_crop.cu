// Crop-and-resize ROI pooling, forward and backward, for sm_100a.
//
// Semantics: Network._crop_pool_layer / _crop_pool_layer_align of the lang2seg reference
// (pyutils/mask-faster-rcnn/lib/nets/network_cycle_response.py:107-182): S x S bilinear samples
// evenly spaced from box start to box end inclusive (align_corners=True), zero padding outside
// the map, optional 2x2 max over a 2S x 2S sample grid.  Closed form: SURVEY.md appendix A.2.
//
// Design (HBM-bound; the output / upstream gradient is ~12x the size of the map):
//   * ROIs are binned by batch index (roi_count/order_kernel) and roi_geom_kernel writes ONE geometry
//     record per ROI: for every sample the four shared-memory slots of its corners (zero slot when the
//     corner is outside the map) and the two bilinear fractions; for the 7x7 variant also a collision
//     rank per sample (samples of equal rank never share a corner cell).  Sample geometry does not depend
//     on the channel, so the 32 channel-chunk CTAs of an expression read it instead of recomputing it.
//   * one CTA owns (batch index b, chunk of CC channels) and stages that slice of the map ONCE in shared
//     memory, transposed to [pixel][channel] with a 16-byte XOR swizzle, then loops over all ROIs of b.
//   * forward: thread = (4-channel group, sample); geometry records arrive by 1-D TMA bulk loads one
//     iteration ahead; four LDS.128 gathers per output float4; results staged in [channel][49] tiles
//     that leave by one TMA bulk store per ROI, double buffered.
//   * backward: upstream-gradient tiles and geometry records arrive by TMA bulk loads through an
//     mbarrier ring fed by a producer warp.  Consumer warp cg exclusively OWNS channel quad cg of the
//     shared-memory accumulator map, so plain vectorised read-modify-write is race free across warps;
//     lanes are samples, the four corners are four phases, and equal-cell samples are serialised by their
//     precomputed rank (fixed order => bit-reproducible sums).  No floating-point atomics anywhere
//     (shared fp32 atomicAdd costs 2 cycles per lane on sm_100).
#include <cstdlib>
#include <initializer_list>

#include "common.cuh"

namespace l2s {
namespace {

constexpr int kPP = 49;             // outputs per (ROI, channel): cfg.POOLING_SIZE^2
constexpr int kFwdItems = 32 * kPP; // (ROI, sample, channel quad) items per forward iteration pass: RPI * CGN == 32
constexpr int kFwdThreads = 800;    // 25 warps: items t and t + 784 ... of an iteration (784 = kFwdItems / 2)
constexpr int kBwdStages = 6;
constexpr int kRecTail = 128;       // bytes after the entries of a geometry record (see roi_geom_kernel)

__device__ __forceinline__ int swz_key(int px) { return (px ^ (px >> 3) ^ (px >> 6)) & 7; }

// ------------------------------------------------------------------ ROI binning (by batch index)
__global__ void roi_count_kernel(const float* __restrict__ rois, int N, int* __restrict__ counts) {
  const int b = blockIdx.x;
  int cnt = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) cnt += ((int)__ldg(rois + 5 * (size_t)n) == b);
  cnt = (int)warp_sum((float)cnt);   // counts < 2^24: exact in fp32
  __shared__ int s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    counts[b] = t;
  }
}

// seg[b]..seg[b+1] indexes `order`, which lists the ROIs of batch b in ascending ROI index.
__global__ void roi_order_kernel(const float* __restrict__ rois, int N, const int* __restrict__ counts,
                                 int* __restrict__ seg, int* __restrict__ order) {
  const int b = blockIdx.x, B = gridDim.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int nw = blockDim.x >> 5;
  __shared__ int s_w[32];
  __shared__ int s_base;
  int part = 0;
  for (int i = t; i < b; i += blockDim.x) part += counts[i];
  part = (int)warp_sum((float)part);
  if (lane == 0) s_w[wid] = part;
  __syncthreads();
  if (t == 0) {
    int base = 0;
    for (int i = 0; i < nw; ++i) base += s_w[i];
    s_base = base;
    seg[b] = base;
    if (b == B - 1) seg[B] = base + counts[b];
  }
  __syncthreads();
  int running = s_base;
  for (int n0 = 0; n0 < N; n0 += blockDim.x) {
    const int n = n0 + t;
    const bool m = n < N && (int)__ldg(rois + 5 * (size_t)n) == b;
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    __syncthreads();
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int i = 0; i < nw; ++i) {
      const int c = s_w[i];
      if (i < wid) off += c;
      tot += c;
    }
    if (m) order[running + off + __popc(bal & ((1u << lane) - 1u))] = n;
    running += tot;
  }
}


struct CropGeom {
  int C, H, W, N, B;
  int S;        // samples per side (pool or 2*pool)
  int maxpool;  // 0/1
  float sx, sy; // image -> feature scale
  int rec;      // bytes per geometry record = S*S*16 + kRecTail
};

// sample coordinate -> corner index / weights.  px = x1 + (x2-x1) * j/(S-1) in feature pixels.
struct Corner {
  int x0, y0;
  float lx, ly;
};
__device__ __forceinline__ Corner sample_at(const float4& box, int i, int j, float inv) {
  const float tx = (float)j * inv, ty = (float)i * inv;
  const float px = fmaf(box.z - box.x, tx, box.x);
  const float py = fmaf(box.w - box.y, ty, box.y);
  const float fx = floorf(px), fy = floorf(py);
  Corner c;
  c.x0 = (int)fx;
  c.y0 = (int)fy;
  c.lx = px - fx;
  c.ly = py - fy;
  return c;
}

// ------------------------------------------------------------------ geometry records
// record r (position in `order`):  entries[S*S] | ranks[49] u8 | ... | maxrank u8 @ +60 | ROI id i32 @ +56
// entry: 4 x u16 float4-slot of the corners (00, 01, 10, 11) for channel quad 0 -- the consumer XORs its
// quad index in -- or the zero slot HW*CGN when the corner lies outside the map; then ly, lx.
struct __align__(16) GeomEntry {
  uint16_t s00, s01, s10, s11;
  float ly, lx;
};

// Separable form of the same geometry for the row-owner backward: sample (i,j) sits at (py_i, px_j).
constexpr int kRowWarps = 31;       // consumer warps of the row-owner backward: warp w owns map rows y = w (mod 31)
constexpr int kRowLd = 33;          // channel stride of its accumulator map (odd: conflict free both ways)
struct __align__(16) SepRec {
  int xoff[8];          // (clamped x0_j) * kRowLd : float offset of the left corner inside a padded accumulator row
  float lx[8];
  float ly[8];
  int16_t y0[8];        // clamped to [-2, 30000]
  uint16_t wmask[32];   // per consumer warp: bit 2i+d set <=> map row y0_i + d is inside the map and owned by the warp
  int n;                // ROI index
  int mode;             // column pattern: 2 = all 14 corner columns distinct, 1 = x0 strictly increasing, 0 = general
  int pad[2];
};
static_assert(sizeof(SepRec) == 192, "SepRec layout");

__global__ void __launch_bounds__(256)
roi_geom_kernel(const float* __restrict__ rois, const int* __restrict__ order, const int* __restrict__ seg,
                unsigned char* __restrict__ table, SepRec* __restrict__ sep, CropGeom g, int CGN) {
  const int r = blockIdx.x, p = threadIdx.x;
  if (r >= seg[g.B]) return;     // ROIs with a batch index outside [0,B) are not listed in `order` (block uniform)
  const int n = __ldg(order + r);
  const int SS = g.S * g.S;
  __shared__ int s_key[64];
  unsigned char* rec = table + (size_t)r * g.rec;
  const float* rp = rois + 5 * (size_t)n;
  float4 box;
  box.x = __ldg(rp + 1) * g.sx;
  box.y = __ldg(rp + 2) * g.sy;
  box.z = __ldg(rp + 3) * g.sx;
  box.w = __ldg(rp + 4) * g.sy;
  const float inv = 1.0f / (float)(g.S - 1);
  int key = 0;
  if (p < SS) {
    const Corner c = sample_at(box, p / g.S, p % g.S, inv);
    const int W = g.W, H = g.H, zs = H * W * CGN;
    const bool vx0 = (unsigned)c.x0 < (unsigned)W, vx1 = (unsigned)(c.x0 + 1) < (unsigned)W;
    const bool vy0 = (unsigned)c.y0 < (unsigned)H, vy1 = (unsigned)(c.y0 + 1) < (unsigned)H;
    auto slot = [&](bool ok, int yy, int xx) -> uint16_t {
      if (!ok) return (uint16_t)zs;
      const int px = yy * W + xx;
      return (uint16_t)(px * CGN + (swz_key(px) & (CGN - 1)));
    };
    GeomEntry e;
    e.s00 = slot(vy0 && vx0, c.y0, c.x0);
    e.s01 = slot(vy0 && vx1, c.y0, c.x0 + 1);
    e.s10 = slot(vy1 && vx0, c.y0 + 1, c.x0);
    e.s11 = slot(vy1 && vx1, c.y0 + 1, c.x0 + 1);
    e.ly = c.ly;
    e.lx = c.lx;
    reinterpret_cast<GeomEntry*>(rec)[p] = e;
    const int y0 = min(max(c.y0, -2), 30000), x0 = min(max(c.x0, -2), 30000);
    key = (y0 << 16) | (x0 & 0xffff);
  }
  if (g.S == 7) {   // launched with 64 threads
    // The ranked backward splits the corners of a sample between two warps by the PARITY h of the map row
    // (row y0 or y0+1).  Inside warp h two samples meet in a pixel through the same (left/right) corner iff they
    // agree in (row_h, x0): rank_h = number of earlier samples with the same key.  tail layout:
    //   [0..48] rank_0 | bit 7: parity of y0      [64..112] rank_1      [56..59] ROI id   [60],[61] max rank_0/1
    const int yq = key >> 16, xq = (int)(short)(key & 0xffff);
    unsigned char* tail = rec + (size_t)SS * 16;
    for (int h = 0; h < 2; ++h) {
      const int row = yq + (((yq & 1) ^ h) & 1);
      const int kh = (p < SS) ? ((row << 16) | (xq & 0xffff)) : (int)(0x80000000u | (unsigned)p);
      __syncthreads();
      s_key[p] = kh;
      __syncthreads();
      int rank = 0;
      if (p < SS)
        for (int q = 0; q < p; ++q) rank += (s_key[q] == kh);
      __syncthreads();
      const int mr = __reduce_max_sync(0xffffffffu, p < SS ? rank : 0);
      if ((p & 31) == 0) s_key[p >> 5] = mr;
      __syncthreads();
      if (p < SS) tail[h * 64 + p] = (unsigned char)(rank | (h == 0 ? ((yq & 1) << 7) : 0));
      if (p == 0) tail[60 + h] = (unsigned char)max(s_key[0], s_key[1]);
    }
  }
  if (p == 0) *reinterpret_cast<int*>(rec + (size_t)SS * 16 + 56) = n;
  if (sep != nullptr && g.S == 7 && p < 32) {
    SepRec* sr = sep + r;
    int y0c[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) y0c[i] = min(max(sample_at(box, i, i, inv).y0, -2), 30000);
    // consumer warp p: which of the 14 (sample row, upper/lower) pairs land in one of its rows
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int row = y0c[i] + d;
        if (row >= 0 && row < g.H && (row % kRowWarps) == p) m |= 1u << (2 * i + d);
      }
    if (p >= kRowWarps) m = 0;
    sr->wmask[p] = (uint16_t)m;
    if (p < 8) {
      const Corner c = sample_at(box, min(p, 6), min(p, 6), inv);
      sr->xoff[p] = min(max(c.x0, -2), g.W) * kRowLd;
      sr->lx[p] = c.lx;
      sr->ly[p] = c.ly;
      sr->y0[p] = (int16_t)min(max(c.y0, -2), 30000);
    }
    if (p == 0) {
      // column pattern on the CLAMPED x0 (several samples clamped to the same guard column fall back to mode 0)
      int mode = 2, prev = 0;
      for (int j = 0; j < 7; ++j) {
        const int xj = min(max(sample_at(box, j, j, inv).x0, -2), g.W);
        if (j > 0) {
          if (xj <= prev) mode = 0;
          else if (xj == prev + 1) mode = min(mode, 1);
        }
        prev = xj;
      }
      sr->n = n;
      sr->mode = mode;
      sr->pad[0] = sr->pad[1] = 0;
    }
  }
}

// ------------------------------------------------------------------ staging the map slice
// smem map: float4 slot (px, cgrp) at px*CGN + ((cgrp ^ key(px)) & (CGN-1)); CGN zero slots follow at px = HW.
template <int CC>
__device__ __forceinline__ int map_off(int px, int cgrp) {
  constexpr int CGN = CC / 4;
  return px * CC + (((cgrp ^ swz_key(px)) & (CGN - 1)) << 2);
}

template <int CC>
__device__ void stage_map(float* __restrict__ map, const float* __restrict__ src /* (C,HW) of batch b */,
                          int c0, int C, int HW) {
  constexpr int CGN = CC / 4;
  constexpr int U = 8;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int pl = lane & 7, cl = lane >> 3;
  const int nblk = (HW + 7) >> 3;
  const int total = nblk * CGN;
  for (int it0 = wid; it0 < total; it0 += nw * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {       // U independent loads in flight per lane
      const int it = it0 + u * nw;
      const int gq = it % CGN, blk = it / CGN;
      const int px = blk * 8 + pl, c = c0 + gq * 4 + cl;
      v[u] = (it < total && px < HW && c < C) ? __ldg(src + (size_t)c * HW + px) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = it0 + u * nw;
      const int gq = it % CGN, blk = it / CGN;
      const int px = blk * 8 + pl;
      if (it < total && px < HW) map[map_off<CC>(px, gq) + cl] = v[u];
    }
  }
  for (int i = threadIdx.x; i < CC; i += blockDim.x) map[(size_t)HW * CC + i] = 0.f;   // the zero slots
}

template <int CC>
__device__ void unstage_map(const float* __restrict__ map, float* __restrict__ dst, int c0, int C, int HW) {
  constexpr int CGN = CC / 4;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int pl = lane & 7, cl = lane >> 3;
  const int nblk = (HW + 7) >> 3;
  for (int it = wid; it < nblk * CGN; it += nw) {
    const int gq = it % CGN, blk = it / CGN;
    const int px = blk * 8 + pl, c = c0 + gq * 4 + cl;
    if (px < HW && c < C) dst[(size_t)c * HW + px] = map[map_off<CC>(px, gq) + cl];
  }
}

__device__ __forceinline__ float4 bilerp4(const float4* __restrict__ map4, const uint4& e, int cg) {
  const float ly = __uint_as_float(e.z), lx = __uint_as_float(e.w);
  const float4 a = map4[(e.x & 0xffffu) ^ cg];
  const float4 b = map4[(e.x >> 16) ^ cg];
  const float4 c = map4[(e.y & 0xffffu) ^ cg];
  const float4 d = map4[(e.y >> 16) ^ cg];
  const float wy0 = 1.f - ly, wx0 = 1.f - lx;
  const float w00 = wy0 * wx0, w01 = wy0 * lx, w10 = ly * wx0, w11 = ly * lx;
  float4 o;
  o.x = fmaf(w11, d.x, fmaf(w10, c.x, fmaf(w01, b.x, w00 * a.x)));
  o.y = fmaf(w11, d.y, fmaf(w10, c.y, fmaf(w01, b.y, w00 * a.y)));
  o.z = fmaf(w11, d.z, fmaf(w10, c.z, fmaf(w01, b.z, w00 * a.z)));
  o.w = fmaf(w11, d.w, fmaf(w10, c.w, fmaf(w01, b.w, w00 * a.w)));
  return o;
}

// ------------------------------------------------------------------ forward
// Iteration = kRPI ROIs: kRPI * 49 * CGN (ROI, sample, channel quad) items, up to two per thread.
// Geometry records arrive through a ring of NT buffers, loaded NT-1 iterations ahead (a bulk load from L2/HBM
// takes longer than one iteration).
constexpr int kRPI = 4;
template <int CC, bool MAXPOOL>
__global__ void __launch_bounds__(kFwdThreads, 1)
roi_crop_fwd_kernel(const float* __restrict__ bottom, const int* __restrict__ seg,
                    const unsigned char* __restrict__ table, float* __restrict__ out,
                    uint8_t* __restrict__ argmax, CropGeom g) {
  constexpr int CGN = CC / 4;
  constexpr int RPI = kRPI;
  constexpr int NT = MAXPOOL ? 2 : 4;      // table ring depth
  constexpr int ITEMS = RPI * kPP * CGN;
  constexpr int HALF = kFwdItems / 2;      // 784
  constexpr int TILE = CC * kPP;           // floats per ROI tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W;
  float* map = reinterpret_cast<float*>(smem_raw);
  float* tiles = map + (size_t)(HW + 1) * CC;                               // [2][RPI][TILE]
  unsigned char* tabs = reinterpret_cast<unsigned char*>(tiles + 2 * RPI * TILE);   // [NT][RPI * rec]
  uint8_t* atile = tabs + (size_t)NT * RPI * g.rec;                         // [2][RPI][TILE] (max-pool only)
  __shared__ uint64_t tab_bar[NT];
  __shared__ int s_n[2][RPI];

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x;
  const int cvalid = min(CC, g.C - c0);
  const int beg = seg[b], end = seg[b + 1];
  const int nit = (end - beg + RPI - 1) / RPI;

  if (t == 0) {
    for (int i = 0; i < NT; ++i) mbar_init(&tab_bar[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto load_tab = [&](int it) {
    const int gi = beg + it * RPI;
    const uint32_t bytes = (uint32_t)(min(RPI, end - gi) * g.rec);
    mbar_arrive_expect_tx(&tab_bar[it % NT], bytes);
    bulk_g2s(tabs + (size_t)(it % NT) * RPI * g.rec, table + (size_t)gi * g.rec, bytes, &tab_bar[it % NT]);
  };
  if (t == 0)
    for (int i = 0; i < NT - 1 && i < nit; ++i) load_tab(i);
  stage_map<CC>(map, bottom + (size_t)b * g.C * HW, c0, g.C, HW);
  __syncthreads();

  const float4* map4 = reinterpret_cast<const float4*>(map);
  // this thread's (up to) two items of every iteration
  int i_slot[2], i_p[2], i_cg[2];
  bool i_on[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int item = t + k * HALF;
    i_on[k] = t < HALF && item < ITEMS;
    i_cg[k] = item % CGN;
    const int sp = item / CGN;
    i_slot[k] = sp / kPP;
    i_p[k] = sp - i_slot[k] * kPP;
  }

  for (int it = 0; it < nit; ++it) {
    const int buf = it & 1;
    const int gi = beg + it * RPI;
    const int cnt = min(RPI, end - gi);
    // ring slot (it-1) % NT was last read in iteration it-1, i.e. before the previous barrier
    if (t == 0 && it + NT - 1 < nit) load_tab(it + NT - 1);
    mbar_wait(&tab_bar[it % NT], (it / NT) & 1);
    const unsigned char* tab = tabs + (size_t)(it % NT) * RPI * g.rec;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int slot = i_slot[k], p = i_p[k], cg = i_cg[k];
      if (i_on[k] && slot < cnt) {
        const unsigned char* rec = tab + (size_t)slot * g.rec;
        const uint4* ent = reinterpret_cast<const uint4*>(rec);
        float* tile = tiles + (size_t)(buf * RPI + slot) * TILE;
        float4 o;
        if (!MAXPOOL) {
          o = bilerp4(map4, ent[p], cg);
        } else {
          // 2x2 block of the 2S x 2S sample grid; first strict maximum in row-major order wins,
          // as in max_pool2d's backward (network_cycle_response.py:144)
          const int pi = p / 7, pj = p - pi * 7;
          const int e0 = (2 * pi) * 14 + 2 * pj;
          o = bilerp4(map4, ent[e0], cg);
          uint32_t am = 0;   // 4 x 8-bit winners
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            const float4 v = bilerp4(map4, ent[e0 + (q >> 1) * 14 + (q & 1)], cg);
            if (v.x > o.x) { o.x = v.x; am = (am & ~0xffu) | (uint32_t)q; }
            if (v.y > o.y) { o.y = v.y; am = (am & ~0xff00u) | ((uint32_t)q << 8); }
            if (v.z > o.z) { o.z = v.z; am = (am & ~0xff0000u) | ((uint32_t)q << 16); }
            if (v.w > o.w) { o.w = v.w; am = (am & ~0xff000000u) | ((uint32_t)q << 24); }
          }
          uint8_t* at = atile + (size_t)(buf * RPI + slot) * TILE;
          const int cb = cg * 4;
          at[(cb + 0) * kPP + p] = (uint8_t)(am & 0xff);
          at[(cb + 1) * kPP + p] = (uint8_t)((am >> 8) & 0xff);
          at[(cb + 2) * kPP + p] = (uint8_t)((am >> 16) & 0xff);
          at[(cb + 3) * kPP + p] = (uint8_t)(am >> 24);
        }
        const int cb = cg * 4;
        tile[(cb + 0) * kPP + p] = o.x;
        tile[(cb + 1) * kPP + p] = o.y;
        tile[(cb + 2) * kPP + p] = o.z;
        tile[(cb + 3) * kPP + p] = o.w;
        if (p == 0 && cg == 0) s_n[buf][slot] = *reinterpret_cast<const int*>(rec + (size_t)g.S * g.S * 16 + 56);
      }
    }
    fence_proxy_async_smem();
    if (t == 0) bulk_wait_read<0>();      // the previous iteration's stores have left the other buffer
    __syncthreads();
    if (t == 0) {
      for (int s = 0; s < cnt; ++s) {
        const int n = s_n[buf][s];
        bulk_s2g(out + ((size_t)n * g.C + c0) * kPP, tiles + (size_t)(buf * RPI + s) * TILE,
                 (uint32_t)(cvalid * kPP * sizeof(float)));
      }
      bulk_commit();
    }
    if (MAXPOOL && argmax != nullptr) {
      // winners: plain 32-bit copies (tile byte count and global offset are multiples of 4)
      const int words = cvalid * kPP / 4;
      for (int w = t; w < cnt * words; w += blockDim.x) {
        const int s = w / words, k = w % words;
        const int n = s_n[buf][s];
        reinterpret_cast<uint32_t*>(argmax + ((size_t)n * g.C + c0) * kPP)[k] =
            reinterpret_cast<const uint32_t*>(atile + (size_t)(buf * RPI + s) * TILE)[k];
      }
    }
  }
  if (t == 0) bulk_wait<0>();
}

// ------------------------------------------------------------------ forward, warp-per-ROI variant (CC = 32, 7x7)
// The block-synchronous kernel above spends 20 % of its issue slots waiting at the per-iteration barrier and keeps the
// shared-memory pipe only ~50 % busy (ncu).  Here nothing is block wide after the map has been staged: every warp owns
// whole ROIs (ROI r = beg + wid, beg + wid + NW, ...), reads the 49 geometry entries of its ROI straight from L2 into
// registers one ROI ahead (13 independent 16-byte loads per lane), gathers into its OWN [32 ch][49] tile and sends the
// tile off with its own TMA bulk store; the only wait is the warp's own previous store having been read out of the
// tile.  Up to kFwdWarps stores are in flight per SM, and warps drift apart so that gathers, tile writes and TMA
// reads of different ROIs overlap in the shared-memory pipe.
constexpr int kFwdWarps = 12;      // 14 fit in shared memory but cap the registers at 128 per thread: measured slower

__global__ void __launch_bounds__(kFwdWarps * 32, 1)
roi_crop_fwd_warp_kernel(const float* __restrict__ bottom, const int* __restrict__ seg,
                         const unsigned char* __restrict__ table, float* __restrict__ out, CropGeom g) {
  constexpr int CC = 32;
  constexpr int CGN = CC / 4;
  constexpr int TILE = CC * kPP;           // floats per ROI tile
  constexpr int NPASS = (kPP * CGN + 31) / 32;      // 13 passes of (4 samples x 8 channel quads)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W;
  float* map = reinterpret_cast<float*>(smem_raw);
  float* tiles = map + (size_t)(HW + 1) * CC;        // [kFwdWarps][TILE]

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cvalid = min(CC, g.C - c0);
  const int beg = seg[b], end = seg[b + 1];

  stage_map<CC>(map, bottom + (size_t)b * g.C * HW, c0, g.C, HW);
  __syncthreads();

  const float4* map4 = reinterpret_cast<const float4*>(map);
  float* tile = tiles + (size_t)wid * TILE;
  const int cg = lane & (CGN - 1), ps = lane >> 3;   // channel quad, sample slot of a pass
  const int cb = cg * 4;
  const uint32_t tile_bytes = (uint32_t)(cvalid * kPP * sizeof(float));

  auto load_entries = [&](int r, uint4 (&e)[NPASS], int* n) {
    const unsigned char* rec = table + (size_t)r * g.rec;
    const uint4* ent = reinterpret_cast<const uint4*>(rec);
#pragma unroll
    for (int k = 0; k < NPASS; ++k) {
      const int p = 4 * k + ps;
      e[k] = (p < kPP) ? __ldg(ent + p) : make_uint4(0, 0, 0, 0);
    }
    *n = __ldg(reinterpret_cast<const int*>(rec + (size_t)kPP * 16 + 56));
  };

  uint4 cur[NPASS];
  int ncur = 0;
  int r = beg + wid;
  if (r < end) load_entries(r, cur, &ncur);
  for (; r < end; r += kFwdWarps) {
    uint4 nxt[NPASS];
    int nnxt = 0;
    const bool more = r + kFwdWarps < end;
    if (more) load_entries(r + kFwdWarps, nxt, &nnxt);       // in flight during this ROI's gathers
    if (lane == 0) bulk_wait_read<0>();                      // my previous store has left the tile
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NPASS; ++k) {
      const int p = 4 * k + ps;
      if (p < kPP) {
        const float4 o = bilerp4(map4, cur[k], cg);
        tile[(cb + 0) * kPP + p] = o.x;
        tile[(cb + 1) * kPP + p] = o.y;
        tile[(cb + 2) * kPP + p] = o.z;
        tile[(cb + 3) * kPP + p] = o.w;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(out + ((size_t)ncur * g.C + c0) * kPP, tile, tile_bytes);
      bulk_commit();
    }
    if (more) {
#pragma unroll
      for (int k = 0; k < NPASS; ++k) cur[k] = nxt[k];
      ncur = nnxt;
    }
  }
  if (lane == 0) bulk_wait<0>();
}

// ------------------------------------------------------------------ backward
// warps: CGN consumers + 1 producer.  Consumer warp cg owns channel quad cg of the accumulator map.
// Stage = gradient tile [CC][49] (+ winners [CC][49] u8 with max-pool) + the ROI's geometry record.
template <int CC, bool MAXPOOL>
__global__ void __launch_bounds__(((MAXPOOL ? 1 : 2) * CC / 4 + 1) * 32, 1)
roi_crop_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ seg,
                    const unsigned char* __restrict__ table, const uint8_t* __restrict__ argmax,
                    float* __restrict__ dbottom, CropGeom g) {
  constexpr int CGN = CC / 4;
  constexpr int NH = MAXPOOL ? 1 : 2;        // consumer warps per channel quad (row-parity split)
  constexpr int NCONS = NH * CGN;
  constexpr int TILE = CC * kPP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W;
  float* map = reinterpret_cast<float*>(smem_raw);
  float* tiles = map + (size_t)(HW + 1) * CC;                                 // [kBwdStages][TILE]
  unsigned char* recs = reinterpret_cast<unsigned char*>(tiles + kBwdStages * TILE);   // [kBwdStages][rec]
  uint8_t* atiles = recs + (size_t)kBwdStages * g.rec;                        // [kBwdStages][TILE] (max-pool)
  __shared__ uint64_t full_bar[kBwdStages], empty_bar[kBwdStages];

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cvalid = min(CC, g.C - c0);

  for (int i = t; i < (HW + 1) * CC; i += blockDim.x) map[i] = 0.f;
  if (t == 0) {
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NCONS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int beg = seg[b], end = seg[b + 1];
  const uint32_t tile_bytes = (uint32_t)(cvalid * kPP * sizeof(float));
  const uint32_t arg_bytes = (uint32_t)(cvalid * kPP);
  // winners travel by TMA only when their rows are 16-byte sized / aligned (C % 16 == 0); otherwise read in place
  const bool arg_smem = MAXPOOL && (g.C % 16 == 0) && (cvalid % 16 == 0);

  if (wid == NCONS) {
    // ---------------- producer warp: streams tiles + records (ROI ids fetched 32 at a time) ----------------
    for (int r0 = beg, k = 0; r0 < end; r0 += 32) {
      const int mine = r0 + lane;
      int nl = 0;
      if (mine < end) nl = __ldg(reinterpret_cast<const int*>(table + (size_t)mine * g.rec + (size_t)g.S * g.S * 16 + 56));
      {   // L2 prefetch of the group after this one (the first group is prefetched before the loop)
        const int ahead = mine + (r0 == beg ? 0 : 32);
        for (int a = ahead; a < end && a <= mine + 32; a += 32) {
          const int na = __ldg(reinterpret_cast<const int*>(table + (size_t)a * g.rec + (size_t)g.S * g.S * 16 + 56));
          bulk_prefetch_l2(dout + ((size_t)na * g.C + c0) * kPP, tile_bytes);
          bulk_prefetch_l2(table + (size_t)a * g.rec, (uint32_t)g.rec);
        }
      }
      const int cnt = min(32, end - r0);
      for (int j = 0; j < cnt; ++j, ++k) {
        const int n = __shfl_sync(0xffffffffu, nl, j);
        if (lane == 0) {
          const int s = k % kBwdStages;
          if (k >= kBwdStages) mbar_wait(&empty_bar[s], ((k / kBwdStages) - 1) & 1);
          const unsigned char* rec = table + (size_t)(r0 + j) * g.rec;
          mbar_arrive_expect_tx(&full_bar[s], tile_bytes + (uint32_t)g.rec + (arg_smem ? arg_bytes : 0u));
          bulk_g2s(recs + (size_t)s * g.rec, rec, (uint32_t)g.rec, &full_bar[s]);
          bulk_g2s(tiles + (size_t)s * TILE, dout + ((size_t)n * g.C + c0) * kPP, tile_bytes, &full_bar[s]);
          if (arg_smem)
            bulk_g2s(atiles + (size_t)s * TILE, argmax + ((size_t)n * g.C + c0) * kPP, arg_bytes, &full_bar[s]);
        }
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int cg = wid % CGN, hpar = wid / CGN;
    const bool ch_ok = cg * 4 < cvalid;
    float4* map4 = reinterpret_cast<float4*>(map);
    const unsigned lt = (1u << lane) - 1u;
    const int zs = HW * CGN;
    const int p0 = lane, p1 = 32 + lane;
    const bool on1 = p1 < kPP;
    for (int ri = beg, k = 0; ri < end; ++ri, ++k) {
      const int s = k % kBwdStages;
      mbar_wait(&full_bar[s], (k / kBwdStages) & 1);
      const float* tile = tiles + (size_t)s * TILE;
      const unsigned char* rec = recs + (size_t)s * g.rec;
      const uint4* ent = reinterpret_cast<const uint4*>(rec);
      const int cb = cg * 4;
      if (!MAXPOOL) {
        // warp (cg, hpar): of every sample the two corners that lie in the map row of parity hpar
        const unsigned char* tail = rec + kPP * 16;
        const int mr = tail[60 + hpar];
        const uint4 e0 = ent[p0];
        const uint4 e1 = on1 ? ent[p1] : make_uint4(0, 0, 0, 0);
        const int top0 = ((tail[p0] >> 7) == hpar), top1 = on1 ? ((tail[p1] >> 7) == hpar) : 0;
        const int r0 = ch_ok ? (int)(tail[hpar * 64 + p0] & 0x7f) : -1;
        const int r1 = (ch_ok && on1) ? (int)(tail[hpar * 64 + p1] & 0x7f) : -1;
        float4 v0, v1 = make_float4(0.f, 0.f, 0.f, 0.f);
        v0.x = tile[(cb + 0) * kPP + p0]; v0.y = tile[(cb + 1) * kPP + p0];
        v0.z = tile[(cb + 2) * kPP + p0]; v0.w = tile[(cb + 3) * kPP + p0];
        if (on1) {
          v1.x = tile[(cb + 0) * kPP + p1]; v1.y = tile[(cb + 1) * kPP + p1];
          v1.z = tile[(cb + 2) * kPP + p1]; v1.w = tile[(cb + 3) * kPP + p1];
        }
        const float ly0 = __uint_as_float(e0.z), lx0 = __uint_as_float(e0.w);
        const float ly1 = __uint_as_float(e1.z), lx1 = __uint_as_float(e1.w);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);     // everything of this stage is in registers now
        const uint32_t pair0 = top0 ? e0.x : e0.y, pair1 = top1 ? e1.x : e1.y;   // slots of (x0, x0+1) in my row
        const float wy0 = top0 ? 1.f - ly0 : ly0, wy1 = top1 ? 1.f - ly1 : ly1;
        // Equal-rank samples differ in (row, x0) but may still meet in a pixel through DIFFERENT corners
        // (A's right pixel is B's left pixel): left and right corners are two read-modify-write phases with a
        // warp barrier between them.  Corners outside the map point at the zero slot, never written back.
        auto rmw = [&](bool on, uint32_t slot, float w, const float4& v) {
          if (on) {
            float4 m = map4[slot ^ cg];
            m.x = fmaf(w, v.x, m.x);
            m.y = fmaf(w, v.y, m.y);
            m.z = fmaf(w, v.z, m.z);
            m.w = fmaf(w, v.w, m.w);
            map4[slot ^ cg] = m;
          }
        };
        for (int rr = 0; rr <= mr; ++rr) {
          const bool a0 = r0 == rr, a1 = r1 == rr;
          rmw(a0, pair0 & 0xffffu, wy0 * (1.f - lx0), v0);
          rmw(a1, pair1 & 0xffffu, wy1 * (1.f - lx1), v1);
          __syncwarp();
          rmw(a0, pair0 >> 16, wy0 * lx0, v0);
          rmw(a1, pair1 >> 16, wy1 * lx1, v1);
          __syncwarp();
        }
      } else {
        // max-pool: the 4 channels of a lane may route to different samples of the 2x2 block, so they
        // are 4 scalar streams; collisions (same slot, same channel) are ranked on the fly by match.any.
        const int n = *reinterpret_cast<const int*>(rec + (size_t)g.S * g.S * 16 + 56);
        const uint8_t* at = arg_smem ? atiles + (size_t)s * TILE : argmax + ((size_t)n * g.C + c0) * kPP;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          const int p = pass * 32 + lane;
          const bool act = p < kPP && ch_ok;
          const int pi = p / 7, pj = p - pi * 7;
#pragma unroll 1
          for (int e = 0; e < 4; ++e) {
            float val = 0.f;
            uint4 en = make_uint4(0, 0, 0, 0);
            if (act) {
              val = tile[(cb + e) * kPP + p];
              const int a = at[(cb + e) * kPP + p] & 3;
              en = ent[(2 * pi + (a >> 1)) * 14 + 2 * pj + (a & 1)];
            }
            const float ly = __uint_as_float(en.z), lx = __uint_as_float(en.w);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t slot = (c == 0) ? (en.x & 0xffffu) : (c == 1) ? (en.x >> 16) : (c == 2) ? (en.y & 0xffffu) : (en.y >> 16);
              const float w = ((c & 2) ? ly : 1.f - ly) * ((c & 1) ? lx : 1.f - lx);
              const bool ok = act && (int)slot < zs;
              const int key = ok ? (int)slot : -1 - lane;
              const unsigned grp = __match_any_sync(0xffffffffu, key);
              const int rank = __popc(grp & lt);
              const int mr = __reduce_max_sync(0xffffffffu, ok ? rank : 0);
              for (int rr = 0; rr <= mr; ++rr) {
                if (ok && rank == rr) {
                  float* mp = map + (size_t)((slot ^ cg) << 2) + e;
                  *mp = fmaf(w, val, *mp);
                }
                __syncwarp();
              }
            }
          }
        }
      }
      if (MAXPOOL) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
      }
    }
  }
  __syncthreads();
  unstage_map<CC>(map, dbottom + (size_t)b * g.C * HW, c0, g.C, HW);
}

// ------------------------------------------------------------------ backward, row-owner variant (CC = 32, 7x7)
// lane = channel, warp w owns map rows y = w (mod kRowWarps): every read-modify-write of the accumulator map is a
// conflict-free 128-byte segment, control flow is warp uniform (geometry does not depend on the channel), there are
// no collisions to rank and no barriers inside a ROI.  The accumulator is stored [row][col+2][33]: the odd channel
// stride keeps both the channel-wise updates and the pixel-wise final read-out conflict free, and two guard columns
// on each side absorb the corners that fall outside the map (x0 is clamped to [-2, W] in the record), so the inner
// loop has no bounds checks.
//
// What bounds it (ncu + scripts/probes/tma_bulk_probe.cu, profiles/r01_tma_bulk_probe.txt): not bandwidth but the
// serial mbarrier round trips.  One wait -> arrive hand-off costs ~400 clk on the thread that does it, so a ring that
// moves ONE 6272-byte tile per barrier tops out at ~15 B/clk/SM even with nothing else to do, and the first versions
// spent ~950 clk per ROI.  Hence:
//   * a stage of the ring carries FOUR ROIs (25 KB: 4 tiles + their 4 records, 5 bulk copies, one barrier round
//     trip) -- the probe sustains 25.6 B/clk/SM at that size, above the HBM share of an SM;
//   * the per-ROI bookkeeping is done ONCE by roi_geom_kernel and shipped in the record: wmask[w] says which
//     (sample row, upper/lower) pairs land in warp w's rows (a warp reads one 16-bit word per ROI and skips the ROI
//     when it is zero -- no row tests, modulos, ballots or shuffles in the loop);
//   * sample rows that fall into the same map row are first combined in registers (v_j = sum_i wy_i g_ij: the column
//     geometry does not depend on i) and written with ONE set of 14 read-modify-writes;
//   * mode: whether the 14 corner columns are pairwise distinct (one batch of 14 independent loads / FMAs / stores),
//     or x0 is strictly increasing (two batches: left corners, right corners), or neither (serial chain).
//     (Merging shared columns in registers -- one RMW per distinct column, pattern shipped in the record -- was built and
//     measured: 1.35 -> 1.63 ms.  A warp sees less than one map row per ROI on average, so the instructions that decode
//     the pattern per (ROI, row) cost more issue slots than the saved wavefronts return; profiles/r02_ab.md.)
// Bit-reproducible (fixed order, no atomics).
constexpr int kRowStages = 3;
constexpr int kRowGroup = 4;        // ROIs per stage

__global__ void __launch_bounds__((kRowWarps + 1) * 32, 1)
roi_crop_bwd_rows_kernel(const float* __restrict__ dout, const int* __restrict__ seg,
                         const SepRec* __restrict__ sep, float* __restrict__ dbottom, CropGeom g) {
  constexpr int CC = 32;
  constexpr int TILE = CC * kPP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W, W = g.W, H = g.H;
  const int WP = W + 4;                                                      // padded row length
  float* map = reinterpret_cast<float*>(smem_raw);                           // [H][WP][33]
  const int map_floats = (H * WP * kRowLd + 3) & ~3;
  float* tiles = map + map_floats;                                           // [kRowStages][kRowGroup][TILE]
  SepRec* recs = reinterpret_cast<SepRec*>(tiles + kRowStages * kRowGroup * TILE);   // [kRowStages][kRowGroup]
  __shared__ uint64_t full_bar[kRowStages], empty_bar[kRowStages];

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cvalid = min(CC, g.C - c0);

  for (int i = t; i < map_floats; i += blockDim.x) map[i] = 0.f;
  if (t == 0) {
    for (int s = 0; s < kRowStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kRowWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int beg = seg[b], end = seg[b + 1];
  const int nst = (end - beg + kRowGroup - 1) / kRowGroup;
  const uint32_t tile_bytes = (uint32_t)(cvalid * kPP * sizeof(float));

  if (wid == kRowWarps) {
    // ---------------- producer warp (ROI ids fetched 32 at a time = 8 stages) ----------------
    int sidx = 0, lap = 0;
    for (int r0 = beg; r0 < end; r0 += 32) {
      const int mine = r0 + lane;
      const int nl = (mine < end) ? __ldg(&sep[mine].n) : 0;
      {   // L2 prefetch of the group after this one (and of the first group itself)
        const int ahead = mine + (r0 == beg ? 0 : 32);
        for (int a = ahead; a < end && a <= mine + 32; a += 32)
          bulk_prefetch_l2(dout + ((size_t)__ldg(&sep[a].n) * g.C + c0) * kPP, tile_bytes);
      }
      const int cnt = min(32, end - r0);
      for (int j = 0; j < cnt; j += kRowGroup) {
        const int gc = min(kRowGroup, cnt - j);               // ROIs in this stage
        int n[kRowGroup];
#pragma unroll
        for (int q = 0; q < kRowGroup; ++q) n[q] = __shfl_sync(0xffffffffu, nl, (j + q) & 31);
        if (lane == 0) {
          if (lap > 0) mbar_wait(&empty_bar[sidx], (lap - 1) & 1);
          mbar_arrive_expect_tx(&full_bar[sidx], (uint32_t)gc * (tile_bytes + (uint32_t)sizeof(SepRec)));
          bulk_g2s(recs + sidx * kRowGroup, sep + r0 + j, (uint32_t)(gc * sizeof(SepRec)), &full_bar[sidx]);
#pragma unroll
          for (int q = 0; q < kRowGroup; ++q)
            if (q < gc)
              bulk_g2s(tiles + (size_t)(sidx * kRowGroup + q) * TILE, dout + ((size_t)n[q] * g.C + c0) * kPP, tile_bytes,
                       &full_bar[sidx]);
        }
        if (++sidx == kRowStages) { sidx = 0; ++lap; }
      }
    }
  } else {
    // ---------------- consumers: warp = rows, lane = channel ----------------
    const int cl = lane < cvalid ? lane : 0;          // idle lanes shadow channel 0 into their own column
    float* mlane = map + 2 * kRowLd + lane;           // (row 0, col 0, my channel); lanes >= cvalid never unstaged
    const int row_stride = WP * kRowLd;
    int sidx = 0, lap = 0;
    for (int k = 0; k < nst; ++k) {
      mbar_wait_parked(&full_bar[sidx], lap & 1);
      const int gc = min(kRowGroup, end - beg - k * kRowGroup);
      // lane q < gc fetches this warp's hit word of ROI q of the stage; only the ROIs with work are visited
      const unsigned mymask = (lane < gc) ? (unsigned)recs[sidx * kRowGroup + lane].wmask[wid] : 0u;
      unsigned act = __ballot_sync(0xffffffffu, mymask != 0u);
      while (act) {
        const int q = __ffs(act) - 1;
        act &= act - 1;
        const SepRec* rec = recs + sidx * kRowGroup + q;
        unsigned hits = __shfl_sync(0xffffffffu, mymask, q);      // same word for every lane: uniform control flow below
        int xoff[7];
        float wa[7], wb[7];
        {
          const int4 xa = *reinterpret_cast<const int4*>(rec->xoff), xb = *reinterpret_cast<const int4*>(rec->xoff + 4);
          const float4 la = *reinterpret_cast<const float4*>(rec->lx), lb = *reinterpret_cast<const float4*>(rec->lx + 4);
          xoff[0] = xa.x; xoff[1] = xa.y; xoff[2] = xa.z; xoff[3] = xa.w; xoff[4] = xb.x; xoff[5] = xb.y; xoff[6] = xb.z;
          wb[0] = la.x; wb[1] = la.y; wb[2] = la.z; wb[3] = la.w; wb[4] = lb.x; wb[5] = lb.y; wb[6] = lb.z;
#pragma unroll
          for (int j = 0; j < 7; ++j) wa[j] = 1.f - wb[j];
        }
        const int mode = rec->mode;
        const float* tcol = tiles + (size_t)(sidx * kRowGroup + q) * TILE + (size_t)cl * kPP;
        while (hits) {
          const int bsel = __ffs(hits) - 1;
          hits &= hits - 1;
          const int yy = (int)rec->y0[bsel >> 1] + (bsel & 1);
          float v[7];
          {
            const float ly = rec->ly[bsel >> 1];
            const float wy = (bsel & 1) ? ly : 1.f - ly;
            const float* trow = tcol + (bsel >> 1) * 7;
#pragma unroll
            for (int j = 0; j < 7; ++j) v[j] = wy * trow[j];
          }
          // further sample rows of this ROI in the same map row (adjacent bits: the rows are monotone in i)
          while (hits) {
            const int b2 = __ffs(hits) - 1;
            if ((int)rec->y0[b2 >> 1] + (b2 & 1) != yy) break;
            hits &= hits - 1;
            const float ly = rec->ly[b2 >> 1];
            const float wy = (b2 & 1) ? ly : 1.f - ly;
            const float* trow = tcol + (b2 >> 1) * 7;
#pragma unroll
            for (int j = 0; j < 7; ++j) v[j] = fmaf(wy, trow[j], v[j]);
          }
          float* mrow = mlane + yy * row_stride;
          if (mode == 2) {
            float m0[7], m1[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) { m0[j] = mrow[xoff[j]]; m1[j] = mrow[xoff[j] + kRowLd]; }
#pragma unroll
            for (int j = 0; j < 7; ++j) { m0[j] = fmaf(wa[j], v[j], m0[j]); m1[j] = fmaf(wb[j], v[j], m1[j]); }
#pragma unroll
            for (int j = 0; j < 7; ++j) { mrow[xoff[j]] = m0[j]; mrow[xoff[j] + kRowLd] = m1[j]; }
          } else if (mode == 1) {
            float m0[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) m0[j] = mrow[xoff[j]];
#pragma unroll
            for (int j = 0; j < 7; ++j) m0[j] = fmaf(wa[j], v[j], m0[j]);
#pragma unroll
            for (int j = 0; j < 7; ++j) mrow[xoff[j]] = m0[j];
#pragma unroll
            for (int j = 0; j < 7; ++j) m0[j] = mrow[xoff[j] + kRowLd];
#pragma unroll
            for (int j = 0; j < 7; ++j) m0[j] = fmaf(wb[j], v[j], m0[j]);
#pragma unroll
            for (int j = 0; j < 7; ++j) mrow[xoff[j] + kRowLd] = m0[j];
          } else {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              float* m = mrow + xoff[j];
              m[0] = fmaf(wa[j], v[j], m[0]);
              m[kRowLd] = fmaf(wb[j], v[j], m[kRowLd]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[sidx]);
      if (++sidx == kRowStages) { sidx = 0; ++lap; }
    }
  }
  __syncthreads();
  // read-out: warp = (channel, 32 consecutive pixels of a row)
  float* dst = dbottom + ((size_t)b * g.C + c0) * HW;
  const int nw = blockDim.x >> 5;
  const int xblocks = (W + 31) >> 5;
  for (int it = wid; it < cvalid * H * xblocks; it += nw) {
    const int c = it % cvalid, rest = it / cvalid;
    const int y = rest / xblocks, x = (rest % xblocks) * 32 + lane;
    if (x < W) dst[(size_t)c * HW + y * W + x] = map[((size_t)y * WP + x + 2) * kRowLd + c];
  }
}

// ------------------------------------------------------------------ backward, generic row-owner variant
// The same ownership idea for the cases the kernel above does not take: the 14x14-sample + 2x2-max crop
// (network_cycle_response.py:140-144; cfg-3 and the VGG path of cfg-5) and 7x7 crops of maps too large for a
// 32-channel accumulator.  The max-pool backward becomes channel independent again by EXPANDING the pooled gradient
// to the 2S x 2S sample grid on the fly: gs[c][i][j] = (winner[c][i/2][j/2] == (i%2, j%2)) ? g[c][i/2][j/2] : 0 -- the
// geometry of sample (i, j) does not depend on the channel, only the VALUE does, so lane = channel keeps control flow
// uniform and the per-lane winner is a select, not a branch.  (The previous kernel -- lanes = samples, four scalar
// channel streams, match.any ranking of colliding corners -- ran at 1 % of the HBM roofline: 17.8 ms at cfg-3.)
//   * CC = 32 or 16 channels per CTA (the accumulator [H][W+4][CC+1] must fit in shared memory: 32x32 maps take 32,
//     37x62 / 38x63 maps take 16); with CC = 16 a warp holds two row owners (half-warps), so the owner count doubles;
//   * per owned map row the contributions of all sample rows that land in it are combined in registers
//     (v_j = sum_i wy_i gs_ij), then adjacent sample columns that share a map column are merged in registers as well
//     (x0_j is non-decreasing in j: a sliding two-column window), so one (ROI, map row) costs one read-modify-write
//     per DISTINCT map column (box width + 1), software pipelined two deep;
//   * a stage of the TMA ring carries kRowGroup ROIs: gradient tiles, winner tiles and records.
// Bit-reproducible (fixed order, no atomics).
struct __align__(16) SepRecX {
  int16_t xoff[16];     // (clamped x0_j) * (CC+1): float offset of the left corner inside a padded accumulator row
  float lx[16];
  float ly[16];
  int16_t y0[16];       // clamped to [-2, 30000]
  uint32_t wmask[64];   // per row owner: bit 2i+d set <=> map row y0_i + d is inside the map and owned by it
  int n;                // ROI index
  int pad[3];
};
static_assert(sizeof(SepRecX) == 464, "SepRecX layout");

__global__ void __launch_bounds__(64)
roi_sepx_kernel(const float* __restrict__ rois, const int* __restrict__ order, const int* __restrict__ seg,
                SepRecX* __restrict__ sep, CropGeom g, int LD, int owners) {
  const int r = blockIdx.x, p = threadIdx.x;
  if (r >= seg[g.B]) return;                 // ROIs whose batch index is outside [0,B) are not in `order`
  const int n = __ldg(order + r);
  const float* rp = rois + 5 * (size_t)n;
  float4 box;
  box.x = __ldg(rp + 1) * g.sx;
  box.y = __ldg(rp + 2) * g.sy;
  box.z = __ldg(rp + 3) * g.sx;
  box.w = __ldg(rp + 4) * g.sy;
  const float inv = 1.0f / (float)(g.S - 1);
  __shared__ int s_y0[16];
  SepRecX* sr = sep + r;
  if (p < 16) {
    const Corner c = sample_at(box, min(p, g.S - 1), min(p, g.S - 1), inv);
    const int y0 = min(max(c.y0, -2), 30000);
    s_y0[p] = y0;
    sr->xoff[p] = (int16_t)(min(max(c.x0, -2), g.W) * LD);
    sr->lx[p] = c.lx;
    sr->ly[p] = c.ly;
    sr->y0[p] = (int16_t)y0;
  }
  __syncthreads();
  unsigned m = 0;
  if (p < owners) {
    for (int i = 0; i < g.S; ++i)
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int row = s_y0[i] + d;
        if (row >= 0 && row < g.H && (row % owners) == p) m |= 1u << (2 * i + d);
      }
  }
  sr->wmask[p] = m;
  if (p == 0) {
    sr->n = n;
    sr->pad[0] = sr->pad[1] = sr->pad[2] = 0;
  }
}

constexpr int kRowXMaxStages = 4;
// ROIs per ring stage.  Finer stages (2 ROIs x up to 8 stages) and a sleep-with-back-off wait instead of the parked
// try_wait were both measured and changed nothing (cfg-3 2.69 vs 2.64 ms, cfg-5 3.47 vs 3.29 ms): the spinning warps are
// the ones that are AHEAD; the kernel is bound by the row owners with the most hits per stage, whose per-row chain
// (record fields -> tile values -> column merge -> read-modify-write) is latency bound.
constexpr int kRowXGroup = 4;

template <int CC, int S, bool MAXPOOL>
__global__ void __launch_bounds__((kRowWarps + 1) * 32, 1)
roi_crop_bwd_rowsx_kernel(const float* __restrict__ dout, const int* __restrict__ seg, const SepRecX* __restrict__ sep,
                          const uint8_t* __restrict__ argmax, float* __restrict__ dbottom, CropGeom g, int nstages) {
  constexpr int LD = CC + 1;
  constexpr int SUB = 32 / CC;                 // row owners per warp
  constexpr int OWNERS = kRowWarps * SUB;
  constexpr int TILE = CC * kPP;               // floats (gradients) / bytes (winners) per ROI tile
  constexpr int ATILE = (TILE + 15) & ~15;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W, W = g.W, H = g.H;
  const int WP = W + 4;                                                      // padded row length
  float* map = reinterpret_cast<float*>(smem_raw);                           // [H][WP][LD]
  const int map_floats = (H * WP * LD + 3) & ~3;
  float* tiles = map + map_floats;                                           // [nstages][kRowXGroup][TILE]
  SepRecX* recs = reinterpret_cast<SepRecX*>(tiles + (size_t)nstages * kRowXGroup * TILE);   // [nstages][kRowXGroup]
  uint8_t* atiles = reinterpret_cast<uint8_t*>(recs + (size_t)nstages * kRowXGroup);         // [nstages][kRowXGroup][ATILE]
  __shared__ uint64_t full_bar[kRowXMaxStages], empty_bar[kRowXMaxStages];

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cvalid = min(CC, g.C - c0);

  for (int i = t; i < map_floats; i += blockDim.x) map[i] = 0.f;
  if (t == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kRowWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int beg = seg[b], end = seg[b + 1];
  const int nst = (end - beg + kRowXGroup - 1) / kRowXGroup;
  const uint32_t tile_bytes = (uint32_t)(cvalid * kPP * sizeof(float));
  const uint32_t arg_bytes = (uint32_t)(cvalid * kPP);
  // winners travel by TMA only when their rows are 16-byte sized / aligned; otherwise they are read in place
  const bool arg_smem = MAXPOOL && (g.C % 16 == 0) && (cvalid % 16 == 0);

  if (wid == kRowWarps) {
    // ---------------- producer warp (ROI ids fetched 32 at a time) ----------------
    int sidx = 0, lap = 0;
    for (int r0 = beg; r0 < end; r0 += 32) {
      const int mine = r0 + lane;
      const int nl = (mine < end) ? __ldg(&sep[mine].n) : 0;
      {   // L2 prefetch of the group after this one (and of the first group itself)
        const int ahead = mine + (r0 == beg ? 0 : 32);
        for (int a = ahead; a < end && a <= mine + 32; a += 32)
          bulk_prefetch_l2(dout + ((size_t)__ldg(&sep[a].n) * g.C + c0) * kPP, tile_bytes);
      }
      const int cnt = min(32, end - r0);
      for (int j = 0; j < cnt; j += kRowXGroup) {
        const int gc = min(kRowXGroup, cnt - j);               // ROIs in this stage
        int n[kRowXGroup];
#pragma unroll
        for (int q = 0; q < kRowXGroup; ++q) n[q] = __shfl_sync(0xffffffffu, nl, (j + q) & 31);
        if (lane == 0) {
          if (lap > 0) mbar_wait(&empty_bar[sidx], (lap - 1) & 1);
          mbar_arrive_expect_tx(&full_bar[sidx],
                                (uint32_t)gc * (tile_bytes + (uint32_t)sizeof(SepRecX) + (arg_smem ? arg_bytes : 0u)));
          bulk_g2s(recs + sidx * kRowXGroup, sep + r0 + j, (uint32_t)(gc * sizeof(SepRecX)), &full_bar[sidx]);
#pragma unroll
          for (int q = 0; q < kRowXGroup; ++q)
            if (q < gc) {
              bulk_g2s(tiles + (size_t)(sidx * kRowXGroup + q) * TILE, dout + ((size_t)n[q] * g.C + c0) * kPP, tile_bytes,
                       &full_bar[sidx]);
              if (arg_smem)
                bulk_g2s(atiles + (size_t)(sidx * kRowXGroup + q) * ATILE, argmax + ((size_t)n[q] * g.C + c0) * kPP,
                         arg_bytes, &full_bar[sidx]);
            }
        }
        if (++sidx == nstages) { sidx = 0; ++lap; }
      }
    }
  } else {
    // ---------------- consumers: (half-)warp = rows, lane = channel ----------------
    const int ch = lane % CC, sub = lane / CC;
    const int owner = wid * SUB + sub;
    const int cl = ch < cvalid ? ch : 0;              // idle lanes shadow channel 0 into their own column
    float* mlane = map + 2 * LD + ch;                 // (row 0, col 0, my channel); channels >= cvalid never unstaged
    const int row_stride = WP * LD;
    int sidx = 0, lap = 0;
    for (int k = 0; k < nst; ++k) {
      mbar_wait_parked(&full_bar[sidx], lap & 1);
      const int gc = min(kRowXGroup, end - beg - k * kRowXGroup);
      for (int q = 0; q < gc; ++q) {
        const SepRecX* rec = recs + sidx * kRowXGroup + q;
        unsigned hits = rec->wmask[owner];
        if (__ballot_sync(0xffffffffu, hits != 0u) == 0u) continue;        // warp uniform
        // column geometry of the ROI (the same for every lane)
        int xoff[S];
        float wa[S], wb[S];
#pragma unroll
        for (int j = 0; j < S; ++j) {
          xoff[j] = (int)rec->xoff[j];
          wb[j] = rec->lx[j];
          wa[j] = 1.f - wb[j];
        }
        const float* tcol = tiles + (size_t)(sidx * kRowXGroup + q) * TILE + (size_t)cl * kPP;
        const uint8_t* acol = nullptr;
        if (MAXPOOL)
          acol = arg_smem ? atiles + (size_t)(sidx * kRowXGroup + q) * ATILE + (size_t)cl * kPP
                          : argmax + ((size_t)rec->n * g.C + c0 + cl) * kPP;
        while (hits) {
          const int bsel = __ffs(hits) - 1;
          const int yy = (int)rec->y0[bsel >> 1] + (bsel & 1);
          float v[S];
#pragma unroll
          for (int j = 0; j < S; ++j) v[j] = 0.f;
          // every sample row of this ROI that lands in map row yy (adjacent bits of this owner's mask).  (Folding the two
          // sample rows of a pooled row into one pass -- winner row bit picks the weight -- measured slower: 2.83 vs 2.64 ms.)
          while (hits) {
            const int b2 = __ffs(hits) - 1;
            const int i = b2 >> 1;
            if ((int)rec->y0[i] + (b2 & 1) != yy) break;
            hits &= hits - 1;
            const float ly = rec->ly[i];
            const float wy = (b2 & 1) ? ly : 1.f - ly;
            if (MAXPOOL) {
              const float* trow = tcol + (i >> 1) * 7;
              const uint8_t* arow = acol + (i >> 1) * 7;
              const int code = (i & 1) << 1;
#pragma unroll
              for (int pj = 0; pj < 7; ++pj) {
                const float gq = wy * trow[pj];
                const int a = (int)arow[pj] & 3;
                v[2 * pj] += (a == code) ? gq : 0.f;
                v[2 * pj + 1] += (a == code + 1) ? gq : 0.f;
              }
            } else {
              const float* trow = tcol + i * 7;
#pragma unroll
              for (int j = 0; j < S; ++j) v[j] = fmaf(wy, trow[j], v[j]);
            }
          }
          // scatter with column merging: window (a0 @ column cx, a1 @ cx + 1); flushes are pipelined two deep
          float* mrow = mlane + yy * row_stride;
          float* pp = nullptr;
          float pv = 0.f, pm = 0.f;
          auto flush = [&](int off, float val) {
            float* addr = mrow + off;
            const float m = *addr;                  // the new column's load is in flight ...
            if (pp) *pp = pm + pv;                  // ... while the previous (different) column completes
            pp = addr; pv = val; pm = m;
          };
          auto drain = [&]() {
            if (pp) *pp = pm + pv;
            pp = nullptr;
          };
          int cx = xoff[0];
          float a0 = wa[0] * v[0], a1 = wb[0] * v[0];
#pragma unroll
          for (int j = 1; j < S; ++j) {
            const int dx = xoff[j] - cx;              // warp uniform
            if (dx == 0) {
              a0 = fmaf(wa[j], v[j], a0);
              a1 = fmaf(wb[j], v[j], a1);
            } else if (dx == LD) {
              flush(cx, a0);
              a0 = fmaf(wa[j], v[j], a1);
              a1 = wb[j] * v[j];
              cx = xoff[j];
            } else {
              flush(cx, a0);
              flush(cx + LD, a1);
              if (dx < 0) drain();                    // degenerate box (x2 < x1): columns are no longer increasing
              a0 = wa[j] * v[j];
              a1 = wb[j] * v[j];
              cx = xoff[j];
            }
          }
          flush(cx, a0);
          flush(cx + LD, a1);
          drain();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[sidx]);
      if (++sidx == nstages) { sidx = 0; ++lap; }
    }
    (void)owner;
  }
  __syncthreads();
  // read-out: warp = (channel, 32 consecutive pixels of a row)
  float* dst = dbottom + ((size_t)b * g.C + c0) * HW;
  const int nw = blockDim.x >> 5;
  const int xblocks = (W + 31) >> 5;
  for (int it = wid; it < cvalid * H * xblocks; it += nw) {
    const int c = it % cvalid, rest = it / cvalid;
    const int y = rest / xblocks, x = (rest % xblocks) * 32 + lane;
    if (x < W) dst[(size_t)c * HW + y * W + x] = map[((size_t)y * WP + x + 2) * LD + c];
  }
}

// ------------------------------------------------------------------ forward, row-mirror variant
// The forward seen the same way as the row-owner backward: lane = channel, the map slice sits in shared memory as
// [row][col + 2][CC + 1] (conflict-free for channel-wise access, zero guard columns absorb corners outside the map),
// and the geometry is the SEPARABLE record (14 or 7 column offsets / fractions, as many row offsets / fractions).  A
// sample row is interpolated as  T[x] = (1 - ly) M[y0][x] + ly M[y0+1][x]  on the distinct map columns it touches (x0_j
// is non-decreasing in j: a sliding two-column window), then  s_j = (1 - lx_j) T[x0_j] + lx_j T[x0_j + 1]  from
// registers -- (box width + 1) column loads per sample row instead of four corner loads per sample, no per-lane
// geometry at all (control flow is warp uniform), and for the 2x2 max variant the two sample rows of a pooled row stay
// in registers.  The table-driven kernels above spend 16 LDS.128 and ~170 instructions per pooled float4; this one
// needs ~2.5x fewer shared-memory wavefronts and ~5x fewer instructions (ncu, profiles/).
// A warp owns whole ROIs; its [CC][49] output tile (and winner tile) leaves by TMA bulk stores.  With CC = 16 the two
// half-warps take alternate pooled rows of the same ROI (same column geometry => still uniform).
constexpr int kFwdRowWarps = 8;

template <int CC, int S, bool MAXPOOL>
__global__ void __launch_bounds__(kFwdRowWarps * 32, 1)
roi_crop_fwd_rows_kernel(const float* __restrict__ bottom, const int* __restrict__ seg, const SepRecX* __restrict__ sep,
                         float* __restrict__ out, uint8_t* __restrict__ argmax, CropGeom g) {
  constexpr int LD = CC + 1;
  constexpr int SUB = 32 / CC;
  constexpr int TILE = CC * kPP;
  constexpr int ATILE = (TILE + 15) & ~15;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W, W = g.W, H = g.H;
  const int WP = W + 4;
  float* map = reinterpret_cast<float*>(smem_raw);                           // [H][WP][LD]
  const int map_floats = (H * WP * LD + 3) & ~3;
  float* tiles = map + map_floats;                                           // [kFwdRowWarps][TILE]
  uint8_t* atiles = reinterpret_cast<uint8_t*>(tiles + kFwdRowWarps * TILE); // [kFwdRowWarps][ATILE]

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cvalid = min(CC, g.C - c0);
  const int beg = seg[b], end = seg[b + 1];

  // ---- stage the map slice: guard columns zero, lanes = consecutive pixels of a row (coalesced), loop over channels
  for (int i = t; i < map_floats; i += blockDim.x) map[i] = 0.f;
  __syncthreads();
  {
    const float* src = bottom + ((size_t)b * g.C + c0) * HW;
    const int xblocks = (W + 31) >> 5;
    for (int it = wid; it < cvalid * H * xblocks; it += kFwdRowWarps) {
      const int c = it % cvalid, rest = it / cvalid;
      const int y = rest / xblocks, x = (rest % xblocks) * 32 + lane;
      if (x < W) map[((size_t)y * WP + x + 2) * LD + c] = __ldg(src + (size_t)c * HW + y * W + x);
    }
  }
  __syncthreads();

  const int ch = lane % CC, sub = lane / CC;
  const float* mlane = map + 2 * LD + ch;             // (row 0, col 0, my channel)
  const int row_stride = WP * LD;
  float* tile = tiles + (size_t)wid * TILE;
  uint8_t* atile = atiles + (size_t)wid * ATILE;
  const uint32_t tile_bytes = (uint32_t)(cvalid * kPP * sizeof(float));
  const bool arg_bulk = MAXPOOL && argmax != nullptr && (g.C % 16 == 0) && (cvalid % 16 == 0);

  struct Rec {
    int4 xo[2];      // 16 x int16
    float4 lx[4], ly[4];
    int4 y0[2];
    int n;
  };
  auto load_rec = [&](int r, Rec& R) {
    const SepRecX* rp = sep + r;
    R.xo[0] = __ldg(reinterpret_cast<const int4*>(rp->xoff));
    R.xo[1] = __ldg(reinterpret_cast<const int4*>(rp->xoff) + 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      R.lx[k] = __ldg(reinterpret_cast<const float4*>(rp->lx) + k);
      R.ly[k] = __ldg(reinterpret_cast<const float4*>(rp->ly) + k);
    }
    R.y0[0] = __ldg(reinterpret_cast<const int4*>(rp->y0));
    R.y0[1] = __ldg(reinterpret_cast<const int4*>(rp->y0) + 1);
    R.n = __ldg(&rp->n);
  };
  auto i16 = [](const int4 (&v)[2], int k) -> int {       // element k of 16 packed int16 (k is a compile-time constant)
    const int w = k >> 1;
    const int word = (w == 0) ? v[0].x : (w == 1) ? v[0].y : (w == 2) ? v[0].z : (w == 3) ? v[0].w
                   : (w == 4) ? v[1].x : (w == 5) ? v[1].y : (w == 6) ? v[1].z : v[1].w;
    return (k & 1) ? (word >> 16) : (int)(short)(word & 0xffff);
  };
  auto f16 = [](const float4 (&v)[4], int k) -> float {
    const float4 q = v[k >> 2];
    return (k & 3) == 0 ? q.x : (k & 3) == 1 ? q.y : (k & 3) == 2 ? q.z : q.w;
  };

  Rec cur, nxt;
  int r = beg + wid;
  if (r < end) load_rec(r, cur);
  for (; r < end; r += kFwdRowWarps) {
    const bool more = r + kFwdRowWarps < end;
    if (more) load_rec(r + kFwdRowWarps, nxt);              // in flight during this ROI
    int xoff[S];
    float wa[S], wb[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
      xoff[j] = i16(cur.xo, j);
      wb[j] = f16(cur.lx, j);
      wa[j] = 1.f - wb[j];
    }
    if (lane == 0) bulk_wait_read<0>();                      // my previous stores have left the tiles
    __syncwarp();
    // one sample row: s[j], j < S, for the map rows (y0, y0 + 1) with fraction ly
    auto sample_row = [&](int y0c, float ly, float (&sv)[S]) {
      // rows outside the map contribute zero: clamp the index, zero the weight
      const bool v0 = (unsigned)y0c < (unsigned)H, v1 = (unsigned)(y0c + 1) < (unsigned)H;
      const float w0 = v0 ? 1.f - ly : 0.f, w1 = v1 ? ly : 0.f;
      const float* r0 = mlane + (v0 ? y0c : 0) * row_stride;
      const float* r1 = mlane + (v1 ? y0c + 1 : 0) * row_stride;
      auto col = [&](int off) { return fmaf(w1, r1[off], w0 * r0[off]); };
      int cx = xoff[0];
      float tl = col(cx), tr = col(cx + LD);
      sv[0] = fmaf(wb[0], tr, wa[0] * tl);
#pragma unroll
      for (int j = 1; j < S; ++j) {
        const int dx = xoff[j] - cx;               // warp uniform
        if (dx == LD) {
          tl = tr;
          tr = col(xoff[j] + LD);
        } else if (dx != 0) {
          tl = col(xoff[j]);
          tr = col(xoff[j] + LD);
        }
        cx = xoff[j];
        sv[j] = fmaf(wb[j], tr, wa[j] * tl);
      }
    };
    constexpr int ROWS = 7;
    for (int pi = sub; pi < ROWS; pi += SUB) {
      if (MAXPOOL) {
        float st[S], sb[S];
        // 2 pi and 2 pi + 1 are not compile-time constants: fetch the row geometry through a small switch-free select
        int y0t = 0, y0b = 0;
        float lyt = 0.f, lyb = 0.f;
#pragma unroll
        for (int k = 0; k < ROWS; ++k)
          if (k == pi) {
            y0t = i16(cur.y0, 2 * k); y0b = i16(cur.y0, 2 * k + 1);
            lyt = f16(cur.ly, 2 * k); lyb = f16(cur.ly, 2 * k + 1);
          }
        sample_row(y0t, lyt, st);
        sample_row(y0b, lyb, sb);
#pragma unroll
        for (int pj = 0; pj < 7; ++pj) {
          // first strict maximum in row-major order of the 2x2 block wins (max_pool2d's backward)
          float m = st[2 * pj];
          int am = 0;
          if (st[2 * pj + 1] > m) { m = st[2 * pj + 1]; am = 1; }
          if (sb[2 * pj] > m) { m = sb[2 * pj]; am = 2; }
          if (sb[2 * pj + 1] > m) { m = sb[2 * pj + 1]; am = 3; }
          if (ch < cvalid) {
            tile[ch * kPP + pi * 7 + pj] = m;
            atile[ch * kPP + pi * 7 + pj] = (uint8_t)am;
          }
        }
      } else {
        float sv[S];
        int y0c = 0;
        float ly = 0.f;
#pragma unroll
        for (int k = 0; k < ROWS; ++k)
          if (k == pi) { y0c = i16(cur.y0, k); ly = f16(cur.ly, k); }
        sample_row(y0c, ly, sv);
#pragma unroll
        for (int j = 0; j < 7; ++j)
          if (ch < cvalid) tile[ch * kPP + pi * 7 + j] = sv[j];
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    const size_t obase = ((size_t)cur.n * g.C + c0) * kPP;
    if (lane == 0) {
      bulk_s2g(out + obase, tile, tile_bytes);
      if (arg_bulk) bulk_s2g(argmax + obase, atile, (uint32_t)(cvalid * kPP));
      bulk_commit();
    }
    if (MAXPOOL && argmax != nullptr && !arg_bulk)
      for (int i = lane; i < cvalid * kPP; i += 32) argmax[obase + i] = atile[i];
    if (more) cur = nxt;
  }
  if (lane == 0) bulk_wait<0>();
}

size_t fwd_rows_smem(int H, int W, int cc, bool maxpool) {
  const size_t map = (((size_t)H * (W + 4) * (cc + 1) + 3) & ~(size_t)3) * 4;
  const size_t tile = (size_t)cc * kPP;
  return map + kFwdRowWarps * (tile * 4 + (maxpool ? ((tile + 15) & ~(size_t)15) : 0)) + 128;
}

// shared memory of the generic row-owner backward for `cc` channels and `nstages` ring stages
size_t rowsx_smem(int H, int W, int cc, bool maxpool, int nstages) {
  const size_t map = (((size_t)H * (W + 4) * (cc + 1) + 3) & ~(size_t)3) * 4;
  const size_t tile = (size_t)cc * kPP;
  const size_t per_roi = tile * 4 + sizeof(SepRecX) + (maxpool ? ((tile + 15) & ~(size_t)15) : 0);
  return map + (size_t)nstages * kRowXGroup * per_roi + 128;
}

// ------------------------------------------------------------------ host side
struct Plan {
  int cc;
  size_t smem_fwd, smem_bwd, smem_rows;
};

bool make_plan(int HW, bool maxpool, int rec, Plan* pl) {
  const size_t cap = (size_t)max_smem_optin() - 1024;
  pl->smem_rows = 0;
  for (int cc = 32; cc >= 4; cc >>= 1) {
    const size_t map = (size_t)(HW + 1) * cc * 4;
    const size_t tile = (size_t)cc * kPP;
    const int rpi = kRPI;
    const size_t fwd = map + 2 * rpi * tile * 4 + (size_t)(maxpool ? 2 : 4) * rpi * rec + (maxpool ? 2 * rpi * tile : 0) + 128;
    const size_t bwd = map + kBwdStages * (tile * 4 + (size_t)rec + (maxpool ? tile : 0)) + 128;
    if (fwd <= cap && bwd <= cap && (size_t)(HW + 1) * (cc / 4) <= 65535) {
      pl->cc = cc;
      pl->smem_fwd = fwd;
      pl->smem_bwd = bwd;
      return true;
    }
  }
  return false;
}

size_t table_offset(int B, int N) {
  return (((size_t)2 * B + 1 + (size_t)(N > 0 ? N : 0)) * sizeof(int) + 255) & ~(size_t)255;
}

int check_common(const void* a, const void* rois, const void* o, int B, int C, int H, int W, int N, int pool,
                 int flags, float im_h, float im_w, void* ws, size_t ws_bytes) {
  L2S_REQUIRE(a && rois && o, L2S_ERR_ARG, "roi_crop: null pointer");
  L2S_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1 && N >= 0, L2S_ERR_SHAPE, "roi_crop: bad shape B=%d C=%d H=%d W=%d N=%d",
              B, C, H, W, N);
  L2S_REQUIRE(pool == 7, L2S_ERR_SHAPE, "roi_crop: pool must be 7 (cfg.POOLING_SIZE), got %d", pool);
  L2S_REQUIRE(C % 4 == 0, L2S_ERR_SHAPE, "roi_crop: C must be a multiple of 4, got %d", C);
  L2S_REQUIRE(aligned16(a) && aligned16(o), L2S_ERR_ALIGN, "roi_crop: map / pooled pointers must be 16-byte aligned");
  L2S_REQUIRE((flags & ~(L2S_CROP_MAX_POOL | L2S_CROP_ALIGN | L2S_CROP_BWD_RANKED | L2S_CROP_WS_PREPARED)) == 0, L2S_ERR_ARG,
              "roi_crop: unknown flags %d", flags);
  if (flags & L2S_CROP_ALIGN)
    L2S_REQUIRE(im_h > 1.f && im_w > 1.f, L2S_ERR_ARG, "roi_crop: align mode needs the image size");
  L2S_REQUIRE(ws && ws_bytes >= l2s_roi_crop_workspace_bytes(B, N, flags) && aligned16(ws), L2S_ERR_WORKSPACE,
              "roi_crop: workspace missing, misaligned or too small (%zu < %zu)", ws_bytes,
              l2s_roi_crop_workspace_bytes(B, N, flags));
  return L2S_OK;
}

CropGeom make_geom(int B, int C, int H, int W, int N, int pool, int flags, float im_h, float im_w) {
  CropGeom g;
  g.B = B; g.C = C; g.H = H; g.W = W; g.N = N;
  g.maxpool = (flags & L2S_CROP_MAX_POOL) ? 1 : 0;
  g.S = g.maxpool ? 2 * pool : pool;
  g.rec = g.S * g.S * 16 + kRecTail;
  if (flags & L2S_CROP_ALIGN) {
    g.sx = (float)(W - 1) / (im_w - 1.f);
    g.sy = (float)(H - 1) / (im_h - 1.f);
  } else {
    g.sx = 1.0f / 16.0f;   // self._feat_stride = [16] : network_cycle_response.py:44,122-125
    g.sy = 1.0f / 16.0f;
  }
  return g;
}

// ROIs whose batch index is outside [0,B) belong to no map: they are skipped by every kernel and their output rows
// are zero (the reference never has them: proposal_layer.py:65 writes batch index 0)
__global__ void roi_invalid_zero_kernel(const float* __restrict__ rois, float* __restrict__ out, int B, int row_floats) {
  const int n = blockIdx.x;
  const float bf = __ldg(rois + 5 * (size_t)n);
  const int b = (int)bf;
  if (bf == bf && b >= 0 && b < B && (float)b <= bf) return;       // a valid index (NaN, negative, >= B fall through)
  float4* o = reinterpret_cast<float4*>(out + (size_t)n * row_floats);
  for (int i = threadIdx.x; i < row_floats / 4; i += blockDim.x) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

struct Prep {
  int* seg;
  int* order;
  unsigned char* table;
  SepRec* sep;
  SepRecX* sepx;
};

// where the binning and the records live inside the caller's workspace
void prep_pointers(const CropGeom& g, void* ws, Prep* p) {
  int* counts = reinterpret_cast<int*>(ws);
  p->seg = counts + g.B;
  p->order = counts + 2 * g.B + 1;
  p->table = reinterpret_cast<unsigned char*>(ws) + table_offset(g.B, g.N);
  p->sep = reinterpret_cast<SepRec*>(p->table + (size_t)g.N * g.rec);
  p->sepx = reinterpret_cast<SepRecX*>(p->table + (size_t)g.N * g.rec + (g.S == 7 ? (size_t)g.N * sizeof(SepRec) : 0));
}

// bins the ROIs by batch index (2 small launches); `reuse`: the forward call already did (L2S_CROP_WS_PREPARED)
int prepare_order(const float* rois, const CropGeom& g, void* ws, cudaStream_t st, Prep* p, bool reuse = false) {
  prep_pointers(g, ws, p);
  if (reuse) return L2S_OK;
  int* counts = reinterpret_cast<int*>(ws);
  roi_count_kernel<<<g.B, 256, 0, st>>>(rois, g.N, counts);
  L2S_LAUNCH_OK("roi_count_kernel");
  roi_order_kernel<<<g.B, 256, 0, st>>>(rois, g.N, counts, p->seg, p->order);
  L2S_LAUNCH_OK("roi_order_kernel");
  count_launch(2);
  return L2S_OK;
}

// ... and writes the per-sample geometry records (forward, ranked backward, 7x7 row-owner backward)
int prepare(const float* rois, const CropGeom& g, int cgn, void* ws, cudaStream_t st, Prep* p, bool reuse = false) {
  int rc = prepare_order(rois, g, ws, st, p, reuse);
  if (rc || reuse) return rc;
  roi_geom_kernel<<<g.N, g.S == 7 ? 64 : 256, 0, st>>>(rois, p->order, p->seg, p->table, g.S == 7 ? p->sep : nullptr, g, cgn);
  L2S_LAUNCH_OK("roi_geom_kernel");
  count_launch();
  return L2S_OK;
}

// whether l2s_roi_crop_fwd takes the opt-in row-mirror kernel for this shape (it then writes the ROI binning and its own
// separable records, but not the geometry table / 7x7 records the default forward leaves in the workspace)
bool fwd_uses_row_mirror(int H, int W, bool maxpool) {
  static const bool rows7 = env_flag("L2S_CROP_FWD_ROWS"), table_fwd = env_flag("L2S_CROP_FWD_TABLE");
  if (!rows7 || table_fwd) return false;
  const size_t cap = (size_t)max_smem_optin() - 1024;
  for (int c : {32, 16})
    if (fwd_rows_smem(H, W, c, maxpool) <= cap) return true;
  return false;
}

template <int CC, bool MP>
int launch_fwd(const float* bottom, const int* seg, const unsigned char* table, float* out, uint8_t* argmax,
               const CropGeom& g, size_t smem, cudaStream_t st) {
  auto kern = roi_crop_fwd_kernel<CC, MP>;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.C + CC - 1) / CC, g.B);
  kern<<<grid, kFwdThreads, smem, st>>>(bottom, seg, table, out, argmax, g);
  L2S_LAUNCH_OK("roi_crop_fwd_kernel");
  count_launch();
  return L2S_OK;
}

template <int CC, bool MP>
int launch_bwd(const float* dout, const int* seg, const unsigned char* table, const uint8_t* argmax, float* dbottom,
               const CropGeom& g, size_t smem, cudaStream_t st) {
  auto kern = roi_crop_bwd_kernel<CC, MP>;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.C + CC - 1) / CC, g.B);
  kern<<<grid, ((MP ? 1 : 2) * CC / 4 + 1) * 32, smem, st>>>(dout, seg, table, argmax, dbottom, g);
  L2S_LAUNCH_OK("roi_crop_bwd_kernel");
  count_launch();
  return L2S_OK;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" size_t l2s_roi_crop_workspace_bytes(int B, int N, int flags) {
  const int S = (flags & L2S_CROP_MAX_POOL) ? 14 : 7;
  return table_offset(B, N) +
         (size_t)(N > 0 ? N : 0) * (S * S * 16 + kRecTail + (S == 7 ? sizeof(SepRec) : 0) + sizeof(SepRecX)) + 256;
}

#define L2S_CROP_DISPATCH(FN, ...)                                                      \
  do {                                                                                  \
    if (g.maxpool) {                                                                    \
      switch (pl.cc) {                                                                  \
        case 32: return FN<32, true>(__VA_ARGS__);                                      \
        case 16: return FN<16, true>(__VA_ARGS__);                                      \
        case 8: return FN<8, true>(__VA_ARGS__);                                        \
        default: return FN<4, true>(__VA_ARGS__);                                       \
      }                                                                                 \
    }                                                                                   \
    switch (pl.cc) {                                                                    \
      case 32: return FN<32, false>(__VA_ARGS__);                                       \
      case 16: return FN<16, false>(__VA_ARGS__);                                       \
      case 8: return FN<8, false>(__VA_ARGS__);                                         \
      default: return FN<4, false>(__VA_ARGS__);                                        \
    }                                                                                   \
  } while (0)

extern "C" int l2s_roi_crop_fwd(const float* bottom, const float* rois, float* out, uint8_t* argmax, int B,
                                int C, int H, int W, int N, int pool, int flags, float im_h, float im_w,
                                void* workspace, size_t workspace_bytes, l2s_stream_t stream) {
  if (N == 0) return L2S_OK;
  int rc = check_common(bottom, rois, out, B, C, H, W, N, pool, flags, im_h, im_w, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const CropGeom g = make_geom(B, C, H, W, N, pool, flags, im_h, im_w);
  Plan pl;
  L2S_REQUIRE(make_plan(H * W, g.maxpool, g.rec, &pl), L2S_ERR_SHAPE,
              "roi_crop: feature map %dx%d does not fit the shared-memory staging", H, W);
  Prep pr;
  roi_invalid_zero_kernel<<<N, 128, 0, st>>>(rois, out, B, C * kPP);
  L2S_LAUNCH_OK("roi_invalid_zero_kernel");
  count_launch();
  // row-mirror kernel (separable records), opt-in: parity green for every variant, but MEASURED SLOWER than the
  // table-driven kernels on B200 (cfg-3 14x14+max 2.52 vs 1.37 ms, cfg-2 7x7 2.13 vs 0.79 ms): with one output tile per
  // warp only 8 warps fit next to the map and their dependent LDS -> FMA chains are latency bound
  {
    const size_t cap = (size_t)max_smem_optin() - 1024;
    static const bool rows7 = env_flag("L2S_CROP_FWD_ROWS");        // A/B switch: row-mirror kernel for all variants
    static const bool table_fwd = env_flag("L2S_CROP_FWD_TABLE");   // diagnostics: force the table-driven kernels
    int cc = 0;
    for (int c : {32, 16})
      if (fwd_rows_smem(H, W, c, g.maxpool) <= cap) { cc = c; break; }
    if (cc && !table_fwd && rows7) {   // == fwd_uses_row_mirror(H, W, g.maxpool)
      rc = prepare_order(rois, g, workspace, st, &pr);
      if (rc) return rc;
      roi_sepx_kernel<<<g.N, 64, 0, st>>>(rois, pr.order, pr.seg, pr.sepx, g, cc + 1, kRowWarps * (32 / cc));
      L2S_LAUNCH_OK("roi_sepx_kernel");
      count_launch();
      const size_t smem = fwd_rows_smem(H, W, cc, g.maxpool);
      dim3 grid((g.C + cc - 1) / cc, g.B);
#define L2S_FROWS(CCV, SV, MPV)                                                                               \
  do {                                                                                                        \
    auto kern = roi_crop_fwd_rows_kernel<CCV, SV, MPV>;                                                       \
    L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    kern<<<grid, kFwdRowWarps * 32, smem, st>>>(bottom, pr.seg, pr.sepx, out, argmax, g);                     \
  } while (0)
      if (g.maxpool) { if (cc == 32) L2S_FROWS(32, 14, true); else L2S_FROWS(16, 14, true); }
      else           { if (cc == 32) L2S_FROWS(32, 7, false); else L2S_FROWS(16, 7, false); }
#undef L2S_FROWS
      L2S_LAUNCH_OK("roi_crop_fwd_rows_kernel");
      count_launch();
      return L2S_OK;
    }
  }
  rc = prepare(rois, g, pl.cc / 4, workspace, st, &pr);
  if (rc) return rc;
  const int* seg = pr.seg;
  const unsigned char* table = pr.table;
  // warp-per-ROI kernel: 7x7 crops of maps whose 32-channel slice + one tile per warp fit in shared memory
  const size_t smem_warp = (size_t)(H * W + 1) * 32 * 4 + (size_t)kFwdWarps * 32 * kPP * 4 + 128;
  static const bool force_block = env_flag("L2S_CROP_FWD_BLOCK");     // diagnostics: force the block-synchronous kernel
  if (!g.maxpool && pl.cc == 32 && smem_warp <= (size_t)max_smem_optin() - 1024 && !force_block) {
    auto kern = roi_crop_fwd_warp_kernel;
    L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_warp));
    dim3 grid((g.C + 31) / 32, g.B);
    kern<<<grid, kFwdWarps * 32, smem_warp, st>>>(bottom, seg, table, out, g);
    L2S_LAUNCH_OK("roi_crop_fwd_warp_kernel");
    count_launch();
    return L2S_OK;
  }
  L2S_CROP_DISPATCH(launch_fwd, bottom, seg, table, out, argmax, g, pl.smem_fwd, st);
}

extern "C" int l2s_roi_crop_bwd(const float* dout, const float* rois, const uint8_t* argmax, float* dbottom,
                                int B, int C, int H, int W, int N, int pool, int flags, float im_h, float im_w,
                                void* workspace, size_t workspace_bytes, l2s_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    L2S_REQUIRE(dbottom && B > 0 && C > 0 && H > 0 && W > 0, L2S_ERR_ARG, "roi_crop_bwd: bad arguments");
    L2S_CUDA_OK(cudaMemsetAsync(dbottom, 0, (size_t)B * C * H * W * sizeof(float), st));
    return L2S_OK;
  }
  int rc = check_common(dbottom, rois, dout ? (const void*)dout : (const void*)dbottom, B, C, H, W, N, pool, flags,
                        im_h, im_w, workspace, workspace_bytes);
  if (rc) return rc;
  L2S_REQUIRE(dout, L2S_ERR_ARG, "roi_crop_bwd: null dout");
  const CropGeom g = make_geom(B, C, H, W, N, pool, flags, im_h, im_w);
  L2S_REQUIRE(!g.maxpool || (argmax && aligned16(argmax)), L2S_ERR_ARG,
              "roi_crop_bwd: max-pool mode needs the (16-byte aligned) argmax of the forward");
  Plan pl;
  L2S_REQUIRE(make_plan(H * W, g.maxpool, g.rec, &pl), L2S_ERR_SHAPE,
              "roi_crop: feature map %dx%d does not fit the shared-memory staging", H, W);
  const size_t cap = (size_t)max_smem_optin() - 1024;
  const bool ranked = (flags & L2S_CROP_BWD_RANKED) != 0;
  // L2S_CROP_WS_PREPARED: the workspace is the one the forward call with the same arguments filled -- the ROI binning
  // (always) and the geometry table / 7x7 separable records (default forward kernels) are reused instead of recomputed
  const bool ws_order = (flags & L2S_CROP_WS_PREPARED) != 0;
  const bool ws_table = ws_order && !fwd_uses_row_mirror(H, W, g.maxpool != 0);
  Prep pr;
  // 7x7 row-owner kernel: needs the whole 32-channel accumulator with guard columns + the tile ring in shared memory
  pl.smem_rows = (((size_t)H * (W + 4) * kRowLd + 3) & ~(size_t)3) * 4 +
                 kRowStages * kRowGroup * ((size_t)32 * kPP * 4 + sizeof(SepRec)) + 128;
  static const bool force_generic = env_flag("L2S_CROP_BWD_ROWSX");   // diagnostics: generic row-owner kernel for 7x7 too
  if (!g.maxpool && pl.cc == 32 && pl.smem_rows <= cap && !ranked && !force_generic) {
    rc = ws_table ? prepare(rois, g, pl.cc / 4, workspace, st, &pr, true) : prepare(rois, g, pl.cc / 4, workspace, st, &pr);
    if (rc) return rc;
    auto kern = roi_crop_bwd_rows_kernel;
    L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_rows));
    dim3 grid((g.C + 31) / 32, g.B);
    kern<<<grid, (kRowWarps + 1) * 32, pl.smem_rows, st>>>(dout, pr.seg, pr.sep, dbottom, g);
    L2S_LAUNCH_OK("roi_crop_bwd_rows_kernel");
    count_launch();
    return L2S_OK;
  }
  // generic row-owner kernel: 14x14 + 2x2 max, and 7x7 on maps that only fit a 16-channel accumulator
  if (!ranked) {
    int cc = 0, nst = 0;
    for (int c : {32, 16})
      if (rowsx_smem(H, W, c, g.maxpool, 2) <= cap) { cc = c; break; }
    if (cc) {
      for (nst = 2; nst < kRowXMaxStages && rowsx_smem(H, W, cc, g.maxpool, nst + 1) <= cap; ++nst) {}
      rc = prepare_order(rois, g, workspace, st, &pr, ws_order);
      if (rc) return rc;
      roi_sepx_kernel<<<g.N, 64, 0, st>>>(rois, pr.order, pr.seg, pr.sepx, g, cc + 1, kRowWarps * (32 / cc));
      L2S_LAUNCH_OK("roi_sepx_kernel");
      count_launch();
      const size_t smem = rowsx_smem(H, W, cc, g.maxpool, nst);
      dim3 grid((g.C + cc - 1) / cc, g.B);
#define L2S_ROWSX(CCV, SV, MPV)                                                                               \
  do {                                                                                                        \
    auto kern = roi_crop_bwd_rowsx_kernel<CCV, SV, MPV>;                                                      \
    L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    kern<<<grid, (kRowWarps + 1) * 32, smem, st>>>(dout, pr.seg, pr.sepx, argmax, dbottom, g, nst);           \
  } while (0)
      if (g.maxpool) { if (cc == 32) L2S_ROWSX(32, 14, true); else L2S_ROWSX(16, 14, true); }
      else           { if (cc == 32) L2S_ROWSX(32, 7, false); else L2S_ROWSX(16, 7, false); }
#undef L2S_ROWSX
      L2S_LAUNCH_OK("roi_crop_bwd_rowsx_kernel");
      count_launch();
      return L2S_OK;
    }
  }
  rc = prepare(rois, g, pl.cc / 4, workspace, st, &pr, ws_table);
  if (rc) return rc;
  const int* seg = pr.seg;
  const unsigned char* table = pr.table;
  L2S_CROP_DISPATCH(launch_bwd, dout, seg, table, argmax, dbottom, g, pl.smem_bwd, st);
}
