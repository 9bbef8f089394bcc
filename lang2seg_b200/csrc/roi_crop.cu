// Crop-and-resize ROI pooling, forward and backward, for sm_100a.
//
// Semantics: Network._crop_pool_layer / _crop_pool_layer_align of the lang2seg reference
// (pyutils/mask-faster-rcnn/lib/nets/network_cycle_response.py:107-182): S x S bilinear samples
// evenly spaced from box start to box end inclusive (align_corners=True), zero padding outside
// the map, optional 2x2 max over a 2S x 2S sample grid.  Closed form: SURVEY.md appendix A.2.
//
// Design (HBM-bound; the output / upstream gradient is ~12x the size of the map):
//   * one CTA owns (batch index b, chunk of CC channels) and stages that slice of the map ONCE
//     in shared memory, transposed to [pixel][channel] with a 16-byte XOR swizzle, then loops
//     over all ROIs of b.  Map bytes therefore cross L2->SM exactly once per CTA.
//   * forward: thread = (4-channel group, sample position); four LDS.128 gathers per output
//     float4, results staged in a [channel][49] tile that is written to HBM with one 1-D TMA
//     bulk store per ROI (cp.async.bulk.global.shared::cta), double buffered.
//   * backward: upstream-gradient tiles arrive by TMA bulk loads (mbarrier pipeline, producer
//     warp).  Every consumer warp exclusively OWNS one (4-channel group, pixel-parity class) of
//     the shared-memory accumulator map, so plain vectorised read-modify-write is race free
//     across warps; lanes are sample positions and collisions inside a warp are resolved with
//     match.any + ranked rounds (warp-aggregated, fixed order => bit-reproducible sums).
//     No floating-point atomics anywhere (shared fp32 atomicAdd is a CAS loop on sm_100).
#include "common.cuh"

namespace l2s {
namespace {

constexpr int kSlotP = 56;          // sample slots per ROI in the forward kernel (49 used)
constexpr int kFwdThreads = 896;    // RPI * CGN * kSlotP with RPI * CGN == 16
constexpr int kBwdStages = 4;
constexpr int kBoxChunk = 256;      // ROI boxes staged per pass in the forward kernel

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

// ------------------------------------------------------------------ staging the map slice
// smem map layout: element (px, c) at px*CC + ((c/4 ^ key(px)) & (CGN-1))*4 + (c&3)
template <int CC>
__device__ __forceinline__ int map_off(int px, int cgrp) {
  constexpr int CGN = CC / 4;
  return px * CC + (((cgrp ^ swz_key(px)) & (CGN - 1)) << 2);
}

template <int CC>
__device__ void stage_map(float* __restrict__ map, const float* __restrict__ src /* (C,HW) of batch b */,
                          int c0, int C, int HW) {
  constexpr int CGN = CC / 4;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int pl = lane & 7, cl = lane >> 3;
  const int nblk = (HW + 7) >> 3;
  for (int it = wid; it < nblk * CGN; it += nw) {
    const int g = it % CGN, blk = it / CGN;
    const int px = blk * 8 + pl, c = c0 + g * 4 + cl;
    if (px < HW) {
      const float v = (c < C) ? __ldg(src + (size_t)c * HW + px) : 0.f;
      map[map_off<CC>(px, g) + cl] = v;
    }
  }
}

template <int CC>
__device__ void unstage_map(const float* __restrict__ map, float* __restrict__ dst, int c0, int C, int HW) {
  constexpr int CGN = CC / 4;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int pl = lane & 7, cl = lane >> 3;
  const int nblk = (HW + 7) >> 3;
  for (int it = wid; it < nblk * CGN; it += nw) {
    const int g = it % CGN, blk = it / CGN;
    const int px = blk * 8 + pl, c = c0 + g * 4 + cl;
    if (px < HW && c < C) dst[(size_t)c * HW + px] = map[map_off<CC>(px, g) + cl];
  }
}

// ------------------------------------------------------------------ forward
template <int CC>
__global__ void __launch_bounds__(kFwdThreads, 1)
roi_crop_fwd_kernel(const float* __restrict__ bottom, const float* __restrict__ rois,
                    const int* __restrict__ seg, const int* __restrict__ order, float* __restrict__ out,
                    uint8_t* __restrict__ argmax, CropGeom g) {
  constexpr int CGN = CC / 4;
  constexpr int RPI = 16 / CGN;            // ROIs per iteration
  constexpr int PP = 49;
  constexpr int TILE = CC * PP;            // floats per ROI tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W;
  float* map = reinterpret_cast<float*>(smem_raw);
  float* tiles = map + (size_t)HW * CC;                        // [2][RPI][TILE]
  uint8_t* atile = reinterpret_cast<uint8_t*>(tiles + 2 * RPI * TILE);   // [2][RPI][TILE] (max-pool only)
  __shared__ int s_n[2][RPI];
  __shared__ float4 s_box[kBoxChunk];
  __shared__ int s_ni[kBoxChunk];

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x;
  const int slot = t / (CGN * kSlotP), r = t % (CGN * kSlotP);
  const int cg = r % CGN, p = r / CGN;
  const int pi = p / 7, pj = p % 7;
  const int cvalid = min(CC, g.C - c0);

  stage_map<CC>(map, bottom + (size_t)b * g.C * HW, c0, g.C, HW);
  __syncthreads();

  const int beg = seg[b], end = seg[b + 1];
  const float inv = 1.0f / (float)(g.S - 1);
  const float4* map4 = reinterpret_cast<const float4*>(map);
  const int W = g.W, H = g.H;

  auto gather = [&](const Corner& c) -> float4 {
    const bool vx0 = (unsigned)c.x0 < (unsigned)W, vx1 = (unsigned)(c.x0 + 1) < (unsigned)W;
    const bool vy0 = (unsigned)c.y0 < (unsigned)H, vy1 = (unsigned)(c.y0 + 1) < (unsigned)H;
    const int xa = min(max(c.x0, 0), W - 1), xb = min(max(c.x0 + 1, 0), W - 1);
    const int ya = min(max(c.y0, 0), H - 1), yb = min(max(c.y0 + 1, 0), H - 1);
    const float wy0 = vy0 ? 1.f - c.ly : 0.f, wy1 = vy1 ? c.ly : 0.f;
    const float wx0 = vx0 ? 1.f - c.lx : 0.f, wx1 = vx1 ? c.lx : 0.f;
    const int p00 = ya * W + xa, p01 = ya * W + xb, p10 = yb * W + xa, p11 = yb * W + xb;
    const float4 a = map4[map_off<CC>(p00, cg) >> 2];
    const float4 bq = map4[map_off<CC>(p01, cg) >> 2];
    const float4 cq = map4[map_off<CC>(p10, cg) >> 2];
    const float4 d = map4[map_off<CC>(p11, cg) >> 2];
    const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
    float4 o;
    o.x = fmaf(w11, d.x, fmaf(w10, cq.x, fmaf(w01, bq.x, w00 * a.x)));
    o.y = fmaf(w11, d.y, fmaf(w10, cq.y, fmaf(w01, bq.y, w00 * a.y)));
    o.z = fmaf(w11, d.z, fmaf(w10, cq.z, fmaf(w01, bq.z, w00 * a.z)));
    o.w = fmaf(w11, d.w, fmaf(w10, cq.w, fmaf(w01, bq.w, w00 * a.w)));
    return o;
  };

  int it = 0;
  for (int cb0 = beg; cb0 < end; cb0 += kBoxChunk) {
    // stage the boxes (already scaled to feature pixels) and ROI ids of this chunk of the segment
    __syncthreads();
    for (int i = t; i < min(kBoxChunk, end - cb0); i += blockDim.x) {
      const int n = __ldg(order + cb0 + i);
      const float* rp = rois + 5 * (size_t)n;
      s_ni[i] = n;
      s_box[i] = make_float4(__ldg(rp + 1) * g.sx, __ldg(rp + 2) * g.sy, __ldg(rp + 3) * g.sx, __ldg(rp + 4) * g.sy);
    }
    __syncthreads();
    const int cend = min(end, cb0 + kBoxChunk);
  for (int gi = cb0; gi < cend; gi += RPI, ++it) {
    const int buf = it & 1;
    const int ri = gi + slot;
    float* tile = tiles + (size_t)(buf * RPI + slot) * TILE;
    uint8_t* at = atile + (size_t)(buf * RPI + slot) * TILE;
    if (ri < cend && p < PP) {
      if (r == 0) s_n[buf][slot] = s_ni[ri - cb0];
      const float4 box = s_box[ri - cb0];
      float4 o;
      if (!g.maxpool) {
        o = gather(sample_at(box, pi, pj, inv));
      } else {
        // 2x2 block of the 2S x 2S sample grid; first strict maximum in row-major order wins,
        // as in max_pool2d's backward (network_cycle_response.py:144)
        o = gather(sample_at(box, 2 * pi, 2 * pj, inv));
        uint32_t am = 0;   // 4 x 8-bit winners
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const float4 v = gather(sample_at(box, 2 * pi + (q >> 1), 2 * pj + (q & 1), inv));
          if (v.x > o.x) { o.x = v.x; am = (am & ~0xffu) | (uint32_t)q; }
          if (v.y > o.y) { o.y = v.y; am = (am & ~0xff00u) | ((uint32_t)q << 8); }
          if (v.z > o.z) { o.z = v.z; am = (am & ~0xff0000u) | ((uint32_t)q << 16); }
          if (v.w > o.w) { o.w = v.w; am = (am & ~0xff000000u) | ((uint32_t)q << 24); }
        }
        const int cb = cg * 4;
        at[(cb + 0) * PP + p] = (uint8_t)(am & 0xff);
        at[(cb + 1) * PP + p] = (uint8_t)((am >> 8) & 0xff);
        at[(cb + 2) * PP + p] = (uint8_t)((am >> 16) & 0xff);
        at[(cb + 3) * PP + p] = (uint8_t)(am >> 24);
      }
      const int cb = cg * 4;
      tile[(cb + 0) * PP + p] = o.x;
      tile[(cb + 1) * PP + p] = o.y;
      tile[(cb + 2) * PP + p] = o.z;
      tile[(cb + 3) * PP + p] = o.w;
    }
    fence_proxy_async_smem();
    if (t == 0) bulk_wait_read<0>();      // the previous iteration's stores have left the other buffer
    __syncthreads();
    if (t == 0) {
      const int cnt = min(RPI, cend - gi);
      for (int s = 0; s < cnt; ++s) {
        const int n = s_n[buf][s];
        bulk_s2g(out + ((size_t)n * g.C + c0) * PP, tiles + (size_t)(buf * RPI + s) * TILE,
                 (uint32_t)(cvalid * PP * sizeof(float)));
      }
      bulk_commit();
    }
    if (g.maxpool && argmax != nullptr) {
      // winners: plain 32-bit copies (tile byte count and global offset are multiples of 4)
      const int cnt = min(RPI, cend - gi);
      const int words = cvalid * PP / 4;
      for (int w = t; w < cnt * words; w += blockDim.x) {
        const int s = w / words, k = w % words;
        const int n = s_n[buf][s];
        reinterpret_cast<uint32_t*>(argmax + ((size_t)n * g.C + c0) * PP)[k] =
            reinterpret_cast<const uint32_t*>(atile + (size_t)(buf * RPI + s) * TILE)[k];
      }
    }
  }
  }
  if (t == 0) bulk_wait<0>();
}

// ------------------------------------------------------------------ backward
// warps: CGN * NCLS consumers + 1 producer.  Consumer warp (cg, q) owns the accumulator entries
// of channel group cg at pixels of parity class q.
template <int CC, int NCLS>
__global__ void __launch_bounds__((CC / 4 * NCLS + 1) * 32, 1)
roi_crop_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ rois,
                    const int* __restrict__ seg, const int* __restrict__ order,
                    const uint8_t* __restrict__ argmax, float* __restrict__ dbottom, CropGeom g) {
  constexpr int CGN = CC / 4;
  constexpr int NCONS = CGN * NCLS;
  constexpr int PP = 49;
  constexpr int TILE = CC * PP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W;
  float* map = reinterpret_cast<float*>(smem_raw);
  float* tiles = map + (size_t)HW * CC;                          // [kBwdStages][TILE]
  __shared__ uint64_t full_bar[kBwdStages], empty_bar[kBwdStages];
  __shared__ int s_n[kBwdStages];

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cvalid = min(CC, g.C - c0);

  for (int i = t; i < HW * CC; i += blockDim.x) map[i] = 0.f;
  if (t == 0) {
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NCONS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int beg = seg[b], end = seg[b + 1];
  const uint32_t tile_bytes = (uint32_t)(cvalid * PP * sizeof(float));

  if (wid == NCONS) {
    // ---------------- producer: one lane streams the upstream-gradient tiles ----------------
    if (lane == 0) {
      for (int ri = beg, k = 0; ri < end; ++ri, ++k) {
        const int s = k % kBwdStages;
        if (k >= kBwdStages) mbar_wait(&empty_bar[s], ((k / kBwdStages) - 1) & 1);
        const int n = __ldg(order + ri);
        s_n[s] = n;
        mbar_arrive_expect_tx(&full_bar[s], tile_bytes);
        bulk_g2s(tiles + (size_t)s * TILE, dout + ((size_t)n * g.C + c0) * PP, tile_bytes, &full_bar[s]);
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int cg = wid % CGN, q = wid / CGN;
    const int qy = (NCLS >= 2) ? (q & 1) : 0;
    const int qx = (NCLS == 4) ? (q >> 1) : 0;
    const float inv = 1.0f / (float)(g.S - 1);
    const int W = g.W, H = g.H;
    const bool ch_ok = cg * 4 < cvalid;
    float4* map4 = reinterpret_cast<float4*>(map);
    const unsigned lt = (1u << lane) - 1u;

    for (int ri = beg, k = 0; ri < end; ++ri, ++k) {
      const int s = k % kBwdStages;
      mbar_wait(&full_bar[s], (k / kBwdStages) & 1);
      const int n = s_n[s];
      const float* tile = tiles + (size_t)s * TILE;
      const float* rp = rois + 5 * (size_t)n;
      float4 box;
      box.x = __ldg(rp + 1) * g.sx;
      box.y = __ldg(rp + 2) * g.sy;
      box.z = __ldg(rp + 3) * g.sx;
      box.w = __ldg(rp + 4) * g.sy;
#pragma unroll 1
      for (int round = 0; round < 2; ++round) {
        const int p = round * 32 + lane;
        const bool act = p < PP && ch_ok;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) {
          const int cb = cg * 4;
          v.x = tile[(cb + 0) * PP + p];
          v.y = tile[(cb + 1) * PP + p];
          v.z = tile[(cb + 2) * PP + p];
          v.w = tile[(cb + 3) * PP + p];
        }
        const int pi = p / 7, pj = p % 7;
        uint32_t am = 0;
        if (g.maxpool && act) {
          const uint8_t* ap = argmax + ((size_t)n * g.C + c0 + cg * 4) * PP + p;
          am = (uint32_t)ap[0] | ((uint32_t)ap[PP] << 8) | ((uint32_t)ap[2 * PP] << 16) |
               ((uint32_t)ap[3 * PP] << 24);
        }
        // with max-pool the 4 channels of a lane may route to different samples: handle them as
        // 4 scalar streams; otherwise one float4 stream.
        const int nstream = g.maxpool ? 4 : 1;
#pragma unroll 1
        for (int e = 0; e < nstream; ++e) {
          const int a = (am >> (8 * e)) & 3;
          const Corner c = g.maxpool ? sample_at(box, 2 * pi + (a >> 1), 2 * pj + (a & 1), inv)
                                     : sample_at(box, pi, pj, inv);
#pragma unroll
          for (int sub = 0; sub < 4 / NCLS; ++sub) {
            // the corner(s) of this sample that fall into this warp's parity class
            int dy, dx;
            if (NCLS == 4) { dy = (qy ^ c.y0) & 1; dx = (qx ^ c.x0) & 1; }
            else if (NCLS == 2) { dy = (qy ^ c.y0) & 1; dx = sub; }
            else { dy = sub >> 1; dx = sub & 1; }
            const int yy = c.y0 + dy, xx = c.x0 + dx;
            const bool ok = act && (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
            const float wgt = (dy ? c.ly : 1.f - c.ly) * (dx ? c.lx : 1.f - c.lx);
            const int px = yy * W + xx;
            const int key = ok ? px : -1 - lane;
            const unsigned grp = __match_any_sync(0xffffffffu, key);
            const int rank = __popc(grp & lt);
            const int maxrank = __reduce_max_sync(0xffffffffu, ok ? rank : 0);
            for (int rr = 0; rr <= maxrank; ++rr) {
              if (ok && rank == rr) {
                const int off = map_off<CC>(px, cg);
                if (g.maxpool) {
                  const float val = e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w));
                  map[off + e] = fmaf(wgt, val, map[off + e]);
                } else {
                  float4 m = map4[off >> 2];
                  m.x = fmaf(wgt, v.x, m.x);
                  m.y = fmaf(wgt, v.y, m.y);
                  m.z = fmaf(wgt, v.z, m.z);
                  m.w = fmaf(wgt, v.w, m.w);
                  map4[off >> 2] = m;
                }
              }
              __syncwarp();
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
  __syncthreads();
  unstage_map<CC>(map, dbottom + (size_t)b * g.C * HW, c0, g.C, HW);
}


// ------------------------------------------------------------------ backward, fast path (no max-pool)
// Geometry is channel independent, so the PRODUCER warp computes it once per ROI: for each of the 49
// samples the corner index, the two fractions and a collision rank (samples whose (y0,x0) coincide get
// ranks 0,1,2,.. by match.any), next to issuing the TMA load of the gradient tile.  Consumer warp cg owns
// channel quad cg of the accumulator map: lanes = samples, 4 corners of a sample are 4 distinct pixels,
// equal-rank samples never share a pixel through the same corner => plain float4 read-modify-write.
struct __align__(16) SampleEnt {
  int yx;        // (y0 << 16) | (x0 & 0xffff)
  float ly, lx;
  int rank;
};

template <int CC>
__global__ void __launch_bounds__((CC / 4 + 1) * 32, 1)
roi_crop_bwd_fast_kernel(const float* __restrict__ dout, const float* __restrict__ rois,
                         const int* __restrict__ seg, const int* __restrict__ order, float* __restrict__ dbottom,
                         CropGeom g) {
  constexpr int CGN = CC / 4;
  constexpr int PP = 49;
  constexpr int TILE = CC * PP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = g.H * g.W;
  float* map = reinterpret_cast<float*>(smem_raw);
  float* tiles = map + (size_t)HW * CC;                          // [kBwdStages][TILE]
  __shared__ uint64_t full_bar[kBwdStages], empty_bar[kBwdStages];
  __shared__ SampleEnt s_tab[kBwdStages][64];
  __shared__ int s_maxrank[kBwdStages][2];

  const int b = blockIdx.y, c0 = blockIdx.x * CC;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cvalid = min(CC, g.C - c0);
  const unsigned lt = (1u << lane) - 1u;

  for (int i = t; i < HW * CC; i += blockDim.x) map[i] = 0.f;
  if (t == 0) {
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&full_bar[s], 2);          // TMA issue (expect_tx) + "table written"
      mbar_init(&empty_bar[s], CGN);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int beg = seg[b], end = seg[b + 1];
  const uint32_t tile_bytes = (uint32_t)(cvalid * PP * sizeof(float));
  const float inv = 1.0f / (float)(g.S - 1);

  if (wid == CGN) {
    // ---------------- producer warp ----------------
    for (int ri = beg, k = 0; ri < end; ++ri, ++k) {
      const int s = k % kBwdStages;
      if (k >= kBwdStages) mbar_wait(&empty_bar[s], ((k / kBwdStages) - 1) & 1);
      const int n = __ldg(order + ri);
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[s], tile_bytes);
        bulk_g2s(tiles + (size_t)s * TILE, dout + ((size_t)n * g.C + c0) * PP, tile_bytes, &full_bar[s]);
      }
      const float* rp = rois + 5 * (size_t)n;
      float4 box;
      box.x = __ldg(rp + 1) * g.sx;
      box.y = __ldg(rp + 2) * g.sy;
      box.z = __ldg(rp + 3) * g.sx;
      box.w = __ldg(rp + 4) * g.sy;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int p = pass * 32 + lane;
        const bool valid = p < PP;
        const Corner c = sample_at(box, p / 7, p % 7, inv);
        const int y0 = min(max(c.y0, -2), 30000), x0 = min(max(c.x0, -2), 30000);
        const int key = valid ? ((y0 << 16) | (x0 & 0xffff)) : (int)(0x80000000u | (unsigned)lane);
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(grp & lt);
        const int mr = __reduce_max_sync(0xffffffffu, valid ? rank : 0);
        SampleEnt e;
        e.yx = key; e.ly = c.ly; e.lx = c.lx; e.rank = rank;
        s_tab[s][p] = e;
        if (lane == 0) s_maxrank[s][pass] = mr;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
  } else {
    // ---------------- consumer warps ----------------
    const int cg = wid;
    const int W = g.W, H = g.H;
    const bool ch_ok = cg * 4 < cvalid;
    float4* map4 = reinterpret_cast<float4*>(map);
    for (int ri = beg, k = 0; ri < end; ++ri, ++k) {
      const int s = k % kBwdStages;
      mbar_wait(&full_bar[s], (k / kBwdStages) & 1);
      const float* tile = tiles + (size_t)s * TILE;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int p = pass * 32 + lane;
        const bool act = p < PP && ch_ok;
        const SampleEnt e = s_tab[s][p];
        const int mr = s_maxrank[s][pass];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) {
          const int cb = cg * 4;
          v.x = tile[(cb + 0) * PP + p];
          v.y = tile[(cb + 1) * PP + p];
          v.z = tile[(cb + 2) * PP + p];
          v.w = tile[(cb + 3) * PP + p];
        }
        const int y0 = e.yx >> 16, x0 = (int)(short)(e.yx & 0xffff);
        const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
        const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
        const float wy0 = 1.f - e.ly, wy1 = e.ly, wx0 = 1.f - e.lx, wx1 = e.lx;
        const int pbase = y0 * W + x0;
        // Two samples of equal rank have different (y0,x0) but may still meet in one pixel through
        // DIFFERENT corners (A's (y0+1,x0+1) is B's (y0,x0)); so the four corners are four separate
        // read-modify-write phases with a warp barrier between them.  Inside one phase a pixel is
        // reached only by samples of the same (y0,x0), which carry distinct ranks.
        const bool b00 = vy0 && vx0, b01 = vy0 && vx1, b10 = vy1 && vx0, b11 = vy1 && vx1;
        auto rmw = [&](bool on, int px, float w) {
          if (on) {
            const int o = map_off<CC>(px, cg) >> 2;
            float4 m = map4[o];
            m.x = fmaf(w, v.x, m.x);
            m.y = fmaf(w, v.y, m.y);
            m.z = fmaf(w, v.z, m.z);
            m.w = fmaf(w, v.w, m.w);
            map4[o] = m;
          }
          __syncwarp();
        };
        for (int rr = 0; rr <= mr; ++rr) {
          const bool sel = act && e.rank == rr;
          rmw(sel && b00, pbase, wy0 * wx0);
          rmw(sel && b01, pbase + 1, wy0 * wx1);
          rmw(sel && b10, pbase + W, wy1 * wx0);
          rmw(sel && b11, pbase + W + 1, wy1 * wx1);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
  __syncthreads();
  unstage_map<CC>(map, dbottom + (size_t)b * g.C * HW, c0, g.C, HW);
}

// ------------------------------------------------------------------ host side
struct Plan {
  int cc;
  size_t smem_fwd, smem_bwd;
};

bool make_plan(int HW, bool maxpool, Plan* pl) {
  const size_t cap = (size_t)max_smem_optin() - 1024;
  for (int cc = 32; cc >= 4; cc >>= 1) {
    const size_t map = (size_t)HW * cc * 4;
    const size_t tile = (size_t)cc * 49;
    const int rpi = 16 / (cc / 4);
    const size_t fwd = map + 2 * rpi * tile * 4 + (maxpool ? 2 * rpi * tile : 0) + 128;
    const size_t bwd = map + kBwdStages * tile * 4 + 128;
    if (fwd <= cap && bwd <= cap) {
      pl->cc = cc;
      pl->smem_fwd = fwd;
      pl->smem_bwd = bwd;
      return true;
    }
  }
  return false;
}

int check_common(const void* a, const void* rois, const void* o, int B, int C, int H, int W, int N, int pool,
                 int flags, float im_h, float im_w, void* ws, size_t ws_bytes) {
  L2S_REQUIRE(a && rois && o, L2S_ERR_ARG, "roi_crop: null pointer");
  L2S_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1 && N >= 0, L2S_ERR_SHAPE, "roi_crop: bad shape B=%d C=%d H=%d W=%d N=%d",
              B, C, H, W, N);
  L2S_REQUIRE(pool == 7, L2S_ERR_SHAPE, "roi_crop: pool must be 7 (cfg.POOLING_SIZE), got %d", pool);
  L2S_REQUIRE(C % 4 == 0, L2S_ERR_SHAPE, "roi_crop: C must be a multiple of 4, got %d", C);
  L2S_REQUIRE(aligned16(a) && aligned16(o), L2S_ERR_ALIGN, "roi_crop: map / pooled pointers must be 16-byte aligned");
  L2S_REQUIRE((flags & ~(L2S_CROP_MAX_POOL | L2S_CROP_ALIGN)) == 0, L2S_ERR_ARG, "roi_crop: unknown flags %d", flags);
  if (flags & L2S_CROP_ALIGN)
    L2S_REQUIRE(im_h > 1.f && im_w > 1.f, L2S_ERR_ARG, "roi_crop: align mode needs the image size");
  L2S_REQUIRE(ws && ws_bytes >= l2s_roi_crop_workspace_bytes(B, N), L2S_ERR_WORKSPACE,
              "roi_crop: workspace too small (%zu < %zu)", ws_bytes, l2s_roi_crop_workspace_bytes(B, N));
  return L2S_OK;
}

CropGeom make_geom(int B, int C, int H, int W, int N, int pool, int flags, float im_h, float im_w) {
  CropGeom g;
  g.B = B; g.C = C; g.H = H; g.W = W; g.N = N;
  g.maxpool = (flags & L2S_CROP_MAX_POOL) ? 1 : 0;
  g.S = g.maxpool ? 2 * pool : pool;
  if (flags & L2S_CROP_ALIGN) {
    g.sx = (float)(W - 1) / (im_w - 1.f);
    g.sy = (float)(H - 1) / (im_h - 1.f);
  } else {
    g.sx = 1.0f / 16.0f;   // self._feat_stride = [16] : network_cycle_response.py:44,122-125
    g.sy = 1.0f / 16.0f;
  }
  return g;
}

int bin_rois(const float* rois, int B, int N, void* ws, cudaStream_t st, int** seg, int** order) {
  int* counts = reinterpret_cast<int*>(ws);
  *seg = counts + B;
  *order = counts + 2 * B + 1;
  roi_count_kernel<<<B, 256, 0, st>>>(rois, N, counts);
  L2S_LAUNCH_OK("roi_count_kernel");
  roi_order_kernel<<<B, 256, 0, st>>>(rois, N, counts, *seg, *order);
  L2S_LAUNCH_OK("roi_order_kernel");
  count_launch(2);
  return L2S_OK;
}

template <int CC>
int launch_fwd(const float* bottom, const float* rois, const int* seg, const int* order, float* out,
               uint8_t* argmax, const CropGeom& g, size_t smem, cudaStream_t st) {
  auto kern = roi_crop_fwd_kernel<CC>;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.C + CC - 1) / CC, g.B);
  kern<<<grid, kFwdThreads, smem, st>>>(bottom, rois, seg, order, out, argmax, g);
  L2S_LAUNCH_OK("roi_crop_fwd_kernel");
  count_launch();
  return L2S_OK;
}

template <int CC, int NCLS>
int launch_bwd(const float* dout, const float* rois, const int* seg, const int* order, const uint8_t* argmax,
               float* dbottom, const CropGeom& g, size_t smem, cudaStream_t st) {
  auto kern = roi_crop_bwd_kernel<CC, NCLS>;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.C + CC - 1) / CC, g.B);
  kern<<<grid, (CC / 4 * NCLS + 1) * 32, smem, st>>>(dout, rois, seg, order, argmax, dbottom, g);
  L2S_LAUNCH_OK("roi_crop_bwd_kernel");
  count_launch();
  return L2S_OK;
}

template <int CC>
int launch_bwd_fast(const float* dout, const float* rois, const int* seg, const int* order, float* dbottom,
                    const CropGeom& g, size_t smem, cudaStream_t st) {
  auto kern = roi_crop_bwd_fast_kernel<CC>;
  L2S_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.C + CC - 1) / CC, g.B);
  kern<<<grid, (CC / 4 + 1) * 32, smem, st>>>(dout, rois, seg, order, dbottom, g);
  L2S_LAUNCH_OK("roi_crop_bwd_fast_kernel");
  count_launch();
  return L2S_OK;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" size_t l2s_roi_crop_workspace_bytes(int B, int N) {
  return ((size_t)2 * B + 1 + (size_t)(N > 0 ? N : 0)) * sizeof(int) + 64;
}

extern "C" int l2s_roi_crop_fwd(const float* bottom, const float* rois, float* out, uint8_t* argmax, int B,
                                int C, int H, int W, int N, int pool, int flags, float im_h, float im_w,
                                void* workspace, size_t workspace_bytes, l2s_stream_t stream) {
  if (N == 0) return L2S_OK;
  int rc = check_common(bottom, rois, out, B, C, H, W, N, pool, flags, im_h, im_w, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const CropGeom g = make_geom(B, C, H, W, N, pool, flags, im_h, im_w);
  Plan pl;
  L2S_REQUIRE(make_plan(H * W, g.maxpool, &pl), L2S_ERR_SHAPE,
              "roi_crop: feature map %dx%d does not fit the shared-memory staging (max 12800 pixels)", H, W);
  int *seg, *order;
  rc = bin_rois(rois, B, N, workspace, st, &seg, &order);
  if (rc) return rc;
  switch (pl.cc) {
    case 32: return launch_fwd<32>(bottom, rois, seg, order, out, argmax, g, pl.smem_fwd, st);
    case 16: return launch_fwd<16>(bottom, rois, seg, order, out, argmax, g, pl.smem_fwd, st);
    case 8: return launch_fwd<8>(bottom, rois, seg, order, out, argmax, g, pl.smem_fwd, st);
    default: return launch_fwd<4>(bottom, rois, seg, order, out, argmax, g, pl.smem_fwd, st);
  }
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
  L2S_REQUIRE(!g.maxpool || argmax, L2S_ERR_ARG, "roi_crop_bwd: max-pool mode needs the argmax of the forward");
  Plan pl;
  L2S_REQUIRE(make_plan(H * W, g.maxpool, &pl), L2S_ERR_SHAPE,
              "roi_crop: feature map %dx%d does not fit the shared-memory staging (max 12800 pixels)", H, W);
  int *seg, *order;
  rc = bin_rois(rois, B, N, workspace, st, &seg, &order);
  if (rc) return rc;
  if (!g.maxpool) {
    switch (pl.cc) {
      case 32: return launch_bwd_fast<32>(dout, rois, seg, order, dbottom, g, pl.smem_bwd, st);
      case 16: return launch_bwd_fast<16>(dout, rois, seg, order, dbottom, g, pl.smem_bwd, st);
      case 8: return launch_bwd_fast<8>(dout, rois, seg, order, dbottom, g, pl.smem_bwd, st);
      default: return launch_bwd_fast<4>(dout, rois, seg, order, dbottom, g, pl.smem_bwd, st);
    }
  }
  switch (pl.cc) {
    case 32: return launch_bwd<32, 2>(dout, rois, seg, order, argmax, dbottom, g, pl.smem_bwd, st);
    case 16: return launch_bwd<16, 4>(dout, rois, seg, order, argmax, dbottom, g, pl.smem_bwd, st);
    case 8: return launch_bwd<8, 4>(dout, rois, seg, order, argmax, dbottom, g, pl.smem_bwd, st);
    default: return launch_bwd<4, 4>(dout, rois, seg, order, argmax, dbottom, g, pl.smem_bwd, st);
  }
}
