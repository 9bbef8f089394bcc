/* Plain-C CPU restatement of the two ROI pooling modes of the lang2seg hot path.
 *
 * TEST INFRASTRUCTURE ONLY (oracle).  Built by oracle/Makefile into
 * oracle/_build/liboracle_c.so and loaded with ctypes by tests/ and bench.py's
 * cpu_baseline leg; never linked into the product library.
 *
 *  - oracle_roi_maxpool_{fwd,bwd}: Caffe RoI max-pool, following the arithmetic of
 *    pyutils/mask-faster-rcnn/lib/layer_utils/roi_pooling/src/cuda/roi_pooling_kernel.cu:15-75
 *    (forward, NCHW, argmax) and :104-179 (backward == scatter of top_diff to argmax).
 *  - oracle_crop_resize_fwd: closed form of Network._crop_pool_layer
 *    (pyutils/mask-faster-rcnn/lib/nets/network_cycle_response.py:107-149), SURVEY.md A.2.
 */
#include <float.h>
#include <math.h>
#include <string.h>

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

int oracle_roi_maxpool_fwd(const float* feat, const float* rois, int B, int C, int H, int W, int N,
                           int ph_n, int pw_n, float scale, float* out, int* argmax) {
  for (int n = 0; n < N; ++n) {
    const float* r = rois + 5 * n;
    int b = (int)r[0];
    if (b < 0 || b >= B) return -1;
    int rsw = (int)roundf(r[1] * scale), rsh = (int)roundf(r[2] * scale);
    int rew = (int)roundf(r[3] * scale), reh = (int)roundf(r[4] * scale);
    int rw = imax(rew - rsw + 1, 1), rh = imax(reh - rsh + 1, 1);
    float bh = (float)rh / (float)ph_n, bw = (float)rw / (float)pw_n;
    for (int c = 0; c < C; ++c)
      for (int ph = 0; ph < ph_n; ++ph)
        for (int pw = 0; pw < pw_n; ++pw) {
          int hs = imin(imax((int)floorf((float)ph * bh) + rsh, 0), H);
          int he = imin(imax((int)ceilf((float)(ph + 1) * bh) + rsh, 0), H);
          int ws = imin(imax((int)floorf((float)pw * bw) + rsw, 0), W);
          int we = imin(imax((int)ceilf((float)(pw + 1) * bw) + rsw, 0), W);
          int empty = (he <= hs) || (we <= ws);
          float best = empty ? 0.f : -FLT_MAX;
          int bi = -1;
          const float* f = feat + (size_t)b * C * H * W;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
              int idx = (c * H + h) * W + w;
              if (f[idx] > best) { best = f[idx]; bi = idx; }
            }
          size_t o = (((size_t)n * C + c) * ph_n + ph) * pw_n + pw;
          out[o] = best;
          argmax[o] = bi;
        }
  }
  return 0;
}

int oracle_roi_maxpool_bwd(const float* top, const float* rois, const int* argmax, int B, int C, int H,
                           int W, int N, int ph_n, int pw_n, float* bottom) {
  memset(bottom, 0, sizeof(float) * (size_t)B * C * H * W);
  size_t per = (size_t)C * ph_n * pw_n;
  for (int n = 0; n < N; ++n) {
    int b = (int)rois[5 * n];
    for (size_t i = 0; i < per; ++i) {
      int a = argmax[n * per + i];
      if (a >= 0) bottom[(size_t)b * C * H * W + a] += top[n * per + i];
    }
  }
  return 0;
}

/* mode 0: x/16 ; mode 1 (align): x/(imW-1)*(W-1).  S samples per side, optional 2x2 max. */
int oracle_crop_resize_fwd(const float* feat, const float* rois, int B, int C, int H, int W, int N, int S,
                           int max_pool, int align, float im_h, float im_w, float* out) {
  int P = max_pool ? S / 2 : S;
  for (int n = 0; n < N; ++n) {
    const float* r = rois + 5 * n;
    int b = (int)r[0];
    if (b < 0 || b >= B) return -1;
    double x1, y1, x2, y2;
    if (align) {
      x1 = r[1] / (im_w - 1) * (W - 1); x2 = r[3] / (im_w - 1) * (W - 1);
      y1 = r[2] / (im_h - 1) * (H - 1); y2 = r[4] / (im_h - 1) * (H - 1);
    } else { x1 = r[1] / 16.0; y1 = r[2] / 16.0; x2 = r[3] / 16.0; y2 = r[4] / 16.0; }
    for (int c = 0; c < C; ++c) {
      const float* f = feat + ((size_t)b * C + c) * H * W;
      float* o = out + ((size_t)n * C + c) * P * P;
      if (max_pool) for (int i = 0; i < P * P; ++i) o[i] = -FLT_MAX;
      for (int i = 0; i < S; ++i)
        for (int j = 0; j < S; ++j) {
          double py = y1 + (y2 - y1) * i / (S - 1), px = x1 + (x2 - x1) * j / (S - 1);
          int y0 = (int)floor(py), x0 = (int)floor(px);
          double ly = py - y0, lx = px - x0, v = 0;
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
              int yy = y0 + dy, xx = x0 + dx;
              if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
              v += (dy ? ly : 1 - ly) * (dx ? lx : 1 - lx) * f[yy * W + xx];
            }
          if (max_pool) { float* q = o + (i / 2) * P + j / 2; if ((float)v > *q) *q = (float)v; }
          else o[i * S + j] = (float)v;
        }
    }
  }
  return 0;
}
