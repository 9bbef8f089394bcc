// Variable-length bidirectional LSTM recurrence of the expression encoder (RNNEncoder.forward,
// lib/layers/lang_encoder.py:38-80: pack_padded_sequence -> nn.LSTM -> pad_packed_sequence) as one library call
// per direction of autograd.
//
// Packing is replaced by masking: sequence b runs for exactly len[b] steps in each direction -- at time t >= len[b]
// the state is left untouched and the output row is zero, which is what pack/unpack computes (the backward direction
// starts at t = len[b]-1 because its state is still the zero initial state before that).  The input projection
// x_t W_ih^T + b for all (b,t) and both directions is ONE GEMM done by the caller (G = xg, shape (B,L,8H), gate order
// i,f,g,o per direction); per time step the library adds h W_hh^T with the skinny exact-fp32 GEMM of decode.cu and runs
// one cell kernel for both directions.  No cuDNN, no host-side sorting, no device->host read of the lengths, so the
// whole encoder is CUDA-graph capturable.
#include "common.cuh"

namespace l2s {
namespace {

struct LstmGeom {
  int L, B, H;
};

// G: (B,L,2,4H) pre-activations (xg + h W_hh^T).  State buffers: (L,2,B,H) indexed by TIME t.
// step s: direction 0 works on t = s, direction 1 on t = L-1-s.
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ G, const int* __restrict__ lens, float* __restrict__ c_all,
                                     float* __restrict__ h_all, float* __restrict__ out, int s, LstmGeom g) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = g.H, B = g.B, L = g.L;
  if (idx >= 2 * B * H) return;
  const int dir = idx / (B * H), r = idx - dir * B * H, b = r / H, j = r - b * H;
  const int t = dir ? L - 1 - s : s;
  const int tp = dir ? t + 1 : t - 1;                       // time of the previous state of this direction
  const bool first = (s == 0);
  const size_t st = ((size_t)(t * 2 + dir) * B + b) * H + j;
  const size_t sp = ((size_t)(tp * 2 + dir) * B + b) * H + j;
  const float cp = first ? 0.f : c_all[sp], hp = first ? 0.f : h_all[sp];
  float c = cp, h = hp, o_out = 0.f;
  if (t < lens[b]) {
    const float* gp = G + (((size_t)b * L + t) * 2 + dir) * 4 * H;
    const float ig = sigmoidf_acc(gp[j]), fg = sigmoidf_acc(gp[H + j]);
    const float gg = tanhf(gp[2 * H + j]), og = sigmoidf_acc(gp[3 * H + j]);
    c = fmaf(fg, cp, ig * gg);
    h = og * tanhf(c);
    o_out = h;
  }
  c_all[st] = c;
  h_all[st] = h;
  out[((size_t)b * L + t) * 2 * H + dir * H + j] = o_out;
}

// dG: (B,L,2,4H) gradient on the pre-activations (zero where masked).  dh_carry / dc_carry: (2,B,H) ping-pong.
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ G, const int* __restrict__ lens,
                                     const float* __restrict__ c_all, const float* __restrict__ dout,
                                     const float* __restrict__ dh_rec /* (2,B,H) from the next step's GEMM, or null */,
                                     const float* __restrict__ dh_pass /* (2,B,H) gradient passed through masked steps */,
                                     const float* __restrict__ dc_in, float* __restrict__ dG,
                                     float* __restrict__ dh_pass_out, float* __restrict__ dc_out, int s, LstmGeom g) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = g.H, B = g.B, L = g.L;
  if (idx >= 2 * B * H) return;
  const int dir = idx / (B * H), r = idx - dir * B * H, b = r / H, j = r - b * H;
  const int t = dir ? L - 1 - s : s;
  const int tp = dir ? t + 1 : t - 1;
  const bool first = (s == 0);
  // gradient arriving at the state AFTER time t of this direction
  const float dh_state = (dh_rec ? dh_rec[idx] : 0.f) + (dh_pass ? dh_pass[idx] : 0.f);
  const float dc_state = dc_in ? dc_in[idx] : 0.f;
  float* dgp = dG + (((size_t)b * L + t) * 2 + dir) * 4 * H;
  if (t < lens[b]) {
    const float* gp = G + (((size_t)b * L + t) * 2 + dir) * 4 * H;
    const float ig = sigmoidf_acc(gp[j]), fg = sigmoidf_acc(gp[H + j]);
    const float gg = tanhf(gp[2 * H + j]), og = sigmoidf_acc(gp[3 * H + j]);
    const size_t st = ((size_t)(t * 2 + dir) * B + b) * H + j;
    const float cp = first ? 0.f : c_all[((size_t)(tp * 2 + dir) * B + b) * H + j];
    const float tc = tanhf(c_all[st]);
    const float dh = dh_state + (dout ? dout[((size_t)b * L + t) * 2 * H + dir * H + j] : 0.f);
    const float dc = dc_state + dh * og * (1.f - tc * tc);
    dgp[j] = dc * gg * ig * (1.f - ig);
    dgp[H + j] = dc * cp * fg * (1.f - fg);
    dgp[2 * H + j] = dc * ig * (1.f - gg * gg);
    dgp[3 * H + j] = dh * tc * og * (1.f - og);
    dc_out[idx] = dc * fg;
    dh_pass_out[idx] = 0.f;          // the previous state receives its dh through dG . W_hh
  } else {
    dgp[j] = 0.f; dgp[H + j] = 0.f; dgp[2 * H + j] = 0.f; dgp[3 * H + j] = 0.f;
    dc_out[idx] = dc_state;          // untouched state: gradients pass straight through
    dh_pass_out[idx] = dh_state;
  }
}

// hidden (B,2H) = [h_fwd after t = L-1 | h_bwd after t = 0]
__global__ void lstm_final_kernel(const float* __restrict__ h_all, float* __restrict__ hidden, LstmGeom g) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = g.H, B = g.B, L = g.L;
  if (idx >= 2 * B * H) return;
  const int dir = idx / (B * H), r = idx - dir * B * H, b = r / H, j = r - b * H;
  const int t = dir ? 0 : L - 1;
  hidden[(size_t)b * 2 * H + dir * H + j] = h_all[((size_t)(t * 2 + dir) * B + b) * H + j];
}

__global__ void lstm_seed_kernel(const float* __restrict__ dhidden, float* __restrict__ dh_pass, LstmGeom g) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = g.H, B = g.B;
  if (idx >= 2 * B * H) return;
  const int dir = idx / (B * H), r = idx - dir * B * H, b = r / H, j = r - b * H;
  dh_pass[idx] = dhidden ? dhidden[(size_t)b * 2 * H + dir * H + j] : 0.f;
}

int check_lstm(int L, int B, int H) {
  L2S_REQUIRE(L > 0 && B > 0 && H > 0 && H % 4 == 0, L2S_ERR_SHAPE, "bilstm: bad shape L=%d B=%d H=%d (H %% 4 == 0)", L, B, H);
  return L2S_OK;
}

size_t lstm_ws_bytes(int B, int H) {
  size_t lin = linear_small_workspace_bytes(B, 4 * H, H);
  const size_t lin2 = linear_small_workspace_bytes(B, H, 4 * H);
  lin = ((lin > lin2 ? lin : lin2) + 255) & ~(size_t)255;
  // two GEMM workspaces + dh_rec, 2x dh_pass, 2x dc, spare + 256 B of grid-barrier counters (persistent kernels)
  return 2 * lin + (size_t)6 * 2 * B * H * sizeof(float) + 256 + 256;
}

}  // namespace
}  // namespace l2s

using namespace l2s;

extern "C" size_t l2s_bilstm_workspace_bytes(int L, int B, int H) {
  (void)L;
  return lstm_ws_bytes(B, H);
}

extern "C" int l2s_bilstm_fwd(float* G, const float* w_hh, const int32_t* lens, float* c_all, float* h_all, float* out,
                              float* hidden, int L, int B, int H, void* workspace, size_t workspace_bytes,
                              l2s_stream_t stream) {
  L2S_REQUIRE(G && w_hh && lens && c_all && h_all && out && hidden, L2S_ERR_ARG, "bilstm_fwd: null pointer");
  int rc = check_lstm(L, B, H);
  if (rc) return rc;
  L2S_REQUIRE(workspace && workspace_bytes >= lstm_ws_bytes(B, H), L2S_ERR_WORKSPACE, "bilstm_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const LstmGeom g{L, B, H};
  char* ws = reinterpret_cast<char*>(workspace);
  if (bilstm_persist_ok(L, B, H))      // one persistent weight-stationary kernel for all L steps (lstm_persist.cu)
    return launch_bilstm_fwd_persist(G, w_hh, lens, c_all, h_all, out, hidden, L, B,
                                     reinterpret_cast<unsigned*>(ws + ((lstm_ws_bytes(B, H) - 256) & ~(size_t)255)), st);
  const size_t lin = (lstm_ws_bytes(B, H) - 512 - (size_t)6 * 2 * B * H * sizeof(float)) / 2;
  L2S_CUDA_OK(cudaMemsetAsync(ws, 0, linear_small_counter_bytes(B, 4 * H), st));
  L2S_CUDA_OK(cudaMemsetAsync(ws + lin, 0, linear_small_counter_bytes(B, 4 * H), st));
  const int threads = 256, blocks = (2 * B * H + threads - 1) / threads;
  const int ldg = L * 8 * H;
  for (int s = 0; s < L; ++s) {
    if (s > 0) {
      for (int dir = 0; dir < 2; ++dir) {
        const int t = dir ? L - 1 - s : s, tp = dir ? t + 1 : t - 1;
        rc = launch_linear_small(h_all + ((size_t)(tp * 2 + dir) * B) * H, H, w_hh + (size_t)dir * 4 * H * H, H, nullptr,
                                 G + ((size_t)t * 2 + dir) * 4 * H, ldg, B, 4 * H, H, 1, ws + dir * lin, lin, st);
        if (rc) return rc;
      }
    }
    L2S_CUDA_OK(launch_chain(lstm_cell_fwd_kernel, dim3(blocks), dim3(threads), 0, st, (const float*)G, lens, c_all, h_all, out, s, g));
    L2S_LAUNCH_OK("lstm_cell_fwd_kernel");
    count_launch();
  }
  L2S_CUDA_OK(launch_chain(lstm_final_kernel, dim3(blocks), dim3(threads), 0, st, (const float*)h_all, hidden, g));
  L2S_LAUNCH_OK("lstm_final_kernel");
  count_launch();
  return L2S_OK;
}

extern "C" int l2s_bilstm_bwd(const float* dout, const float* dhidden, const float* G, const float* w_hh_t,
                              const int32_t* lens, const float* c_all, float* dG, int L, int B, int H, void* workspace,
                              size_t workspace_bytes, l2s_stream_t stream) {
  L2S_REQUIRE(G && w_hh_t && lens && c_all && dG && (dout || dhidden), L2S_ERR_ARG, "bilstm_bwd: null pointer");
  int rc = check_lstm(L, B, H);
  if (rc) return rc;
  L2S_REQUIRE(workspace && workspace_bytes >= lstm_ws_bytes(B, H), L2S_ERR_WORKSPACE, "bilstm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const LstmGeom g{L, B, H};
  char* ws = reinterpret_cast<char*>(workspace);
  if (bilstm_persist_ok(L, B, H))
    return launch_bilstm_bwd_persist(dout, dhidden, G, w_hh_t, lens, c_all, dG, L, B,
                                     reinterpret_cast<unsigned*>(ws + ((lstm_ws_bytes(B, H) - 256) & ~(size_t)255)), st);
  const size_t lin = (lstm_ws_bytes(B, H) - 512 - (size_t)6 * 2 * B * H * sizeof(float)) / 2;
  L2S_CUDA_OK(cudaMemsetAsync(ws, 0, linear_small_counter_bytes(B, 4 * H), st));
  L2S_CUDA_OK(cudaMemsetAsync(ws + lin, 0, linear_small_counter_bytes(B, 4 * H), st));
  const size_t n = (size_t)2 * B * H;
  float* fb = reinterpret_cast<float*>(ws + 2 * lin);
  float* dh_rec = fb;
  float* dh_pass[2] = {fb + n, fb + 2 * n};
  float* dc[2] = {fb + 3 * n, fb + 4 * n};
  const int threads = 256, blocks = (int)((n + threads - 1) / threads);
  const int ldg = L * 8 * H;
  L2S_CUDA_OK(launch_chain(lstm_seed_kernel, dim3(blocks), dim3(threads), 0, st, dhidden, dh_pass[0], g));
  L2S_LAUNCH_OK("lstm_seed_kernel");
  count_launch();
  for (int s = L - 1, k = 0; s >= 0; --s, ++k) {
    const bool last = (s == L - 1);
    L2S_CUDA_OK(launch_chain(lstm_cell_bwd_kernel, dim3(blocks), dim3(threads), 0, st, G, lens, c_all, dout,
                             (const float*)(last ? nullptr : dh_rec), (const float*)dh_pass[k & 1],
                             (const float*)(last ? nullptr : dc[k & 1]), dG, dh_pass[(k & 1) ^ 1], dc[(k & 1) ^ 1], s, g));
    L2S_LAUNCH_OK("lstm_cell_bwd_kernel");
    count_launch();
    if (s > 0) {   // gradient on the previous state: dG_t . W_hh   (w_hh_t = W_hh^T stored (2, H, 4H))
      for (int dir = 0; dir < 2; ++dir) {
        const int t = dir ? L - 1 - s : s;
        rc = launch_linear_small(dG + ((size_t)t * 2 + dir) * 4 * H, ldg, w_hh_t + (size_t)dir * H * 4 * H, 4 * H, nullptr,
                                 dh_rec + (size_t)dir * B * H, H, B, H, 4 * H, 0, ws + dir * lin, lin, st);
        if (rc) return rc;
      }
    }
  }
  return L2S_OK;
}
