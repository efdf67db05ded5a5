// One training step of FACT on one GPU: forward with saved activations, MSE loss, full backward into Keras-layout fp32
// gradients.  Reference semantics: mint/ctl/single_task_trainer.py:138-187 (forward, loss / num_replicas,
// tape.gradient) with FACTModel.call / loss (mint/core/fact_model.py:72-101, 134-148).  The all-reduce and the Adam
// update are separate entry points (fact_adam_step) so the host can put ONE NCCL all-reduce of the flat gradient
// bucket between them.
//
// Product path: bf16 operands, fp32 accumulation (BASELINE.json configs[2]: "training step bf16"); residual stream,
// LayerNorm statistics, softmax statistics, loss and all gradients of parameters are fp32.
#include <math.h>

#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {
int sdpa_run(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo, float* lse, int batch, int n,
             int heads, int head_dim, int q_rows, cudaStream_t st);
int wgrad_gemm(const void* x_bf16, int ldx, const void* dy_bf16, int ldy, float* dW, int ldw, int tokens, int in_dim,
               int out_dim, cudaStream_t st);
int sdpa_backward(const void* qkv, const void* o, const void* d_o, const float* lse, float* Dv, float* dq_acc,
                  void* dqkv, int batch, int n, int heads, int head_dim, float scale, cudaStream_t st);
int ln_backward(const float* x, const float* gamma, const float* dy, const float* dres, float* dx, void* dx_b,
                float* dgamma, float* dbeta, float* colsum, int rows, int d, cudaStream_t st);
int cast_colsum(const float* x, int ldx, void* y, int ldy, float* colsum, int rows, int n, cudaStream_t st);
int colsum_bf16(const void* x, int ldx, float* colsum, int rows, int n, cudaStream_t st);
int slice_rows(const float* src, float* dst, long long rows_dst, int d, int seq_dst, int seq_src, int seq_off,
               cudaStream_t st);
int embed_backward(const float* x, long long x_batch_stride, const float* dy, float* dW, float* dbias, float* dpos,
                   int batch, int n_tok, int f, int d, cudaStream_t st);
int pos_grad(const float* dy, float* dpos, int batch, int n_tok, int d, cudaStream_t st);
int embed_prep(const float* x, long long batch_stride, const int* step_ptr, void* hi, void* lo, int rows, int n_tok,
               int f, int kp, cudaStream_t st);
int pack_weight_ld(const float* w, void* hi, void* lo, int k_in, int n_out, int ld, cudaStream_t st);

struct SavedLayer {
  float* x_in;   // [M, d]   input of the layer (residual stream)
  bf16* ln1;     // [M, d]
  bf16* qkv;     // [M, 3d]
  float* lse;    // [B, H, N]
  bf16* ao;      // [M, d]
  float* x_mid;  // [M, d]
  bf16* ln2;     // [M, d]
  bf16* hpre;    // [M, ff]
  bf16* h;       // [M, ff]
};

struct TrainWs {
  SavedLayer* layers;  // host array [motion + audio + cross]
  float* xf;           // [Mc, d] output of the cross stack
  bf16* xf_b;          // [Mc, d]
  float* pred;         // [Mc, out]
  float* dpred;        // [Mc, out]
  bf16* dpred_b;       // [Mc, 256]
  float* partial;      // 1024
  float *dy, *dym, *dya, *dln;  // [Mc, d] fp32 (dym / dya: encoder-sized)
  bf16 *dy_b, *dz_b, *dao_b, *dqkv_b;
  float *Dscr, *dq_scr, *ln_stats;
  // LinearEmbedding on the tensor path: bf16 inputs [tokens, kp] (kept for the weight gradient), kernels [d, kp]
  bf16 *em_x, *ea_x, *em_w, *ea_w;
  int em_kp, ea_kp;
  int* ln_sync;  // arrival counters of the residual GEMMs that normalise their own rows (fact_gemm_epilogue.ln_sync)
  size_t ln_sync_bytes;
};

static size_t align_up(size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); }

static int head_pad(int out_dim) { return (out_dim + 63) / 64 * 64; }

// Lays out the training workspace; `layers_out` (host array) receives the per-layer saved-activation pointers.
static size_t carve_train(const fact_dims* dm, int batch, void* base, TrainWs* ws, SavedLayer* layers_out) {
  const size_t d = dm->d_model, ff = dm->d_ff, H = dm->n_heads;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return base ? static_cast<uint8_t*>(base) + o : nullptr;
  };
  const int nl[3] = {dm->motion_layers, dm->audio_layers, dm->cross_layers};
  const size_t seq[3] = {static_cast<size_t>(dm->motion_seq), static_cast<size_t>(dm->audio_seq),
                         static_cast<size_t>(dm->motion_seq + dm->audio_seq)};
  int li = 0;
  for (int stack = 0; stack < 3; ++stack)
    for (int l = 0; l < nl[stack]; ++l, ++li) {
      const size_t M = batch * seq[stack];
      SavedLayer s;
      s.x_in = reinterpret_cast<float*>(take(M * d * 4));
      s.ln1 = reinterpret_cast<bf16*>(take(M * d * 2));
      s.qkv = reinterpret_cast<bf16*>(take(M * 3 * d * 2));
      s.lse = reinterpret_cast<float*>(take(static_cast<size_t>(batch) * H * seq[stack] * 4));
      s.ao = reinterpret_cast<bf16*>(take(M * d * 2));
      s.x_mid = reinterpret_cast<float*>(take(M * d * 4));
      s.ln2 = reinterpret_cast<bf16*>(take(M * d * 2));
      s.hpre = reinterpret_cast<bf16*>(take(M * ff * 2));
      s.h = reinterpret_cast<bf16*>(take(M * ff * 2));
      if (layers_out) layers_out[li] = s;
    }
  const size_t Mc = batch * seq[2], Mm = batch * seq[0], Ma = batch * seq[1];
  const size_t od = dm->out_dim, odp = head_pad(dm->out_dim);
  TrainWs w{};
  w.xf = reinterpret_cast<float*>(take(Mc * d * 4));
  w.xf_b = reinterpret_cast<bf16*>(take(Mc * d * 2));
  w.pred = reinterpret_cast<float*>(take(Mc * od * 4));
  w.dpred = reinterpret_cast<float*>(take(Mc * od * 4));
  w.dpred_b = reinterpret_cast<bf16*>(take(Mc * odp * 2));
  w.partial = reinterpret_cast<float*>(take(1024 * 4));
  w.dy = reinterpret_cast<float*>(take(Mc * d * 4));
  w.dym = reinterpret_cast<float*>(take(Mm * d * 4));
  w.dya = reinterpret_cast<float*>(take(Ma * d * 4));
  w.dln = reinterpret_cast<float*>(take(Mc * d * 4));
  w.dy_b = reinterpret_cast<bf16*>(take(Mc * d * 2));
  w.dz_b = reinterpret_cast<bf16*>(take(Mc * ff * 2));
  w.dao_b = reinterpret_cast<bf16*>(take(Mc * d * 2));
  w.dqkv_b = reinterpret_cast<bf16*>(take(Mc * 3 * d * 2));
  w.Dscr = reinterpret_cast<float*>(take(static_cast<size_t>(batch) * H * seq[2] * 4));
  w.dq_scr = reinterpret_cast<float*>(take(Mc * d * 4));
  w.ln_stats = reinterpret_cast<float*>(take(Mc * 16));
  w.em_kp = (dm->motion_dim + 7) / 8 * 8;
  w.ea_kp = (dm->audio_dim + 7) / 8 * 8;
  w.em_x = reinterpret_cast<bf16*>(take(Mm * w.em_kp * 2));
  w.ea_x = reinterpret_cast<bf16*>(take(Ma * w.ea_kp * 2));
  w.em_w = reinterpret_cast<bf16*>(take(d * w.em_kp * 2));
  w.ea_w = reinterpret_cast<bf16*>(take(d * w.ea_kp * 2));
  w.ln_sync_bytes = 2 * ((Mc + 31) / 32 + 1) * sizeof(int);
  w.ln_sync = reinterpret_cast<int*>(take(w.ln_sync_bytes));
  if (ws) *ws = w;
  return off;
}

static int bf16_gemm(const bf16* a, int lda, const void* w, int ldw, int m, int n, int k, const fact_gemm_epilogue* e,
                     cudaStream_t st) {
  return fact_gemm(a, nullptr, lda, w, nullptr, ldw, m, n, k, e, st);
}

// LinearEmbedding + bias + PositionEmbedding (base_models.py:130-156) as a tensor-core GEMM on the bf16 inputs; the
// position table enters as a residual broadcast over the clips.  x_b is kept for the weight gradient.
static int embed_fwd_train(const float* x, const float* w_f32, const float* bias, const float* pos, float* y,
                           bf16* x_b, bf16* w_b, int kp, int batch, int n_tok, int f, int d, cudaStream_t st) {
  const int rows = batch * n_tok;
  int rc;
  if ((rc = embed_prep(x, static_cast<long long>(n_tok) * f, nullptr, x_b, nullptr, rows, n_tok, f, kp, st))) return rc;
  if ((rc = pack_weight_ld(w_f32, w_b, nullptr, f, d, kp, st))) return rc;
  fact_gemm_epilogue e{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = y;
  e.ldo = d;
  e.bias = bias;
  e.resid = pos;
  e.ldr = d;
  e.resid_rows = n_tok;
  return bf16_gemm(x_b, kp, w_b, kp, rows, d, f, &e, st);
}

// dy_b / the bias gradient were already produced by the LayerNorm backward of the encoder's first layer
static int embed_bwd_train(const bf16* x_b, int kp, const float* dy, const bf16* dy_b, float* dW, float* dpos, int batch,
                           int n_tok, int f, int d, cudaStream_t st) {
  int rc;
  if ((rc = wgrad_gemm(x_b, kp, dy_b, d, dW, d, batch * n_tok, f, d, st))) return rc;
  return pos_grad(dy, dpos, batch, n_tok, d, st);
}

// forward of one layer, saving what the backward needs; `out` = where the layer's output residual stream goes
// ln1_done: S.ln1 already holds LayerNorm1(S.x_in) (the FF2 call of the layer below produced it); next / next_ln1: the
// layer above in the same stack, whose LayerNorm1 this layer's FF2 call produces (NULL: nobody, or the output is
// remapped into the concatenated buffer).  The LayerNorms ride on the residual GEMM calls (fact_gemm_epilogue.ln_*):
// at training batch sizes the GEMM's own LayerNorm warps write them inside the launch.
static int layer_fwd_train(const fact_dims* dm, const fact_layer_weights& L, const SavedLayer& S, int batch, int seq,
                           float* out, int out_seq, int out_off, bool ln1_done, const fact_layer_weights* next,
                           bf16* next_ln1, int* ln_sync, cudaStream_t st) {
  const int d = dm->d_model, ff = dm->d_ff, H = dm->n_heads, dh = d / H, M = batch * seq;
  int rc;
  if (!ln1_done && (rc = fact_layernorm_split(S.x_in, L.ln1_gamma, L.ln1_beta, S.ln1, nullptr, M, d, st))) return rc;
  fact_gemm_epilogue e{};
  e.kind = FACT_EPI_SPLIT;
  e.out_hi = S.qkv;
  e.ldo = 3 * d;
  e.scale = static_cast<float>(1.0 / sqrt(static_cast<double>(d)) * 1.4426950408889634);
  e.scale_cols = d;
  if ((rc = bf16_gemm(S.ln1, d, L.wqkv_hi, d, M, 3 * d, d, &e, st))) return rc;
  if ((rc = sdpa_run(S.qkv, nullptr, S.ao, nullptr, S.lse, batch, seq, H, dh, seq, st))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = S.x_mid;
  e.ldo = d;
  e.bias = L.bo;
  e.resid = S.x_in;
  e.ldr = d;
  e.ln_gamma = L.ln2_gamma;  // LayerNorm2 of the rows this call writes
  e.ln_beta = L.ln2_beta;
  e.ln_hi = S.ln2;
  e.ln_sync = ln_sync;
  if ((rc = bf16_gemm(S.ao, d, L.wo_hi, d, M, d, d, &e, st))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_GELU_SAVE;
  e.out_hi = S.h;
  e.out_lo = S.hpre;
  e.ldo = ff;
  e.bias = L.b1;
  if ((rc = bf16_gemm(S.ln2, d, L.w1_hi, d, M, ff, d, &e, st))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = out;
  e.ldo = d;
  e.bias = L.b2;
  e.resid = S.x_mid;
  e.ldr = d;
  if (out_seq) {
    e.seq_in = seq;
    e.seq_out = out_seq;
    e.seq_off = out_off;
  } else if (next) {  // LayerNorm1 of the layer above
    e.ln_gamma = next->ln1_gamma;
    e.ln_beta = next->ln1_beta;
    e.ln_hi = next_ln1;
    e.ln_sync = ln_sync;
  }
  return bf16_gemm(S.h, ff, L.w2_hi, ff, M, d, ff, &e, st);
}

// backward of one layer: dy [M, d] fp32 holds dL/d(output) on entry and dL/d(input) on exit.
// dy_b_ready: ws.dy_b already holds bf16(dy) and G.b2 already has its column sums (the LayerNorm backward of the layer
// above wrote them); next_b2: bias gradient the layer below wants from this layer's output (NULL: it casts itself).
static int layer_bwd(const fact_dims* dm, const fact_layer_weights& L, const fact_layer_grads& G, const SavedLayer& S,
                     int batch, int seq, float* dy, const TrainWs& ws, bool dy_b_ready, float* next_b2,
                     cudaStream_t st) {
  const int d = dm->d_model, ff = dm->d_ff, H = dm->n_heads, dh = d / H, M = batch * seq;
  int rc;
  fact_gemm_epilogue e{};
  // ---- Residual(Norm(MLP))
  if (!dy_b_ready && (rc = cast_colsum(dy, d, ws.dy_b, d, G.b2, M, d, st))) return rc;
  if ((rc = wgrad_gemm(S.h, ff, ws.dy_b, d, G.w2, d, M, ff, d, st))) return rc;
  e.kind = FACT_EPI_GELU_GRAD;
  e.out_hi = ws.dz_b;
  e.ldo = ff;
  e.aux = S.hpre;
  e.ldaux = ff;
  e.colsum = G.b1;  // d b1 = column sums of dz, taken in the same epilogue
  if ((rc = bf16_gemm(ws.dy_b, d, L.w2_kl, d, M, ff, d, &e, st))) return rc;  // dh = dy . W2^T, then * gelu'(z)
  if ((rc = wgrad_gemm(S.ln2, d, ws.dz_b, ff, G.w1, ff, M, d, ff, st))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_F32;
  e.out_f32 = ws.dln;
  e.ldo = d;
  if ((rc = bf16_gemm(ws.dz_b, ff, L.w1_kl, ff, M, d, ff, &e, st))) return rc;  // d ln2 = dz . W1^T
  // dy becomes dL/d x_mid; its bf16 copy and column sums (= d bo) come out of the same pass
  if ((rc = ln_backward(S.x_mid, L.ln2_gamma, ws.dln, dy, dy, ws.dy_b, G.ln2_gamma, G.ln2_beta, G.bo, M, d, st)))
    return rc;
  // ---- Residual(Norm(Attention))
  if ((rc = wgrad_gemm(S.ao, d, ws.dy_b, d, G.wo, d, M, d, d, st))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_SPLIT;
  e.out_hi = ws.dao_b;
  e.ldo = d;
  e.scale = 1.f;
  e.scale_cols = 0;
  if ((rc = bf16_gemm(ws.dy_b, d, L.wo_kl, d, M, d, d, &e, st))) return rc;  // d ao = dy . Wo^T
  if ((rc = sdpa_backward(S.qkv, S.ao, ws.dao_b, S.lse, ws.Dscr, ws.dq_scr, ws.dqkv_b, batch, seq, H, dh,
                          static_cast<float>(1.0 / sqrt(static_cast<double>(d))), st)))
    return rc;
  if ((rc = wgrad_gemm(S.ln1, d, ws.dqkv_b, 3 * d, G.wqkv, 3 * d, M, d, 3 * d, st))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_F32;
  e.out_f32 = ws.dln;
  e.ldo = d;
  if ((rc = bf16_gemm(ws.dqkv_b, 3 * d, L.wqkv_kl, 3 * d, M, d, 3 * d, &e, st))) return rc;  // d ln1 = dqkv . Wqkv^T
  return ln_backward(S.x_in, L.ln1_gamma, ws.dln, dy, dy, next_b2 ? ws.dy_b : nullptr, G.ln1_gamma, G.ln1_beta, next_b2,
                     M, d, st);
}

}  // namespace fact

using namespace fact;

extern "C" size_t fact_train_workspace_bytes(const fact_dims* dims, int batch) {
  if (!dims || batch <= 0) return 0;
  return carve_train(dims, batch, nullptr, nullptr, nullptr);
}

extern "C" int fact_train_step(const fact_dims* dims, const fact_weights* w, const fact_grads* g, const float* motion,
                               const float* audio, const float* target, int target_len, int batch, float loss_scale,
                               float* loss_out, void* workspace, size_t workspace_bytes,
                               void* const* stage_events, int n_stage_events, void* stream) {
  FACT_REQUIRE(dims && w && g && motion && audio && target && loss_out && batch > 0, FACT_ERR_BAD_SHAPE,
               "fact_train_step: bad arguments");
  const int d = dims->d_model, ns = dims->motion_seq + dims->audio_seq, od = dims->out_dim, odp = head_pad(od);
  FACT_REQUIRE(d % dims->n_heads == 0 && d % 8 == 0 && dims->d_ff % 8 == 0 && d <= 1024, FACT_ERR_BAD_SHAPE,
               "fact_train_step: unsupported dims");
  FACT_REQUIRE(target_len > 0 && target_len <= ns, FACT_ERR_BAD_SHAPE, "target_len %d", target_len);
  FACT_REQUIRE(dims->motion_layers > 0 && dims->audio_layers > 0 && dims->cross_layers > 0, FACT_ERR_UNSUPPORTED,
               "every stack needs at least one layer");
  FACT_REQUIRE(w->out_w_kl && w->out_w_hi, FACT_ERR_BAD_SHAPE, "training needs the packed head weights");
  FACT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, FACT_ERR_BAD_ALIGN,
               "workspace must be 1024-B aligned");
  const int nl = dims->motion_layers + dims->audio_layers + dims->cross_layers;
  FACT_REQUIRE(nl <= 64, FACT_ERR_UNSUPPORTED, "too many layers");
  SavedLayer saved[64];
  TrainWs ws;
  const size_t need = carve_train(dims, batch, workspace, &ws, saved);
  FACT_REQUIRE(workspace && workspace_bytes >= need, FACT_ERR_WORKSPACE, "workspace too small: %zu < %zu",
               workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  const SavedLayer* Sm = saved;
  const SavedLayer* Sa = saved + dims->motion_layers;
  const SavedLayer* Sc = Sa + dims->audio_layers;
  const int Mc = batch * ns;
  int rc;
  // stage i of the backward is complete -> record stage_events[i] (see the header for the order)
  int stage = 0;
  auto stage_done = [&]() -> int {
    const int i = stage++;
    if (stage_events && i < n_stage_events && stage_events[i])
      FACT_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(stage_events[i]), st));
    return FACT_OK;
  };

  FACT_CUDA_CHECK(cudaMemsetAsync(ws.ln_sync, 0, ws.ln_sync_bytes, st));
  // ------------------------------------------------------------------ forward (saves activations)
  if ((rc = embed_fwd_train(motion, w->motion_embed_w, w->motion_embed_b, w->motion_pos, Sm[0].x_in, ws.em_x, ws.em_w,
                            ws.em_kp, batch, dims->motion_seq, dims->motion_dim, d, st)))
    return rc;
  for (int l = 0; l < dims->motion_layers; ++l) {
    const bool last = l + 1 == dims->motion_layers;
    if ((rc = layer_fwd_train(dims, w->motion_layers[l], Sm[l], batch, dims->motion_seq,
                              last ? Sc[0].x_in : Sm[l + 1].x_in, last ? ns : 0, 0, l > 0,
                              last ? nullptr : &w->motion_layers[l + 1], last ? nullptr : Sm[l + 1].ln1, ws.ln_sync, st)))
      return rc;
  }
  if ((rc = embed_fwd_train(audio, w->audio_embed_w, w->audio_embed_b, w->audio_pos, Sa[0].x_in, ws.ea_x, ws.ea_w,
                            ws.ea_kp, batch, dims->audio_seq, dims->audio_dim, d, st)))
    return rc;
  for (int l = 0; l < dims->audio_layers; ++l) {
    const bool last = l + 1 == dims->audio_layers;
    if ((rc = layer_fwd_train(dims, w->audio_layers[l], Sa[l], batch, dims->audio_seq,
                              last ? Sc[0].x_in : Sa[l + 1].x_in, last ? ns : 0, last ? dims->motion_seq : 0, l > 0,
                              last ? nullptr : &w->audio_layers[l + 1], last ? nullptr : Sa[l + 1].ln1, ws.ln_sync, st)))
      return rc;
  }
  for (int l = 0; l < dims->cross_layers; ++l) {
    const bool last = l + 1 == dims->cross_layers;
    if ((rc = layer_fwd_train(dims, w->cross_layers[l], Sc[l], batch, ns, last ? ws.xf : Sc[l + 1].x_in, 0, 0, l > 0,
                              last ? nullptr : &w->cross_layers[l + 1], last ? nullptr : Sc[l + 1].ln1, ws.ln_sync, st)))
      return rc;
  }
  if ((rc = fact_layernorm_split(ws.xf, nullptr, nullptr, ws.xf_b, nullptr, Mc, d, st))) return rc;
  fact_gemm_epilogue e{};
  e.kind = FACT_EPI_BIAS_F32;
  e.out_f32 = ws.pred;
  e.ldo = od;
  e.bias = w->out_b;
  if ((rc = bf16_gemm(ws.xf_b, d, w->out_w_hi, d, Mc, od, d, &e, st))) return rc;

  // ------------------------------------------------------------------ loss (fact_model.py:143-148) and dL/dpred
  if ((rc = fact_mse(target, ws.pred, loss_out, ws.dpred, ws.partial, batch, target_len, ns, od, loss_scale, st)))
    return rc;

  // ------------------------------------------------------------------ backward
  if ((rc = cast_colsum(ws.dpred, od, ws.dpred_b, odp, g->out_b, Mc, od, st))) return rc;
  if ((rc = wgrad_gemm(ws.xf_b, d, ws.dpred_b, odp, g->out_w, od, Mc, d, od, st))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_F32;
  e.out_f32 = ws.dy;
  e.ldo = d;
  if ((rc = bf16_gemm(ws.dpred_b, odp, w->out_w_kl, odp, Mc, d, odp, &e, st))) return rc;  // d xf = dpred . Wout^T
  if ((rc = stage_done())) return rc;  // stage 0: the output head
  for (int l = dims->cross_layers - 1; l >= 0; --l) {
    if ((rc = layer_bwd(dims, w->cross_layers[l], g->cross_layers[l], Sc[l], batch, ns, ws.dy, ws,
                        l + 1 < dims->cross_layers, l > 0 ? g->cross_layers[l - 1].b2 : nullptr, st)))
      return rc;
    // layer l's slice is final (its b2 sums were written by layer l+1's call; this call also wrote layer l-1's b2,
    // which belongs to a slice that is reduced later)
    if ((rc = stage_done())) return rc;
  }
  // tf.concat backward (base_models.py:192-193): rows [0, motion_seq) -> motion encoder, the rest -> audio encoder
  if ((rc = slice_rows(ws.dy, ws.dym, static_cast<long long>(batch) * dims->motion_seq, d, dims->motion_seq, ns, 0, st)))
    return rc;
  if ((rc = slice_rows(ws.dy, ws.dya, static_cast<long long>(batch) * dims->audio_seq, d, dims->audio_seq, ns,
                       dims->motion_seq, st)))
    return rc;
  // each encoder's first-layer LayerNorm backward also emits bf16(dy) and its column sums = the embedding bias gradient
  for (int l = dims->motion_layers - 1; l >= 0; --l) {
    if ((rc = layer_bwd(dims, w->motion_layers[l], g->motion_layers[l], Sm[l], batch, dims->motion_seq, ws.dym, ws,
                        l + 1 < dims->motion_layers, l > 0 ? g->motion_layers[l - 1].b2 : g->motion_embed_b, st)))
      return rc;
    if (l > 0 && (rc = stage_done())) return rc;
  }
  if ((rc = embed_bwd_train(ws.em_x, ws.em_kp, ws.dym, ws.dy_b, g->motion_embed_w, g->motion_pos, batch,
                            dims->motion_seq, dims->motion_dim, d, st)))
    return rc;
  if ((rc = stage_done())) return rc;  // motion layer 0 + its position table and LinearEmbedding
  for (int l = dims->audio_layers - 1; l >= 0; --l) {
    if ((rc = layer_bwd(dims, w->audio_layers[l], g->audio_layers[l], Sa[l], batch, dims->audio_seq, ws.dya, ws,
                        l + 1 < dims->audio_layers, l > 0 ? g->audio_layers[l - 1].b2 : g->audio_embed_b, st)))
      return rc;
    if (l > 0 && (rc = stage_done())) return rc;
  }
  if ((rc = embed_bwd_train(ws.ea_x, ws.ea_kp, ws.dya, ws.dy_b, g->audio_embed_w, g->audio_pos, batch, dims->audio_seq,
                            dims->audio_dim, d, st)))
    return rc;
  return stage_done();  // audio layer 0 + its position table and LinearEmbedding: the whole bucket is final
}
