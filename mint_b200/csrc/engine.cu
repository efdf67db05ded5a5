// Whole-model sequencing of the FACT hot path on one GPU (C++ runtime above the kernels):
//   fact_forward                -- FACTModel.call                  (mint/core/fact_model.py:72-101)
//   fact_infer_auto_regressive  -- FACTModel.infer_auto_regressive (mint/core/fact_model.py:103-132)
// One transformer layer (base_models.py:102-106) is 7 launches:
//   LN+split -> QKV GEMM (split epilogue, q pre-scaled) -> attention core -> out-proj GEMM (+bias +residual, in place)
//   LN+split -> FF1 GEMM (+bias, GELU, split) -> FF2 GEMM (+bias +residual)
// The last FF2 of each modality encoder writes straight into its slice of the [B, 360, d] cross-modal buffer
// (the tf.concat of base_models.py:192-193 costs nothing).  The AR loop keeps a device-side step counter: the
// shift-by-one of fact_model.py:131 and the audio window slice of :124 are row offsets read by the embedding
// kernel, so one captured CUDA graph is replayed once per generated frame.
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {
int gemm_simt_split(const void* a_hi, const void* a_lo, int lda, const float* w_keras, int m, int n, int k,
                    const fact_gemm_epilogue* epi, cudaStream_t st);
int step_set(int* p, int v, cudaStream_t st);
int sdpa_run(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo, float* lse, int batch, int n,
             int heads, int head_dim, int q_rows, cudaStream_t st);
int step_inc(int* p, cudaStream_t st);

struct Workspace {
  float *xm, *xa, *xc, *xr;
  bf16 *ln_hi, *ln_lo, *ao_hi, *ao_lo, *qkv_hi, *qkv_lo, *h_hi, *h_lo;
  float* splitk;        // split-K partial slabs (small-M GEMMs)
  size_t splitk_bytes;
  // second scratch set so that the audio encoder can run concurrently with the motion encoder (small batches only)
  bf16 *a_ln_hi, *a_ln_lo, *a_ao_hi, *a_ao_lo, *a_qkv_hi, *a_qkv_lo, *a_h_hi, *a_h_lo;
  float* a_splitk;
  size_t a_splitk_bytes;
  // arrival counters of the GEMMs that normalise their own output rows (fact_gemm_epilogue.ln_sync): one int per 32
  // rows, zeroed at the start of every API call, left zero by every GEMM; the concurrent audio encoder has its own
  int *ln_sync, *a_ln_sync;
  size_t ln_sync_bytes, a_ln_sync_bytes;
  // tensor-core embedding (batch * seq >= kEmbedTcMinRows): split inputs and packed LinearEmbedding kernels
  bf16 *em_x_hi, *em_x_lo, *ea_x_hi, *ea_x_lo;   // [tokens, kp]
  bf16 *em_w_hi, *em_w_lo, *ea_w_hi, *ea_w_lo;   // [d, kp]
  int em_kp, ea_kp;
};
constexpr size_t kEmbedTcMinRows = 2048;  // below this the single CUDA-core launch wins (batch-1 decode)

int g_dual_stream = 1;  // fact_set_flag("dual_stream", 0): run both modality encoders on the caller's stream
constexpr size_t kDualStreamMaxTokens = 4096;  // audio tokens (batch * 240) up to which the encoders run concurrently

static size_t align_up(size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); }

static size_t carve(const fact_dims* dm, int batch, int mode, void* base, Workspace* ws) {
  const size_t d = dm->d_model, ff = dm->d_ff;
  const size_t tm = static_cast<size_t>(batch) * dm->motion_seq, ta = static_cast<size_t>(batch) * dm->audio_seq;
  const size_t tc = tm + ta;
  const bool lo = mode != FACT_MODE_BF16;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return base ? static_cast<uint8_t*>(base) + o : nullptr;
  };
  Workspace w{};
  w.xm = reinterpret_cast<float*>(take(tm * d * 4));
  w.xa = reinterpret_cast<float*>(take(ta * d * 4));
  w.xc = reinterpret_cast<float*>(take(tc * d * 4));
  w.xr = reinterpret_cast<float*>(take(static_cast<size_t>(batch) * d * 4));  // row 0 of every clip (AR tail)
  w.ln_hi = reinterpret_cast<bf16*>(take(tc * d * 2));
  w.ln_lo = lo ? reinterpret_cast<bf16*>(take(tc * d * 2)) : nullptr;
  w.ao_hi = reinterpret_cast<bf16*>(take(tc * d * 2));
  w.ao_lo = lo ? reinterpret_cast<bf16*>(take(tc * d * 2)) : nullptr;
  w.qkv_hi = reinterpret_cast<bf16*>(take(tc * 3 * d * 2));
  w.qkv_lo = lo ? reinterpret_cast<bf16*>(take(tc * 3 * d * 2)) : nullptr;
  w.h_hi = reinterpret_cast<bf16*>(take(tc * ff * 2));
  w.h_lo = lo ? reinterpret_cast<bf16*>(take(tc * ff * 2)) : nullptr;
  // up to 16 slabs of the widest small-M GEMM output (GEMMs with m <= 1024: batch-1/2 decode, the row-0 AR tail)
  w.splitk_bytes = 16 * (tc < 1024 ? tc : 1024) * (3 * d > ff ? 3 * d : ff) * 4;
  w.splitk = reinterpret_cast<float*>(take(w.splitk_bytes));
  w.ln_sync_bytes = 2 * ((tc + 31) / 32 + 1) * sizeof(int);
  w.ln_sync = reinterpret_cast<int*>(take(w.ln_sync_bytes));
  if (ta <= kDualStreamMaxTokens) {
    w.a_ln_hi = reinterpret_cast<bf16*>(take(ta * d * 2));
    w.a_ln_lo = lo ? reinterpret_cast<bf16*>(take(ta * d * 2)) : nullptr;
    w.a_ao_hi = reinterpret_cast<bf16*>(take(ta * d * 2));
    w.a_ao_lo = lo ? reinterpret_cast<bf16*>(take(ta * d * 2)) : nullptr;
    w.a_qkv_hi = reinterpret_cast<bf16*>(take(ta * 3 * d * 2));
    w.a_qkv_lo = lo ? reinterpret_cast<bf16*>(take(ta * 3 * d * 2)) : nullptr;
    w.a_h_hi = reinterpret_cast<bf16*>(take(ta * ff * 2));
    w.a_h_lo = lo ? reinterpret_cast<bf16*>(take(ta * ff * 2)) : nullptr;
    w.a_splitk_bytes = 16 * (ta < 1024 ? ta : 1024) * (3 * d > ff ? 3 * d : ff) * 4;
    w.a_splitk = reinterpret_cast<float*>(take(w.a_splitk_bytes));
    w.a_ln_sync_bytes = 2 * ((ta + 31) / 32 + 1) * sizeof(int);
    w.a_ln_sync = reinterpret_cast<int*>(take(w.a_ln_sync_bytes));
  }
  if (mode != FACT_MODE_FP32_SIMT && tm >= kEmbedTcMinRows) {
    w.em_kp = (dm->motion_dim + 7) / 8 * 8;
    w.ea_kp = (dm->audio_dim + 7) / 8 * 8;
    w.em_x_hi = reinterpret_cast<bf16*>(take(tm * w.em_kp * 2));
    w.em_x_lo = lo ? reinterpret_cast<bf16*>(take(tm * w.em_kp * 2)) : nullptr;
    w.ea_x_hi = reinterpret_cast<bf16*>(take(ta * w.ea_kp * 2));
    w.ea_x_lo = lo ? reinterpret_cast<bf16*>(take(ta * w.ea_kp * 2)) : nullptr;
    w.em_w_hi = reinterpret_cast<bf16*>(take(d * w.em_kp * 2));
    w.em_w_lo = lo ? reinterpret_cast<bf16*>(take(d * w.em_kp * 2)) : nullptr;
    w.ea_w_hi = reinterpret_cast<bf16*>(take(d * w.ea_kp * 2));
    w.ea_w_lo = lo ? reinterpret_cast<bf16*>(take(d * w.ea_kp * 2)) : nullptr;
  }
  if (ws) *ws = w;
  return off;
}

// once per API call, outside any captured frame
static int zero_ln_sync(const Workspace& ws, cudaStream_t st) {
  FACT_CUDA_CHECK(cudaMemsetAsync(ws.ln_sync, 0, ws.ln_sync_bytes, st));
  if (ws.a_ln_sync) FACT_CUDA_CHECK(cudaMemsetAsync(ws.a_ln_sync, 0, ws.a_ln_sync_bytes, st));
  return FACT_OK;
}

static int check_dims(const fact_dims* dm) {
  FACT_REQUIRE(dm != nullptr, FACT_ERR_BAD_SHAPE, "null dims");
  FACT_REQUIRE(dm->d_model > 0 && dm->n_heads > 0 && dm->d_model % dm->n_heads == 0, FACT_ERR_BAD_SHAPE,
               "d_model %d not divisible by heads %d", dm->d_model, dm->n_heads);
  FACT_REQUIRE(dm->d_model % 8 == 0 && dm->d_ff % 8 == 0, FACT_ERR_BAD_ALIGN,
               "d_model and d_ff must be multiples of 8 (TMA 16-B row pitch)");
  FACT_REQUIRE(dm->d_model <= 1024, FACT_ERR_UNSUPPORTED, "d_model > 1024 not supported by the LayerNorm kernel");
  const int dh = dm->d_model / dm->n_heads;
  FACT_REQUIRE(dh == 16 || dh == 32 || dh == 64 || dh == 80, FACT_ERR_UNSUPPORTED, "head_dim %d not instantiated", dh);
  return FACT_OK;
}

// GEMM dispatch on mode
static int dense(int mode, const bf16* a_hi, const bf16* a_lo, int lda, const void* w_hi, const void* w_lo,
                 const float* w_f32, int m, int n, int k, const fact_gemm_epilogue* e_in, cudaStream_t st,
                 const Workspace* ws = nullptr) {
  fact_gemm_epilogue e_copy = *e_in;
  if (ws && ws->splitk) {
    e_copy.splitk_scratch = ws->splitk;
    e_copy.splitk_scratch_bytes = ws->splitk_bytes;
  }
  const fact_gemm_epilogue* e = &e_copy;
  if (mode == FACT_MODE_FP32_SIMT) return gemm_simt_split(a_hi, a_lo, lda, w_f32, m, n, k, e, st);
  if (mode == FACT_MODE_PRECISE) {
    FACT_REQUIRE(w_lo != nullptr, FACT_ERR_BAD_SHAPE, "precise mode needs the lo half of every packed weight");
    return fact_gemm(a_hi, a_lo, lda, w_hi, w_lo, k, m, n, k, e, st);
  }
  return fact_gemm(a_hi, nullptr, lda, w_hi, nullptr, k, m, n, k, e, st);
}

// One transformer layer on x [tokens, d] (in place unless `dst` remaps the final residual write).
// ln1_done: ws.ln_hi / ln_lo already hold LayerNorm1(x) (the previous layer's FF2 call produced it); next: the layer
// that follows in the same stack, whose LayerNorm1 the FF2 call of this layer produces (NULL: nobody, or remapped dst).
// The LayerNorm rides on the GEMM call (fact_gemm_epilogue.ln_*): fused into the split-K finish kernel for small
// batches, a separate launch inside the call otherwise -- the same kernels as before at large batch.
static int run_layer(const fact_dims* dm, const fact_layer_weights& L, float* x, int batch, int seq, int mode,
                     const Workspace& ws, float* dst, int dst_seq, int dst_off, bool ln1_done,
                     const fact_layer_weights* next, cudaStream_t st) {
  const int d = dm->d_model, ff = dm->d_ff, H = dm->n_heads, dh = d / H;
  const int M = batch * seq;
  const bool lo = mode != FACT_MODE_BF16;
  const bool ride = mode != FACT_MODE_FP32_SIMT;  // the CUDA-core debug GEMM has no LayerNorm option
  int rc;
  // --- Residual(Norm(Attention))
  if (!ln1_done &&
      (rc = fact_layernorm_split(x, L.ln1_gamma, L.ln1_beta, ws.ln_hi, lo ? ws.ln_lo : nullptr, M, d, st)))
    return rc;
  fact_gemm_epilogue e{};
  e.kind = FACT_EPI_SPLIT;
  e.out_hi = ws.qkv_hi;
  e.out_lo = lo ? ws.qkv_lo : nullptr;
  e.ldo = 3 * d;
  e.scale = static_cast<float>(1.0 / sqrt(static_cast<double>(d)) * 1.4426950408889634);  // d_model^-0.5 * log2(e)
  e.scale_cols = d;
  if ((rc = dense(mode, ws.ln_hi, ws.ln_lo, d, L.wqkv_hi, L.wqkv_lo, L.wqkv_f32, M, 3 * d, d, &e, st, &ws))) return rc;
  if ((rc = fact_sdpa(ws.qkv_hi, lo ? ws.qkv_lo : nullptr, ws.ao_hi, lo ? ws.ao_lo : nullptr, batch, seq, H, dh, st)))
    return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = x;
  e.ldo = d;
  e.bias = L.bo;
  e.resid = x;
  e.ldr = d;
  if (ride) {  // --- Residual(Norm(MLP)): LayerNorm2 of the rows this call writes
    e.ln_gamma = L.ln2_gamma;
    e.ln_beta = L.ln2_beta;
    e.ln_hi = ws.ln_hi;
    e.ln_lo = lo ? ws.ln_lo : nullptr;
    e.ln_sync = ws.ln_sync;
  }
  if ((rc = dense(mode, ws.ao_hi, ws.ao_lo, d, L.wo_hi, L.wo_lo, L.wo_f32, M, d, d, &e, st, &ws))) return rc;
  if (!ride &&
      (rc = fact_layernorm_split(x, L.ln2_gamma, L.ln2_beta, ws.ln_hi, lo ? ws.ln_lo : nullptr, M, d, st)))
    return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_GELU_SPLIT;
  e.out_hi = ws.h_hi;
  e.out_lo = lo ? ws.h_lo : nullptr;
  e.ldo = ff;
  e.bias = L.b1;
  if ((rc = dense(mode, ws.ln_hi, ws.ln_lo, d, L.w1_hi, L.w1_lo, L.w1_f32, M, ff, d, &e, st, &ws))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = dst ? dst : x;
  e.ldo = d;
  e.bias = L.b2;
  e.resid = x;
  e.ldr = d;
  if (dst) {
    e.seq_in = seq;
    e.seq_out = dst_seq;
    e.seq_off = dst_off;
  } else if (next && ride) {  // LayerNorm1 of the next layer of this stack
    e.ln_gamma = next->ln1_gamma;
    e.ln_beta = next->ln1_beta;
    e.ln_hi = ws.ln_hi;
    e.ln_lo = lo ? ws.ln_lo : nullptr;
    e.ln_sync = ws.ln_sync;
  }
  return dense(mode, ws.h_hi, ws.h_lo, ff, L.w2_hi, L.w2_lo, L.w2_f32, M, d, ff, &e, st, &ws);
}

// Last cross-modal layer of an AR step: infer_auto_regressive keeps only output row 0 of each clip
// (fact_model.py:128), and rows do not mix after the attention core, so everything downstream of it runs on
// `batch` rows instead of batch*seq: the attention core computes query block 0 only, out-proj / LN / FFN read row 0 of
// each clip through a row pitch of seq*d and write the compact buffer ws.xr [batch, d].  Results for row 0 are
// bit-identical to the full layer.
static int run_layer_row0(const fact_dims* dm, const fact_layer_weights& L, float* x, int batch, int seq, int mode,
                          const Workspace& ws, cudaStream_t st) {
  const int d = dm->d_model, ff = dm->d_ff, H = dm->n_heads, dh = d / H;
  const int M = batch * seq;
  const bool lo = mode != FACT_MODE_BF16;
  int rc;
  if ((rc = fact_layernorm_split(x, L.ln1_gamma, L.ln1_beta, ws.ln_hi, lo ? ws.ln_lo : nullptr, M, d, st))) return rc;
  fact_gemm_epilogue e{};
  e.kind = FACT_EPI_SPLIT;
  e.out_hi = ws.qkv_hi;
  e.out_lo = lo ? ws.qkv_lo : nullptr;
  e.ldo = 3 * d;
  e.scale = static_cast<float>(1.0 / sqrt(static_cast<double>(d)) * 1.4426950408889634);
  e.scale_cols = d;
  if ((rc = dense(mode, ws.ln_hi, ws.ln_lo, d, L.wqkv_hi, L.wqkv_lo, L.wqkv_f32, M, 3 * d, d, &e, st, &ws))) return rc;
  if ((rc = sdpa_run(ws.qkv_hi, lo ? ws.qkv_lo : nullptr, ws.ao_hi, lo ? ws.ao_lo : nullptr, nullptr, batch, seq, H, dh, 1, st)))
    return rc;
  const int pitch = seq * d;  // row 0 of clip b lives at b * seq * d
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = ws.xr;
  e.ldo = d;
  e.bias = L.bo;
  e.resid = x;
  e.ldr = pitch;
  const bool ride = mode != FACT_MODE_FP32_SIMT;
  if (ride) {
    e.ln_gamma = L.ln2_gamma;
    e.ln_beta = L.ln2_beta;
    e.ln_hi = ws.ln_hi;
    e.ln_lo = lo ? ws.ln_lo : nullptr;
  }
  if ((rc = dense(mode, ws.ao_hi, ws.ao_lo, pitch, L.wo_hi, L.wo_lo, L.wo_f32, batch, d, d, &e, st, &ws))) return rc;
  if (!ride &&
      (rc = fact_layernorm_split(ws.xr, L.ln2_gamma, L.ln2_beta, ws.ln_hi, lo ? ws.ln_lo : nullptr, batch, d, st)))
    return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_GELU_SPLIT;
  e.out_hi = ws.h_hi;
  e.out_lo = lo ? ws.h_lo : nullptr;
  e.ldo = ff;
  e.bias = L.b1;
  if ((rc = dense(mode, ws.ln_hi, ws.ln_lo, d, L.w1_hi, L.w1_lo, L.w1_f32, batch, ff, d, &e, st, &ws))) return rc;
  e = fact_gemm_epilogue{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = ws.xr;
  e.ldo = d;
  e.bias = L.b2;
  e.resid = ws.xr;
  e.ldr = d;
  return dense(mode, ws.h_hi, ws.h_lo, ff, L.w2_hi, L.w2_lo, L.w2_f32, batch, d, ff, &e, st, &ws);
}

static int run_stack(const fact_dims* dm, const fact_layer_weights* layers, int n_layers, float* x, int batch,
                     int seq, int mode, const Workspace& ws, float* dst, int dst_seq, int dst_off, cudaStream_t st) {
  const bool ride = mode != FACT_MODE_FP32_SIMT;
  for (int i = 0; i < n_layers; ++i) {
    const bool last = i + 1 == n_layers;
    // layer i's FF2 call leaves LayerNorm1 of layer i + 1 in ws.ln_hi / ln_lo
    int rc = run_layer(dm, layers[i], x, batch, seq, mode, ws, last ? dst : nullptr, dst_seq, dst_off, ride && i > 0,
                       last ? nullptr : &layers[i + 1], st);
    if (rc) return rc;
  }
  return FACT_OK;
}

int embed_prep(const float* x, long long batch_stride, const int* step_ptr, void* hi, void* lo, int rows, int n_tok,
               int f, int kp, cudaStream_t st);                                                       // elementwise.cu
int pack_weight_ld(const float* w, void* hi, void* lo, int k_in, int n_out, int ld, cudaStream_t st);  // elementwise.cu

// Once per fact_forward / fact_infer_auto_regressive call (outside the captured frame): K-major split copies of the two
// LinearEmbedding kernels for the tensor-core embedding.
static int pack_embed_weights(const fact_dims* dm, const fact_weights* w, const Workspace& ws, cudaStream_t st) {
  if (!ws.em_x_hi) return FACT_OK;
  int rc;
  if ((rc = pack_weight_ld(w->motion_embed_w, ws.em_w_hi, ws.em_w_lo, dm->motion_dim, dm->d_model, ws.em_kp, st)))
    return rc;
  return pack_weight_ld(w->audio_embed_w, ws.ea_w_hi, ws.ea_w_lo, dm->audio_dim, dm->d_model, ws.ea_kp, st);
}

// LinearEmbedding + bias + PositionEmbedding (base_models.py:130-156).  Large batches: inputs split to bf16 hi / lo
// and the product runs on the tensor path (the position table enters as a broadcast residual); otherwise one
// CUDA-core fp32 launch.
static int embed(const float* x, long long bs, const int* step_ptr, const float* w_f32, const float* bias,
                 const float* pos, float* y, int batch, int n_tok, int f, int d, int mode, bf16* x_hi, bf16* x_lo,
                 const bf16* w_hi, const bf16* w_lo, int kp, cudaStream_t st) {
  if (!x_hi) return fact_embed(x, bs, step_ptr, w_f32, bias, pos, y, batch, n_tok, f, d, st);
  const int rows = batch * n_tok;
  int rc;
  if ((rc = embed_prep(x, bs, step_ptr, x_hi, x_lo, rows, n_tok, f, kp, st))) return rc;
  fact_gemm_epilogue e{};
  e.kind = FACT_EPI_BIAS_RESID_F32;
  e.out_f32 = y;
  e.ldo = d;
  e.bias = bias;
  e.resid = pos;
  e.ldr = d;
  e.resid_rows = n_tok;
  return fact_gemm(x_hi, mode == FACT_MODE_PRECISE ? x_lo : nullptr, kp, w_hi, mode == FACT_MODE_PRECISE ? w_lo : nullptr,
                   kp, rows, d, f, &e, st);
}

// embeddings + modality encoders + 12-layer cross-modal stack; leaves the result in ws.xc
static int run_trunk(const fact_dims* dm, const fact_weights* w, const float* motion, long long motion_bs,
                     const float* audio, long long audio_bs, const int* step_ptr, int batch, int mode,
                     const Workspace& ws, bool row0_only, cudaStream_t st) {
  const int d = dm->d_model, ns = dm->motion_seq + dm->audio_seq;
  int rc;
  FACT_REQUIRE(dm->motion_layers > 0 && dm->audio_layers > 0 && dm->cross_layers > 0, FACT_ERR_UNSUPPORTED,
               "every stack needs at least one layer");
  // The two modality encoders are independent (fact_model.py:88-96).  At small batch each is a chain of tiny,
  // latency-bound launches, so the audio encoder runs on a second stream with its own scratch and joins before the
  // cross-modal stack (a fork / join inside the captured graph).  At large batch the GPU is full: one stream.
  const bool dual = g_dual_stream && ws.a_ln_hi != nullptr;
  cudaStream_t sa = st;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  Workspace wa = ws;
  if (dual) {
    // side stream + fork / join events of the calling thread, one set per device (a stream or event of another device
    // is an invalid resource handle on this one)
    struct Side {
      cudaStream_t s2 = nullptr;
      cudaEvent_t fork = nullptr, join = nullptr;
    };
    static thread_local std::map<int, Side> sides;
    int device = 0;
    FACT_CUDA_CHECK(cudaGetDevice(&device));
    Side& side = sides[device];
    if (!side.s2) {
      FACT_CUDA_CHECK(cudaStreamCreateWithFlags(&side.s2, cudaStreamNonBlocking));
      FACT_CUDA_CHECK(cudaEventCreateWithFlags(&side.fork, cudaEventDisableTiming));
      FACT_CUDA_CHECK(cudaEventCreateWithFlags(&side.join, cudaEventDisableTiming));
    }
    sa = side.s2;
    ev_fork = side.fork;
    ev_join = side.join;
    wa.ln_hi = ws.a_ln_hi; wa.ln_lo = ws.a_ln_lo; wa.ao_hi = ws.a_ao_hi; wa.ao_lo = ws.a_ao_lo;
    wa.qkv_hi = ws.a_qkv_hi; wa.qkv_lo = ws.a_qkv_lo; wa.h_hi = ws.a_h_hi; wa.h_lo = ws.a_h_lo;
    wa.splitk = ws.a_splitk; wa.splitk_bytes = ws.a_splitk_bytes;
    wa.ln_sync = ws.a_ln_sync;
    FACT_CUDA_CHECK(cudaEventRecord(ev_fork, st));
    FACT_CUDA_CHECK(cudaStreamWaitEvent(sa, ev_fork, 0));
    if ((rc = embed(audio, audio_bs, step_ptr, w->audio_embed_w, w->audio_embed_b, w->audio_pos, ws.xa, batch,
                    dm->audio_seq, dm->audio_dim, d, mode, ws.ea_x_hi, ws.ea_x_lo, ws.ea_w_hi, ws.ea_w_lo, ws.ea_kp, sa)))
      return rc;
    if ((rc = run_stack(dm, w->audio_layers, dm->audio_layers, ws.xa, batch, dm->audio_seq, mode, wa, ws.xc, ns,
                        dm->motion_seq, sa)))
      return rc;
    FACT_CUDA_CHECK(cudaEventRecord(ev_join, sa));
  }
  if ((rc = embed(motion, motion_bs, step_ptr, w->motion_embed_w, w->motion_embed_b, w->motion_pos, ws.xm, batch,
                  dm->motion_seq, dm->motion_dim, d, mode, ws.em_x_hi, ws.em_x_lo, ws.em_w_hi, ws.em_w_lo, ws.em_kp, st)))
    return rc;
  if ((rc = run_stack(dm, w->motion_layers, dm->motion_layers, ws.xm, batch, dm->motion_seq, mode, ws, ws.xc, ns, 0,
                      st)))
    return rc;
  if (dual) {
    FACT_CUDA_CHECK(cudaStreamWaitEvent(st, ev_join, 0));   // join before the cross-modal stack reads xc
  } else {
    if ((rc = embed(audio, audio_bs, step_ptr, w->audio_embed_w, w->audio_embed_b, w->audio_pos, ws.xa, batch,
                    dm->audio_seq, dm->audio_dim, d, mode, ws.ea_x_hi, ws.ea_x_lo, ws.ea_w_hi, ws.ea_w_lo, ws.ea_kp, st)))
      return rc;
    if ((rc = run_stack(dm, w->audio_layers, dm->audio_layers, ws.xa, batch, dm->audio_seq, mode, ws, ws.xc, ns,
                        dm->motion_seq, st)))
      return rc;
  }
  if (!row0_only)
    return run_stack(dm, w->cross_layers, dm->cross_layers, ws.xc, batch, ns, mode, ws, nullptr, 0, 0, st);
  if ((rc = run_stack(dm, w->cross_layers, dm->cross_layers - 1, ws.xc, batch, ns, mode, ws, nullptr, 0, 0, st)))
    return rc;
  return run_layer_row0(dm, w->cross_layers[dm->cross_layers - 1], ws.xc, batch, ns, mode, ws, st);
}

static int check_weights(const fact_dims* dm, const fact_weights* w) {
  FACT_REQUIRE(w && w->motion_layers && w->audio_layers && w->cross_layers, FACT_ERR_BAD_SHAPE, "null weights");
  FACT_REQUIRE(w->motion_embed_w && w->motion_embed_b && w->motion_pos && w->audio_embed_w && w->audio_embed_b &&
                   w->audio_pos && w->out_w && w->out_b,
               FACT_ERR_BAD_SHAPE, "null embedding / head weight");
  (void)dm;
  return FACT_OK;
}

// ---- graph cache for the AR loop
int g_ar_prune = 1;  // fact_set_flag("ar_prune", 0): run the full last layer (A-B check of the row-0 pruning)
int g_ar_fused = 1;  // reserved for the fused small-batch decode path (fact_set_flag("ar_fused", 0) disables it)
extern int g_gemm_pair, g_gemm_splitk, g_sdpa_legacy, g_gemm_tma_store, g_gemm_bn, g_gemm_finish_ln, g_sdpa_wide,
    g_gemm_fuse_ln;

// One captured frame is valid for exactly the pointers and sizes it was captured with, so the key is EVERYTHING the
// capture bakes in: the whole weight table (every pointer of every layer), the dims, every developer flag, the call's
// buffers and sizes.
struct GraphKey {
  std::vector<uintptr_t> v;
  bool operator==(const GraphKey& o) const { return v == o.v; }
  void add(uintptr_t x) { v.push_back(x); }
  void add_bytes(const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i += sizeof(uintptr_t)) {
      uintptr_t w = 0;
      memcpy(&w, b + i, n - i < sizeof(uintptr_t) ? n - i : sizeof(uintptr_t));
      v.push_back(w);
    }
  }
};

// Graph executables belong to a session (fact_ar_session_create): a model owns one and destroys it with itself, so a
// stale graph cannot outlive the buffers it points into.  Calls without a session share a process-wide default one.
// Bounded, least-recently-used eviction of ONE entry at a time; destroying an executable that is still running on a
// stream is legal (its resources are released when the launch completes).
struct ArSession {
  struct Entry {
    GraphKey key;
    cudaGraphExec_t exec;
    long long kernels;  // kernel nodes of one frame
    unsigned long long last_use;
  };
  std::mutex mu;
  std::vector<Entry> entries;
  unsigned long long tick = 0;
  static constexpr size_t kCapacity = 12;
  ~ArSession() {
    for (auto& e : entries) cudaGraphExecDestroy(e.exec);
  }
};
static ArSession g_default_session;

}  // namespace fact

using namespace fact;

extern "C" size_t fact_workspace_bytes(const fact_dims* dims, int batch, int mode) {
  if (!dims || batch <= 0) return 0;
  return carve(dims, batch, mode, nullptr, nullptr);
}

extern "C" int fact_forward(const fact_dims* dims, const fact_weights* w, const float* motion, const float* audio,
                            float* out, int batch, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_weights(dims, w))) return rc;
  FACT_REQUIRE(motion && audio && out && batch > 0, FACT_ERR_BAD_SHAPE, "fact_forward: bad arguments");
  FACT_REQUIRE(mode >= 0 && mode <= 2, FACT_ERR_UNSUPPORTED, "unknown mode %d", mode);
  FACT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, FACT_ERR_BAD_ALIGN,
               "workspace must be 1024-B aligned");
  Workspace ws;
  const size_t need = carve(dims, batch, mode, workspace, &ws);
  FACT_REQUIRE(workspace && workspace_bytes >= need, FACT_ERR_WORKSPACE, "workspace too small: %zu < %zu",
               workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  const long long mbs = static_cast<long long>(dims->motion_seq) * dims->motion_dim;
  const long long abs_ = static_cast<long long>(dims->audio_seq) * dims->audio_dim;
  if ((rc = pack_embed_weights(dims, w, ws, st))) return rc;
  if ((rc = zero_ln_sync(ws, st))) return rc;
  if ((rc = run_trunk(dims, w, motion, mbs, audio, abs_, nullptr, batch, mode, ws, false, st))) return rc;
  // output Dense on all 360 rows (base_models.py:200)
  const int d = dims->d_model, ns = dims->motion_seq + dims->audio_seq, M = batch * ns;
  const bool lo = mode != FACT_MODE_BF16;
  if ((rc = fact_layernorm_split(ws.xc, nullptr, nullptr, ws.ln_hi, lo ? ws.ln_lo : nullptr, M, d, st))) return rc;
  fact_gemm_epilogue e{};
  e.kind = FACT_EPI_BIAS_F32;
  e.out_f32 = out;
  e.ldo = dims->out_dim;
  e.bias = w->out_b;
  if (mode == FACT_MODE_FP32_SIMT)
    return gemm_simt_split(ws.ln_hi, ws.ln_lo, d, w->out_w, M, dims->out_dim, d, &e, st);
  FACT_REQUIRE(w->out_w_hi && (mode == FACT_MODE_BF16 || w->out_w_lo), FACT_ERR_BAD_SHAPE, "packed head weight missing");
  return fact_gemm(ws.ln_hi, lo ? ws.ln_lo : nullptr, d, w->out_w_hi, lo ? w->out_w_lo : nullptr, d, M,
                   dims->out_dim, d, &e, st);
}

extern "C" void* fact_ar_session_create(void) { return new (std::nothrow) ArSession(); }

extern "C" int fact_ar_session_destroy(void* session) {
  delete static_cast<ArSession*>(session);
  return FACT_OK;
}

extern "C" int fact_ar_session_graphs(void* session) {
  ArSession* s = session ? static_cast<ArSession*>(session) : &g_default_session;
  std::lock_guard<std::mutex> lk(s->mu);
  return static_cast<int>(s->entries.size());
}

extern "C" int fact_infer_auto_regressive(const fact_dims* dims, const fact_weights* w, float* motion_hist,
                                          int hist_capacity, const float* audio, int audio_len, int batch,
                                          int start_frame, int n_frames, int* step_counter, void* workspace,
                                          size_t workspace_bytes, int mode, int use_graph, void* session,
                                          void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_weights(dims, w))) return rc;
  FACT_REQUIRE(motion_hist && audio && step_counter && batch > 0 && n_frames > 0 && start_frame >= 0,
               FACT_ERR_BAD_SHAPE, "fact_infer_auto_regressive: bad arguments");
  FACT_REQUIRE(start_frame + n_frames <= hist_capacity, FACT_ERR_BAD_SHAPE,
               "frames [%d, %d) exceed the history capacity %d", start_frame, start_frame + n_frames, hist_capacity);
  FACT_REQUIRE(audio_len - dims->audio_seq + 1 >= start_frame + n_frames, FACT_ERR_BAD_SHAPE,
               "audio_len %d supports only %d frames, %d requested (fact_model.py:125-126 early stop is the caller's)",
               audio_len, audio_len - dims->audio_seq + 1, start_frame + n_frames);
  FACT_REQUIRE(mode >= 0 && mode <= 2, FACT_ERR_UNSUPPORTED, "unknown mode %d", mode);
  FACT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, FACT_ERR_BAD_ALIGN,
               "workspace must be 1024-B aligned");
  Workspace ws;
  const size_t need = carve(dims, batch, mode, workspace, &ws);
  FACT_REQUIRE(workspace && workspace_bytes >= need, FACT_ERR_WORKSPACE, "workspace too small: %zu < %zu",
               workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  const int ns = dims->motion_seq + dims->audio_seq;
  const long long hist_bs = static_cast<long long>(dims->motion_seq + hist_capacity) * dims->motion_dim;
  const long long audio_bs = static_cast<long long>(audio_len) * dims->audio_dim;
  FACT_REQUIRE(dims->out_dim == dims->motion_dim, FACT_ERR_BAD_SHAPE,
               "AR feedback needs out_dim == motion feature dim (fact_model.py:131)");
  {  // outside the per-frame graph: the embedding kernels may have changed since the last call
    int rc = pack_embed_weights(dims, w, ws, st);
    if (rc) return rc;
    if ((rc = zero_ln_sync(ws, st))) return rc;
  }

  auto one_frame = [&](cudaStream_t s) -> int {
    int r;
    if ((r = run_trunk(dims, w, motion_hist, hist_bs, audio, audio_bs, step_counter, batch, mode, ws, g_ar_prune != 0,
                       s)))
      return r;
    // keep row 0 of each sample (fact_model.py:128), append it to the history at row motion_seq + step
    if ((r = fact_head_rows(g_ar_prune ? ws.xr : ws.xc, g_ar_prune ? 1 : ns, w->out_w, w->out_b,
                            motion_hist + static_cast<size_t>(dims->motion_seq) * dims->motion_dim, hist_bs,
                            step_counter, batch, dims->d_model, dims->out_dim, s)))
      return r;
    return step_inc(step_counter, s);
  };

  if ((rc = step_set(step_counter, start_frame, st))) return rc;
  if (!use_graph) {
    for (int i = 0; i < n_frames; ++i)
      if ((rc = one_frame(st))) return rc;
    return FACT_OK;
  }
  FACT_REQUIRE(st != nullptr, FACT_ERR_UNSUPPORTED,
               "graph replay needs a non-default stream (capture is illegal on the legacy stream)");
  GraphKey key;
  key.add_bytes(dims, sizeof(*dims));
  key.add_bytes(w->motion_layers, sizeof(fact_layer_weights) * dims->motion_layers);
  key.add_bytes(w->audio_layers, sizeof(fact_layer_weights) * dims->audio_layers);
  key.add_bytes(w->cross_layers, sizeof(fact_layer_weights) * dims->cross_layers);
  {
    fact_weights flat = *w;  // the three host-array pointers are not device state: their contents were added above
    flat.motion_layers = flat.audio_layers = flat.cross_layers = nullptr;
    key.add_bytes(&flat, sizeof(flat));
  }
  for (uintptr_t x : {reinterpret_cast<uintptr_t>(motion_hist), reinterpret_cast<uintptr_t>(audio),
                      reinterpret_cast<uintptr_t>(step_counter), reinterpret_cast<uintptr_t>(workspace),
                      static_cast<uintptr_t>(audio_len), static_cast<uintptr_t>(batch),
                      static_cast<uintptr_t>(hist_capacity), static_cast<uintptr_t>(mode)})
    key.add(x);
  int device = 0;
  FACT_CUDA_CHECK(cudaGetDevice(&device));
  key.add(static_cast<uintptr_t>(device));
  for (int f : {g_ar_prune, g_gemm_pair, g_gemm_splitk, g_sdpa_legacy, g_dual_stream, g_gemm_tma_store, g_gemm_bn,
                g_gemm_finish_ln, g_ar_fused, g_sdpa_wide, g_gemm_fuse_ln, g_pdl})   // g_pdl stays LAST: the capture fallback below rewrites it
    key.add(static_cast<uintptr_t>(f));
  ArSession* sess = session ? static_cast<ArSession*>(session) : &g_default_session;
  cudaGraphExec_t exec = nullptr;
  long long frame_kernels = 0;
  {
    std::lock_guard<std::mutex> lk(sess->mu);
    for (auto& e : sess->entries)
      if (e.key == key) {
        exec = e.exec;
        frame_kernels = e.kernels;
        e.last_use = ++sess->tick;
        break;
      }
  }
  if (!exec) {
    // Capture one frame.  With programmatic dependent launches the graph carries programmatic edges; should this
    // driver refuse them at capture or instantiation, the frame is captured once more with plain launches (the flag is
    // left off for the rest of the process and the key of this entry says so).
    for (int attempt = 0; attempt < 2 && !exec; ++attempt) {
      cudaGraph_t graph = nullptr;
      const long long before = g_launch_count;
      FACT_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      rc = one_frame(st);
      frame_kernels = g_launch_count - before;
      g_launch_count = before;  // captured, not executed: replays are counted below
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc == FACT_OK && ce == cudaSuccess) {
        ce = cudaGraphInstantiate(&exec, graph, 0);
        if (ce != cudaSuccess) exec = nullptr;
      }
      if (graph) cudaGraphDestroy(graph);
      if (exec) break;
      if (attempt == 0 && g_pdl) {
        cudaGetLastError();  // clear the sticky capture error
        g_pdl = 0;
        key.v.back() = 0;    // g_pdl is the last key word
        continue;
      }
      if (rc) return rc;
      return cuda_fail(ce, "AR frame graph capture / instantiate");
    }
    std::lock_guard<std::mutex> lk(sess->mu);
    if (sess->entries.size() >= ArSession::kCapacity) {  // evict the least recently used entry only
      size_t victim = 0;
      for (size_t i = 1; i < sess->entries.size(); ++i)
        if (sess->entries[i].last_use < sess->entries[victim].last_use) victim = i;
      cudaGraphExecDestroy(sess->entries[victim].exec);
      sess->entries.erase(sess->entries.begin() + victim);
    }
    sess->entries.push_back(ArSession::Entry{key, exec, frame_kernels, ++sess->tick});
  }
  for (int i = 0; i < n_frames; ++i) FACT_CUDA_CHECK(cudaGraphLaunch(exec, st));
  g_launch_count += frame_kernels * n_frames;
  return FACT_OK;
}
