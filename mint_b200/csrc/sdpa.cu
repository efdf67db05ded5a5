// Attention core of the FACT transformer block (reference: mint/core/base_models.py:76-86):
//   q,k,v = split(qkv) "b n (qkv h d) -> qkv b h n d";  out = softmax(q.k^T * d_model^-0.5) . v ; merge heads.
//
// Flash-style, no score tensor in HBM: per CTA one (batch, head, 128-query block); 8 warps x 16 query rows;
// keys/values stream through a 2-stage cp.async ring in blocks of 64; S = Q.K^T and O += P.V run on the
// warp-level tensor path (mma.sync m16n8k16 bf16, fp32 accumulate) with an online softmax in registers
// (exp2 domain: q arrives pre-multiplied by scale*log2 e from the QKV GEMM epilogue).
// Precise mode (NPART == 2): q, k, v, p are hi+lo bf16 pairs and every product is hi.hi + lo.hi + hi.lo.
// Sequence lengths here are 120 / 240 / 360 (not multiples of 64): tail keys are masked to -inf, tail query
// rows are computed on zero inputs and never stored.
#include <string.h>

#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {

constexpr int SDPA_QB = 128;
constexpr int SDPA_KB = 64;
constexpr int SDPA_THREADS = 256;

template <int DH, int NPART>
__global__ void __launch_bounds__(SDPA_THREADS, (NPART == 1 ? 2 : 1))
sdpa_kernel(const bf16* __restrict__ qkv_hi, const bf16* __restrict__ qkv_lo, bf16* __restrict__ o_hi,
            bf16* __restrict__ o_lo, float* __restrict__ lse, int N, int H) {
  constexpr int DHP = DH + 8;            // row pitch: 16-B rotation per row -> conflict-free ldmatrix
  constexpr int KSTEPS = DH / 16;
  constexpr int DT = DH / 8;
  constexpr int CH = DH / 8;             // 16-byte chunks per row
  constexpr int TILE = SDPA_KB * DHP;    // elements of one K or V part tile
  constexpr int STAGE = 2 * NPART * TILE;  // [K parts][V parts]
  static_assert(DH % 16 == 0, "head_dim must be a multiple of 16");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* smem = reinterpret_cast<bf16*>(smem_raw);

  const int D = H * DH, LD = 3 * D;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t tok0 = static_cast<size_t>(b) * N;
  const bf16* src_part[2] = {qkv_hi, qkv_lo};

  // ---- stage Q (128 x DH per part) into the stage-1 region, K/V block 0 into stage 0
  {
    bf16* qs = smem + STAGE;
    for (int p = 0; p < NPART; ++p)
      for (int c = threadIdx.x; c < SDPA_QB * CH; c += SDPA_THREADS) {
        const int r = c / CH, cc = c % CH;
        const int row = qb * SDPA_QB + r;
        const bool ok = row < N;
        const bf16* src = src_part[p] + (tok0 + (ok ? row : 0)) * LD + h * DH + cc * 8;
        cp_async16(smem_u32(qs + p * (SDPA_QB * DHP) + r * DHP + cc * 8), src, ok);
      }
    cp_async_commit();
  }
  auto load_kv = [&](int kb, int stage) {
    bf16* base = smem + stage * STAGE;
    const int j0 = kb * SDPA_KB;
    for (int p = 0; p < NPART; ++p)
      for (int c = threadIdx.x; c < 2 * SDPA_KB * CH; c += SDPA_THREADS) {
        const int which = c / (SDPA_KB * CH);  // 0 = K, 1 = V
        const int rem = c % (SDPA_KB * CH);
        const int r = rem / CH, cc = rem % CH;
        const int row = j0 + r;
        const bool ok = row < N;
        const bf16* src = src_part[p] + (tok0 + (ok ? row : 0)) * LD + (1 + which) * D + h * DH + cc * 8;
        cp_async16(smem_u32(base + (which * NPART + p) * TILE + r * DHP + cc * 8), src, ok);
      }
  };
  load_kv(0, 0);
  cp_async_commit();

  cp_async_wait<1>();  // Q landed
  __syncthreads();
  uint32_t qf[NPART][KSTEPS][4];
  {
    const int mat = lane >> 3, r = lane & 7;
    const int row = warp * 16 + (mat & 1) * 8 + r;
#pragma unroll
    for (int p = 0; p < NPART; ++p)
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int col = ks * 16 + (mat >> 1) * 8;
        ldsm_x4(smem_u32(smem + STAGE + p * (SDPA_QB * DHP) + row * DHP + col), qf[p][ks][0], qf[p][ks][1],
                qf[p][ks][2], qf[p][ks][3]);
      }
  }
  __syncthreads();  // everyone holds its Q fragments: the stage-1 region is free for K/V block 1

  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  const int nkb = (N + SDPA_KB - 1) / SDPA_KB;
  for (int kb = 0; kb < nkb; ++kb) {
    if (kb + 1 < nkb) load_kv(kb + 1, (kb + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();  // block kb landed (the newest group may still be in flight)
    __syncthreads();
    const bf16* Ks = smem + (kb & 1) * STAGE;
    const bf16* Vs = Ks + NPART * TILE;

    // ---- S = Q.K^T  (16 x 64 per warp)
    float s[SDPA_KB / 8][4];
#pragma unroll
    for (int i = 0; i < SDPA_KB / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    {
      // products are issued tile-major (8 independent accumulators between two updates of the same tile), so the
      // hi.hi / lo.hi / hi.lo passes never wait on each other's HMMA latency
      const int mat = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        uint32_t kh[SDPA_KB / 16][4], kl[SDPA_KB / 16][4];
        const int col = ks * 16 + (mat & 1) * 8;
#pragma unroll
        for (int ntp = 0; ntp < SDPA_KB / 16; ++ntp) {
          const int krow = ntp * 16 + (mat >> 1) * 8 + r;
          ldsm_x4(smem_u32(Ks + krow * DHP + col), kh[ntp][0], kh[ntp][1], kh[ntp][2], kh[ntp][3]);
          if (NPART == 2)
            ldsm_x4(smem_u32(Ks + TILE + krow * DHP + col), kl[ntp][0], kl[ntp][1], kl[ntp][2], kl[ntp][3]);
        }
#pragma unroll
        for (int ntp = 0; ntp < SDPA_KB / 16; ++ntp) {
          mma_bf16_16816(s[2 * ntp], qf[0][ks], kh[ntp][0], kh[ntp][1]);
          mma_bf16_16816(s[2 * ntp + 1], qf[0][ks], kh[ntp][2], kh[ntp][3]);
        }
        if (NPART == 2) {
#pragma unroll
          for (int ntp = 0; ntp < SDPA_KB / 16; ++ntp) {  // q_lo . k_hi
            mma_bf16_16816(s[2 * ntp], qf[NPART - 1][ks], kh[ntp][0], kh[ntp][1]);
            mma_bf16_16816(s[2 * ntp + 1], qf[NPART - 1][ks], kh[ntp][2], kh[ntp][3]);
          }
#pragma unroll
          for (int ntp = 0; ntp < SDPA_KB / 16; ++ntp) {  // q_hi . k_lo
            mma_bf16_16816(s[2 * ntp], qf[0][ks], kl[ntp][0], kl[ntp][1]);
            mma_bf16_16816(s[2 * ntp + 1], qf[0][ks], kl[ntp][2], kl[ntp][3]);
          }
        }
      }
    }
    // ---- mask the key tail, online softmax (rows g and g+8 of this warp's 16)
    const int j0 = kb * SDPA_KB;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < SDPA_KB / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = j0 + nt * 8 + 2 * t + (e & 1);
        if (col >= N) s[nt][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    float alpha[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
      const float mnew = fmaxf(m_run[i], mx[i]);  // finite: block 0 always has a valid key
      alpha[i] = exp2f(m_run[i] - mnew);
      m_run[i] = mnew;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < SDPA_KB / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[nt][e] - m_run[e >> 1]);
        s[nt][e] = p;
        psum[e >> 1] += p;
      }
    l_run[0] = l_run[0] * alpha[0] + psum[0];
    l_run[1] = l_run[1] * alpha[1] + psum[1];
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      o[dt][0] *= alpha[0];
      o[dt][1] *= alpha[0];
      o[dt][2] *= alpha[1];
      o[dt][3] *= alpha[1];
    }
    // ---- O += P.V
    {
      const int mat = lane >> 3, r = lane & 7;
#pragma unroll
      for (int kk = 0; kk < SDPA_KB / 16; ++kk) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // A fragment regs: {tile 2kk: c0c1, c2c3, tile 2kk+1: c0c1, c2c3}
          const float p0 = s[2 * kk + (i >> 1)][(i & 1) * 2], p1 = s[2 * kk + (i >> 1)][(i & 1) * 2 + 1];
          bf16 h0, l0, h1, l1;
          split_bf16(p0, h0, l0);
          split_bf16(p1, h1, l1);
          ah[i] = pack_bf16x2(h0, h1);
          al[i] = pack_bf16x2(l0, l1);
        }
        const int vrow = kk * 16 + (mat & 1) * 8 + r;
        uint32_t vh[DT / 2][4];
#pragma unroll
        for (int dtp = 0; dtp < DT / 2; ++dtp) {
          const int col = (2 * dtp + (mat >> 1)) * 8;
          ldsm_x4_t(smem_u32(Vs + vrow * DHP + col), vh[dtp][0], vh[dtp][1], vh[dtp][2], vh[dtp][3]);
          mma_bf16_16816(o[2 * dtp], ah, vh[dtp][0], vh[dtp][1]);
          mma_bf16_16816(o[2 * dtp + 1], ah, vh[dtp][2], vh[dtp][3]);
        }
        if (NPART == 2) {
#pragma unroll
          for (int dtp = 0; dtp < DT / 2; ++dtp) {  // p_lo . v_hi
            mma_bf16_16816(o[2 * dtp], al, vh[dtp][0], vh[dtp][1]);
            mma_bf16_16816(o[2 * dtp + 1], al, vh[dtp][2], vh[dtp][3]);
          }
#pragma unroll
          for (int dtp = 0; dtp < DT / 2; ++dtp) {  // p_hi . v_lo
            const int col = (2 * dtp + (mat >> 1)) * 8;
            uint32_t b00, b01, b10, b11;
            ldsm_x4_t(smem_u32(Vs + TILE + vrow * DHP + col), b00, b01, b10, b11);
            mma_bf16_16816(o[2 * dtp], ah, b00, b01);
            mma_bf16_16816(o[2 * dtp + 1], ah, b10, b11);
          }
        }
      }
    }
    __syncthreads();  // this stage is overwritten by the prefetch of the next iteration
  }

  // ---- normalise and store "b h n d -> b n (h d)"
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
  }
  const float inv[2] = {1.f / l_run[0], 1.f / l_run[1]};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = qb * SDPA_QB + warp * 16 + g + i * 8;
    if (row < N) {
      if (lse && t == 0) lse[(static_cast<size_t>(b) * H + h) * N + row] = m_run[i] + log2f(l_run[i]);
      const size_t base = (tok0 + row) * D + h * DH + 2 * t;
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        bf16 h0, l0, h1, l1;
        split_bf16(o[dt][2 * i] * inv[i], h0, l0);
        split_bf16(o[dt][2 * i + 1] * inv[i], h1, l1);
        *reinterpret_cast<uint32_t*>(o_hi + base + dt * 8) = pack_bf16x2(h0, h1);
        if (NPART == 2) *reinterpret_cast<uint32_t*>(o_lo + base + dt * 8) = pack_bf16x2(l0, l1);
      }
    }
  }
}

template <int DH, int NPART>
static int launch_sdpa(const bf16* qh, const bf16* ql, bf16* oh, bf16* ol, float* lse, int batch, int n, int heads,
                       int q_rows, cudaStream_t st) {
  constexpr int smem = 2 * (2 * NPART * SDPA_KB * (DH + 8)) * 2;
  auto kern = sdpa_kernel<DH, NPART>;
  static bool attr_done = false;
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  dim3 grid((q_rows + SDPA_QB - 1) / SDPA_QB, heads, batch);
  kern<<<grid, SDPA_THREADS, smem, st>>>(qh, ql, oh, ol, lse, n, heads);
  FACT_LAUNCH_CHECK("sdpa_kernel launch");
  return FACT_OK;
}

int sdpa_tc_try(const bf16* qh, const bf16* ql, bf16* oh, bf16* ol, float* lse, int batch, int n, int heads,
                int head_dim, int q_rows, cudaStream_t st);
extern int g_gemm_pair;
extern int g_gemm_splitk;
extern int g_gemm_bn;
extern int g_gemm_tma_store;
extern int g_wgrad_pair;
extern int g_gemm_finish_ln;
extern int g_sdpa_bwd_tc;
extern int g_dual_stream;
extern int g_ar_prune;
extern int g_ar_fused;
extern int g_sdpa_wide;
extern int g_wgrad_order;
extern int g_gemm_fuse_ln;
int g_sdpa_legacy = 0;  // fact_set_flag("sdpa_legacy", 1): force the mma.sync kernel (tests / A-B timing)

}  // namespace fact

using namespace fact;

// Developer switches: process-global on purpose (A/B timing of kernel variants, tests).  Plain ints written here and
// read at dispatch time: set them while no other thread is inside the library.  Every one is part of the AR graph key.
extern "C" int fact_set_flag(const char* name, int value) {
  struct Flag {
    const char* name;
    int* slot;
  };
  const Flag flags[] = {{"sdpa_legacy", &g_sdpa_legacy},   {"ar_prune", &g_ar_prune},
                        {"dual_stream", &g_dual_stream},   {"gemm_bn", &g_gemm_bn},
                        {"sdpa_bwd_tc", &g_sdpa_bwd_tc},   {"gemm_finish_ln", &g_gemm_finish_ln},
                        {"wgrad_pair", &g_wgrad_pair},     {"gemm_tma_store", &g_gemm_tma_store},
                        {"gemm_splitk", &g_gemm_splitk},   {"gemm_pair", &g_gemm_pair},
                        {"ar_fused", &g_ar_fused},         {"pdl", &g_pdl},
                        {"sdpa_wide", &g_sdpa_wide},       {"gemm_fuse_ln", &g_gemm_fuse_ln},
                        {"wgrad_order", &g_wgrad_order}};
  for (const Flag& f : flags)
    if (name && strcmp(name, f.name) == 0) {
      *f.slot = value;
      return FACT_OK;
    }
  set_error("fact_set_flag: unknown flag %s", name ? name : "(null)");
  return FACT_ERR_UNSUPPORTED;
}

namespace fact {

// q_rows < n: only the first q_rows query rows of every (batch, head) are needed (whole 128-row blocks are computed)
int sdpa_run(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo, float* lse, int batch, int n,
             int heads, int head_dim, int q_rows, cudaStream_t st) {
  FACT_REQUIRE(qkv_hi && out_hi, FACT_ERR_BAD_SHAPE, "fact_sdpa: null buffer");
  FACT_REQUIRE((qkv_lo == nullptr) == (out_lo == nullptr), FACT_ERR_BAD_SHAPE,
               "fact_sdpa: qkv_lo and out_lo must both be given (precise) or both NULL (bf16)");
  FACT_REQUIRE(batch > 0 && batch <= 65535 && n > 0 && heads > 0 && heads <= 65535 && q_rows > 0 && q_rows <= n,
               FACT_ERR_BAD_SHAPE, "fact_sdpa: bad shape batch=%d n=%d heads=%d q_rows=%d", batch, n, heads, q_rows);
  const bf16* qh = static_cast<const bf16*>(qkv_hi);
  const bf16* ql = static_cast<const bf16*>(qkv_lo);
  bf16* oh = static_cast<bf16*>(out_hi);
  bf16* ol = static_cast<bf16*>(out_lo);
  const bool precise = qkv_lo != nullptr;
  if (!g_sdpa_legacy) {  // B200-native path: tcgen05 + TMEM-resident scores (head_dim 80, n <= 384)
    const int rc = sdpa_tc_try(qh, ql, oh, ol, lse, batch, n, heads, head_dim, q_rows, st);
    if (rc != FACT_ERR_UNSUPPORTED) return rc;
  }
#define FACT_SDPA_CASE(DHV)                                                            \
  case DHV:                                                                            \
    return precise ? launch_sdpa<DHV, 2>(qh, ql, oh, ol, lse, batch, n, heads, q_rows, st)  \
                   : launch_sdpa<DHV, 1>(qh, ql, oh, ol, lse, batch, n, heads, q_rows, st);
  switch (head_dim) {
    FACT_SDPA_CASE(16)
    FACT_SDPA_CASE(32)
    FACT_SDPA_CASE(64)
    FACT_SDPA_CASE(80)
  }
#undef FACT_SDPA_CASE
  set_error("fact_sdpa: head_dim %d not instantiated (16, 32, 64, 80)", head_dim);
  return FACT_ERR_UNSUPPORTED;
}

}  // namespace fact

extern "C" int fact_sdpa(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo, int batch, int n,
                         int heads, int head_dim, void* stream) {
  return fact::sdpa_run(qkv_hi, qkv_lo, out_hi, out_lo, nullptr, batch, n, heads, head_dim, n,
                        fact::as_stream(stream));
}

extern "C" int fact_sdpa_lse(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo, float* lse, int batch,
                             int n, int heads, int head_dim, void* stream) {
  return fact::sdpa_run(qkv_hi, qkv_lo, out_hi, out_lo, lse, batch, n, heads, head_dim, n, fact::as_stream(stream));
}
