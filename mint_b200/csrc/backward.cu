// Backward kernels of the FACT training step (reference semantics: mint/ctl/single_task_trainer.py:138-199 --
// tape.gradient of FACTModel.loss through FACTModel.call; Keras Adam at trainer.py:150).  Training runs the bf16
// product path (BASELINE.json configs[2] "training step bf16"): bf16 operands, fp32 accumulation, fp32 master weights.
//
//   gemm_wgrad_kernel   dW[in,out] += X^T . dY   tcgen05, BOTH operands MN-major through TMA (no transposes in HBM),
//                       split-K over tokens, fp32 red.add epilogue straight into the Keras-layout gradient
//   (dX = dY . W^T reuses gemm_tc_kernel with the Keras-layout bf16 weight as the K-major B operand)
//   sdpa_bwd_*          attention-core backward, flash style (mma.sync), recomputing P from the saved log-sum-exp
//   ln_bwd_kernel       LayerNorm backward + residual-gradient add, per-block partial dgamma / dbeta
//   cast_colsum_kernel  fp32 -> bf16 operand cast fused with the bias gradient (column sums)
//   embed_bwd_kernel    LinearEmbedding / PositionEmbedding gradients
//   adam_kernel         Keras Adam (epsilon OUTSIDE the square root, bias correction folded into the step size)
#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {

// =========================================================================================== weight-gradient GEMM
constexpr int WG_BM = 128, WG_BN = 128, WG_BK = 64, WG_STAGES = 6, WG_THREADS = 192;
constexpr int WG_BOX_BYTES = 64 * 64 * 2;                 // one {64 MN, 64 K} box
constexpr int WG_STAGE_BYTES = 4 * WG_BOX_BYTES;          // A: 2 boxes, B: 2 boxes
constexpr int WG_SMEM_BYTES = 1024 + WG_STAGES * WG_STAGE_BYTES + 256;

// C[in, out] (+)= sum_t X[t, in] * dY[t, out].  tmX: tensor {in (inner), tokens}, tmY: tensor {out (inner), tokens}.
// grid = (tiles_out, tiles_in, splits); CTA handles K blocks [z * kb_per, min((z+1) * kb_per, num_kb)).
__global__ void __launch_bounds__(WG_THREADS, 1) gemm_wgrad_kernel(
    const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, float* __restrict__ dW, int ldw,
    int IN, int OUT, int num_kb, int kb_per) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + WG_STAGES * WG_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * WG_STAGES);
  const uint32_t tmem_ptr_addr = done_bar + 8;
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + WG_STAGES * WG_STAGE_BYTES + 8 * (2 * WG_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * WG_BN, m0 = blockIdx.y * WG_BM;
  const int kb0 = blockIdx.z * kb_per;
  const int kb1 = min(kb0 + kb_per, num_kb);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  auto sA = [&](int s) { return smem_base + s * WG_STAGE_BYTES; };
  auto sB = [&](int s) { return smem_base + s * WG_STAGE_BYTES + 2 * WG_BOX_BYTES; };

  if (nkb > 0) {
    if (warp == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % WG_STAGES;
        mbar_wait(empty_bar(s), ((i / WG_STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), WG_STAGE_BYTES);
          const int t0 = (kb0 + i) * WG_BK;
          tma_load_2d(sA(s), &tmX, m0, t0, full_bar(s));
          tma_load_2d(sA(s) + WG_BOX_BYTES, &tmX, m0 + 64, t0, full_bar(s));
          tma_load_2d(sB(s), &tmY, n0, t0, full_bar(s));
          tma_load_2d(sB(s) + WG_BOX_BYTES, &tmY, n0 + 64, t0, full_bar(s));
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32_maj(WG_BM, WG_BN, 1, 1);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % WG_STAGES;
        mbar_wait(full_bar(s), (i / WG_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a0 = umma_desc_mn_sw128(sA(s), WG_BOX_BYTES, 1024);
          const uint64_t b0 = umma_desc_mn_sw128(sB(s), WG_BOX_BYTES, 1024);
#pragma unroll
          for (int kk = 0; kk < WG_BK / 16; ++kk) {  // 16 tokens = two 8-row groups = 2048 bytes
            const uint64_t off = static_cast<uint64_t>(kk * 2048) >> 4;
            umma_bf16(tmem_base, a0 + off, b0 + off, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          if (i == nkb - 1) umma_commit(done_bar);
        }
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < WG_BN; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (row < IN) {
          float* dst = dW + static_cast<size_t>(row) * ldw + n0 + c0;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (n0 + c0 + c < OUT) atomicAdd(dst + c, v[c]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// =========================================================================================== attention backward
// D[b,h,row] = sum_d dO * O  (softmax-backward row term).  One warp per token: 16-byte loads across the whole
// [H * DH] row (DH % 8 == 0, so a chunk never straddles heads), per-chunk partial dots reduced per head through
// shared memory.
__global__ void __launch_bounds__(256) sdpa_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                            float* __restrict__ Dv, int tokens, int N, int H, int DH) {
  __shared__ float part[8][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 8 + warp;
  if (tok >= tokens) return;
  const int chunks = H * DH / 8, cph = DH / 8;  // chunks <= 128 (d_model <= 1024)
  const uint4* po = reinterpret_cast<const uint4*>(o + static_cast<size_t>(tok) * H * DH);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + static_cast<size_t>(tok) * H * DH);
  for (int c = lane; c < chunks; c += 32) {
    const uint4 a = po[c], b = pd[c];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(bw[i] << 16), s);
      s = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(bw[i] & 0xffff0000u), s);
    }
    part[warp][c] = s;
  }
  __syncwarp();
  const int b = tok / N, r = tok % N;
  for (int h = lane; h < H; h += 32) {
    float s = 0.f;
    for (int i = 0; i < cph; ++i) s += part[warp][h * cph + i];
    Dv[(static_cast<size_t>(b) * H + h) * N + r] = s;
  }
}

// One CTA per (batch, head, 64-key block); 4 warps x 16 keys.  For every 64-query block: recompute P^T from the saved
// log2-sum-exp, accumulate dV, dK in registers; dQ contributions go through a shared transposed tile and fp32 atomics.
// qkv holds q' = q * scale * log2(e) (as written by the QKV epilogue), k, v.  Outputs: dqkv[:, d:2d] = dK, [:, 2d:3d] = dV
// (bf16), dq_acc (fp32, += dZ . K, unscaled).
template <int DH>
__global__ void __launch_bounds__(128, 1) sdpa_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                          const float* __restrict__ lse, const float* __restrict__ Dv,
                                                          bf16* __restrict__ dqkv, float* __restrict__ dq_acc, int N,
                                                          int H, float k_scale) {
  constexpr int DHP = DH + 8, KS = DH / 16, DT = DH / 8, CH = DH / 8, BLK = 64, ZP = BLK + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
  bf16* Vs = Ks + BLK * DHP;
  bf16* Qs = Vs + BLK * DHP;          // 2 stages
  bf16* Os = Qs + 2 * BLK * DHP;      // dO, 2 stages
  bf16* Zs = Os + 2 * BLK * DHP;      // dZ^T [key][query]
  float* Ls = reinterpret_cast<float*>(Zs + BLK * ZP);  // lse, 2 stages
  float* Ds = Ls + 2 * BLK;                              // D, 2 stages

  const int jb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int D = H * DH, LD = 3 * D;
  const size_t tok0 = static_cast<size_t>(b) * N;
  const int j0 = jb * BLK;

  for (int c = threadIdx.x; c < 2 * BLK * CH; c += 128) {
    const int which = c / (BLK * CH), rem = c % (BLK * CH), r = rem / CH, cc = rem % CH;
    const bool ok = j0 + r < N;
    const bf16* src = qkv + (tok0 + (ok ? j0 + r : 0)) * LD + (1 + which) * D + h * DH + cc * 8;
    cp_async16(smem_u32((which ? Vs : Ks) + r * DHP + cc * 8), src, ok);
  }
  auto load_q = [&](int ib, int st) {
    const int i0 = ib * BLK;
    for (int c = threadIdx.x; c < 2 * BLK * CH; c += 128) {
      const int which = c / (BLK * CH), rem = c % (BLK * CH), r = rem / CH, cc = rem % CH;
      const bool ok = i0 + r < N;
      const size_t tok = tok0 + (ok ? i0 + r : 0);
      const bf16* src = which ? d_o + tok * D + h * DH + cc * 8 : qkv + tok * LD + h * DH + cc * 8;
      cp_async16(smem_u32((which ? Os : Qs) + st * BLK * DHP + r * DHP + cc * 8), src, ok);
    }
    if (threadIdx.x < BLK) {
      const int r = i0 + threadIdx.x;
      const size_t idx = (static_cast<size_t>(b) * H + h) * N + (r < N ? r : 0);
      Ls[st * BLK + threadIdx.x] = r < N ? lse[idx] : 0.f;
      Ds[st * BLK + threadIdx.x] = r < N ? Dv[idx] : 0.f;
    }
  };
  load_q(0, 0);
  cp_async_commit();

  float dk[DT][4], dv[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;

  const int nqb = (N + BLK - 1) / BLK;
  const int mat = lane >> 3, r8 = lane & 7;
  for (int ib = 0; ib < nqb; ++ib) {
    const int st = ib & 1;
    if (ib + 1 < nqb) load_q(ib + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* Q = Qs + st * BLK * DHP;
    const bf16* dO = Os + st * BLK * DHP;
    const float* L = Ls + st * BLK;
    const float* Dd = Ds + st * BLK;
    const int i0 = ib * BLK;

    // (1) S^T = K_w . Q^T   and (4) dP^T = V_w . dO^T      [16 keys x 64 queries each]
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[i][e] = dp[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ak[4], av[4];
      const int arow = warp * 16 + (mat & 1) * 8 + r8, acol = ks * 16 + (mat >> 1) * 8;
      ldsm_x4(smem_u32(Ks + arow * DHP + acol), ak[0], ak[1], ak[2], ak[3]);
      ldsm_x4(smem_u32(Vs + arow * DHP + acol), av[0], av[1], av[2], av[3]);
#pragma unroll
      for (int ntp = 0; ntp < 4; ++ntp) {
        const int brow = ntp * 16 + (mat >> 1) * 8 + r8, bcol = ks * 16 + (mat & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(Q + brow * DHP + bcol), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * ntp], ak, b0, b1);
        mma_bf16_16816(s[2 * ntp + 1], ak, b2, b3);
        ldsm_x4(smem_u32(dO + brow * DHP + bcol), b0, b1, b2, b3);
        mma_bf16_16816(dp[2 * ntp], av, b0, b1);
        mma_bf16_16816(dp[2 * ntp + 1], av, b2, b3);
      }
    }
    // (2) P^T = 2^(S^T - lse[query]) ; (5) dZ^T = P^T * (dP^T - D[query]) ; masked outside [0, N)
    uint32_t pa[4][4], za[4][4];  // A fragments (16 keys x 16 queries per k-step) of P^T and dZ^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pz[4], zz[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = nt * 8 + 2 * t + (e & 1);
        const int key = j0 + warp * 16 + g + (e >> 1) * 8;
        const bool ok = (i0 + qi < N) && (key < N);
        const float p = ok ? ex2_approx(s[nt][e] - L[qi]) : 0.f;
        pz[e] = p;
        zz[e] = p * (dp[nt][e] - Dd[qi]);
      }
      const int kk = nt >> 1, hi = nt & 1;
      pa[kk][hi * 2 + 0] = cvt_bf16x2(pz[0], pz[1]);
      pa[kk][hi * 2 + 1] = cvt_bf16x2(pz[2], pz[3]);
      za[kk][hi * 2 + 0] = cvt_bf16x2(zz[0], zz[1]);
      za[kk][hi * 2 + 1] = cvt_bf16x2(zz[2], zz[3]);
      // transposed tile for the dQ product: Zs[key][query]
      const int kr = warp * 16 + g, qc = nt * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(Zs + kr * ZP + qc) = za[kk][hi * 2 + 0];
      *reinterpret_cast<uint32_t*>(Zs + (kr + 8) * ZP + qc) = za[kk][hi * 2 + 1];
    }
    // (3) dV_w += P^T . dO     (6) dK_w += dZ^T . Q        [B operand: rows = query, n = dh -> transposed ldmatrix]
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int vrow = kk * 16 + (mat & 1) * 8 + r8;
#pragma unroll
      for (int dtp = 0; dtp < DT / 2; ++dtp) {
        const int col = (2 * dtp + (mat >> 1)) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(dO + vrow * DHP + col), b0, b1, b2, b3);
        mma_bf16_16816(dv[2 * dtp], pa[kk], b0, b1);
        mma_bf16_16816(dv[2 * dtp + 1], pa[kk], b2, b3);
        ldsm_x4_t(smem_u32(Q + vrow * DHP + col), b0, b1, b2, b3);
        mma_bf16_16816(dk[2 * dtp], za[kk], b0, b1);
        mma_bf16_16816(dk[2 * dtp + 1], za[kk], b2, b3);
      }
    }
    __syncthreads();  // Zs complete
    // (7) dQ[16 queries of this warp x DH] = dZ[query, key] . K[key, dh]
    {
      float dq[DT][4];
#pragma unroll
      for (int i = 0; i < DT; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
        // A[m = query][k = key] read transposed from Zs[key][query]
        uint32_t a[4];
        {
          const int krow = kk * 16 + (mat >> 1) * 8 + r8;       // matrices 0,1: keys 0-7 ; 2,3: keys 8-15
          const int qcol = warp * 16 + (mat & 1) * 8;           // matrices 0,2: queries 0-7 ; 1,3: queries 8-15
          ldsm_x4_t(smem_u32(Zs + krow * ZP + qcol), a[0], a[1], a[2], a[3]);
        }
        const int vrow = kk * 16 + (mat & 1) * 8 + r8;
#pragma unroll
        for (int dtp = 0; dtp < DT / 2; ++dtp) {
          const int col = (2 * dtp + (mat >> 1)) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(Ks + vrow * DHP + col), b0, b1, b2, b3);
          mma_bf16_16816(dq[2 * dtp], a, b0, b1);
          mma_bf16_16816(dq[2 * dtp + 1], a, b2, b3);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int qrow = i0 + warp * 16 + g + i * 8;
        if (qrow < N) {
          float* dst = dq_acc + (tok0 + qrow) * D + h * DH + 2 * t;
#pragma unroll
          for (int dt = 0; dt < DT; ++dt) {
            atomicAdd(dst + dt * 8, dq[dt][2 * i]);
            atomicAdd(dst + dt * 8 + 1, dq[dt][2 * i + 1]);
          }
        }
      }
    }
    __syncthreads();  // Q/dO stage and Zs are reused by the next iteration
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int key = j0 + warp * 16 + g + i * 8;
    if (key < N) {
      bf16* pk = dqkv + (tok0 + key) * LD + D + h * DH + 2 * t;
      bf16* pv = pk + D;
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        *reinterpret_cast<uint32_t*>(pk + dt * 8) = cvt_bf16x2(dk[dt][2 * i] * k_scale, dk[dt][2 * i + 1] * k_scale);
        *reinterpret_cast<uint32_t*>(pv + dt * 8) = cvt_bf16x2(dv[dt][2 * i], dv[dt][2 * i + 1]);
      }
    }
  }
}

// dqkv[:, 0:d] = bf16(dq_acc * scale)
__global__ void __launch_bounds__(256) sdpa_bwd_finish_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv,
                                                              long long tokens, int D, float scale) {
  const long long total = tokens * (D / 4);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long tok = i / (D / 4);
    const int c = static_cast<int>(i % (D / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(dq_acc + tok * D + c);
    uint2 o;
    o.x = cvt_bf16x2(v.x * scale, v.y * scale);
    o.y = cvt_bf16x2(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(dqkv + tok * 3 * D + c) = o;
  }
}

// =========================================================================================== LayerNorm backward
// One persistent launch (a block per SM).  Rows are processed in 8-row slabs; a slab's x and dy rows (contiguous in
// memory, like dres) arrive in shared memory through bulk async copies into a two-deep ring, so HBM sees x, dy and
// dres exactly once, the next slab is in flight while the current one is processed, and the compute phases never wait
// on a global load (ncu on the non-pipelined version: 20 % of the warp slots active, 50 % of the HBM peak, all stalls
// on loads; with dres still read directly in phase 2 the slab time was two global round trips).
//  (1) one warp per row: stats[row] = {mean, rstd, mean(g*dy), rstd * mean(g*dy*xhat)}
//  (2) thread = one float4 column group: dx = dres + rstd*(g*dy - m1 - xhat*m2); dgamma / dbeta (and, optionally, the
//      column sums of dx = the bias gradient of the dense layer feeding this residual branch) stay in registers over
//      ALL the block's slabs and leave through one atomicAdd per column per block; dx is optionally also written as
//      the bf16 GEMM operand.
constexpr int LN_SLAB = 8;
__global__ void __launch_bounds__(256, 1) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ dy, const float* dres, float* dx,
                                                        bf16* __restrict__ dx_b, float* __restrict__ dgamma,
                                                        float* __restrict__ dbeta, float* __restrict__ colsum,
                                                        int rows, int d, int stages) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int slab_floats = LN_SLAB * d;                               // one array of one slab
  float* ring = reinterpret_cast<float*>(ln_smem);                   // [stages][3][LN_SLAB][d]: x, dy, dres
  float4* sstats = reinterpret_cast<float4*>(ring + static_cast<size_t>(stages) * 3 * slab_floats);
  const uint32_t bar0 = smem_u32(sstats + LN_SLAB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = d >> 2;
  const int num_slabs = (rows + LN_SLAB - 1) / LN_SLAB;
  const int my_slabs = blockIdx.x < num_slabs ? (num_slabs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int j) {  // thread 0: slab j of this block -> ring slot j % stages
    const int r0 = (blockIdx.x + j * gridDim.x) * LN_SLAB;
    const uint32_t bytes = static_cast<uint32_t>(min(rows - r0, LN_SLAB)) * d * 4;
    const int slot = j % stages;
    float* dst = ring + static_cast<size_t>(slot) * 3 * slab_floats;
    mbar_arrive_expect_tx(bar0 + 8u * slot, (dres ? 3 : 2) * bytes);
    bulk_load_1d(smem_u32(dst), x + static_cast<size_t>(r0) * d, bytes, bar0 + 8u * slot);
    bulk_load_1d(smem_u32(dst + slab_floats), dy + static_cast<size_t>(r0) * d, bytes, bar0 + 8u * slot);
    if (dres) bulk_load_1d(smem_u32(dst + 2 * slab_floats), dres + static_cast<size_t>(r0) * d, bytes, bar0 + 8u * slot);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_barrier_init();
    for (int j = 0; j < stages - 1 && j < my_slabs; ++j) issue(j);
  }
  __syncthreads();
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const int cg = threadIdx.x;  // column group of phase 2 (d <= 1024 -> nv <= 256)
  const float4 gg = cg < nv ? __ldg(g4 + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, ac = ag;
  for (int j = 0; j < my_slabs; ++j) {
    // slot (j - 1) % stages was released by the __syncthreads that ended the previous iteration
    if (threadIdx.x == 0 && j + stages - 1 < my_slabs) issue(j + stages - 1);
    const int slot = j % stages;
    const int r0 = (blockIdx.x + j * gridDim.x) * LN_SLAB;
    const int nr = min(rows - r0, LN_SLAB);
    const float* xs = ring + static_cast<size_t>(slot) * 3 * slab_floats;
    const float* ys = xs + slab_floats;
    const float* rs = ys + slab_floats;
    mbar_wait(bar0 + 8u * slot, (j / stages) & 1);
    // ---- phase 1: warp w owns row w of the slab
    if (warp < nr) {
      const float4* xr = reinterpret_cast<const float4*>(xs + warp * d);
      const float4* yr = reinterpret_cast<const float4*>(ys + warp * d);
      float4 v[8];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = lane + i * 32;
        v[i] = idx < nv ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / d;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (lane + i * 32 < nv) {
          v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
          q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q / d + 1e-5f);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = lane + i * 32;
        if (idx < nv) {
          const float4 gl = __ldg(g4 + idx);
          const float4 u = yr[idx];
          const float a0 = gl.x * u.x, a1 = gl.y * u.y, a2 = gl.z * u.z, a3 = gl.w * u.w;
          s1 += (a0 + a1) + (a2 + a3);
          s2 += (a0 * v[i].x + a1 * v[i].y) + (a2 * v[i].z + a3 * v[i].w);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) sstats[warp] = make_float4(mean, rstd, s1 / d, s2 * rstd / d);
    }
    __syncthreads();
    // ---- phase 2
    if (cg < nv) {
#pragma unroll 4
      for (int rl = 0; rl < nr; ++rl) {
        const size_t off = static_cast<size_t>(r0 + rl) * nv + cg;
        const float4 xv = reinterpret_cast<const float4*>(xs + rl * d)[cg];
        const float4 dyv = reinterpret_cast<const float4*>(ys + rl * d)[cg];
        const float4 st = sstats[rl];  // mean, rstd, m1, m2
        float4 xh;
        xh.x = (xv.x - st.x) * st.y; xh.y = (xv.y - st.x) * st.y; xh.z = (xv.z - st.x) * st.y; xh.w = (xv.w - st.x) * st.y;
        float4 o;
        o.x = st.y * (gg.x * dyv.x - st.z - xh.x * st.w);
        o.y = st.y * (gg.y * dyv.y - st.z - xh.y * st.w);
        o.z = st.y * (gg.z * dyv.z - st.z - xh.z * st.w);
        o.w = st.y * (gg.w * dyv.w - st.z - xh.w * st.w);
        if (dres) {
          const float4 rr = reinterpret_cast<const float4*>(rs + rl * d)[cg];
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        reinterpret_cast<float4*>(dx)[off] = o;
        if (dx_b) reinterpret_cast<uint2*>(dx_b)[off] = make_uint2(cvt_bf16x2(o.x, o.y), cvt_bf16x2(o.z, o.w));
        ag.x += dyv.x * xh.x; ag.y += dyv.y * xh.y; ag.z += dyv.z * xh.z; ag.w += dyv.w * xh.w;
        ab.x += dyv.x; ab.y += dyv.y; ab.z += dyv.z; ab.w += dyv.w;
        ac.x += o.x; ac.y += o.y; ac.z += o.z; ac.w += o.w;
      }
    }
    __syncthreads();  // the slot and sstats may be overwritten
  }
  if (cg >= nv || my_slabs == 0) return;
  atomicAdd(dgamma + cg * 4 + 0, ag.x); atomicAdd(dgamma + cg * 4 + 1, ag.y);
  atomicAdd(dgamma + cg * 4 + 2, ag.z); atomicAdd(dgamma + cg * 4 + 3, ag.w);
  atomicAdd(dbeta + cg * 4 + 0, ab.x); atomicAdd(dbeta + cg * 4 + 1, ab.y);
  atomicAdd(dbeta + cg * 4 + 2, ab.z); atomicAdd(dbeta + cg * 4 + 3, ab.w);
  if (colsum) {
    atomicAdd(colsum + cg * 4 + 0, ac.x); atomicAdd(colsum + cg * 4 + 1, ac.y);
    atomicAdd(colsum + cg * 4 + 2, ac.z); atomicAdd(colsum + cg * 4 + 3, ac.w);
  }
}

// =========================================================================================== cast + column sums
// y_bf16[r, c] = bf16(x[r, c]) (row pitch ldy, zero padding of columns [n, ldy)); colsum[c] += sum_r x[r, c] (optional)
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* __restrict__ x, int ldx, bf16* __restrict__ y,
                                                          int ldy, float* __restrict__ colsum, int rows, int n,
                                                          int rows_per_block) {
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < ldy; c += 256) {
    float s = 0.f;
    if (c < n) {
      for (int r = r0; r < r1; ++r) {
        const float v = x[static_cast<size_t>(r) * ldx + c];
        s += v;
        if (y) y[static_cast<size_t>(r) * ldy + c] = __float2bfloat16_rn(v);
      }
      if (colsum) atomicAdd(colsum + c, s);
    } else if (y) {
      for (int r = r0; r < r1; ++r) y[static_cast<size_t>(r) * ldy + c] = __float2bfloat16_rn(0.f);
    }
  }
}
// vector path (ldx, ldy, n multiples of 4, n == ldy): thread = float4 column group, block = 64 rows x 1024 columns
__global__ void __launch_bounds__(256) cast_colsum_vec_kernel(const float* __restrict__ x, int ldx, bf16* __restrict__ y,
                                                              int ldy, float* __restrict__ colsum, int rows, int n,
                                                              int rows_per_block) {
  const int cg = blockIdx.y * 256 + threadIdx.x;
  if (cg * 4 >= n) return;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int r = r0; r < r1; ++r) {
    const float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ldx + cg * 4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    if (y) *reinterpret_cast<uint2*>(y + static_cast<size_t>(r) * ldy + cg * 4) =
        make_uint2(cvt_bf16x2(v.x, v.y), cvt_bf16x2(v.z, v.w));
  }
  if (colsum) {
    atomicAdd(colsum + cg * 4 + 0, s.x); atomicAdd(colsum + cg * 4 + 1, s.y);
    atomicAdd(colsum + cg * 4 + 2, s.z); atomicAdd(colsum + cg * 4 + 3, s.w);
  }
}
// column sums of a bf16 matrix, 8 columns per thread (n % 8 == 0, 16-byte aligned rows)
__global__ void __launch_bounds__(256) colsum_bf16_vec_kernel(const bf16* __restrict__ x, int ldx,
                                                              float* __restrict__ colsum, int rows, int n,
                                                              int rows_per_block) {
  const int cg = blockIdx.y * 256 + threadIdx.x;
  if (cg * 8 >= n) return;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int r = r0; r < r1; ++r) {
    const uint4 q = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * ldx + cg * 8);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[2 * i] += __uint_as_float(w[i] << 16);
      s[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(colsum + cg * 8 + i, s[i]);
}

// column sums of a bf16 matrix (bias gradient of the FFN hidden layer)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ x, int ldx,
                                                          float* __restrict__ colsum, int rows, int n,
                                                          int rows_per_block) {
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < n; c += 256) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += __bfloat162float(x[static_cast<size_t>(r) * ldx + c]);
    atomicAdd(colsum + c, s);
  }
}

// gather / scatter of the per-clip row slices between [B*seq_in, d] and [B*seq_out, d] (concat backward)
__global__ void __launch_bounds__(256) slice_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                         long long rows_dst, int d, int seq_dst, int seq_src,
                                                         int seq_off) {
  const long long total = rows_dst * (d / 4);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long r = i / (d / 4);
    const int c = static_cast<int>(i % (d / 4));
    const long long sr = (r / seq_dst) * seq_src + seq_off + r % seq_dst;
    reinterpret_cast<float4*>(dst)[r * (d / 4) + c] = reinterpret_cast<const float4*>(src)[sr * (d / 4) + c];
  }
}

// =========================================================================================== embedding backward
// dW[f, c] += sum_r x[r, f] * dy[r, c] ; dpos[r % n_tok, c] += dy[r, c] ; dbias[c] += sum_r dy[r, c]
// grid (d / 64, rows / 256): each block owns 64 output columns and 256 rows; threads: 64 columns x 4 row lanes
__global__ void __launch_bounds__(256) embed_bwd_kernel(const float* __restrict__ x, long long x_batch_stride,
                                                        const float* __restrict__ dy, float* __restrict__ dW,
                                                        float* __restrict__ dbias, float* __restrict__ dpos, int rows,
                                                        int n_tok, int f, int d) {
  extern __shared__ float xs[];  // [64 rows][f]
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int rl = threadIdx.x >> 6;  // 0..3
  const int rbase = blockIdx.y * 256;
  float bsum = 0.f;
  for (int chunk = 0; chunk < 4; ++chunk) {
    const int r0 = rbase + chunk * 64;
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * f; e += 256) {
      const int r = r0 + e / f, k = e % f;
      xs[e] = r < rows ? x[static_cast<size_t>(r / n_tok) * x_batch_stride + static_cast<size_t>(r % n_tok) * f + k] : 0.f;
    }
    __syncthreads();
    // each thread: its column c, rows rl, rl+4, ... of the chunk; accumulate dW[k, c] over those rows per k
    float dyv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = r0 + rl + 4 * i;
      dyv[i] = (r < rows && c < d) ? dy[static_cast<size_t>(r) * d + c] : 0.f;
      bsum += dyv[i];
      if (r < rows && c < d && dpos) atomicAdd(dpos + static_cast<size_t>(r % n_tok) * d + c, dyv[i]);
    }
    if (c < d) {
      for (int k = 0; k < f; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = fmaf(xs[(rl + 4 * i) * f + k], dyv[i], acc);
        atomicAdd(dW + static_cast<size_t>(k) * d + c, acc);
      }
    }
  }
  if (c < d) atomicAdd(dbias + c, bsum);
}

// =========================================================================================== Adam (Keras semantics)
// tf.keras.optimizers.Adam (trainer.py:150): m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
// w -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)      (eps = 1e-7 outside the root)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float lr_t, float b1, float b2, float eps, float grad_scale,
                                                   bf16* __restrict__ w_bf16) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float wi = w[i] - lr_t * mi / (sqrtf(vi) + eps);
    w[i] = wi;
    if (w_bf16) w_bf16[i] = __float2bfloat16_rn(wi);
  }
}

// ---- data-parallel optimizer step fused with its collective, over NVLink peer memory --------------------------------
// What the reference does in three framework steps -- cross-replica SUM of the gradients (MirroredStrategy inside
// apply_gradients, single_task_trainer.py:186-187), Keras Adam (trainer.py:150), variables mirrored on every replica --
// is ONE kernel here.  Every rank owns a 1/world shard of the flat bucket and, for the elements of its shard:
//   reduce     g = sum over replicas of their local gradient       (P2P loads from every peer's bucket, or ONE
//                                                                    multimem.ld_reduce through the NVSwitch multicast)
//   update     m, v, w  (the Adam moments of a shard live only on its owner: optimizer state is sharded)
//   broadcast  w (fp32 master) and bf16(w) (the operand mirror) stored into EVERY replica's buffers
//                                                                   (P2P stores, or ONE multimem.st each)
// so a gradient element crosses the fabric once and the new weight once; there is no separate all-reduce, no second
// pass over the bucket, and replicas stay bit-identical because all of them receive the same stored values.
// The host brackets the launch with two device-side barriers over the symmetric memory (all gradients final before,
// all stores landed after).
constexpr int DP_MAX_WORLD = 16;
struct DpPeers {
  char* base[DP_MAX_WORLD];  // arena of rank p as mapped into THIS process
};

__device__ __forceinline__ float4 ld_f4(const char* p) { return *reinterpret_cast<const float4*>(p); }

template <bool MC>
__global__ void __launch_bounds__(256) dp_adam_kernel(DpPeers peers, char* mc_base, long long grad_off, long long w_off,
                                                      long long wb_off, float* __restrict__ m, float* __restrict__ v,
                                                      long long lo, long long hi, int rank, int world, float lr_t,
                                                      float b1, float b2, float eps, float grad_scale) {
  // elements [lo, hi) of the bucket, four per thread (lo and hi are multiples of 4)
  for (long long i = lo + 4 * (blockIdx.x * 256ll + threadIdx.x); i < hi; i += 4 * 256ll * gridDim.x) {
    float4 g;
    if (MC) {
      const char* a = mc_base + grad_off + 4 * i;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w)
                   : "l"(a)
                   : "memory");
    } else {
      g = ld_f4(peers.base[0] + grad_off + 4 * i);
      for (int p = 1; p < world; ++p) {  // fixed order: every run sums the replicas the same way
        const float4 t = ld_f4(peers.base[p] + grad_off + 4 * i);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
    }
    const float4 w0 = ld_f4(peers.base[rank] + w_off + 4 * i);
    float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    float gg[4] = {g.x * grad_scale, g.y * grad_scale, g.z * grad_scale, g.w * grad_scale};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w}, wa[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ma[e] = b1 * ma[e] + (1.f - b1) * gg[e];
      va[e] = b2 * va[e] + (1.f - b2) * gg[e] * gg[e];
      wa[e] = wa[e] - lr_t * ma[e] / (sqrtf(va[e]) + eps);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(va[0], va[1], va[2], va[3]);
    const uint32_t h0 = cvt_bf16x2(wa[0], wa[1]), h1 = cvt_bf16x2(wa[2], wa[3]);
    if (MC) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_base + w_off + 4 * i),
                   "f"(wa[0]), "f"(wa[1]), "f"(wa[2]), "f"(wa[3])
                   : "memory");
      asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc_base + wb_off + 2 * i),
                   "f"(__uint_as_float(h0)), "f"(__uint_as_float(h1))
                   : "memory");
    } else {
      for (int p = 0; p < world; ++p) {
        *reinterpret_cast<float4*>(peers.base[p] + w_off + 4 * i) = make_float4(wa[0], wa[1], wa[2], wa[3]);
        *reinterpret_cast<uint2*>(peers.base[p] + wb_off + 2 * i) = make_uint2(h0, h1);
      }
    }
  }
  __threadfence_system();
}

// sum of squares (global-norm clipping, single_task_trainer.py:180-183)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) s = fmaf(g[i], g[i], s);
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

// bf16 copy of a Keras-layout kernel without transposing (B operand of the dX = dY . W^T GEMMs), optional column pad
__global__ void __launch_bounds__(256) cast_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out,
                                                          int rows, int cols, int ld_out) {
  const long long total = static_cast<long long>(rows) * ld_out;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int r = static_cast<int>(i / ld_out), c = static_cast<int>(i % ld_out);
    out[i] = __float2bfloat16_rn(c < cols ? w[static_cast<size_t>(r) * cols + c] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------- internal launchers
int wgrad_gemm_pair(const void* x_bf16, int ldx, const void* dy_bf16, int ldy, float* dW, int ldw, int tokens, int in_dim,
                    int out_dim, cudaStream_t st, bool* done);  // wgrad_tc.cu

int wgrad_gemm(const void* x_bf16, int ldx, const void* dy_bf16, int ldy, float* dW, int ldw, int tokens, int in_dim,
               int out_dim, cudaStream_t st) {
  FACT_REQUIRE(x_bf16 && dy_bf16 && dW && tokens > 0 && in_dim > 0 && out_dim > 0, FACT_ERR_BAD_SHAPE,
               "wgrad_gemm: bad arguments");
  {
    bool done = false;  // large problems: 256 x 256 CTA-pair tiles
    const int rc2 = wgrad_gemm_pair(x_bf16, ldx, dy_bf16, ldy, dW, ldw, tokens, in_dim, out_dim, st, &done);
    if (rc2 != FACT_OK || done) return rc2;
  }
  static bool attr_done = false;
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
    attr_done = true;
  }
  CUtensorMap tmx, tmy;
  int rc;
  if ((rc = make_tmap_bf16(&tmx, x_bf16, tokens, in_dim, ldx, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tmy, dy_bf16, tokens, out_dim, ldy, 64, 64))) return rc;
  const int tiles_in = (in_dim + WG_BM - 1) / WG_BM, tiles_out = (out_dim + WG_BN - 1) / WG_BN;
  const int num_kb = (tokens + WG_BK - 1) / WG_BK;
  // enough (tile, K-slice) work items for >= 4 full waves (168 tiles on 148 SMs with one slice each would leave the
  // second wave 13 % full), but at least 16 K blocks per slice so the fp32 red.add epilogue stays amortised
  const int tiles = tiles_in * tiles_out;
  int splits = (4 * num_sms() + tiles - 1) / tiles;
  if (splits > num_kb / 16) splits = num_kb / 16;
  if (splits < 1) splits = 1;
  const int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;
  dim3 grid(tiles_out, tiles_in, splits);
  gemm_wgrad_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, st>>>(tmx, tmy, dW, ldw, in_dim, out_dim, num_kb, kb_per);
  FACT_LAUNCH_CHECK("gemm_wgrad_kernel launch");
  return FACT_OK;
}

int sdpa_bwd_tc_try(const bf16* qkv, const bf16* d_o, const float* lse, const float* Dv, bf16* dqkv, float* dq_acc,
                    int batch, int n, int heads, int head_dim, float k_scale, cudaStream_t st, bool* done);  // sdpa_bwd_tc.cu
int sdpa_bwd_tc2_try(const bf16* qkv, const bf16* d_o, const float* lse, const float* Dv, bf16* dqkv, float* dq_acc,
                     int batch, int n, int heads, int head_dim, float k_scale, cudaStream_t st, bool* done);  // sdpa_bwd_tc2.cu
extern int g_sdpa_bwd_tc;

int sdpa_backward(const void* qkv, const void* o, const void* d_o, const float* lse, float* Dv, float* dq_acc,
                  void* dqkv, int batch, int n, int heads, int head_dim, float scale, cudaStream_t st) {
  FACT_REQUIRE(qkv && o && d_o && lse && Dv && dq_acc && dqkv, FACT_ERR_BAD_SHAPE, "sdpa_backward: null buffer");
  FACT_REQUIRE(head_dim == 80 || head_dim == 64 || head_dim == 32 || head_dim == 16, FACT_ERR_UNSUPPORTED,
               "sdpa_backward: head_dim %d", head_dim);
  const int d = heads * head_dim;
  const long long tokens = static_cast<long long>(batch) * n;
  FACT_CUDA_CHECK(cudaMemsetAsync(dq_acc, 0, tokens * d * sizeof(float), st));
  {
    FACT_REQUIRE(d <= 1024, FACT_ERR_UNSUPPORTED, "sdpa_backward: d_model %d", d);
    sdpa_bwd_prep_kernel<<<static_cast<int>((tokens + 7) / 8), 256, 0, st>>>(
        static_cast<const bf16*>(o), static_cast<const bf16*>(d_o), Dv, static_cast<int>(tokens), n, heads, head_dim);
    FACT_LAUNCH_CHECK("sdpa_bwd_prep_kernel");
  }
  dim3 grid((n + 63) / 64, heads, batch);
  const float k_scale = 0.6931471805599453f;  // dK = ln2 * dZ^T q'   (q' = q * scale * log2 e)
  bool tc_done = false;  // FACT shapes: the tcgen05 kernels (flag sdpa_bwd_tc: 1 = pipelined, 2 = first generation)
  if (g_sdpa_bwd_tc == 1) {
    const int rc = sdpa_bwd_tc2_try(static_cast<const bf16*>(qkv), static_cast<const bf16*>(d_o), lse, Dv,
                                    static_cast<bf16*>(dqkv), dq_acc, batch, n, heads, head_dim, k_scale, st, &tc_done);
    if (rc != FACT_OK) return rc;
  }
  if (!tc_done) {
    const int rc = sdpa_bwd_tc_try(static_cast<const bf16*>(qkv), static_cast<const bf16*>(d_o), lse, Dv,
                                   static_cast<bf16*>(dqkv), dq_acc, batch, n, heads, head_dim, k_scale, st, &tc_done);
    if (rc != FACT_OK) return rc;
  }
#define FACT_BWD_CASE(DHV)                                                                                         \
  case DHV: {                                                                                                      \
    constexpr int smem = (6 * 64 * (DHV + 8) + 64 * 72) * 2 + 4 * 64 * 4;                                          \
    static bool done = false;                                                                                      \
    if (!done) {                                                                                                   \
      FACT_CUDA_CHECK(cudaFuncSetAttribute(sdpa_bwd_kernel<DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      done = true;                                                                                                 \
    }                                                                                                              \
    sdpa_bwd_kernel<DHV><<<grid, 128, smem, st>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(d_o), lse, \
                                                  Dv, static_cast<bf16*>(dqkv), dq_acc, n, heads, k_scale);        \
    break;                                                                                                         \
  }
  if (!tc_done) {
    switch (head_dim) {
      FACT_BWD_CASE(16)
      FACT_BWD_CASE(32)
      FACT_BWD_CASE(64)
      FACT_BWD_CASE(80)
    }
    FACT_LAUNCH_CHECK("sdpa_bwd_kernel");
  }
#undef FACT_BWD_CASE
  long long total = tokens * (d / 4);
  int g2 = static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  sdpa_bwd_finish_kernel<<<g2, 256, 0, st>>>(dq_acc, static_cast<bf16*>(dqkv), tokens, d, scale);
  FACT_LAUNCH_CHECK("sdpa_bwd_finish_kernel");
  return FACT_OK;
}

// dx_b (bf16 copy of dx) and colsum (+= column sums of dx) are optional: they fold the cast_colsum pass that would
// otherwise re-read dx for the next dense layer's weight / bias gradients into this kernel
int ln_backward(const float* x, const float* gamma, const float* dy, const float* dres, float* dx, void* dx_b,
                float* dgamma, float* dbeta, float* colsum, int rows, int d, cudaStream_t st) {
  FACT_REQUIRE(d % 4 == 0 && d <= 1024, FACT_ERR_BAD_SHAPE, "ln_backward: d %d", d);
  FACT_REQUIRE(dy != dx, FACT_ERR_BAD_SHAPE, "ln_backward: dx must not alias dy (it may alias dres)");
  FACT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, FACT_ERR_BAD_ALIGN,
               "ln_backward: x and dy must be 16-byte aligned");
  FACT_REQUIRE(!dres || (reinterpret_cast<uintptr_t>(dres) & 15) == 0, FACT_ERR_BAD_ALIGN,
               "ln_backward: dres must be 16-byte aligned");
  const int slab_bytes = 3 * LN_SLAB * d * 4;
  int stages = (220 * 1024) / slab_bytes;
  stages = stages > 3 ? 3 : stages;
  FACT_REQUIRE(stages >= 2, FACT_ERR_UNSUPPORTED, "ln_backward: d %d too wide for the shared-memory ring", d);
  const int smem = stages * slab_bytes + LN_SLAB * 16 + stages * 8 + 16;
  static int smem_set = 0;
  if (smem > smem_set) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  const int num_slabs = (rows + LN_SLAB - 1) / LN_SLAB;
  const int grid = num_slabs < num_sms() ? num_slabs : num_sms();
  ln_bwd_kernel<<<grid, 256, smem, st>>>(x, gamma, dy, dres, dx, static_cast<bf16*>(dx_b), dgamma, dbeta, colsum, rows,
                                         d, stages);
  FACT_LAUNCH_CHECK("ln_bwd_kernel");
  return FACT_OK;
}

int cast_colsum(const float* x, int ldx, void* y, int ldy, float* colsum, int rows, int n, cudaStream_t st) {
  const int rpb = 64;
  const bool vec = (n % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && n == ldy &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (!y || (reinterpret_cast<uintptr_t>(y) & 7) == 0);
  if (vec) {
    dim3 grid((rows + rpb - 1) / rpb, (n / 4 + 255) / 256);
    cast_colsum_vec_kernel<<<grid, 256, 0, st>>>(x, ldx, static_cast<bf16*>(y), ldy, colsum, rows, n, rpb);
  } else {
    cast_colsum_kernel<<<(rows + rpb - 1) / rpb, 256, 0, st>>>(x, ldx, static_cast<bf16*>(y), ldy, colsum, rows, n, rpb);
  }
  FACT_LAUNCH_CHECK("cast_colsum_kernel");
  return FACT_OK;
}

int colsum_bf16(const void* x, int ldx, float* colsum, int rows, int n, cudaStream_t st) {
  const int rpb = 64;
  if (n % 8 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 grid((rows + rpb - 1) / rpb, (n / 8 + 255) / 256);
    colsum_bf16_vec_kernel<<<grid, 256, 0, st>>>(static_cast<const bf16*>(x), ldx, colsum, rows, n, rpb);
  } else {
    colsum_bf16_kernel<<<(rows + rpb - 1) / rpb, 256, 0, st>>>(static_cast<const bf16*>(x), ldx, colsum, rows, n, rpb);
  }
  FACT_LAUNCH_CHECK("colsum_bf16_kernel");
  return FACT_OK;
}

int slice_rows(const float* src, float* dst, long long rows_dst, int d, int seq_dst, int seq_src, int seq_off,
               cudaStream_t st) {
  long long total = rows_dst * (d / 4);
  int grid = static_cast<int>((total + 255) / 256 > 8192 ? 8192 : (total + 255) / 256);
  slice_rows_kernel<<<grid, 256, 0, st>>>(src, dst, rows_dst, d, seq_dst, seq_src, seq_off);
  FACT_LAUNCH_CHECK("slice_rows_kernel");
  return FACT_OK;
}

// dpos[t, c] += sum over clips of dy[clip * n_tok + t, c]  (PositionEmbedding gradient, base_models.py:148-156)
__global__ void __launch_bounds__(256) pos_grad_kernel(const float* __restrict__ dy, float* __restrict__ dpos, int batch,
                                                       int n_tok, int d) {
  const int nv = d >> 2;
  const int idx = blockIdx.x * 256 + threadIdx.x;  // (t, column group)
  if (idx >= n_tok * nv) return;
  const int t = idx / nv, cg = idx % nv;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int b = 0; b < batch; ++b) {
    const float4 v = reinterpret_cast<const float4*>(dy + (static_cast<size_t>(b) * n_tok + t) * d)[cg];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4* o = reinterpret_cast<float4*>(dpos + static_cast<size_t>(t) * d) + cg;
  float4 cur = *o;
  cur.x += s.x; cur.y += s.y; cur.z += s.z; cur.w += s.w;
  *o = cur;
}

int pos_grad(const float* dy, float* dpos, int batch, int n_tok, int d, cudaStream_t st) {
  FACT_REQUIRE(d % 4 == 0, FACT_ERR_BAD_SHAPE, "pos_grad: d %d", d);
  const int total = n_tok * (d / 4);
  pos_grad_kernel<<<(total + 255) / 256, 256, 0, st>>>(dy, dpos, batch, n_tok, d);
  FACT_LAUNCH_CHECK("pos_grad_kernel");
  return FACT_OK;
}

int embed_backward(const float* x, long long x_batch_stride, const float* dy, float* dW, float* dbias, float* dpos,
                   int batch, int n_tok, int f, int d, cudaStream_t st) {
  const int rows = batch * n_tok;
  dim3 grid((d + 63) / 64, (rows + 255) / 256);
  const size_t smem = static_cast<size_t>(64) * f * sizeof(float);
  static size_t max_set = 0;
  if (smem > 48 * 1024 && smem > max_set) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    max_set = smem;
  }
  embed_bwd_kernel<<<grid, 256, smem, st>>>(x, x_batch_stride, dy, dW, dbias, dpos, rows, n_tok, f, d);
  FACT_LAUNCH_CHECK("embed_bwd_kernel");
  return FACT_OK;
}

int cast_weight(const float* w, void* out, int rows, int cols, int ld_out, cudaStream_t st) {
  long long total = static_cast<long long>(rows) * ld_out;
  int grid = static_cast<int>((total + 255) / 256 > 8192 ? 8192 : (total + 255) / 256);
  cast_weight_kernel<<<grid, 256, 0, st>>>(w, static_cast<bf16*>(out), rows, cols, ld_out);
  FACT_LAUNCH_CHECK("cast_weight_kernel");
  return FACT_OK;
}

// g *= clip / max(sqrt(*sumsq), clip): the factor of tf.clip_by_global_norm from a device-resident squared norm
__global__ void clip_scale_kernel(float* __restrict__ g, long long n, const float* __restrict__ sumsq, float clip) {
  const float norm = sqrtf(*sumsq);
  const float f = clip / fmaxf(norm, clip);
  if (f == 1.f) return;
  const long long n4 = n / 4;
  float4* g4 = reinterpret_cast<float4*>(g);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = g4[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    g4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) g[n4 * 4 + threadIdx.x] *= f;
}

}  // namespace fact

using namespace fact;

// ------------------------------------------------------------------------------------------- C ABI (training blocks)
extern "C" int fact_clip_scale(float* g, long long n, const float* sum_squares, float clip_norm, void* stream) {
  FACT_REQUIRE(g && sum_squares && n > 0 && clip_norm > 0.f, FACT_ERR_BAD_SHAPE, "fact_clip_scale: bad arguments");
  FACT_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, FACT_ERR_BAD_ALIGN, "fact_clip_scale: g must be 16-B aligned");
  const long long blocks = (n / 4 + 255) / 256;
  const int grid = static_cast<int>(blocks > 8192 ? 8192 : (blocks < 1 ? 1 : blocks));
  clip_scale_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, n, sum_squares, clip_norm);
  FACT_LAUNCH_CHECK("clip_scale_kernel");
  return FACT_OK;
}

extern "C" int fact_wgrad_gemm(const void* x_bf16, int ldx, const void* dy_bf16, int ldy, float* dw, int ldw,
                               int tokens, int in_dim, int out_dim, void* stream) {
  return wgrad_gemm(x_bf16, ldx, dy_bf16, ldy, dw, ldw, tokens, in_dim, out_dim, as_stream(stream));
}

extern "C" int fact_sdpa_backward(const void* qkv, const void* o, const void* d_o, const float* lse, float* d_scratch,
                                  float* dq_scratch, void* dqkv, int batch, int n, int heads, int head_dim,
                                  float scale, void* stream) {
  return sdpa_backward(qkv, o, d_o, lse, d_scratch, dq_scratch, dqkv, batch, n, heads, head_dim, scale,
                       as_stream(stream));
}

extern "C" int fact_layernorm_backward(const float* x, const float* gamma, const float* dy, const float* dres,
                                       float* dx, float* dgamma, float* dbeta, float* stats_scratch, int rows, int d,
                                       void* stream) {
  FACT_REQUIRE(x && gamma && dy && dx && dgamma && dbeta && stats_scratch && rows > 0, FACT_ERR_BAD_SHAPE,
               "fact_layernorm_backward: bad arguments");
  (void)stats_scratch;  // kept in the signature; the row statistics now live in shared memory
  return ln_backward(x, gamma, dy, dres, dx, nullptr, dgamma, dbeta, nullptr, rows, d, as_stream(stream));
}

extern "C" int fact_embed_backward(const float* x, long long x_batch_stride, const float* dy, float* dw, float* dbias,
                                   float* dpos, int batch, int n_tok, int f, int d, void* stream) {
  FACT_REQUIRE(x && dy && dw && dbias && batch > 0 && n_tok > 0 && f > 0 && d > 0, FACT_ERR_BAD_SHAPE,
               "fact_embed_backward: bad arguments");
  return embed_backward(x, x_batch_stride, dy, dw, dbias, dpos, batch, n_tok, f, d, as_stream(stream));
}

extern "C" int fact_cast_colsum(const float* x, int ldx, void* y_bf16, int ldy, float* colsum, int rows, int n,
                                void* stream) {
  FACT_REQUIRE(x && rows > 0 && n > 0 && ldy >= n, FACT_ERR_BAD_SHAPE, "fact_cast_colsum: bad arguments");
  return cast_colsum(x, ldx, y_bf16, ldy, colsum, rows, n, as_stream(stream));
}

extern "C" int fact_adam_step(float* w, const float* g, float* m, float* v, long long n, float lr, float beta1,
                              float beta2, float eps, long long step, float grad_scale, void* w_bf16, void* stream) {
  FACT_REQUIRE(w && g && m && v && n > 0 && step >= 1, FACT_ERR_BAD_SHAPE, "fact_adam_step: bad arguments");
  const double t = static_cast<double>(step);
  const float lr_t = static_cast<float>(lr * sqrt(1.0 - pow(static_cast<double>(beta2), t)) /
                                        (1.0 - pow(static_cast<double>(beta1), t)));
  long long blocks = (n + 255) / 256;
  int grid = static_cast<int>(blocks > 16384 ? 16384 : blocks);
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(w, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale,
                                                   static_cast<bf16*>(w_bf16));
  FACT_LAUNCH_CHECK("adam_kernel");
  return FACT_OK;
}

extern "C" int fact_dp_adam_range(void* const* peer_base, void* mc_base, long long grad_off, long long w_off,
                                  long long wb_off, float* m, float* v, long long off, long long count, int rank,
                                  int world, float lr, float beta1, float beta2, float eps, long long step,
                                  float grad_scale, int max_blocks, void* stream) {
  FACT_REQUIRE(peer_base && m && v && off >= 0 && count > 0 && step >= 1, FACT_ERR_BAD_SHAPE,
               "fact_dp_adam_range: bad arguments");
  FACT_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, FACT_ERR_BAD_SHAPE,
               "fact_dp_adam_range: rank %d of %d (at most %d replicas)", rank, world, DP_MAX_WORLD);
  FACT_REQUIRE(grad_off % 16 == 0 && w_off % 16 == 0 && wb_off % 16 == 0 && off % 4 == 0 && count % 4 == 0,
               FACT_ERR_BAD_ALIGN, "fact_dp_adam_range: byte offsets must be 16-byte aligned, off and count multiples of 4");
  DpPeers peers{};
  for (int p = 0; p < world; ++p) {
    FACT_REQUIRE(peer_base[p] != nullptr && (reinterpret_cast<uintptr_t>(peer_base[p]) & 15) == 0, FACT_ERR_BAD_ALIGN,
                 "fact_dp_adam_range: arena of rank %d is null or misaligned", p);
    peers.base[p] = static_cast<char*>(peer_base[p]);
  }
  // shard r of the range = elements off + [r * per, min(count, (r + 1) * per)), per a multiple of 8
  long long per = (count + world - 1) / world;
  per = (per + 7) / 8 * 8;
  const long long lo = off + (per * rank < count ? per * rank : count);
  const long long hi = lo + per < off + count ? lo + per : off + count;
  if (hi <= lo) return FACT_OK;
  const double t = static_cast<double>(step);
  const float lr_t = static_cast<float>(lr * sqrt(1.0 - pow(static_cast<double>(beta2), t)) /
                                        (1.0 - pow(static_cast<double>(beta1), t)));
  const long long threads = (hi - lo) / 4;
  long long blocks = (threads + 255) / 256;
  const long long cap = max_blocks > 0 ? max_blocks : 4096;
  const int grid = static_cast<int>(blocks > cap ? cap : blocks);
  if (mc_base)
    dp_adam_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(peers, static_cast<char*>(mc_base), grad_off, w_off, wb_off,
                                                              m, v, lo, hi, rank, world, lr_t, beta1, beta2, eps,
                                                              grad_scale);
  else
    dp_adam_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(peers, nullptr, grad_off, w_off, wb_off, m, v, lo, hi, rank,
                                                               world, lr_t, beta1, beta2, eps, grad_scale);
  FACT_LAUNCH_CHECK("dp_adam_kernel");
  return FACT_OK;
}

extern "C" int fact_dp_adam_step(void* const* peer_base, void* mc_base, long long grad_off, long long w_off,
                                 long long wb_off, float* m, float* v, long long n, int rank, int world, float lr,
                                 float beta1, float beta2, float eps, long long step, float grad_scale, void* stream) {
  return fact_dp_adam_range(peer_base, mc_base, grad_off, w_off, wb_off, m, v, 0, n, rank, world, lr, beta1, beta2, eps,
                            step, grad_scale, 0, stream);
}

extern "C" int fact_sum_squares(const float* g, long long n, float* out, void* stream) {
  FACT_REQUIRE(g && out && n > 0, FACT_ERR_BAD_SHAPE, "fact_sum_squares: bad arguments");
  FACT_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float), as_stream(stream)));
  long long blocks = (n + 255) / 256;
  int grid = static_cast<int>(blocks > 4096 ? 4096 : blocks);
  sumsq_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, n, out);
  FACT_LAUNCH_CHECK("sumsq_kernel");
  return FACT_OK;
}

extern "C" int fact_cast_weight(const float* w_keras, void* out_bf16, int rows, int cols, int ld_out, void* stream) {
  FACT_REQUIRE(w_keras && out_bf16 && rows > 0 && cols > 0 && ld_out >= cols, FACT_ERR_BAD_SHAPE,
               "fact_cast_weight: bad arguments");
  return cast_weight(w_keras, out_bf16, rows, cols, ld_out, as_stream(stream));
}
