// tcgen05 attention-core backward for the FACT shapes (N <= 384, head_dim 80, bf16 products): the gradient of
// softmax(q.k^T * d_model^-0.5) . v (mint/core/base_models.py:76-86) that tape.gradient produces in
// mint/ctl/single_task_trainer.py:176-178.  Same algebra as the mma.sync kernel in backward.cu (flash style, P
// recomputed from the forward's log2-sum-exp), restated for TMEM:
//
//   work item = (batch, head, 128-key block j); for every 128-query block i:
//     S^T  = K_j . Q_i^T          [keys x queries]  tcgen05.mma, both operands K-major SWIZZLE_32B sub-tiles
//     dP^T = V_j . dO_i^T         [keys x queries]
//     P^T  = 2^(S^T - lse[q]),  dZ^T = P^T * (dP^T - D[q])      softmax warps: TMEM lane = key, thread-local
//     dV_j += P^T . dO_i          P^T stays in TMEM (bf16, written in place over S^T) as the A operand
//     dK_j += dZ^T . Q_i          dZ^T likewise over dP^T
//     dQ_i  = dZ . K_j            A = dZ^T staged in shared memory as an MN-major SWIZZLE_128B tile
//   dV_j / dK_j accumulate in TMEM over i and leave as bf16 rows of dqkv; the dQ_i partial of each (i, j) goes through
//   swizzled shared memory and a bulk tensor reduction (fp32 add in the L2) into dq_acc.
// TMEM columns: S^T / P^T [0,128)  dP^T / dZ^T [128,256)  dV [256,336)  dK [336,416)  dQ [416,496).
// Keys / queries past N inside a block are other rows of the buffers (or TMA zero fill): their P and dZ entries are
// written as exact zeros, so they add nothing anywhere.
#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {

constexpr int BT_B = 128;             // block edge (keys per item, queries per inner step)
constexpr int BT_DH = 80;
constexpr int BT_KS = BT_DH / 16;     // 16-wide head_dim slices
constexpr int BT_SUB = BT_B * 32;     // bytes of one [128 rows][16 el] sub-tile
constexpr int BT_OP = BT_KS * BT_SUB; // 20 KB: one 128 x 80 operand
constexpr int BT_SM_WARPS = 8;
constexpr int BT_THREADS = 64 + BT_SM_WARPS * 32;
constexpr int BT_COL_S = 0, BT_COL_DP = 128, BT_COL_DV = 256, BT_COL_DK = 336, BT_COL_DQ = 416;

// shared memory map (offsets from the 1024-aligned base)
constexpr int BT_OFF_K = 0;
constexpr int BT_OFF_V = BT_OFF_K + BT_OP;
constexpr int BT_OFF_QO = BT_OFF_V + BT_OP;              // 2 stages x (Q_i, dO_i)
constexpr int BT_OFF_DS = BT_OFF_QO + 4 * BT_OP;         // dZ^T, MN-major SW128: 2 atoms x [128 keys][128 B]
constexpr int BT_OFF_DQ = BT_OFF_DS + 2 * 16384;         // dQ staging: 8 warps x 3 x [32 rows][64 B]
constexpr int BT_OFF_LD = BT_OFF_DQ + BT_SM_WARPS * 3 * 2048;  // lse[128], D[128]
constexpr int BT_OFF_BAR = BT_OFF_LD + 2 * BT_B * 4;
constexpr int BT_SMEM_BYTES = 1024 + BT_OFF_BAR + 256;
static_assert(BT_SMEM_BYTES <= 232448, "exceeds 227 KB");
static_assert(BT_OFF_DS % 1024 == 0 && BT_OFF_DQ % 1024 == 0, "swizzled tiles need 1024-byte alignment");

__global__ void __launch_bounds__(BT_THREADS, 1)
sdpa_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dq, const float* __restrict__ lse,
                   const float* __restrict__ Dv, bf16* __restrict__ dqkv, int N, int H, int nblk, int num_items,
                   float k_scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t k_smem = smem_base + BT_OFF_K, v_smem = smem_base + BT_OFF_V;
  auto q_smem = [&](int st) { return smem_base + BT_OFF_QO + st * 2 * BT_OP; };
  auto do_smem = [&](int st) { return smem_base + BT_OFF_QO + st * 2 * BT_OP + BT_OP; };
  const uint32_t ds_smem = smem_base + BT_OFF_DS;
  const uint32_t bar_base = smem_base + BT_OFF_BAR;
  // barriers: kv_full kv_empty qo_full[2] qo_empty[2] sdp_full p_ready dq_full dq_empty dkv_full dkv_empty | tmem ptr
  const uint32_t kv_full = bar_base, kv_empty = bar_base + 8;
  auto qo_full = [&](int st) { return bar_base + 16u + 8u * st; };
  auto qo_empty = [&](int st) { return bar_base + 32u + 8u * st; };
  const uint32_t sdp_full = bar_base + 48, p_ready = bar_base + 56, dq_full = bar_base + 64, dq_empty = bar_base + 72;
  const uint32_t dkv_full = bar_base + 80, dkv_empty = bar_base + 88, tmem_ptr_addr = bar_base + 96;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + BT_OFF_BAR + 96);
  float* lse_s = reinterpret_cast<float*>(smem_gen + BT_OFF_LD);
  float* d_s = lse_s + BT_B;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * BT_DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int st = 0; st < 2; ++st) {
      mbar_init(qo_full(st), 1);
      mbar_init(qo_empty(st), 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(p_ready, BT_SM_WARPS);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, BT_SM_WARPS);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_empty, BT_SM_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto item_coords = [&](int item, int& b, int& h, int& jb) {
    jb = item % nblk;
    const int bh = item / nblk;
    h = bh % H;
    b = bh / H;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    auto load_op = [&](uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
#pragma unroll
      for (int ks = 0; ks < BT_KS; ++ks) tma_load_2d(dst + ks * BT_SUB, tm, col + ks * 16, row, bar);
    };
    uint32_t it = 0, t = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
      int b, h, jb;
      item_coords(item, b, h, jb);
      const int row0 = b * N;
      mbar_wait(kv_empty, (t & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * BT_OP);
        load_op(k_smem, &tm_qkv, D + h * BT_DH, row0 + jb * BT_B, kv_full);
        load_op(v_smem, &tm_qkv, 2 * D + h * BT_DH, row0 + jb * BT_B, kv_full);
      }
      __syncwarp();
      for (int ib = 0; ib < nblk; ++ib, ++it) {
        const int st = it & 1;
        mbar_wait(qo_empty(st), ((it >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(qo_full(st), 2 * BT_OP);
          load_op(q_smem(st), &tm_qkv, h * BT_DH, row0 + ib * BT_B, qo_full(st));
          load_op(do_smem(st), &tm_do, h * BT_DH, row0 + ib * BT_B, qo_full(st));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16_f32_ex(BT_B, BT_B, 0);          // K-major x K-major
    constexpr uint32_t idesc_ts = umma_idesc_bf16_f32_ex(BT_B, BT_DH, 1);        // TMEM A, MN-major B
    constexpr uint32_t idesc_dq = umma_idesc_bf16_f32_maj(BT_B, BT_DH, 1, 1);    // MN-major A and B
    uint32_t it = 0, t = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
      int b, h, jb;
      item_coords(item, b, h, jb);
      const int ksteps_k = (min(BT_B, N - jb * BT_B) + 15) >> 4;
      mbar_wait(kv_full, t & 1);
      for (int ib = 0; ib < nblk; ++ib, ++it) {
        const int st = it & 1;
        const int ksteps_q = (min(BT_B, N - ib * BT_B) + 15) >> 4;
        mbar_wait(qo_full(st), (it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t kd = umma_desc_k_sw32(k_smem), vd = umma_desc_k_sw32(v_smem);
          const uint64_t qd = umma_desc_k_sw32(q_smem(st)), od = umma_desc_k_sw32(do_smem(st));
#pragma unroll
          for (int ks = 0; ks < BT_KS; ++ks) {
            const uint64_t off = static_cast<uint64_t>(ks * BT_SUB) >> 4;
            umma_bf16(tmem_base + BT_COL_S, kd + off, qd + off, idesc_s, ks > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < BT_KS; ++ks) {
            const uint64_t off = static_cast<uint64_t>(ks * BT_SUB) >> 4;
            umma_bf16(tmem_base + BT_COL_DP, vd + off, od + off, idesc_s, ks > 0 ? 1u : 0u);
          }
          umma_commit(sdp_full);
        }
        __syncwarp();
        mbar_wait(p_ready, it & 1);
        if (ib == 0) mbar_wait(dkv_empty, (t & 1) ^ 1);  // previous item's dV / dK have been read out
        mbar_wait(dq_empty, (it & 1) ^ 1);               // previous step's dQ has been read out
        tc_fence_after();
        if (elect_one()) {
          const uint64_t od_mn = umma_desc_mn_sw32(do_smem(st), BT_SUB, 256);
          const uint64_t qd_mn = umma_desc_mn_sw32(q_smem(st), BT_SUB, 256);
          const uint64_t kd_mn = umma_desc_mn_sw32(k_smem, BT_SUB, 256);
          const uint64_t zd = umma_desc_mn_sw128(ds_smem, 16384, 1024);
#pragma unroll
          for (int j = 0; j < BT_B / 16; ++j)
            if (j < ksteps_q) {
              const uint32_t a_col = 32 * (j >> 1) + 8 * (j & 1);  // 16 queries = 8 packed columns
              const uint64_t off = static_cast<uint64_t>(j * 512) >> 4;
              umma_bf16_ts(tmem_base + BT_COL_DV, tmem_base + BT_COL_S + a_col, od_mn + off, idesc_ts,
                           (ib > 0 || j > 0) ? 1u : 0u);
              umma_bf16_ts(tmem_base + BT_COL_DK, tmem_base + BT_COL_DP + a_col, qd_mn + off, idesc_ts,
                           (ib > 0 || j > 0) ? 1u : 0u);
            }
#pragma unroll
          for (int j = 0; j < BT_B / 16; ++j)
            if (j < ksteps_k)
              umma_bf16(tmem_base + BT_COL_DQ, zd + (static_cast<uint64_t>(j * 2048) >> 4),
                        kd_mn + (static_cast<uint64_t>(j * 512) >> 4), idesc_dq, j > 0 ? 1u : 0u);
          umma_commit(qo_empty(st));
          umma_commit(dq_full);
          if (ib == nblk - 1) {
            umma_commit(dkv_full);
            umma_commit(kv_empty);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps 2..9
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int st_tid = (warp - 2) * 32 + lane;  // 0..255
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    const int rl = q4 * 32 + lane;              // row (TMEM lane) inside the block
    const uint32_t dq_stage = smem_base + BT_OFF_DQ + static_cast<uint32_t>(warp - 2) * 3 * 2048;
    auto sm_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    uint32_t it = 0, t = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
      int b, h, jb;
      item_coords(item, b, h, jb);
      const size_t tok0 = static_cast<size_t>(b) * N;
      const int key = jb * BT_B + rl;
      const bool key_ok = key < N;
      for (int ib = 0; ib < nblk; ++ib, ++it) {
        const int i0 = ib * BT_B;
        // per-query log2-sum-exp and D of this block (zeros past N); the previous step's readers are done: every
        // softmax thread passed its p_ready arrival after its last read
        sm_sync();
        {
          const int qi = i0 + (st_tid & 127);
          const size_t idx = (static_cast<size_t>(b) * H + h) * N + (qi < N ? qi : 0);
          const float val = qi < N ? (st_tid < 128 ? lse[idx] : Dv[idx]) : 0.f;
          (st_tid < 128 ? lse_s : d_s)[st_tid & 127] = val;
        }
        sm_sync();
        mbar_wait(sdp_full, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int g = half; g < 4; g += 2) {
          float s[32], dp[32];
          tmem_ld_32x32(lane_base + BT_COL_S + g * 32, s);
          tmem_ld_32x32(lane_base + BT_COL_DP + g * 32, dp);
          tmem_ld_wait();
          uint32_t pk[16], zk[16];
          const int qlim = N - i0 - g * 32;  // columns [0, qlim) of this group are real queries
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float2 l2 = *reinterpret_cast<const float2*>(lse_s + g * 32 + 2 * c);
            const float2 d2 = *reinterpret_cast<const float2*>(d_s + g * 32 + 2 * c);
            const float p0 = (key_ok && 2 * c < qlim) ? ex2_approx(s[2 * c] - l2.x) : 0.f;
            const float p1 = (key_ok && 2 * c + 1 < qlim) ? ex2_approx(s[2 * c + 1] - l2.y) : 0.f;
            pk[c] = cvt_bf16x2(p0, p1);
            zk[c] = cvt_bf16x2(p0 * (dp[2 * c] - d2.x), p1 * (dp[2 * c + 1] - d2.y));
          }
          tmem_st_32x16(lane_base + BT_COL_S + g * 32, pk);
          tmem_st_32x16(lane_base + BT_COL_DP + g * 32, zk);
          // dZ^T row `rl` (key), queries g*32 .. +31 -> atom g/2, 16-byte chunks (g&1)*4 .. +3, SWIZZLE_128B
          const uint32_t rbase = ds_smem + (g >> 1) * 16384 + rl * 128;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            st_shared_v4(rbase + ((((g & 1) * 4 + c4) ^ (rl & 7)) << 4), zk[4 * c4], zk[4 * c4 + 1], zk[4 * c4 + 2],
                         zk[4 * c4 + 3]);
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);

        // ---- dQ_i partial of this (i, j): TMEM lane = query; half 0 takes columns [0, 48), half 1 [48, 80)
        mbar_wait(dq_full, it & 1);
        tc_fence_after();
        constexpr int C0 = 48;
        float dq[C0];
        if (half == 0) {
          tmem_ld_32x32(lane_base + BT_COL_DQ, dq);
          tmem_ld_32x16(lane_base + BT_COL_DQ + 32, dq + 32);
        } else {
          tmem_ld_32x32(lane_base + BT_COL_DQ + C0, dq);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(dq_empty);
          bulk_wait_group_read0();  // the previous step's boxes have left this warp's staging slot
        }
        __syncwarp();
        {
          const int nsub = half == 0 ? 3 : 2;
          const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
          for (int sb = 0; sb < 3; ++sb)
            if (sb < nsub) {
              const uint32_t rb = dq_stage + sb * 2048 + lane * 64;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4)
                st_shared_v4(rb + ((c4 ^ sw) << 4), __float_as_uint(dq[sb * 16 + 4 * c4]),
                             __float_as_uint(dq[sb * 16 + 4 * c4 + 1]), __float_as_uint(dq[sb * 16 + 4 * c4 + 2]),
                             __float_as_uint(dq[sb * 16 + 4 * c4 + 3]));
            }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && i0 + q4 * 32 < N) {  // rows past N hold exact zeros: skip wholly empty boxes
            const int col = h * BT_DH + (half == 0 ? 0 : C0);
            const int row = static_cast<int>(tok0) + i0 + q4 * 32;
            for (int sb = 0; sb < nsub; ++sb) tma_reduce_add_2d(&tm_dq, dq_stage + sb * 2048, col + sb * 16, row);
            bulk_commit_group();
          }
        }
      }
      // ---- dV_j (half 0) / dK_j (half 1): TMEM lane = key
      mbar_wait(dkv_full, t & 1);
      tc_fence_after();
      {
        float o[BT_DH];
        const uint32_t src = lane_base + (half == 0 ? BT_COL_DV : BT_COL_DK);
        tmem_ld_32x32(src, o);
        tmem_ld_32x32(src + 32, o + 32);
        tmem_ld_32x16(src + 64, o + 64);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_empty);
        if (key_ok) {
          const float sc = half == 0 ? 1.f : k_scale;
          uint4* dst = reinterpret_cast<uint4*>(dqkv + (tok0 + key) * 3 * D + (half == 0 ? 2 * D : D) + h * BT_DH);
#pragma unroll
          for (int i = 0; i < BT_DH / 8; ++i)
            dst[i] = make_uint4(cvt_bf16x2(o[8 * i] * sc, o[8 * i + 1] * sc),
                                cvt_bf16x2(o[8 * i + 2] * sc, o[8 * i + 3] * sc),
                                cvt_bf16x2(o[8 * i + 4] * sc, o[8 * i + 5] * sc),
                                cvt_bf16x2(o[8 * i + 6] * sc, o[8 * i + 7] * sc));
        }
      }
    }
    if (lane == 0) bulk_wait_group0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int g_sdpa_bwd_tc = 1;  // fact_set_flag("sdpa_bwd_tc", 0): keep the mma.sync kernel (A/B timing, tests)

// Returns FACT_OK with *done = true when the tcgen05 kernel took the problem (dq_acc must be zeroed by the caller, D
// already computed); *done = false -> the caller runs the mma.sync kernel.
int sdpa_bwd_tc_try(const bf16* qkv, const bf16* d_o, const float* lse, const float* Dv, bf16* dqkv, float* dq_acc,
                    int batch, int n, int heads, int head_dim, float k_scale, cudaStream_t st, bool* done) {
  *done = false;
  if (!g_sdpa_bwd_tc || head_dim != BT_DH || n > 3 * BT_B) return FACT_OK;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(dqkv) |
       reinterpret_cast<uintptr_t>(dq_acc)) & 15)
    return FACT_OK;
  static bool attr_done = false;
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(sdpa_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM_BYTES));
    attr_done = true;
  }
  const int d = heads * BT_DH;
  const int tokens = batch * n;
  CUtensorMap tq, to, tdq;
  int rc;
  if ((rc = make_tmap_bf16(&tq, qkv, tokens, 3 * d, 3 * d, BT_B, 16))) return rc;
  if ((rc = make_tmap_bf16(&to, d_o, tokens, d, d, BT_B, 16))) return rc;
  if ((rc = make_tmap_box(&tdq, dq_acc, tokens, d, d, 4, 32, 16))) return rc;
  const int nblk = (n + BT_B - 1) / BT_B;
  const int num_items = batch * heads * nblk;
  const int grid = num_items < num_sms() ? num_items : num_sms();
  sdpa_bwd_tc_kernel<<<grid, BT_THREADS, BT_SMEM_BYTES, st>>>(tq, to, tdq, lse, Dv, dqkv, n, heads, nblk, num_items,
                                                              k_scale);
  FACT_LAUNCH_CHECK("sdpa_bwd_tc_kernel launch");
  *done = true;
  return FACT_OK;
}

}  // namespace fact
