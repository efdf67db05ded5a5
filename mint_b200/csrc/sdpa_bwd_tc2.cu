// tcgen05 attention-core backward, software-pipelined (second generation of sdpa_bwd_tc.cu; same algebra, same
// reference: the gradient of softmax(q.k^T * d_model^-0.5) . v, mint/core/base_models.py:76-86, as tape.gradient
// produces it in mint/ctl/single_task_trainer.py:176-178).
//
// sdpa_bwd_tc.cu walks 128-query blocks with ONE score buffer in TMEM, so every step is a chain
//   S^T / dP^T MMAs -> softmax warps -> dV / dK / dQ MMAs -> dQ read-out -> next step
// and the tensor pipe is busy 18 % of the time (ncu, profiles/r1_ncu_full_summary_final.txt).  Here the inner step is
// 64 queries, which halves the score columns and lets TWO score buffers live in TMEM next to the accumulators:
//   TMEM columns: buffer b: S^T / P^T [128 b, 128 b + 64)  dP^T / dZ^T [128 b + 64, 128 b + 128)   (b = 0, 1)
//                 dV [256, 336)   dK [336, 416)   dQ [416, 496)
// so the issuer runs the score MMAs of step i + 1 while the softmax warps turn step i into P^T / dZ^T, and two
// dedicated warps drain dQ of step i out of TMEM (-> shared memory -> bulk reduction) while the softmax warps are
// already on step i + 1: in steady state no role waits for another.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = softmax (TMEM lane quadrant = warp % 4, two warps per quadrant
// split the 64 query columns), 12 / 13 = dQ read-out (quadrants 0 / 1 = the 64 query lanes), 10 / 11 idle.
//
//   work item = (batch, head, 128-key block j); for every 64-query block i:
//     S^T  = K_j . Q_i^T   [128 keys x 64 queries]     dP^T = V_j . dO_i^T
//     P^T  = 2^(S^T - lse[q]),  dZ^T = P^T * (dP^T - D[q])      softmax warps, TMEM lane = key
//     dV_j += P^T . dO_i    dK_j += dZ^T . Q_i                  P^T / dZ^T stay in TMEM as bf16 A operands
//     dQ_i  = dZ . K_j      A = dZ^T staged in shared memory (MN-major SWIZZLE_128B, one 64-query atom; the M = 128
//                           MMA reads the atom twice, TMEM lanes 64..127 are never read)
//   dV_j / dK_j leave as bf16 rows of dqkv after the last query block; each dQ_i partial goes through swizzled shared
//   memory and a bulk tensor reduction (fp32 add in the L2) into dq_acc.
// Q_i, dO_i and the 64 lse / D values of a step travel together in one shared-memory stage (TMA + two 1-D bulk copies).
#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {

constexpr int B2_KB = 128;                    // keys per work item
constexpr int B2_QB = 64;                     // queries per inner step
constexpr int B2_DH = 80;
constexpr int B2_KS = B2_DH / 16;             // 16-wide head_dim slices
constexpr int B2_SUBK = B2_KB * 32;           // bytes of one [128 rows][16 el] sub-tile
constexpr int B2_SUBQ = B2_QB * 32;           // bytes of one [64 rows][16 el] sub-tile
constexpr int B2_OPK = B2_KS * B2_SUBK;       // 20 KB: K_j or V_j
constexpr int B2_OPQ = B2_KS * B2_SUBQ;       // 10 KB: Q_i or dO_i
constexpr int B2_QSTAGES = 4;
constexpr int B2_STAGE = 21 * 1024;           // Q_i | dO_i | lse[64] | D[64] (+ pad to a 1024-byte multiple)
constexpr int B2_SM_WARPS = 8;
constexpr int B2_RO_WARP0 = 12;                // read-out warps 12 / 13 own TMEM lane quadrants 0 / 1 (warp % 4)
constexpr int B2_THREADS = 14 * 32;
constexpr int B2_COL_DV = 256, B2_COL_DK = 336, B2_COL_DQ = 416;

constexpr int B2_OFF_K = 0;
constexpr int B2_OFF_V = B2_OFF_K + B2_OPK;
constexpr int B2_OFF_QO = B2_OFF_V + B2_OPK;
constexpr int B2_OFF_DS = B2_OFF_QO + B2_QSTAGES * B2_STAGE;     // 2 x dZ^T atom [128 keys][128 B]
constexpr int B2_OFF_DQ = B2_OFF_DS + 2 * 16384;                 // dQ staging: 8 warp slots x 3 x [32 rows][64 B]
constexpr int B2_OFF_BAR = B2_OFF_DQ + B2_SM_WARPS * 3 * 2048;
constexpr int B2_SMEM_BYTES = 1024 + B2_OFF_BAR + 256;
static_assert(B2_SMEM_BYTES <= 232448, "exceeds 227 KB");
static_assert(B2_OFF_QO % 1024 == 0 && B2_OFF_DS % 1024 == 0 && B2_OFF_DQ % 1024 == 0, "swizzled tiles: 1024-B aligned");
static_assert(2 * B2_OPQ + 2 * B2_QB * 4 <= B2_STAGE, "stage too small");

__global__ void __launch_bounds__(B2_THREADS, 1)
sdpa_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                    const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_dq,
                    const float* __restrict__ lse, const float* __restrict__ Dv, bf16* __restrict__ dqkv, int N, int H,
                    int nkb, int nqb, int num_items, float k_scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t k_smem = smem_base + B2_OFF_K, v_smem = smem_base + B2_OFF_V;
  auto q_smem = [&](int s) { return smem_base + B2_OFF_QO + s * B2_STAGE; };
  auto do_smem = [&](int s) { return smem_base + B2_OFF_QO + s * B2_STAGE + B2_OPQ; };
  auto ld_off = [&](int s) { return B2_OFF_QO + s * B2_STAGE + 2 * B2_OPQ; };  // lse[64] then D[64] (generic offset)
  auto ds_smem = [&](int b) { return smem_base + B2_OFF_DS + b * 16384; };
  const uint32_t bar_base = smem_base + B2_OFF_BAR;
  const uint32_t kv_full = bar_base, kv_empty = bar_base + 8;
  auto qo_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto qo_empty = [&](int s) { return bar_base + 48u + 8u * s; };
  auto sdp_full = [&](int b) { return bar_base + 80u + 8u * b; };
  auto p_ready = [&](int b) { return bar_base + 96u + 8u * b; };
  const uint32_t dq_full = bar_base + 112, dq_empty = bar_base + 120, dkv_full = bar_base + 128;
  const uint32_t dkv_empty = bar_base + 136, tmem_ptr_addr = bar_base + 144;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + B2_OFF_BAR + 144);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * B2_DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int s = 0; s < B2_QSTAGES; ++s) {
      mbar_init(qo_full(s), 1);
      mbar_init(qo_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sdp_full(b), 1);
      mbar_init(p_ready(b), B2_SM_WARPS);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 2);  // the two read-out warps (TMEM lanes 0..63 = the 64 queries of a step)
    mbar_init(dkv_full, 1);
    mbar_init(dkv_empty, B2_SM_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto item_coords = [&](int item, int& b, int& h, int& jb) {
    jb = item % nkb;
    const int bh = item / nkb;
    h = bh % H;
    b = bh / H;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t it = 0, t = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
      int b, h, jb;
      item_coords(item, b, h, jb);
      const int row0 = b * N;
      mbar_wait(kv_empty, (t & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * B2_OPK);
#pragma unroll
        for (int ks = 0; ks < B2_KS; ++ks) {
          tma_load_2d(k_smem + ks * B2_SUBK, &tm_kv, D + h * B2_DH + ks * 16, row0 + jb * B2_KB, kv_full);
          tma_load_2d(v_smem + ks * B2_SUBK, &tm_kv, 2 * D + h * B2_DH + ks * 16, row0 + jb * B2_KB, kv_full);
        }
      }
      __syncwarp();
      for (int ib = 0; ib < nqb; ++ib, ++it) {
        const int s = it % B2_QSTAGES;
        mbar_wait(qo_empty(s), ((it / B2_QSTAGES) & 1) ^ 1);
        if (elect_one()) {
          const int i0 = ib * B2_QB;
          const uint32_t nval = static_cast<uint32_t>(min(B2_QB, N - i0)) * 4u;  // bytes of lse / D (N % 4 == 0)
          mbar_arrive_expect_tx(qo_full(s), 2 * B2_OPQ + 2 * nval);
#pragma unroll
          for (int ks = 0; ks < B2_KS; ++ks) {
            tma_load_2d(q_smem(s) + ks * B2_SUBQ, &tm_q, h * B2_DH + ks * 16, row0 + i0, qo_full(s));
            tma_load_2d(do_smem(s) + ks * B2_SUBQ, &tm_do, h * B2_DH + ks * 16, row0 + i0, qo_full(s));
          }
          const size_t idx = (static_cast<size_t>(b) * H + h) * N + i0;
          bulk_load_1d(smem_base + ld_off(s), lse + idx, nval, qo_full(s));
          bulk_load_1d(smem_base + ld_off(s) + B2_QB * 4, Dv + idx, nval, qo_full(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16_f32_ex(B2_KB, B2_QB, 0);        // [128 keys x 64 queries], K-major
    constexpr uint32_t idesc_ts = umma_idesc_bf16_f32_ex(B2_KB, B2_DH, 1);       // TMEM A, MN-major B
    constexpr uint32_t idesc_dq = umma_idesc_bf16_f32_maj(B2_KB, B2_DH, 1, 1);   // MN-major A and B
    uint32_t it = 0, t = 0;
    // score MMAs of global step `g` (buffer g & 1, stage g % QSTAGES)
    auto issue_sdp = [&](uint32_t g) {
      const int s = g % B2_QSTAGES;
      const uint32_t buf = g & 1;
      mbar_wait(qo_full(s), (g / B2_QSTAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kd = umma_desc_k_sw32(k_smem), vd = umma_desc_k_sw32(v_smem);
        const uint64_t qd = umma_desc_k_sw32(q_smem(s)), od = umma_desc_k_sw32(do_smem(s));
#pragma unroll
        for (int ks = 0; ks < B2_KS; ++ks)
          umma_bf16(tmem_base + buf * 128, kd + (static_cast<uint64_t>(ks * B2_SUBK) >> 4),
                    qd + (static_cast<uint64_t>(ks * B2_SUBQ) >> 4), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < B2_KS; ++ks)
          umma_bf16(tmem_base + buf * 128 + 64, vd + (static_cast<uint64_t>(ks * B2_SUBK) >> 4),
                    od + (static_cast<uint64_t>(ks * B2_SUBQ) >> 4), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(sdp_full(buf));
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
      int b, h, jb;
      item_coords(item, b, h, jb);
      const int ksteps_k = (min(B2_KB, N - jb * B2_KB) + 15) >> 4;
      mbar_wait(kv_full, t & 1);
      issue_sdp(it);
      for (int ib = 0; ib < nqb; ++ib, ++it) {
        const int s = it % B2_QSTAGES;
        const uint32_t buf = it & 1;
        const int ksteps_q = (min(B2_QB, N - ib * B2_QB) + 15) >> 4;
        if (ib + 1 < nqb) issue_sdp(it + 1);  // overlaps the softmax warps' pass over step `it`
        mbar_wait(p_ready(buf), (it >> 1) & 1);
        if (ib == 0) mbar_wait(dkv_empty, (t & 1) ^ 1);  // previous item's dV / dK have been read out
        tc_fence_after();
        if (elect_one()) {
          const uint64_t od_mn = umma_desc_mn_sw32(do_smem(s), B2_SUBQ, 256);
          const uint64_t qd_mn = umma_desc_mn_sw32(q_smem(s), B2_SUBQ, 256);
#pragma unroll
          for (int j = 0; j < B2_QB / 16; ++j)
            if (j < ksteps_q) {
              const uint32_t a_col = buf * 128 + 32 * (j >> 1) + 8 * (j & 1);  // 16 queries = 8 packed columns
              const uint64_t off = static_cast<uint64_t>(j * 512) >> 4;
              umma_bf16_ts(tmem_base + B2_COL_DV, tmem_base + a_col, od_mn + off, idesc_ts,
                           (ib > 0 || j > 0) ? 1u : 0u);
              umma_bf16_ts(tmem_base + B2_COL_DK, tmem_base + a_col + 64, qd_mn + off, idesc_ts,
                           (ib > 0 || j > 0) ? 1u : 0u);
            }
        }
        __syncwarp();
        mbar_wait(dq_empty, (it & 1) ^ 1);  // dQ of the previous step has been read out of TMEM
        tc_fence_after();
        if (elect_one()) {
          const uint64_t kd_mn = umma_desc_mn_sw32(k_smem, B2_SUBK, 256);
          const uint64_t zd = umma_desc_mn_sw128(ds_smem(buf), 0, 1024);  // LBO 0: rows 64..127 mirror rows 0..63
#pragma unroll
          for (int j = 0; j < B2_KB / 16; ++j)
            if (j < ksteps_k)
              umma_bf16(tmem_base + B2_COL_DQ, zd + (static_cast<uint64_t>(j * 2048) >> 4),
                        kd_mn + (static_cast<uint64_t>(j * 512) >> 4), idesc_dq, j > 0 ? 1u : 0u);
          umma_commit(qo_empty(s));
          umma_commit(dq_full);
          if (ib == nqb - 1) {
            umma_commit(dkv_full);
            umma_commit(kv_empty);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= B2_RO_WARP0) {
    // ------------------------------------------------------------------ dQ read-out warps 12, 13
    // dQ_i [64 queries x 80] sits in TMEM lanes 0..63: warp 12 owns lanes 0..31, warp 13 lanes 32..63 (a warp reaches
    // the lane quadrant warp % 4).  Each step: TMEM -> registers -> five swizzled [32 rows][16 col] fp32 boxes in shared
    // memory -> bulk tensor reduction into dq_acc; the softmax warps never wait for any of it.
    const int q4 = warp & 3;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    const uint32_t dq_stage = smem_base + B2_OFF_DQ + static_cast<uint32_t>(warp - B2_RO_WARP0) * B2_KS * 2048;
    uint32_t g = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int b, h, jb;
      item_coords(item, b, h, jb);
      const size_t tok0 = static_cast<size_t>(b) * N;
      for (int ib = 0; ib < nqb; ++ib, ++g) {
        const int q0 = ib * B2_QB;
        mbar_wait(dq_full, g & 1);
        tc_fence_after();
        float dq[B2_DH];
        tmem_ld_32x32(lane_base + B2_COL_DQ, dq);
        tmem_ld_32x32(lane_base + B2_COL_DQ + 32, dq + 32);
        tmem_ld_32x16(lane_base + B2_COL_DQ + 64, dq + 64);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(dq_empty);
          bulk_wait_group_read0();  // the previous step's boxes have left the staging slot
        }
        __syncwarp();
        const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
        for (int sb = 0; sb < B2_KS; ++sb) {
          const uint32_t rb = dq_stage + sb * 2048 + lane * 64;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            st_shared_v4(rb + ((c4 ^ sw) << 4), __float_as_uint(dq[sb * 16 + 4 * c4]),
                         __float_as_uint(dq[sb * 16 + 4 * c4 + 1]), __float_as_uint(dq[sb * 16 + 4 * c4 + 2]),
                         __float_as_uint(dq[sb * 16 + 4 * c4 + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && q0 + q4 * 32 < N) {  // rows past N hold exact zeros: skip wholly empty boxes
          const int row = static_cast<int>(tok0) + q0 + q4 * 32;
#pragma unroll
          for (int sb = 0; sb < B2_KS; ++sb) tma_reduce_add_2d(&tm_dq, dq_stage + sb * 2048, h * B2_DH + sb * 16, row);
          bulk_commit_group();
        }
      }
    }
    if (lane == 0) bulk_wait_group0();
  } else if (warp >= 2 && warp < 2 + B2_SM_WARPS) {
    // ------------------------------------------------------------------ softmax / epilogue warps 2..9
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    const int rl = q4 * 32 + lane;  // row (TMEM lane) inside the block: a key for S^T / dP^T / dV / dK, a query for dQ
    uint32_t it = 0, t = 0;

    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
      int b, h, jb;
      item_coords(item, b, h, jb);
      const size_t tok0 = static_cast<size_t>(b) * N;
      const int key = jb * B2_KB + rl;
      const bool key_ok = key < N;
      for (int ib = 0; ib < nqb; ++ib, ++it) {
        const int i0 = ib * B2_QB;
        const int s = it % B2_QSTAGES;
        const uint32_t buf = it & 1;
        mbar_wait(qo_full(s), (it / B2_QSTAGES) & 1);  // the stage's lse / D (async-proxy writes) are visible to us
        const float4* l4 = reinterpret_cast<const float4*>(smem_gen + ld_off(s)) + half * 8;
        const float4* d4 = reinterpret_cast<const float4*>(smem_gen + ld_off(s) + B2_QB * 4) + half * 8;
        mbar_wait(sdp_full(buf), (it >> 1) & 1);
        tc_fence_after();
        {
          float sc[32], dp[32];
          tmem_ld_32x32(lane_base + buf * 128 + half * 32, sc);
          tmem_ld_32x32(lane_base + buf * 128 + 64 + half * 32, dp);
          tmem_ld_wait();
          uint32_t pk[16], zk[16];
          const int qlim = N - i0 - half * 32;  // columns [0, qlim) of this thread's 32 are real queries
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 lv = l4[c], dv = d4[c];
            const float lq[4] = {lv.x, lv.y, lv.z, lv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
            float p[4], z[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = 4 * c + e;
              p[e] = (key_ok && col < qlim) ? ex2_approx(sc[col] - lq[e]) : 0.f;
              z[e] = (key_ok && col < qlim) ? p[e] * (dp[col] - dd[e]) : 0.f;
            }
            pk[2 * c] = cvt_bf16x2(p[0], p[1]);
            pk[2 * c + 1] = cvt_bf16x2(p[2], p[3]);
            zk[2 * c] = cvt_bf16x2(z[0], z[1]);
            zk[2 * c + 1] = cvt_bf16x2(z[2], z[3]);
          }
          tmem_st_32x16(lane_base + buf * 128 + half * 32, pk);
          tmem_st_32x16(lane_base + buf * 128 + 64 + half * 32, zk);
          // dZ^T row `rl` (key), queries half*32 .. +31 -> 16-byte chunks half*4 .. +3 of the 128-byte row, SWIZZLE_128B
          const uint32_t rbase = ds_smem(buf) + rl * 128;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            st_shared_v4(rbase + (((half * 4 + c4) ^ (rl & 7)) << 4), zk[4 * c4], zk[4 * c4 + 1], zk[4 * c4 + 2],
                         zk[4 * c4 + 3]);
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready(buf));
      }
      // ---- dV_j (half 0) / dK_j (half 1): TMEM lane = key
      mbar_wait(dkv_full, t & 1);
      tc_fence_after();
      {
        float o[B2_DH];
        const uint32_t src = lane_base + (half == 0 ? B2_COL_DV : B2_COL_DK);
        tmem_ld_32x32(src, o);
        tmem_ld_32x32(src + 32, o + 32);
        tmem_ld_32x16(src + 64, o + 64);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_empty);
        if (key_ok) {
          const float scl = half == 0 ? 1.f : k_scale;
          uint4* dst = reinterpret_cast<uint4*>(dqkv + (tok0 + key) * 3 * D + (half == 0 ? 2 * D : D) + h * B2_DH);
#pragma unroll
          for (int i = 0; i < B2_DH / 8; ++i)
            dst[i] = make_uint4(cvt_bf16x2(o[8 * i] * scl, o[8 * i + 1] * scl),
                                cvt_bf16x2(o[8 * i + 2] * scl, o[8 * i + 3] * scl),
                                cvt_bf16x2(o[8 * i + 4] * scl, o[8 * i + 5] * scl),
                                cvt_bf16x2(o[8 * i + 6] * scl, o[8 * i + 7] * scl));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Returns FACT_OK with *done = true when this kernel took the problem (dq_acc zeroed by the caller, D already
// computed); *done = false -> the caller tries the first-generation kernel.
int sdpa_bwd_tc2_try(const bf16* qkv, const bf16* d_o, const float* lse, const float* Dv, bf16* dqkv, float* dq_acc,
                     int batch, int n, int heads, int head_dim, float k_scale, cudaStream_t st, bool* done) {
  *done = false;
  if (head_dim != B2_DH || n > 3 * B2_KB || n % 4 != 0) return FACT_OK;  // lse / D rows move as 16-byte bulk copies
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(dqkv) |
       reinterpret_cast<uintptr_t>(dq_acc) | reinterpret_cast<uintptr_t>(lse) | reinterpret_cast<uintptr_t>(Dv)) & 15)
    return FACT_OK;
  static bool attr_done = false;
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(sdpa_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM_BYTES));
    attr_done = true;
  }
  const int d = heads * B2_DH;
  const int tokens = batch * n;
  CUtensorMap tkv, tq, to, tdq;
  int rc;
  if ((rc = make_tmap_bf16(&tkv, qkv, tokens, 3 * d, 3 * d, B2_KB, 16))) return rc;
  if ((rc = make_tmap_bf16(&tq, qkv, tokens, 3 * d, 3 * d, B2_QB, 16))) return rc;
  if ((rc = make_tmap_bf16(&to, d_o, tokens, d, d, B2_QB, 16))) return rc;
  if ((rc = make_tmap_box(&tdq, dq_acc, tokens, d, d, 4, 32, 16))) return rc;
  const int nkb = (n + B2_KB - 1) / B2_KB, nqb = (n + B2_QB - 1) / B2_QB;
  const int num_items = batch * heads * nkb;
  const int grid = num_items < num_sms() ? num_items : num_sms();
  sdpa_bwd_tc2_kernel<<<grid, B2_THREADS, B2_SMEM_BYTES, st>>>(tkv, tq, to, tdq, lse, Dv, dqkv, n, heads, nkb, nqb,
                                                               num_items, k_scale);
  FACT_LAUNCH_CHECK("sdpa_bwd_tc2_kernel launch");
  *done = true;
  return FACT_OK;
}

}  // namespace fact
