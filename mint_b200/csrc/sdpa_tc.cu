// tcgen05 attention core for the FACT sequence lengths (N <= 384, head_dim 80); reference math:
// mint/core/base_models.py:76-86 (softmax(q.k^T * d_model^-0.5) . v, no mask, heads merged "b h n d -> b n (h d)").
//
// B200 mapping (persistent, one CTA per SM, a tile = one (batch, head, 128-query block)):
//   * the whole score block lives in TMEM: S[128 x 384] fp32 (384 columns) + O[128 x 80] fp32 (80 columns) <= 512;
//     row i of the tile is TMEM lane i, so the softmax is thread-local (no shuffles, no online rescaling);
//   * S = Q.K^T and O = P.V are tcgen05.mma (UMMA 128 x 128 x 16 and 128 x 80 x 16) issued by one thread; in precise
//     mode each product is hi.hi + lo.hi + hi.lo accumulated in the same TMEM tile;
//   * P is written back IN PLACE over S (32 fp32 columns -> 16 columns bf16 P_hi | 16 columns bf16 P_lo) with
//     tcgen05.st and consumed by the PV MMA as a TMEM A-operand -- probabilities never touch shared or global memory;
//   * Q/K/V head slices arrive by TMA.  WIDE = 0: five [rows][16]-element sub-tiles per operand part (SWIZZLE_32B;
//     80 = 5 x 16 so no padding); K sub-tiles are K-major B operands, V sub-tiles MN-major B operands of the PV product.
//     WIDE >= 1: a Q / K part is ONE [rows][64] box (SWIZZLE_128B, 128-byte rows: a quarter of the TMA row requests and
//     conflict-free operand reads) + one [rows][16] box (SWIZZLE_32B) for head dims 64..79 -- four K steps read the wide
//     box, the fifth the narrow one;
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = softmax + output (TMEM lane quadrant = warp % 4, two
//     warps per quadrant split the score columns and exchange row max / sum through shared memory).
// Keys beyond N inside the last 128-key block are other rows of the qkv buffer (or TMA zero fill): their scores are
// never read and their probabilities are written as exact zeros.
#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {

constexpr int TC_QB = 128;           // query rows per tile (UMMA M)
constexpr int TC_KB = 128;           // keys per block (UMMA N of the score MMA)
constexpr int TC_MAXBLK = 3;         // N <= 384
constexpr int TC_DH = 80;
constexpr int TC_KS = TC_DH / 16;    // 16-wide head_dim slices
constexpr int TC_SUB = TC_KB * 32;   // bytes of one [128 rows][16 el] sub-tile
constexpr int TC_KV_STAGES = 4;
constexpr int TC_SM_WARPS = 8;        // softmax / output warps: two per TMEM lane quadrant, splitting the columns
constexpr int TC_THREADS = 64 + TC_SM_WARPS * 32;
constexpr int TC_O_COL = TC_MAXBLK * TC_KB;  // 384
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_OC = TC_DH / 2;                  // output columns per softmax warp (its half of the head)
constexpr int TC_OUT_BYTES = 32 * TC_OC * 2;      // 2560: one [32 rows][40 el] bf16 box (80-byte rows: conflict-free)

template <int NPART>
struct SdpaTcCfg {
  static constexpr int PART_BYTES = TC_KS * TC_SUB;          // 20 KB: one 128 x 80 operand part
  static constexpr int TILE_BYTES = NPART * PART_BYTES;      // Q tile or one K / V block
  static constexpr int BAR_OFF = (1 + TC_KV_STAGES) * TILE_BYTES;
  static constexpr int XCHG_OFF = BAR_OFF + 256;             // 2 x [2 halves][128 rows] floats: row max, row sum
  static constexpr int OUT_OFF = XCHG_OFF + 2 * 2 * TC_QB * 4;  // output staging: per softmax warp [32 rows][40 el] bf16
  static constexpr int SMEM_BYTES = 1024 + OUT_OFF + TC_SM_WARPS * TC_OUT_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
};

constexpr int TC_WIDE_BYTES = TC_KB * 128;  // the [128 rows][64 el] SWIZZLE_128B box of a wide operand part (16 KB)

template <int NPART, int WIDE>
__global__ void __launch_bounds__(TC_THREADS, 1)
sdpa_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const __grid_constant__ CUtensorMap tw_hi, const __grid_constant__ CUtensorMap tw_lo,
               const __grid_constant__ CUtensorMap to_hi, const __grid_constant__ CUtensorMap to_lo,
               float* __restrict__ lse, int N, int H, int q_blocks, int num_tiles) {
  using Cfg = SdpaTcCfg<NPART>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base;
  auto kv_smem = [&](int s) { return smem_base + (1 + s) * Cfg::TILE_BYTES; };
  const uint32_t bar_base = smem_base + Cfg::BAR_OFF;
  // barriers: kv_full[4] kv_empty[4] q_full q_empty s_full[3] p_full[3] o_full o_empty | tmem_ptr
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (TC_KV_STAGES + s); };
  const uint32_t q_full = bar_base + 8u * (2 * TC_KV_STAGES);
  const uint32_t q_empty = q_full + 8;
  auto s_full = [&](int c) { return q_full + 16u + 8u * c; };
  auto p_full = [&](int c) { return q_full + 16u + 8u * (TC_MAXBLK + c); };
  const uint32_t o_full = q_full + 16u + 8u * (2 * TC_MAXBLK);
  const uint32_t o_empty = o_full + 8;
  const uint32_t tmem_ptr_addr = o_empty + 8;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + Cfg::BAR_OFF + 8 * (2 * TC_KV_STAGES) + 16 + 8 * (2 * TC_MAXBLK) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * TC_DH;
  const int nblk = (N + TC_KB - 1) / TC_KB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    if (NPART == 2) tma_prefetch_desc(&tm_lo);
    if (WIDE > 0) {
      tma_prefetch_desc(&tw_hi);
      if (NPART == 2) tma_prefetch_desc(&tw_lo);
    }
    tma_prefetch_desc(&to_hi);
    if (NPART == 2) tma_prefetch_desc(&to_lo);
    for (int s = 0; s < TC_KV_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int c = 0; c < TC_MAXBLK; ++c) {
      mbar_init(s_full(c), 1);
      mbar_init(p_full(c), TC_SM_WARPS);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, TC_SM_WARPS);
    fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) tmem_alloc<TC_TMEM_COLS>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();   // q / k / v come from the QKV GEMM before this launch; o_hi / o_lo may still be read by the one before

  auto tile_coords = [&](int tile, int& b, int& h, int& qb) {
    qb = tile % q_blocks;
    const int bh = tile / q_blocks;
    h = bh % H;
    b = bh / H;
  };

  // Key / value blocks travel through the shared-memory ring per tile as K_0..K_n, V_0..V_n.  (Issuing the score MMA
  // of block c of tile t + 1 right behind the PV MMA of block c of tile t -- it overwrites P_c, and MMAs of one thread
  // execute in order -- was built and measured: 234 us instead of 223 at batch 128.  The softmax warps are what a
  // tile waits for, and score MMAs that run during their pass compete with them for TMEM; in this order the score MMAs
  // of the next tile overlap the output epilogue instead.)
  const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  auto tile_of = [&](int t) { return static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x); };

  if (warp == 0) {
    {  // ------------------------------------------------------------ TMA producer (whole warp, elected lane issues)
      uint32_t it = 0;
      auto load_block = [&](uint32_t dst, uint32_t bar, int col, int row, bool wide) {
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, Cfg::TILE_BYTES);
          if (wide) {  // [rows][64] SWIZZLE_128B box + [rows][16] SWIZZLE_32B box per part (same 20 KB)
            tma_load_2d(dst, &tw_hi, col, row, bar);
            tma_load_2d(dst + TC_WIDE_BYTES, &tm_hi, col + 64, row, bar);
            if (NPART == 2) {
              tma_load_2d(dst + Cfg::PART_BYTES, &tw_lo, col, row, bar);
              tma_load_2d(dst + Cfg::PART_BYTES + TC_WIDE_BYTES, &tm_lo, col + 64, row, bar);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < TC_KS; ++ks) {
              tma_load_2d(dst + ks * TC_SUB, &tm_hi, col + ks * 16, row, bar);
              if (NPART == 2) tma_load_2d(dst + Cfg::PART_BYTES + ks * TC_SUB, &tm_lo, col + ks * 16, row, bar);
            }
          }
        }
        __syncwarp();
      };
      auto load_q = [&](int t) {
        int b, h, qb;
        tile_coords(tile_of(t), b, h, qb);
        mbar_wait(q_empty, (t & 1) ^ 1);
        load_block(q_smem, q_full, h * TC_DH, b * N + qb * TC_QB, WIDE >= 1);
      };
      auto load_kv = [&](int t, int which, int c) {  // which: 1 = K, 2 = V
        int b, h, qb;
        tile_coords(tile_of(t), b, h, qb);
        const int s = it % TC_KV_STAGES;
        mbar_wait(kv_empty(s), ((it / TC_KV_STAGES) & 1) ^ 1);
        load_block(kv_smem(s), kv_full(s), which * D + h * TC_DH, b * N + c * TC_KB, which == 1 && WIDE >= 1);
        ++it;
      };
      for (int t = 0; t < my_tiles; ++t) {
        load_q(t);
        for (int which = 1; which <= 2; ++which)      // all K blocks, then all V blocks
          for (int c = 0; c < nblk; ++c) load_kv(t, which, c);
      }
    }
  } else if (warp == 1) {
    {  // ------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32_ex(TC_QB, TC_KB, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16_f32_ex(TC_QB, TC_DH, 1);
      uint32_t it = 0;
      // ---- S[:, 128c : 128c+128] = Q . K_c^T of tile t
      auto issue_s = [&](int t, int c) {
        const int s = it % TC_KV_STAGES;
        if (c == 0) mbar_wait(q_full, t & 1);
        mbar_wait(kv_full(s), (it / TC_KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + c * TC_KB;
        if (elect_one()) {
          const uint32_t ql_s = q_smem + (NPART - 1) * Cfg::PART_BYTES;
          const uint32_t kh_s = kv_smem(s), kl_s = kv_smem(s) + (NPART - 1) * Cfg::PART_BYTES;
          // WIDE: K steps 0..3 walk the 128-byte rows of the wide box (32 bytes per step), step 4 is the narrow box
          const uint64_t qh0 = WIDE ? umma_desc_k_sw128(q_smem) : umma_desc_k_sw32(q_smem);
          const uint64_t kh0 = WIDE ? umma_desc_k_sw128(kh_s) : umma_desc_k_sw32(kh_s);
          const uint64_t ql0 = WIDE ? umma_desc_k_sw128(ql_s) : umma_desc_k_sw32(ql_s);
          const uint64_t kl0 = WIDE ? umma_desc_k_sw128(kl_s) : umma_desc_k_sw32(kl_s);
          const uint64_t qh4 = umma_desc_k_sw32(q_smem + TC_WIDE_BYTES), kh4 = umma_desc_k_sw32(kh_s + TC_WIDE_BYTES);
          const uint64_t ql4 = umma_desc_k_sw32(ql_s + TC_WIDE_BYTES), kl4 = umma_desc_k_sw32(kl_s + TC_WIDE_BYTES);
#pragma unroll
          for (int ks = 0; ks < TC_KS; ++ks) {
            const uint64_t off = WIDE ? static_cast<uint64_t>(ks * 32) >> 4 : static_cast<uint64_t>(ks * TC_SUB) >> 4;
            const bool tail = WIDE && ks == TC_KS - 1;
            const uint64_t aqh = tail ? qh4 : qh0 + off, akh = tail ? kh4 : kh0 + off;
            const uint64_t aql = tail ? ql4 : ql0 + off, akl = tail ? kl4 : kl0 + off;
            umma_bf16(d, aqh, akh, idesc_s, ks > 0 ? 1u : 0u);
            if (NPART == 2) {
              umma_bf16(d, aql, akh, idesc_s, 1u);
              umma_bf16(d, aqh, akl, idesc_s, 1u);
            }
          }
          umma_commit(kv_empty(s));
          umma_commit(s_full(c));
          if (c == nblk - 1) umma_commit(q_empty);  // Q tile reusable once every score MMA has read it
        }
        __syncwarp();
        ++it;
      };
      // ---- O (+)= P_c . V_c of tile t   (P from TMEM, V MN-major from shared)
      auto issue_pv = [&](int t, int c) {
        const uint32_t tph = t & 1;
        const int s = it % TC_KV_STAGES;
        mbar_wait(kv_full(s), (it / TC_KV_STAGES) & 1);
        mbar_wait(p_full(c), tph);
        if (c == 0) mbar_wait(o_empty, tph ^ 1);  // previous tile's output has been read out of TMEM
        tc_fence_after();
        const int nvalid = min(TC_KB, N - c * TC_KB);
        const int ksteps = (nvalid + 15) >> 4;
        if (elect_one()) {
          const uint32_t vh_s = kv_smem(s), vl_s = kv_smem(s) + (NPART - 1) * Cfg::PART_BYTES;
          const uint64_t vh0 = umma_desc_mn_sw32(vh_s, TC_SUB, 256);
          const uint64_t vl0 = umma_desc_mn_sw32(vl_s, TC_SUB, 256);
#pragma unroll
          for (int j = 0; j < TC_KB / 16; ++j) {
            if (j < ksteps) {
              const uint32_t a_hi = tmem_base + c * TC_KB + 32 * (j >> 1) + 8 * (j & 1);
              const uint64_t off = static_cast<uint64_t>(j * 512) >> 4;
              umma_bf16_ts(tmem_base + TC_O_COL, a_hi, vh0 + off, idesc_o, (c > 0 || j > 0) ? 1u : 0u);
              if (NPART == 2) {
                umma_bf16_ts(tmem_base + TC_O_COL, a_hi + 16, vh0 + off, idesc_o, 1u);
                umma_bf16_ts(tmem_base + TC_O_COL, a_hi, vl0 + off, idesc_o, 1u);
              }
            }
          }
          umma_commit(kv_empty(s));
          if (c == nblk - 1) umma_commit(o_full);
        }
        __syncwarp();
        ++it;
      };
      for (int t = 0; t < my_tiles; ++t) {
        for (int c = 0; c < nblk; ++c) issue_s(t, c);
        for (int c = 0; c < nblk; ++c) issue_pv(t, c);
      }
    }
  } else {  // ------------------------------------------------------------------------ softmax + output warps 2..9
    // Two warps share each TMEM lane quadrant (rows); they split the 32-column groups of every key block by parity
    // and exchange their partial row max / row sum through shared memory with a 64-thread named barrier.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* xmax = reinterpret_cast<float*>(smem_gen + Cfg::XCHG_OFF);   // [2][128]
    float* xsum = xmax + 2 * TC_QB;                                     // [2][128]
    const int rt = q * 32 + lane;                                       // row inside the tile
    const uint32_t stg = smem_base + Cfg::OUT_OFF + static_cast<uint32_t>(warp - 2) * TC_OUT_BYTES;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); };
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const uint32_t tph = t & 1;
      int b, h, qb;
      tile_coords(tile, b, h, qb);
      const int row = qb * TC_QB + rt;
      // ---- pass 1: row max over this warp's column groups (scores are already in the exp2 domain)
      float mx = -INFINITY;
      for (int c = 0; c < nblk; ++c) {
        mbar_wait(s_full(c), tph);
        tc_fence_after();
        const int nvalid = min(TC_KB, N - c * TC_KB);
        for (int g = half; g * 32 < nvalid; g += 2) {
          float v[32];
          tmem_ld_32x32(lane_base + c * TC_KB + g * 32, v);
          tmem_ld_wait();
          const int lim = nvalid - g * 32;
          if (lim >= 32) {  // only the last group of the last key block is partial
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i < lim ? v[i] : -INFINITY);
          }
        }
      }
      xmax[half * TC_QB + rt] = mx;
      pair_sync();
      mx = fmaxf(mx, xmax[(half ^ 1) * TC_QB + rt]);
      // ---- pass 2: p = 2^(s - max); P_hi | P_lo overwrite the score columns they came from.
      // (Keeping half of the scores in registers between the passes -- three of a thread's six 32-column groups, 96
      // registers -- was measured: 262 us instead of 223 at batch 128.  The second TMEM read is not what the pass waits
      // for.)
      float lsum = 0.f;
      for (int c = 0; c < nblk; ++c) {
        const int nvalid = min(TC_KB, N - c * TC_KB);
        for (int g = half; g * 32 < nvalid; g += 2) {
          float v[32];
          const uint32_t addr = lane_base + c * TC_KB + g * 32;
          tmem_ld_32x32(addr, v);
          tmem_ld_wait();
          const int lim = nvalid - g * 32;
          const bool full = lim >= 32;
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = (full || 2 * i < lim) ? ex2_approx(v[2 * i] - mx) : 0.f;
            const float p1 = (full || 2 * i + 1 < lim) ? ex2_approx(v[2 * i + 1] - mx) : 0.f;
            lsum += p0 + p1;
            if (NPART == 2) {
              // hi = the upper 16 bits of p (p >= 0: truncation), lo = bf16(p - hi): hi + lo still carries 16 mantissa
              // bits and the pass needs one conversion per pair instead of two (the packing is a byte permute)
              const uint32_t u0 = __float_as_uint(p0), u1 = __float_as_uint(p1);
              pk[i] = __byte_perm(u0, u1, 0x7632);
              const float h0 = __uint_as_float(u0 & 0xffff0000u), h1 = __uint_as_float(u1 & 0xffff0000u);
              pk[16 + i] = cvt_bf16x2(p0 - h0, p1 - h1);
            } else {
              pk[i] = cvt_bf16x2(p0, p1);
            }
          }
          if (NPART == 2) tmem_st_32x32(addr, pk);
          else tmem_st_32x16(addr, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(c));
      }
      xsum[half * TC_QB + rt] = lsum;
      pair_sync();
      lsum += xsum[(half ^ 1) * TC_QB + rt];
      // ---- output: O / l, split, "b h n d -> b n (h d)"; each warp of a pair takes 40 of the head's 80 columns.
      // The rows leave through shared memory and the TMA engine (one [32 rows][40 columns] bulk tensor store per part;
      // the {columns, rows, clips} map clips the rows past the end of the clip).  Row-per-lane 16-byte global stores
      // -- 32 separate lines per instruction -- kept these warps waiting on the LSU for a fifth of their time (ncu).
      const float inv = 1.f / lsum;
      mbar_wait(o_full, tph);
      tc_fence_after();
      float ov[TC_OC];
      tmem_ld_32x32(lane_base + TC_O_COL + half * TC_OC, ov);
      tmem_ld_32x8(lane_base + TC_O_COL + half * TC_OC + 32, ov + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (lse && half == 0 && row < N) lse[(static_cast<size_t>(b) * H + h) * N + row] = mx + log2f(lsum);
      uint32_t hh[TC_OC / 2];
#pragma unroll
      for (int i = 0; i < TC_OC / 2; ++i) {
        ov[2 * i] *= inv;
        ov[2 * i + 1] *= inv;
        hh[i] = cvt_bf16x2(ov[2 * i], ov[2 * i + 1]);
      }
      const uint32_t srow = stg + lane * (TC_OC * 2);
      const int c0 = h * TC_DH + half * TC_OC, r0 = qb * TC_QB + q * 32;
      if (lane == 0) bulk_wait_group_read0();  // the previous tile's boxes have left the staging slot
      __syncwarp();
#pragma unroll
      for (int i = 0; i < TC_OC / 8; ++i) st_shared_v4(srow + i * 16, hh[4 * i], hh[4 * i + 1], hh[4 * i + 2], hh[4 * i + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && r0 < N) {
        tma_store_3d(&to_hi, stg, c0, r0, b);
        bulk_commit_group();
      }
      if (NPART == 2) {
        // the lo halves are computed while the engine reads the hi box out of the staging slot
        uint32_t ll[TC_OC / 2];
#pragma unroll
        for (int i = 0; i < TC_OC / 2; ++i)
          ll[i] = cvt_bf16x2(ov[2 * i] - __uint_as_float(hh[i] << 16), ov[2 * i + 1] - __uint_as_float(hh[i] & 0xffff0000u));
        if (lane == 0) bulk_wait_group_read0();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < TC_OC / 8; ++i)
          st_shared_v4(srow + i * 16, ll[4 * i], ll[4 * i + 1], ll[4 * i + 2], ll[4 * i + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && r0 < N) {
          tma_store_3d(&to_lo, stg, c0, r0, b);
          bulk_commit_group();
        }
      }
    }
    if (lane == 0) bulk_wait_group0();  // all boxes written before the CTA (and its shared memory) goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TC_TMEM_COLS>(tmem_base);
  }
}

// fact_set_flag("sdpa_wide", v): 0 = five 16-column SWIZZLE_32B sub-tiles per operand part, 1 = Q / K as a 64-column
// SWIZZLE_128B box + a 16-column box.  (V the same way, with the PV product split into an N = 64 and an N = 16 MMA, was
// measured too: 220 vs 221 us at batch 128 -- not kept.)
int g_sdpa_wide = 1;
template <int NPART, int WIDE>
static int launch_sdpa_tc(const bf16* qh, const bf16* ql, bf16* oh, bf16* ol, float* lse, int batch, int n,
                          int heads, int q_rows, cudaStream_t st) {
  using Cfg = SdpaTcCfg<NPART>;
  auto kern = sdpa_tc_kernel<NPART, WIDE>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int d = heads * TC_DH;
  CUtensorMap tmh, tml, twh, twl;
  int rc;
  if ((rc = make_tmap_bf16(&tmh, qh, batch * n, 3 * d, 3 * d, TC_KB, 16))) return rc;
  if (NPART == 2) {
    if ((rc = make_tmap_bf16(&tml, ql, batch * n, 3 * d, 3 * d, TC_KB, 16))) return rc;
  } else {
    tml = tmh;
  }
  twh = tmh;
  twl = tml;
  if (WIDE > 0) {
    if ((rc = make_tmap_bf16(&twh, qh, batch * n, 3 * d, 3 * d, TC_KB, 64))) return rc;
    twl = twh;
    if (NPART == 2 && (rc = make_tmap_bf16(&twl, ql, batch * n, 3 * d, 3 * d, TC_KB, 64))) return rc;
  }
  CUtensorMap toh, tol;
  if ((rc = make_tmap_rows3d(&toh, oh, batch, n, d, d, 32, TC_OC))) return rc;
  tol = toh;
  if (NPART == 2 && (rc = make_tmap_rows3d(&tol, ol, batch, n, d, d, 32, TC_OC))) return rc;
  const int q_blocks = (q_rows + TC_QB - 1) / TC_QB;
  const int num_tiles = q_blocks * heads * batch;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  FACT_CUDA_CHECK(launch_k(kern, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, true, tmh, tml, twh, twl, toh, tol, lse,
                           n, heads, q_blocks, num_tiles));
  FACT_LAUNCH_CHECK("sdpa_tc_kernel launch");
  return FACT_OK;
}

// tcgen05 path: head_dim 80, n <= 384, 16-byte aligned buffers.  Returns FACT_ERR_UNSUPPORTED (without setting an
// error) when the shape is outside that envelope so that fact_sdpa falls back to the generic mma.sync kernel.
int sdpa_tc_try(const bf16* qh, const bf16* ql, bf16* oh, bf16* ol, float* lse, int batch, int n, int heads,
                int head_dim, int q_rows, cudaStream_t st) {
  if (head_dim != TC_DH || n > TC_MAXBLK * TC_KB) return FACT_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qh) | reinterpret_cast<uintptr_t>(oh) | reinterpret_cast<uintptr_t>(ql) |
       reinterpret_cast<uintptr_t>(ol)) & 15)
    return FACT_ERR_UNSUPPORTED;
#define FACT_SDPA_TC(W)                                                                   \
  return ql ? launch_sdpa_tc<2, W>(qh, ql, oh, ol, lse, batch, n, heads, q_rows, st)      \
            : launch_sdpa_tc<1, W>(qh, ql, oh, ol, lse, batch, n, heads, q_rows, st)
  if (g_sdpa_wide) { FACT_SDPA_TC(1); }
  FACT_SDPA_TC(0);
#undef FACT_SDPA_TC
}

}  // namespace fact
