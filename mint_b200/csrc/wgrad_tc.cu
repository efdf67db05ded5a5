// Weight-gradient GEMM on CTA pairs: dW[in, out] += X^T . dY over the tokens (reference semantics: tape.gradient of the
// Dense kernels, mint/ctl/single_task_trainer.py:176-178).
//
// Tile = 256 (M side, fixed by cta_group::2) x NT (N side, any multiple of 16 up to 256).  FACT's Dense kernels are
// 800 wide on one side (d_model) and 800 / 2400 / 3072 on the other, and 800 = 3.125 x 256: with square 256-tiles a
// fifth of the MMAs multiplied zero padding.  So the N side is cut into EQUAL tiles just wide enough (800 -> 4 x 208,
// 2400 -> 10 x 240, 3072 -> 12 x 256), and when it is the `in` dimension that is 800 wide the product is computed
// transposed (M side = out, N side = in; the epilogue transposes its 32 x 32 boxes on the way to shared memory).
// Executed MMA work per layer drops by 19 % (dW1, dW2: 1.28 -> 1.04 x the useful flops; dWqkv 1.37 -> 1.11).
//
// The 1-SM kernel in backward.cu (128 x 128 tiles) needs 128 bytes of operands per MMA clock and is bound by the
// L2 -> SM path at a third of the tensor peak.  Here a cluster of two CTAs owns a 256 x 256 tile (cta_group::2: each CTA
// stages its own 128 `in` rows of X^T and 128 of the 256 `out` columns of dY), which halves the operand bytes per flop:
//   * both operands MN-major straight from the row-major activations through TMA ({64 MN, 64 tokens} SWIZZLE_128B boxes;
//     boxes past the matrix edge are zero-filled on chip and cost no L2 traffic),
//   * persistent clusters walk (tile, token-slice) work items; two 256-column TMEM accumulators, so the epilogue of one
//     item overlaps the main loop of the next,
//   * epilogue: each warp stages a 32 x 32 fp32 box in shared memory and the TMA engine adds it into dW in the L2
//     (cp.reduce.async.bulk .add.f32) -- no row-strided atomics through the LSU.
#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {

constexpr int W2_BK = 64, W2_STAGES = 6, W2_EPI_WARPS = 8, W2_THREADS = 64 + W2_EPI_WARPS * 32;
constexpr int W2_BOX_BYTES = 64 * 64 * 2;           // one {64 MN, 64 tokens} box
constexpr int W2_STAGE_BYTES = 4 * W2_BOX_BYTES;    // per CTA: A 128 rows (2 boxes) + B 128 columns (2 boxes)
constexpr int W2_STAGING_BYTES = W2_EPI_WARPS * 4096;
constexpr int W2_SMEM_BYTES = 1024 + W2_STAGES * W2_STAGE_BYTES + W2_STAGING_BYTES + 256;
static_assert(W2_SMEM_BYTES <= 232448, "exceeds 227 KB");

// Work item w -> (tile = w / splits, slice = w % splits); slice z covers token blocks [z * kb_per, (z + 1) * kb_per).
// tmX: A operand {M-side dim (inner), tokens}, tmY: B operand {N-side dim (inner), tokens} (X and dY, or dY and X when
// `swapped`), tmW: fp32 dW {out (inner), in}, box 32 x 32 SWIZZLE_128B.  MDIM / NDIM: extents of the M / N side; NT: N
// tile width (multiple of 16, <= 256; each CTA stages NT / 2 columns of the B operand).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(W2_THREADS, 1) gemm_wgrad2_kernel(
    const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
    const __grid_constant__ CUtensorMap tmW, int MDIM, int NDIM, int NT, int swapped, int tiles_out, int num_items,
    int splits, int num_kb, int kb_per, int tiles, int slice_major) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t staging_base = smem_base + W2_STAGES * W2_STAGE_BYTES;
  const uint32_t bar_base = staging_base + W2_STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (W2_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * W2_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * W2_STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * W2_STAGES + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + W2_STAGES * W2_STAGE_BYTES + W2_STAGING_BYTES + 8 * (2 * W2_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < W2_STAGES; ++s) {
      mbar_init(full_bar(s), 2);   // leader's arrive.expect_tx + peer's remote arrive
      mbar_init(empty_bar(s), 1);  // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 2 * W2_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  auto sA = [&](int s) { return smem_base + s * W2_STAGE_BYTES; };
  auto sB = [&](int s) { return smem_base + s * W2_STAGE_BYTES + 2 * W2_BOX_BYTES; };
  // slice_major: item = slice * tiles + tile -- the clusters that run side by side work on the SAME token slice of
  // different tiles, so the X / dY panels of that slice are fetched from HBM once and then hit in the L2 (each X block
  // is wanted by tiles_out tiles, each dY block by tiles / tiles_out).  Tile-major order (item = tile * splits + slice)
  // had concurrent clusters on different token ranges: 3.4x the algorithmic bytes came from DRAM (ncu, round 1).
  auto tile_of = [&](int item) { return slice_major ? item % tiles : item / splits; };
  auto item_kb = [&](int item, int& kb0) {
    kb0 = (slice_major ? item / tiles : item % splits) * kb_per;
    const int kb1 = min(kb0 + kb_per, num_kb);
    return kb1 - kb0;  // >= 1 by construction of splits / kb_per
  };

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs)
    uint32_t it = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int tile = tile_of(item);
      const int m0 = (tile / tiles_out) * 256 + rank * 128;       // this CTA's rows of the M side
      const int n0 = (tile % tiles_out) * NT + rank * (NT >> 1);  // this CTA's half of the tile's N-side columns
      int kb0;
      const int nkb = item_kb(item, kb0);
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % W2_STAGES;
        mbar_wait(empty_bar(s), ((it / W2_STAGES) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * W2_STAGE_BYTES);
          else mbar_arrive_leader(full_bar(s));
          const int t0 = (kb0 + i) * W2_BK;
          tma_load_2d_2sm(sA(s), &tmX, m0, t0, full_bar(s));
          tma_load_2d_2sm(sA(s) + W2_BOX_BYTES, &tmX, m0 + 64, t0, full_bar(s));
          tma_load_2d_2sm(sB(s), &tmY, n0, t0, full_bar(s));
          tma_load_2d_2sm(sB(s) + W2_BOX_BYTES, &tmY, n0 + 64, t0, full_bar(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {  // ---------------- MMA issuer
      const uint32_t idesc = umma_idesc_bf16_f32_maj(256, NT, 1, 1);
      uint32_t it = 0, t = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++t) {
        const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
        int kb0;
        const int nkb = item_kb(item, kb0);
        mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % W2_STAGES;
          mbar_wait(full_bar(s), (it / W2_STAGES) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a0 = umma_desc_mn_sw128(sA(s), W2_BOX_BYTES, 1024);
            const uint64_t b0 = umma_desc_mn_sw128(sB(s), W2_BOX_BYTES, 1024);
#pragma unroll
            for (int kk = 0; kk < W2_BK / 16; ++kk) {  // 16 tokens = two 8-row groups = 2048 bytes
              const uint64_t off = static_cast<uint64_t>(kk * 2048) >> 4;
              umma_bf16_2sm(tmem_d, a0 + off, b0 + off, idesc, (i > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit_2sm(empty_bar(s), 0x3);
            if (i == nkb - 1) umma_commit_2sm(tmem_full_bar(acc), 0x3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------- epilogue warps (both CTAs): this CTA's 128 M-side rows x NT N-side columns
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t stg = staging_base + static_cast<uint32_t>(warp - 2) * 4096;
    uint32_t t = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++t) {
      const int tile = tile_of(item);
      const int row0 = (tile / tiles_out) * 256 + rank * 128 + q * 32;
      const int n0 = (tile % tiles_out) * NT;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < 8; c += 2) {
        const int col0 = n0 + c * 32;
        const int valid = NT - c * 32;  // tile columns of this 32-wide chunk the MMA wrote (the rest is stale TMEM)
        if (valid <= 0 || col0 >= NDIM || row0 >= MDIM) continue;
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + c * 32, v);
        tmem_ld_wait();
        if (valid < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= valid) v[i] = 0.f;  // these columns belong to the next tile: add zero there
        }
        if (lane == 0) bulk_wait_group_read0();
        __syncwarp();
        if (!swapped) {
          // box[row = lane (in)][col = i (out)]: 16-byte chunks, swizzled like the tensor map expects
          const uint32_t rbase = stg + lane * 128, sw = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            st_shared_v4(rbase + ((i ^ sw) << 4), __float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                         __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
        } else {
          // transposed: this lane holds out = row0 + lane, in = col0 + i -> box[row = i (in)][col = lane (out)].
          // For a fixed i the 32 lanes write 32 distinct words of one 128-byte row: conflict-free scalar stores.
          const uint32_t cchunk = static_cast<uint32_t>(lane) >> 2, cword = (static_cast<uint32_t>(lane) & 3) << 2;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t addr = stg + i * 128 + ((cchunk ^ (i & 7)) << 4) + cword;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v[i]) : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (!swapped) tma_reduce_add_2d(&tmW, stg, col0, row0);   // {out, in} = {N side, M side}
          else tma_reduce_add_2d(&tmW, stg, row0, col0);            // {out, in} = {M side, N side}
          bulk_commit_group();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tmem_empty_bar(acc));
        else mbar_arrive_leader(tmem_empty_bar(acc));
      }
    }
    if (lane == 0) bulk_wait_group0();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// fact_set_flag("wgrad_pair", v): 0 keeps the 1-SM 128 x 128 kernel (A/B timing, tests), 1 = pair kernel with the cheaper
// orientation, 2 / 3 force the plain / transposed orientation (tests)
int g_wgrad_pair = 1;
// fact_set_flag("wgrad_order", 0): tile-major work items (A/B timing of the L2 reuse the slice-major order buys)
int g_wgrad_order = 1;

// returns FACT_OK and sets *done when the pair kernel took the problem; *done = false -> caller uses the 1-SM kernel
int wgrad_gemm_pair(const void* x_bf16, int ldx, const void* dy_bf16, int ldy, float* dW, int ldw, int tokens, int in_dim,
                    int out_dim, cudaStream_t st, bool* done) {
  *done = false;
  // bulk reductions need a 16-byte pitch; tiny or skinny problems stay on the 128 x 128 kernel
  if (!g_wgrad_pair || (static_cast<long long>(ldw) * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(dW) & 15) != 0 ||
      in_dim < 256 || out_dim < 256 || tokens < 64 * 64)
    return FACT_OK;
  static bool attr_done = false;
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(gemm_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W2_SMEM_BYTES));
    attr_done = true;
  }
  // Orientation: the M side is cut into 256-row tiles, the N side into equal tiles of up to 256 columns; pick the
  // orientation (and with it the padding) that executes fewer MMA columns.  swapped = M side is `out`.
  auto pad_m = [](int v) { return (v + 255) / 256 * 256; };
  auto tiles_n_of = [](int v) { return (v + 255) / 256; };
  auto nt_of = [&](int v) { return ((v + tiles_n_of(v) - 1) / tiles_n_of(v) + 15) / 16 * 16; };
  const long long cost_plain = static_cast<long long>(pad_m(in_dim)) * tiles_n_of(out_dim) * nt_of(out_dim);
  const long long cost_swap = static_cast<long long>(pad_m(out_dim)) * tiles_n_of(in_dim) * nt_of(in_dim);
  const int swapped = (g_wgrad_pair != 2 && cost_swap < cost_plain) || g_wgrad_pair == 3 ? 1 : 0;
  const int m_dim = swapped ? out_dim : in_dim, n_dim = swapped ? in_dim : out_dim;
  const int nt = nt_of(n_dim);
  CUtensorMap tmx, tmy, tmw;
  int rc;
  if ((rc = make_tmap_bf16(&tmx, x_bf16, tokens, in_dim, ldx, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tmy, dy_bf16, tokens, out_dim, ldy, 64, 64))) return rc;
  if ((rc = make_tmap_out(&tmw, dW, in_dim, out_dim, ldw, 4))) return rc;
  const int tiles_m = (m_dim + 255) / 256, tiles_out = tiles_n_of(n_dim);
  const int tiles = tiles_m * tiles_out;
  const int num_kb = (tokens + W2_BK - 1) / W2_BK;
  const int clusters_max = num_sms() / 2;
  // token slices: the split that fills whole rounds of the persistent clusters best, with >= 24 token blocks per item
  // so the 256 KB reduction per item stays a small fraction of the main loop
  int best_s = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 64 && num_kb / s >= 24; ++s) {
    const int kbp = (num_kb + s - 1) / s;
    const int s_eff = (num_kb + kbp - 1) / kbp;
    const int items = tiles * s_eff;
    const int rounds = (items + clusters_max - 1) / clusters_max;
    const double eff = static_cast<double>(items) / (static_cast<double>(rounds) * clusters_max);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_s = s_eff;
    }
  }
  const int kb_per = (num_kb + best_s - 1) / best_s;
  const int splits = (num_kb + kb_per - 1) / kb_per;
  const int num_items = tiles * splits;
  const int clusters = num_items < clusters_max ? num_items : clusters_max;
  gemm_wgrad2_kernel<<<2 * clusters, W2_THREADS, W2_SMEM_BYTES, st>>>(swapped ? tmy : tmx, swapped ? tmx : tmy, tmw, m_dim,
                                                                      n_dim, nt, swapped, tiles_out, num_items, splits,
                                                                      num_kb, kb_per, tiles, g_wgrad_order);
  FACT_LAUNCH_CHECK("gemm_wgrad2_kernel launch");
  *done = true;
  return FACT_OK;
}

}  // namespace fact
