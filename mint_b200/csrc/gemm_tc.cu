// tcgen05 GEMM for the FACT Dense layers (reference: tf.keras.layers.Dense at mint/core/base_models.py:51-53,
// 68-69, 176-180):  C[m,n] = A[m,k] . W[n,k]^T  with fused epilogues.
//
//  * operands are bf16, K-major, staged global -> shared by TMA (cp.async.bulk.tensor, SWIZZLE_128B, zero fill
//    out of bounds so K = 800 = 12.5 x 64 needs no padding);
//  * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1), fp32 accumulators live in TMEM;
//  * FACT_MODE_PRECISE ("bf16x3"): A and W each arrive as hi + lo bf16 halves; per K block the issuer runs
//    hi.hi + lo.hi + hi.lo into the same accumulator -> fp32-grade products at 3 MMAs per 2x operand bytes;
//  * warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2..5 epilogue
//    (tcgen05.ld -> shared transpose -> coalesced global stores with bias / GELU / residual / split fused).
//
// Pipeline: STAGES-deep ring of {A parts, B parts} with full/empty mbarriers; tcgen05.commit releases a stage
// when the MMAs that read it retire, and signals the epilogue after the last K block.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "fact_internal.h"
#include "fact_ptx.cuh"

namespace fact {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // one 128-byte swizzle span of bf16
constexpr int UMMA_K = 16;    // bf16
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quadrant
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int A_TILE_BYTES = BM * BK * 2;

struct EpiArgs {
  float* out_f32;
  bf16* out_hi;
  bf16* out_lo;
  int ldo;
  const float* bias;
  const float* resid;
  int ldr;
  float scale;
  int scale_cols;
  int seq_in, seq_out, seq_off;
  int vec_ok;  // all row pitches / base pointers allow 16-byte row-chunk accesses
  const bf16* aux;  // GELU_GRAD: saved pre-activation
  int ldaux;
  int reduce_add;  // TMA-store epilogue, BIAS_RESID_F32 with out == resid: the add is a bulk reduction in the L2
  float* colsum;   // TMA-store epilogue, GELU_GRAD: += column sums of the staged bf16 box (bias gradient)
  int resid_rows;  // > 0: resid is a [resid_rows, ldr] table read at row % resid_rows (the position embedding)
  const float* ln_gamma;  // split-K finish: also LayerNorm(out row) -> ln_hi / ln_lo (see fact_gemm_epilogue)
  const float* ln_beta;
  bf16* ln_hi;
  bf16* ln_lo;
  int* ln_sync;  // pair kernel with LayerNorm warps: arrival counters, one per 32-row group (zero between launches)
};

template <int BN, int NPART, int STAGES>
struct GemmCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = NPART * (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int ACC_COLS = 2 * BN;  // double-buffered accumulator
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128
                                   : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + BAR_BYTES;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "UMMA N constraint for M=128 / 32-column epilogue chunks");
  static_assert(ACC_COLS <= 512, "TMEM has 512 columns");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
  static_assert((2 * STAGES + 5) * 8 <= BAR_BYTES, "barrier block too small");
};

__device__ __forceinline__ size_t resid_row(int row, const EpiArgs& e) {
  return static_cast<size_t>(e.resid_rows > 0 ? row % e.resid_rows : row);
}
__device__ __forceinline__ int map_row(int row, const EpiArgs& e) {
  if (e.seq_in == 0) return row;
  return (row / e.seq_in) * e.seq_out + e.seq_off + (row % e.seq_in);
}

// One 32-column chunk of one accumulator row (this thread's row): fused epilogue + global store.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const float* v, const float4* rv, int row, int col0, bool row_ok,
                                               bool fast, int N, const EpiArgs& ep) {
  if (EPI == FACT_EPI_BIAS_RESID_F32 || EPI == FACT_EPI_BIAS_F32) {
    if (!row_ok) return;
    float* orow = ep.out_f32 + static_cast<size_t>(map_row(row, ep)) * ep.ldo + col0;
    if (fast) {
      const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b = ep.bias ? __ldg(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = v[4 * i + 0] + b.x;
        o.y = v[4 * i + 1] + b.y;
        o.z = v[4 * i + 2] + b.z;
        o.w = v[4 * i + 3] + b.w;
        if (EPI == FACT_EPI_BIAS_RESID_F32) {
          o.x += rv[i].x;
          o.y += rv[i].y;
          o.z += rv[i].z;
          o.w += rv[i].w;
        }
        reinterpret_cast<float4*>(orow)[i] = o;
      }
    } else {
      const float* rrow = ep.resid + resid_row(row, ep) * ep.ldr + col0;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (col0 + c < N) {
          float x = v[c] + (ep.bias ? ep.bias[col0 + c] : 0.f);
          if (EPI == FACT_EPI_BIAS_RESID_F32) x += rrow[c];
          orow[c] = x;
        }
      }
    }
  } else {
    if (!row_ok) return;
    const size_t o = static_cast<size_t>(row) * ep.ldo + col0;
    float x[32];
    if (EPI == FACT_EPI_BIAS_GELU_SAVE) {  // training forward: h = gelu(z) and z itself, both bf16 (hi only)
      uint32_t hz[16], zz[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float b0 = (fast || col0 + 2 * c < N) ? __ldg(ep.bias + col0 + 2 * c) : 0.f;
        const float b1 = (fast || col0 + 2 * c + 1 < N) ? __ldg(ep.bias + col0 + 2 * c + 1) : 0.f;
        const float z0 = v[2 * c] + b0, z1 = v[2 * c + 1] + b1;
        zz[c] = cvt_bf16x2(z0, z1);
        hz[c] = cvt_bf16x2(gelu_tanh(z0), gelu_tanh(z1));
      }
      if (fast) {
        uint4* ph = reinterpret_cast<uint4*>(ep.out_hi + o);
        uint4* pz = reinterpret_cast<uint4*>(ep.out_lo + o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ph[i] = make_uint4(hz[4 * i], hz[4 * i + 1], hz[4 * i + 2], hz[4 * i + 3]);
          pz[i] = make_uint4(zz[4 * i], zz[4 * i + 1], zz[4 * i + 2], zz[4 * i + 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (col0 + c < N) {
            const uint32_t hv = hz[c >> 1], zv = zz[c >> 1];
            ep.out_hi[o + c] = __ushort_as_bfloat16(static_cast<unsigned short>((c & 1) ? hv >> 16 : hv & 0xffff));
            ep.out_lo[o + c] = __ushort_as_bfloat16(static_cast<unsigned short>((c & 1) ? zv >> 16 : zv & 0xffff));
          }
      }
      return;
    }
    if (EPI == FACT_EPI_GELU_GRAD) {  // backward: dz = dh * gelu'(z), z read back as bf16
      const bf16* zrow = ep.aux + static_cast<size_t>(row) * ep.ldaux + col0;
      uint32_t zraw[16];
      if (fast) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 q4 = reinterpret_cast<const uint4*>(zrow)[i];
          zraw[4 * i] = q4.x; zraw[4 * i + 1] = q4.y; zraw[4 * i + 2] = q4.z; zraw[4 * i + 3] = q4.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint32_t a = (col0 + 2 * c < N) ? __bfloat16_as_ushort(zrow[2 * c]) : 0u;
          const uint32_t b = (col0 + 2 * c + 1 < N) ? __bfloat16_as_ushort(zrow[2 * c + 1]) : 0u;
          zraw[c] = a | (b << 16);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const uint32_t raw = zraw[c >> 1];
        const float z = __uint_as_float((c & 1) ? (raw & 0xffff0000u) : (raw << 16));
        x[c] = v[c] * gelu_tanh_grad(z);
      }
    } else if (EPI == FACT_EPI_BIAS_GELU_SPLIT) {
      if (fast) {  // bias as 8 x 16-byte loads (warp-uniform addresses), not 32 scalar ones
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(b4 + i);
          x[4 * i + 0] = gelu_tanh(v[4 * i + 0] + b.x);
          x[4 * i + 1] = gelu_tanh(v[4 * i + 1] + b.y);
          x[4 * i + 2] = gelu_tanh(v[4 * i + 2] + b.z);
          x[4 * i + 3] = gelu_tanh(v[4 * i + 3] + b.w);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = gelu_tanh(v[c] + (col0 + c < N ? __ldg(ep.bias + col0 + c) : 0.f));
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = (col0 + c < ep.scale_cols) ? v[c] * ep.scale : v[c];
    }
    if (fast) {
      uint32_t h[16], l[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {  // packed conversions: one cvt per pair for hi, one for lo
        const uint32_t hp = cvt_bf16x2(x[2 * c], x[2 * c + 1]);
        h[c] = hp;
        l[c] = cvt_bf16x2(x[2 * c] - __uint_as_float(hp << 16), x[2 * c + 1] - __uint_as_float(hp & 0xffff0000u));
      }
      uint4* ph = reinterpret_cast<uint4*>(ep.out_hi + o);
#pragma unroll
      for (int i = 0; i < 4; ++i) ph[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
      if (ep.out_lo) {
        uint4* pl = reinterpret_cast<uint4*>(ep.out_lo + o);
#pragma unroll
        for (int i = 0; i < 4; ++i) pl[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (col0 + c < N) {
          bf16 h0, l0;
          split_bf16(x[c], h0, l0);
          ep.out_hi[o + c] = h0;
          if (ep.out_lo) ep.out_lo[o + c] = l0;
        }
      }
    }
  }
}

// Persistent: grid = min(#tiles, #SMs); CTA c walks tiles c, c + grid, ... (n fastest, so co-resident CTAs share A
// rows in L2).  TMEM holds two BN-column accumulators: the MMA warp fills one while the epilogue warps drain the other.
template <int BN, int NPART, int STAGES, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(
    const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, int M, int N, int K,
    int tiles_n, int num_tiles, EpiArgs ep) {
  using Cfg = GemmCfg<BN, NPART, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier block: full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (NPART == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), EPI_WARPS);
    }
    fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto sA = [&](int s, int part) { return smem_base + s * Cfg::STAGE_BYTES + part * A_TILE_BYTES; };
  auto sB = [&](int s, int part) {
    return smem_base + s * Cfg::STAGE_BYTES + NPART * A_TILE_BYTES + part * Cfg::B_TILE_BYTES;
  };

  if (warp == 0) {
    // ---------------- TMA producer: whole warp walks the ring, one elected lane issues
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
          tma_load_2d(sA(s, 0), &tmA0, kb * BK, m0, full_bar(s));
          tma_load_2d(sB(s, 0), &tmB0, kb * BK, n0, full_bar(s));
          if (NPART == 2) {
            tma_load_2d(sA(s, 1), &tmA1, kb * BK, m0, full_bar(s));
            tma_load_2d(sB(s, 1), &tmB1, kb * BK, n0, full_bar(s));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp walks the pipeline, one elected lane issues
    constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
    uint32_t it = 0, t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          // descriptors differ only in the start-address field: build once per stage, bump by 32 B per K step
          const uint64_t a_hi0 = umma_desc_k_sw128(sA(s, 0)), b_hi0 = umma_desc_k_sw128(sB(s, 0));
          const uint64_t a_lo0 = umma_desc_k_sw128(sA(s, NPART - 1)), b_lo0 = umma_desc_k_sw128(sB(s, NPART - 1));
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t koff = static_cast<uint64_t>(kk * UMMA_K * 2) >> 4;
            umma_bf16(tmem_d, a_hi0 + koff, b_hi0 + koff, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            if (NPART == 2) {
              umma_bf16(tmem_d, a_lo0 + koff, b_hi0 + koff, idesc, 1u);
              umma_bf16(tmem_d, a_hi0 + koff, b_lo0 + koff, idesc, 1u);
            }
          }
          umma_commit(empty_bar(s));  // stage reusable once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(tmem_full_bar(acc));
        }
        __syncwarp();
      }
    }
  } else {  // ---------------- epilogue warps 2..9 ; TMEM lane quadrant = warp % 4, chunk parity = (warp - 2) / 4
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      // prefetch the first residual chunk while the MMAs of this tile are still running
      float4 rv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* rrow = (EPI == FACT_EPI_BIAS_RESID_F32 && row_ok) ? ep.resid + resid_row(row, ep) * ep.ldr
                                                                      : nullptr;
      if (EPI == FACT_EPI_BIAS_RESID_F32 && ep.vec_ok && row_ok && n0 + half * 32 + 32 <= N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = reinterpret_cast<const float4*>(rrow + n0 + half * 32)[i];
      }
      if (EPI == FACT_EPI_BIAS_RESID_F32) {
        // the residual tile of the NEXT tile is HBM-cold (written a layer ago): pull this thread's lines into L2 now,
        // a whole mainloop ahead, so the epilogue loads below only pay L2 latency
        const int nt = tile + static_cast<int>(gridDim.x);
        if (nt < num_tiles) {
          const int nrow = (nt / tiles_n) * BM + 0 + q * 32 + lane;
          if (nrow < M) {
            const float* nr = ep.resid + resid_row(nrow, ep) * ep.ldr + (nt % tiles_n) * BN;
#pragma unroll
            for (int c = half; c < BN / 32; c += 2) prefetch_l2(nr + c * 32);
          }
        }
      }
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        const bool fast = ep.vec_ok && col0 + 32 <= N;
        float4 rcur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rcur[i] = rv[i];
        // next chunk's residual goes in flight before this chunk's stores
        if (EPI == FACT_EPI_BIAS_RESID_F32 && c + 2 < BN / 32 && ep.vec_ok && row_ok && col0 + 64 + 32 <= N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rv[i] = reinterpret_cast<const float4*>(rrow + col0 + 64)[i];
        }
        epilogue_chunk<EPI>(v, rcur, row, col0, row_ok, fast, N, ep);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------- split-K (small M)
// Batch-1 decode has M = 360 tokens: 15-45 output tiles cannot fill 148 SMs, and every CTA would pull its whole A
// panel + W tile through one SM's TMA port.  Here the K blocks of every tile are dealt across `splits` CTAs
// (grid = tiles x splits); each writes its fp32 partial tile to its own slab part[z][M][N] with plain 16-byte stores,
// and gemm_finish_kernel sums the slabs and applies the fused epilogue -- no atomics, bit-reproducible.
template <int BN, int NPART, int STAGES>
__global__ void __launch_bounds__(192, 1) gemm_tc_splitk_kernel(
    const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, int M, int N, int K,
    int tiles_n, int kb_per, float* __restrict__ part) {
  using Cfg = GemmCfg<BN, NPART, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, z = blockIdx.y;
  const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
  const int num_kb = (K + BK - 1) / BK;
  const int kb0 = z * kb_per, nkb = min(kb_per, num_kb - kb0);
  constexpr int TCOLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // one accumulator
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TCOLS>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  auto sA = [&](int s, int part_) { return smem_base + s * Cfg::STAGE_BYTES + part_ * A_TILE_BYTES; };
  auto sB = [&](int s, int part_) {
    return smem_base + s * Cfg::STAGE_BYTES + NPART * A_TILE_BYTES + part_ * Cfg::B_TILE_BYTES;
  };
  if (warp == 0) {
    // The W operand is constant (packed weights): the first ring fill of W goes out BEFORE the dependency wait, so
    // under a programmatic dependent launch it streams from HBM while the previous kernel is still finishing; the
    // activations (A) follow once that kernel has completed.
    const int pre = nkb < STAGES ? nkb : STAGES;
    if (elect_one()) {
      for (int i = 0; i < pre; ++i) {
        const int kc = (kb0 + i) * BK;
        mbar_arrive_expect_tx(full_bar(i), Cfg::STAGE_BYTES);
        tma_load_2d(sB(i, 0), &tmB0, kc, n0, full_bar(i));
        if (NPART == 2) tma_load_2d(sB(i, 1), &tmB1, kc, n0, full_bar(i));
      }
    }
    __syncwarp();
    pdl_wait();
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      if (i >= pre) mbar_wait(empty_bar(s), ((i / STAGES) & 1) ^ 1);
      if (elect_one()) {
        const int kc = (kb0 + i) * BK;
        if (i >= pre) {
          mbar_arrive_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
          tma_load_2d(sB(s, 0), &tmB0, kc, n0, full_bar(s));
          if (NPART == 2) tma_load_2d(sB(s, 1), &tmB1, kc, n0, full_bar(s));
        }
        tma_load_2d(sA(s, 0), &tmA0, kc, m0, full_bar(s));
        if (NPART == 2) tma_load_2d(sA(s, 1), &tmA1, kc, m0, full_bar(s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      mbar_wait(full_bar(s), (i / STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_hi0 = umma_desc_k_sw128(sA(s, 0)), b_hi0 = umma_desc_k_sw128(sB(s, 0));
        const uint64_t a_lo0 = umma_desc_k_sw128(sA(s, NPART - 1)), b_lo0 = umma_desc_k_sw128(sB(s, NPART - 1));
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t koff = static_cast<uint64_t>(kk * UMMA_K * 2) >> 4;
          umma_bf16(tmem_base, a_hi0 + koff, b_hi0 + koff, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          if (NPART == 2) {
            umma_bf16(tmem_base, a_lo0 + koff, b_hi0 + koff, idesc, 1u);
            umma_bf16(tmem_base, a_hi0 + koff, b_lo0 + koff, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));
        if (i == nkb - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    pdl_wait();   // the slabs below may still be read by the finish kernel of the previous GEMM
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
    float* dst = part + (static_cast<size_t>(z) * M + row) * N + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (row < M) {
        if (n0 + c0 + 32 <= N && (N & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(dst + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (n0 + c0 + c < N) dst[c0 + c] = v[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}

// acc[row, col] = sum_z part[z][row][col], then the same fused epilogues as the in-kernel path (inference kinds)
__global__ void __launch_bounds__(256) gemm_finish_kernel(const float* __restrict__ part, int splits, int M, int N,
                                                          int kind, EpiArgs ep) {
  const long long total = static_cast<long long>(M) * N;
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int row = static_cast<int>(i / N), col = static_cast<int>(i % N);
    float x = 0.f;
    for (int zz = 0; zz < splits; ++zz) x += part[static_cast<size_t>(zz) * total + i];
    if (kind == FACT_EPI_SPLIT) {
      if (col < ep.scale_cols) x *= ep.scale;
    } else if (ep.bias) {
      x += ep.bias[col];
    }
    if (kind == FACT_EPI_BIAS_GELU_SPLIT) x = gelu_tanh(x);
    if (kind == FACT_EPI_BIAS_RESID_F32) x += ep.resid[resid_row(row, ep) * ep.ldr + col];
    if (kind == FACT_EPI_SPLIT || kind == FACT_EPI_BIAS_GELU_SPLIT) {
      bf16 h, l;
      split_bf16(x, h, l);
      const size_t o = static_cast<size_t>(row) * ep.ldo + col;
      ep.out_hi[o] = h;
      if (ep.out_lo) ep.out_lo[o] = l;
    } else {
      ep.out_f32[static_cast<size_t>(map_row(row, ep)) * ep.ldo + col] = x;
    }
  }
}

// the same, four columns per thread (n, pitches and pointers 16-byte friendly): the batch-1 decode runs 64 of these
// per frame, so the scalar version's per-element div / mod and 4-byte accesses showed up as 28 % of the frame
__global__ void __launch_bounds__(256) gemm_finish_vec_kernel(const float* __restrict__ part, int splits, int M, int N,
                                                              int kind, EpiArgs ep) {
  const int nv = N >> 2;
  const long long total4 = static_cast<long long>(M) * nv;
  const long long slab4 = total4;  // float4 elements per split slab
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total4; i += 256ll * gridDim.x) {
    const int row = static_cast<int>(i / nv), col = static_cast<int>(i % nv) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int zz = 0; zz < splits; ++zz) {
      const float4 v = reinterpret_cast<const float4*>(part)[zz * slab4 + i];
      x.x += v.x; x.y += v.y; x.z += v.z; x.w += v.w;
    }
    if (kind == FACT_EPI_SPLIT) {
      if (col + 0 < ep.scale_cols) x.x *= ep.scale;
      if (col + 1 < ep.scale_cols) x.y *= ep.scale;
      if (col + 2 < ep.scale_cols) x.z *= ep.scale;
      if (col + 3 < ep.scale_cols) x.w *= ep.scale;
    } else if (ep.bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
      x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
    }
    if (kind == FACT_EPI_BIAS_GELU_SPLIT) {
      x.x = gelu_tanh(x.x); x.y = gelu_tanh(x.y); x.z = gelu_tanh(x.z); x.w = gelu_tanh(x.w);
    }
    if (kind == FACT_EPI_BIAS_RESID_F32) {
      const float4 r = *reinterpret_cast<const float4*>(ep.resid + resid_row(row, ep) * ep.ldr + col);
      x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
    }
    if (kind == FACT_EPI_SPLIT || kind == FACT_EPI_BIAS_GELU_SPLIT) {
      const uint32_t h0 = cvt_bf16x2(x.x, x.y), h1 = cvt_bf16x2(x.z, x.w);
      const size_t o = static_cast<size_t>(row) * ep.ldo + col;
      *reinterpret_cast<uint2*>(ep.out_hi + o) = make_uint2(h0, h1);
      if (ep.out_lo) {
        const uint32_t l0 = cvt_bf16x2(x.x - __uint_as_float(h0 << 16), x.y - __uint_as_float(h0 & 0xffff0000u));
        const uint32_t l1 = cvt_bf16x2(x.z - __uint_as_float(h1 << 16), x.w - __uint_as_float(h1 & 0xffff0000u));
        *reinterpret_cast<uint2*>(ep.out_lo + o) = make_uint2(l0, l1);
      }
    } else {
      *reinterpret_cast<float4*>(ep.out_f32 + static_cast<size_t>(map_row(row, ep)) * ep.ldo + col) = x;
    }
  }
}

// BIAS_RESID_F32 finish fused with the LayerNorm that follows it in a pre-norm layer: one block per row (n <= 1024),
// thread = one float4 column group, the value stays in registers between the residual write and the normalisation.
// The reductions reproduce ln_split_kernel's order (per-lane partials over the eight 32-group strides, then the lane
// butterfly), so the result is bit-identical to running that kernel on the written row.
__global__ void __launch_bounds__(256) gemm_finish_ln_kernel(const float* __restrict__ part, int splits, int M, int N,
                                                             EpiArgs ep) {
  __shared__ float ps[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x, idx = threadIdx.x;
  const int nv = N >> 2;
  const bool on = idx < nv;
  const size_t slab4 = static_cast<size_t>(M) * nv;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  pdl_trigger();
  pdl_wait();
  if (on) {
    const float4* p4 = reinterpret_cast<const float4*>(part) + static_cast<size_t>(row) * nv + idx;
    for (int zz = 0; zz < splits; ++zz) {
      const float4 t = p4[zz * slab4];
      x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
    }
    if (ep.bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias) + idx);
      x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
    }
    const float4 r = reinterpret_cast<const float4*>(ep.resid + resid_row(row, ep) * ep.ldr)[idx];
    x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
    reinterpret_cast<float4*>(ep.out_f32 + static_cast<size_t>(row) * ep.ldo)[idx] = x;
  }
  ps[warp][lane] = (x.x + x.y) + (x.z + x.w);
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += ps[i][lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(N);
  const float a = x.x - mean, b = x.y - mean, c = x.z - mean, e = x.w - mean;
  __syncthreads();
  ps[warp][lane] = on ? (a * a + b * b) + (c * c + e * e) : 0.f;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) q += ps[i][lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(N) + 1e-5f);
  if (!on) return;
  const float4 gg = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma) + idx);
  const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.ln_beta) + idx);
  float4 y;
  y.x = (x.x - mean) * rstd * gg.x + bb.x;
  y.y = (x.y - mean) * rstd * gg.y + bb.y;
  y.z = (x.z - mean) * rstd * gg.z + bb.z;
  y.w = (x.w - mean) * rstd * gg.w + bb.w;
  bf16 h0, h1, h2, h3, l0, l1, l2, l3;
  split_bf16(y.x, h0, l0);
  split_bf16(y.y, h1, l1);
  split_bf16(y.z, h2, l2);
  split_bf16(y.w, h3, l3);
  reinterpret_cast<uint2*>(ep.ln_hi + static_cast<size_t>(row) * N)[idx] =
      make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
  if (ep.ln_lo)
    reinterpret_cast<uint2*>(ep.ln_lo + static_cast<size_t>(row) * N)[idx] =
        make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
}

// ------------------------------------------------------------------------------------------- CTA-pair variant
// cta_group::2: a cluster of two CTAs (one TPC) computes a 256 x BN tile.  CTA r stages A rows [128 r, 128 r + 128)
// and W rows [BN/2 r, BN/2 r + BN/2); the leader's single thread issues tcgen05.mma.cta_group::2 (UMMA 256 x BN x 16)
// which reads both CTAs' shared memory and writes each CTA's half of the accumulator into that CTA's TMEM.  Per SM and
// K block this stages A + W/2 instead of A + W: the measured TMA ingest limit (~64 B/cycle/SM) stops binding.
// Barriers: full[] and tmem_empty[] are the LEADER's (peer arrives remotely, peer TMA credits the leader's barrier);
// empty[] and tmem_full[] exist in both CTAs and are signalled by multicast tcgen05.commit.
constexpr int TS_WARP_BYTES = 4096;  // per epilogue warp: 32 rows x 128 B (one fp32 box, or two bf16 boxes)
template <int BN, int NPART, int STAGES, bool TS = false>
struct Gemm2Cfg {
  static constexpr int B_TILE_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = NPART * (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int STAGING_BYTES = TS ? EPI_WARPS * TS_WARP_BYTES : 0;
  static constexpr int ACC_COLS = 2 * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128
                                   : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;
  static constexpr int LN_GB_OFF = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;  // LNW: gamma | beta, fp32 [1024] each
  static constexpr int SMEM_BYTES_LN = SMEM_BYTES + 2 * 1024 * 4;
  static_assert(STAGE_BYTES % 1024 == 0, "staging boxes need the 1024-byte alignment the stages keep");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "UMMA N constraint for M=256 / 32-column epilogue chunks");
  static_assert((BN / 2) % 8 == 0, "W half must be whole 8-row swizzle atoms");
  static_assert(ACC_COLS <= 512, "TMEM has 512 columns");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
  static_assert((2 * STAGES + 5) * 8 <= BAR_BYTES, "barrier block too small");
};


// ---- TMA-store epilogue (pair kernel): the same arithmetic as epilogue_chunk, but the results stay in registers
// (packed bf16 pairs o0/o1, or fp32 of) and the caller stages them in shared memory for one bulk tensor store per
// 32 x 32 box.  Rows / columns outside [M, N] are computed on zeros and clipped by the store.
__device__ __forceinline__ void load_bias32(float* b, const EpiArgs& ep, int col0, int N) {
  if (ep.bias == nullptr) {
#pragma unroll
    for (int c = 0; c < 32; ++c) b[c] = 0.f;
  } else if (ep.vec_ok && col0 + 32 <= N) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(b4 + i);
      b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) b[c] = (col0 + c < N) ? __ldg(ep.bias + col0 + c) : 0.f;
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_compute(const float* v, int row, int col0, bool row_ok, int N,
                                                 const EpiArgs& ep, uint32_t* o0, uint32_t* o1, float* of) {
  if (EPI == FACT_EPI_BIAS_RESID_F32 || EPI == FACT_EPI_BIAS_F32) {
    float b[32];
    load_bias32(b, ep, col0, N);
#pragma unroll
    for (int c = 0; c < 32; ++c) of[c] = v[c] + b[c];
    if (EPI == FACT_EPI_BIAS_RESID_F32 && !ep.reduce_add && row_ok) {
      const float* rrow = ep.resid + resid_row(row, ep) * ep.ldr + col0;
      if (ep.vec_ok && col0 + 32 <= N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 r = reinterpret_cast<const float4*>(rrow)[i];
          of[4 * i] += r.x; of[4 * i + 1] += r.y; of[4 * i + 2] += r.z; of[4 * i + 3] += r.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (col0 + c < N) of[c] += rrow[c];
      }
    }
    return;
  }
  float x[32];
  if (EPI == FACT_EPI_BIAS_GELU_SAVE) {
    float b[32];
    load_bias32(b, ep, col0, N);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float z0 = v[2 * c] + b[2 * c], z1 = v[2 * c + 1] + b[2 * c + 1];
      o1[c] = cvt_bf16x2(z0, z1);
      o0[c] = cvt_bf16x2(gelu_tanh(z0), gelu_tanh(z1));
    }
    return;
  }
  if (EPI == FACT_EPI_GELU_GRAD) {
    uint32_t zraw[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) zraw[c] = 0u;
    if (row_ok) {
      const bf16* zrow = ep.aux + static_cast<size_t>(row) * ep.ldaux + col0;
      if (ep.vec_ok && col0 + 32 <= N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 q4 = reinterpret_cast<const uint4*>(zrow)[i];
          zraw[4 * i] = q4.x; zraw[4 * i + 1] = q4.y; zraw[4 * i + 2] = q4.z; zraw[4 * i + 3] = q4.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint32_t a = (col0 + 2 * c < N) ? __bfloat16_as_ushort(zrow[2 * c]) : 0u;
          const uint32_t b = (col0 + 2 * c + 1 < N) ? __bfloat16_as_ushort(zrow[2 * c + 1]) : 0u;
          zraw[c] = a | (b << 16);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const uint32_t raw = zraw[c >> 1];
      const float z = __uint_as_float((c & 1) ? (raw & 0xffff0000u) : (raw << 16));
      x[c] = v[c] * gelu_tanh_grad(z);
    }
  } else if (EPI == FACT_EPI_BIAS_GELU_SPLIT) {
    float b[32];
    load_bias32(b, ep, col0, N);
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = gelu_tanh(v[c] + b[c]);
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = (col0 + c < ep.scale_cols) ? v[c] * ep.scale : v[c];
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const uint32_t hp = cvt_bf16x2(x[2 * c], x[2 * c + 1]);
    o0[c] = hp;
    o1[c] = cvt_bf16x2(x[2 * c] - __uint_as_float(hp << 16), x[2 * c + 1] - __uint_as_float(hp & 0xffff0000u));
  }
}

// LNW (BIAS_RESID_F32 through the bulk-store epilogue only): eight more warps per CTA normalise the finished rows.
// A residual-stream row is complete when the tiles_n column tiles of its 256-row block have all landed, and those are
// handled by tiles_n different clusters: every epilogue warp counts its 32-row group in ep.ln_sync once its bulk stores
// / reductions of a tile have completed, and the LayerNorm warps that own rows of the group wait for the count
// (2 warps x tiles_n), then read the rows back from the L2 (one row in flight behind the one being normalised; gamma
// and beta sit in shared memory) and write LayerNorm(row) split into bf16 hi / lo -- the A operand of the next GEMM --
// while the MMA and epilogue warps are already on their next tiles.  Same arithmetic as
// ln_split_kernel (ln_row_finish), which this replaces: one launch and one HBM read of the residual stream fewer.
constexpr int LN_WARPS = 8;
template <int BN, int NPART, int STAGES, int EPI, bool TS, bool LNW = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS + (LNW ? LN_WARPS * 32 : 0), 1) gemm_tc2_kernel(
    const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
    const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1, int M, int N, int K,
    int tiles_n, int num_tiles, EpiArgs ep) {
  using Cfg = Gemm2Cfg<BN, NPART, STAGES, TS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t staging_base = smem_base + STAGES * Cfg::STAGE_BYTES;  // TS: 8 x 4 KB, one slot per epilogue warp
  const uint32_t bar_base = staging_base + Cfg::STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (NPART == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    if (TS) {
      tma_prefetch_desc(&tmO0);
      tma_prefetch_desc(&tmO1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);   // leader's arrive.expect_tx + peer's remote arrive
      mbar_init(empty_bar(s), 1);  // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 2 * EPI_WARPS);  // both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  pdl_trigger();
  if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail

  auto sA = [&](int s, int part) { return smem_base + s * Cfg::STAGE_BYTES + part * A_TILE_BYTES; };
  auto sB = [&](int s, int part) {
    return smem_base + s * Cfg::STAGE_BYTES + NPART * A_TILE_BYTES + part * Cfg::B_TILE_BYTES;
  };

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs): whole warp walks the ring, one elected lane issues
    uint32_t it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / tiles_n) * (2 * BM) + rank * BM;
      const int n0 = (tile % tiles_n) * BN + rank * (BN / 2);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
          else mbar_arrive_leader(full_bar(s));
          tma_load_2d_2sm(sA(s, 0), &tmA0, kb * BK, m0, full_bar(s));
          tma_load_2d_2sm(sB(s, 0), &tmB0, kb * BK, n0, full_bar(s));
          if (NPART == 2) {
            tma_load_2d_2sm(sA(s, 1), &tmA1, kb * BK, m0, full_bar(s));
            tma_load_2d_2sm(sB(s, 1), &tmB1, kb * BK, n0, full_bar(s));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {  // ---------------- MMA issuer (leader CTA only): whole warp walks, one elected lane issues
      constexpr uint32_t idesc = umma_idesc_bf16_f32(2 * BM, BN);
      uint32_t it = 0, t = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++t) {
        const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
        mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_hi0 = umma_desc_k_sw128(sA(s, 0)), b_hi0 = umma_desc_k_sw128(sB(s, 0));
            const uint64_t a_lo0 = umma_desc_k_sw128(sA(s, NPART - 1)), b_lo0 = umma_desc_k_sw128(sB(s, NPART - 1));
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              const uint64_t koff = static_cast<uint64_t>(kk * UMMA_K * 2) >> 4;
              umma_bf16_2sm(tmem_d, a_hi0 + koff, b_hi0 + koff, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              if (NPART == 2) {
                umma_bf16_2sm(tmem_d, a_lo0 + koff, b_hi0 + koff, idesc, 1u);
                umma_bf16_2sm(tmem_d, a_hi0 + koff, b_lo0 + koff, idesc, 1u);
              }
            }
            umma_commit_2sm(empty_bar(s), 0x3);  // both CTAs' stage s is free
            if (kb == num_kb - 1) umma_commit_2sm(tmem_full_bar(acc), 0x3);
          }
          __syncwarp();
        }
      }
    }
  } else if (LNW && warp >= 2 + EPI_WARPS) {  // ---------------- LayerNorm warps (both CTAs)
    // The 256 rows of a row block are dealt over the LayerNorm warps of ALL the clusters that work on its column
    // tiles (tiles_n clusters x 2 CTAs x LN_WARPS warps = 80 units of 3 or 4 rows at N = 800): every warp has a small
    // job per tile instead of a long one every tiles_n-th tile, so the rows of the last tiles are normalised a few
    // microseconds after they land (the kernel's tail) instead of a whole 128-row job later.
    const int j = warp - (2 + EPI_WARPS);
    const int target = 2 * tiles_n;  // two epilogue warps (column halves) per column tile
    const int nv = N >> 2;
    const int units = tiles_n * 2 * LN_WARPS;
    int* done = ep.ln_sync + ((M + 31) >> 5) + 1;  // second half of ln_sync: units that are past their wait
    auto unit_of_row = [&](int r) { return ((r + 1) * units + 2 * BM - 1) / (2 * BM) - 1; };
    float* gb = reinterpret_cast<float*>(smem_gen + Cfg::LN_GB_OFF);
    for (int i = j * 32 + lane; i < N; i += LN_WARPS * 32) {
      gb[i] = __ldg(ep.ln_gamma + i);
      gb[1024 + i] = __ldg(ep.ln_beta + i);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(LN_WARPS * 32) : "memory");
    const float* sg = gb;
    const float* sb = gb + 1024;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int base = (tile / tiles_n) * (2 * BM);
      const int u = (tile % tiles_n) * (2 * LN_WARPS) + static_cast<int>(rank) * LN_WARPS + j;
      const int r_lo = (u * 2 * BM) / units, r_hi = ((u + 1) * 2 * BM) / units;  // rows of the block owned by unit u
      if (r_hi == r_lo || base + (r_lo & ~31) >= M) continue;
      const int g_lo = r_lo >> 5, g_hi = (r_hi - 1) >> 5;
      if (lane == 0) {
        for (int g = g_lo; g <= g_hi && base + g * 32 < M; ++g) {
          const int* ctr = ep.ln_sync + ((base + g * 32) >> 5);
          // bounded: a GEMM launch lasts about a millisecond; a count that has not arrived after ~a second never will
          // (dirty counters, or two streams sharing them) -> fail the launch instead of hanging
          unsigned spins = 0;
          while (ld_relaxed_gpu(ctr) < target) {
            __nanosleep(100);
            if (++spins > (1u << 23)) __trap();
          }
        }
        __threadfence();  // acquire: the rows counted above are visible to the loads below
        // the last unit past its wait on a group clears the group's counters for the next launch
        for (int g = g_lo; g <= g_hi && base + g * 32 < M; ++g) {
          const int gi = (base + g * 32) >> 5;
          const int waiters = unit_of_row(g * 32 + 31) - unit_of_row(g * 32) + 1;
          if (atomicAdd(done + gi, 1) == waiters - 1) {
            ep.ln_sync[gi] = 0;
            done[gi] = 0;
          }
        }
      }
      __syncwarp();
      const int row0 = base + r_lo;
      const int nrows = min(r_hi - r_lo, M - row0);
      if (nrows <= 0) continue;
      auto xrow = [&](int r) { return ep.out_f32 + static_cast<size_t>(row0 + r) * ep.ldo; };
      auto finish = [&](const float4* v, int r) {
        ln_row_finish<true, false>(v, sg, sb, ep.ln_hi + static_cast<size_t>(row0 + r) * N,
                                   ep.ln_lo ? ep.ln_lo + static_cast<size_t>(row0 + r) * N : nullptr, N, lane);
      };
      float4 va[8], vb[8];
      ln_row_load<true>(xrow(0), nv, lane, va);
      for (int r = 0; r < nrows; r += 2) {  // the next row's loads (from the L2) are in flight while this one is normalised
        if (r + 1 < nrows) ln_row_load<true>(xrow(r + 1), nv, lane, vb);
        finish(va, r);
        if (r + 1 < nrows) {
          if (r + 2 < nrows) ln_row_load<true>(xrow(r + 2), nv, lane, va);
          finish(vb, r + 1);
        }
      }
    }
  } else if (TS) {  // ---------------- epilogue warps, bulk-store flavour: stage a 32 x 32 box, one TMA store per box
    const int q = warp & 3, half = (warp - 2) >> 2;
    int ln_pending = -1;  // LNW: 32-row group whose stores of the previous tile are not yet counted in ep.ln_sync
    const uint32_t stg = staging_base + static_cast<uint32_t>(warp - 2) * TS_WARP_BYTES;
    constexpr bool F32_OUT = EPI == FACT_EPI_BIAS_RESID_F32 || EPI == FACT_EPI_BIAS_F32;
    const bool two = !F32_OUT && ep.out_lo != nullptr;
    uint32_t t = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++t) {
      const int m0 = (tile / tiles_n) * (2 * BM) + rank * BM, n0 = (tile % tiles_n) * BN;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      const int row0 = m0 + q * 32, row = row0 + lane;
      const bool row_ok = row < M;
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
      if (LNW && ln_pending >= 0) {
        // the previous tile's boxes were issued a whole main loop ago: by now this wait is free.  Completed bulk
        // stores / reductions + gpu-scope fence + atomic = the rows are visible to whoever reads the full count.
        if (lane == 0) {
          bulk_wait_group0();
          __threadfence();
          atomicAdd(ep.ln_sync + (ln_pending >> 5), 1);
        }
        ln_pending = -1;
      }
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int col0 = n0 + c * 32;
        if (col0 >= N || row0 >= M) continue;  // warp-uniform: nothing of this box is inside the matrix
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        uint32_t o0[16], o1[16];
        float of[32];
        epilogue_compute<EPI>(v, row, col0, row_ok, N, ep, o0, o1, of);
        if (lane == 0) bulk_wait_group_read0();  // the previous box of this warp has left the staging slot
        __syncwarp();
        if (F32_OUT) {  // 128-byte rows, SWIZZLE_128B: 16-byte chunk i of row r sits at chunk i ^ (r & 7)
          const uint32_t rbase = stg + lane * 128, sw = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            st_shared_v4(rbase + ((i ^ sw) << 4), __float_as_uint(of[4 * i]), __float_as_uint(of[4 * i + 1]),
                         __float_as_uint(of[4 * i + 2]), __float_as_uint(of[4 * i + 3]));
        } else {  // 64-byte rows, SWIZZLE_64B: chunk i of row r sits at chunk i ^ ((r >> 1) & 3)
          const uint32_t rbase = stg + lane * 64, sw = (lane >> 1) & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            st_shared_v4(rbase + ((i ^ sw) << 4), o0[4 * i], o0[4 * i + 1], o0[4 * i + 2], o0[4 * i + 3]);
            if (two)
              st_shared_v4(rbase + 2048 + ((i ^ sw) << 4), o1[4 * i], o1[4 * i + 1], o1[4 * i + 2], o1[4 * i + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (EPI == FACT_EPI_BIAS_RESID_F32 && ep.reduce_add) tma_reduce_add_2d(&tmO0, stg, col0, row0);
          else tma_store_2d(&tmO0, stg, col0, row0);
          if (two) tma_store_2d(&tmO1, stg + 2048, col0, row0);
          bulk_commit_group();
        }
        if (EPI == FACT_EPI_GELU_GRAD && ep.colsum != nullptr) {
          // bias gradient: lane c sums column c of the staged box (rows past M hold zeros), one atomic per column
          float s = 0.f;
          const uint32_t cb = stg + (lane & 7) * 2;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            uint16_t hv;
            asm volatile("ld.shared.u16 %0, [%1];"
                         : "=h"(hv)
                         : "r"(cb + r * 64 + (((lane >> 3) ^ ((r >> 1) & 3)) << 4)));
            s += __uint_as_float(static_cast<uint32_t>(hv) << 16);
          }
          if (col0 + lane < N) atomicAdd(ep.colsum + col0 + lane, s);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tmem_empty_bar(acc));
        else mbar_arrive_leader(tmem_empty_bar(acc));
      }
      if (LNW && row0 < M) ln_pending = row0;
    }
    if (lane == 0) bulk_wait_group0();  // all boxes written before the CTA (and its shared memory) goes away
    if (LNW && ln_pending >= 0 && lane == 0) {
      __threadfence();
      atomicAdd(ep.ln_sync + (ln_pending >> 5), 1);
    }
  } else {  // ---------------- epilogue warps (both CTAs): this CTA's 128 rows of the pair tile
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint32_t t = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++t) {
      const int m0 = (tile / tiles_n) * (2 * BM) + rank * BM, n0 = (tile % tiles_n) * BN;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      float4 rv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* rrow = (EPI == FACT_EPI_BIAS_RESID_F32 && row_ok) ? ep.resid + resid_row(row, ep) * ep.ldr
                                                                      : nullptr;
      if (EPI == FACT_EPI_BIAS_RESID_F32 && ep.vec_ok && row_ok && n0 + half * 32 + 32 <= N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = reinterpret_cast<const float4*>(rrow + n0 + half * 32)[i];
      }
      if (EPI == FACT_EPI_BIAS_RESID_F32) {
        // the residual tile of the NEXT tile is HBM-cold (written a layer ago): pull this thread's lines into L2 now,
        // a whole mainloop ahead, so the epilogue loads below only pay L2 latency
        const int nt = tile + num_clusters;
        if (nt < num_tiles) {
          const int nrow = (nt / tiles_n) * (2 * BM) + static_cast<int>(rank) * BM + q * 32 + lane;
          if (nrow < M) {
            const float* nr = ep.resid + resid_row(nrow, ep) * ep.ldr + (nt % tiles_n) * BN;
#pragma unroll
            for (int c = half; c < BN / 32; c += 2) prefetch_l2(nr + c * 32);
          }
        }
      }
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        const bool fast = ep.vec_ok && col0 + 32 <= N;
        float4 rcur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rcur[i] = rv[i];
        if (EPI == FACT_EPI_BIAS_RESID_F32 && c + 2 < BN / 32 && ep.vec_ok && row_ok && col0 + 64 + 32 <= N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rv[i] = reinterpret_cast<const float4*>(rrow + col0 + 64)[i];
        }
        epilogue_chunk<EPI>(v, rcur, row, col0, row_ok, fast, N, ep);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tmem_empty_bar(acc));
        else mbar_arrive_leader(tmem_empty_bar(acc));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody frees TMEM / exits while the peer may still touch this CTA's barriers or TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int rows, cols, ld, box_rows, box_cols;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           box_cols == o.box_cols;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) ^ (static_cast<size_t>(k.cols) << 20) ^
         (static_cast<size_t>(k.ld) << 40) ^ (static_cast<size_t>(k.box_rows) << 52) ^
         (static_cast<size_t>(k.box_cols) << 8);
    return h;
  }
};

// bf16 row-major [rows, cols] with row pitch ld elements -> 2-D tiled map, box {box_cols, box_rows}; the swizzle
// span equals the box width (64 cols -> SWIZZLE_128B, 16 cols -> SWIZZLE_32B).
int make_tmap_bf16(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int box_rows, int box_cols) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, box_rows, box_cols};
  FACT_REQUIRE(box_cols == 64 || box_cols == 16, FACT_ERR_UNSUPPORTED, "tensor-map box width %d", box_cols);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return FACT_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  FACT_REQUIRE(enc != nullptr, FACT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  FACT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, FACT_ERR_BAD_ALIGN, "TMA base must be 16-B aligned");
  FACT_REQUIRE((static_cast<long long>(ld) * 2) % 16 == 0, FACT_ERR_BAD_ALIGN,
               "TMA row pitch (%d bf16) must be a multiple of 16 bytes", ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FACT_REQUIRE(r == CUDA_SUCCESS, FACT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%d",
               static_cast<int>(r), rows, cols, ld, box_rows);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return FACT_OK;
}

// Row-major [rows, cols] matrix of bf16 (elem_bytes 2) or fp32 (4) -> map with a {box_cols, box_rows} box whose swizzle
// span equals the box's row bytes (32 / 64 / 128), the layout the staging writes of the bulk-store epilogues assume.
int make_tmap_box(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int elem_bytes, int box_rows,
                  int box_cols) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, box_rows, box_cols * 8 + elem_bytes};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return FACT_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  FACT_REQUIRE(enc != nullptr, FACT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const int span = box_cols * elem_bytes;
  FACT_REQUIRE(span == 32 || span == 64 || span == 128, FACT_ERR_UNSUPPORTED, "tensor-map box row of %d bytes", span);
  FACT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (static_cast<long long>(ld) * elem_bytes) % 16 == 0,
               FACT_ERR_BAD_ALIGN, "bulk tensor store needs a 16-byte aligned base and row pitch");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = span == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FACT_REQUIRE(r == CUDA_SUCCESS, FACT_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (%d) rows=%d cols=%d ld=%d",
               static_cast<int>(r), rows, cols, ld);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return FACT_OK;
}
int make_tmap_rows3d(CUtensorMap* out, const void* ptr, int batches, int rows, int cols, int ld, int box_rows,
                     int box_cols) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, box_rows + (batches << 8), box_cols};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return FACT_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  FACT_REQUIRE(enc != nullptr, FACT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  FACT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (static_cast<long long>(ld) * 2) % 16 == 0 &&
                   (box_cols * 2) % 16 == 0,
               FACT_ERR_BAD_ALIGN, "bulk tensor store needs 16-byte aligned base, row pitch and box rows");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batches)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * rows};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FACT_REQUIRE(r == CUDA_SUCCESS, FACT_ERR_CUDA, "cuTensorMapEncodeTiled (3-D rows) failed (%d) batches=%d rows=%d cols=%d",
               static_cast<int>(r), batches, rows, cols);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return FACT_OK;
}
// the 32 x 32 box of the GEMM epilogues
int make_tmap_out(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int elem_bytes) {
  return make_tmap_box(out, ptr, rows, cols, ld, elem_bytes, 32, 32);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

template <int BN, int NPART, int STAGES, int EPI>
static int launch_cfg(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                      int m, int n, int k, const EpiArgs& ep, cudaStream_t st) {
  using Cfg = GemmCfg<BN, NPART, STAGES>;
  auto kern = gemm_tc_kernel<BN, NPART, STAGES, EPI>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int tiles_n = (n + BN - 1) / BN, tiles_m = (m + BM - 1) / BM;
  const int num_tiles = tiles_n * tiles_m;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  FACT_CUDA_CHECK(launch_k(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, true, a0, a1, b0, b1, m, n, k,
                           tiles_n, num_tiles, ep));
  FACT_LAUNCH_CHECK("gemm_tc_kernel launch");
  return FACT_OK;
}

template <int BN, int NPART, int STAGES>
static int launch_splitk(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                         int m, int n, int k, int splits, int kb_per, int kind, const EpiArgs& ep, float* part,
                         cudaStream_t st) {
  using Cfg = GemmCfg<BN, NPART, STAGES>;
  auto kern = gemm_tc_splitk_kernel<BN, NPART, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int tiles_n = (n + BN - 1) / BN, tiles_m = (m + BM - 1) / BM;
  dim3 grid(tiles_n * tiles_m, splits);
  FACT_CUDA_CHECK(launch_k(kern, grid, dim3(192), Cfg::SMEM_BYTES, st, true, a0, a1, b0, b1, m, n, k, tiles_n, kb_per,
                           part));
  FACT_LAUNCH_CHECK("gemm_tc_splitk_kernel launch");
  const long long total = static_cast<long long>(m) * n;
  if (ep.ln_hi != nullptr) {  // eligibility was checked by the caller (fact_gemm)
    FACT_CUDA_CHECK(launch_k(gemm_finish_ln_kernel, dim3(m), dim3(256), 0, st, true, static_cast<const float*>(part),
                             splits, m, n, ep));
    FACT_LAUNCH_CHECK("gemm_finish_ln_kernel launch");
    return FACT_OK;
  }
  if (ep.vec_ok && n % 4 == 0 && ep.ldo % 4 == 0) {  // vec_ok: pitches / pointers allow 16-byte accesses
    const long long total4 = total / 4;
    int fgrid = static_cast<int>((total4 + 255) / 256 > 2048 ? 2048 : (total4 + 255) / 256);
    FACT_CUDA_CHECK(launch_k(gemm_finish_vec_kernel, dim3(fgrid), dim3(256), 0, st, true,
                             static_cast<const float*>(part), splits, m, n, kind, ep));
  } else {
    int fgrid = static_cast<int>((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    FACT_CUDA_CHECK(launch_k(gemm_finish_kernel, dim3(fgrid), dim3(256), 0, st, true, static_cast<const float*>(part),
                             splits, m, n, kind, ep));
  }
  FACT_LAUNCH_CHECK("gemm_finish_kernel launch");
  return FACT_OK;
}

struct OutMaps {
  CUtensorMap o0, o1;
};

template <int BN, int NPART, int STAGES, int EPI, bool TS, bool LNW = false>
static int launch_cfg2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                       const OutMaps& om, int m, int n, int k, const EpiArgs& ep, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN, NPART, STAGES, TS>;
  auto kern = gemm_tc2_kernel<BN, NPART, STAGES, EPI, TS, LNW>;
  constexpr int smem_bytes = LNW ? Cfg::SMEM_BYTES_LN : Cfg::SMEM_BYTES;
  static_assert(smem_bytes <= 232448, "exceeds 227 KB");
  static bool attr_done = false;
  if (!attr_done) {
    FACT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_done = true;
  }
  const int tiles_n = (n + BN - 1) / BN, tiles_m = (m + 2 * BM - 1) / (2 * BM);
  const int num_tiles = tiles_n * tiles_m;
  int clusters = num_sms() / 2;
  if (num_tiles < clusters) clusters = num_tiles;
  FACT_CUDA_CHECK(launch_k(kern, dim3(2 * clusters), dim3(GEMM_THREADS + (LNW ? LN_WARPS * 32 : 0)), smem_bytes, st, true,
                           a0, a1, b0, b1, om.o0, om.o1, m, n, k, tiles_n, num_tiles, ep));
  FACT_LAUNCH_CHECK("gemm_tc2_kernel launch");
  return FACT_OK;
}

template <int BN, int NPART, int STAGES, bool TS>
static int launch_epi2(int kind, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                       const CUtensorMap& b1, const OutMaps& om, int m, int n, int k, const EpiArgs& ep,
                       cudaStream_t st) {
  switch (kind) {
    case FACT_EPI_SPLIT:
      return launch_cfg2<BN, NPART, STAGES, FACT_EPI_SPLIT, TS>(a0, a1, b0, b1, om, m, n, k, ep, st);
    case FACT_EPI_BIAS_GELU_SPLIT:
      return launch_cfg2<BN, NPART, STAGES, FACT_EPI_BIAS_GELU_SPLIT, TS>(a0, a1, b0, b1, om, m, n, k, ep, st);
    case FACT_EPI_BIAS_RESID_F32:
      return launch_cfg2<BN, NPART, STAGES, FACT_EPI_BIAS_RESID_F32, TS>(a0, a1, b0, b1, om, m, n, k, ep, st);
    case FACT_EPI_BIAS_F32:
      return launch_cfg2<BN, NPART, STAGES, FACT_EPI_BIAS_F32, TS>(a0, a1, b0, b1, om, m, n, k, ep, st);
    case FACT_EPI_BIAS_GELU_SAVE:
      if (NPART == 1)
        return launch_cfg2<BN, 1, STAGES, FACT_EPI_BIAS_GELU_SAVE, TS>(a0, a1, b0, b1, om, m, n, k, ep, st);
      break;
    case FACT_EPI_GELU_GRAD:
      if (NPART == 1) return launch_cfg2<BN, 1, STAGES, FACT_EPI_GELU_GRAD, TS>(a0, a1, b0, b1, om, m, n, k, ep, st);
      break;
  }
  set_error("epilogue kind %d unknown, or a training epilogue (4, 5) used with split (precise) operands", kind);
  return FACT_ERR_UNSUPPORTED;
}

template <int BN, int NPART, int STAGES>
static int launch_epi(int kind, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                      const CUtensorMap& b1, int m, int n, int k, const EpiArgs& ep, cudaStream_t st) {
  switch (kind) {
    case FACT_EPI_SPLIT:
      return launch_cfg<BN, NPART, STAGES, FACT_EPI_SPLIT>(a0, a1, b0, b1, m, n, k, ep, st);
    case FACT_EPI_BIAS_GELU_SPLIT:
      return launch_cfg<BN, NPART, STAGES, FACT_EPI_BIAS_GELU_SPLIT>(a0, a1, b0, b1, m, n, k, ep, st);
    case FACT_EPI_BIAS_RESID_F32:
      return launch_cfg<BN, NPART, STAGES, FACT_EPI_BIAS_RESID_F32>(a0, a1, b0, b1, m, n, k, ep, st);
    case FACT_EPI_BIAS_F32:
      return launch_cfg<BN, NPART, STAGES, FACT_EPI_BIAS_F32>(a0, a1, b0, b1, m, n, k, ep, st);
    case FACT_EPI_BIAS_GELU_SAVE:
      if (NPART == 1) return launch_cfg<BN, 1, STAGES, FACT_EPI_BIAS_GELU_SAVE>(a0, a1, b0, b1, m, n, k, ep, st);
      break;
    case FACT_EPI_GELU_GRAD:
      if (NPART == 1) return launch_cfg<BN, 1, STAGES, FACT_EPI_GELU_GRAD>(a0, a1, b0, b1, m, n, k, ep, st);
      break;
  }
  set_error("epilogue kind %d unknown, or a training epilogue (4, 5) used with split (precise) operands", kind);
  return FACT_ERR_UNSUPPORTED;
}

int g_gemm_bn = 0;  // fact_set_flag("gemm_bn", 128 | 160 | 256): force the N tile (experiments); 0 = automatic
int gemm_tile_n(int n) {
  if (g_gemm_bn == 128 || g_gemm_bn == 160 || g_gemm_bn == 256) return g_gemm_bn;
  if (n % 160 == 0) return 160;  // 2400 = 15 x 160, 800 = 5 x 160
  if (n % 256 == 0) return 256;  // 3072 = 12 x 256
  return 128;
}

int g_gemm_tma_store = 1;  // fact_set_flag("gemm_tma_store", 0 | 1 | 2): pair-kernel epilogue through bulk tensor stores
                           // (0 = direct row-per-lane stores, 2 = bulk stores but no in-place bulk reduction)
int g_gemm_fuse_ln = 1;    // fact_set_flag("gemm_fuse_ln", 0): large batches keep the LayerNorm as its own launch
int g_gemm_finish_ln = 1;  // fact_set_flag("gemm_finish_ln", 0): keep the LayerNorm out of the split-K finish kernel
                           // (batch 1: 722 frames/s fused, block per row, vs 656 with a separate LayerNorm launch)
int g_pdl = 1;          // fact_set_flag("pdl", 0): plain serialized launches for the small-batch decode chain
int g_gemm_pair = 1;    // fact_set_flag("gemm_pair", 0) forces the 1-SM kernel
int g_gemm_splitk = 1;  // fact_set_flag("gemm_splitk", 0) disables the small-M split-K path

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace fact

using namespace fact;

namespace fact {
int colsum_bf16(const void* x, int ldx, float* colsum, int rows, int n, cudaStream_t st);  // backward.cu
}

static int gemm_dispatch(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                         int m, int n, int k, const fact_gemm_epilogue* epi, void* stream, bool* colsum_fused);

extern "C" int fact_gemm(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                         int m, int n, int k, const fact_gemm_epilogue* epi, void* stream) {
  FACT_REQUIRE(a_hi && w_hi && epi, FACT_ERR_BAD_SHAPE, "fact_gemm: null operand");
  bool colsum_fused = false;  // also reports a fused LayerNorm (the two options belong to different epilogue kinds)
  if (epi->ln_hi) {
    FACT_REQUIRE(epi->kind == FACT_EPI_BIAS_RESID_F32 && epi->seq_in == 0 && epi->ln_gamma && epi->ln_beta &&
                     epi->ldo == n && n % 4 == 0 && n <= 1024,
                 FACT_ERR_BAD_SHAPE, "fact_gemm: the LayerNorm option needs BIAS_RESID_F32, no row remap, ldo == n <= 1024");
  }
  int rc = gemm_dispatch(a_hi, a_lo, lda, w_hi, w_lo, ldw, m, n, k, epi, stream, &colsum_fused);
  if (rc == FACT_OK && epi->kind == FACT_EPI_GELU_GRAD && epi->colsum && !colsum_fused)
    rc = colsum_bf16(epi->out_hi, epi->ldo, epi->colsum, m, n, as_stream(stream));  // kernels without the fused sums
  if (rc == FACT_OK && epi->ln_hi && !colsum_fused)
    rc = fact_layernorm_split(epi->out_f32, epi->ln_gamma, epi->ln_beta, epi->ln_hi, epi->ln_lo, m, n, stream);
  return rc;
}

static int gemm_dispatch(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                         int m, int n, int k, const fact_gemm_epilogue* epi, void* stream, bool* colsum_fused) {
  FACT_REQUIRE(m > 0 && n > 0 && k > 0, FACT_ERR_BAD_SHAPE, "fact_gemm: bad shape m=%d n=%d k=%d", m, n, k);
  FACT_REQUIRE((a_lo == nullptr) == (w_lo == nullptr), FACT_ERR_BAD_SHAPE,
               "fact_gemm: a_lo and w_lo must both be given (precise) or both be NULL (bf16)");
  const bool split_out = epi->kind == FACT_EPI_SPLIT || epi->kind == FACT_EPI_BIAS_GELU_SPLIT ||
                         epi->kind == FACT_EPI_BIAS_GELU_SAVE || epi->kind == FACT_EPI_GELU_GRAD;
  if (split_out) {
    FACT_REQUIRE(epi->out_hi != nullptr && (epi->ldo % 2) == 0, FACT_ERR_BAD_SHAPE,
                 "fact_gemm: split epilogue needs out_hi and an even ldo");
    FACT_REQUIRE((epi->kind != FACT_EPI_BIAS_GELU_SPLIT && epi->kind != FACT_EPI_BIAS_GELU_SAVE) || epi->bias,
                 FACT_ERR_BAD_SHAPE, "gelu epilogue needs bias");
    FACT_REQUIRE(epi->kind != FACT_EPI_BIAS_GELU_SAVE || epi->out_lo, FACT_ERR_BAD_SHAPE,
                 "GELU_SAVE writes the pre-activation through out_lo");
    FACT_REQUIRE(epi->kind != FACT_EPI_GELU_GRAD || epi->aux, FACT_ERR_BAD_SHAPE, "GELU_GRAD needs aux");
  } else {
    FACT_REQUIRE(epi->out_f32 != nullptr, FACT_ERR_BAD_SHAPE, "fact_gemm: f32 epilogue needs out_f32");
    FACT_REQUIRE(epi->kind != FACT_EPI_BIAS_RESID_F32 || epi->resid, FACT_ERR_BAD_SHAPE, "resid epilogue needs resid");
  }
  EpiArgs ep;
  ep.out_f32 = epi->out_f32;
  ep.out_hi = static_cast<bf16*>(epi->out_hi);
  ep.out_lo = static_cast<bf16*>(epi->out_lo);
  ep.ldo = epi->ldo;
  ep.bias = epi->bias;
  ep.resid = epi->resid;
  ep.ldr = epi->ldr;
  ep.scale = epi->scale;
  ep.scale_cols = epi->scale_cols;
  ep.seq_in = epi->seq_in;
  ep.seq_out = epi->seq_out;
  ep.seq_off = epi->seq_off;
  ep.aux = static_cast<const bf16*>(epi->aux);
  ep.ldaux = epi->ldaux;
  ep.reduce_add = 0;
  ep.colsum = nullptr;
  ep.resid_rows = epi->resid_rows;
  ep.ln_gamma = nullptr;
  ep.ln_beta = nullptr;
  ep.ln_hi = nullptr;
  ep.ln_lo = nullptr;
  ep.ln_sync = nullptr;
  if (split_out)
    ep.vec_ok = (ep.ldo % 8 == 0) && aligned16(ep.out_hi) && (!ep.out_lo || aligned16(ep.out_lo)) &&
                (!ep.bias || aligned16(ep.bias)) && (!ep.aux || (aligned16(ep.aux) && ep.ldaux % 8 == 0));
  else
    ep.vec_ok = (ep.ldo % 4 == 0) && aligned16(ep.out_f32) && (!ep.bias || aligned16(ep.bias)) &&
                (!ep.resid || ((ep.ldr % 4 == 0) && aligned16(ep.resid)));

  const int bn = gemm_tile_n(n);
  const bool precise = a_lo != nullptr;
  // CTA-pair kernel (256 x BN tiles) once there is at least one pair tile per cluster; the 1-SM kernel keeps the
  // small problems (batch-1 decode) on more SMs
  const bool pair = g_gemm_pair && bn != 128 &&
                    (g_gemm_pair == 2 ||  // 2 = forced (tests)
                     static_cast<long long>((m + 2 * BM - 1) / (2 * BM)) * ((n + bn - 1) / bn) >= num_sms() / 2);
  const int b_box = pair ? bn / 2 : bn;
  CUtensorMap a0, a1, b0, b1;
  int rc;
  if ((rc = make_tmap_bf16(&a0, a_hi, m, k, lda, BM, BK))) return rc;
  if ((rc = make_tmap_bf16(&b0, w_hi, n, k, ldw, b_box, BK))) return rc;
  if (precise) {
    if ((rc = make_tmap_bf16(&a1, a_lo, m, k, lda, BM, BK))) return rc;
    if ((rc = make_tmap_bf16(&b1, w_lo, n, k, ldw, b_box, BK))) return rc;
  } else {
    a1 = a0;
    b1 = b0;
  }
  cudaStream_t st = as_stream(stream);
  // ---- small M: split-K over all SMs (per-split partial slabs + fused finish kernel)
  if (!pair && g_gemm_splitk && epi->splitk_scratch && epi->kind <= FACT_EPI_BIAS_F32 && m <= 1024) {
    const int tiles = ((m + BM - 1) / BM) * ((n + bn - 1) / bn);
    const int num_kb = (k + BK - 1) / BK;
    // only when the tiles alone fill less than ~1/3 of the SMs (else the extra finish launch costs more than it buys)
    int splits = tiles <= 48 ? num_sms() / tiles : 1;   // at most ONE wave of CTAs (a 2nd wave doubles the kernel)
    if (splits > num_kb / 2) splits = num_kb / 2;          // at least 2 K blocks per CTA
    const size_t slab = static_cast<size_t>(m) * n * sizeof(float);
    if (splits > 1 && static_cast<size_t>(splits) * slab > epi->splitk_scratch_bytes)
      splits = static_cast<int>(epi->splitk_scratch_bytes / slab);
    if (splits > 1) {
      const int kb_per = (num_kb + splits - 1) / splits;
      splits = (num_kb + kb_per - 1) / kb_per;
      float* part = static_cast<float*>(epi->splitk_scratch);
      if (epi->ln_hi && g_gemm_finish_ln && ep.vec_ok && aligned16(epi->ln_gamma) && aligned16(epi->ln_beta) &&
          (reinterpret_cast<uintptr_t>(epi->ln_hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(epi->ln_lo) & 7) == 0) {
        ep.ln_gamma = epi->ln_gamma;   // the finish kernel normalises the rows it has just written
        ep.ln_beta = epi->ln_beta;
        ep.ln_hi = static_cast<bf16*>(epi->ln_hi);
        ep.ln_lo = static_cast<bf16*>(epi->ln_lo);
        *colsum_fused = true;
      }
      if (precise) {
        if (bn == 160) return launch_splitk<160, 2, 3>(a0, a1, b0, b1, m, n, k, splits, kb_per, epi->kind, ep, part, st);
        if (bn == 256) return launch_splitk<256, 2, 2>(a0, a1, b0, b1, m, n, k, splits, kb_per, epi->kind, ep, part, st);
        return launch_splitk<128, 2, 3>(a0, a1, b0, b1, m, n, k, splits, kb_per, epi->kind, ep, part, st);
      }
      if (bn == 160) return launch_splitk<160, 1, 6>(a0, a1, b0, b1, m, n, k, splits, kb_per, epi->kind, ep, part, st);
      if (bn == 256) return launch_splitk<256, 1, 4>(a0, a1, b0, b1, m, n, k, splits, kb_per, epi->kind, ep, part, st);
      return launch_splitk<128, 1, 6>(a0, a1, b0, b1, m, n, k, splits, kb_per, epi->kind, ep, part, st);
    }
  }
  if (pair) {
    // epilogue through shared memory + bulk tensor stores when the output is a plain 16-byte-pitched matrix (no row
    // remap); BIAS_RESID_F32 written in place becomes a bulk reduction (x += acc + bias inside the L2)
    OutMaps om;
    om.o0 = a0;
    om.o1 = a0;
    const int esz = split_out ? 2 : 4;
    const void* out0 = split_out ? static_cast<const void*>(ep.out_hi) : static_cast<const void*>(ep.out_f32);
    bool ts = g_gemm_tma_store != 0 && ep.seq_in == 0 && ep.vec_ok &&
              (static_cast<long long>(ep.ldo) * esz) % 16 == 0 && aligned16(out0);
    ep.reduce_add = 0;
    if (ts && epi->kind == FACT_EPI_BIAS_RESID_F32 && g_gemm_tma_store == 1 && ep.resid == ep.out_f32 &&
        ep.ldr == ep.ldo)
      ep.reduce_add = 1;
    if (ts) {
      if (epi->kind == FACT_EPI_GELU_GRAD && epi->colsum) {
        ep.colsum = epi->colsum;
        *colsum_fused = true;
      }
      if ((rc = make_tmap_out(&om.o0, out0, m, n, ep.ldo, esz))) return rc;
      if (split_out && ep.out_lo && (rc = make_tmap_out(&om.o1, ep.out_lo, m, n, ep.ldo, 2))) return rc;
      // LayerNorm of the finished rows inside this launch (dedicated warps) instead of a launch after it
      if (epi->ln_hi && epi->ln_sync && g_gemm_fuse_ln && epi->kind == FACT_EPI_BIAS_RESID_F32 && bn == 160 &&
          n % 160 == 0 && aligned16(epi->ln_gamma) && aligned16(epi->ln_beta) &&
          (reinterpret_cast<uintptr_t>(epi->ln_hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(epi->ln_lo) & 7) == 0) {
        ep.ln_gamma = epi->ln_gamma;
        ep.ln_beta = epi->ln_beta;
        ep.ln_hi = static_cast<bf16*>(epi->ln_hi);
        ep.ln_lo = static_cast<bf16*>(epi->ln_lo);
        ep.ln_sync = epi->ln_sync;
        *colsum_fused = true;
        if (precise)
          return launch_cfg2<160, 2, 3, FACT_EPI_BIAS_RESID_F32, true, true>(a0, a1, b0, b1, om, m, n, k, ep, st);
        return launch_cfg2<160, 1, 7, FACT_EPI_BIAS_RESID_F32, true, true>(a0, a1, b0, b1, om, m, n, k, ep, st);
      }
      if (precise) {
        if (bn == 160) return launch_epi2<160, 2, 3, true>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
        return launch_epi2<256, 2, 3, true>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
      }
      if (bn == 160) return launch_epi2<160, 1, 7, true>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
      return launch_epi2<256, 1, 6, true>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
    }
    if (precise) {
      if (bn == 160) return launch_epi2<160, 2, 4, false>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
      return launch_epi2<256, 2, 3, false>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
    }
    if (bn == 160) return launch_epi2<160, 1, 8, false>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
    return launch_epi2<256, 1, 6, false>(epi->kind, a0, a1, b0, b1, om, m, n, k, ep, st);
  }
  if (precise) {
    if (bn == 160) return launch_epi<160, 2, 3>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
    if (bn == 256) return launch_epi<256, 2, 2>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
    return launch_epi<128, 2, 3>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
  }
  if (bn == 160) return launch_epi<160, 1, 6>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
  if (bn == 256) return launch_epi<256, 1, 4>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
  return launch_epi<128, 1, 6>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
}
