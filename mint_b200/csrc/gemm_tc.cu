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
constexpr int GEMM_THREADS = 192;
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
};

template <int BN, int NPART, int STAGES>
struct GemmCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = NPART * (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + BAR_BYTES;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(STAGE_BYTES >= 4 * 32 * 33 * 4, "epilogue scratch aliases stage 0");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
};

__device__ __forceinline__ int map_row(int row, const EpiArgs& e) {
  if (e.seq_in == 0) return row;
  return (row / e.seq_in) * e.seq_out + e.seq_off + (row % e.seq_in);
}

template <int BN, int NPART, int STAGES, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_tc_kernel(
    const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, int M, int N, int K,
    EpiArgs ep) {
  using Cfg = GemmCfg<BN, NPART, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier block: full[STAGES] | empty[STAGES] | tmem_full | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
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
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_addr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto sA = [&](int s, int part) { return smem_base + s * Cfg::STAGE_BYTES + part * A_TILE_BYTES; };
  auto sB = [&](int s, int part) {
    return smem_base + s * Cfg::STAGE_BYTES + NPART * A_TILE_BYTES + part * Cfg::B_TILE_BYTES;
  };

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
        tma_load_2d(sA(s, 0), &tmA0, kb * BK, m0, full_bar(s));
        tma_load_2d(sB(s, 0), &tmB0, kb * BK, n0, full_bar(s));
        if (NPART == 2) {
          tma_load_2d(sA(s, 1), &tmA1, kb * BK, m0, full_bar(s));
          tma_load_2d(sB(s, 1), &tmB1, kb * BK, n0, full_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint32_t koff = kk * UMMA_K * 2;  // bytes inside the 128-B swizzle span
          const uint64_t a_hi = umma_desc_k_sw128(sA(s, 0) + koff);
          const uint64_t b_hi = umma_desc_k_sw128(sB(s, 0) + koff);
          umma_bf16(tmem_base, a_hi, b_hi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          if (NPART == 2) {
            const uint64_t a_lo = umma_desc_k_sw128(sA(s, 1) + koff);
            const uint64_t b_lo = umma_desc_k_sw128(sB(s, 1) + koff);
            umma_bf16(tmem_base, a_lo, b_hi, idesc, 1u);
            umma_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));  // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);
    }
  } else {  // ---------------- epilogue warps 2..5 ; TMEM lane quadrant = warp % 4
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // all MMAs (hence all TMA loads) are complete: stage 0 is free to serve as the transpose scratch
    float* scratch = reinterpret_cast<float*>(smem_gen) + q * (32 * 33);
    const int row_base = m0 + q * 32;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) scratch[lane * 33 + c] = v[c];
      __syncwarp();
      if (EPI == FACT_EPI_BIAS_RESID_F32 || EPI == FACT_EPI_BIAS_F32) {
        const int col = n0 + c0 + lane;
        const bool cok = col < N;
        const float b = (cok && ep.bias) ? ep.bias[col] : 0.f;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int row = row_base + r;
          if (row < M && cok) {
            float x = scratch[r * 33 + lane] + b;
            if (EPI == FACT_EPI_BIAS_RESID_F32) x += ep.resid[static_cast<size_t>(row) * ep.ldr + col];
            ep.out_f32[static_cast<size_t>(map_row(row, ep)) * ep.ldo + col] = x;
          }
        }
      } else {
        const int cp = (lane & 15) * 2;
        const int col = n0 + c0 + cp;
        float b0 = 0.f, b1 = 0.f;
        if (EPI == FACT_EPI_BIAS_GELU_SPLIT) {
          if (col < N) b0 = ep.bias[col];
          if (col + 1 < N) b1 = ep.bias[col + 1];
        }
        const float s0 = (EPI == FACT_EPI_SPLIT && col < ep.scale_cols) ? ep.scale : 1.f;
        const float s1 = (EPI == FACT_EPI_SPLIT && col + 1 < ep.scale_cols) ? ep.scale : 1.f;
#pragma unroll 4
        for (int rr = 0; rr < 16; ++rr) {
          const int r = rr * 2 + (lane >> 4);
          const int row = row_base + r;
          if (row < M && col < N) {
            float x0 = scratch[r * 33 + cp], x1 = scratch[r * 33 + cp + 1];
            if (EPI == FACT_EPI_BIAS_GELU_SPLIT) {
              x0 = gelu_tanh(x0 + b0);
              x1 = gelu_tanh(x1 + b1);
            } else {
              x0 *= s0;
              x1 *= s1;
            }
            bf16 h0, l0, h1, l1;
            split_bf16(x0, h0, l0);
            split_bf16(x1, h1, l1);
            const size_t o = static_cast<size_t>(row) * ep.ldo + col;
            if (col + 1 < N) {
              *reinterpret_cast<uint32_t*>(ep.out_hi + o) = pack_bf16x2(h0, h1);
              if (ep.out_lo) *reinterpret_cast<uint32_t*>(ep.out_lo + o) = pack_bf16x2(l0, l1);
            } else {
              ep.out_hi[o] = h0;
              if (ep.out_lo) ep.out_lo[o] = l0;
            }
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
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
  int rows, cols, ld, box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) ^ (static_cast<size_t>(k.cols) << 20) ^
         (static_cast<size_t>(k.ld) << 40) ^ (static_cast<size_t>(k.box_rows) << 52);
    return h;
  }
};

// bf16 row-major [rows, cols] with row pitch ld elements -> 2-D tiled map, box {64 cols, box_rows}, 128B swizzle.
static int make_tmap(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int box_rows) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, box_rows};
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
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FACT_REQUIRE(r == CUDA_SUCCESS, FACT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%d",
               static_cast<int>(r), rows, cols, ld, box_rows);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return FACT_OK;
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
  dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM);
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(a0, a1, b0, b1, m, n, k, ep);
  FACT_LAUNCH_CHECK("gemm_tc_kernel launch");
  return FACT_OK;
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
  }
  set_error("unknown epilogue kind %d", kind);
  return FACT_ERR_UNSUPPORTED;
}

int gemm_tile_n(int n) {
  if (n % 160 == 0) return 160;  // 2400 = 15 x 160, 800 = 5 x 160
  if (n % 256 == 0) return 256;  // 3072 = 12 x 256
  return 128;
}

}  // namespace fact

using namespace fact;

extern "C" int fact_gemm(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo, int ldw,
                         int m, int n, int k, const fact_gemm_epilogue* epi, void* stream) {
  FACT_REQUIRE(a_hi && w_hi && epi, FACT_ERR_BAD_SHAPE, "fact_gemm: null operand");
  FACT_REQUIRE(m > 0 && n > 0 && k > 0, FACT_ERR_BAD_SHAPE, "fact_gemm: bad shape m=%d n=%d k=%d", m, n, k);
  FACT_REQUIRE((a_lo == nullptr) == (w_lo == nullptr), FACT_ERR_BAD_SHAPE,
               "fact_gemm: a_lo and w_lo must both be given (precise) or both be NULL (bf16)");
  const bool split_out = epi->kind == FACT_EPI_SPLIT || epi->kind == FACT_EPI_BIAS_GELU_SPLIT;
  if (split_out) {
    FACT_REQUIRE(epi->out_hi != nullptr && (epi->ldo % 2) == 0, FACT_ERR_BAD_SHAPE,
                 "fact_gemm: split epilogue needs out_hi and an even ldo");
    FACT_REQUIRE(epi->kind != FACT_EPI_BIAS_GELU_SPLIT || epi->bias, FACT_ERR_BAD_SHAPE, "gelu epilogue needs bias");
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

  const int bn = gemm_tile_n(n);
  const bool precise = a_lo != nullptr;
  CUtensorMap a0, a1, b0, b1;
  int rc;
  if ((rc = make_tmap(&a0, a_hi, m, k, lda, BM))) return rc;
  if ((rc = make_tmap(&b0, w_hi, n, k, ldw, bn))) return rc;
  if (precise) {
    if ((rc = make_tmap(&a1, a_lo, m, k, lda, BM))) return rc;
    if ((rc = make_tmap(&b1, w_lo, n, k, ldw, bn))) return rc;
  } else {
    a1 = a0;
    b1 = b0;
  }
  cudaStream_t st = as_stream(stream);
  if (precise) {
    if (bn == 160) return launch_epi<160, 2, 3>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
    if (bn == 256) return launch_epi<256, 2, 2>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
    return launch_epi<128, 2, 3>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
  }
  if (bn == 160) return launch_epi<160, 1, 3>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
  if (bn == 256) return launch_epi<256, 1, 4>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
  return launch_epi<128, 1, 3>(epi->kind, a0, a1, b0, b1, m, n, k, ep, st);
}
