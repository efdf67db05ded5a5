// Internal (non-ABI) declarations shared by the translation units of libfact_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/fact_sm100.h"

namespace fact {

extern thread_local long long g_launch_count;  // kernels launched (graph replays are added by the AR engine)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define FACT_CUDA_CHECK(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::fact::cuda_fail(_e, #expr); \
  } while (0)

// every kernel launch of the library passes through here: also feeds fact_launch_count()
#define FACT_LAUNCH_CHECK(what)                                 \
  do {                                                          \
    ++::fact::g_launch_count;                                   \
    cudaError_t _e = cudaGetLastError();                        \
    if (_e != cudaSuccess) return ::fact::cuda_fail(_e, what);  \
  } while (0)

#define FACT_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      ::fact::set_error(__VA_ARGS__); \
      return (code);                  \
    }                                 \
  } while (0)

int make_tmap_bf16(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int box_rows, int box_cols);
// fp32 / bf16 row-major output matrix -> {32, 32} box map for bulk tensor stores / reductions (gemm_tc.cu)
int make_tmap_out(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int elem_bytes);
int make_tmap_box(CUtensorMap* out, const void* ptr, int rows, int cols, int ld, int elem_bytes, int box_rows,
                  int box_cols);
// bf16 [batches, rows, cols] (row pitch ld elements, dense over the batches) as a 3-D map {cols, rows, batches} with a
// {box_cols, box_rows, 1} box, no swizzle: bulk stores of a row block that are clipped at the end of ITS clip
int make_tmap_rows3d(CUtensorMap* out, const void* ptr, int batches, int rows, int cols, int ld, int box_rows,
                     int box_cols);
int num_sms();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// fact_set_flag("pdl", 0 | 1): launch the small-batch decode chain (split-K GEMMs, finish kernels, attention core,
// LayerNorm, embedding, head) as programmatic dependent launches -- every one of those kernels calls pdl_wait()
extern int g_pdl;

// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-dependent-launch attribute when `pdl` and g_pdl
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl && g_pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace fact
