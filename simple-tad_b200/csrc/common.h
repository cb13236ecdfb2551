// Host-side plumbing shared by every translation unit of libstad.so: error reporting, tensor-map
// encoding (driver entry point fetched through the runtime so the library links only libcudart).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stad.h"

namespace stad {

typedef __nv_bfloat16 bf16;

// Sets the thread-local message returned by stad_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);
const char* last_error();

#define STAD_CHECK_ARG(cond, ...)                        \
  do {                                                   \
    if (!(cond)) return ::stad::fail(STAD_E_SHAPE, __VA_ARGS__); \
  } while (0)

#define STAD_CUDA_OK(expr)                                                                                \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return ::stad::fail(STAD_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Launch-time check: catches bad configuration (too much smem, ...) without synchronising.
#define STAD_LAUNCH_OK(what)                                                                      \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) return ::stad::fail(STAD_E_CUDA, "%s launch: %s", what, cudaGetErrorString(_e)); \
  } while (0)

int sm_count();  // SMs of the current device (cached)

// Programmatic dependent launch: the kernel may start (prologue: barrier init, TMEM allocation, descriptor prefetch)
// while the kernel before it in the stream is still draining; it must execute griddepcontrol.wait (pdl_wait() in
// ptx.cuh) before it touches any global memory.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// rank-N bf16 tensor map. dims/strides innermost-first; strides in BYTES for dims 1..rank-1.
// swizzle_bytes: 128 (default) or 32.  NOTE: TMA pads a box whose inner extent is narrower than the swizzle span up
// to the span, so the inner box extent must equal the span for a dense shared-memory tile.
int make_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes = 128);

// Per-launch event timing (stad_profile_*). ProfScope records an event pair around one launch when enabled.
void prof_begin(int kind, int epi, int m, int n, int k, cudaStream_t stream);
void prof_end(cudaStream_t stream);
int prof_enable(int capacity);
int prof_read(stad_profile_record* out, int max_records);
struct ProfScope {
  cudaStream_t s;
  ProfScope(int kind, int epi, int m, int n, int k, cudaStream_t stream) : s(stream) { prof_begin(kind, epi, m, n, k, s); }
  ~ProfScope() { prof_end(s); }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace stad
