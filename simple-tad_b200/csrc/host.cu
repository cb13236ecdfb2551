// Host-side utilities of libstad.so: thread-local error message, device query, TMA descriptor encoding.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.h"

namespace stad {

namespace {
thread_local char g_err[512] = "";
int g_sm_count[64] = {};
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  // The driver symbol is fetched through the runtime: libstad.so has no link-time dependency on libcuda.
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}
}  // namespace

// ---------------------------------------------------------------- per-launch event timing
namespace {
struct ProfEntry {
  stad_profile_record rec;
  cudaEvent_t e0, e1;
};
std::mutex g_prof_mutex;
std::vector<ProfEntry> g_prof;   // pre-created event pairs
int g_prof_used = 0;
bool g_prof_on = false;
bool g_prof_open = false;        // a begin without its end yet
}  // namespace

void prof_begin(int kind, int epi, int m, int n, int k, cudaStream_t stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (g_prof_used >= static_cast<int>(g_prof.size())) return;
  ProfEntry& e = g_prof[g_prof_used];
  e.rec.kind = kind;
  e.rec.epi = epi;
  e.rec.m = m;
  e.rec.n = n;
  e.rec.k = k;
  e.rec.ms = 0.f;
  cudaEventRecord(e.e0, stream);
  g_prof_open = true;
}

void prof_end(cudaStream_t stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (!g_prof_open) return;
  cudaEventRecord(g_prof[g_prof_used].e1, stream);
  ++g_prof_used;
  g_prof_open = false;
}

int prof_enable(int capacity) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  for (auto& e : g_prof) {
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
  }
  g_prof.clear();
  g_prof_used = 0;
  g_prof_open = false;
  g_prof_on = capacity > 0;
  if (capacity > 0) {
    g_prof.resize(capacity);
    for (auto& e : g_prof) {
      if (cudaEventCreate(&e.e0) != cudaSuccess || cudaEventCreate(&e.e1) != cudaSuccess)
        return fail(STAD_E_CUDA, "stad_profile_enable: cudaEventCreate failed");
    }
  }
  return STAD_OK;
}

int prof_read(stad_profile_record* out, int max_records) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  int n = g_prof_used < max_records ? g_prof_used : max_records;
  for (int i = 0; i < n; ++i) {
    ProfEntry& e = g_prof[i];
    if (cudaEventSynchronize(e.e1) != cudaSuccess) return fail(STAD_E_CUDA, "stad_profile_read: event sync failed");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.e0, e.e1);
    e.rec.ms = ms;
    out[i] = e.rec;
  }
  g_prof_used = 0;
  return n;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const char* last_error() { return g_err; }

bool pdl_enabled() { return true; }  // programmatic dependent launch on every GEMM / attention / finalize launch

// SM count of the CURRENT device (cached per device: one process may drive several GPUs through the C ABI)
int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int n = g_sm_count[dev];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return n;
}

int make_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return fail(STAD_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (reinterpret_cast<uintptr_t>(gptr) & 15) return fail(STAD_E_ALIGN, "tensor map base must be 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] & 15) return fail(STAD_E_ALIGN, "tensor map stride %d (%llu B) not a multiple of 16", i,
                                        (unsigned long long)gstr[i - 1]);
    }
  }
  if (swizzle_bytes != 128 && swizzle_bytes != 64 && swizzle_bytes != 32)
    return fail(STAD_E_SHAPE, "tensor map: swizzle %d unsupported", swizzle_bytes);
  if (box[0] * 2 != (uint32_t)swizzle_bytes)
    return fail(STAD_E_SHAPE, "tensor map: inner box (%u B) must equal the swizzle span (%d B)", box[0] * 2, swizzle_bytes);
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr,
                        bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(STAD_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)", (int)r,
                rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return STAD_OK;
}

}  // namespace stad
