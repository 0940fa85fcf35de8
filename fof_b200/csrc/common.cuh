// Shared plumbing of libfofb200: error handling, grow-only device buffers, the handle.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <map>

#include "../../include/fofb.h"
#include "cosmo.cuh"

namespace fofb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FOFB_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw ::fofb::Error(FOFB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +   \
                                             " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

#define FOFB_REQUIRE(cond, msg)                                           \
  do {                                                                    \
    if (!(cond)) throw ::fofb::Error(FOFB_ERR_INVALID, std::string(msg)); \
  } while (0)

// Grow-only device buffer: benchmark steps and repeated API calls reuse the allocation.
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  T* alloc(size_t n) {
    if (n > cap) {
      if (p) cudaFree(p);
      p = nullptr;
      cap = 0;
      size_t want = n + n / 8 + 64;
      FOFB_CUDA(cudaMalloc(&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  ~DBuf() { release(); }
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};

// Pinned host staging buffer (grow-only) for small device->host result reads.
template <class T>
struct HBuf {
  T* p = nullptr;
  size_t cap = 0;
  T* alloc(size_t n) {
    if (n > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      size_t want = n + n / 8 + 64;
      FOFB_CUDA(cudaMallocHost(&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
  ~HBuf() {
    if (p) cudaFreeHost(p);
  }
};

struct View {  // strided float64 matrix resident on the device
  const double* data = nullptr;
  int64_t rows = 0;
  int32_t cols = 0;
  int64_t rs = 0, cs = 0;
  __host__ __device__ __forceinline__ double at(int64_t i, int c) const { return data[i * rs + c * cs]; }
};

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

struct Catalog;    // catalog.cu
struct FofState;   // fof_stage.cu
struct CleanState; // clean.cu
struct Pipeline;   // pipeline.cu
struct Transitive; // transitive.cu
struct Bubble;     // bubble.cu
struct Matcher;    // match.cu

}  // namespace fofb

namespace fofb {
// Optional per-kernel CUDA-event timing (bench.py reads it for the roofline of the dominant kernel).
struct KernelProfile {
  bool on = false;
  struct Pending { const char* name; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, std::pair<double, int64_t>> acc;  // name -> (ms, launches)
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void resolve() {  // call after the stream is synchronised
    for (auto& p : pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { auto& r = acc[p.name]; r.first += ms; r.second += 1; }
      pool.push_back(p.a); pool.push_back(p.b);
    }
    pending.clear();
  }
};
}  // namespace fofb

struct fofb_handle {
  int device = 0;
  fofb::KernelProfile prof;
  int sm_count = 148;
  cudaStream_t stream = nullptr;      // the stream every call runs on: own_stream, or one borrowed through fofb_set_stream
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // second half of a split upload (to_device_split)
  cudaEvent_t upload_done = nullptr, upload_go = nullptr;
  bool upload_pending = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string last_error;
  double last_ms = 0.0;
  int64_t launches = 0;       // kernels launched by the current/last compute call
  fofb::DBuf<double> cosmo_table;
  double cosmo_key[2] = {-1.0, -1.0};
  fofb::Catalog* catalog = nullptr;
  fofb::FofState* fof = nullptr;
  fofb::CleanState* clean = nullptr;
  fofb::Pipeline* pipe = nullptr;
  fofb::Transitive* trans = nullptr;
  fofb::Bubble* bubble = nullptr;
  fofb::Matcher* matcher = nullptr;
  fofb::DBuf<unsigned char> cub_tmp;
  fofb::DBuf<double> upload;      // staging for host matrices
  fofb::DBuf<double> upload2;
  fofb::HBuf<int64_t> pinned_i64;
};

namespace fofb {

// Device cosmology for the given parameters (table uploaded once per (Om0, Ode0)).
Cosmo device_cosmo(fofb_handle* h, const fofb_params& p);
Cosmo host_cosmo(const fofb_params& p);

// Bring a caller matrix onto the device (copying the underlying block if it lives on the host).
View to_device(fofb_handle* h, const fofb_matrix& m, DBuf<double>& staging);

inline void count_launch(fofb_handle* h, int k = 1) { h->launches += k; }

// FOFB_PROFILED(h, "name", kernel<<<...>>>(...)): launch + check, with event timing when enabled
#define FOFB_PROFILED(h, name, ...)                                   \
  do {                                                                \
    cudaEvent_t _ea = nullptr, _eb = nullptr;                         \
    if ((h)->prof.on) { _ea = (h)->prof.get(); _eb = (h)->prof.get(); cudaEventRecord(_ea, (h)->stream); } \
    __VA_ARGS__;                                                      \
    if ((h)->prof.on) { cudaEventRecord(_eb, (h)->stream); (h)->prof.pending.push_back({name, _ea, _eb}); } \
    ::fofb::count_launch(h);                                          \
    FOFB_CUDA(cudaGetLastError());                                    \
  } while (0)

// One device scalar read back through the handle's pinned staging word (a pageable destination makes the driver
// stage and block; the priority rounds read a dozen scalars each).
template <class T>
inline T d2h_scalar(fofb_handle* h, const T* dptr) {
  static_assert(sizeof(T) <= 16, "scalar read-back");
  int64_t* stage = h->pinned_i64.alloc(16);  // words 0-1: this helper; word 8: fofb_dist_round_advance2's flag
  FOFB_CUDA(cudaMemcpyAsync(stage, dptr, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  FOFB_CUDA(cudaStreamSynchronize(h->stream));
  T v;
  memcpy(&v, stage, sizeof(T));
  return v;
}

#define FOFB_LAUNCH_CHECK(h)          \
  do {                                \
    ::fofb::count_launch(h);          \
    FOFB_CUDA(cudaGetLastError());    \
  } while (0)

}  // namespace fofb
