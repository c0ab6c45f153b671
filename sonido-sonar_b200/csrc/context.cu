// Context, error plumbing, memory helpers and the host-arithmetic entry points of
// the C ABI (include/sonar.h).  No kernels live here.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>

#include "common.h"

namespace sonar {

namespace {
thread_local std::string g_err;
thread_local sonar_ctx* g_cur = nullptr;
}  // namespace

int set_error(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_error(cudaError_t e, const char* what) {
  std::ostringstream os;
  os << "CUDA error: " << cudaGetErrorString(e) << " (" << what << ")";
  g_err = os.str();
  return e == cudaErrorMemoryAllocation ? SONAR_ERR_NOMEM : SONAR_ERR_CUDA;
}
const std::string& last_error_string() { return g_err; }
namespace {
thread_local sonar_ctx::ProfRec g_open{nullptr, nullptr, nullptr};
thread_local cudaStream_t g_open_st = nullptr;
}  // namespace
void prof_begin(const char* kernel, cudaStream_t st) {
  g_open = sonar_ctx::ProfRec{nullptr, nullptr, nullptr};
  if (!g_cur) return;
  g_cur->launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_cur->profiling.load(std::memory_order_relaxed)) return;
  sonar_ctx::ProfRec r{kernel, nullptr, nullptr};
  {
    std::lock_guard<std::mutex> lk(g_cur->prof_mu);  // events come from a pool: creating them per launch costs host time
    if (g_cur->prof_pool.size() >= 2) {
      r.a = g_cur->prof_pool.back();
      g_cur->prof_pool.pop_back();
      r.b = g_cur->prof_pool.back();
      g_cur->prof_pool.pop_back();
    }
  }
  if (!r.a && (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess)) {
    cudaGetLastError();
    return;
  }
  cudaEventRecord(r.a, st);
  g_open = r;
  g_open_st = st;
}
void prof_end() {
  if (!g_open.name || !g_cur) return;
  cudaEventRecord(g_open.b, g_open_st);
  std::lock_guard<std::mutex> lk(g_cur->prof_mu);
  g_cur->prof.push_back(g_open);
  g_open = sonar_ctx::ProfRec{nullptr, nullptr, nullptr};
}
void prof_count_launch() {
  if (g_cur) g_cur->launches.fetch_add(1, std::memory_order_relaxed);
}
void set_current_ctx(sonar_ctx* c) { g_cur = c; }

int DevCtx::ensure_dev(Buf& b, size_t bytes) {
  if (bytes <= b.bytes) return SONAR_OK;
  if (b.p) {
    SONAR_CUDA(cudaDeviceSynchronize());
    SONAR_CUDA(cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
  }
  const size_t want = bytes + bytes / 8 + 256;  // head-room against ping-pong regrowth
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    std::ostringstream os;
    os << "device allocation of " << want << " bytes failed: " << cudaGetErrorString(e);
    return set_error(SONAR_ERR_NOMEM, os.str());
  }
  b.bytes = want;
  return SONAR_OK;
}

int DevCtx::ensure_host(Buf& b, size_t bytes) {
  if (bytes <= b.bytes) return SONAR_OK;
  if (b.p) {
    SONAR_CUDA(cudaDeviceSynchronize());
    SONAR_CUDA(cudaFreeHost(b.p));
    b.p = nullptr;
    b.bytes = 0;
  }
  const size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMallocHost(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    std::ostringstream os;
    os << "pinned host allocation of " << want << " bytes failed: " << cudaGetErrorString(e);
    return set_error(SONAR_ERR_NOMEM, os.str());
  }
  b.bytes = want;
  return SONAR_OK;
}

}  // namespace sonar

using namespace sonar;

extern "C" {

int sonar_init(int n_devices, const int* device_ids, sonar_ctx** out) {
  if (!out) return set_error(SONAR_ERR_INVALID, "nil argument");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return set_error(SONAR_ERR_CUDA,
                     std::string("no usable CUDA device (there is no CPU fallback): ") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  std::vector<int> ids;
  if (n_devices <= 0) {
    int cur = 0;
    SONAR_CUDA(cudaGetDevice(&cur));
    ids.push_back(cur);
  } else {
    for (int i = 0; i < n_devices; i++) ids.push_back(device_ids ? device_ids[i] : i);
  }
  for (int id : ids)
    if (id < 0 || id >= count) return set_error(SONAR_ERR_INVALID, "device id out of range");
  int restore = 0;
  cudaGetDevice(&restore);
  auto* ctx = new sonar_ctx();
  ctx->devs.resize(ids.size());
  for (size_t i = 0; i < ids.size(); i++) {
    DevCtx& d = ctx->devs[i];
    d.device = ids[i];
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(d.device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, d.device)) != cudaSuccess) {
      delete ctx;
      return cuda_error(e, "cudaSetDevice");
    }
    if (prop.major < 10) {
      delete ctx;
      return set_error(SONAR_ERR_CUDA, "libsonar.so is built for sm_100a (B200) only; device is older");
    }
    int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (auto& s : d.slot) {
      if ((e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking)) != cudaSuccess ||
          (e = cudaStreamCreateWithPriority(&s.st2, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s.mid, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s.fpdone, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaStreamCreateWithFlags(&s.st3, cudaStreamNonBlocking)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming)) != cudaSuccess) {
        delete ctx;
        return cuda_error(e, "cudaStreamCreate");
      }
    }
  }
  cudaSetDevice(restore);
  *out = ctx;
  return SONAR_OK;
}

void sonar_destroy(sonar_ctx* ctx) {
  if (!ctx) return;
  int restore = 0;
  cudaGetDevice(&restore);
  if (!ctx->devs.empty()) {
    cudaSetDevice(ctx->devs[0].device);
    cudaDeviceSynchronize();
    sonar::nccl_release(ctx);
  }
  for (auto& d : ctx->devs) {
    cudaSetDevice(d.device);
    cudaDeviceSynchronize();
    for (auto& s : d.slot) {
      if (s.d_in.p) cudaFree(s.d_in.p);
      if (s.d_out.p) cudaFree(s.d_out.p);
      if (s.d_tmp.p) cudaFree(s.d_tmp.p);
      if (s.d_raw.p) cudaFree(s.d_raw.p);
      if (s.h_in.p) cudaFreeHost(s.h_in.p);
      if (s.h_out.p) cudaFreeHost(s.h_out.p);
      if (s.done) cudaEventDestroy(s.done);
      if (s.mid) cudaEventDestroy(s.mid);
      if (s.fpdone) cudaEventDestroy(s.fpdone);
      sonar::stft_workspace_release(d.device, s.st);  // the STFT kernel pair's row workspace of this stream
      sonar::stft_workspace_release(d.device, s.st2);
      if (s.st) cudaStreamDestroy(s.st);
      if (s.st2) cudaStreamDestroy(s.st2);
      if (s.fork) cudaEventDestroy(s.fork);
      if (s.join) cudaEventDestroy(s.join);
      if (s.st3) cudaStreamDestroy(s.st3);
    }
  }
  if (!ctx->devs.empty()) cudaSetDevice(ctx->devs[0].device);
  ctx->plans.clear();
  cudaSetDevice(restore);
  delete ctx;
}

const char* sonar_last_error(void) { return last_error_string().c_str(); }
int sonar_abi_version(void) { return SONAR_ABI_VERSION; }
const char* sonar_backend(void) { return "cuda-sm100a"; }

int sonar_host_alloc(sonar_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return set_error(SONAR_ERR_INVALID, "nil argument");
  SONAR_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return SONAR_OK;
}
int sonar_host_free(sonar_ctx*, void* p) {
  if (p) SONAR_CUDA(cudaFreeHost(p));
  return SONAR_OK;
}
int sonar_host_register(sonar_ctx* ctx, void* p, uint64_t bytes) {
  if (!ctx || !p) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (bytes == 0) return SONAR_OK;
  SONAR_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return SONAR_OK;
}
int sonar_host_unregister(sonar_ctx* ctx, void* p) {
  if (!ctx || !p) return set_error(SONAR_ERR_INVALID, "nil argument");
  SONAR_CUDA(cudaHostUnregister(p));
  return SONAR_OK;
}
int sonar_dev_alloc(sonar_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return set_error(SONAR_ERR_INVALID, "nil argument");
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  SONAR_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return SONAR_OK;
}
int sonar_dev_free(sonar_ctx* ctx, void* p) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  if (p) SONAR_CUDA(cudaFree(p));
  return SONAR_OK;
}
int sonar_memcpy_h2d(sonar_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  SONAR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->devs[0].slot[0].st));
  SONAR_CUDA(cudaStreamSynchronize(ctx->devs[0].slot[0].st));
  return SONAR_OK;
}
int sonar_memcpy_d2h(sonar_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  SONAR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->devs[0].slot[0].st));
  SONAR_CUDA(cudaStreamSynchronize(ctx->devs[0].slot[0].st));
  return SONAR_OK;
}
int sonar_synchronize(sonar_ctx* ctx) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  for (auto& d : ctx->devs) {
    SONAR_CUDA(cudaSetDevice(d.device));
    for (auto& s : d.slot) {
      SONAR_CUDA(cudaStreamSynchronize(s.st));
      SONAR_CUDA(cudaStreamSynchronize(s.st2));
    }
  }
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  return SONAR_OK;
}
uint64_t sonar_kernel_launches(sonar_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

void* sonar_stream(sonar_ctx* ctx) { return ctx ? (void*)ctx->devs[0].slot[0].st : nullptr; }

int sonar_profile_enable(sonar_ctx* ctx, int on) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (on) {
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    while (ctx->prof_pool.size() < 512) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return cuda_error(cudaGetLastError(), "cudaEventCreate");
      ctx->prof_pool.push_back(e);
    }
  }
  ctx->profiling.store(on != 0);
  return SONAR_OK;
}

int sonar_profile_read(sonar_ctx* ctx, sonar_kernel_time* out, int cap, int* n_out) {
  if (!ctx || !n_out) return set_error(SONAR_ERR_INVALID, "nil argument");
  int rc = sonar_synchronize(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->prof_mu);
  int n = 0;
  if (std::getenv("SONAR_PROFILE_TIMELINE") && !ctx->prof.empty()) {  // diagnostic: when each launch ran
    for (auto& r : ctx->prof) {
      float t0 = 0.f, t1 = 0.f;
      cudaEventElapsedTime(&t0, ctx->prof.front().a, r.a);
      cudaEventElapsedTime(&t1, ctx->prof.front().a, r.b);
      std::fprintf(stderr, "[timeline] %-32s %9.3f -> %9.3f ms\n", r.name, t0, t1);
    }
    cudaGetLastError();
  }
  for (auto& r : ctx->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) {
      cudaGetLastError();
      ms = 0.f;
    }
    int k = 0;
    for (; k < n; k++)
      if (std::strcmp(out[k].kernel, r.name) == 0) break;
    if (k == n) {
      if (n >= cap) continue;
      std::memset(&out[n], 0, sizeof(out[n]));
      std::strncpy(out[n].kernel, r.name, sizeof(out[n].kernel) - 1);
      n++;
    }
    out[k].total_ms += (double)ms;
    out[k].launches += 1;
    ctx->prof_pool.push_back(r.a);
    ctx->prof_pool.push_back(r.b);
  }
  ctx->prof.clear();
  *n_out = n;
  return SONAR_OK;
}

int sonar_window_f64(int type, int size, int symmetric, int normalize, double beta, double alpha, double* out) {
  if (!out) return set_error(SONAR_ERR_INVALID, "nil argument");
  return host_window(type, size, symmetric != 0, normalize != 0, beta, alpha, out);
}

void sonar_fp_params_default(sonar_fp_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->window_size = 1024;
  p->hop_size = 256;
  p->window_type = SONAR_WINDOW_HANN;
  p->algo_sample_rate = 0;  // what stock GenerateFingerprint passes (content_config.go:87-103)
  p->call_sample_rate = 44100;
  p->energy_frame = 1024;
  p->energy_hop = 256;
  p->n_mfcc = 13;
  p->n_mel = 26;
  p->use_liftering = 1;
  p->enable = SONAR_FP_ENABLE_MFCC;
  p->lifter = 22.0;
  p->pre_emph_alpha = 0.97;
}

int sonar_fp_sizes(const sonar_fp_params* p, int64_t n, sonar_fp_sizes_t* out) { return host_fp_sizes(p, n, out); }

int sonar_fp_dev_layout(const sonar_fp_params* p, int64_t n, sonar_fp_dev_layout_t* out) {
  if (!p || !out) return set_error(SONAR_ERR_INVALID, "nil argument");
  return host_fp_layout(p, n, out);
}

int sonar_align_dtw_scalars(const sonar_dtw_out* d, int n, int m, int sr, sonar_align_result* out) {
  return host_align_dtw_scalars(d, n, m, sr, out);
}

}  // extern "C"
