// Lag-range sharding of ONE long cross-correlation over the ranks of a job, entirely inside the library:
//   z-score both sequences (every rank; a dependent-add chain in the reference's order, it does not shard) ->
//   this rank's lags [lo, hi) of CrossCorrelation.computeTimeDomain (correlation.go:373-449) in reference order ->
//   ONE ncclAllGather of the curve shards, device buffers, on the library's stream (in place: the rank's shard already
//   sits at its final position of the full curve) -> findPeak / SNR / sharpness / second peak / side lobe
//   (correlation.go:526-661) over the full curve on every rank.
// No host round trip between the phases (r01 did two torch.distributed all_gathers from Python with .cpu() hops and was
// 2.4 x SLOWER on 8 ranks than one GPU).  NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already
// loaded in a Python host, the system one in a Go host), so libsonar.so has no link-time dependency on it and the
// single-GPU paths never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "common.h"

namespace sonar {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.handle, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
    return a;
  }();
  return api;
}

int nccl_error(ncclResult_t r, const char* what) {
  return set_error(SONAR_ERR_CUDA, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
}
#define SONAR_NCCL(call)                                 \
  do {                                                   \
    ncclResult_t r__ = (call);                           \
    if (r__ != ncclSuccess) return nccl_error(r__, #call); \
  } while (0)

inline int64_t even2(int64_t n) { return (n + 1) & ~(int64_t)1; }

}  // namespace

void nccl_release(sonar_ctx* ctx) {
  if (ctx->nccl_comm && nccl().ok) nccl().CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  ctx->nccl_world = 1;
  ctx->nccl_rank = 0;
  if (ctx->shard_buf) cudaFree(ctx->shard_buf);
  ctx->shard_buf = nullptr;
  ctx->shard_bytes = 0;
}

}  // namespace sonar

using namespace sonar;

extern "C" {

int sonar_nccl_unique_id(unsigned char* id, int cap) {
  if (!id || cap < (int)sizeof(ncclUniqueId)) return set_error(SONAR_ERR_INVALID, "id buffer must hold 128 bytes");
  if (!nccl().ok) return set_error(SONAR_ERR_UNSUPPORTED, "libnccl.so.2 not found");
  ncclUniqueId u;
  SONAR_NCCL(nccl().GetUniqueId(&u));
  std::memcpy(id, &u, sizeof(u));
  return SONAR_OK;
}

int sonar_nccl_init(sonar_ctx* ctx, int world, int rank, const unsigned char* id) {
  if (!ctx || !id || world <= 0 || rank < 0 || rank >= world) return set_error(SONAR_ERR_INVALID, "bad communicator arguments");
  if (!nccl().ok) return set_error(SONAR_ERR_UNSUPPORTED, "libnccl.so.2 not found");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  SONAR_CUDA(cudaSetDevice(ctx->devs[0].device));
  nccl_release(ctx);
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  SONAR_NCCL(nccl().CommInitRank(&comm, world, u, rank));
  ctx->nccl_comm = comm;
  ctx->nccl_world = world;
  ctx->nccl_rank = rank;
  return SONAR_OK;
}

int sonar_nccl_shutdown(sonar_ctx* ctx) {
  if (!ctx) return set_error(SONAR_ERR_INVALID, "nil argument");
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  cudaSetDevice(ctx->devs[0].device);
  nccl_release(ctx);
  return SONAR_OK;
}

int sonar_xcorr_lag_sharded(sonar_ctx* ctx, const double* a, int64_t na, const double* b, int64_t nb, int max_lag,
                            int inputs_on_device, double* corr_host, sonar_xcorr_summary* out) {
  if (!ctx || !out) return set_error(SONAR_ERR_INVALID, "nil argument");
  if (!a || !b || na <= 0 || nb <= 0) return set_error(SONAR_ERR_EMPTY, "empty signals provided");  // correlation.go:132
  std::lock_guard<std::mutex> call_lock(ctx->call_mu);
  set_current_ctx(ctx);
  DevCtx& dev = ctx->devs[0];
  SONAR_CUDA(cudaSetDevice(dev.device));
  cudaStream_t st = dev.slot[0].st;
  const int world = ctx->nccl_comm ? ctx->nccl_world : 1, rank = ctx->nccl_comm ? ctx->nccl_rank : 0;
  const int aml = actual_max_lag(max_lag, na, nb);
  const int64_t nl = 2 * (int64_t)aml + 1;
  const int64_t step = even2((nl + world - 1) / world);  // lags per rank (even: 16-byte aligned shards)
  const int64_t lo = std::min<int64_t>(nl, (int64_t)rank * step), hi = std::min<int64_t>(nl, lo + step);
  // workspace: a | b | za | zb | curve (world * step) | descriptors
  const size_t doubles = (size_t)(2 * (even2(na) + even2(nb)) + world * step + 2);
  const size_t desc = sizeof(XcorrSeq) * 2 + sizeof(XcorrPair) * 2 + sizeof(XcorrPairOut);
  const size_t bytes = sizeof(double) * doubles + desc + 64;
  if (ctx->shard_bytes < bytes) {
    if (ctx->shard_buf) cudaFree(ctx->shard_buf);
    ctx->shard_buf = nullptr;
    ctx->shard_bytes = 0;
    SONAR_CUDA(cudaMalloc(&ctx->shard_buf, bytes));
    ctx->shard_bytes = bytes;
  }
  double* d_a = static_cast<double*>(ctx->shard_buf);
  double* d_b = d_a + even2(na);
  double* d_za = d_b + even2(nb);
  double* d_zb = d_za + even2(na);
  double* d_curve = d_zb + even2(nb);
  unsigned char* dd = reinterpret_cast<unsigned char*>(d_curve + world * step + 2);
  XcorrSeq* d_seqs = reinterpret_cast<XcorrSeq*>(dd);
  XcorrPair* d_pairs = reinterpret_cast<XcorrPair*>(dd + sizeof(XcorrSeq) * 2);
  XcorrPairOut* d_out = reinterpret_cast<XcorrPairOut*>(dd + sizeof(XcorrSeq) * 2 + sizeof(XcorrPair) * 2);
  const double* src_a = inputs_on_device ? a : d_a;
  const double* src_b = inputs_on_device ? b : d_b;
  if (!inputs_on_device) {
    SONAR_CUDA(cudaMemcpyAsync(d_a, a, sizeof(double) * (size_t)na, cudaMemcpyHostToDevice, st));
    SONAR_CUDA(cudaMemcpyAsync(d_b, b, sizeof(double) * (size_t)nb, cudaMemcpyHostToDevice, st));
  }
  const XcorrSeq seqs[2] = {{src_a, d_za, na}, {src_b, d_zb, nb}};
  const XcorrPair pairs[2] = {{d_za, d_zb, d_curve + lo, na, nb, lo, hi, aml, 0},   // this rank's shard, at its final place
                              {d_za, d_zb, d_curve, na, nb, 0, nl, aml, 0}};         // the gathered curve
  SONAR_CUDA(cudaMemcpyAsync(d_seqs, seqs, sizeof(seqs), cudaMemcpyHostToDevice, st));
  SONAR_CUDA(cudaMemcpyAsync(d_pairs, pairs, sizeof(pairs), cudaMemcpyHostToDevice, st));
  int rc;
  if ((rc = launch_znorm(d_seqs, 2, st))) return rc;
  if (hi > lo && (rc = launch_xcorr(d_pairs, 1, hi - lo, st))) return rc;
  if (world > 1)
    SONAR_NCCL(nccl().AllGather(d_curve + (int64_t)rank * step, d_curve, (size_t)step, ncclDouble,
                                static_cast<ncclComm_t>(ctx->nccl_comm), st));
  if ((rc = launch_xcorr_finalize(d_pairs + 1, 1, -1, d_out, st))) return rc;
  XcorrPairOut o;
  SONAR_CUDA(cudaMemcpyAsync(&o, d_out, sizeof(o), cudaMemcpyDeviceToHost, st));
  if (corr_host) SONAR_CUDA(cudaMemcpyAsync(corr_host, d_curve, sizeof(double) * (size_t)nl, cudaMemcpyDeviceToHost, st));
  SONAR_CUDA(cudaStreamSynchronize(st));
  summarize_xcorr(o, aml, na, nb, nl, out);
  return SONAR_OK;
}

}  // extern "C"
