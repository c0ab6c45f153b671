// TEMPORARY: entry points not yet implemented on the GPU return SONAR_ERR_UNSUPPORTED (never a CPU fallback).
#include "common.h"
using namespace sonar;
#define TODO(name) return set_error(SONAR_ERR_UNSUPPORTED, name " is not implemented yet")
namespace sonar {
int fingerprint_temporal_tail(sonar_ctx*, const double* const*, const int64_t*, int, const sonar_fp_params*, sonar_fp_out*) { TODO("temporal feature group"); }
}
extern "C" {
int sonar_xcorr_ncc_f64(sonar_ctx*, const double*, int64_t, const double*, int64_t, int, double*, sonar_xcorr_summary*) { TODO("sonar_xcorr_ncc_f64"); }
int sonar_xcorr_batch_f64(sonar_ctx*, const double* const*, const int64_t*, const double* const*, const int64_t*, int, int, double* const*, sonar_xcorr_summary*) { TODO("sonar_xcorr_batch_f64"); }
int sonar_xcorr_batch_dev(sonar_ctx*, const double*, int64_t, const double*, int64_t, int, int, double*, sonar_xcorr_summary*) { TODO("sonar_xcorr_batch_dev"); }
int sonar_xcorr_shard_open(sonar_ctx*, const double*, int64_t, const double*, int64_t, int, int64_t, int64_t, sonar_xcorr_shard**, sonar_xcorr_shard_peak*) { TODO("sonar_xcorr_shard_open"); }
int sonar_xcorr_shard_metrics_f64(sonar_xcorr_shard*, int64_t, sonar_xcorr_shard_metrics*) { TODO("sonar_xcorr_shard_metrics_f64"); }
int sonar_xcorr_shard_corr(sonar_xcorr_shard*, double*) { TODO("sonar_xcorr_shard_corr"); }
void sonar_xcorr_shard_close(sonar_xcorr_shard*) {}
int sonar_xcorr_merge_peaks(const sonar_xcorr_shard_peak*, int, int64_t*) { TODO("sonar_xcorr_merge_peaks"); }
int sonar_xcorr_merge_metrics(const sonar_xcorr_shard_metrics*, int, int64_t, int64_t, int, int64_t, sonar_xcorr_summary*) { TODO("sonar_xcorr_merge_metrics"); }
int sonar_align_xcorr_f64(sonar_ctx*, const double*, int64_t, const double*, int64_t, int, int, int, double*, sonar_xcorr_summary*, sonar_align_result*) { TODO("sonar_align_xcorr_f64"); }
int sonar_dtw_f64(sonar_ctx*, const double*, int, const double*, int, int, int, int, int, sonar_dtw_out*) { TODO("sonar_dtw_f64"); }
int sonar_dtw_batch_f64(sonar_ctx*, const double* const*, const double* const*, int, int, int, int, int, int, int, sonar_dtw_out*) { TODO("sonar_dtw_batch_f64"); }
int sonar_colstats_cosine_f64(sonar_ctx*, const double*, int64_t, const double*, int64_t, int, double*) { TODO("sonar_colstats_cosine_f64"); }
int sonar_colstats_f64(sonar_ctx*, const double*, int64_t, int, double*) { TODO("sonar_colstats_f64"); }
int sonar_compare_f64(sonar_ctx*, const sonar_cmp_features*, const sonar_cmp_features*, const sonar_cmp_weights*, int, sonar_cmp_result*) { TODO("sonar_compare_f64"); }
}
