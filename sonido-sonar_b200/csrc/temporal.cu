// Temporal feature group of the speech extractor (fingerprint/extractors/speech.go:370-408,587-777):
// SURVEY.md §8(f1) "next" row — not built yet.  There is deliberately no CPU fallback: asking for
// SONAR_FP_ENABLE_TEMPORAL fails loudly until the device kernels (order-statistic silence
// threshold, onset peak-pick, 512/256 envelope) exist.
#include "common.h"

namespace sonar {
int fingerprint_temporal_tail(sonar_ctx*, const double* const*, const int64_t*, int, const sonar_fp_params*,
                              sonar_fp_out*) {
  return set_error(SONAR_ERR_UNSUPPORTED, "temporal feature group (speech.go:370-408) is not implemented on the GPU path yet");
}
}  // namespace sonar
