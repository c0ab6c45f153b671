"""ctypes binding of include/sonar.h.

This is the Python-side mirror of the C ABI that the Go cgo shim binds (see
INTEGRATION.md and go/).  It is plumbing for the test harness, smoke() and
bench.py: `SonarLib()` loads the CUDA product library (libsonar.so, built
in-tree by __graft_entry__.build()) and fails loudly when it is missing —
there is no CPU fallback.  The same binding can be pointed at another library
implementing the same ABI by passing an explicit path; the tests use that to
drive the CPU oracle (oracle/libsonar_oracle.so) as the checker.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libsonar.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)

# sonar_status
OK, ERR_INVALID, ERR_EMPTY, ERR_TOO_SHORT, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED = range(7)

# sonar_window
WINDOWS = {
    "hann": 0, "hamming": 1, "blackman": 2, "blackman_harris": 3, "kaiser": 4,
    "tukey": 5, "rectangular": 6, "bartlett": 7, "welch": 8,
}
FP_ENABLE_MFCC, FP_ENABLE_TEMPORAL, FP_ENABLE_SPEECH = 1, 2, 4
STEP_SYMMETRIC2, STEP_ASYMMETRIC, STEP_SYMMETRIC1 = 0, 1, 2


class SonarError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code
        self.msg = msg


class FpParams(C.Structure):
    _fields_ = [
        ("window_size", C.c_int32), ("hop_size", C.c_int32), ("window_type", C.c_int32),
        ("algo_sample_rate", C.c_int32), ("call_sample_rate", C.c_int32),
        ("energy_frame", C.c_int32), ("energy_hop", C.c_int32), ("n_mfcc", C.c_int32),
        ("n_mel", C.c_int32), ("use_liftering", C.c_int32), ("enable", C.c_uint32),
        ("reserved0", C.c_int32), ("low_hz", C.c_double), ("high_hz", C.c_double),
        ("lifter", C.c_double), ("pre_emph_alpha", C.c_double),
    ]


class FpSizes(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_frames", "n_bins", "n_flux", "n_energy_frames", "n_pitch_frames", "n_mfcc", "n_envelope")]


_FP_ARRAYS = (
    "mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness",
    "spectral_crest", "spectral_slope", "spectral_flux", "zero_crossing_rate", "short_time_energy",
    "energy_entropy", "low_energy_ratio", "high_energy_ratio", "pitch_estimate", "pitch_confidence",
    "voicing_strength", "harmonic_ratio", "inharmonicity_ratio", "tonal_centroid", "rms_energy",
    "envelope_shape", "attack_time",
)
_FP_SCALARS = ("energy_variance", "loudness_range", "dynamic_range", "silence_ratio",
               "peak_amplitude", "average_amplitude", "onset_density")


class FpOut(C.Structure):
    _fields_ = ([(n, c_double_p) for n in _FP_ARRAYS] + [("attack_time_cap", C.c_int64)] +
                [(n, C.c_double) for n in _FP_SCALARS] + [("n_attack_time", C.c_int64)])


class FpDevLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "mfcc", "spectral_centroid", "spectral_rolloff", "spectral_bandwidth", "spectral_flatness",
        "spectral_crest", "spectral_slope", "spectral_flux", "zero_crossing_rate",
        "short_time_energy", "energy_entropy", "low_energy_ratio", "high_energy_ratio",
        "pitch_estimate", "pitch_confidence", "voicing_strength", "harmonic_ratio",
        "inharmonicity_ratio", "tonal_centroid", "scalars", "total")]


class XcorrSummary(C.Structure):
    _fields_ = [
        ("peak_correlation", C.c_double), ("p_value", C.c_double), ("snr", C.c_double),
        ("sharpness", C.c_double), ("second_peak", C.c_double), ("peak_to_sidelobe", C.c_double),
        ("peak_lag", C.c_int32), ("peak_index", C.c_int32), ("actual_max_lag", C.c_int32),
        ("overlap_length", C.c_int32), ("is_significant", C.c_int32), ("n_candidates", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SpeechOut(C.Structure):
    """sonar_speech_out (include/sonar.h)."""
    _fields_ = [("voicing_probability", c_double_p), ("spectral_tilt", c_double_p), ("pause_duration", c_double_p),
                ("pause_cap", C.c_int64), ("n_pause", C.c_int64), ("n_frames", C.c_int64), ("is_speech", C.c_int32),
                ("reserved", C.c_int32), ("speech_rate", C.c_double)]


class XcorrShardPeak(C.Structure):
    _fields_ = [("abs_peak", C.c_double), ("index", C.c_int64)]


class XcorrShardMetrics(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "noise_sum", "noise_count", "max_sidelobe", "second_abs", "second_val", "second_index",
        "c_peak", "c_prev", "c_next")]


class AlignResult(C.Structure):
    _fields_ = [
        ("method", C.c_int32), ("offset", C.c_int32), ("offset_seconds", C.c_double),
        ("confidence", C.c_double), ("similarity", C.c_double), ("alignment_quality", C.c_double),
        ("noise_level", C.c_double), ("stability", C.c_double), ("query_length", C.c_int32),
        ("reference_length", C.c_int32), ("sample_rate", C.c_int32), ("reserved0", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class DtwOut(C.Structure):
    _fields_ = [
        ("path_query", c_int32_p), ("path_ref", c_int32_p), ("path_cost", c_double_p),
        ("path_cap", C.c_int64), ("path_len", C.c_int64), ("distance", C.c_double),
        ("total_cost", C.c_double), ("cost_matrix", c_double_p),
    ]


class PairOut(C.Structure):
    _fields_ = [("query", FpOut), ("reference", FpOut), ("xcorr", XcorrSummary), ("corr_alignment", AlignResult),
                ("corr", c_double_p), ("dtw", DtwOut), ("dtw_length", C.c_int32), ("reserved0", C.c_int32)]


class KernelTime(C.Structure):
    _fields_ = [("kernel", C.c_char * 48), ("total_ms", C.c_double), ("launches", C.c_int64)]


class CmpFeatures(C.Structure):
    _fields_ = [
        ("mfcc", c_double_p), ("mfcc_frames", C.c_int64), ("mfcc_dim", C.c_int32),
        ("content_type", C.c_int32),
        ("spectral_centroid", c_double_p), ("n_centroid", C.c_int64),
        ("spectral_rolloff", c_double_p), ("n_rolloff", C.c_int64),
        ("spectral_flux", c_double_p), ("n_flux", C.c_int64),
        ("has_spectral", C.c_int32), ("has_harmonic", C.c_int32),
        ("harmonic_ratio", c_double_p), ("n_harmonic_ratio", C.c_int64),
        ("pitch_estimate", c_double_p), ("n_pitch", C.c_int64),
        ("rms_energy", c_double_p), ("n_rms", C.c_int64),
        ("has_temporal", C.c_int32), ("reserved0", C.c_int32),
        ("dynamic_range", C.c_double), ("silence_ratio", C.c_double), ("onset_density", C.c_double),
    ]


class CmpWeights(C.Structure):
    _fields_ = [("w", C.c_double * 7)]


class CmpResult(C.Structure):
    _fields_ = [
        ("overall_similarity", C.c_double), ("feature_similarity", C.c_double),
        ("confidence", C.c_double), ("dist_mfcc", C.c_double), ("dist_spectral", C.c_double),
        ("dist_temporal", C.c_double), ("dist_harmonic", C.c_double),
        ("content_type_match", C.c_int32), ("n_features", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/sonar.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "sonar_init", "sonar_destroy", "sonar_last_error", "sonar_abi_version", "sonar_backend",
    "sonar_host_alloc", "sonar_host_free", "sonar_host_register", "sonar_host_unregister", "sonar_dev_alloc", "sonar_dev_free", "sonar_memcpy_h2d",
    "sonar_memcpy_d2h", "sonar_synchronize", "sonar_kernel_launches", "sonar_stream",
    "sonar_profile_enable", "sonar_profile_read", "sonar_fp_exact_counts", "sonar_window_f64",
    "sonar_stft_stream_open", "sonar_stft_stream_frames", "sonar_stft_stream_buffered", "sonar_stft_stream_process",
    "sonar_stft_stream_close",
    "sonar_fp_params_default", "sonar_fp_sizes", "sonar_fingerprint_f64", "sonar_fingerprint_speech_f64",
    "sonar_fingerprint_batch_f64", "sonar_fingerprint_batch_pcm", "sonar_fingerprint_batch_dev", "sonar_fp_dev_layout",
    "sonar_stft_f64", "sonar_xcorr_ncc_f64", "sonar_xcorr_batch_f64", "sonar_xcorr_batch_dev",
    "sonar_xcorr_shard_open", "sonar_xcorr_shard_metrics_f64", "sonar_xcorr_shard_corr",
    "sonar_xcorr_shard_close", "sonar_xcorr_merge_peaks", "sonar_xcorr_merge_metrics",
    "sonar_nccl_unique_id", "sonar_nccl_init", "sonar_nccl_shutdown", "sonar_xcorr_lag_sharded",
    "sonar_truncate_to_alignment",
    "sonar_align_xcorr_f64", "sonar_dtw_f64", "sonar_dtw_batch_f64", "sonar_align_dtw_scalars",
    "sonar_colstats_cosine_f64", "sonar_colstats_f64", "sonar_compare_f64",
    "sonar_music_spectral_f64", "sonar_compare_batch_f64", "sonar_align_pairs_sizes", "sonar_align_pairs_f64", "sonar_align_pairs_pcm", "sonar_align_pairs_dev",
)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a: np.ndarray | None):
    return a.ctypes.data_as(c_double_p) if a is not None else None


@dataclass
class Fingerprint:
    """Flat mirror of extractors.ExtractedFeatures (fingerprint/extractors/features.go:5-124)."""
    sizes: dict
    arrays: dict = field(default_factory=dict)
    scalars: dict = field(default_factory=dict)

    def __getattr__(self, k):
        d = self.__dict__
        if k in d.get("arrays", {}):
            return d["arrays"][k]
        if k in d.get("scalars", {}):
            return d["scalars"][k]
        raise AttributeError(k)


class StftStream:
    """STFTStreamer (analyzers/spectral.go:312-374) over sonar_stft_stream_*."""

    def __init__(self, lib, handle, win):
        self.lib, self.h, self.bins = lib, handle, win // 2 + 1

    def buffered(self) -> int:
        return int(self.lib.lib.sonar_stft_stream_buffered(self.h))

    def process_chunk(self, chunk, phase=True, cplx=True):
        """-> (magnitude [T][B], phase [T][B] | None, complex [T][B][2] | None); T may be 0."""
        chunk = _f64(chunk)
        T = int(self.lib.lib.sonar_stft_stream_frames(self.h, chunk.size))
        mag = np.zeros((T, self.bins))
        ph = np.zeros((T, self.bins)) if phase else None
        cx = np.zeros((T, self.bins, 2)) if cplx else None
        n = C.c_int64()
        self.lib._chk(self.lib.lib.sonar_stft_stream_process(self.h, _dp(chunk), chunk.size, _dp(mag), _dp(ph), _dp(cx), T,
                                                             C.byref(n)))
        assert n.value == T
        return mag, ph, cx

    def close(self):
        if self.h:
            self.lib.lib.sonar_stft_stream_close(self.h)
            self.h = None


class SonarLib:
    """One loaded implementation of include/sonar.h plus one sonar_ctx."""

    def __init__(self, path: str | None = None, n_devices: int = 0, init: bool = True):
        path = path or PRODUCT_LIB
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " — there is no CPU fallback")
        self.path = path
        self.lib = L = C.CDLL(path)
        L.sonar_last_error.restype = C.c_char_p
        L.sonar_backend.restype = C.c_char_p
        L.sonar_kernel_launches.restype = C.c_uint64
        L.sonar_kernel_launches.argtypes = [C.c_void_p]
        L.sonar_stream.restype = C.c_void_p
        L.sonar_stream.argtypes = [C.c_void_p]
        L.sonar_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.sonar_profile_read.argtypes = [C.c_void_p, C.POINTER(KernelTime), C.c_int, C.POINTER(C.c_int)]
        if hasattr(L, "sonar_fp_exact_counts"):
            L.sonar_fp_exact_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.sonar_destroy.restype = None
        L.sonar_destroy.argtypes = [C.c_void_p]
        L.sonar_xcorr_shard_close.restype = None
        L.sonar_xcorr_shard_close.argtypes = [C.c_void_p]
        L.sonar_fp_params_default.restype = None
        L.sonar_stft_stream_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sonar_stft_stream_frames.restype = C.c_int64
        L.sonar_stft_stream_frames.argtypes = [C.c_void_p, C.c_int64]
        L.sonar_stft_stream_buffered.restype = C.c_int64
        L.sonar_stft_stream_buffered.argtypes = [C.c_void_p]
        L.sonar_stft_stream_process.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p,
                                                C.c_int64, C.POINTER(C.c_int64)]
        L.sonar_stft_stream_close.restype = None
        L.sonar_stft_stream_close.argtypes = [C.c_void_p]
        L.sonar_init.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
        L.sonar_synchronize.argtypes = [C.c_void_p]
        L.sonar_host_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.sonar_host_free.argtypes = [C.c_void_p, C.c_void_p]
        L.sonar_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.sonar_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
        L.sonar_dev_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.sonar_dev_free.argtypes = [C.c_void_p, C.c_void_p]
        L.sonar_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.sonar_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.sonar_window_f64.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, c_double_p]
        L.sonar_fp_sizes.argtypes = [C.POINTER(FpParams), C.c_int64, C.POINTER(FpSizes)]
        L.sonar_fingerprint_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.POINTER(FpParams), C.POINTER(FpOut)]
        L.sonar_fingerprint_batch_f64.argtypes = [C.c_void_p, C.POINTER(c_double_p), c_int64_p, C.c_int,
                                                  C.POINTER(FpParams), C.POINTER(FpOut)]
        L.sonar_fingerprint_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                                  C.POINTER(FpParams), C.c_void_p]
        L.sonar_fp_dev_layout.argtypes = [C.POINTER(FpParams), C.c_int64, C.POINTER(FpDevLayout)]
        L.sonar_stft_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     c_double_p, c_double_p, c_double_p]
        L.sonar_xcorr_ncc_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int,
                                          c_double_p, C.POINTER(XcorrSummary)]
        L.sonar_xcorr_batch_f64.argtypes = [C.c_void_p, C.POINTER(c_double_p), c_int64_p, C.POINTER(c_double_p),
                                            c_int64_p, C.c_int, C.c_int, C.POINTER(c_double_p),
                                            C.POINTER(XcorrSummary)]
        L.sonar_xcorr_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                            C.c_int, C.c_void_p, C.POINTER(XcorrSummary)]
        L.sonar_xcorr_shard_open.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int,
                                             C.c_int64, C.c_int64, C.POINTER(C.c_void_p),
                                             C.POINTER(XcorrShardPeak)]
        L.sonar_xcorr_shard_metrics_f64.argtypes = [C.c_void_p, C.c_int64, C.POINTER(XcorrShardMetrics)]
        L.sonar_xcorr_shard_corr.argtypes = [C.c_void_p, c_double_p]
        L.sonar_xcorr_merge_peaks.argtypes = [C.POINTER(XcorrShardPeak), C.c_int, c_int64_p]
        L.sonar_nccl_unique_id.argtypes = [C.c_char_p, C.c_int]
        L.sonar_nccl_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        L.sonar_nccl_shutdown.argtypes = [C.c_void_p]
        L.sonar_xcorr_lag_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                              c_double_p, C.POINTER(XcorrSummary)]
        L.sonar_xcorr_merge_metrics.argtypes = [C.POINTER(XcorrShardMetrics), C.c_int, C.c_int64, C.c_int64,
                                                C.c_int, C.c_int64, C.POINTER(XcorrSummary)]
        L.sonar_align_xcorr_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int,
                                            C.c_int, C.c_int, c_double_p, C.POINTER(XcorrSummary),
                                            C.POINTER(AlignResult)]
        L.sonar_dtw_f64.argtypes = [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.POINTER(DtwOut)]
        L.sonar_dtw_batch_f64.argtypes = [C.c_void_p, C.POINTER(c_double_p), C.POINTER(c_double_p), C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(DtwOut)]
        L.sonar_align_dtw_scalars.argtypes = [C.POINTER(DtwOut), C.c_int, C.c_int, C.c_int, C.POINTER(AlignResult)]
        L.sonar_colstats_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int, c_double_p]
        L.sonar_colstats_cosine_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64,
                                                C.c_int, c_double_p]
        L.sonar_compare_f64.argtypes = [C.c_void_p, C.POINTER(CmpFeatures), C.POINTER(CmpFeatures),
                                        C.POINTER(CmpWeights), C.c_int, C.POINTER(CmpResult)]
        L.sonar_align_pairs_sizes.argtypes = [C.POINTER(FpParams), C.c_int64, C.c_double, c_int32_p, c_int32_p]
        L.sonar_align_pairs_f64.argtypes = [C.c_void_p, C.POINTER(c_double_p), C.POINTER(c_double_p), C.c_int64, C.c_int,
                                            C.POINTER(FpParams), C.c_double, C.c_int, C.POINTER(PairOut)]
        L.sonar_music_spectral_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               c_double_p, c_double_p, C.c_int, C.c_double, C.c_double, c_double_p]
        L.sonar_align_pairs_pcm.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int64,
                                            C.c_int, C.POINTER(FpParams), C.c_double, C.c_int, C.POINTER(PairOut)]
        L.sonar_align_pairs_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.POINTER(FpParams),
                                            C.c_double, C.c_int, C.POINTER(PairOut)]
        self.ctx = C.c_void_p()
        if init:
            self._chk(L.sonar_init(n_devices, None, C.byref(self.ctx)))

    # -- helpers ---------------------------------------------------------------
    def _chk(self, rc: int):
        if rc != 0:
            raise SonarError(rc, (self.lib.sonar_last_error() or b"").decode())

    def close(self):
        if self.ctx:
            self.lib.sonar_destroy(self.ctx)
            self.ctx = C.c_void_p()

    @property
    def backend(self) -> str:
        return self.lib.sonar_backend().decode()

    def kernel_launches(self) -> int:
        return int(self.lib.sonar_kernel_launches(self.ctx))

    def stream(self) -> int:
        return int(self.lib.sonar_stream(self.ctx) or 0)

    def profile_enable(self, on: bool = True):
        self._chk(self.lib.sonar_profile_enable(self.ctx, int(on)))

    def profile_read(self) -> dict:
        """{kernel name: (total ms, launches)} since the last read; synchronises."""
        buf = (KernelTime * 64)()
        n = C.c_int()
        self._chk(self.lib.sonar_profile_read(self.ctx, buf, 64, C.byref(n)))
        return {buf[i].kernel.decode(): (buf[i].total_ms, int(buf[i].launches)) for i in range(n.value)}

    def exact_counts(self) -> tuple[int, int]:
        """(STFT frames, YIN frames) of the last fingerprint batch that took the float64 re-evaluation; synchronises."""
        a, b = C.c_int64(), C.c_int64()
        self._chk(self.lib.sonar_fp_exact_counts(self.ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def synchronize(self):
        self._chk(self.lib.sonar_synchronize(self.ctx))

    def host_register(self, arr: np.ndarray):
        """sonar_host_register: page-locks a caller-owned (pageable) numpy array in place."""
        self._chk(self.lib.sonar_host_register(self.ctx, arr.ctypes.data, arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self._chk(self.lib.sonar_host_unregister(self.ctx, arr.ctypes.data))

    def default_params(self, **kw) -> FpParams:
        p = FpParams()
        self.lib.sonar_fp_params_default(C.byref(p))
        for k, v in kw.items():
            if k == "window_type" and isinstance(v, str):
                v = WINDOWS[v]
            setattr(p, k, v)
        return p

    # -- windows ---------------------------------------------------------------
    def window(self, wtype, size, symmetric=True, normalize=True, beta=8.6, alpha=0.5) -> np.ndarray:
        out = np.empty(max(size, 0), dtype=np.float64)
        t = WINDOWS[wtype] if isinstance(wtype, str) else wtype
        self._chk(self.lib.sonar_window_f64(t, size, int(symmetric), int(normalize), beta, alpha, _dp(out)))
        return out

    # -- fingerprint -----------------------------------------------------------
    def fp_sizes(self, p: FpParams, n: int) -> FpSizes:
        s = FpSizes()
        self._chk(self.lib.sonar_fp_sizes(C.byref(p), n, C.byref(s)))
        return s

    def _alloc_fp(self, p: FpParams, n: int):
        s = self.fp_sizes(p, n)
        T, Te, Tp = s.n_frames, s.n_energy_frames, s.n_pitch_frames
        shapes = {
            "mfcc": (T, s.n_mfcc), "spectral_centroid": (T,), "spectral_rolloff": (T,),
            "spectral_bandwidth": (T,), "spectral_flatness": (T,), "spectral_crest": (T,),
            "spectral_slope": (T,), "spectral_flux": (s.n_flux,), "zero_crossing_rate": (T,),
            "short_time_energy": (Te,), "energy_entropy": (Te,), "low_energy_ratio": (Te,),
            "high_energy_ratio": (Te,), "pitch_estimate": (Tp,), "pitch_confidence": (Tp,),
            "voicing_strength": (Tp,), "harmonic_ratio": (Tp,), "inharmonicity_ratio": (Tp,),
            "tonal_centroid": (Tp,),
        }
        if p.enable & FP_ENABLE_TEMPORAL:
            shapes.update({"rms_energy": (Te,), "envelope_shape": (s.n_envelope,),
                           "attack_time": (max(Te, 1),)})
        arrays = {k: np.zeros(shape, dtype=np.float64) for k, shape in shapes.items()}
        out = FpOut()
        for k, a in arrays.items():
            setattr(out, k, _dp(a))
        out.attack_time_cap = arrays["attack_time"].size if "attack_time" in arrays else 0
        sizes = {k: getattr(s, k) for k, _ in s._fields_}
        return sizes, arrays, out

    @staticmethod
    def _finish_fp(sizes, arrays, out) -> Fingerprint:
        scal = {k: getattr(out, k) for k in _FP_SCALARS}
        scal["n_attack_time"] = out.n_attack_time
        if "attack_time" in arrays:
            arrays["attack_time"] = arrays["attack_time"][: out.n_attack_time]
        return Fingerprint(sizes=sizes, arrays=arrays, scalars=scal)

    def fingerprint(self, pcm, p: FpParams) -> Fingerprint:
        pcm = _f64(pcm)
        sizes, arrays, out = self._alloc_fp(p, pcm.size)
        self._chk(self.lib.sonar_fingerprint_f64(self.ctx, _dp(pcm), pcm.size, C.byref(p), C.byref(out)))
        return self._finish_fp(sizes, arrays, out)

    def fingerprint_speech(self, pcm, p: FpParams):
        """sonar_fingerprint_speech_f64: (Fingerprint, speech dict) with the speech-specific group enabled."""
        pcm = _f64(pcm)
        sizes, arrays, out = self._alloc_fp(p, pcm.size)
        tp = max(int(sizes["n_pitch_frames"]), 0)
        cap = max(int(sizes["n_energy_frames"]) // 2 + 1, 1)
        voi, tilt, pauses = np.zeros(tp), np.zeros(tp), np.zeros(cap)
        sp = SpeechOut(_dp(voi), _dp(tilt), _dp(pauses), cap, 0, 0, 0, 0, 0.0)
        self.lib.sonar_fingerprint_speech_f64.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.POINTER(FpParams),
                                                          C.POINTER(FpOut), C.POINTER(SpeechOut)]
        self._chk(self.lib.sonar_fingerprint_speech_f64(self.ctx, _dp(pcm), pcm.size, C.byref(p), C.byref(out), C.byref(sp)))
        nf, npause = int(sp.n_frames), int(sp.n_pause)
        speech = {"is_speech": bool(sp.is_speech), "speech_rate": float(sp.speech_rate),
                  "voicing_probability": voi[:nf].copy(), "spectral_tilt": tilt[:nf].copy(),
                  "pause_duration": pauses[:min(npause, cap)].copy(), "n_pause": npause}
        return self._finish_fp(sizes, arrays, out), speech

    def alloc_batch_outputs(self, lengths, p: FpParams):
        """Caller-owned output buffers for fingerprint_batch (reusable across calls, like the Go shim's slices)."""
        ns = len(lengths)
        outs = (FpOut * ns)()
        keep = []
        for i, n in enumerate(lengths):
            sizes, arrays, o = self._alloc_fp(p, n)
            outs[i] = o
            keep.append((sizes, arrays))
        return outs, keep

    def fingerprint_batch(self, pcms, p: FpParams, buffers=None) -> list[Fingerprint]:
        pcms = [_f64(x) for x in pcms]
        ns = len(pcms)
        ptrs = (c_double_p * ns)(*[_dp(x) for x in pcms])
        lens = (C.c_int64 * ns)(*[x.size for x in pcms])
        outs, keep = buffers if buffers is not None else self.alloc_batch_outputs([x.size for x in pcms], p)
        self._chk(self.lib.sonar_fingerprint_batch_f64(self.ctx, ptrs, lens, ns, C.byref(p), outs))
        return [self._finish_fp(k[0], dict(k[1]), outs[i]) for i, k in enumerate(keep)]

    def fingerprint_batch_pcm(self, pcms, p: FpParams, buffers=None) -> list[Fingerprint]:
        """sonar_fingerprint_batch_pcm: float64, float32 or int16 arrays (one dtype for the whole batch)."""
        pcms = [np.ascontiguousarray(x) for x in pcms]
        fmt = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int16): 2}[pcms[0].dtype]
        assert all(x.dtype == pcms[0].dtype for x in pcms)
        ns = len(pcms)
        ptrs = (C.c_void_p * ns)(*[x.ctypes.data for x in pcms])
        lens = (C.c_int64 * ns)(*[x.size for x in pcms])
        outs, keep = buffers if buffers is not None else self.alloc_batch_outputs([x.size for x in pcms], p)
        self.lib.sonar_fingerprint_batch_pcm.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, c_int64_p, C.c_int,
                                                         C.POINTER(FpParams), C.POINTER(FpOut)]
        self._chk(self.lib.sonar_fingerprint_batch_pcm(self.ctx, ptrs, fmt, lens, ns, C.byref(p), outs))
        return [self._finish_fp(k[0], dict(k[1]), outs[i]) for i, k in enumerate(keep)]

    def fp_dev_layout(self, p: FpParams, n: int) -> FpDevLayout:
        L = FpDevLayout()
        self._chk(self.lib.sonar_fp_dev_layout(C.byref(p), n, C.byref(L)))
        return L

    def fingerprint_batch_dev(self, pcm_dev: int, n: int, stride: int, n_streams: int, p: FpParams,
                              feat_dev: int):
        self._chk(self.lib.sonar_fingerprint_batch_dev(self.ctx, pcm_dev, n, stride, n_streams,
                                                       C.byref(p), feat_dev))

    def stft(self, pcm, win, hop, wtype="hann", phase=False, cplx=False):
        pcm = _f64(pcm)
        T = int((pcm.size - win) / hop) + 1 if (win > 0 and hop > 0) else 0  # Go's int division truncates toward zero
        B = win // 2 + 1
        T = max(T, 0)
        mag = np.zeros((T, B))
        ph = np.zeros((T, B)) if phase else None
        cx = np.zeros((T, B, 2)) if cplx else None
        t = WINDOWS[wtype] if isinstance(wtype, str) else wtype
        self._chk(self.lib.sonar_stft_f64(self.ctx, _dp(pcm), pcm.size, win, hop, t, _dp(mag), _dp(ph), _dp(cx)))
        return mag, ph, cx

    def stft_stream(self, win, hop, wtype="hann"):
        """SpectralAnalyzer.ComputeSTFTStreaming (analyzers/spectral.go:289-310): returns an StftStream."""
        t = WINDOWS[wtype] if isinstance(wtype, str) else wtype
        h = C.c_void_p()
        self._chk(self.lib.sonar_stft_stream_open(self.ctx, win, hop, t, C.byref(h)))
        return StftStream(self, h, win)

    # -- cross-correlation -----------------------------------------------------
    def xcorr(self, a, b, max_lag, want_corr=True):
        a, b = _f64(a), _f64(b)
        corr = np.zeros(2 * max(max_lag, 0) + 1) if want_corr else None
        s = XcorrSummary()
        self._chk(self.lib.sonar_xcorr_ncc_f64(self.ctx, _dp(a), a.size, _dp(b), b.size, max_lag, _dp(corr),
                                               C.byref(s)))
        if corr is not None:
            corr = corr[: 2 * s.actual_max_lag + 1]
        return corr, s

    def xcorr_batch(self, As, Bs, max_lag, want_corr=False):
        As, Bs = [_f64(x) for x in As], [_f64(x) for x in Bs]
        n = len(As)
        pa = (c_double_p * n)(*[_dp(x) for x in As])
        pb = (c_double_p * n)(*[_dp(x) for x in Bs])
        la = (C.c_int64 * n)(*[x.size for x in As])
        lb = (C.c_int64 * n)(*[x.size for x in Bs])
        corrs = [np.zeros(2 * max_lag + 1) for _ in range(n)] if want_corr else None
        pc = (c_double_p * n)(*[_dp(x) for x in corrs]) if want_corr else None
        outs = (XcorrSummary * n)()
        self._chk(self.lib.sonar_xcorr_batch_f64(self.ctx, pa, la, pb, lb, n, max_lag, pc, outs))
        return corrs, list(outs)

    def nccl_unique_id(self) -> bytes:
        """sonar_nccl_unique_id: 128 bytes rank 0 hands to every rank (any transport)."""
        buf = C.create_string_buffer(128)
        self._chk(self.lib.sonar_nccl_unique_id(buf, 128))
        return buf.raw

    def nccl_init(self, world: int, rank: int, uid: bytes):
        self._chk(self.lib.sonar_nccl_init(self.ctx, world, rank, uid))

    def nccl_shutdown(self):
        self._chk(self.lib.sonar_nccl_shutdown(self.ctx))

    def xcorr_lag_sharded(self, a, b, max_lag, want_corr=False, dev_ptrs=None):
        """sonar_xcorr_lag_sharded: this rank's lags + one ncclAllGather + the peak analysis, inside the library.
        dev_ptrs = (a_ptr, na, b_ptr, nb): the sequences are already on the device."""
        s = XcorrSummary()
        corr = None
        if dev_ptrs is not None:
            pa, na, pb, nb = dev_ptrs
            if want_corr:
                corr = np.zeros(2 * max(0, min(max_lag, na - 1, nb - 1)) + 1)
            self._chk(self.lib.sonar_xcorr_lag_sharded(self.ctx, pa, na, pb, nb, max_lag, 1, _dp(corr), C.byref(s)))
        else:
            a, b = _f64(a), _f64(b)
            if want_corr:
                corr = np.zeros(2 * max(0, min(max_lag, a.size - 1, b.size - 1)) + 1)
            self._chk(self.lib.sonar_xcorr_lag_sharded(self.ctx, a.ctypes.data, a.size, b.ctypes.data, b.size, max_lag, 0,
                                                       _dp(corr), C.byref(s)))
        return s, corr

    def xcorr_shard(self, a, b, max_lag, lo, hi):
        a, b = _f64(a), _f64(b)
        sh = C.c_void_p()
        pk = XcorrShardPeak()
        self._chk(self.lib.sonar_xcorr_shard_open(self.ctx, _dp(a), a.size, _dp(b), b.size, max_lag, lo, hi,
                                                  C.byref(sh), C.byref(pk)))
        return sh, pk

    def xcorr_shard_metrics(self, sh, gpeak) -> XcorrShardMetrics:
        m = XcorrShardMetrics()
        self._chk(self.lib.sonar_xcorr_shard_metrics_f64(sh, gpeak, C.byref(m)))
        return m

    def xcorr_shard_corr(self, sh, count) -> np.ndarray:
        out = np.zeros(count)
        self._chk(self.lib.sonar_xcorr_shard_corr(sh, _dp(out)))
        return out

    def xcorr_shard_close(self, sh):
        self.lib.sonar_xcorr_shard_close(sh)

    def xcorr_merge_peaks(self, peaks) -> int:
        arr = (XcorrShardPeak * len(peaks))(*peaks)
        g = C.c_int64()
        self._chk(self.lib.sonar_xcorr_merge_peaks(arr, len(peaks), C.byref(g)))
        return g.value

    def xcorr_merge_metrics(self, parts, na, nb, max_lag, gpeak) -> XcorrSummary:
        arr = (XcorrShardMetrics * len(parts))(*parts)
        s = XcorrSummary()
        self._chk(self.lib.sonar_xcorr_merge_metrics(arr, len(parts), na, nb, max_lag, gpeak, C.byref(s)))
        return s

    def align_xcorr(self, q, r, max_lag_frames, hop, sr, want_corr=False):
        q, r = _f64(q), _f64(r)
        corr = np.zeros(2 * max(max_lag_frames, 0) + 1) if want_corr else None
        xs, ar = XcorrSummary(), AlignResult()
        self._chk(self.lib.sonar_align_xcorr_f64(self.ctx, _dp(q), q.size, _dp(r), r.size, max_lag_frames, hop,
                                                 sr, _dp(corr), C.byref(xs), C.byref(ar)))
        return corr, xs, ar

    # -- DTW -------------------------------------------------------------------
    def _dtw_out(self, n, m, want_matrix):
        cap = n + m + 2
        pq = np.zeros(cap, dtype=np.int32)
        pr = np.zeros(cap, dtype=np.int32)
        pc = np.zeros(cap, dtype=np.float64)
        cm = np.zeros((n, m + 1)) if want_matrix else None
        o = DtwOut()
        o.path_query = pq.ctypes.data_as(c_int32_p)
        o.path_ref = pr.ctypes.data_as(c_int32_p)
        o.path_cost = _dp(pc)
        o.path_cap = cap
        o.cost_matrix = _dp(cm)
        return o, (pq, pr, pc, cm)

    @staticmethod
    def _dtw_result(o, bufs):
        pq, pr, pc, cm = bufs
        L = o.path_len
        return {"path_query": pq[:L], "path_ref": pr[:L], "path_cost": pc[:L], "distance": o.distance,
                "total_cost": o.total_cost, "cost_matrix": cm, "_out": o, "_bufs": bufs}

    def dtw(self, q, r, band=-1, step=STEP_SYMMETRIC2, want_matrix=False):
        q, r = _f64(q), _f64(r)
        if q.ndim == 1:
            q = q[:, None]
        if r.ndim == 1:
            r = r[:, None]
        q, r = _f64(q), _f64(r)
        n, m, dim = q.shape[0], r.shape[0], q.shape[1] if q.ndim == 2 else 1
        o, bufs = self._dtw_out(n, m, want_matrix)
        self._chk(self.lib.sonar_dtw_f64(self.ctx, _dp(q), n, _dp(r), m, dim, band, step, 0, C.byref(o)))
        return self._dtw_result(o, bufs)

    def dtw_batch(self, qs, rs, band=-1, step=STEP_SYMMETRIC2):
        qs = [_f64(x if np.ndim(x) == 2 else np.asarray(x)[:, None]) for x in qs]
        rs = [_f64(x if np.ndim(x) == 2 else np.asarray(x)[:, None]) for x in rs]
        npairs = len(qs)
        n, m, dim = qs[0].shape[0], rs[0].shape[0], qs[0].shape[1]
        pq = (c_double_p * npairs)(*[_dp(x) for x in qs])
        pr = (c_double_p * npairs)(*[_dp(x) for x in rs])
        outs = (DtwOut * npairs)()
        keep = []
        for i in range(npairs):
            o, bufs = self._dtw_out(n, m, False)
            outs[i] = o
            keep.append(bufs)
        self._chk(self.lib.sonar_dtw_batch_f64(self.ctx, pq, pr, npairs, n, m, dim, band, step, 0, outs))
        return [self._dtw_result(outs[i], keep[i]) for i in range(npairs)]

    def align_dtw_scalars(self, dtw_res, n, m, sr) -> AlignResult:
        ar = AlignResult()
        self._chk(self.lib.sonar_align_dtw_scalars(C.byref(dtw_res["_out"]), n, m, sr, C.byref(ar)))
        return ar

    # -- chained pair pipeline ---------------------------------------------------
    def align_pairs_sizes(self, p: FpParams, n: int, max_lag_seconds: float):
        nl, dl = C.c_int32(), C.c_int32()
        self._chk(self.lib.sonar_align_pairs_sizes(C.byref(p), n, max_lag_seconds, C.byref(nl), C.byref(dl)))
        return nl.value, dl.value

    def alloc_pair_outputs(self, n_pairs: int, n: int, p: FpParams, max_lag_seconds: float, features=True, corr=True):
        """Caller-owned result buffers for align_pairs / align_pairs_dev (reusable across calls)."""
        nl, dl = self.align_pairs_sizes(p, n, max_lag_seconds)
        outs = (PairOut * n_pairs)()
        keep = []
        for i in range(n_pairs):
            k = {}
            if features:
                sq, aq, oq = self._alloc_fp(p, n)
                sr, ar, orr = self._alloc_fp(p, n)
                outs[i].query, outs[i].reference = oq, orr
                k["query"], k["reference"] = (sq, aq), (sr, ar)
            if corr:
                k["corr"] = np.zeros(nl)
                outs[i].corr = _dp(k["corr"])
            cap = 2 * dl
            k["pq"], k["pr"], k["pc"] = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
            outs[i].dtw.path_query = k["pq"].ctypes.data_as(c_int32_p)
            outs[i].dtw.path_ref = k["pr"].ctypes.data_as(c_int32_p)
            outs[i].dtw.path_cost = _dp(k["pc"])
            outs[i].dtw.path_cap = cap
            keep.append(k)
        return outs, keep

    def _pair_results(self, outs, keep):
        res = []
        for i, k in enumerate(keep):
            o = outs[i]
            L = o.dtw.path_len
            d = {"xcorr": o.xcorr, "corr_alignment": o.corr_alignment, "corr": k.get("corr"),
                 "dtw_length": o.dtw_length, "path_query": k["pq"][:L], "path_ref": k["pr"][:L],
                 "path_cost": k["pc"][:L], "distance": o.dtw.distance, "total_cost": o.dtw.total_cost}
            if "query" in k:
                d["query"] = self._finish_fp(k["query"][0], dict(k["query"][1]), o.query)
                d["reference"] = self._finish_fp(k["reference"][0], dict(k["reference"][1]), o.reference)
            res.append(d)
        return res

    def align_pairs(self, queries, references, p: FpParams, max_lag_seconds: float, dtw_band: int, buffers=None):
        qs, rs = [_f64(x) for x in queries], [_f64(x) for x in references]
        npairs, n = len(qs), qs[0].size
        outs, keep = buffers if buffers is not None else self.alloc_pair_outputs(npairs, n, p, max_lag_seconds)
        pq = (c_double_p * npairs)(*[_dp(x) for x in qs])
        pr = (c_double_p * npairs)(*[_dp(x) for x in rs])
        self._chk(self.lib.sonar_align_pairs_f64(self.ctx, pq, pr, n, npairs, C.byref(p), max_lag_seconds, dtw_band, outs))
        return self._pair_results(outs, keep)

    def music_spectral(self, pcm, win=1024, hop=256, window_type="hann", sample_rate=44100, n_bands=6, n_bark=24,
                       bark_low=0.0, bark_high=None):
        """sonar_music_spectral_f64: (contrast [T][n_bands], chroma [T][12], bark [T][n_bark])."""
        x = _f64(pcm)
        T = int((x.size - win) / hop) + 1  # Go's int division truncates toward zero
        if T <= 0:
            T = 0
        contrast, chroma, bark = np.zeros((T, n_bands)), np.zeros((T, 12)), np.zeros((T, n_bark))
        wt = WINDOWS[window_type] if isinstance(window_type, str) else window_type
        hi = float(sample_rate) / 2 if bark_high is None else bark_high
        self._chk(self.lib.sonar_music_spectral_f64(self.ctx, _dp(x), x.size, win, hop, wt, sample_rate, n_bands,
                                                    _dp(contrast), _dp(chroma), n_bark, bark_low, hi, _dp(bark)))
        return contrast, chroma, bark

    PCM_FORMATS = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int16): 2}

    def align_pairs_pcm(self, queries, references, p: FpParams, max_lag_seconds: float, dtw_band: int, buffers=None):
        """sonar_align_pairs_pcm: PCM as float64, float32 or int16 arrays (the decoder's own sample format)."""
        fmt = self.PCM_FORMATS[np.asarray(queries[0]).dtype]
        qs = [np.ascontiguousarray(x) for x in queries]
        rs = [np.ascontiguousarray(x) for x in references]
        assert all(x.dtype == qs[0].dtype for x in qs + rs)
        npairs, n = len(qs), qs[0].size
        outs, keep = buffers if buffers is not None else self.alloc_pair_outputs(npairs, n, p, max_lag_seconds)
        pq = (C.c_void_p * npairs)(*[x.ctypes.data for x in qs])
        pr = (C.c_void_p * npairs)(*[x.ctypes.data for x in rs])
        self._chk(self.lib.sonar_align_pairs_pcm(self.ctx, pq, pr, fmt, n, npairs, C.byref(p), max_lag_seconds, dtw_band,
                                                 outs))
        return self._pair_results(outs, keep)

    def align_pairs_dev(self, pcm_dev: int, n: int, stride: int, n_pairs: int, p: FpParams, max_lag_seconds: float,
                        dtw_band: int, buffers=None):
        outs, keep = buffers if buffers is not None else self.alloc_pair_outputs(n_pairs, n, p, max_lag_seconds,
                                                                                 features=False)
        self._chk(self.lib.sonar_align_pairs_dev(self.ctx, pcm_dev, n, stride, n_pairs, C.byref(p), max_lag_seconds,
                                                 dtw_band, outs))
        return self._pair_results(outs, keep)

    # -- comparison ------------------------------------------------------------
    def colstats(self, x, dim=None) -> np.ndarray:
        x = _f64(x)
        if x.ndim == 1:
            x = x[:, None]
        t, d = x.shape
        st = np.zeros(2 * d)
        self._chk(self.lib.sonar_colstats_f64(self.ctx, _dp(x), t, d, _dp(st)))
        return st

    def colstats_cosine(self, x, y) -> float:
        x, y = _f64(x), _f64(y)
        if x.ndim == 1:
            x = x[:, None]
        if y.ndim == 1:
            y = y[:, None]
        sim = C.c_double()
        self._chk(self.lib.sonar_colstats_cosine_f64(self.ctx, _dp(x), x.shape[0], _dp(y), y.shape[0], x.shape[1],
                                                     C.byref(sim)))
        return sim.value

    @staticmethod
    def cmp_features(fp: Fingerprint, content_type: int = 0, spectral=True, harmonic=True,
                     temporal=False):
        f = CmpFeatures()
        keep = []

        def put(name, cnt_name, arr):
            arr = _f64(arr)
            keep.append(arr)
            setattr(f, name, _dp(arr))
            setattr(f, cnt_name, arr.shape[0])

        a = fp.arrays
        put("mfcc", "mfcc_frames", a["mfcc"])
        f.mfcc_dim = a["mfcc"].shape[1]
        f.content_type = content_type
        f.has_spectral, f.has_harmonic, f.has_temporal = int(spectral), int(harmonic), int(temporal)
        put("spectral_centroid", "n_centroid", a["spectral_centroid"])
        put("spectral_rolloff", "n_rolloff", a["spectral_rolloff"])
        put("spectral_flux", "n_flux", a["spectral_flux"])
        put("harmonic_ratio", "n_harmonic_ratio", a["harmonic_ratio"])
        put("pitch_estimate", "n_pitch", a["pitch_estimate"])
        if temporal and "rms_energy" in a:
            put("rms_energy", "n_rms", a["rms_energy"])
            f.dynamic_range = fp.scalars["dynamic_range"]
            f.silence_ratio = fp.scalars["silence_ratio"]
            f.onset_density = fp.scalars["onset_density"]
        return f, keep

    def truncate_to_alignment(self, n1: int, n2: int, sample_rate: int, offset_seconds: float):
        """sonar_truncate_to_alignment: (start1, start2, length) of TruncateToAlignmentPCM's segments."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.lib.sonar_truncate_to_alignment.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_double, c_int64_p, c_int64_p,
                                                         c_int64_p]
        self._chk(self.lib.sonar_truncate_to_alignment(n1, n2, sample_rate, offset_seconds, C.byref(a), C.byref(b),
                                                       C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def compare_batch(self, query: CmpFeatures, candidates, weights, content_filter=False):
        """sonar_compare_batch_f64: FingerprintComparator.BatchCompare (None candidates are skipped)."""
        w = CmpWeights()
        for i, v in enumerate(weights):
            w.w[i] = v
        n = len(candidates)
        ptrs = (C.POINTER(CmpFeatures) * n)(*[C.pointer(c) if c is not None else None for c in candidates])
        res = (CmpResult * n)()
        self.lib.sonar_compare_batch_f64.argtypes = [C.c_void_p, C.POINTER(CmpFeatures),
                                                     C.POINTER(C.POINTER(CmpFeatures)), C.c_int, C.POINTER(CmpWeights),
                                                     C.c_int, C.POINTER(CmpResult)]
        self._chk(self.lib.sonar_compare_batch_f64(self.ctx, C.byref(query), ptrs, n, C.byref(w), int(content_filter),
                                                   res))
        return list(res)

    def compare(self, f1: CmpFeatures, f2: CmpFeatures, weights, content_filter=False) -> CmpResult:
        w = CmpWeights()
        for i, v in enumerate(weights):
            w.w[i] = v
        r = CmpResult()
        self._chk(self.lib.sonar_compare_f64(self.ctx, C.byref(f1), C.byref(f2), C.byref(w),
                                             int(content_filter), C.byref(r)))
        return r
