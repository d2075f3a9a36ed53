/*
 * qampy_b200 -- C ABI of the B200-native coherent-receiver DSP hot path.
 *
 * This is the drop-in boundary for the ONE path this project re-implements: the adaptive
 * MIMO FIR equaliser (train + apply) and the blind-phase-search carrier recovery of
 * ChalmersPhotonicsLab/QAMpy.  Each entry point replaces one Pythran-exported kernel of the
 * reference (paths relative to the reference checkout):
 *
 *   qb_train_equaliser_*        qampy/core/equalisation/pythran_equalisation.py:128-173  train_equaliser
 *   qb_apply_filter_to_signal_* qampy/core/equalisation/pythran_equalisation.py:33-76    apply_filter_to_signal
 *   qb_bps_*                    qampy/core/pythran_dsp.py:45-85 bps (+ :26-42 select_angle_index) and,
 *                               when ph/Eout are requested, the L2 tail
 *                               qampy/core/phaserecovery.py:150-159 (select_angles, unwrap*4/4, rotate)
 *   qb_select_angles_*          qampy/core/pythran_dsp.py:133-153 select_angles
 *
 * Conventions
 *   - plain C: pointers, sizes, no C++/torch types.  All functions return 0 on success or a
 *     negative qb_status; qb_last_error() gives the message of the last failure on the calling
 *     thread.  Nothing here falls back to a CPU implementation: without a usable CUDA device
 *     every compute entry point fails with QB_ERR_CUDA.
 *   - complex arrays are interleaved (re, im) pairs of float (QB_C64) or double (QB_C128),
 *     row-major, exactly the NumPy layout the reference hands to its kernels.
 *   - `*_host` entry points take HOST pointers, are synchronous, and do their own H2D/D2H
 *     copies: they are what a ctypes/cffi binding of the reference's L1 seam calls.
 *   - `*_dev` entry points take DEVICE pointers (every array argument except `modes`), enqueue
 *     work on `stream` (a cudaStream_t passed as void*, NULL = legacy default stream) and return
 *     without synchronising: they are CUDA-graph capturable and keep the signal resident in HBM
 *     across train -> train -> apply -> bps.
 *   - batched ("segment") layout: `nseg` independent time segments are processed by one launch.
 *     Segment s of the signal starts at E + s*seg_stride samples and its mode rows are
 *     row_stride samples apart, so overlapping segments of one long capture need no copy.
 *     nseg = 1, seg_stride = nmodes*L, row_stride = L is the reference call.
 */
#ifndef QAMPY_B200_H
#define QAMPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_VERSION 100 /* 0.1.0 */

typedef enum {
    QB_OK = 0,
    QB_ERR_ARG = -1,         /* invalid argument (the reference raises ValueError/AssertionError) */
    QB_ERR_UNSUPPORTED = -2, /* valid in the reference but not implemented here */
    QB_ERR_CUDA = -3,        /* CUDA runtime/driver error, or no device */
    QB_ERR_NOMEM = -4
} qb_status;

typedef enum { QB_C64 = 0, QB_C128 = 1 } qb_dtype;

/* error functions, pythran_equalisation.py:131-152 (string dispatch) */
typedef enum {
    QB_CMA = 0,      /* :178 cma_error        */
    QB_CMA2 = 1,     /* :182 cma2_error       */
    QB_SGNCMA = 2,   /* :133-134 mapped to cma_error by the reference */
    QB_MCMA = 3,     /* :190 mcma_error       */
    QB_RDE = 4,      /* :196 rde_error        */
    QB_MRDE = 5,     /* :203 mrde_error       */
    QB_SBD = 6,      /* :213 sbd_error        */
    QB_SBD_DATA = 7, /* :218 sbd_data_error   */
    QB_MDDMA = 8,    /* :224 mddma_error      */
    QB_DD = 9,       /* :229 ddlms_error      */
    /* error functions of train_equaliser_realvalued (:80-125).  The real-valued trainer runs on the complex
     * kernels with the imaginary parts held at exactly zero (E, wx and symbols widened by the caller);
     * these ids select its error functions and the real form of the step-size rule (:18-22). */
    QB_CMA_REAL = 10,     /* :113 cma_error_real     */
    QB_SGNCMA_REAL = 11,  /* :117 sgncma_error_real  */
    QB_DD_REAL = 12,      /* :121 dd_error_real (det_symbol_argmin :232-235) */
    QB_DD_DATA_REAL = 13  /* :125 dd_data_error_real */
} qb_method;

#define QB_MAX_MODES 8     /* nmodes (input rows / output modes) per segment            */
#define QB_MAX_TAPDIM 512  /* nmodes * ntaps handled by the trainer                     */
#define QB_MAX_ANGLES 256  /* BPS test angles                                           */

int qb_version(void);
const char *qb_last_error(void);
/* number of CUDA devices visible, or a negative qb_status */
int qb_device_count(void);
/* method name ("mcma", "mrde", ...) -> qb_method, or QB_ERR_ARG ("Unknown method") */
int qb_method_from_name(const char *name);

/* ---- equaliser training ---------------------------------------------------------------
 * replaces train_equaliser(E, TrSyms, Niter, os, mu, wx, modes, adaptive, symbols, method)
 *   E        (nseg, nmodes, >= (TrSyms-1)*os + ntaps) complex
 *   wx       (nseg, nmodes, nmodes, ntaps) complex, IN/OUT (the reference updates wx in place)
 *   modes    HOST array of nsel output-mode numbers (< nmodes)
 *   symbols  (nmodes, K) complex per-method constants / alphabet / training sequence
 *   mu       (nseg, nsel) real, IN/OUT: step size per trained stream (final value written back)
 *   err      (nseg, nmodes, TrSyms*Niter) complex or NULL; only rows of selected modes are written
 * Each (segment, mode) stream is an independent serial recurrence.                         */
int qb_train_equaliser_dev(int dtype, const void *E, int64_t nseg, int64_t seg_stride,
                           int64_t row_stride, int64_t nmodes, int64_t TrSyms, int64_t Niter,
                           int64_t os, void *wx, int64_t ntaps, const int64_t *modes, int64_t nsel,
                           int adaptive, const void *symbols, int64_t K, int method, void *mu,
                           void *err, void *stream);

/* HOST-pointer form with the reference's exact call shape (nseg = 1).  `mu` points to ONE real
 * (in: step size, out: final step size).  mu_shared != 0 reproduces the interpreted reference,
 * where one mu is carried through the modes in list order when adaptive (:130, :162-172);
 * mu_shared == 0 starts every mode from *mu (deterministic; what an OpenMP build intends).   */
int qb_train_equaliser_host(int dtype, const void *E, int64_t nmodes, int64_t L, int64_t TrSyms,
                            int64_t Niter, int64_t os, void *mu, void *wx, int64_t ntaps,
                            const int64_t *modes, int64_t nsel, int adaptive, const void *symbols,
                            int64_t K, int method, int mu_shared, void *err);

/* ---- static filter + decimation ---------------------------------------------------------
 * replaces apply_filter_to_signal(E, os, wx, modes):  out (nseg, nsel, N), N = (L-ntaps+1)/os */
int qb_apply_filter_to_signal_dev(int dtype, const void *E, int64_t nseg, int64_t seg_stride,
                                  int64_t row_stride, int64_t nmodes, int64_t L, int64_t os,
                                  const void *wx, int64_t ntaps, const int64_t *modes, int64_t nsel,
                                  void *out, void *stream);
int qb_apply_filter_to_signal_host(int dtype, const void *E, int64_t nmodes, int64_t L, int64_t os,
                                   const void *wx, int64_t ntaps, const int64_t *modes,
                                   int64_t nsel, void *out);

/* ---- blind phase search -------------------------------------------------------------------
 * replaces bps(E, testangles, symbols, N) for `nstream` independent 1-D streams (the reference
 * is called once per polarisation) and optionally the L2 tail.
 *   E        (nstream, L) complex, stream s at E + s*stream_stride
 *   comp     (A) complex  = exp(1j*testangles)   (pythran_dsp.py:72; computed by the caller so the
 *                                                 table is bit-identical to NumPy's)
 *   angles   (A) real test angles (only read when ph/Eout are requested)
 *   symbols  (M) complex alphabet
 *   lev_re/lev_im  sorted, uniformly spaced per-axis levels when the alphabet is a full
 *            rectangular grid (see qb_detect_grid_host), else n_re = n_im = 0 -> brute force.
 *            The slicer result is bit-identical to the brute-force minimum distance.
 *   idx      (nstream, L) int32 or NULL   -- select_angle_index output (edges 0)
 *   ph       (nstream, L) real  or NULL   -- angles[idx], [N:L-N] unwrapped as np.unwrap(ph*4)/4
 *   Eout     (nstream, L) complex or NULL -- E * exp(+1j*ph)
 * qb_bps_* take ONE angle table (testangles.shape[0] == 1, pythran_dsp.py:76-77 ph_idx = 0).
 * qb_bps_rows_* take a PER-SYMBOL table, comp and angles of shape (nstream, L, A) -- the p == L form
 * of pythran_dsp.py:74-75 that the second stage of two-stage BPS uses
 * (qampy/core/phaserecovery.py:276-281).  Only the index search is fused there: ph and Eout must be
 * NULL (the two-stage tail unwraps the whole array, :282, unlike bps); use qb_select_angles_*.   */
int qb_bps_dev(int dtype, const void *E, int64_t nstream, int64_t stream_stride, int64_t L,
               const void *comp, const void *angles, int64_t A, const void *symbols, int64_t M,
               const void *lev_re, int64_t n_re, const void *lev_im, int64_t n_im, int64_t N,
               int32_t *idx, void *ph, void *Eout, void *stream);
int qb_bps_host(int dtype, const void *E, int64_t nstream, int64_t L, const void *comp,
                const void *angles, int64_t A, const void *symbols, int64_t M, int64_t N,
                int32_t *idx, void *ph, void *Eout);
int qb_bps_rows_dev(int dtype, const void *E, int64_t nstream, int64_t stream_stride, int64_t L,
                    const void *comp, const void *angles, int64_t A, const void *symbols, int64_t M,
                    const void *lev_re, int64_t n_re, const void *lev_im, int64_t n_im, int64_t N,
                    int32_t *idx, void *ph, void *Eout, void *stream);
int qb_bps_rows_host(int dtype, const void *E, int64_t nstream, int64_t L, const void *comp,
                     const void *angles, int64_t A, const void *symbols, int64_t M, int64_t N,
                     int32_t *idx, void *ph, void *Eout);

/* HOST helper: if `symbols` (M complex) is exactly the product set of n_re real levels and n_im
 * imaginary levels (each uniformly spaced), write the sorted levels (capacity 64 each, real type
 * of dtype) and return 1; return 0 if not a grid; negative on error.                          */
int qb_detect_grid_host(int dtype, const void *symbols, int64_t M, void *lev_re, int64_t *n_re,
                        void *lev_im, int64_t *n_im);

/* ---- decisions and quality metrics -----------------------------------------------------
 * qb_make_decision_*: make_decision, qampy/core/equalisation/pythran_equalisation.py:306-334 -- for every
 *     E[i]: idx = first minimum of |E[i] - symbols[j]| (np.argmin of np.abs, :232-235), det = symbols[idx],
 *     dist = that modulus.  det / dist / idx may be NULL.
 * qb_soft_l_value_demapper_host: soft_l_value_demapper (minmax = 0, pythran_dsp.py:95-108) and
 *     soft_l_value_demapper_minmax (minmax = 1, :110-131); bits_map (nbits_map, K, 2) complex, L_values
 *     (N, num_bits) float64 like the reference's output array.
 * qb_estimate_snr_host: estimate_snr, pythran_dsp.py:244-286; out3 = (snr, signal power, noise power). */
int qb_make_decision_dev(int dtype, const void *E, int64_t L, const void *symbols, int64_t M, void *det, void *dist,
                         int32_t *idx, void *stream);
int qb_make_decision_host(int dtype, const void *E, int64_t L, const void *symbols, int64_t M, void *det, void *dist,
                          int32_t *idx);
int qb_soft_l_value_demapper_host(int dtype, const void *rx, int64_t N, int64_t num_bits, double snr,
                                  const void *bits_map, int64_t nbits_map, int64_t K, int minmax, double *L_values);
int qb_estimate_snr_host(int dtype, const void *signal_rx, const void *symbols_tx, int64_t n,
                         const void *gray_symbols, int64_t ngray, double *out3);

/* ---- element-wise stages of the pilot-based receiver (device pointers only) ----------------------
 * qb_freq_shift_dev: comp_freq_offset, qampy/core/phaserecovery.py:438-473:
 *     out[r, t] = E[r, t] * exp(-2j*pi*(t0 + t + 1)*freq[r]/os),  freq (nrows) float64 on the device;
 *     t0 = 0 for a whole signal, the window's first sample index when only a window is compensated.
 * qb_pilot_cpe_dev: pilot_based_cpe_new, qampy/core/pilotbased_receiver.py:258-327, one frame per row:
 *     pilot_idx (npilots) int64 sorted positions inside the row, pilots (nrows, npilots) reference pilots at
 *     pilots + r*pilot_stride; residual pilot phase -> unwrap -> moving average over num_average (odd) ->
 *     linear interpolation to every symbol -> out = E*exp(-1j*phase); trace (real, optional) = phase.     */
int qb_freq_shift_dev(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, const double *freq,
                      int64_t os, int64_t t0, void *out, int64_t out_stride, void *stream);
int qb_pilot_cpe_dev(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t nlen,
                     const int64_t *pilot_idx, const void *pilots, int64_t pilot_stride, int64_t npilots,
                     int64_t num_average, void *out, int64_t out_stride, void *trace, int64_t trace_stride,
                     void *stream);

/* ---- signal synthesis on the device (device pointers only) ---------------------------------
 * The spectral / element-wise stages of the reference's generators around cuFFT (the caller owns the FFTs):
 *   qampy/core/resample.py:73-126 rrcos_resample (+ core/filter.py:177-212 rrcos_pulseshaping, fftconvolve "same"),
 *   qampy/core/impairments.py:94-131 apply_PMD_to_field, :133-186 phase_noise / apply_phase_noise, :188-233 add_awgn /
 *   change_snr.  All arrays complex128 (the reference generates in double) except the final output.
 * qb_synth_upsample_dev   out[r, up*k] = symbols[r, k], zero elsewhere up to out_stride (the FFT length).
 * qb_synth_specmul_dev    X[r, k] *= H[k]: spectrum of a row times the spectrum of the zero-padded tap vector.
 * qb_synth_crop_norm_dev  out[r, t] = x[r, first + t*down], t < n ("same" crop of the convolution + decimation); with
 *                         renormalise: centred and scaled to power target_power[r] (normalise_and_center * sqrt(p)).
 * qb_synth_pmd_dev        spectra (2, n) in FFT order of the two polarisations: rotate(theta), axis delays
 *                         exp(-+ i omega t_dgd / 2) with omega on the reference's grid, rotate(-theta); in place.
 * qb_synth_tail_dev       out[r, t] = (x[r, t] + noise_sigma[r] (N + iN)/sqrt(2)) * exp(i phase[r, t]), phase = phase0[r]
 *                         + running sum of N(0, walk_sigma^2) steps; deviates from a counter-based generator keyed by
 *                         (seed, row0 + r, index0 + t) so that blocks of one capture can be made independently
 *                         (index0 even); noise_sigma NULL / walk_sigma 0 switch a stage off; phase_out optional.   */
int qb_synth_upsample_dev(const void *symbols, int64_t nrows, int64_t n, int64_t up, void *out, int64_t out_stride,
                          void *stream);
int qb_synth_specmul_dev(void *X, int64_t nrows, int64_t nfft, const void *H, void *stream);
int qb_synth_crop_norm_dev(const void *x, int64_t nrows, int64_t row_stride, int64_t first, int64_t down, int64_t n,
                           const double *target_power, int renormalise, void *out, void *stream);
int qb_synth_pmd_dev(void *spectra, int64_t n, double theta, double t_dgd, double fs, void *stream);
int qb_synth_tail_dev(int dtype, const void *x, int64_t nrows, int64_t n, const double *noise_sigma, double walk_sigma,
                      uint64_t seed, int64_t row0, int64_t index0, const double *phase0, void *out, int64_t out_stride,
                      double *phase_out, void *stream);

/* ---- Viterbi-Viterbi M-th power phase recovery, qampy/core/phaserecovery.py:40-79 ---------------------
 * For every row r of E (nrows, L): ph[r, w] = (unwrap(angle(sum_{t=w..w+N-1} (E[r,t]/|E[r,t]|)^M)) - pi)/M for the
 * L-N+1 windows, out[r, o+w] = E[r, o+w]*exp(-1j*ph[r, w]) with o = (N-1)/2 and zeros where no full window
 * exists.  ph rows are ph_stride (dev) / L-N+1 (host) real elements apart.  1 <= N <= 512.             */
int qb_viterbiviterbi_dev(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, int64_t N, int64_t M,
                          void *out, int64_t out_stride, void *ph, int64_t ph_stride, void *stream);
int qb_viterbiviterbi_host(int dtype, const void *E, int64_t nrows, int64_t L, int64_t N, int64_t M, void *out,
                           void *ph);

/* ---- select_angles(angles, idx): out[i] = angles[p > 1 ? i : 0][idx[i]] ----------------------- */
int qb_select_angles_dev(int dtype, const void *angles, int64_t p, int64_t A, const int64_t *idx,
                         int64_t L, void *out, void *stream);
int qb_select_angles_host(int dtype, const void *angles, int64_t p, int64_t A, const int64_t *idx,
                          int64_t L, void *out);

/* ---- layout of the training kernels --------------------------------------------------------------------
 * Sets the layout used by the CALLING THREAD's subsequent qb_train_equaliser_dev calls and returns the
 * previous one.  QB_LAYOUT_THROUGHPUT (default): four streams per warp, the layout that fills the machine
 * when a launch holds hundreds of (segment, mode) streams.  QB_LAYOUT_LATENCY: one stream per warp (taps
 * spread over all 32 lanes, warp-wide fixed-point reduction of the tap dot) -- for calls whose time is the
 * serial depth of a few streams, i.e. the reference's own call shape: train_equaliser on ONE capture
 * (pythran_equalisation.py:130-173 has one stream per trained mode).  qb_train_equaliser_host always runs
 * in the latency layout.  The layout is never chosen from the size of a launch, so a stream's result does
 * not depend on what else shares its launch; both layouts meet the same parity tolerance.                */
enum { QB_LAYOUT_THROUGHPUT = 0, QB_LAYOUT_LATENCY = 1 };
int qb_set_train_layout(int layout);

/* ---- accumulation mode of the blind phase search ------------------------------------------------------
 * Sets the mode used by the CALLING THREAD's subsequent qb_bps_* calls and returns the previous one.
 * QB_BPS_EXACT (default): the reference's sequential running sum per test angle in the signal's precision and
 * its window difference (pythran_dsp.py:26-42) -- phase indices bit-identical to the reference, including where its
 * fp32 sum has grown to 2e5 after 1e7 symbols and no longer resolves neighbouring angles.
 * QB_BPS_WINDOWED: every window sum is formed directly from its 2N distances (accumulated in double): the
 * numerically sound variant, a flagged deviation from the reference's bits; on complex64 input
 * it follows the reference's complex128 result instead of its complex64 one.                                  */
enum { QB_BPS_EXACT = 0, QB_BPS_WINDOWED = 1 };
int qb_set_bps_accumulation(int mode);

/* ---- kernel-selection overrides (tests and tuning only; results are the same to the parity bounds) ------
 * Process-wide.  name: "TRAIN_KERNEL" (direct | warp | gla), "TRAIN_LPS" (8 | 16), "TRAIN_GLA" (0), "LA_TILE" (symbols
 * per staged tile of the look-ahead trainer), "BPS_KERNEL" (ws | simple), "BPS_SPLIT" (0 | 1 | 2: fused, producer / chain, phase-parallel mapping).  value NULL or "" clears
 * the override.  Initial values come from the environment variables QB_<name>, read once at the first use, never per
 * launch.  Returns QB_OK, or QB_ERR_ARG for an unknown name.  Nothing in the reference corresponds to this.        */
int qb_set_option(const char *name, const char *value);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t qb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QAMPY_B200_H */
