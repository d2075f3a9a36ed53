// C ABI of qampy_b200 (see include/qampy_b200.h): argument validation, error strings, the
// device-pointer entry points (thin launch wrappers) and the host-pointer entry points (own H2D /
// D2H copies around the same launches).  No CPU fallback exists anywhere in this file: without a
// CUDA device every compute entry point returns QB_ERR_CUDA.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "qb_common.cuh"

namespace qb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches += n; }

// kernel-selection overrides: a small table behind a mutex, seeded once from the environment
static const char *const OPTION_NAMES[OPT_COUNT] = {"TRAIN_KERNEL", "TRAIN_LPS", "TRAIN_GLA", "LA_TILE", "BPS_KERNEL", "BPS_SPLIT"};
static std::mutex g_opt_mutex;
static bool g_opt_seeded = false;
static bool g_opt_set[OPT_COUNT];
static char g_opt_val[OPT_COUNT][24];
static void seed_options_locked()
{
    if (g_opt_seeded) return;
    g_opt_seeded = true;
    for (int o = 0; o < OPT_COUNT; o++) {
        char name[48];
        snprintf(name, sizeof(name), "QB_%s", OPTION_NAMES[o]);
        const char *e = getenv(name);
        g_opt_set[o] = e && e[0];
        if (g_opt_set[o]) snprintf(g_opt_val[o], sizeof(g_opt_val[o]), "%s", e);
    }
}
char option_char(Option o)
{
    std::lock_guard<std::mutex> lk(g_opt_mutex);
    seed_options_locked();
    return g_opt_set[o] ? g_opt_val[o][0] : 0;
}
int option_int(Option o, int unset)
{
    std::lock_guard<std::mutex> lk(g_opt_mutex);
    seed_options_locked();
    return g_opt_set[o] ? atoi(g_opt_val[o]) : unset;
}
static int set_option(const char *name, const char *value)
{
    if (!name) return set_error(QB_ERR_ARG, "qb_set_option: name is NULL");
    std::lock_guard<std::mutex> lk(g_opt_mutex);
    seed_options_locked();
    for (int o = 0; o < OPT_COUNT; o++) {
        if (strcmp(name, OPTION_NAMES[o]) != 0) continue;
        g_opt_set[o] = value && value[0];
        if (g_opt_set[o]) snprintf(g_opt_val[o], sizeof(g_opt_val[o]), "%s", value);
        return QB_OK;
    }
    return set_error(QB_ERR_ARG, "qb_set_option: unknown option '%s'", name);
}

int set_train_layout(int layout);
int set_bps_accumulation(int mode);
int apply_dispatch(int dtype, const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                   int64_t nmodes, int64_t L, int64_t os, const void *wx, int64_t ntaps,
                   const int64_t *modes, int64_t nsel, void *out, cudaStream_t st);
int train_dispatch(int dtype, const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                   int64_t nmodes, int64_t TrSyms, int64_t Niter, int64_t os, void *wx, int64_t ntaps,
                   const int64_t *modes, int64_t nsel, int adaptive, const void *symbols, int64_t K,
                   int method, void *mu, void *err, cudaStream_t st);
int bps_dispatch(int dtype, const void *E, int64_t nstream, int64_t stream_stride, int64_t L,
                 const void *comp, const void *angles, int64_t A, const void *symbols, int64_t M,
                 const void *lev_re, int64_t n_re, const void *lev_im, int64_t n_im, int64_t N,
                 int32_t *idx, void *ph, void *Eout, int comp_rows, cudaStream_t st);
int select_angles_dispatch(int dtype, const void *angles, int64_t p, int64_t A, const int64_t *idx,
                           int64_t L, void *out, cudaStream_t st);

int freq_shift_dispatch(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, const double *freq,
                        int64_t os, int64_t t0, void *out, int64_t out_stride, cudaStream_t st);
int pilot_cpe_dispatch(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t nlen, const int64_t *pidx,
                       const void *pilots, int64_t pilot_stride, int64_t nph, int64_t navg, void *out,
                       int64_t out_stride, void *trace, int64_t trace_stride, cudaStream_t st);

int synth_upsample_dispatch(const void *sym, int64_t rows, int64_t n, int64_t up, void *out, int64_t out_stride, cudaStream_t st);
int synth_specmul_dispatch(void *X, int64_t rows, int64_t nfft, const void *H, cudaStream_t st);
int synth_crop_norm_dispatch(const void *x, int64_t rows, int64_t row_stride, int64_t first, int64_t down, int64_t n,
                             const double *target_power, int renorm, double *mom, void *out, cudaStream_t st);
int synth_pmd_dispatch(void *S, int64_t n, double theta, double t_dgd, double fs, cudaStream_t st);
int synth_tail_dispatch(int dtype, const void *x, int64_t rows, int64_t n, const double *noise_sigma, double walk_sigma,
                        uint64_t seed, int64_t row0, uint64_t index0, const double *phase0, double *tile_buf, void *out,
                        int64_t out_stride, double *phase_out, cudaStream_t st);
int vv_dispatch(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, int64_t N, int64_t M, void *out,
                int64_t out_stride, void *ph, int64_t ph_stride, void *work, cudaStream_t st);
size_t vv_work_bytes(int dtype, int64_t nrows, int64_t L, int64_t N);

int make_decision_dispatch(int dtype, const void *E, int64_t L, const void *symbols, int64_t M, void *det, void *dist,
                           int32_t *idx, cudaStream_t st);
int demapper_dispatch(int dtype, const void *rx, int64_t N, int64_t num_bits, double snr, const void *bits_map, int64_t K,
                      int minmax, double *Lv, cudaStream_t st);
int snr_pass_dispatch(int dtype, const void *rx, const void *tx, int64_t n, const void *gray, int64_t Ncls,
                      const double *means, double *acc, int pass, cudaStream_t st);

static inline size_t csize(int dtype) { return dtype == QB_C64 ? 8 : 16; }
static inline size_t rsize(int dtype) { return dtype == QB_C64 ? 4 : 8; }

static int check_dtype(int dtype)
{
    if (dtype != QB_C64 && dtype != QB_C128)
        return set_error(QB_ERR_ARG, "dtype must be QB_C64 or QB_C128, got %d", dtype);
    return QB_OK;
}

static int check_modes(const int64_t *modes, int64_t nsel, int64_t nmodes)
{
    QB_REQUIRE(nmodes >= 1 && nmodes <= QB_MAX_MODES, "nmodes must be in [1, %d], got %lld", QB_MAX_MODES,
               (long long)nmodes);
    QB_REQUIRE(nsel >= 0 && nsel <= QB_MAX_MODES, "number of selected modes must be in [0, %d]", QB_MAX_MODES);
    QB_REQUIRE(nsel == 0 || modes != nullptr, "modes must not be NULL");
    for (int64_t j = 0; j < nsel; j++)
        QB_REQUIRE(modes[j] >= 0 && modes[j] < nmodes,
                   "Maximum mode number must not be higher than number of modes (mode %lld, nmodes %lld)",
                   (long long)modes[j], (long long)nmodes);
    return QB_OK;
}

// scratch device buffer that frees itself; stream-ordered pool allocations (cached by the driver)
struct DevBuf {
    void *p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    int alloc(size_t bytes)
    {
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(&p, bytes, st);
        if (e != cudaSuccess) {
            p = nullptr;
            return set_error(e == cudaErrorMemoryAllocation ? QB_ERR_NOMEM : QB_ERR_CUDA,
                             "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
        }
        return QB_OK;
    }
    ~DevBuf()
    {
        if (p) cudaFreeAsync(p, st);
    }
};

static int host_stream(cudaStream_t *st)
{
    // one library-owned stream (and one memory-pool set-up) per DEVICE, created under a lock: streams and pools belong
    // to the device that is current when they are made, and two threads may make their first call at the same time
    static std::mutex mtx;
    static cudaStream_t streams[64] = {nullptr};
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_error(QB_ERR_CUDA, "no CUDA device available (%s); qampy_b200 has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = 0;
    QB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(QB_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mtx);
    if (!streams[dev]) {
        cudaStream_t s = nullptr;
        QB_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        QB_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t thr = ~0ull;  // keep freed scratch cached between calls
        QB_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        streams[dev] = s;
    }
    *st = streams[dev];
    return QB_OK;
}

#define QB_TRY(expr)             \
    do {                         \
        int _rc = (expr);        \
        if (_rc != QB_OK) return _rc; \
    } while (0)

template <typename T>
static int detect_grid(const T *sy, int64_t M, T *lre, int64_t *n_re, T *lim, int64_t *n_im)
{
    std::vector<T> re, im;
    for (int64_t j = 0; j < M; j++) {
        bool fr = false, fi = false;
        for (T v : re) fr = fr || v == sy[2 * j];
        for (T v : im) fi = fi || v == sy[2 * j + 1];
        if (!fr) re.push_back(sy[2 * j]);
        if (!fi) im.push_back(sy[2 * j + 1]);
        if (re.size() > 64 || im.size() > 64) return 0;
    }
    if ((int64_t)(re.size() * im.size()) != M) return 0;
    auto sortv = [](std::vector<T> &v) {
        for (size_t a = 1; a < v.size(); a++)
            for (size_t b = a; b > 0 && v[b] < v[b - 1]; b--) {
                T t = v[b];
                v[b] = v[b - 1];
                v[b - 1] = t;
            }
    };
    sortv(re);
    sortv(im);
    // every (re, im) combination must be present exactly once
    for (T a : re)
        for (T b : im) {
            int cnt = 0;
            for (int64_t j = 0; j < M; j++) cnt += (sy[2 * j] == a && sy[2 * j + 1] == b);
            if (cnt != 1) return 0;
        }
    auto uniform = [](const std::vector<T> &v) {
        if (v.size() < 2) return true;
        const double step = ((double)v.back() - (double)v.front()) / (double)(v.size() - 1);
        if (!(step > 0) || !isfinite(step)) return false;
        for (size_t a = 0; a < v.size(); a++)
            if (fabs((double)v[a] - ((double)v.front() + step * (double)a)) > 1e-3 * step) return false;
        return true;
    };
    if (!uniform(re) || !uniform(im)) return 0;
    for (T v : re)
        if (!isfinite((double)v)) return 0;
    for (T v : im)
        if (!isfinite((double)v)) return 0;
    for (size_t a = 0; a < re.size(); a++) lre[a] = re[a];
    for (size_t a = 0; a < im.size(); a++) lim[a] = im[a];
    *n_re = (int64_t)re.size();
    *n_im = (int64_t)im.size();
    return 1;
}

}  // namespace qb

using namespace qb;

extern "C" {

int qb_version(void) { return QB_VERSION; }
const char *qb_last_error(void) { return g_err; }
int64_t qb_launch_count(void) { return g_launches.load(); }
int qb_set_train_layout(int layout) { return qb::set_train_layout(layout); }
int qb_set_bps_accumulation(int mode) { return qb::set_bps_accumulation(mode); }
int qb_set_option(const char *name, const char *value) { return qb::set_option(name, value); }

int qb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return set_error(QB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int qb_method_from_name(const char *name)
{
    static const char *names[] = {"cma", "cma2", "sgncma", "mcma", "rde", "mrde", "sbd", "sbd_data", "mddma", "dd"};
    if (name)
        for (int i = 0; i < 10; i++)
            if (strcmp(name, names[i]) == 0) return i;
    return set_error(QB_ERR_ARG, "Unknown method %s", name ? name : "(null)");
}

static int train_check(int dtype, const void *E, int64_t nseg, int64_t nmodes, int64_t TrSyms, int64_t Niter,
                       int64_t os, const void *wx, int64_t ntaps, const int64_t *modes, int64_t nsel,
                       const void *symbols, int64_t K, int method, const void *mu)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(method >= QB_CMA && method <= QB_DD_DATA_REAL, "Unknown method %d", method);
    QB_REQUIRE(E && wx && symbols && mu, "E, wx, symbols and mu must not be NULL");
    QB_REQUIRE(nseg >= 0 && TrSyms >= 0 && Niter >= 0, "nseg, TrSyms and Niter must be non-negative");
    QB_REQUIRE(os >= 1, "oversampling factor must be larger than 0");
    QB_REQUIRE(ntaps >= 1, "ntaps must be >= 1");
    QB_TRY(check_modes(modes, nsel, nmodes));
    QB_REQUIRE(K >= 1, "symbols must hold at least one value per mode");
    QB_REQUIRE((method != QB_SBD_DATA && method != QB_DD_DATA_REAL) || K >= TrSyms,
               "data-aided methods need at least TrSyms training symbols per mode");
    if (nmodes * ntaps > QB_MAX_TAPDIM)
        return set_error(QB_ERR_UNSUPPORTED, "nmodes*ntaps = %lld exceeds %d", (long long)(nmodes * ntaps),
                         QB_MAX_TAPDIM);
    return QB_OK;
}

int qb_train_equaliser_dev(int dtype, const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                           int64_t nmodes, int64_t TrSyms, int64_t Niter, int64_t os, void *wx, int64_t ntaps,
                           const int64_t *modes, int64_t nsel, int adaptive, const void *symbols, int64_t K,
                           int method, void *mu, void *err, void *stream)
{
    QB_TRY(train_check(dtype, E, nseg, nmodes, TrSyms, Niter, os, wx, ntaps, modes, nsel, symbols, K, method, mu));
    return train_dispatch(dtype, E, nseg, seg_stride, row_stride, nmodes, TrSyms, Niter, os, wx, ntaps, modes,
                          nsel, adaptive, symbols, K, method, mu, err, (cudaStream_t)stream);
}

int qb_train_equaliser_host(int dtype, const void *E, int64_t nmodes, int64_t L, int64_t TrSyms, int64_t Niter,
                            int64_t os, void *mu, void *wx, int64_t ntaps, const int64_t *modes, int64_t nsel,
                            int adaptive, const void *symbols, int64_t K, int method, int mu_shared, void *err)
{
    QB_TRY(train_check(dtype, E, 1, nmodes, TrSyms, Niter, os, wx, ntaps, modes, nsel, symbols, K, method, mu));
    QB_REQUIRE(TrSyms == 0 || (TrSyms - 1) * os + ntaps <= L,
               "Field must be longer than the number of training symbols");
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    const size_t cs = csize(dtype), rs = rsize(dtype);
    const size_t nE = (size_t)nmodes * L, nW = (size_t)nmodes * nmodes * ntaps, nS = (size_t)nmodes * K;
    const size_t nErr = (size_t)nmodes * TrSyms * Niter;
    DevBuf dE(st), dW(st), dS(st), dMu(st), dErr(st);
    QB_TRY(dE.alloc(nE * cs));
    QB_TRY(dW.alloc(nW * cs));
    QB_TRY(dS.alloc(nS * cs));
    QB_TRY(dMu.alloc(QB_MAX_MODES * rs));
    if (err) QB_TRY(dErr.alloc(nErr * cs));
    QB_CUDA_CHECK(cudaMemcpyAsync(dE.p, E, nE * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dW.p, wx, nW * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dS.p, symbols, nS * cs, cudaMemcpyHostToDevice, st));
    if (err) QB_CUDA_CHECK(cudaMemsetAsync(dErr.p, 0, nErr * cs, st));  // unselected rows stay 0 (:161)
    unsigned char mubuf[QB_MAX_MODES * 8];
    // one capture, one stream per trained mode: the call takes as long as a stream is deep -> latency layout
    struct LayoutGuard {
        int old;
        LayoutGuard() : old(qb::set_train_layout(QB_LAYOUT_LATENCY)) {}
        ~LayoutGuard() { qb::set_train_layout(old); }
    } layout_guard;
    if (!adaptive || !mu_shared || nsel <= 1) {
        for (int64_t j = 0; j < nsel; j++) memcpy(mubuf + j * rs, mu, rs);
        QB_CUDA_CHECK(cudaMemcpyAsync(dMu.p, mubuf, (nsel ? nsel : 1) * rs, cudaMemcpyHostToDevice, st));
        QB_TRY(train_dispatch(dtype, dE.p, 1, nE, L, nmodes, TrSyms, Niter, os, dW.p, ntaps, modes, nsel, adaptive,
                              dS.p, K, method, dMu.p, err ? dErr.p : nullptr, st));
        if (nsel > 0)
            QB_CUDA_CHECK(cudaMemcpyAsync(mu, (char *)dMu.p + (nsel - 1) * rs, rs, cudaMemcpyDeviceToHost, st));
    } else {
        // interpreted-reference semantics: one mu carried through the modes in list order; the device
        // keeps mu, so mode j+1 simply starts from the slot mode j finished in
        QB_CUDA_CHECK(cudaMemcpyAsync(dMu.p, mu, rs, cudaMemcpyHostToDevice, st));
        for (int64_t j = 0; j < nsel; j++)
            QB_TRY(train_dispatch(dtype, dE.p, 1, nE, L, nmodes, TrSyms, Niter, os, dW.p, ntaps, modes + j, 1,
                                  adaptive, dS.p, K, method, dMu.p, err ? dErr.p : nullptr, st));
        QB_CUDA_CHECK(cudaMemcpyAsync(mu, dMu.p, rs, cudaMemcpyDeviceToHost, st));
    }
    QB_CUDA_CHECK(cudaMemcpyAsync(wx, dW.p, nW * cs, cudaMemcpyDeviceToHost, st));
    if (err) QB_CUDA_CHECK(cudaMemcpyAsync(err, dErr.p, nErr * cs, cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

static int apply_check(int dtype, const void *E, int64_t nseg, int64_t nmodes, int64_t L, int64_t os,
                       const void *wx, int64_t ntaps, const int64_t *modes, int64_t nsel, const void *out)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(os >= 1, "oversampling factor must be larger than 0");
    QB_REQUIRE(ntaps >= 1 && L >= 0 && nseg >= 0, "ntaps must be >= 1 and L, nseg non-negative");
    QB_TRY(check_modes(modes, nsel, nmodes));
    const int64_t N = (L - ntaps + 1) / os;
    QB_REQUIRE(N <= 0 || nsel == 0 || nseg == 0 || (E && wx && out), "E, wx and out must not be NULL");
    return QB_OK;
}

int qb_apply_filter_to_signal_dev(int dtype, const void *E, int64_t nseg, int64_t seg_stride, int64_t row_stride,
                                  int64_t nmodes, int64_t L, int64_t os, const void *wx, int64_t ntaps,
                                  const int64_t *modes, int64_t nsel, void *out, void *stream)
{
    QB_TRY(apply_check(dtype, E, nseg, nmodes, L, os, wx, ntaps, modes, nsel, out));
    return apply_dispatch(dtype, E, nseg, seg_stride, row_stride, nmodes, L, os, wx, ntaps, modes, nsel, out,
                          (cudaStream_t)stream);
}

int qb_apply_filter_to_signal_host(int dtype, const void *E, int64_t nmodes, int64_t L, int64_t os,
                                   const void *wx, int64_t ntaps, const int64_t *modes, int64_t nsel, void *out)
{
    QB_TRY(apply_check(dtype, E, 1, nmodes, L, os, wx, ntaps, modes, nsel, out));
    const int64_t N = (L - ntaps + 1) / os;
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    if (N <= 0 || nsel == 0) return QB_OK;
    const size_t cs = csize(dtype);
    const size_t nE = (size_t)nmodes * L, nW = (size_t)nmodes * nmodes * ntaps, nO = (size_t)nsel * N;
    DevBuf dE(st), dW(st), dO(st);
    QB_TRY(dE.alloc(nE * cs));
    QB_TRY(dW.alloc(nW * cs));
    QB_TRY(dO.alloc(nO * cs));
    QB_CUDA_CHECK(cudaMemcpyAsync(dE.p, E, nE * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dW.p, wx, nW * cs, cudaMemcpyHostToDevice, st));
    QB_TRY(apply_dispatch(dtype, dE.p, 1, nE, L, nmodes, L, os, dW.p, ntaps, modes, nsel, dO.p, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(out, dO.p, nO * cs, cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

static int bps_check(int dtype, const void *E, int64_t nstream, int64_t L, const void *comp, const void *angles,
                     int64_t A, const void *symbols, int64_t M, int64_t n_re, int64_t n_im, int64_t N,
                     const void *ph, const void *Eout)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(nstream >= 0 && L >= 0, "nstream and L must be non-negative");
    QB_REQUIRE(A >= 1 && A <= QB_MAX_ANGLES, "number of test angles must be in [1, %d], got %lld", QB_MAX_ANGLES,
               (long long)A);
    QB_REQUIRE(N >= 1, "averaging block length N must be >= 1");
    QB_REQUIRE(M >= 1 && symbols, "the symbol alphabet must not be empty");
    QB_REQUIRE(comp && (E || L == 0 || nstream == 0), "E and comp must not be NULL");
    QB_REQUIRE((n_re == 0) == (n_im == 0) && n_re >= 0 && n_re <= 64 && n_im <= 64, "invalid slicer levels");
    QB_REQUIRE(n_re == 0 || n_re * n_im == M, "slicer levels do not match the alphabet size");
    QB_REQUIRE(!Eout || ph, "Eout requires ph");
    QB_REQUIRE(!ph || angles, "ph requires the angle table");
    return QB_OK;
}

int qb_bps_dev(int dtype, const void *E, int64_t nstream, int64_t stream_stride, int64_t L, const void *comp,
               const void *angles, int64_t A, const void *symbols, int64_t M, const void *lev_re, int64_t n_re,
               const void *lev_im, int64_t n_im, int64_t N, int32_t *idx, void *ph, void *Eout, void *stream)
{
    QB_TRY(bps_check(dtype, E, nstream, L, comp, angles, A, symbols, M, n_re, n_im, N, ph, Eout));
    return bps_dispatch(dtype, E, nstream, stream_stride, L, comp, angles, A, symbols, M, lev_re, n_re, lev_im,
                        n_im, N, idx, ph, Eout, 0, (cudaStream_t)stream);
}

int qb_bps_rows_dev(int dtype, const void *E, int64_t nstream, int64_t stream_stride, int64_t L, const void *comp,
                    const void *angles, int64_t A, const void *symbols, int64_t M, const void *lev_re,
                    int64_t n_re, const void *lev_im, int64_t n_im, int64_t N, int32_t *idx, void *ph, void *Eout,
                    void *stream)
{
    QB_TRY(bps_check(dtype, E, nstream, L, comp, angles, A, symbols, M, n_re, n_im, N, ph, Eout));
    if (ph || Eout)   // the reference's two-stage tail differs from bps' (whole-array unwrap, phaserecovery.py:282)
        return set_error(QB_ERR_UNSUPPORTED, "qb_bps_rows: only the index search is fused; ph and Eout must be NULL");
    return bps_dispatch(dtype, E, nstream, stream_stride, L, comp, angles, A, symbols, M, lev_re, n_re, lev_im,
                        n_im, N, idx, ph, Eout, 1, (cudaStream_t)stream);
}

int qb_detect_grid_host(int dtype, const void *symbols, int64_t M, void *lev_re, int64_t *n_re, void *lev_im,
                        int64_t *n_im)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(symbols && lev_re && lev_im && n_re && n_im && M >= 1, "invalid arguments");
    *n_re = *n_im = 0;
    if (dtype == QB_C64)
        return detect_grid<float>((const float *)symbols, M, (float *)lev_re, n_re, (float *)lev_im, n_im);
    return detect_grid<double>((const double *)symbols, M, (double *)lev_re, n_re, (double *)lev_im, n_im);
}

static int bps_host_impl(int dtype, const void *E, int64_t nstream, int64_t L, const void *comp, const void *angles,
                         int64_t A, const void *symbols, int64_t M, int64_t N, int32_t *idx, void *ph, void *Eout,
                         int comp_rows);

int qb_bps_host(int dtype, const void *E, int64_t nstream, int64_t L, const void *comp, const void *angles,
                int64_t A, const void *symbols, int64_t M, int64_t N, int32_t *idx, void *ph, void *Eout)
{
    return bps_host_impl(dtype, E, nstream, L, comp, angles, A, symbols, M, N, idx, ph, Eout, 0);
}

int qb_bps_rows_host(int dtype, const void *E, int64_t nstream, int64_t L, const void *comp, const void *angles,
                     int64_t A, const void *symbols, int64_t M, int64_t N, int32_t *idx, void *ph, void *Eout)
{
    if (ph || Eout)
        return set_error(QB_ERR_UNSUPPORTED, "qb_bps_rows: only the index search is fused; ph and Eout must be NULL");
    return bps_host_impl(dtype, E, nstream, L, comp, angles, A, symbols, M, N, idx, ph, Eout, 1);
}

static int bps_host_impl(int dtype, const void *E, int64_t nstream, int64_t L, const void *comp, const void *angles,
                         int64_t A, const void *symbols, int64_t M, int64_t N, int32_t *idx, void *ph, void *Eout,
                         int comp_rows)
{
    QB_TRY(bps_check(dtype, E, nstream, L, comp, angles, A, symbols, M, 0, 0, N, ph, Eout));
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    if (nstream == 0 || L == 0) return QB_OK;
    const size_t cs = csize(dtype), rs = rsize(dtype);
    unsigned char lre[64 * 8], lim[64 * 8];
    int64_t n_re = 0, n_im = 0;
    const int grid = qb_detect_grid_host(dtype, symbols, M, lre, &n_re, lim, &n_im);
    if (grid < 0) return grid;
    if (grid == 0) n_re = n_im = 0;
    const size_t nE = (size_t)nstream * L;
    DevBuf dE(st), dC(st), dA(st), dS(st), dLr(st), dLi(st), dI(st), dP(st), dO(st);
    QB_TRY(dE.alloc(nE * cs));
    const size_t nT = comp_rows ? nE * (size_t)A : (size_t)A;   // one table, or one per symbol and stream
    QB_TRY(dC.alloc(nT * cs));
    QB_TRY(dA.alloc(nT * rs));
    QB_TRY(dS.alloc(M * cs));
    QB_TRY(dLr.alloc(64 * rs));
    QB_TRY(dLi.alloc(64 * rs));
    if (idx) QB_TRY(dI.alloc(nE * sizeof(int32_t)));
    if (ph) QB_TRY(dP.alloc(nE * rs));
    if (Eout) QB_TRY(dO.alloc(nE * cs));
    QB_CUDA_CHECK(cudaMemcpyAsync(dE.p, E, nE * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dC.p, comp, nT * cs, cudaMemcpyHostToDevice, st));
    if (angles) QB_CUDA_CHECK(cudaMemcpyAsync(dA.p, angles, nT * rs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dS.p, symbols, M * cs, cudaMemcpyHostToDevice, st));
    if (n_re) {
        QB_CUDA_CHECK(cudaMemcpyAsync(dLr.p, lre, n_re * rs, cudaMemcpyHostToDevice, st));
        QB_CUDA_CHECK(cudaMemcpyAsync(dLi.p, lim, n_im * rs, cudaMemcpyHostToDevice, st));
    }
    QB_TRY(bps_dispatch(dtype, dE.p, nstream, L, L, dC.p, angles ? dA.p : nullptr, A, dS.p, M, dLr.p, n_re, dLi.p,
                        n_im, N, idx ? (int32_t *)dI.p : nullptr, ph ? dP.p : nullptr, Eout ? dO.p : nullptr,
                        comp_rows, st));
    if (idx) QB_CUDA_CHECK(cudaMemcpyAsync(idx, dI.p, nE * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (ph) QB_CUDA_CHECK(cudaMemcpyAsync(ph, dP.p, nE * rs, cudaMemcpyDeviceToHost, st));
    if (Eout) QB_CUDA_CHECK(cudaMemcpyAsync(Eout, dO.p, nE * cs, cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

int qb_make_decision_dev(int dtype, const void *E, int64_t L, const void *symbols, int64_t M, void *det, void *dist,
                         int32_t *idx, void *stream)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(L >= 0 && M >= 1 && symbols && (E || L == 0), "invalid arguments");
    return make_decision_dispatch(dtype, E, L, symbols, M, det, dist, idx, (cudaStream_t)stream);
}

int qb_make_decision_host(int dtype, const void *E, int64_t L, const void *symbols, int64_t M, void *det, void *dist,
                          int32_t *idx)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(L >= 0 && M >= 1 && symbols && (E || L == 0), "invalid arguments");
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    if (L == 0) return QB_OK;
    const size_t cs = csize(dtype), rs = rsize(dtype);
    DevBuf dE(st), dS(st), dD(st), dR(st), dI(st);
    QB_TRY(dE.alloc(L * cs));
    QB_TRY(dS.alloc(M * cs));
    if (det) QB_TRY(dD.alloc(L * cs));
    if (dist) QB_TRY(dR.alloc(L * rs));
    if (idx) QB_TRY(dI.alloc(L * sizeof(int32_t)));
    QB_CUDA_CHECK(cudaMemcpyAsync(dE.p, E, L * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dS.p, symbols, M * cs, cudaMemcpyHostToDevice, st));
    QB_TRY(make_decision_dispatch(dtype, dE.p, L, dS.p, M, det ? dD.p : nullptr, dist ? dR.p : nullptr,
                                  idx ? (int32_t *)dI.p : nullptr, st));
    if (det) QB_CUDA_CHECK(cudaMemcpyAsync(det, dD.p, L * cs, cudaMemcpyDeviceToHost, st));
    if (dist) QB_CUDA_CHECK(cudaMemcpyAsync(dist, dR.p, L * rs, cudaMemcpyDeviceToHost, st));
    if (idx) QB_CUDA_CHECK(cudaMemcpyAsync(idx, dI.p, L * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

int qb_soft_l_value_demapper_host(int dtype, const void *rx, int64_t N, int64_t num_bits, double snr, const void *bits_map,
                                  int64_t nbits_map, int64_t K, int minmax, double *L_values)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(N >= 0 && num_bits >= 0 && K >= 1 && bits_map && L_values, "invalid arguments");
    QB_REQUIRE(nbits_map >= num_bits, "bits_map has fewer bits than num_bits");   // pythran_dsp.py:96, :124
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    if (N == 0 || num_bits == 0) return QB_OK;
    const size_t cs = csize(dtype);
    DevBuf dR(st), dB(st), dL(st);
    QB_TRY(dR.alloc(N * cs));
    QB_TRY(dB.alloc((size_t)num_bits * K * 2 * cs));
    QB_TRY(dL.alloc((size_t)N * num_bits * sizeof(double)));
    QB_CUDA_CHECK(cudaMemcpyAsync(dR.p, rx, N * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dB.p, bits_map, (size_t)num_bits * K * 2 * cs, cudaMemcpyHostToDevice, st));
    QB_TRY(demapper_dispatch(dtype, dR.p, N, num_bits, snr, dB.p, K, minmax, (double *)dL.p, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(L_values, dL.p, (size_t)N * num_bits * sizeof(double), cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

int qb_estimate_snr_host(int dtype, const void *signal_rx, const void *symbols_tx, int64_t n, const void *gray_symbols,
                         int64_t ngray, double *out3)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(n >= 1 && ngray >= 1 && ngray <= 1024 && signal_rx && symbols_tx && gray_symbols && out3, "invalid arguments");
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    const size_t cs = csize(dtype);
    DevBuf dR(st), dT(st), dG(st), dA(st), dM(st);
    QB_TRY(dR.alloc(n * cs));
    QB_TRY(dT.alloc(n * cs));
    QB_TRY(dG.alloc(ngray * cs));
    QB_TRY(dA.alloc((size_t)ngray * 4 * sizeof(double)));
    QB_TRY(dM.alloc((size_t)ngray * 2 * sizeof(double)));
    QB_CUDA_CHECK(cudaMemcpyAsync(dR.p, signal_rx, n * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dT.p, symbols_tx, n * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dG.p, gray_symbols, ngray * cs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemsetAsync(dA.p, 0, (size_t)ngray * 4 * sizeof(double), st));
    double *acc1 = (double *)dA.p, *acc2 = acc1 + 3 * ngray;
    QB_TRY(snr_pass_dispatch(dtype, dR.p, dT.p, n, dG.p, ngray, nullptr, acc1, 1, st));
    std::vector<double> h1(3 * ngray), means(2 * ngray), h2(ngray);
    QB_CUDA_CHECK(cudaMemcpyAsync(h1.data(), acc1, 3 * ngray * sizeof(double), cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int64_t c = 0; c < ngray; c++) {
        means[2 * c] = h1[3 * c + 1] / h1[3 * c];          // 0/0 = NaN for an unused alphabet point, like np.mean([])
        means[2 * c + 1] = h1[3 * c + 2] / h1[3 * c];
    }
    QB_CUDA_CHECK(cudaMemcpyAsync(dM.p, means.data(), 2 * ngray * sizeof(double), cudaMemcpyHostToDevice, st));
    QB_TRY(snr_pass_dispatch(dtype, dR.p, dT.p, n, dG.p, ngray, (const double *)dM.p, acc2, 2, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(h2.data(), acc2, ngray * sizeof(double), cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    double in_pow = 0., N0 = 0.;
    for (int64_t c = 0; c < ngray; c++) {                  // pythran_dsp.py:272-281
        const double K = h1[3 * c], Px = K / (double)n;
        N0 += (h2[c] / K) * Px;
        in_pow += (means[2 * c] * means[2 * c] + means[2 * c + 1] * means[2 * c + 1]) * Px;
    }
    out3[0] = in_pow / N0;
    out3[1] = in_pow;
    out3[2] = N0;
    return QB_OK;
}

int qb_freq_shift_dev(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, const double *freq,
                      int64_t os, int64_t t0, void *out, int64_t out_stride, void *stream)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(nrows >= 0 && L >= 0 && os >= 1 && t0 >= 0, "invalid sizes");
    QB_REQUIRE(nrows == 0 || L == 0 || (E && freq && out), "E, freq and out must not be NULL");
    QB_REQUIRE(nrows <= 65535, "at most 65535 rows");
    return freq_shift_dispatch(dtype, E, nrows, row_stride, L, freq, os, t0, out, out_stride, (cudaStream_t)stream);
}

int qb_pilot_cpe_dev(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t nlen, const int64_t *pilot_idx,
                     const void *pilots, int64_t pilot_stride, int64_t npilots, int64_t num_average, void *out,
                     int64_t out_stride, void *trace, int64_t trace_stride, void *stream)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(nrows >= 0 && nlen >= 0, "invalid sizes");
    QB_REQUIRE(num_average > 1 && (num_average % 2) == 1, "num_average must be odd and at least 3");
    QB_REQUIRE(npilots >= num_average, "Larger averaging block size than total number of pilot symbols");
    QB_REQUIRE(nrows == 0 || nlen == 0 || (E && pilot_idx && pilots && out), "E, pilot_idx, pilots and out must not be NULL");
    return pilot_cpe_dispatch(dtype, E, nrows, row_stride, nlen, pilot_idx, pilots, pilot_stride, npilots, num_average,
                              out, out_stride, trace, trace_stride, (cudaStream_t)stream);
}

static int vv_check(int64_t nrows, int64_t L, int64_t N, int64_t M)
{
    QB_REQUIRE(nrows >= 0 && nrows <= 65535, "between 0 and 65535 rows");
    QB_REQUIRE(N >= 1 && L >= N, "the averaging length N must be between 1 and the signal length");
    QB_REQUIRE(M >= 1 && M <= 1024, "PSK order M out of range");
    return QB_OK;
}

int qb_synth_upsample_dev(const void *symbols, int64_t nrows, int64_t n, int64_t up, void *out, int64_t out_stride,
                          void *stream)
{
    QB_REQUIRE(nrows >= 0 && n >= 0 && up >= 1 && out_stride >= n * up, "invalid sizes");
    QB_REQUIRE(nrows <= 65535, "at most 65535 rows");
    QB_REQUIRE(nrows == 0 || out_stride == 0 || (symbols && out), "symbols and out must not be NULL");
    return synth_upsample_dispatch(symbols, nrows, n, up, out, out_stride, (cudaStream_t)stream);
}

int qb_synth_specmul_dev(void *X, int64_t nrows, int64_t nfft, const void *H, void *stream)
{
    QB_REQUIRE(nrows >= 0 && nrows <= 65535 && nfft >= 0, "invalid sizes");
    QB_REQUIRE(nrows == 0 || nfft == 0 || (X && H), "X and H must not be NULL");
    return synth_specmul_dispatch(X, nrows, nfft, H, (cudaStream_t)stream);
}

int qb_synth_crop_norm_dev(const void *x, int64_t nrows, int64_t row_stride, int64_t first, int64_t down, int64_t n,
                           const double *target_power, int renormalise, void *out, void *stream)
{
    QB_REQUIRE(nrows >= 0 && nrows <= 65535 && n >= 0 && first >= 0 && down >= 1, "invalid sizes");
    QB_REQUIRE(n == 0 || first + (n - 1) * down < row_stride, "crop reaches past the row");
    QB_REQUIRE(nrows == 0 || n == 0 || (x && out && (!renormalise || target_power)), "x, out (and target_power) must not be NULL");
    if (nrows == 0 || n == 0) return QB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf mom(st);
    QB_TRY(mom.alloc((size_t)nrows * 3 * sizeof(double)));
    return synth_crop_norm_dispatch(x, nrows, row_stride, first, down, n, target_power, renormalise, (double *)mom.p, out, st);
}

int qb_synth_pmd_dev(void *spectra, int64_t n, double theta, double t_dgd, double fs, void *stream)
{
    QB_REQUIRE(n >= 0, "invalid sizes");
    QB_REQUIRE(n == 0 || spectra, "spectra must not be NULL");
    return synth_pmd_dispatch(spectra, n, theta, t_dgd, fs, (cudaStream_t)stream);
}

int qb_synth_tail_dev(int dtype, const void *x, int64_t nrows, int64_t n, const double *noise_sigma, double walk_sigma,
                      uint64_t seed, int64_t row0, int64_t index0, const double *phase0, void *out, int64_t out_stride,
                      double *phase_out, void *stream)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(nrows >= 0 && nrows <= 65535 && n >= 0 && out_stride >= n && row0 >= 0 && index0 >= 0, "invalid sizes");
    QB_REQUIRE(index0 % 2 == 0, "index0 must be even (walk steps are drawn in pairs)");
    QB_REQUIRE(nrows == 0 || n == 0 || (x && out), "x and out must not be NULL");
    if (nrows == 0 || n == 0) return QB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf tiles(st);
    QB_TRY(tiles.alloc((size_t)nrows * ((n + 2047) / 2048) * sizeof(double)));
    return synth_tail_dispatch(dtype, x, nrows, n, noise_sigma, walk_sigma, seed, row0, (uint64_t)index0, phase0,
                               (double *)tiles.p, out, out_stride, phase_out, st);
}

int qb_viterbiviterbi_dev(int dtype, const void *E, int64_t nrows, int64_t row_stride, int64_t L, int64_t N, int64_t M,
                          void *out, int64_t out_stride, void *ph, int64_t ph_stride, void *stream)
{
    QB_TRY(check_dtype(dtype));
    QB_TRY(vv_check(nrows, L, N, M));
    QB_REQUIRE(nrows == 0 || (E && out && ph), "E, out and ph must not be NULL");
    if (nrows == 0) return QB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf work(st);                                        // per-tile turn counts; stream-ordered, freed after the kernels
    QB_TRY(work.alloc(vv_work_bytes(dtype, nrows, L, N)));
    return vv_dispatch(dtype, E, nrows, row_stride, L, N, M, out, out_stride, ph, ph_stride, work.p, st);
}

int qb_viterbiviterbi_host(int dtype, const void *E, int64_t nrows, int64_t L, int64_t N, int64_t M, void *out, void *ph)
{
    QB_TRY(check_dtype(dtype));
    QB_TRY(vv_check(nrows, L, N, M));
    QB_REQUIRE(nrows == 0 || (E && out && ph), "E, out and ph must not be NULL");
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    if (nrows == 0) return QB_OK;
    const size_t cs = csize(dtype), rs = rsize(dtype);
    const int64_t nwin = L - N + 1;
    DevBuf dE(st), dO(st), dP(st), work(st);
    QB_TRY(dE.alloc((size_t)nrows * L * cs));
    QB_TRY(dO.alloc((size_t)nrows * L * cs));
    QB_TRY(dP.alloc((size_t)nrows * nwin * rs));
    QB_TRY(work.alloc(vv_work_bytes(dtype, nrows, L, N)));
    QB_CUDA_CHECK(cudaMemcpyAsync(dE.p, E, (size_t)nrows * L * cs, cudaMemcpyHostToDevice, st));
    QB_TRY(vv_dispatch(dtype, dE.p, nrows, L, L, N, M, dO.p, L, dP.p, nwin, work.p, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(out, dO.p, (size_t)nrows * L * cs, cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(ph, dP.p, (size_t)nrows * nwin * rs, cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

int qb_select_angles_dev(int dtype, const void *angles, int64_t p, int64_t A, const int64_t *idx, int64_t L,
                         void *out, void *stream)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(angles && idx && out && A >= 1 && p >= 1 && L >= 0, "invalid arguments");
    return select_angles_dispatch(dtype, angles, p, A, idx, L, out, (cudaStream_t)stream);
}

int qb_select_angles_host(int dtype, const void *angles, int64_t p, int64_t A, const int64_t *idx, int64_t L,
                          void *out)
{
    QB_TRY(check_dtype(dtype));
    QB_REQUIRE(A >= 1 && p >= 1 && L >= 0, "invalid arguments");
    QB_REQUIRE(L == 0 || (angles && idx && out), "angles, idx and out must not be NULL");
    for (int64_t i = 0; i < L; i++)
        QB_REQUIRE(idx[i] >= 0 && idx[i] < A, "angle index out of range at %lld", (long long)i);
    cudaStream_t st;
    QB_TRY(host_stream(&st));
    if (L == 0) return QB_OK;
    const size_t rs = rsize(dtype);
    DevBuf dA(st), dI(st), dO(st);
    QB_TRY(dA.alloc((size_t)p * A * rs));
    QB_TRY(dI.alloc((size_t)L * sizeof(int64_t)));
    QB_TRY(dO.alloc((size_t)L * rs));
    QB_CUDA_CHECK(cudaMemcpyAsync(dA.p, angles, (size_t)p * A * rs, cudaMemcpyHostToDevice, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(dI.p, idx, (size_t)L * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    QB_TRY(select_angles_dispatch(dtype, dA.p, p, A, (const int64_t *)dI.p, L, dO.p, st));
    QB_CUDA_CHECK(cudaMemcpyAsync(out, dO.p, (size_t)L * rs, cudaMemcpyDeviceToHost, st));
    QB_CUDA_CHECK(cudaStreamSynchronize(st));
    return QB_OK;
}

}  // extern "C"
