/*
 * xrft_b200 -- C-ABI of the B200-native spectral engine behind the xrft API.
 *
 * The reference (xgcm/xrft @ efc1c30) is pure Python and has no FFI.  Its numeric seam --
 * the only place raw buffers are handed to third-party code -- is:
 *
 *   (S1) the module object returned by `_fft_module`            xrft/xrft.py:32-36
 *        used as  fftm.fftn / rfftn / ifftn / irfftn / fftshift / ifftshift
 *                                                              xrft/xrft.py:398-404, 439-447, 586-591, 612-621
 *   (S2) the detrend ufuncs   `da.mean`, `scipy.signal.detrend`,
 *        `_detrend_2d_ufunc(arr)`, `_detrend_3d_ufunc(arr)`     xrft/detrend.py:55, 65-71, 100-113, 116-138
 *   (S3) the window product   `scipy_win_func(n, sym=False)` outer product multiply
 *                                                              xrft/xrft.py:83-103
 *   (S4) spectrum algebra     |F|^2, F conj(G), angle, scalings xrft/xrft.py:740-748, 825-833, 865-869
 *   (S5) `_binned_agg(array, indices, num_bins, func=sum)`      xrft/xrft.py:877-907
 *
 * Every entry point below replaces one of those seams (cited per function) or fuses several
 * of them into one pass.  Conventions: plain pointers and sizes only; all data pointers are
 * DEVICE pointers owned by the caller (torch.Tensor.data_ptr() in the Python binding); inputs
 * are never modified; no hidden device allocations except cached twiddle tables; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream); every function returns 0 on success or
 * a negative XRFTB_E* code with a message retrievable by xrftb_last_error() (thread-local).
 * Arrays are C-contiguous (row-major).  dtype: 0 = float32 / complex64, 1 = float64 / complex128.
 * There is NO CPU fallback: without a CUDA device the compute entry points return XRFTB_ECUDA.
 */
#ifndef XRFT_B200_H
#define XRFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XRFTB_VERSION 100 /* 0.1.0 */

enum { XRFTB_OK = 0, XRFTB_EINVAL = -1, XRFTB_EUNSUPPORTED = -2, XRFTB_ECUDA = -3, XRFTB_EWORKSPACE = -4 };
enum { XRFTB_F32 = 0, XRFTB_F64 = 1 };
/* transform kinds, numpy semantics (forward unnormalised, inverse 1/N) */
enum { XRFTB_C2C_FWD = 0, XRFTB_C2C_INV = 1, XRFTB_R2C = 2, XRFTB_C2R = 3 };
/* epilogue modes */
enum {
    XRFTB_EPI_COMPLEX = 0,    /* F * scale * ramps                         (xrft.fft,            xrft.py:446-472) */
    XRFTB_EPI_POWER = 1,      /* |F|^2 * scale * weight                    (power_spectrum,      xrft.py:740-748) */
    XRFTB_EPI_CROSS = 2,      /* F1 conj(F2) * scale * ramps * weight      (cross_spectrum,      xrft.py:825-833) */
    XRFTB_EPI_PHASE = 3,      /* angle(F1 conj(F2) * ramps)                (cross_phase,         xrft.py:865-869) */
    XRFTB_EPI_BINS_POWER = 4, /* radial-bin sum of POWER, no spectrum out  (isotropize,          xrft.py:895-906) */
    XRFTB_EPI_BINS_CROSS = 5  /* radial-bin sum of CROSS (re,im), no spectrum out */
};

int xrftb_version(void);
const char* xrftb_last_error(void);
/* sm_count, compute capability and opt-in shared memory per block of the current device */
int xrftb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);
/* number of kernels this library has launched since the last reset (bench.py's gpu_launches) */
long xrftb_launch_count(int reset);
/* per-kernel-class CUDA-event timing of xrftb_spectrum2d: begin() arms it, end() synchronises and returns
 * summed milliseconds and launch counts for {moments, row pass, column pass, mirror fill}. */
int xrftb_profile_begin(void);
int xrftb_profile_end(double ms[4], long counts[4]);
/* Chain-selection options (process-wide; every value selects a kernel chain that is also the regular path of some shape or
 * dtype, and tests/test_gpu_variants.py runs each of them against the reference formulas):
 *   "cols_first"  1  columns-first chain for the full-width power spectrum and the radial bins (0: rows first + mirror pass)
 *   "zpack"       1  pass 1 leaves packed column spectra ("z mode") where pass 2 supports it
 *   "ztma"        1  TMA tensor stores of the packed column spectra
 *   "cols_async"  2  TMA-fed column kernels: 0 never, 1 always, 2 where measured faster
 *   "rowline"     1  float32 detrend as exact per-line subtraction + rank-2 completion (0: moments pass + fp64 plane)
 *   "cross_z"     1  two-field z-mode chain for cross spectrum / phase (0: rows-first two-field chain)
 *   "bins_static" 1  radial-bin kernels with a static cell-to-bin mapping (0: generic LUT epilogue)
 * The environment variable XRFTB_<NAME> (upper case) only supplies the value an option starts with. */
int xrftb_set_option(const char* name, int value);
int xrftb_get_option(const char* name, int* value);
/* which kernel chain the last xrftb_spectrum2d call OF THE CALLING THREAD took: 0 = rows first (moments | row R2C | column pass | Hermitian
 * mirror), 1 = columns first (column R2C with column-line detrend | completion tables | row C2C + epilogue, no mirror pass),
 * 2 = columns first in "z mode" (pass 1 leaves the packed column spectra, pass 2 separates the real columns in its loads),
 * 3 = the two-field z-mode chain (cross spectrum / phase).
 * In chains 1 and 2 the profile classes read {completion tables, row pass = pass 2, column pass = pass 1, unused}. */
int xrftb_spectrum2d_last_path(void);

/* ---- (S1) np.fft.fftn / ifftn / rfftn / irfftn --------------------------------------------------
 * N-D transform over `axes` of a C-contiguous array.  `shape[ndim]` is the REAL-SPACE shape.
 *   C2C_FWD/INV : in, out complex, same shape; any distinct axes.
 *   R2C         : in real `shape`; out complex with last dim shape[ndim-1]/2+1; the last listed axis
 *                 must be ndim-1 (the reference always moves real_dim last: xrft.py:386,396).
 *   C2R         : in complex (last dim N/2+1), out real `shape`; same axis rule; needs a workspace of
 *                 the input's size when naxes > 1 (input is never modified).
 * Lengths: powers of two up to 2^14 (f32) / 2^13 (f64) on the contiguous axis (R2C/C2R: twice that) and
 * up to 8192 on strided axes run in one Stockham pass; longer powers of two (to 2^26) run the four-step
 * decomposition n = n1*n2 on the same kernels; non-smooth lengths <= 64 run a direct DFT; lengths 2^a 3^b 5^c 7^d up to
 * 12800 (f32) / 6400 (f64) run a mixed-radix shared-memory kernel (one pass); every other length runs Bluestein's
 * chirp-z on top of the power-of-two machinery.  `work` must hold xrftb_fftn_workspace(...) bytes
 * (0 for single-pass power-of-two C2C/R2C).  in == out is allowed for C2C. */
size_t xrftb_fftn_workspace(int dtype, int kind, int ndim, const int64_t* shape, int naxes, const int* axes);
int xrftb_fftn(const void* in, void* out, void* work, size_t work_bytes, int dtype, int kind, int ndim,
               const int64_t* shape, int naxes, const int* axes, void* stream);

/* ---- (S1) + neighbours: 2-D real transform with xrft.fft / xrft.ifft's elementwise steps folded into its passes -------
 * One call == pad -> rfftn -> x phase ramps x prod(dx)        (forward: xrft.py:398-404, 462-472; padding.py:157-181), or
 *             x phase ramps, ifftshift -> irfftn -> fftshift / ifftshift, / prod(df), unpad crop
 *                                                               (inverse: xrft.py:574-621, 641-642; padding.py:425-446)
 * over the two trailing axes of [batch][ny][nx] real <-> [batch][ny][nx/2+1] complex, numpy normalisation (inverse 1/(ny nx)).
 *   forward (inverse = 0): `in` is real [batch][in_ny][in_nx] sitting at (in_off_y, in_off_x) inside a zero [ny][nx] grid that is
 *     never materialised (in_ny = 0: `in` has the full shape); out = rfft2 * ramp_y[ky] * ramp_x[kx] * scale.
 *   inverse (inverse = 1): transform row r reads input row (r + in_roll_y) % ny (sortby + ifftshift of xrft.ifft folded into
 *     the loads) multiplied by ramp_y[that input row] * ramp_x[kx]; real result element (i, j) lands at ((i + out_roll_y) % ny,
 *     (j + out_roll_x) % nx), multiplied by scale; only the out_ny x out_nx box at (out_off_y, out_off_x) of that rolled
 *     result is stored, densely (out_ny = 0: everything).
 * ramps are complex vectors of the transform dtype, nullable.  Power-of-two sizes (nx within the single-pass row limits, ny
 * up to 2^26 by the four-step decomposition); other sizes return XRFTB_EUNSUPPORTED.  `in` is never modified. */
typedef struct {
    int dtype;
    int inverse;
    int64_t batch, ny, nx;
    const void* in;
    void* out;
    int64_t in_ny, in_nx, in_off_y, in_off_x;
    const void* ramp_y;
    const void* ramp_x;
    double scale;
    int64_t in_roll_y;
    int64_t out_roll_y, out_roll_x;
    int64_t out_ny, out_nx, out_off_y, out_off_x;
    void* work;
    size_t work_bytes;
} xrftb_fft2r_desc;
size_t xrftb_fft2r_workspace(const xrftb_fft2r_desc* desc);
int xrftb_fft2r(const xrftb_fft2r_desc* desc, void* stream);

/* ---- (S2) detrend -------------------------------------------------------------------------------
 * Real input viewed as [batch][n0][n1][n2] (use 1 for unused leading dims).  `moments` receives
 * per item {S, S0, S1, S2}: the sum and the centred first moments sum((i_d - (n_d-1)/2) * x).
 * On a full regular grid the least-squares (hyper)plane of xrft/detrend.py:100-138 and
 * scipy.signal.detrend (detrend.py:65-71) is   mean + sum_d S_d / V_d * (i_d - (n_d-1)/2),
 * V_d = npts * (n_d^2 - 1) / 12  (orthogonal regressors), which detrend_apply evaluates in fp64. */
int xrftb_moments(const void* in, double* moments, int dtype, int64_t batch, int64_t n0, int64_t n1, int64_t n2,
                  void* stream);
/* out = (in - trend) * w0[i0] * w1[i1] * w2[i2];  detrend: 0 none, 1 constant (da.mean, detrend.py:55),
 * 2 linear; windows nullable (S3, xrft.py:96-103).  in == out allowed. */
int xrftb_detrend_window(const void* in, void* out, const double* moments, int detrend, const void* w0, const void* w1,
                         const void* w2, int dtype, int64_t batch, int64_t n0, int64_t n1, int64_t n2, void* stream);

/* ---- (S4) generic spectral epilogue over up to 3 trailing transform axes --------------------------
 * in1 (and in2 for CROSS/PHASE): complex [batch][k0][k1][k2in]; hermitian != 0 means the inputs are
 * rfftn half spectra (k2in = k2/2+1) of REAL fields, expanded to the full k2 width on output unless
 * keep_half.  shift[d]: 0 none, 1 fftshift (xrft.py:446-447), 2 ifftshift (xrft.py:612-614) on axis d.  ramp[d]: complex vectors indexed by the
 * UNSHIFTED frequency index of axis d (phase ramp xrft.py:462-469; for CROSS the caller passes
 * ramp1*conj(ramp2)), nullable.  weight: real vector on axis 2 (one-sided x2, xrft.py:673-682), nullable.
 * out: complex (COMPLEX, CROSS) or real (POWER, PHASE), [batch][k0][k1][W], W = keep_half ? k2/2+1 : k2. */
int xrftb_spectral_post(const void* in1, const void* in2, void* out, int dtype, int mode, int64_t batch, int64_t k0,
                        int64_t k1, int64_t k2, int hermitian, int keep_half, const int* shift, const void* const* ramp,
                        const void* weight, double scale, void* stream);

/* The same epilogue with the Welch segment mean folded in (chunks_to_segments=True followed by .mean over the `<dim>_segment`
 * axis: xrft.py:106-136, 390-391; tests/test_xrft.py:273-337): the batch is viewed as [outer][seg_n][seg_inner] and `out`
 * ([outer][seg_inner][k0][k1][W]) receives the mean over seg_n -- the per-segment spectra are never written.  Modes COMPLEX,
 * POWER, CROSS (a mean of phases is not a phase).  `out` is zeroed by the call; accumulation uses global atomics. */
int xrftb_spectral_post_segmean(const void* in1, const void* in2, void* out, int dtype, int mode, int64_t batch, int64_t k0,
                                int64_t k1, int64_t k2, int hermitian, int keep_half, const int* shift, const void* const* ramp,
                                const void* weight, double scale, int64_t seg_n, int64_t seg_inner, void* stream);

/* circular roll + scale over up to 3 trailing axes, real or complex: out[(i + s) % n] = in[i] * scale.
 * Replaces the output-side fftm.ifftshift / fftm.fftshift and the `/ prod(spacing)` of xrft.ifft
 * (xrft.py:617-621, 641-642).  Out of place only. */
int xrftb_roll_scale(const void* in, void* out, int dtype, int is_complex, int64_t batch, int64_t n0, int64_t n1, int64_t n2,
                     int64_t s0, int64_t s1, int64_t s2, double scale, void* stream);

/* ---- the data movement around the transforms that the reference does with xarray / numpy views: moving the transform axes
 * last (da.transpose, xrft.py:386-396, 474-476) and reversing axes with decreasing coordinates (xrft.py:436-441).
 * out (C-contiguous) has shape in_shape[perm[0]], in_shape[perm[1]], ...; axis p of the INPUT is reversed when flip[p] != 0
 * (flip nullable).  ndim <= 6 (fold unpermuted neighbours); elem_bytes 4 / 8 / 16.  Out of place. */
int xrftb_permute(const void* in, void* out, int elem_bytes, int ndim, const int64_t* in_shape, const int* perm, const int* flip,
                  void* stream);

/* ---- xrft.pad (padding.py:157-181 -> DataArray.pad -> numpy.pad) on device-resident data ---------------
 * out[ndim] = in surrounded by pad_before[a] / pad_after[a] cells on axis a (ndim <= 4; fold leading axes).  mode: 0 constant
 * (`fill` points at one element, NULL = zeros), 1 edge, 2 reflect, 3 symmetric, 4 wrap (numpy semantics, reflect_type "even",
 * pads wider than the array allowed).  elem_bytes 4 / 8 / 16 (float32; float64 or complex64; complex128).  unpad needs no
 * kernel: it is a strided view of the result (padding.py:425-446). */
int xrftb_pad(const void* in, void* out, int elem_bytes, int ndim, const int64_t* in_shape, const int64_t* pad_before,
              const int64_t* pad_after, int mode, const void* fill, void* stream);

/* ---- (S5) _binned_agg(func="sum") ---------------------------------------------------------------
 * array: real (is_complex = 0) or complex [batch][ncell]; lut: int32 [ncell], negative = masked
 * (the NaN mask of xrft.py:895-896); bins: float64 [batch][nbins] or [batch][nbins][2], ACCUMULATED
 * into (caller zeroes). */
int xrftb_binned_sum(const void* array, const int32_t* lut, double* bins, int dtype, int is_complex, int64_t batch,
                     int64_t ncell, int nbins, void* stream);

/* ---- fused hot path: detrend + window + 2-D real FFT + spectrum epilogue ---------------------------
 * One call == power_spectrum / cross_spectrum / cross_phase / isotropic_* / fft of REAL fields over the
 * two trailing axes (xrft.py:685-750 and callees): S2+S3 fused into the row-pass loads, S1 as an
 * R2C row pass + strided column pass through a blocked half-spectrum intermediate, S4 (+S5) fused
 * into the column-pass stores.  in1/in2: real [batch][ny][nx]; ny, nx powers of two,
 * 2 <= ny <= 8192 (4096 for two-field modes), 4 <= nx <= 32768 (f32) / 16384 (f64).
 * Output position/shape as xrftb_spectral_post with (k1, k2) = (ny, nx).
 * BINS modes: lut int32 [ny][W] (bin of each OUTPUT cell, negative = skip), bins accumulated. */
typedef struct {
    int dtype;
    int64_t batch;
    int ny, nx;
    const void* in1;
    const void* in2;
    int detrend;          /* 0 none, 1 constant, 2 linear */
    const void* win_y;    /* T[ny] or NULL */
    const void* win_x;    /* T[nx] or NULL */
    int mode;             /* XRFTB_EPI_* */
    int keep_half;        /* 1: real_dim semantics, output width nx/2+1, no x shift */
    int shift_y, shift_x;
    double scale;
    const void* ramp_y;   /* complex T[ny], unshifted index, or NULL */
    const void* ramp_x;   /* complex T[W], unshifted index, or NULL */
    const void* weight_x; /* T[nx/2+1] or NULL (keep_half only) */
    void* out;
    const int32_t* lut;
    double* bins;
    int nbins;
    int lut_symmetric;    /* 1 if lut[-ky][-kx] == lut[ky][kx] for all cells (radial bins): lets mirrored cells reuse the bin */
    void* work;
    size_t work_bytes;
    void* out2;           /* mode CROSS only, nullable: ALSO write angle(F1 conj(F2) ramps), real [batch][ny][W] -- cross_spectrum and
                             cross_phase (xrft.py:753-874) from one read of the two fields instead of the reference's two pipelines */
} xrftb_spectrum2d_desc;

/* minimum workspace (one batch item in flight) and the size that keeps `batch` items in flight */
size_t xrftb_spectrum2d_workspace(int dtype, int ny, int nx, int two_fields, int64_t batch_in_flight);
int xrftb_spectrum2d(const xrftb_spectrum2d_desc* desc, void* stream);

/* ---- (e) multi-GPU: the one exchange step of the path -----------------------------------------------
 * The reference shards chunks of the non-transform axes over dask workers (xrft.py:32-36); transforms never cross
 * workers, so fft / power_spectrum / cross_* need no collective.  Averaging an isotropic spectrum over the sharded axis
 * (isotropic_power_spectrum(...).mean(dim), xrft.py:1013-1095; tests/test_xrft.py:1011-1013) needs ONE sum of
 * nbins (+1 count) float64 over the ranks: xrftb_allreduce_bins == ncclAllReduce(sum, float64, in place) on `stream`.
 * One process per GPU; rank 0 calls xrftb_comm_unique_id and ships the 128 bytes to the other ranks through any side
 * channel; every rank then calls xrftb_comm_init on its current device.  NCCL is bound at run time (dlopen), so these
 * return XRFTB_EUNSUPPORTED when libnccl.so.2 is absent. */
#define XRFTB_COMM_ID_BYTES 128
int xrftb_comm_unique_id(void* id128);
int xrftb_comm_init(void** comm, int nranks, int rank, const void* id128);
int xrftb_comm_destroy(void* comm);
int xrftb_comm_nccl_version(void);
int xrftb_allreduce_bins(void* comm, double* buf, size_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XRFT_B200_H */
