// xrft_b200 -- templated launchers (included by the per-dtype instantiation units).
#pragma once
#include "internal.h"

namespace xrftb {

// Per-kernel launch preparation, cached PER DEVICE: the opt-in dynamic shared-memory attribute is a per-device property of
// the function, and the occupancy (hence the persistent grid size) belongs to the device that answered the query.  A
// process may drive several GPUs (tensors on any device); concurrent first calls write the same values (benign).
constexpr int kMaxDevices = 64;
struct DevOcc { int v[kMaxDevices]; DevOcc() { for (int& x : v) x = -1; } };
template <class Kern>
static inline int prepare_kernel(Kern kern, int threads, size_t smem, DevOcc* cache, int* occ_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { set_error("no current CUDA device"); return -3; }
    if (cache->v[dev] < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e)); return -3; }
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
        if (e != cudaSuccess || occ < 1) { set_error("occupancy query failed (threads=%d smem=%zu): %s", threads, smem, cudaGetErrorString(e)); return -3; }
        cache->v[dev] = occ;
    }
    *occ_out = cache->v[dev];
    return 0;
}

constexpr int kMaxFusedBins = 1024;
constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// sequences per CTA on the contiguous-axis pass
template <int LOG2L, int LOGE> constexpr int rows_seq_generic() { return cmax(1, cmin(128, 256 >> (LOG2L - LOGE))); }
#ifndef XRFTB_FUSED_CTA_THREADS
#define XRFTB_FUSED_CTA_THREADS 512
#endif
template <int LOG2L, int LOGE> constexpr int rows_seq_fused() { return cmax(rows_seq_generic<LOG2L, LOGE>(), cmax(1, cmin(4, XRFTB_FUSED_CTA_THREADS >> (LOG2L - LOGE)))); }

template <typename T, int LOG2L, int SEQ, class IO>
static int launch_rows(const IO& io, long nseq, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<T>::LOGE, LOG2L);
    using G_ = Geometry<LOG2L, LOGE>;
    auto kern = rows_kernel<T, LOG2L, LOGE, SEQ, IO>;
    constexpr int threads = G_::NT * SEQ;
    constexpr size_t smem = (size_t)SEQ * (G_::LPAD + IO::kSeqSkew) * sizeof(cplx<T>);
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const cplx<T>* tw = twiddle_fft<T>(LOG2L);
    if (!tw) return -3;
    long ngroups = (nseq + SEQ - 1) / SEQ;
    long grid = (long)sm_count() * occ;
    if (grid > ngroups) grid = ngroups;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, nseq);
    return check_launch("rows_kernel");
}

template <typename T, int LOG2L, int PAIRS, bool ROWLINE>
static int launch_rows2(const RowsR2CFused<T>& io, long nseq, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<T>::LOGE, LOG2L);
    using G_ = Geometry<LOG2L, LOGE>;
    auto kern = rows2_kernel<T, LOG2L, LOGE, PAIRS, ROWLINE>;
    constexpr int threads = G_::NT * PAIRS;
    // + row-line partial sums [PAIRS][WPP][4] and lines [PAIRS][4] (floats)
    constexpr int WPP = G_::NT < 32 ? 1 : G_::NT / 32;
    constexpr size_t smem = (size_t)PAIRS * (2 * G_::LPAD + 8) * sizeof(cplx<T>) + (size_t)PAIRS * (WPP + 1) * 4 * sizeof(float);
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const cplx<T>* tw = twiddle_fft<T>(LOG2L);
    if (!tw) return -3;
    long ngroups = nseq / (2 * PAIRS);
    long grid = (long)sm_count() * occ;
    if (grid > ngroups) grid = ngroups;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, nseq);
    return check_launch("rows2_kernel");
}

template <typename T, int LOG2L, int PAIRS>
static int launch_rows2c_power(const RowsC2CPower<T>& io, long nseq, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<T>::LOGE, LOG2L);
    using G_ = Geometry<LOG2L, LOGE>;
    auto kern = rows2c_power_kernel<T, LOG2L, LOGE, PAIRS>;
    constexpr int threads = G_::NT * PAIRS;
    constexpr size_t smem = (size_t)PAIRS * (2 * G_::LPAD + 8) * sizeof(cplx<T>);
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const cplx<T>* tw = twiddle_fft<T>(LOG2L);
    if (!tw) return -3;
    long ngroups = (nseq + 2 * PAIRS - 1) / (2 * PAIRS);
    long grid = (long)sm_count() * occ;
    if (grid > ngroups) grid = ngroups;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, nseq);
    return check_launch("rows2c_power_kernel");
}

// pass 2 + radial bins (rows_bins_kernel).  Returns 1 when the launch shape does not fit (caller falls back).
template <int LOG2L, int SEQ>
static int launch_rows_bins(const RowsBins& io, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<float>::LOGE, LOG2L);
    using G_ = Geometry<LOG2L, LOGE>;
    auto kern = rows_bins_kernel<LOG2L, LOGE, SEQ>;
    constexpr int threads = G_::NT * SEQ;
    const int P = io.rows == 1 ? SEQ : 1;
    const int nseg = P * io.nbins;
    if (nseg > 4096 || (io.rows != 1 && io.rows % SEQ != 0)) return 1;
    constexpr size_t smem_x = (size_t)SEQ * G_::LPAD * sizeof(float2);
    static_assert(smem_x >= (size_t)SEQ * (1 << LOG2L) * sizeof(float), "the staging array fits the exchange buffer");
    constexpr size_t smem = smem_x + (size_t)SEQ * (1 << LOG2L) * sizeof(unsigned short) + 4096 * sizeof(int);   // + slot keys + counters
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const float2* tw = twiddle_fft<float>(LOG2L);
    if (!tw) return -3;
    const long groups = io.rows == 1 ? 1 : io.rows / SEQ;
    const long tiles = io.rows == 1 ? (io.nplanes + SEQ - 1) / SEQ : io.nplanes;
    long grid = (long)sm_count() * occ;
    if (grid > tiles * groups) grid = tiles * groups;
    grid -= grid % groups;
    if (grid < groups) return 1;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw);
    return check_launch("rows_bins_kernel");
}

// two-field pass 2 + radial bins (rowszx_bins_kernel).  Returns 1 when the launch shape does not fit (caller falls back).
template <int LOG2M, int ROWS>
static int launch_rowszx_bins(const RowsZCrossBins& io, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<float>::LOGE, LOG2M);
    using G_ = Geometry<LOG2M, LOGE>;
    auto kern = rowszx_bins_kernel<LOG2M, LOGE, ROWS>;
    constexpr int threads = G_::NT * 2 * ROWS;
    const int P = io.rows == 1 ? ROWS : 1;
    if (P * io.nbins > 4096 || (io.rows != 1 && io.rows % ROWS != 0)) return 1;
    constexpr size_t smem = ((size_t)2 * ROWS * (2 * G_::LPAD + 8) + (size_t)(1 << LOG2M)) * sizeof(float2)
                            + (size_t)ROWS * (2 << LOG2M) * sizeof(unsigned short) + 4096 * sizeof(int);
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const float2* tw = twiddle_fft<float>(LOG2M);
    if (!tw) return -3;
    const long groups = io.rows == 1 ? 1 : io.rows / ROWS;
    const long tiles = io.rows == 1 ? (io.nplanes + ROWS - 1) / ROWS : io.nplanes;
    long grid = (long)sm_count() * occ;
    if (grid > tiles * groups) grid = tiles * groups;
    grid -= grid % groups;
    if (grid < groups) return 1;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw);
    return check_launch("rowszx_bins_kernel");
}

// pass 2 on packed column spectra (rowsz_power_kernel): ROWS half-spectrum rows per CTA, two M-point sequences per thread
template <typename T, int LOG2M, int ROWS>
static int launch_rowsz_power(const RowsZPower<T>& io, long nseq, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<T>::LOGE, LOG2M);
    using G_ = Geometry<LOG2M, LOGE>;
    auto kern = rowsz_power_kernel<T, LOG2M, LOGE, ROWS>;
    constexpr int threads = G_::NT * ROWS;
    constexpr size_t smem = ((size_t)ROWS * (2 * G_::LPAD + 8) + (size_t)(1 << LOG2M)) * sizeof(cplx<T>);
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const cplx<T>* tw = twiddle_fft<T>(LOG2M);
    if (!tw) return -3;
    long ngroups = (nseq + ROWS - 1) / ROWS;
    long grid = (long)sm_count() * occ;
    if (grid > ngroups) grid = ngroups;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, nseq);
    return check_launch("rowsz_power_kernel");
}

// two-field pass 2 on packed column spectra (rowszx_kernel): ROWS rows per CTA, one NT-thread group per (row, field)
template <int LOG2M, int ROWS, int MODE>
static int launch_rowszx(const RowsZCross<float>& io, long nseq, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<float>::LOGE, LOG2M);
    using G_ = Geometry<LOG2M, LOGE>;
    auto kern = rowszx_kernel<LOG2M, LOGE, ROWS, MODE>;
    constexpr int threads = G_::NT * 2 * ROWS;
    constexpr size_t smem = ((size_t)2 * ROWS * (2 * G_::LPAD + 8) + (size_t)(1 << LOG2M)) * sizeof(float2);
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const float2* tw = twiddle_fft<float>(LOG2M);
    if (!tw) return -3;
    long ngroups = (nseq + ROWS - 1) / ROWS;
    long grid = (long)sm_count() * occ;
    if (grid > ngroups) grid = ngroups;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, nseq);
    return check_launch("rowszx_kernel");
}

template <typename T, int LOG2L, int C, class IO>
static int launch_cols(const IO& io, long ntiles, cudaStream_t st, size_t extra_smem = 0) {
    constexpr int LOGE = cmin(TypeCfg<T>::LOGE, LOG2L);
    constexpr int V = cmin(TypeCfg<T>::V, C);
    using G_ = Geometry<LOG2L, LOGE>;
    auto kern = [] { if constexpr (IO::kTwoFields) return cols2f_kernel<T, LOG2L, LOGE, C, IO>; else return cols_kernel<T, LOG2L, LOGE, C, V, IO>; }();
    constexpr int threads = IO::kTwoFields ? G_::NT * C : G_::NT * (C / V);
    constexpr size_t smem_fixed = (size_t)G_::LPAD * C * sizeof(cplx<T>) * (IO::kTwoFields ? 2 : 1) + IO::kExtraSmemBytes;
    // bins modes append a histogram of up to kMaxFusedBins (x2 for complex) doubles
    constexpr size_t smem_cap = smem_fixed + (IO::kBins ? (size_t)kMaxFusedBins * 2 * sizeof(double) : 0);
    const size_t smem = smem_fixed + extra_smem;
    if (smem > smem_cap) { set_error("launch_cols: too many bins for the fused path"); return -2; }
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem_cap, &occ_, &occ)) return rc;
    const cplx<T>* tw = twiddle_fft<T>(LOG2L);
    if (!tw) return -3;
    long grid = (long)sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, ntiles);
    return check_launch("cols_kernel");
}

// column pass + radial-bin epilogue with the bins of each thread's cells held in registers (cols_bins_kernel): float32,
// symmetric LUT, nbins <= 255; the grid is a multiple of the tiles per item.  Returns 1 if the launch shape does not allow
// it (fewer resident CTAs than tiles per item): the caller takes the generic kernel.
template <int LOG2L, int C>
static int launch_cols_bins(ColsFused<float, EPI_BINS_POWER> io, long ntiles, cudaStream_t st) {
    constexpr int LOGE = cmin(TypeCfg<float>::LOGE, LOG2L);
    using G_ = Geometry<LOG2L, LOGE>;
    auto kern = cols_bins_kernel<LOG2L, LOGE, C>;
    constexpr int threads = G_::NT * (C / 2);
    constexpr size_t smem_fixed = (size_t)G_::LPAD * C * sizeof(float2);
    constexpr size_t smem = smem_fixed + 256 * sizeof(float);
    io.hist_off = (int)smem_fixed;
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem, &occ_, &occ)) return rc;
    const float2* tw = twiddle_fft<float>(LOG2L);
    if (!tw) return -3;
    long grid = (long)sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    grid -= grid % io.ntile;
    if (grid < io.ntile) return 1;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, ntiles);
    return check_launch("cols_bins_kernel");
}

// asynchronous (bulk-copy fed) column pass: landing buffer + half-size exchange buffer + mbarrier (+ histogram)
template <typename T, int LOG2L, int C, class IO>
static int launch_cols_async(IO io, long ntiles, cudaStream_t st, size_t extra_smem = 0) {
    constexpr int LOGE = cmin(TypeCfg<T>::LOGE, LOG2L);
    using G_ = Geometry<LOG2L, LOGE>;
    auto kern = cols_async_kernel<T, LOG2L, LOGE, C, IO>;
    constexpr int threads = G_::NT * (C / 2);
    constexpr size_t smem_fixed = (size_t)G_::LPAD * (C + C / 2) * sizeof(cplx<T>) + 16 + IO::kExtraSmemBytes;
    constexpr size_t smem_cap = smem_fixed + (IO::kBins ? (size_t)kMaxFusedBins * 2 * sizeof(double) : 0);
    const size_t smem = smem_fixed + extra_smem;
    if (smem > smem_cap) { set_error("launch_cols_async: too many bins for the fused path"); return -2; }
    if constexpr (IO::kBins) io.hist_off = (int)((size_t)G_::LPAD * (C / 2) * sizeof(cplx<T>) + 16);   // relative to the X buffer
    static DevOcc occ_;
    int occ = 0;
    if (int rc = prepare_kernel(kern, threads, smem_cap, &occ_, &occ)) return rc;
    const cplx<T>* tw = twiddle_fft<T>(LOG2L);
    if (!tw) return -3;
    long grid = (long)sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(io, tw, ntiles);
    return check_launch("cols_async_kernel");
}

}  // namespace xrftb
