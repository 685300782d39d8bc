// xrft_b200 -- C-ABI entry points, twiddle cache, and the bandwidth-bound elementwise kernels
// (moments reduce, detrend+window, generic spectral epilogue, radial-bin sum).
// See include/xrft_b200.h for the reference seams each entry point replaces.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <vector>
#include "../../include/xrft_b200.h"
#include "internal.h"
#include <cudaTypedefs.h>

namespace xrftb {

// ------------------------------------------------------------------------------------------------
// errors / device
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long> g_launches{0};
int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s launch failed: %s", what, cudaGetErrorString(e)); return XRFTB_ECUDA; }
    return 0;
}
// ------------------------------------------------------------------------------------------------
// chain-selection options: process-wide, set explicitly through xrftb_set_option (the environment variable XRFTB_<NAME> in
// upper case only provides the value an option starts with)
// ------------------------------------------------------------------------------------------------
struct OptionDef { const char* name; int def; const char* env; };
static const OptionDef kOptions[OPT_COUNT] = {
    {"cols_first", 1, "XRFTB_COLS_FIRST"},   // 1: columns-first chain for the full-width power spectrum / radial bins; 0: rows first
    {"zpack", 1, "XRFTB_ZPACK"},             // 1: pass 1 leaves packed column spectra (z mode) where pass 2 supports it
    {"ztma", 1, "XRFTB_ZTMA"},               // 1: TMA tensor stores of the packed column spectra (rows >= 32 bytes)
    {"cols_async", 2, "XRFTB_COLS_ASYNC"},   // column kernels fed by TMA: 0 never, 1 always, 2 where measured faster
    {"rowline", 1, "XRFTB_ROWLINE"},         // 1: per-line exact detrend + rank-2 completion (float32); 0: moments pass + plane subtract
    {"cross_z", 1, "XRFTB_CROSS_Z"},         // 1: two-field z-mode chain for cross spectrum / phase; 0: rows-first two-field chain
    {"bins_static", 1, "XRFTB_BINS_STATIC"}, // 1: radial-bin kernels with a static cell-to-bin mapping; 0: generic LUT epilogue
};
static std::atomic<int> g_opt[OPT_COUNT];
static std::once_flag g_opt_once;
static void options_init() {
    for (int i = 0; i < OPT_COUNT; ++i) {
        const char* e = getenv(kOptions[i].env);
        g_opt[i].store(e ? atoi(e) : kOptions[i].def);
    }
}
int option(Option o) {
    std::call_once(g_opt_once, options_init);
    return g_opt[o].load(std::memory_order_relaxed);
}
static int option_index(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, kOptions[i].name) == 0) return i;
    return -1;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 1;
}

// ------------------------------------------------------------------------------------------------
// optional per-kernel-class timing of the fused path (CUDA events on the launching stream); used by
// bench.py to report the dominant kernel's achieved bandwidth.  Off by default (zero overhead).
// ------------------------------------------------------------------------------------------------
enum { PROF_MOMENTS = 0, PROF_ROWS = 1, PROF_COLS = 2, PROF_MIRROR = 3, PROF_NKIND = 4 };
struct ProfRec { cudaEvent_t a, b; int kind; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
struct ProfScope {
    ProfRec r{}; bool on; cudaStream_t st;
    ProfScope(int kind, cudaStream_t s) : on(g_prof_on), st(s) {
        if (on) { r.kind = kind; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, st); }
    }
    ~ProfScope() { if (on) { cudaEventRecord(r.b, st); g_prof.push_back(r); } }
};

// ------------------------------------------------------------------------------------------------
// twiddle cache
// ------------------------------------------------------------------------------------------------
static std::mutex g_tw_mu;
static std::map<std::tuple<int, int, int, int>, void*> g_tw;  // (device, kind, dtype, log2) -> device ptr

template <typename T> static const cplx<T>* get_table(int kind, int log2n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return nullptr; }
    const int dt = sizeof(T) == 4 ? 0 : 1;
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto key = std::make_tuple(dev, kind, dt, log2n);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) return reinterpret_cast<const cplx<T>*>(it->second);
    const long n = 1L << log2n;
    const long count = kind == 0 ? n : kind == 1 ? n / 2 + 1 : (n < 8192 ? n : 8192);
    std::vector<cplx<T>> h(count);
    for (long m = 0; m < count; ++m) {
        // exact octant symmetries are not needed at 1e-6 / 1e-3 tolerances; double sincos is exact enough
        double a = -2.0 * M_PI * (double)m / (double)n;
        h[m].x = (T)cos(a);
        h[m].y = (T)sin(a);
    }
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, count * sizeof(cplx<T>));
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), count * sizeof(cplx<T>), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("twiddle table upload failed: %s", cudaGetErrorString(e)); return nullptr; }
    g_tw[key] = d;
    return reinterpret_cast<const cplx<T>*>(d);
}
template <> const cplx<float>* twiddle_fft<float>(int l) { return get_table<float>(0, l); }
template <> const cplx<double>* twiddle_fft<double>(int l) { return get_table<double>(0, l); }
template <> const cplx<float>* twiddle_r2c<float>(int l) { return get_table<float>(1, l); }
template <> const cplx<double>* twiddle_r2c<double>(int l) { return get_table<double>(1, l); }
template <> const cplx<float>* twiddle_head<float>(int l) { return get_table<float>(2, l); }
template <> const cplx<double>* twiddle_head<double>(int l) { return get_table<double>(2, l); }

// ------------------------------------------------------------------------------------------------
// K1: moments.  item = [n0][n1][n2]; rows = n0*n1 of length n2.  fp64 accumulation.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512) moments_kernel(const T* __restrict__ in, double* __restrict__ mom, long n0, long n1,
                                                      long n2, int chunks, long batch) {
    const long rows = n0 * n1;
    const long r0 = rows * blockIdx.x / chunks, r1 = rows * (blockIdx.x + 1) / chunks;
    const double c0m = 0.5 * (double)(n0 - 1), c1m = 0.5 * (double)(n1 - 1), c2m = 0.5 * (double)(n2 - 1);
    __shared__ double red[4][16];
    for (long b = blockIdx.y; b < batch; b += gridDim.y) {   // gridDim.y <= 65535: items beyond it are strided over
    const T* base = in + b * rows * n2;
    double S = 0, S0 = 0, S1 = 0, S2 = 0;
    const bool vec = (sizeof(T) == 4) && (n2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
    if (vec && n2 / 4 < 2 * (long)blockDim.x) {
        // short rows: flat float4 index over the CTA's row range so every thread has work
        const long n4 = n2 / 4;
        const float4* p4 = reinterpret_cast<const float4*>(base + r0 * n2);
        const long tot4 = (r1 - r0) * n4;
        const bool p2 = (n4 & (n4 - 1)) == 0;
        int sh = 0;
        while ((1L << sh) < n4) ++sh;
        double s = 0, sx = 0, sy0 = 0, sy1 = 0;
        for (long i = threadIdx.x; i < tot4; i += blockDim.x) {
            float4 x = __ldg(p4 + i);
            const long rr = p2 ? (i >> sh) : (i / n4);
            const long c4 = i - rr * n4;
            const float c = (float)((double)(4 * c4) - c2m);
            const double s4 = (double)((x.x + x.y) + (x.z + x.w));
            const long r = r0 + rr;
            const long i0 = r / n1, i1 = r - i0 * n1;
            s += s4;
            sx += (double)((c * x.x + (c + 1.f) * x.y) + ((c + 2.f) * x.z + (c + 3.f) * x.w));
            sy0 += ((double)i0 - c0m) * s4;
            sy1 += ((double)i1 - c1m) * s4;
        }
        S = s; S0 = sy0; S1 = sy1; S2 = sx;
    } else if (vec) {
        // thread-private column position is fixed when blockDim divides the row: weights hoisted
        const long n4 = n2 / 4;
        for (long r = r0; r < r1; ++r) {
            const float4* p4 = reinterpret_cast<const float4*>(base + r * n2);
            float s4 = 0.f, w4 = 0.f;  // fp32 partials over <= 4*ceil(n4/blockDim) values, folded into fp64 per row
            double s = 0, sx = 0;
            long i = threadIdx.x;
            for (; i + 3 * blockDim.x < n4; i += 4 * blockDim.x) {
                float4 x0 = __ldg(p4 + i), x1 = __ldg(p4 + i + blockDim.x), x2 = __ldg(p4 + i + 2 * blockDim.x), x3 = __ldg(p4 + i + 3 * blockDim.x);
                const float c = (float)((double)(4 * i) - c2m), dc = (float)(4 * blockDim.x);
                s4 = ((x0.x + x0.y) + (x0.z + x0.w)) + ((x1.x + x1.y) + (x1.z + x1.w)) + ((x2.x + x2.y) + (x2.z + x2.w)) + ((x3.x + x3.y) + (x3.z + x3.w));
                w4 = (c * x0.x + (c + 1.f) * x0.y) + ((c + 2.f) * x0.z + (c + 3.f) * x0.w);
                w4 += ((c + dc) * x1.x + (c + dc + 1.f) * x1.y) + ((c + dc + 2.f) * x1.z + (c + dc + 3.f) * x1.w);
                w4 += ((c + 2 * dc) * x2.x + (c + 2 * dc + 1.f) * x2.y) + ((c + 2 * dc + 2.f) * x2.z + (c + 2 * dc + 3.f) * x2.w);
                w4 += ((c + 3 * dc) * x3.x + (c + 3 * dc + 1.f) * x3.y) + ((c + 3 * dc + 2.f) * x3.z + (c + 3 * dc + 3.f) * x3.w);
                s += (double)s4;
                sx += (double)w4;
            }
            for (; i < n4; i += blockDim.x) {
                float4 x = __ldg(p4 + i);
                const float c = (float)((double)(4 * i) - c2m);
                s += (double)((x.x + x.y) + (x.z + x.w));
                sx += (double)((c * x.x + (c + 1.f) * x.y) + ((c + 2.f) * x.z + (c + 3.f) * x.w));
            }
            const long i0 = r / n1, i1 = r - i0 * n1;
            S += s;
            S0 += ((double)i0 - c0m) * s;
            S1 += ((double)i1 - c1m) * s;
            S2 += sx;
        }
    } else {
        for (long r = r0; r < r1; ++r) {
            const T* p = base + r * n2;
            double s = 0, sx = 0;
            for (long i = threadIdx.x; i < n2; i += blockDim.x) {
                double x = (double)p[i];
                s += x;
                sx += ((double)i - c2m) * x;
            }
            const long i0 = r / n1, i1 = r - i0 * n1;
            S += s;
            S0 += ((double)i0 - c0m) * s;
            S1 += ((double)i1 - c1m) * s;
            S2 += sx;
        }
    }
    double vals[4] = {S, S0, S1, S2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double x = vals[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double x = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += red[threadIdx.x][w];
        atomicAdd(mom + b * 4 + threadIdx.x, x);
    }
    __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// row-line detrend completion (fused float32 path).  The row pass subtracted from row i of item b the exactly known
// line ph_i(j) = A0 + B j and recorded (A0, B, sum_j r, sum_j (j - jc) r) of the residual r = x - ph.
//   row sums of x        S_i = Nx A0 + B Nx (Nx-1)/2 + sum r ;   T_i = sum_j (j - jc) x = B vx + sum (j - jc) r
//   least-squares plane  a = sum S_i / (Ny Nx), b = sum (i - ic) S_i / (Nx vy), c = sum T_i / (Ny vx)   (detrend=linear;
//                        b = c = 0 for detrend=constant) -- the same centred-moment closed form as moments_kernel
//   still to subtract    plane - ph_i = -(alpha_i + gamma_i (j - jc)),  alpha_i = A0 + B jc - a - b (i - ic), gamma_i = B - c
// Output: ag[b][i] = w_y(i) (alpha_i, gamma_i); the column pass adds ag.x What(kx) + ag.y Jhat(kx) to row i of the
// half-spectrum (What, Jhat = transforms of w_x(j) and w_x(j)(j - jc)).  Grid (items, splits): every CTA redoes the
// small reduction over the item's lines (L2-resident) and completes its share of them; fp64 throughout.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rowline_fix_kernel(const float4* __restrict__ rowstats, cplx<T>* __restrict__ ag, const T* __restrict__ wy,
                                                          int ny, int nx, int detrend) {
    const long b = blockIdx.x;
    const float4* rs = rowstats + b * ny;
    const double ic = 0.5 * (double)(ny - 1), jc = 0.5 * (double)(nx - 1);
    const double vx = (double)nx * ((double)nx * nx - 1.0) / 12.0, vy = (double)ny * ((double)ny * ny - 1.0) / 12.0;
    const double tri = 0.5 * (double)nx * (double)(nx - 1);
    double S = 0, Sy = 0, Sx = 0;
    for (int i = threadIdx.x; i < ny; i += blockDim.x) {
        const float4 r = rs[i];
        const double Si = (double)nx * (double)r.x + (double)r.y * tri + (double)r.z;
        S += Si;
        Sy += ((double)i - ic) * Si;
        Sx += (double)r.y * vx + (double)r.w;
    }
    __shared__ double red[3][8];
    __shared__ double plane[3];
    double vals[3] = {S, Sy, Sx};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double x = vals[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double x = 0;
        for (int w = 0; w < 8; ++w) x += red[threadIdx.x][w];
        plane[threadIdx.x] = x;
    }
    __syncthreads();
    const double a = plane[0] / ((double)ny * (double)nx);
    const double bb = (detrend == 2 && ny > 1) ? plane[1] / ((double)nx * vy) : 0.0;
    const double c = (detrend == 2) ? plane[2] / ((double)ny * vx) : 0.0;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < ny; i += gridDim.y * blockDim.x) {
        const float4 r = rs[i];
        const double w = wy ? (double)wy[i] : 1.0;
        const double alpha = (double)r.x + (double)r.y * jc - a - bb * ((double)i - ic);
        const double gamma = (double)r.y - c;
        ag[b * ny + i] = mk<T>((T)(w * alpha), (T)(w * gamma));
    }
}
// rows [w_x(j)] and [w_x(j) (j - jc)] in double, input of the one-off transform that yields What, Jhat
template <typename T>
__global__ void __launch_bounds__(256) rowline_wj_in_kernel(const T* __restrict__ wx, double* __restrict__ rows, int nx) {
    const double jc = 0.5 * (double)(nx - 1);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nx; j += gridDim.x * blockDim.x) {
        const double w = wx ? (double)wx[j] : 1.0;
        rows[j] = w;
        rows[nx + j] = w * ((double)j - jc);
    }
}
// wj[2 k] = What(k), wj[2 k + 1] = Jhat(k) for k <= M; zero for the padding columns of the last tile
template <typename T>
__global__ void __launch_bounds__(256) rowline_wj_out_kernel(const double2* __restrict__ spec, cplx<T>* __restrict__ wj, int M, int ncols) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ncols; k += gridDim.x * blockDim.x) {
        cplx<T> w = mk<T>(0, 0), j = mk<T>(0, 0);
        if (k <= M) {
            const double2 a = spec[k], b = spec[(M + 1) + k];
            w = mk<T>((T)a.x, (T)a.y);
            j = mk<T>((T)b.x, (T)b.y);
        }
        wj[2 * k] = w;
        wj[2 * k + 1] = j;
    }
}

// ------------------------------------------------------------------------------------------------
// Index of a grid-stride loop over a [batch][n0][n1][n2] array, decomposed ONCE per thread (four 64-bit divisions) and advanced
// by the decomposed stride with carries: the elementwise kernels below spent most of their instructions on a 64-bit
// division chain per element (1 TB/s on 4-byte elements; profiles/r02_pass1_source_profile.md).
// ------------------------------------------------------------------------------------------------
struct GridIndex {
    long b; int i0, i1, i2;
    long sb; int s0, s1, s2;
    int n0, n1, n2;
    __device__ __forceinline__ GridIndex(long start, long stride, long n0_, long n1_, long n2_) : n0((int)n0_), n1((int)n1_), n2((int)n2_) {
        i2 = (int)(start % n2_); long r = start / n2_; i1 = (int)(r % n1_); r /= n1_; i0 = (int)(r % n0_); b = r / n0_;
        s2 = (int)(stride % n2_); r = stride / n2_; s1 = (int)(r % n1_); r /= n1_; s0 = (int)(r % n0_); sb = r / n0_;
    }
    __device__ __forceinline__ void next() {
        int c;
        i2 += s2; c = i2 >= n2; i2 -= c ? n2 : 0;
        i1 += s1 + c; c = i1 >= n1; i1 -= c ? n1 : 0;
        i0 += s0 + c; c = i0 >= n0; i0 -= c ? n0 : 0;
        b += sb + c;
    }
};
// (o + add) mod n for 0 <= o, add < n
__device__ __forceinline__ int wrap_add(int o, int add, int n) { const int x = o + add; return x >= n ? x - n : x; }

// ------------------------------------------------------------------------------------------------
// detrend + window (non-fused path and the public xrft.detrend)
// ------------------------------------------------------------------------------------------------
// VEC consecutive elements of the last axis per thread and trip (VEC = 4 needs n2 % 4 == 0 and 16-byte aligned rows): four
// independent loads in flight per thread -- with one 4-byte load per thread the kernel could not keep HBM busy -- and vector stores
__device__ __forceinline__ void load_vec4(const float* p, float (&v)[4]) { const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
__device__ __forceinline__ void load_vec4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store_vec4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store_vec4(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
template <typename T, int VEC>
__global__ void __launch_bounds__(256) detrend_window_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                             const double* __restrict__ mom, int detrend, const T* w0,
                                                             const T* w1, const T* w2, long n0, long n1, long n2, long total) {
    const double npts = (double)n0 * (double)n1 * (double)n2;
    // plane = m0 / npts + sum_a m_a c_a (i_a - (n_a - 1) / 2): the reciprocals are per launch, the products per element
    const double inv_npts = 1.0 / npts;
    const double c0 = n0 > 1 ? 1.0 / (npts * ((double)n0 * n0 - 1.0) / 12.0) : 0.0;
    const double c1 = n1 > 1 ? 1.0 / (npts * ((double)n1 * n1 - 1.0) / 12.0) : 0.0;
    const double c2 = n2 > 1 ? 1.0 / (npts * ((double)n2 * n2 - 1.0) / 12.0) : 0.0;
    const double h0 = 0.5 * (double)(n0 - 1), h1 = 0.5 * (double)(n1 - 1), h2 = 0.5 * (double)(n2 - 1);
    const long stride = (long)gridDim.x * blockDim.x;
    const long ngroups = total / VEC;
    long iv = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (iv >= ngroups) return;
    GridIndex g(iv, stride, n0, n1, n2 / VEC);
    for (; iv < ngroups; iv += stride, g.next()) {
        const long e = iv * VEC;
        T v[VEC];
        if constexpr (VEC == 4) load_vec4(in + e, v); else v[0] = in[e];
        double prow = 0.0, pcol = 0.0;
        if (detrend) {
            const double* m = mom + g.b * 4;
            prow = m[0] * inv_npts;
            if (detrend == 2) { prow += m[1] * c0 * ((double)g.i0 - h0) + m[2] * c1 * ((double)g.i1 - h1); pcol = m[3] * c2; }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int i2 = g.i2 * VEC + k;
            double x = (double)v[k];
            if (detrend) x -= prow + pcol * ((double)i2 - h2);
            T y = (T)x;  // the reference rounds the detrended field to the input dtype (output_dtypes=[da.dtype])
            if (w0) y *= w0[g.i0];
            if (w1) y *= w1[g.i1];
            if (w2) y *= w2[i2];
            v[k] = y;
        }
        if constexpr (VEC == 4) store_vec4(out + e, v); else out[e] = v[0];
    }
}

// ------------------------------------------------------------------------------------------------
// generic spectral epilogue (non-fused path)
// ------------------------------------------------------------------------------------------------
struct PostDesc {
    int mode;
    long k0, k1, k2, k2in, W;
    int hermitian;
    int shift[3];
    const void* ramp[3];
    const void* weight;
    double scale;
    long seg_n, seg_inner;   // seg_n > 0: mean over an axis of the batch (Welch segments): batch = [outer][seg_n][seg_inner]
};

// VEC = 4 (W % 4 == 0, aligned rows, no segment mean): four outputs of a row per thread and trip -- four independent source
// loads in flight and one vector store
template <typename T, int VEC>
__global__ void __launch_bounds__(256) spectral_post_kernel(const cplx<T>* __restrict__ in1, const cplx<T>* __restrict__ in2,
                                                            void* __restrict__ out, PostDesc d, long total) {
    const long stride = (long)gridDim.x * blockDim.x;
    const long ngroups = total / VEC;
    long iv = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (iv >= ngroups) return;
    const int k0 = (int)d.k0, k1 = (int)d.k1, k2 = (int)d.k2;
    // o = (f + N/2) % N (fftshift) <=> f = (o + N - N/2) % N ; ifftshift: o = (f + N - N/2) % N <=> f = (o + N/2) % N
    const int a0 = d.shift[0] == 1 ? k0 - k0 / 2 : d.shift[0] == 2 ? k0 / 2 : 0;
    const int a1 = d.shift[1] == 1 ? k1 - k1 / 2 : d.shift[1] == 2 ? k1 / 2 : 0;
    const int a2 = d.shift[2] == 1 ? k2 - k2 / 2 : d.shift[2] == 2 ? k2 / 2 : 0;
    const bool two = !(d.mode == EPI_COMPLEX || d.mode == EPI_POWER);
    GridIndex gi(iv, stride, d.k0, d.k1, d.W / VEC);
    for (; iv < ngroups; iv += stride, gi.next()) {
        const int o0 = gi.i0, o1 = gi.i1;
        const long b = gi.b;
        const int f0 = wrap_add(o0, a0 == k0 ? 0 : a0, k0), f1 = wrap_add(o1, a1 == k1 ? 0 : a1, k1);
        const long rowd = (b * d.k0 + f0) * d.k1 + f1;                                // direct source row
        const long rowm = (b * d.k0 + (f0 ? k0 - f0 : 0)) * d.k1 + (f1 ? k1 - f1 : 0);   // its Hermitian partner
        int f2[VEC];
        bool cj[VEC];
        cplx<T> a[VEC], g[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int o2 = gi.i2 * VEC + k;
            f2[k] = d.shift[2] ? wrap_add(o2, a2 == k2 ? 0 : a2, k2) : o2;   // (without a shift o2 runs over W <= k2 columns)
            cj[k] = d.hermitian && f2[k] > k2 / 2;
            const long src = cj[k] ? rowm * d.k2in + (k2 - f2[k]) : rowd * d.k2in + f2[k];
            a[k] = in1[src];
            if (two) g[k] = in2[src];
        }
        const long i = iv * VEC;
        T re[VEC], im[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            cplx<T> x = a[k];
            if (cj[k]) x.y = -x.y;
            cplx<T> val;
            if (d.mode == EPI_COMPLEX) val = x;
            else if (d.mode == EPI_POWER) val = mk<T>(x.x * x.x + x.y * x.y, 0);
            else {
                cplx<T> y = g[k];
                if (cj[k]) y.y = -y.y;
                val = cmulc(x, y);
            }
            if (d.ramp[0]) val = cmul(val, reinterpret_cast<const cplx<T>*>(d.ramp[0])[f0]);
            if (d.ramp[1]) val = cmul(val, reinterpret_cast<const cplx<T>*>(d.ramp[1])[f1]);
            if (d.ramp[2]) val = cmul(val, reinterpret_cast<const cplx<T>*>(d.ramp[2])[f2[k]]);
            T sc = (T)d.scale;
            if (d.weight) sc *= reinterpret_cast<const T*>(d.weight)[f2[k]];
            val = cscale(val, sc);
            if constexpr (VEC == 1) {
                if (d.seg_n > 0) {
                    // segment mean as an epilogue reduction: the per-segment spectra are never written; every value is added
                    // (already divided by the number of segments) to the cell of the reduced array
                    const long span = d.seg_n * d.seg_inner;
                    const long outer = b / span;
                    const long ob = outer * d.seg_inner + (b - outer * span) % d.seg_inner;
                    const long oi = ((ob * d.k0 + o0) * d.k1 + o1) * d.W + gi.i2;
                    if (d.mode == EPI_POWER) atomicAdd(reinterpret_cast<T*>(out) + oi, val.x);
                    else { atomicAdd(reinterpret_cast<T*>(out) + 2 * oi, val.x); atomicAdd(reinterpret_cast<T*>(out) + 2 * oi + 1, val.y); }
                    continue;
                }
            }
            re[k] = val.x; im[k] = val.y;
            if (d.mode == EPI_PHASE) re[k] = xatan2(val.y, val.x);
        }
        if constexpr (VEC == 1) {
            if (d.seg_n > 0) continue;
            if (d.mode == EPI_COMPLEX || d.mode == EPI_CROSS) reinterpret_cast<cplx<T>*>(out)[i] = mk<T>(re[0], im[0]);
            else reinterpret_cast<T*>(out)[i] = re[0];
        } else {
            if (d.mode == EPI_COMPLEX || d.mode == EPI_CROSS) {
                T lo[4] = {re[0], im[0], re[1], im[1]}, hi[4] = {re[2], im[2], re[3], im[3]};
                store_vec4(reinterpret_cast<T*>(out) + 2 * i, lo);
                store_vec4(reinterpret_cast<T*>(out) + 2 * i + 4, hi);
            } else {
                store_vec4(reinterpret_cast<T*>(out) + i, re);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// circular roll + scale over up to 3 trailing axes: out[(i + s) % n] = in[i] * scale (np.fft.fftshift /
// ifftshift of xrft.py:617-621 and the 1/prod(spacing) of xrft.py:641-642 on the inverse path)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) roll_scale_kernel(const T* __restrict__ in, T* __restrict__ out, long n0, long n1, long n2,
                                                         long s0, long s1, long s2, int width, T scale, long total) {
    const long stride = (long)gridDim.x * blockDim.x;
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r0 = (int)(s0 % n0), r1 = (int)(s1 % n1), r2 = (int)(s2 % n2);   // shifts reduced once: 0 <= r < n
    GridIndex g(i, stride, n0, n1, n2);
    for (; i < total; i += stride, g.next()) {
        const long o = ((g.b * n0 + wrap_add(g.i0, r0, (int)n0)) * n1 + wrap_add(g.i1, r1, (int)n1)) * n2 + wrap_add(g.i2, r2, (int)n2);
        for (int w = 0; w < width; ++w) out[o * width + w] = in[i * width + w] * scale;
    }
}

// ------------------------------------------------------------------------------------------------
// Hermitian completion of a full-width spectrum whose direct half (unshifted kx <= Nx/2) is written:
//   out[-ky][-kx] = conj(out[ky][kx]) (* ramp correction for unit-modulus ramps), rows fully coalesced.
// ------------------------------------------------------------------------------------------------
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256) mirror_fill_kernel(void* __restrict__ out_, int logNy, int logNx, int shift_y, int shift_x,
                                                          const cplx<T>* __restrict__ ramp_y, const cplx<T>* __restrict__ ramp_x,
                                                          long nrows_total, int negate) {
    using E = typename std::conditional<CPLX, cplx<T>, T>::type;
    const T sgn = negate ? (T)-1 : (T)1;   // cross-phase: angle(conj z) = -angle(z)
    E* out = reinterpret_cast<E*>(out_);
    const int Ny = 1 << logNy, Nx = 1 << logNx, M = Nx >> 1;
    const int sy = shift_y ? Ny / 2 : 0, sx = shift_x ? M : 0;
    // target cells: unshifted kx' in (M, Nx)  <=>  ox' = (kx' + sx) & (Nx-1); there are M-1 of them per row
    const int per_row = M - 1;
    for (long row = blockIdx.x; row < nrows_total; row += gridDim.x) {
        const long b = row >> logNy;
        const int oy = (int)(row & (Ny - 1));
        const int kyt = (oy - sy) & (Ny - 1);             // target unshifted row index
        const int kys = (Ny - kyt) & (Ny - 1);            // source (direct) row
        const int oys = (kys + sy) & (Ny - 1);
        const E* src = out + ((b << logNy) | oys) * (long)Nx;
        E* dst = out + ((b << logNy) | oy) * (long)Nx;
        if constexpr (!CPLX && sizeof(T) == 4) {
            if (Nx >= 64) {
                // target column c (in the half being filled) takes source column Nx - c of the other half (both in the
                // SHIFTED or both in the unshifted frame: (kxs + sx) + (kxt + sx) == Nx (mod Nx)).
                // Aligned 16 B stores of targets [4g, 4g+3] need sources [Nx-4g-3, Nx-4g]: one aligned float4
                // [Nx-4g-4, Nx-4g-1] plus the first element of the previous (higher) float4 -> one warp shuffle.
                const int t0 = shift_x ? 0 : M;            // first column of the target half (column t0 itself is direct)
                const float* srcf = reinterpret_cast<const float*>(src);
                float* dstf = reinterpret_cast<float*>(dst);
                const int ngroups = M / 4;
                for (int g0 = 0; g0 < ngroups; g0 += blockDim.x) {
                    const int g = g0 + threadIdx.x;
                    const bool act = g < ngroups;
                    const int c = t0 + 4 * g;              // target columns c .. c+3
                    const int s_hi = (Nx - c) & (Nx - 1);  // source of target c (wraps to 0 when c == 0)
                    float4 lo = make_float4(0, 0, 0, 0);
                    if (act) lo = *reinterpret_cast<const float4*>(srcf + ((Nx - c - 4) & (Nx - 1)));  // sources of c+4 .. c+1 (reversed)
                    // source of target c is element x of the float4 held by the thread handling group g-1
                    float prev = __shfl_up_sync(0xffffffffu, lo.x, 1);
                    if ((threadIdx.x & 31) == 0 && act) prev = srcf[s_hi];
                    if (act) {
                        float4 o = make_float4(sgn * prev, sgn * lo.w, sgn * lo.z, sgn * lo.y);
                        if (g == 0) {  // column t0 is a direct cell (kx = 0 or Nyquist): keep it
                            dstf[c + 1] = o.y; dstf[c + 2] = o.z; dstf[c + 3] = o.w;
                        } else {
                            *reinterpret_cast<float4*>(dstf + c) = o;
                        }
                    }
                }
                continue;
            }
        }
        if constexpr (CPLX && sizeof(T) == 4) {
            if (Nx >= 64 && !ramp_y && !ramp_x) {
                // complex64, no ramps: out[-k] = conj(out[k]); two cells (16 B) per thread, same shuffle realignment
                const int t0 = shift_x ? 0 : M;
                const float2* srcz = reinterpret_cast<const float2*>(src);
                float2* dstz = reinterpret_cast<float2*>(dst);
                const int ngroups = M / 2;
                for (int g0 = 0; g0 < ngroups; g0 += blockDim.x) {
                    const int g = g0 + threadIdx.x;
                    const bool act = g < ngroups;
                    const int c = t0 + 2 * g;              // target columns c, c+1
                    float4 lo = make_float4(0, 0, 0, 0);
                    if (act) lo = *reinterpret_cast<const float4*>(srcz + ((Nx - c - 2) & (Nx - 1)));  // sources of c+2 (x,y), c+1 (z,w)
                    float px = __shfl_up_sync(0xffffffffu, lo.x, 1), py = __shfl_up_sync(0xffffffffu, lo.y, 1);
                    if ((threadIdx.x & 31) == 0 && act) { float2 z = srcz[(Nx - c) & (Nx - 1)]; px = z.x; py = z.y; }
                    if (act) {
                        if (g == 0) dstz[c + 1] = make_float2(lo.z, -lo.w);   // column t0 is a direct cell
                        else *reinterpret_cast<float4*>(dstz + c) = make_float4(px, -py, lo.z, -lo.w);
                    }
                }
                continue;
            }
        }
        cplx<T> ry = mk<T>(1, 0);
        if constexpr (CPLX) { if (ramp_y) ry = cmul(__ldg(ramp_y + kys), __ldg(ramp_y + kyt)); }
#pragma unroll 4
        for (int j = threadIdx.x; j < per_row; j += blockDim.x) {
            const int kxt = M + 1 + j;                    // target unshifted column index
            const int kxs = Nx - kxt;
            E v = src[(kxs + sx) & (Nx - 1)];
            if constexpr (CPLX) {
                v.y = -v.y;
                if (ramp_y) v = cmul(v, ry);
                if (ramp_x) v = cmul(v, cmul(__ldg(ramp_x + kxs), __ldg(ramp_x + kxt)));
            } else {
                v = v * sgn;
            }
            dst[(kxt + sx) & (Nx - 1)] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// radial-bin sum: per-CTA fp64 histogram in shared memory, flushed with one global atomic per bin
// ------------------------------------------------------------------------------------------------
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256) binned_sum_kernel(const T* __restrict__ arr, const int* __restrict__ lut,
                                                         double* __restrict__ bins, long ncell, int nbins, int chunks, long batch) {
    extern __shared__ double hist[];
    const int width = CPLX ? 2 : 1;
    const long c0 = ncell * blockIdx.x / chunks, c1 = ncell * (blockIdx.x + 1) / chunks;
    for (long b = blockIdx.y; b < batch; b += gridDim.y) {   // gridDim.y <= 65535: items beyond it are strided over
        for (int i = threadIdx.x; i < nbins * width; i += blockDim.x) hist[i] = 0.0;
        __syncthreads();
        const T* p = arr + b * ncell * width;
        // Every thread takes RUN consecutive cells: their loads are independent (RUN in flight), neighbouring cells mostly share
        // the radial bin, so a thread adds up runs of equal bins in registers and pays one shared-memory atomic per run
        // (fp64 atomics on shared memory are compare-and-swap loops: one per cell was the cost of this kernel)
        constexpr int RUN = 8;
        for (long i0 = c0 + (long)threadIdx.x * RUN; i0 < c1; i0 += (long)blockDim.x * RUN) {
            int bins_[RUN];
            T re[RUN], im[RUN];
#pragma unroll
            for (int k = 0; k < RUN; ++k) {
                const long i = i0 + k;
                const bool ok = i < c1;
                bins_[k] = ok ? lut[i] : -1;
                re[k] = ok ? p[width * i] : (T)0;
                im[k] = (CPLX && ok) ? p[width * i + 1] : (T)0;
            }
            int cur = -1;
            double ar = 0.0, ai = 0.0;
#pragma unroll
            for (int k = 0; k < RUN; ++k) {
                if (bins_[k] != cur) {
                    if (cur >= 0) { atomicAdd(&hist[width * cur], ar); if (CPLX) atomicAdd(&hist[2 * cur + 1], ai); }
                    cur = bins_[k]; ar = 0.0; ai = 0.0;
                }
                ar += (double)re[k]; ai += (double)im[k];
            }
            if (cur >= 0) { atomicAdd(&hist[width * cur], ar); if (CPLX) atomicAdd(&hist[2 * cur + 1], ai); }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nbins * width; i += blockDim.x)
            if (hist[i] != 0.0) atomicAdd(bins + b * nbins * width + i, hist[i]);
        __syncthreads();
    }
}
// more bins than a CTA histogram holds (min(N) / nfactor > 2048): fp64 atomics straight to the item's global bins
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256) binned_sum_global_kernel(const T* __restrict__ arr, const int* __restrict__ lut,
                                                                double* __restrict__ bins, long ncell, int nbins, long batch) {
    const int width = CPLX ? 2 : 1;
    for (long b = blockIdx.y; b < batch; b += gridDim.y) {
        const T* p = arr + b * ncell * width;
        double* bb = bins + b * (long)nbins * width;
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < ncell; i += (long)gridDim.x * blockDim.x) {
            const int bin = lut[i];
            if (bin < 0) continue;
            if (CPLX) { atomicAdd(bb + 2 * bin, (double)p[2 * i]); atomicAdd(bb + 2 * bin + 1, (double)p[2 * i + 1]); }
            else atomicAdd(bb + bin, (double)p[i]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// arbitrary lengths (the reference's tests use 10, 15, 19, 20, 30, 40, 100, 1000 ...)
//   N <= kSmallDft : direct O(N^2) DFT, one thread per sequence, twiddles in shared memory (non-smooth lengths, tiny strided axes)
//   2^a 3^b 5^c 7^d: mixed-radix Stockham kernel in shared memory (smooth.cu), up to 12800 (float32) / 6400 (float64) points
//   otherwise      : Bluestein chirp-z on top of the power-of-two passes (no cuFFT, no CPU fallback)
// All operate on a [A][N][B] row-major view (FFT along the middle axis), in-place safe.
// ------------------------------------------------------------------------------------------------
constexpr int kSmallDft = 64;
// Short smooth lengths run the mixed-radix kernel too, except on a strided axis with a handful of points or columns (its tiles
// would be nearly empty there).  Measured (tools/probe_small.py): 400000 x 40 complex64 1.38 -> 0.11 ms, 200000 x 60 2.65 -> 0.10,
// 100000 x 48 complex128 0.84 -> 0.07, rfft of 400000 x 40 float32 0.70 -> 0.08; 65536 x 12 x 20 (strided 12 x 11 columns) stays direct.
template <typename T> static inline bool small_direct(long n, long B) {
    return n <= kSmallDft && !(n >= 6 && smooth_len_ok<T>(n) && (B == 1 || (n >= 16 && B >= 16)));
}
template <typename T> static inline bool small_direct_real(long N) { return N <= kSmallDft && !(N >= 12 && smooth_real_ok<T>(N)); }

// mode: 0 C2C forward, 1 C2C inverse (x scale), 2 R2C (real in, k <= N/2 out), 3 C2R (half in, real out, x scale)
template <typename T>
__global__ void __launch_bounds__(128) dft_small_kernel(const void* __restrict__ in_, void* __restrict__ out_, int N, long A, long B,
                                                        int mode, T scale) {
    __shared__ double2 tw[kSmallDft];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)i / (double)N, &s, &c);
        tw[i] = make_double2(c, s);
    }
    __syncthreads();
    const long nseq = A * B;
    const int H = N / 2 + 1;
    for (long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < nseq; q += (long)gridDim.x * blockDim.x) {
        const long a = q / B, b = q - a * B;
        double2 x[kSmallDft];
        if (mode == 2) {
            const T* p = reinterpret_cast<const T*>(in_) + a * N * B + b;
            for (int n = 0; n < N; ++n) x[n] = make_double2((double)p[n * B], 0.0);
        } else if (mode == 3) {
            const cplx<T>* p = reinterpret_cast<const cplx<T>*>(in_) + a * H * B + b;
            for (int k = 0; k < H; ++k) { cplx<T> v = p[k * B]; x[k] = make_double2((double)v.x, (double)v.y); }
            x[0].y = 0.0;
            if (N % 2 == 0) x[N / 2].y = 0.0;
            for (int k = H; k < N; ++k) x[k] = make_double2(x[N - k].x, -x[N - k].y);
        } else {
            const cplx<T>* p = reinterpret_cast<const cplx<T>*>(in_) + a * N * B + b;
            for (int n = 0; n < N; ++n) { cplx<T> v = p[n * B]; x[n] = make_double2((double)v.x, (double)v.y); }
        }
        const bool inv = (mode == 1 || mode == 3);
        const int K = (mode == 2) ? H : N;
        for (int k = 0; k < K; ++k) {
            double re = 0.0, im = 0.0;
            int idx = 0;
            for (int n = 0; n < N; ++n) {
                double2 w = tw[idx];
                if (inv) w.y = -w.y;
                re += x[n].x * w.x - x[n].y * w.y;
                im += x[n].x * w.y + x[n].y * w.x;
                idx += k;
                if (idx >= N) idx -= N;
            }
            if (mode == 3) reinterpret_cast<T*>(out_)[(a * N + k) * B + b] = (T)(re * (double)scale);
            else if (mode == 2) reinterpret_cast<cplx<T>*>(out_)[(a * H + k) * B + b] = mk<T>((T)re, (T)im);
            else reinterpret_cast<cplx<T>*>(out_)[(a * N + k) * B + b] = mk<T>((T)(re * (double)scale), (T)(im * (double)scale));
        }
    }
}

// Bluestein: X[k] = b[k] * sum_n (x[n] b[n]) conj(b[k-n]),  b[n] = exp(-i pi n^2 / N)
// pre : work[a][m][b] = (conj? x)[a][m][b] * chirp[m]   (m < N), 0 (N <= m < M);  real or complex x
// mul : work[a][m][b] *= filt[m]                        (filt = FFT_M of the wrapped conj chirp)
// post: out[a][k][b]  = (conj?)(work[a][k][b] * chirp[k]) * scale, k < N
template <typename T, bool REAL_IN>
__global__ void __launch_bounds__(256) bluestein_pre_kernel(const void* __restrict__ in_, cplx<T>* __restrict__ work, const cplx<T>* __restrict__ chirp,
                                                            long A, long N, long M, long B, int conj_in, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i % B;
        long r = i / B;
        const long m = r % M;
        const long a = r / M;
        cplx<T> v = mk<T>(0, 0);
        if (m < N) {
            if (REAL_IN) v = mk<T>(reinterpret_cast<const T*>(in_)[(a * N + m) * B + b], 0);
            else v = reinterpret_cast<const cplx<T>*>(in_)[(a * N + m) * B + b];
            if (conj_in) v.y = -v.y;
            v = cmul(v, chirp[m]);
        }
        work[i] = v;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) bluestein_mul_kernel(cplx<T>* __restrict__ work, const cplx<T>* __restrict__ filt, long M, long B, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long m = (i / B) % M;
        work[i] = cmul(work[i], filt[m]);
    }
}
template <typename T, bool REAL_OUT>
__global__ void __launch_bounds__(256) bluestein_post_kernel(const cplx<T>* __restrict__ work, void* __restrict__ out_, const cplx<T>* __restrict__ chirp,
                                                             long A, long N, long Nout, long M, long B, int conj_out, T scale, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i % B;
        long r = i / B;
        const long k = r % Nout;
        const long a = r / Nout;
        cplx<T> v = cmul(work[(a * M + k) * B + b], chirp[k]);
        if (conj_out) v.y = -v.y;
        if (REAL_OUT) reinterpret_cast<T*>(out_)[i] = v.x * scale;
        else reinterpret_cast<cplx<T>*>(out_)[i] = cscale(v, scale);
    }
}
// Hermitian extension of a half spectrum along the last-but-B axis: [A][H][B] -> [A][N][B] (for the non-pow2 C2R path)
template <typename T>
__global__ void __launch_bounds__(256) herm_extend_kernel(const cplx<T>* __restrict__ in, cplx<T>* __restrict__ out, long N, long B, long total) {
    const long H = N / 2 + 1;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i % B;
        long r = i / B;
        const long k = r % N;
        const long a = r / N;
        cplx<T> v;
        if (k < H) { v = in[(a * H + k) * B + b]; if (k == 0 || (N % 2 == 0 && k == N / 2)) v.y = 0; }
        else { v = in[(a * H + (N - k)) * B + b]; v.y = -v.y; }
        out[i] = v;
    }
}

// four-step decomposition of power-of-two lengths beyond one CTA's shared memory: n = n1 * n2,
//   x[i1*n2 + i2] --FFT over i1--> Y[k1][i2] --* w_n^(k1 i2)--> --FFT over i2--> Z[k1][k2] --transpose--> X[k1 + n1 k2]
// (twiddle and transposition are store hooks of ColsC2C; the kernel below serves the contiguous case B == 1)
// out[a][k2][k1][b] = in[a][k1][k2][b] * scale
template <typename T>
__global__ void __launch_bounds__(256) fourstep_transpose_kernel(const cplx<T>* __restrict__ in, cplx<T>* __restrict__ out, long n1, long n2, long B,
                                                                 T scale, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i % B;
        long r = i / B;
        const long k1 = r % n1;
        r /= n1;
        const long k2 = r % n2;
        const long a = r / n2;
        out[i] = cscale(in[((a * n1 + k1) * n2 + k2) * B + b], scale);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) real_to_cplx_kernel(const T* __restrict__ in, cplx<T>* __restrict__ out, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) out[i] = mk<T>(in[i], 0);
}
// out[s][k] = in[s][k] for k < H (complex crop) or out[s][n] = Re in[s][n] (REAL_OUT, H == N)
template <typename T, bool REAL_OUT>
__global__ void __launch_bounds__(256) crop_kernel(const cplx<T>* __restrict__ in, void* __restrict__ out, long N, long H, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long k = i % H, sq = i / H;
        cplx<T> v = in[sq * N + k];
        if (REAL_OUT) reinterpret_cast<T*>(out)[i] = v.x;
        else reinterpret_cast<cplx<T>*>(out)[i] = v;
    }
}


// ------------------------------------------------------------------------------------------------
// pad (xrft/padding.py:157-181 -> xarray.DataArray.pad -> numpy.pad): out = in surrounded by pad_before / pad_after cells per
// axis, the new cells filled by a constant or by an index map of the input (edge / reflect / symmetric / wrap, numpy
// semantics incl. pads wider than the array).  Pure data movement: one gather per output element, any element size.
// ------------------------------------------------------------------------------------------------
struct PadDesc { long in_n[4], out_n[4], before[4]; int mode; };
__device__ __forceinline__ long pad_src_index(long j, long n, int mode) {   // j = output index - pad_before, may be outside [0, n)
    if (j >= 0 && j < n) return j;
    if (mode == 1) return j < 0 ? 0 : n - 1;                       // edge
    if (n == 1) return 0;
    if (mode == 4) { long r = j % n; return r < 0 ? r + n : r; }    // wrap
    const long p = mode == 2 ? 2 * (n - 1) : 2 * n;                 // reflect (edge not repeated) / symmetric (edge repeated)
    long r = j % p;
    if (r < 0) r += p;
    return r < n ? r : (mode == 2 ? p - r : p - 1 - r);
}
template <typename E>
__global__ void __launch_bounds__(256) pad_kernel(const E* __restrict__ in, E* __restrict__ out, PadDesc d, E fill, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i, src = 0, mul = 1;
        bool inside = true;
        long idx[4];
#pragma unroll
        for (int a = 3; a >= 0; --a) { idx[a] = r % d.out_n[a]; r /= d.out_n[a]; }
#pragma unroll
        for (int a = 3; a >= 0; --a) {
            const long j = idx[a] - d.before[a];
            long sidx = j;
            if (j < 0 || j >= d.in_n[a]) {
                if (d.mode == 0) inside = false; else sidx = pad_src_index(j, d.in_n[a], d.mode);
            }
            src += sidx * mul;
            mul *= d.in_n[a];
        }
        out[i] = inside ? in[src] : fill;
    }
}

// ------------------------------------------------------------------------------------------------
// axis permutation + per-axis reversal (the transpose that moves the transform axes last, xrft.py:386-396, and the flip of
// decreasing coordinates under true_phase, xrft.py:436-441): out[o0..] = in[perm / flip of the indices], any element size
// ------------------------------------------------------------------------------------------------
struct PermDesc { long out_n[6]; long in_stride[6]; long in_off; int ndim; };   // in_stride: of the input axis feeding output axis a (negative when flipped)
template <typename E>
__global__ void __launch_bounds__(256) permute_kernel(const E* __restrict__ in, E* __restrict__ out, PermDesc d, long total) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i, src = d.in_off;
#pragma unroll
        for (int a = 5; a >= 0; --a) {
            if (a < d.ndim) { const long idx = r % d.out_n[a]; r /= d.out_n[a]; src += idx * d.in_stride[a]; }
        }
        out[i] = in[src];
    }
}

// Permutations that collapse to a batched 2-D transpose out[b][q][p] = in[b][p][q] (the usual case: ONE block of axes moved
// behind the others, e.g. dim="time" of a [time][y][x] array): 32 x 32 tiles through shared memory, reads along q and writes
// along p both coalesced -- the generic gather above reads with a stride of Q elements.
template <typename E>
__global__ void __launch_bounds__(256) transpose_kernel(const E* __restrict__ in, E* __restrict__ out, long P, long Q, long tiles_p, long tiles_q, long ntiles) {
    __shared__ E tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long tq = t % tiles_q, r = t / tiles_q;
        const long tp = r % tiles_p, b = r / tiles_p;
        const E* src = in + b * P * Q;
        E* dst = out + b * P * Q;
        const long q = tq * 32 + tx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long p_ = tp * 32 + ty + 8 * k;
            if (p_ < P && q < Q) tile[ty + 8 * k][tx] = src[p_ * Q + q];
        }
        __syncthreads();
        const long p2 = tp * 32 + tx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long q2 = tq * 32 + ty + 8 * k;
            if (q2 < Q && p2 < P) dst[q2 * P + p2] = tile[tx][ty + 8 * k];
        }
        __syncthreads();
    }
}

static inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
static inline int next_pow2_log(long n) { int l = 0; while ((1L << l) < n) ++l; return l; }

static inline int ilog2_exact(int64_t n) {
    if (n < 1 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1LL << l) < n) ++l;
    return l;
}

static inline unsigned mirror_grid(long nrows) {
    long cap = (long)sm_count() * 16;
    return (unsigned)(nrows < cap ? nrows : cap);
}

static inline unsigned ew_grid(long total) {
    long g = (total + 255) / 256;
    long cap = (long)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ------------------------------------------------------------------------------------------------
// fftn composition
// ------------------------------------------------------------------------------------------------
static std::map<std::tuple<int, int, long, int>, std::pair<void*, void*>> g_chirp;  // (device, dtype, N, log2M) -> (chirp, filter)

template <typename T> static bool fast_len(long n, bool contiguous) {
    const int l2 = ilog2_exact(n);
    if (l2 < 1) return false;
    return l2 <= (contiguous ? TypeCfg<T>::MAX_ROWS_LOG2 : TypeCfg<T>::MAX_COLS_LOG2);
}
// workspace (bytes) one C2C pass of length n over an [A][n][B] view needs
template <typename T> static size_t pass_workspace(long A, long n, long B) {
    if (n <= 1 || fast_len<T>(n, B == 1) || n <= kSmallDft || smooth_len_ok<T>(n)) return 0;
    const size_t al = 256;
    if (ilog2_exact(n) > 0) return (((size_t)A * n * B * sizeof(cplx<T>)) + al) & ~(al - 1);   // four-step scratch
    const int lm = next_pow2_log(2 * n - 1);
    const long M = 1L << lm;
    return ((((size_t)A * M * B * sizeof(cplx<T>)) + al) & ~(al - 1)) + pass_workspace<T>(A, M, B);
}

template <typename T>
static int c2c_pass(const cplx<T>* src, cplx<T>* dst, long A, long n, long B, int inverse, T scale, void* work, size_t work_bytes,
                    cudaStream_t st, const ColsC2C<T>* hooks = nullptr);

template <typename T> static int get_chirp(long N, int log2M, const cplx<T>** chirp, const cplx<T>** filt, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    const int dt = sizeof(T) == 4 ? 0 : 1;
    auto key = std::make_tuple(dev, dt, N, log2M);
    {
        std::lock_guard<std::mutex> lk(g_tw_mu);
        auto it = g_chirp.find(key);
        if (it != g_chirp.end()) { *chirp = (const cplx<T>*)it->second.first; *filt = (const cplx<T>*)it->second.second; return 0; }
    }
    const long M = 1L << log2M;
    std::vector<cplx<T>> hb(N), hc(M);
    for (long m = 0; m < M; ++m) hc[m] = mk<T>(0, 0);
    for (long n = 0; n < N; ++n) {
        const long r = (long)(((__int128)n * n) % (2 * N));  // exp(-i pi n^2 / N) has period 2N in n^2
        const double ang = -M_PI * (double)r / (double)N;
        hb[n] = mk<T>((T)cos(ang), (T)sin(ang));
        cplx<T> cj = mk<T>(hb[n].x, -hb[n].y);
        hc[n] = cj;
        if (n > 0) hc[M - n] = cj;
    }
    cplx<T>*db = nullptr, *dc = nullptr, *df = nullptr;
    void* tmp = nullptr;
    const size_t ws = pass_workspace<T>(1, M, 1);
    cudaError_t e = cudaMalloc(&db, N * sizeof(cplx<T>));
    if (e == cudaSuccess) e = cudaMalloc(&dc, M * sizeof(cplx<T>));
    if (e == cudaSuccess) e = cudaMalloc(&df, M * sizeof(cplx<T>));
    if (e == cudaSuccess && ws) e = cudaMalloc(&tmp, ws);
    if (e == cudaSuccess) e = cudaMemcpy(db, hb.data(), N * sizeof(cplx<T>), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dc, hc.data(), M * sizeof(cplx<T>), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("chirp table: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
    int rc = c2c_pass<T>(dc, df, 1, M, 1, 0, (T)1, tmp, ws, st);
    if (rc) return rc;
    cudaStreamSynchronize(st);
    cudaFree(dc);
    if (tmp) cudaFree(tmp);
    std::lock_guard<std::mutex> lk(g_tw_mu);
    g_chirp[key] = std::make_pair((void*)db, (void*)df);
    *chirp = db; *filt = df;
    return 0;
}

// one C2C pass along the middle axis of [A][n][B]; src may equal dst
template <typename T>
static int c2c_pass(const cplx<T>* src, cplx<T>* dst, long A, long n, long B, int inverse, T scale, void* work, size_t work_bytes,
                    cudaStream_t st, const ColsC2C<T>* hooks) {
    if (hooks && !(ilog2_exact(n) > 0 && B > 1 && (double)n * (double)B < 2147483648.0 && (double)A * (double)n < 2147483648.0)) {
        set_error("fft2r: strided power-of-two axis with fewer than 2^31 elements per item required"); return XRFTB_EUNSUPPORTED;
    }
    if (n == 1) {
        if (src != dst || scale != (T)1) {
            const long total = A * B;
            roll_scale_kernel<T><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const T*>(src), reinterpret_cast<T*>(dst), 1, 1, total, 0, 0, 0, 2, scale, total);
            return check_launch("roll_scale_kernel");
        }
        return 0;
    }
    if (fast_len<T>(n, B == 1)) {
        const int l2 = ilog2_exact(n);
        if (B == 1) return rows_c2c<T>(src, dst, l2, A, n, n, inverse, scale, st);
        return cols_c2c<T>(src, dst, l2, A, B, inverse, scale, st, hooks);
    }
    if (small_direct<T>(n, B)) {
        const long nseq = A * B;
        dft_small_kernel<T><<<(unsigned)((nseq + 127) / 128 > 4096 ? 4096 : (nseq + 127) / 128), 128, 0, st>>>(src, dst, (int)n, A, B, inverse ? 1 : 0, scale);
        return check_launch("dft_small_kernel");
    }
    if (smooth_len_ok<T>(n)) return smooth_c2c<T>(src, dst, A, n, B, inverse, scale, st);   // 2^a 3^b 5^c 7^d: mixed-radix kernel
    const size_t need = pass_workspace<T>(A, n, B);
    if (!work || work_bytes < need) { set_error("fftn: workspace too small for length %ld (%zu < %zu)", n, work_bytes, need); return XRFTB_EWORKSPACE; }
    cplx<T>* w = reinterpret_cast<cplx<T>*>(work);
    const int l2 = ilog2_exact(n);
    if (l2 > 0) {
        // ---- four-step: n = n1 * n2, both within the single-pass limits
        const long n1 = 1L << (l2 / 2), n2 = n / n1;
        if (ilog2_exact(n2) > TypeCfg<T>::MAX_COLS_LOG2) { set_error("fftn: length %ld too long", n); return XRFTB_EUNSUPPORTED; }
        if ((double)n2 * (double)B >= 2147483648.0 || (double)A * (double)n1 >= 2147483648.0) { set_error("fftn: four-step view of length %ld x %ld columns too large", n, B); return XRFTB_EUNSUPPORTED; }
        const long total = A * n * B;
        // the twiddle product w_n^(k1 i2) rides on the stores of the first pass; with B > 1 (strided second pass) the final
        // transposition rides on the stores of the second: two passes over the data instead of four
        // (w_n^m from a table of n entries up to n = 2^13, else from a coarse and a fine table of n / 8192 and 8192 entries)
        const cplx<T>* tw4 = l2 <= 13 ? twiddle_fft<T>(l2) : twiddle_head<T>(l2);
        const cplx<T>* tw4_hi = l2 <= 13 ? nullptr : twiddle_fft<T>(l2 - 13);
        if (!tw4 || (l2 > 13 && !tw4_hi)) return XRFTB_ECUDA;
        ColsC2C<T> ea{}, eb{};
        if (hooks) { ea = *hooks; eb = *hooks; }
        ea.tw4 = tw4; ea.tw4_div = B; ea.tw4_hi = tw4_hi; ea.row_mul = n2; ea.row_div = B;   // step A: rows r = i1 n2 + i2; store hooks belong to step B
        ea.out_ramp = nullptr; ea.out_roll = 0; ea.out_hi = 0; ea.out_conj = 0;
        eb.tr_n1 = (int)n1; eb.in_ramp = nullptr; eb.in_roll = 0; eb.in_hi = 0; eb.in_rows = 0; eb.in_conj = 0;   // step B: load hooks belong to step A
        if (int rc = cols_c2c<T>(src, w, ilog2_exact(n1), A, n2 * B, inverse, (T)1, st, &ea)) return rc;            // FFT over i1, x twiddle
        if (B > 1) return cols_c2c<T>(w, dst, ilog2_exact(n2), A * n1, B, inverse, scale, st, &eb);                  // FFT over i2, transposed store
        if (hooks) { set_error("fft2r: contiguous four-step pass has no hooks"); return XRFTB_EUNSUPPORTED; }
        if (int rc = rows_c2c<T>(w, w, ilog2_exact(n2), A * n1, n2, n2, inverse, (T)1, st)) return rc;                // FFT over i2 (contiguous)
        fourstep_transpose_kernel<T><<<ew_grid(total), 256, 0, st>>>(w, dst, n1, n2, B, scale, total);
        return check_launch("fourstep_transpose");
    }
    // ---- Bluestein on top of the power-of-two machinery (recursive: M may itself need the four-step)
    const int lm = next_pow2_log(2 * n - 1);
    const long M = 1L << lm;
    if (lm > 26) { set_error("fftn: length %ld too long", n); return XRFTB_EUNSUPPORTED; }
    const size_t mine = (((size_t)A * M * B * sizeof(cplx<T>)) + 256) & ~(size_t)255;
    char* sub = reinterpret_cast<char*>(work) + mine;
    const size_t sub_bytes = work_bytes - mine;
    const cplx<T>*chirp, *filt;
    if (int rc = get_chirp<T>(n, lm, &chirp, &filt, st)) return rc;
    // The chirp products, the zero padding to M, the filter product and the final crop ride on the loads and stores of the two
    // power-of-two transforms (hooks of ColsC2C / RowsC2C): two passes over the data instead of five.
    const T s2 = (T)((double)scale / (double)M);
    if (B > 1 && (double)M * (double)B < 2147483648.0 && (double)A * (double)M < 2147483648.0) {
        ColsC2C<T> h1{}, h2{};
        h1.hook_n = M; h1.in_rows = n; h1.in_hi = n; h1.in_ramp = chirp; h1.in_conj = inverse; h1.out_ramp = filt;
        h2.hook_n = M; h2.out_ramp = chirp; h2.out_conj = inverse; h2.out_hi = n;
        if (int rc = c2c_pass<T>(src, w, A, M, B, 0, (T)1, sub, sub_bytes, st, &h1)) return rc;
        return c2c_pass<T>(w, dst, A, M, B, 1, s2, sub, sub_bytes, st, &h2);
    }
    if (B == 1 && fast_len<T>(M, true)) {
        RowsC2C<T> r1{}, r2{};
        r1.in_len = (int)n; r1.in_conj = inverse; r1.in_ramp = chirp; r1.out_ramp = filt;
        r2.out_ramp = chirp; r2.out_conj = inverse; r2.out_len = (int)n;
        if (int rc = rows_c2c<T>(src, w, lm, A, n, M, 0, (T)1, st, &r1)) return rc;
        return rows_c2c<T>(w, dst, lm, A, M, n, 1, s2, st, &r2);
    }
    // contiguous sequences longer than one CTA's shared memory: separate chirp / filter passes around the four-step transforms
    const long tw_ = A * M * B;
    bluestein_pre_kernel<T, false><<<ew_grid(tw_), 256, 0, st>>>(src, w, chirp, A, n, M, B, inverse, tw_);
    if (int rc = check_launch("bluestein_pre")) return rc;
    if (int rc = c2c_pass<T>(w, w, A, M, B, 0, (T)1, sub, sub_bytes, st)) return rc;
    bluestein_mul_kernel<T><<<ew_grid(tw_), 256, 0, st>>>(w, filt, M, B, tw_);
    if (int rc = check_launch("bluestein_mul")) return rc;
    if (int rc = c2c_pass<T>(w, w, A, M, B, 1, (T)(1.0 / M), sub, sub_bytes, st)) return rc;
    const long to = A * n * B;
    bluestein_post_kernel<T, false><<<ew_grid(to), 256, 0, st>>>(w, dst, chirp, A, n, n, M, B, inverse, scale, to);
    return check_launch("bluestein_post");
}

template <typename T> static bool fast_real_len(long N) {
    const int l2 = ilog2_exact(N);
    return l2 >= 2 && l2 - 1 <= TypeCfg<T>::MAX_ROWS_LOG2;
}

template <typename T>
static size_t fftn_workspace_impl(int kind, int ndim, const int64_t* shape, int naxes, const int* axes) {
    std::vector<int64_t> cshape(shape, shape + ndim);
    const bool real_kind = (kind == XRFTB_R2C || kind == XRFTB_C2R);
    const long N = shape[ndim - 1];
    if (real_kind) cshape[ndim - 1] = N / 2 + 1;
    size_t need = 0;
    auto view = [&](int axis, long& A, long& B) { A = 1; B = 1; for (int d = 0; d < axis; ++d) A *= cshape[d]; for (int d = axis + 1; d < ndim; ++d) B *= cshape[d]; };
    for (int i = 0; i < naxes - (real_kind ? 1 : 0); ++i) {
        long A, B; view(axes[i], A, B);
        size_t w = pass_workspace<T>(A, cshape[axes[i]], B);
        if (w > need) need = w;
    }
    size_t base = 0;
    if (real_kind) {
        long nseq = 1; for (int d = 0; d < ndim - 1; ++d) nseq *= shape[d];
        if (!fast_real_len<T>(N) && N > kSmallDft) {
            // generic real path: full-length complex copy + whatever the complex pass needs
            size_t last = (((size_t)nseq * N * sizeof(cplx<T>)) + 256) & ~(size_t)255;
            last += pass_workspace<T>(nseq, N, 1);
            if (last > need) need = last;
        }
        if (kind == XRFTB_C2R && naxes > 1) base = (((size_t)nseq * (N / 2 + 1) * sizeof(cplx<T>)) + 256) & ~(size_t)255;  // private copy of the input
    }
    return base + need + 512;
}

template <typename T>
static int fftn_impl(const void* in, void* out, void* work, size_t work_bytes, int kind, int ndim, const int64_t* shape,
                     int naxes, const int* axes, cudaStream_t st) {
    using C = cplx<T>;
    std::vector<int64_t> cshape(shape, shape + ndim);
    const bool real_kind = (kind == XRFTB_R2C || kind == XRFTB_C2R);
    if (real_kind) {
        if (axes[naxes - 1] != ndim - 1) { set_error("R2C/C2R: the last listed axis must be the last array axis"); return XRFTB_EINVAL; }
        cshape[ndim - 1] = shape[ndim - 1] / 2 + 1;
    }
    double norm = 1.0;
    for (int i = 0; i < naxes; ++i) norm *= (double)shape[axes[i]];
    const bool inverse = (kind == XRFTB_C2C_INV || kind == XRFTB_C2R);
    const T inv_scale = inverse ? (T)(1.0 / norm) : (T)1;
    char* wbase = reinterpret_cast<char*>(work);
    size_t wleft = work ? work_bytes : 0;
    auto view = [&](int axis, long& A, long& B) { A = 1; B = 1; for (int d = 0; d < axis; ++d) A *= cshape[d]; for (int d = axis + 1; d < ndim; ++d) B *= cshape[d]; };

    if (!real_kind) {
        const C* src = reinterpret_cast<const C*>(in);
        C* dst = reinterpret_cast<C*>(out);
        for (int i = 0; i < naxes; ++i) {
            long A, B; view(axes[i], A, B);
            int rc = c2c_pass<T>(src, dst, A, cshape[axes[i]], B, inverse ? 1 : 0, (i == naxes - 1) ? inv_scale : (T)1, wbase, wleft, st);
            if (rc) return rc;
            src = dst;
        }
        return 0;
    }
    const long N = shape[ndim - 1], H = N / 2 + 1;
    long nseq = 1;
    for (int d = 0; d < ndim - 1; ++d) nseq *= shape[d];
    const int l2 = ilog2_exact(N);
    const bool fast_real = fast_real_len<T>(N);
    const unsigned small_grid = (unsigned)((nseq + 127) / 128 > 4096 ? 4096 : (nseq + 127) / 128);
    if (N < 2) { set_error("real transforms need at least 2 points on the real axis"); return XRFTB_EINVAL; }
    if (kind == XRFTB_R2C) {
        C* dst = reinterpret_cast<C*>(out);
        if (fast_real) {
            RowsR2CFused<T> io{};
            io.in = reinterpret_cast<const T*>(in); io.in_row_stride = N; io.logNy = 0; io.detrend = 0; io.moments = nullptr;
            io.wy = nullptr; io.wx = nullptr; io.out = dst; io.logC = -1; io.out_seq_stride = H;
            if (int rc = rows_r2c<T>(io, l2 - 1, nseq, st)) return rc;
        } else if (small_direct_real<T>(N)) {
            dft_small_kernel<T><<<small_grid, 128, 0, st>>>(in, dst, (int)N, nseq, 1, 2, (T)1);
            if (int rc = check_launch("dft_small_kernel")) return rc;
        } else if (smooth_real_ok<T>(N)) {
            // even smooth length: packed half-length mixed-radix transform, split in its stores (one pass, no promotion)
            if (int rc = smooth_r2c<T>(reinterpret_cast<const T*>(in), dst, nseq, N, st)) return rc;
        } else {
            // promote to complex, full-length complex pass (mixed radix / four-step / Bluestein), keep k <= N/2
            const size_t cb = (((size_t)nseq * N * sizeof(C)) + 256) & ~(size_t)255;
            if (wleft < cb) { set_error("rfftn: workspace too small (%zu < %zu)", wleft, cb); return XRFTB_EWORKSPACE; }
            C* full = reinterpret_cast<C*>(wbase);
            real_to_cplx_kernel<T><<<ew_grid(nseq * N), 256, 0, st>>>(reinterpret_cast<const T*>(in), full, nseq * N);
            if (int rc = c2c_pass<T>(full, full, nseq, N, 1, 0, (T)1, wbase + cb, wleft - cb, st)) return rc;
            crop_kernel<T, false><<<ew_grid(nseq * H), 256, 0, st>>>(full, dst, N, H, nseq * H);
            if (int rc = check_launch("crop_kernel")) return rc;
        }
        for (int i = 0; i < naxes - 1; ++i) {
            long A, B; view(axes[i], A, B);
            if (int rc = c2c_pass<T>(dst, dst, A, cshape[axes[i]], B, 0, (T)1, wbase, wleft, st)) return rc;
        }
        return 0;
    }
    // ---- C2R
    const C* src = reinterpret_cast<const C*>(in);
    if (naxes > 1) {
        const size_t copy_bytes = (((size_t)nseq * H * sizeof(C)) + 256) & ~(size_t)255;
        if (wleft < copy_bytes) { set_error("irfftn: workspace too small"); return XRFTB_EWORKSPACE; }
        C* w = reinterpret_cast<C*>(wbase);
        wbase += copy_bytes;
        wleft -= copy_bytes;
        for (int i = 0; i < naxes - 1; ++i) {
            long A, B; view(axes[i], A, B);
            if (int rc = c2c_pass<T>(src, w, A, cshape[axes[i]], B, 1, (T)1, wbase, wleft, st)) return rc;
            src = w;
        }
    }
    if (fast_real) return rows_c2r<T>(src, H, reinterpret_cast<T*>(out), N, l2 - 1, nseq, inv_scale * (T)2, st);  // half-length inverse: 1/M = 2/N
    if (small_direct_real<T>(N)) {
        dft_small_kernel<T><<<small_grid, 128, 0, st>>>(src, out, (int)N, nseq, 1, 3, inv_scale);
        return check_launch("dft_small_kernel");
    }
    if (smooth_real_ok<T>(N)) return smooth_c2r<T>(src, reinterpret_cast<T*>(out), nseq, N, inv_scale * (T)2, st);   // 1 / (N/2)
    {
        const size_t cb = (((size_t)nseq * N * sizeof(C)) + 256) & ~(size_t)255;
        if (wleft < cb) { set_error("irfftn: workspace too small (%zu < %zu)", wleft, cb); return XRFTB_EWORKSPACE; }
        C* ext = reinterpret_cast<C*>(wbase);
        herm_extend_kernel<T><<<ew_grid(nseq * N), 256, 0, st>>>(src, ext, N, 1, nseq * N);
        if (int rc = c2c_pass<T>(ext, ext, nseq, N, 1, 1, inv_scale, wbase + cb, wleft - cb, st)) return rc;
        crop_kernel<T, true><<<ew_grid(nseq * N), 256, 0, st>>>(ext, out, N, N, nseq * N);
        return check_launch("crop_kernel");
    }
}


// ------------------------------------------------------------------------------------------------
// 2-D real transform with the neighbouring elementwise steps of xrft.fft / xrft.ifft folded into its passes (xrftb_fft2r):
//   forward: zero padding (xrft.pad) as load predicates + skipped padding rows | R2C rows x ramp_x x scale | strided passes,
//            the last one x ramp_y                                              (xrft.py:398-404, 462-472; padding.py:157-181)
//   inverse: strided passes reading rolled rows x ramp_y, the last one storing rolled + cropped rows | C2R rows x ramp_x on
//            the way in, rolled + cropped + scaled on the way out                (xrft.py:574-621, 641-642; padding.py:425-446)
// Power-of-two sizes on the single-pass / four-step kernels; anything else reports XRFTB_EUNSUPPORTED and the caller composes
// xrftb_pad / xrftb_spectral_post / xrftb_fftn / xrftb_roll_scale instead.
// ------------------------------------------------------------------------------------------------
template <typename T> static bool fft2r_supported(const xrftb_fft2r_desc& q) {
    const int ly = ilog2_exact(q.ny), lx = ilog2_exact(q.nx);
    if (ly < 1 || lx < 2 || !fast_real_len<T>(q.nx)) return false;
    if (fast_len<T>(q.ny, false)) return true;
    const long n1 = 1L << (ly / 2), n2 = q.ny / n1;
    return ilog2_exact(n2) <= TypeCfg<T>::MAX_COLS_LOG2 && ly <= 26;
}
template <typename T> static size_t fft2r_workspace_impl(const xrftb_fft2r_desc& q) {
    if (!fft2r_supported<T>(q)) return 0;
    const size_t H = (size_t)q.nx / 2 + 1;
    size_t need = pass_workspace<T>(q.batch, q.ny, (long)H);
    if (q.inverse) need += align256((size_t)q.batch * (size_t)(q.out_ny > 0 ? q.out_ny : q.ny) * H * sizeof(cplx<T>));
    return need + 512;
}
template <typename T> static int fft2r_impl(const xrftb_fft2r_desc& q, cudaStream_t st) {
    using C = cplx<T>;
    if (!fft2r_supported<T>(q)) { set_error("fft2r: size %ld x %ld is not covered by the fused passes", (long)q.ny, (long)q.nx); return XRFTB_EUNSUPPORTED; }
    const long H = q.nx / 2 + 1;
    const int lx = ilog2_exact(q.nx);
    char* wbase = reinterpret_cast<char*>(q.work);
    if (q.work_bytes < fft2r_workspace_impl<T>(q) - 512 || (!q.work && fft2r_workspace_impl<T>(q) > 512)) { set_error("fft2r: workspace too small"); return XRFTB_EWORKSPACE; }
    if (!q.inverse) {
        const bool padded = q.in_ny > 0;
        if (padded && (q.in_off_y < 0 || q.in_off_x < 0 || q.in_off_y + q.in_ny > q.ny || q.in_off_x + q.in_nx > q.nx || q.in_nx < 1)) { set_error("fft2r: input placement outside the transform"); return XRFTB_EINVAL; }
        RowsR2CFused<T> io{};
        io.in = reinterpret_cast<const T*>(q.in); io.in_row_stride = padded ? q.in_nx : q.nx; io.logNy = 0; io.detrend = 0; io.moments = nullptr;
        io.wy = nullptr; io.wx = nullptr; io.out = reinterpret_cast<C*>(q.out); io.logC = -1; io.out_seq_stride = H; io.rowstats = nullptr;
        if (padded) { io.rows_in = q.in_ny; io.rows_out = q.ny; io.row_off = q.in_off_y; io.col_lo = (int)q.in_off_x; io.col_hi = (int)(q.in_off_x + q.in_nx); }
        io.out_ramp = reinterpret_cast<const C*>(q.ramp_x); io.out_scale = (T)q.scale;
        if (int rc = rows_r2c<T>(io, lx - 1, q.batch * (padded ? q.in_ny : q.ny), st)) return rc;
        ColsC2C<T> hk{};
        hk.hook_n = q.ny;
        if (padded) { hk.in_lo = q.in_off_y; hk.in_hi = q.in_off_y + q.in_ny; }
        hk.out_ramp = reinterpret_cast<const C*>(q.ramp_y);
        const bool any = padded || q.ramp_y;
        return c2c_pass<T>(reinterpret_cast<C*>(q.out), reinterpret_cast<C*>(q.out), q.batch, q.ny, H, 0, (T)1, wbase, q.work_bytes, st, any ? &hk : nullptr);
    }
    // ---- inverse
    const long cy = q.out_ny > 0 ? q.out_ny : q.ny, cx = q.out_nx > 0 ? q.out_nx : q.nx;
    const long oy = q.out_ny > 0 ? q.out_off_y : 0, ox = q.out_nx > 0 ? q.out_off_x : 0;
    if (oy < 0 || ox < 0 || oy + cy > q.ny || ox + cx > q.nx) { set_error("fft2r: crop outside the transform"); return XRFTB_EINVAL; }
    const size_t w2_bytes = align256((size_t)q.batch * cy * H * sizeof(C));
    C* w2 = reinterpret_cast<C*>(wbase);
    ColsC2C<T> hk{};
    hk.hook_n = q.ny;
    hk.in_roll = ((q.in_roll_y % q.ny) + q.ny) % q.ny;
    hk.in_ramp = reinterpret_cast<const C*>(q.ramp_y);
    hk.out_roll = ((q.out_roll_y % q.ny) + q.ny) % q.ny;
    if (q.out_ny > 0) { hk.out_lo = oy; hk.out_hi = oy + cy; }
    const bool any = hk.in_roll || hk.in_ramp || hk.out_roll || hk.out_hi;
    if (int rc = c2c_pass<T>(reinterpret_cast<const C*>(q.in), w2, q.batch, q.ny, H, 1, (T)1, wbase + w2_bytes, q.work_bytes - w2_bytes, st, any ? &hk : nullptr)) return rc;
    RowsC2R<T> ex{};
    ex.in_ramp = reinterpret_cast<const C*>(q.ramp_x);
    ex.out_roll = (int)(((q.out_roll_x % q.nx) + q.nx) % q.nx);
    if (q.out_nx > 0) { ex.out_lo = (int)ox; ex.out_hi = (int)(ox + cx); }
    const double norm = (double)q.ny * (double)q.nx;
    return rows_c2r<T>(w2, H, reinterpret_cast<T*>(q.out), cx, lx - 1, q.batch * cy, (T)(2.0 * q.scale / norm), st, &ex);
}

// ------------------------------------------------------------------------------------------------
// fused 2-D real spectrum
// ------------------------------------------------------------------------------------------------
// TMA descriptor of the POWER output viewed as [rows][W] float32, box = [box_rows][box_cols]
static bool encode_out_tmap(CUtensorMap* tm, void* base, long rows, long W, int box_cols, int box_rows) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)W * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T> static size_t interm_bytes_per_item(int ny, int nx, int C) {
    const long ntile = (nx / 2) / C + 1;
    return (size_t)ntile * ny * C * sizeof(cplx<T>);
}

// row-line detrend: per-item rowstats (float4 / row) + ag (cplx<T> / row); fixed: wj table + its fp64 transform scratch
template <typename T> static size_t rowline_item_bytes(int ny) { return (size_t)ny * (sizeof(float4) + sizeof(cplx<T>)); }
template <typename T> static size_t rowline_fixed_bytes(int nx, int C) {
    const size_t ncols = (size_t)((nx / 2) / C + 1) * C;
    return align256(ncols * 2 * sizeof(cplx<T>)) + align256((size_t)2 * nx * sizeof(double)) + align256((size_t)2 * (nx / 2 + 1) * sizeof(double2));
}
static bool rowline_enabled() { return option(OPT_ROWLINE) != 0; }

// "columns first" order (see ColsR2CPack / RowsC2CPower): eligible for the full-width power spectrum
// chain taken by this thread's last xrftb_spectrum2d call: 0 rows first (+ Hermitian mirror pass), 1 columns first, 2 columns
// first with packed column spectra (z mode), 3 two-field z mode
struct LastPath { int v = 0; void store(int x) { v = x; } int load() const { return v; } };
static thread_local LastPath g_last_path;
static bool colsfirst_enabled() { return option(OPT_COLS_FIRST) != 0; }
template <typename T> static bool colsfirst_eligible(const xrftb_spectrum2d_desc& q, int ly, int lx) {
    if (q.keep_half || q.weight_x) return false;
    if (q.mode == XRFTB_EPI_BINS_POWER) {
        // radial bins in pass 2 (rows_bins_kernel): float32, symmetric LUT (radial bins), shapes with a static row mapping
        if (!std::is_same<T, float>::value || !q.lut_symmetric || !bins_static_enabled()) return false;
        if (!rows_bins_shape_ok(lx, q.ny) || q.nbins > q.nx || q.nbins > 4096) return false;
    } else if (q.mode != XRFTB_EPI_POWER) {
        return false;
    }
    const int C = colsfirst_tile_width<T>(ly);
    return C >= 1 && ly >= 1 && ly <= TypeCfg<T>::MAX_COLS_LOG2 && lx >= 1 && lx <= TypeCfg<T>::MAX_ROWS_LOG2 && q.nx >= 2 * C;
}
template <typename T> static size_t colsfirst_item_bytes(int ny, int nx) { return (size_t)(ny / 2 + 1) * nx * sizeof(cplx<T>); }

// column-line detrend tables of the columns-first order: per item colstats (float4 / column) + ag (cplx<T> / column);
// fixed: wj table (ny/2+1 pairs) + its fp64 transform scratch
template <typename T> static size_t colline_item_bytes(int nx) { return (size_t)nx * (sizeof(float4) + sizeof(cplx<T>)); }
template <typename T> static size_t colline_fixed_bytes(int ny) {
    return align256((size_t)(ny / 2 + 1) * 2 * sizeof(cplx<T>)) + align256((size_t)2 * ny * sizeof(double)) + align256((size_t)2 * (ny / 2 + 1) * sizeof(double2));
}

template <typename T>
static int spectrum2d_colsfirst(const xrftb_spectrum2d_desc& q, int ly, int lx, cudaStream_t st) {
    using C_ = cplx<T>;
    const int C = colsfirst_tile_width<T>(ly);
    const int H = q.ny / 2 + 1;
    // column-line detrend (no moments pass): float32 with two packed columns per thread, ny long enough for the fp64 table transform
    const bool colline = std::is_same<T, float>::value && q.detrend && rowline_enabled() && C >= 2 && ly >= 3;
    const size_t per_item = colsfirst_item_bytes<T>(q.ny, q.nx) + (colline ? colline_item_bytes<T>(q.nx) : 0);
    const size_t mom_bytes = (size_t)q.batch * 4 * sizeof(double);
    const size_t mom_region = align256(mom_bytes) + (colline ? colline_fixed_bytes<T>(q.ny) : 0);
    if (!q.work || q.work_bytes < mom_region + per_item) { set_error("spectrum2d: workspace too small (%zu < %zu)", q.work_bytes, mom_region + per_item); return XRFTB_EWORKSPACE; }
    long bchunk = (long)((q.work_bytes - mom_region) / per_item);
    if (bchunk > q.batch) bchunk = q.batch;
    { const long nch = (q.batch + bchunk - 1) / bchunk; bchunk = (q.batch + nch - 1) / nch; }
    double* mom = reinterpret_cast<double*>(q.work);
    C_* interm = reinterpret_cast<C_*>(reinterpret_cast<char*>(q.work) + mom_region);
    const long item = (long)q.ny * q.nx;
    const T* in = reinterpret_cast<const T*>(q.in1);
    float4* colstats = nullptr;
    C_* ag = nullptr;
    C_* wj = nullptr;
    if (colline) {
        char* fx = reinterpret_cast<char*>(q.work) + align256(mom_bytes);
        const int My = q.ny / 2;
        wj = reinterpret_cast<C_*>(fx);
        double* wrows = reinterpret_cast<double*>(fx + align256((size_t)H * 2 * sizeof(C_)));
        double2* wspec = reinterpret_cast<double2*>(reinterpret_cast<char*>(wrows) + align256((size_t)2 * q.ny * sizeof(double)));
        colstats = reinterpret_cast<float4*>(reinterpret_cast<char*>(interm) + (size_t)bchunk * colsfirst_item_bytes<T>(q.ny, q.nx));
        ag = reinterpret_cast<C_*>(colstats + (size_t)bchunk * q.nx);
        // transforms of w_y(i) and w_y(i)(i - ic), k <= ny/2: one fp64 real transform of two rows, once per call
        rowline_wj_in_kernel<T><<<(q.ny + 255) / 256, 256, 0, st>>>(reinterpret_cast<const T*>(q.win_y), wrows, q.ny);
        if (int rc = check_launch("rowline_wj_in_kernel")) return rc;
        RowsR2CFused<double> wio{};
        wio.in = wrows; wio.in_row_stride = q.ny; wio.logNy = 0; wio.detrend = 0; wio.moments = nullptr; wio.wy = nullptr; wio.wx = nullptr;
        wio.out = wspec; wio.logC = -1; wio.out_seq_stride = H; wio.rowstats = nullptr;
        if (int rc = rows_r2c<double>(wio, ly - 1, 2, st)) return rc;
        rowline_wj_out_kernel<T><<<(unsigned)((H + 255) / 256), 256, 0, st>>>(wspec, wj, My, H);
        if (int rc = check_launch("rowline_wj_out_kernel")) return rc;
    } else if (q.detrend) {
        cudaError_t e = cudaMemsetAsync(mom, 0, mom_bytes, st);
        if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
        ProfScope ps_(PROF_MOMENTS, st);
        int chunks = (int)((8L * sm_count() + q.batch - 1) / q.batch);
        if (chunks < 1) chunks = 1;
        if (chunks > q.ny) chunks = q.ny;
        dim3 grid(chunks, (unsigned)q.batch);
        moments_kernel<T><<<grid, 256, 0, st>>>(in, mom, 1, q.ny, q.nx, chunks, q.batch);
        if (int rc = check_launch("moments_kernel")) return rc;
    }
    const int tiles_per_item = q.nx / (2 * C);
    for (long b0 = 0; b0 < q.batch; b0 += bchunk) {
        const long nb = (q.batch - b0 < bchunk) ? q.batch - b0 : bchunk;
        bool zmode = false;
        {
            ColsR2CPack<T> io{in + b0 * item, q.nx, tiles_per_item, q.detrend, mom + b0 * 4,
                              reinterpret_cast<const T*>(q.win_y), reinterpret_cast<const T*>(q.win_x), interm, colstats, 0, {}};
            // tensor-map fed (asynchronous) variant: float32, >= 2 exchange stages, 16-byte aligned rows
            const int async_on = option(OPT_COLS_ASYNC);
            bool use_async = false;
            if (async_on && std::is_same<T, float>::value && ly > TypeCfg<T>::LOGE && C >= 2 && (q.nx * sizeof(T)) % 16 == 0) {
                const int box_rows = q.ny < 256 ? q.ny : 256;
                if (encode_out_tmap(&io.tmap, const_cast<T*>(in + b0 * item), nb * q.ny, q.nx, 2 * C, box_rows)) { io.box_rows = box_rows; use_async = true; }
            }
            // "z" mode: pass 1 stores the packed column spectra, pass 2 separates the real columns in its loads (RowsZPower)
            const int zpack_on = option(OPT_ZPACK);
            // tiles of >= 4 packed columns only: with 2 the rows of Z are 16 bytes (half sectors: measured 2.7x slower stores)
            zmode = use_async && zpack_on && rows_z_supported(lx - 1) && C >= 4 && q.mode == XRFTB_EPI_POWER;
            io.zout = zmode ? interm : nullptr;
            if (zmode) g_last_path.store(2);
            io.ztma = 0;
            if (zmode && C * sizeof(C_) >= 32 && q.ny >= 512) {   // tensor stores of Z (XRFTB_ZTMA=0: 16-byte stores from registers)
                const int ztma_on = option(OPT_ZTMA);
                const int zbox = q.ny / 4 < 256 ? q.ny / 4 : 256;
                if (ztma_on && encode_out_tmap(&io.ztmap, interm, nb * q.ny, q.nx, 2 * C, zbox)) { io.ztma = 1; io.zbox_rows = zbox; }
            }
            ProfScope ps_(PROF_COLS, st);
            if (int rc = cols_r2c_pack<T>(io, ly, C, nb * tiles_per_item, use_async, st)) return rc;
        }
        if (colline) {
            // the row-line completion kernel with the roles of the axes swapped: nx lines (columns) of length ny
            ProfScope ps_(PROF_MOMENTS, st);
            rowline_fix_kernel<T><<<dim3((unsigned)nb, (unsigned)((q.nx + 1023) / 1024)), 256, 0, st>>>(colstats, ag, reinterpret_cast<const T*>(q.win_x), q.nx, q.ny, q.detrend);
            if (int rc = check_launch("rowline_fix_kernel")) return rc;
        }
        if (q.mode == XRFTB_EPI_BINS_POWER) {
            // rows ky in [0, Ny/2) of every plane, then the Nyquist row: |F|^2 goes straight into the planes' radial bins
            if constexpr (std::is_same<T, float>::value) {
                RowsBins io{{interm, nullptr, ly, H, q.shift_y, q.shift_x, (T)q.scale, colline ? ag : nullptr, wj}, q.lut, q.bins + b0 * q.nbins, q.nbins, 0, q.ny / 2, nb};
                ProfScope ps_(PROF_ROWS, st);
                if (int rc = rows_bins(io, lx, st)) { if (rc > 0) set_error("spectrum2d: radial-bin pass does not cover %d x %d", q.ny, q.nx); return rc > 0 ? XRFTB_EUNSUPPORTED : rc; }
                io.ky0 = q.ny / 2; io.rows = 1;
                if (int rc = rows_bins(io, lx, st)) { if (rc > 0) set_error("spectrum2d: radial-bin pass does not cover %d x %d", q.ny, q.nx); return rc > 0 ? XRFTB_EUNSUPPORTED : rc; }
            }
        } else if (zmode) {
            RowsZPower<T> io{interm, reinterpret_cast<T*>(q.out) + b0 * item, ly, H, q.shift_y, q.shift_x, (T)q.scale, colline ? ag : nullptr, wj, nullptr};
            ProfScope ps_(PROF_ROWS, st);
            if (int rc = rows_z_power<T>(io, lx - 1, nb * H, st)) return rc;
        } else {
            RowsC2CPower<T> io{interm, reinterpret_cast<T*>(q.out) + b0 * item, ly, H, q.shift_y, q.shift_x, (T)q.scale, colline ? ag : nullptr, wj};
            ProfScope ps_(PROF_ROWS, st);
            if (int rc = rows_c2c_power<T>(io, lx, nb * H, st)) return rc;
        }
    }
    return 0;
}

// The z-mode chain for the cross spectrum / cross phase of two real fields (config 3).  Pass 1 (the config-2 kernel,
// unchanged) runs once per field and leaves Z1, Z2; the completion tables are built per field; one two-field pass 2
// combines them (and writes the phase next to the cross spectrum when desc.out2 is set).  Per item the workspace holds two
// Z arrays and two sets of column-line tables; the chunk adapts to the workspace the caller sized for the rows-first chain.
// Option cross_z = 0 selects the rows-first two-field chain.
static bool crossz_enabled() { return option(OPT_CROSS_Z) != 0; }
template <typename T> static bool crossz_eligible(const xrftb_spectrum2d_desc& q, int ly, int lx) {
    if (!std::is_same<T, float>::value) return false;
    if (q.keep_half || q.weight_x || q.ramp_y || q.ramp_x) return false;
    if (q.mode == XRFTB_EPI_BINS_CROSS) {   // radial bins of the cross spectrum in the two-field pass 2 (rowszx_bins_kernel)
        if (!q.lut_symmetric || !bins_static_enabled() || !rows_zx_bins_shape_ok(lx - 1, q.ny) || q.nbins > q.nx || q.nbins > 4096) return false;
    } else if (!(q.mode == XRFTB_EPI_CROSS || q.mode == XRFTB_EPI_PHASE)) {
        return false;
    }
    const int C = cols_tile_width<T>(ly, false);
    if (C < 4 || ly <= TypeCfg<T>::LOGE || ly > TypeCfg<T>::MAX_COLS_LOG2 || !rows_z_supported(lx - 1)) return false;
    if ((q.nx * sizeof(T)) % 16 != 0) return false;
    return !q.detrend || (rowline_enabled() && ly >= 3);   // detrend only through the column-line tables
}
template <typename T>
static int spectrum2d_crossz(const xrftb_spectrum2d_desc& q, int ly, int lx, cudaStream_t st) {
    using C_ = cplx<T>;
    const int C = cols_tile_width<T>(ly, false);
    const int H = q.ny / 2 + 1;
    const bool colline = q.detrend != 0;
    const size_t zbytes = (size_t)q.ny * (q.nx / 2) * sizeof(C_);
    const size_t per_item = 2 * (zbytes + (colline ? colline_item_bytes<T>(q.nx) : 0));
    const size_t fixed = colline ? colline_fixed_bytes<T>(q.ny) : 0;
    if (!q.work || q.work_bytes < fixed + per_item) return XRFTB_EWORKSPACE;   // caller falls back to the rows-first chain
    long bchunk = (long)((q.work_bytes - fixed) / per_item);
    if (bchunk > q.batch) bchunk = q.batch;
    { const long nch = (q.batch + bchunk - 1) / bchunk; bchunk = (q.batch + nch - 1) / nch; }
    char* base = reinterpret_cast<char*>(q.work);
    C_* wj = nullptr;
    if (colline) {
        const int My = q.ny / 2;
        wj = reinterpret_cast<C_*>(base);
        double* wrows = reinterpret_cast<double*>(base + align256((size_t)H * 2 * sizeof(C_)));
        double2* wspec = reinterpret_cast<double2*>(reinterpret_cast<char*>(wrows) + align256((size_t)2 * q.ny * sizeof(double)));
        rowline_wj_in_kernel<T><<<(q.ny + 255) / 256, 256, 0, st>>>(reinterpret_cast<const T*>(q.win_y), wrows, q.ny);
        if (int rc = check_launch("rowline_wj_in_kernel")) return rc;
        RowsR2CFused<double> wio{};
        wio.in = wrows; wio.in_row_stride = q.ny; wio.logNy = 0; wio.detrend = 0; wio.moments = nullptr; wio.wy = nullptr; wio.wx = nullptr;
        wio.out = wspec; wio.logC = -1; wio.out_seq_stride = H; wio.rowstats = nullptr;
        if (int rc = rows_r2c<double>(wio, ly - 1, 2, st)) return rc;
        rowline_wj_out_kernel<T><<<(unsigned)((H + 255) / 256), 256, 0, st>>>(wspec, wj, My, H);
        if (int rc = check_launch("rowline_wj_out_kernel")) return rc;
    }
    char* items = base + fixed;
    C_* zf[2] = {reinterpret_cast<C_*>(items), reinterpret_cast<C_*>(items + (size_t)bchunk * zbytes)};
    char* tabs = items + 2 * (size_t)bchunk * zbytes;
    float4* colstats[2] = {nullptr, nullptr};
    C_* ag[2] = {nullptr, nullptr};
    if (colline) {
        for (int f = 0; f < 2; ++f) {
            colstats[f] = reinterpret_cast<float4*>(tabs + (size_t)f * bchunk * colline_item_bytes<T>(q.nx));
            ag[f] = reinterpret_cast<C_*>(colstats[f] + (size_t)bchunk * q.nx);
        }
    }
    const long item = (long)q.ny * q.nx;
    const T* ins[2] = {reinterpret_cast<const T*>(q.in1), reinterpret_cast<const T*>(q.in2)};
    const int tiles_per_item = q.nx / (2 * C);
    const int box_rows = q.ny < 256 ? q.ny : 256;
    const size_t out_elem = q.mode == XRFTB_EPI_PHASE ? sizeof(T) : sizeof(C_);
    const bool bins = q.mode == XRFTB_EPI_BINS_CROSS;
    for (long b0 = 0; b0 < q.batch; b0 += bchunk) {
        const long nb = (q.batch - b0 < bchunk) ? q.batch - b0 : bchunk;
        for (int f = 0; f < 2; ++f) {
            ColsR2CPack<T> io{ins[f] + b0 * item, q.nx, tiles_per_item, q.detrend, nullptr,
                              reinterpret_cast<const T*>(q.win_y), reinterpret_cast<const T*>(q.win_x), zf[f], colstats[f], 0, {}};
            if (!encode_out_tmap(&io.tmap, const_cast<T*>(ins[f] + b0 * item), nb * q.ny, q.nx, 2 * C, box_rows)) {
                set_error("spectrum2d (cross z): tensor map encoding failed"); return XRFTB_ECUDA;
            }
            io.box_rows = box_rows;
            io.zout = zf[f];
            io.ztma = 0;
            if (q.ny >= 512) {
                const int ztma_on = option(OPT_ZTMA);
                const int zbox = q.ny / 4 < 256 ? q.ny / 4 : 256;
                if (ztma_on && encode_out_tmap(&io.ztmap, zf[f], nb * q.ny, q.nx, 2 * C, zbox)) { io.ztma = 1; io.zbox_rows = zbox; }
            }
            {
                ProfScope ps_(PROF_COLS, st);
                if (int rc = cols_r2c_pack<T>(io, ly, C, nb * tiles_per_item, true, st)) return rc;
            }
            if (colline) {
                ProfScope ps_(PROF_MOMENTS, st);
                rowline_fix_kernel<T><<<dim3((unsigned)nb, (unsigned)((q.nx + 1023) / 1024)), 256, 0, st>>>(colstats[f], ag[f], reinterpret_cast<const T*>(q.win_x), q.nx, q.ny, q.detrend);
                if (int rc = check_launch("rowline_fix_kernel")) return rc;
            }
        }
        if (bins) {
            if constexpr (std::is_same<T, float>::value) {
                RowsZCrossBins bio{{zf[0], zf[1], nullptr, nullptr, ly, H, q.shift_y, q.shift_x, (T)q.scale, ag[0], ag[1], wj, nullptr},
                                   q.lut, q.bins + (size_t)b0 * q.nbins * 2, q.nbins, 0, q.ny / 2, nb};
                ProfScope ps_(PROF_ROWS, st);
                for (int part = 0; part < 2; ++part) {   // rows ky in [0, Ny/2) of every plane, then the Nyquist row
                    if (part) { bio.ky0 = q.ny / 2; bio.rows = 1; }
                    const int rc = rows_zx_bins(bio, lx - 1, st);
                    if (rc > 0) { set_error("spectrum2d: radial-bin pass does not cover %d x %d", q.ny, q.nx); return XRFTB_EUNSUPPORTED; }
                    if (rc) return rc;
                }
            }
            continue;
        }
        RowsZCross<T> io{zf[0], zf[1], reinterpret_cast<char*>(q.out) + (size_t)b0 * item * out_elem,
                         q.out2 ? reinterpret_cast<char*>(q.out2) + (size_t)b0 * item * sizeof(T) : nullptr, ly, H, q.shift_y, q.shift_x, (T)q.scale,
                         ag[0], ag[1], wj, nullptr};
        ProfScope ps_(PROF_ROWS, st);
        if (int rc = rows_z_cross<T>(io, lx - 1, nb * H, (q.mode == XRFTB_EPI_CROSS && q.out2) ? (int)EPI_CROSS_AND_PHASE : q.mode, st)) return rc;
    }
    return 0;
}

template <typename T>
static int spectrum2d_impl(const xrftb_spectrum2d_desc& q, cudaStream_t st) {
    using C_ = cplx<T>;
    const int ly = ilog2_exact(q.ny), lx = ilog2_exact(q.nx);
    if (ly < 1 || lx < 2) { set_error("spectrum2d: ny, nx must be powers of two (ny>=2, nx>=4), got %d x %d", q.ny, q.nx); return XRFTB_EUNSUPPORTED; }
    const bool two = (q.mode == XRFTB_EPI_CROSS || q.mode == XRFTB_EPI_PHASE || q.mode == XRFTB_EPI_BINS_CROSS);
    if (two && !q.in2) { set_error("spectrum2d: mode %d needs in2", q.mode); return XRFTB_EINVAL; }
    if (colsfirst_enabled() && colsfirst_eligible<T>(q, ly, lx)) { g_last_path.store(1); return spectrum2d_colsfirst<T>(q, ly, lx, st); }
    if (crossz_enabled() && crossz_eligible<T>(q, ly, lx)) {
        const int rc = spectrum2d_crossz<T>(q, ly, lx, st);
        if (rc != XRFTB_EWORKSPACE) { g_last_path.store(3); return rc; }
    }
    if (q.mode == XRFTB_EPI_CROSS && q.out2) {   // no fused chain for this shape: cross spectrum, then its phase
        xrftb_spectrum2d_desc a = q, b = q;
        a.out2 = nullptr;
        if (int rc = spectrum2d_impl<T>(a, st)) return rc;
        b.mode = XRFTB_EPI_PHASE; b.out = q.out2; b.out2 = nullptr;
        return spectrum2d_impl<T>(b, st);
    }
    g_last_path.store(0);
    const int C = cols_tile_width<T>(ly, two);
    if (C < 1 || ly > TypeCfg<T>::MAX_COLS_LOG2 || lx - 1 > TypeCfg<T>::MAX_ROWS_LOG2) { set_error("spectrum2d: size %d x %d unsupported", q.ny, q.nx); return XRFTB_EUNSUPPORTED; }
    if (q.shift_x && q.keep_half) { set_error("spectrum2d: shift_x is incompatible with keep_half"); return XRFTB_EINVAL; }
    const int fields = two ? 2 : 1;
    const long ntile = (q.nx / 2) / C + 1;
    const size_t per_item = interm_bytes_per_item<T>(q.ny, q.nx, C);
    const size_t mom_bytes = (size_t)q.batch * fields * 4 * sizeof(double);
    // row-line detrend (no moments pass): float32, one field, and the two-rows-per-thread row kernel is the one that runs
    const bool rowline = std::is_same<T, float>::value && q.detrend && !two && rowline_enabled() && rows2_eligible(lx - 1, ly);
    const size_t mom_region = align256(mom_bytes) + (rowline ? rowline_fixed_bytes<T>(q.nx, C) : 0);
    const size_t item_total = per_item * fields + (rowline ? rowline_item_bytes<T>(q.ny) : 0);
    if (!q.work || q.work_bytes < mom_region + item_total) { set_error("spectrum2d: workspace too small (%zu < %zu)", q.work_bytes, mom_region + item_total); return XRFTB_EWORKSPACE; }
    long bchunk = (long)((q.work_bytes - mom_region) / item_total);
    if (bchunk > q.batch) bchunk = q.batch;
    { const long nch = (q.batch + bchunk - 1) / bchunk; bchunk = (q.batch + nch - 1) / nch; }  // balanced chunks
    double* mom = reinterpret_cast<double*>(q.work);
    C_* interm = reinterpret_cast<C_*>(reinterpret_cast<char*>(q.work) + mom_region);
    float4* rowstats = nullptr;
    C_* ag = nullptr;
    C_* wj = nullptr;
    if (rowline) {
        char* fx = reinterpret_cast<char*>(q.work) + align256(mom_bytes);
        const int M = q.nx / 2;
        const size_t ncols = (size_t)(M / C + 1) * C;
        wj = reinterpret_cast<C_*>(fx);
        double* wrows = reinterpret_cast<double*>(fx + align256(ncols * 2 * sizeof(C_)));
        double2* wspec = reinterpret_cast<double2*>(reinterpret_cast<char*>(wrows) + align256((size_t)2 * q.nx * sizeof(double)));
        rowstats = reinterpret_cast<float4*>(reinterpret_cast<char*>(interm) + (size_t)bchunk * per_item * fields);
        ag = reinterpret_cast<C_*>(rowstats + (size_t)bchunk * q.ny);
        // What, Jhat: one fp64 real transform of two rows of length nx, once per call
        rowline_wj_in_kernel<T><<<(q.nx + 255) / 256, 256, 0, st>>>(reinterpret_cast<const T*>(q.win_x), wrows, q.nx);
        if (int rc = check_launch("rowline_wj_in_kernel")) return rc;
        RowsR2CFused<double> wio{};
        wio.in = wrows; wio.in_row_stride = q.nx; wio.logNy = 0; wio.detrend = 0; wio.moments = nullptr; wio.wy = nullptr; wio.wx = nullptr;
        wio.out = wspec; wio.logC = -1; wio.out_seq_stride = M + 1; wio.rowstats = nullptr;
        if (int rc = rows_r2c<double>(wio, lx - 1, 2, st)) return rc;
        rowline_wj_out_kernel<T><<<(unsigned)((ncols + 255) / 256), 256, 0, st>>>(wspec, wj, M, (int)ncols);
        if (int rc = check_launch("rowline_wj_out_kernel")) return rc;
    }
    const long item = (long)q.ny * q.nx;
    const T* ins[2] = {reinterpret_cast<const T*>(q.in1), reinterpret_cast<const T*>(q.in2)};

    if (q.detrend && !rowline) {   // global-plane detrend: one moments reduce per field ahead of the passes
        cudaError_t e = cudaMemsetAsync(mom, 0, mom_bytes, st);
        if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
        for (int f = 0; f < fields; ++f) {
            ProfScope ps_(PROF_MOMENTS, st);
            int chunks = (int)((8L * sm_count() + q.batch - 1) / q.batch);
            if (chunks < 1) chunks = 1;
            if (chunks > q.ny) chunks = q.ny;
            dim3 grid(chunks, (unsigned)q.batch);
            moments_kernel<T><<<grid, 256, 0, st>>>(ins[f], mom + (size_t)f * q.batch * 4, 1, q.ny, q.nx, chunks, q.batch);
            if (int rc = check_launch("moments_kernel")) return rc;
        }
    }
    EpilogueDesc d{};
    const bool phase_plain = (q.mode == XRFTB_EPI_PHASE && !q.ramp_y && !q.ramp_x);  // angle(conj z) = -angle(z)
    const bool mirror_pass = !q.keep_half && (q.mode == XRFTB_EPI_POWER || q.mode == XRFTB_EPI_COMPLEX || q.mode == XRFTB_EPI_CROSS || phase_plain) && q.nx >= 8;
    d.logNx = lx; d.full = q.keep_half ? 0 : (mirror_pass ? 2 : 1); d.shift_y = q.shift_y; d.shift_x = q.shift_x;
    d.scale = q.scale; d.ramp_y = q.ramp_y; d.ramp_x = q.ramp_x; d.weight_x = q.weight_x; d.lut = q.lut; d.nbins = q.nbins; d.lut_symmetric = q.lut_symmetric;
    const long W = q.keep_half ? q.nx / 2 + 1 : q.nx;
    const bool bins_mode = (q.mode == XRFTB_EPI_BINS_POWER || q.mode == XRFTB_EPI_BINS_CROSS);
    const size_t out_elem = (q.mode == XRFTB_EPI_COMPLEX || q.mode == XRFTB_EPI_CROSS) ? sizeof(C_) : sizeof(T);
    for (long b0 = 0; b0 < q.batch; b0 += bchunk) {
        const long nb = (q.batch - b0 < bchunk) ? q.batch - b0 : bchunk;
        for (int f = 0; f < fields; ++f) {
            RowsR2CFused<T> io{};
            io.in = ins[f] + b0 * item; io.in_row_stride = q.nx; io.logNy = ly; io.detrend = q.detrend;
            io.moments = mom + ((size_t)f * q.batch + b0) * 4;
            io.wy = reinterpret_cast<const T*>(q.win_y); io.wx = reinterpret_cast<const T*>(q.win_x);
            io.out = interm + (size_t)f * bchunk * (per_item / sizeof(C_)); io.logC = ilog2_exact(C); io.out_seq_stride = 0;
            io.rowstats = rowline ? rowstats : nullptr;
            ProfScope ps_(PROF_ROWS, st);
            if (int rc = rows_r2c<T>(io, lx - 1, nb * q.ny, st)) return rc;
        }
        if (rowline) {
            ProfScope ps_(PROF_MOMENTS, st);
            rowline_fix_kernel<T><<<dim3((unsigned)nb, (unsigned)((q.ny + 1023) / 1024)), 256, 0, st>>>(rowstats, ag, reinterpret_cast<const T*>(q.win_y), q.ny, q.nx, q.detrend);
            if (int rc = check_launch("rowline_fix_kernel")) return rc;
        }
        d.out = bins_mode ? nullptr : reinterpret_cast<char*>(q.out) + (size_t)b0 * q.ny * W * out_elem;
        d.bins = bins_mode ? q.bins + (size_t)b0 * q.nbins * (q.mode == XRFTB_EPI_BINS_CROSS ? 2 : 1) : nullptr;
        const C_* i1 = interm;
        const C_* i2 = two ? interm + (size_t)bchunk * (per_item / sizeof(C_)) : nullptr;
        d.fix_ag = rowline ? ag : nullptr;
        d.fix_wj = rowline ? wj : nullptr;
        CUtensorMap tmap;
        const CUtensorMap* ptm = nullptr;
        d.use_tma = 0;
        if (q.mode == XRFTB_EPI_POWER && std::is_same<T, float>::value && C >= 8 /* >= 32-byte rows: 16-byte boxes measured slower than LSU stores */ && (mirror_pass || q.keep_half) && !q.weight_x
            && (W * sizeof(float)) % 16 == 0 && q.ny >= 4 && q.nx / 2 >= C /* the Nyquist column starts its own tile: no box wraps around the fftshift */) {
            const int box_rows = q.ny / 2 < 256 ? q.ny / 2 : 256;
            if (encode_out_tmap(&tmap, d.out, nb * q.ny, W, C, box_rows)) { d.use_tma = 1; d.tma_box_rows = box_rows; ptm = &tmap; }
        }
        { ProfScope ps_(PROF_COLS, st);
        if (int rc = cols_fused<T>(q.mode, i1, i2, ly, nb * ntile, (int)ntile, d, ptm, st)) return rc; }
        if (mirror_pass) {
            ProfScope ps_(PROF_MIRROR, st);
            const long nrows = nb * q.ny;
            if (q.mode == XRFTB_EPI_POWER || q.mode == XRFTB_EPI_PHASE)
                mirror_fill_kernel<T, false><<<mirror_grid(nrows), 256, 0, st>>>(d.out, ly, lx, q.shift_y, q.shift_x, nullptr, nullptr, nrows,
                                                                                 q.mode == XRFTB_EPI_PHASE ? 1 : 0);
            else
                mirror_fill_kernel<T, true><<<mirror_grid(nrows), 256, 0, st>>>(d.out, ly, lx, q.shift_y, q.shift_x,
                    reinterpret_cast<const C_*>(q.ramp_y), reinterpret_cast<const C_*>(q.ramp_x), nrows, 0);
            if (int rc = check_launch("mirror_fill_kernel")) return rc;
        }
    }
    return 0;
}

}  // namespace xrftb

using namespace xrftb;

extern "C" {

int xrftb_version(void) { return XRFTB_VERSION; }
long xrftb_launch_count(int reset) {
    long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}
int xrftb_profile_begin(void) {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
    return 0;
}
int xrftb_profile_end(double* ms, long* counts) {
    g_prof_on = false;
    for (int k = 0; k < PROF_NKIND; ++k) { ms[k] = 0.0; counts[k] = 0; }
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.b);
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.kind] += t; counts[r.kind] += 1; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return 0;
}
const char* xrftb_last_error(void) { return g_err; }
int xrftb_set_option(const char* name, int value) {
    const int i = option_index(name);
    if (i < 0) { set_error("unknown option '%s'", name ? name : "(null)"); return XRFTB_EINVAL; }
    std::call_once(g_opt_once, options_init);
    g_opt[i].store(value);
    return 0;
}
int xrftb_get_option(const char* name, int* value) {
    const int i = option_index(name);
    if (i < 0 || !value) { set_error("unknown option '%s'", name ? name : "(null)"); return XRFTB_EINVAL; }
    *value = option((Option)i);
    return 0;
}
int xrftb_spectrum2d_last_path(void) { return g_last_path.load(); }

int xrftb_device_info(int* sms, int* major, int* minor, size_t* smem_optin) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
    if (sms) *sms = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
    return 0;
}

size_t xrftb_fftn_workspace(int dtype, int kind, int ndim, const int64_t* shape, int naxes, const int* axes) {
    if (ndim < 1 || ndim > 8 || naxes < 1 || naxes > ndim || !shape || !axes) return 0;
    return dtype == XRFTB_F32 ? fftn_workspace_impl<float>(kind, ndim, shape, naxes, axes) : fftn_workspace_impl<double>(kind, ndim, shape, naxes, axes);
}

int xrftb_fftn(const void* in, void* out, void* work, size_t work_bytes, int dtype, int kind, int ndim, const int64_t* shape,
               int naxes, const int* axes, void* stream) {
    if (!in || !out || ndim < 1 || ndim > 8 || naxes < 1 || naxes > ndim) { set_error("fftn: bad arguments"); return XRFTB_EINVAL; }
    for (int i = 0; i < naxes; ++i) {
        if (axes[i] < 0 || axes[i] >= ndim) { set_error("fftn: axis out of range"); return XRFTB_EINVAL; }
        for (int j = 0; j < i; ++j)
            if (axes[j] == axes[i]) { set_error("fftn: repeated axis"); return XRFTB_EINVAL; }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == XRFTB_F32) return fftn_impl<float>(in, out, work, work_bytes, kind, ndim, shape, naxes, axes, st);
    if (dtype == XRFTB_F64) return fftn_impl<double>(in, out, work, work_bytes, kind, ndim, shape, naxes, axes, st);
    set_error("fftn: bad dtype %d", dtype);
    return XRFTB_EINVAL;
}

int xrftb_moments(const void* in, double* moments, int dtype, int64_t batch, int64_t n0, int64_t n1, int64_t n2, void* stream) {
    if (!in || !moments || batch < 1 || n0 < 1 || n1 < 1 || n2 < 1) { set_error("moments: bad arguments"); return XRFTB_EINVAL; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(moments, 0, (size_t)batch * 4 * sizeof(double), st);
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
    long rows = n0 * n1;
    int chunks = (int)((8L * sm_count() + batch - 1) / batch);
    if (chunks < 1) chunks = 1;
    if (chunks > rows) chunks = (int)rows;
    dim3 grid(chunks, (unsigned)(batch < 65535 ? batch : 65535));
    if (dtype == XRFTB_F32) moments_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(in), moments, n0, n1, n2, chunks, batch);
    else if (dtype == XRFTB_F64) moments_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(in), moments, n0, n1, n2, chunks, batch);
    else { set_error("moments: bad dtype"); return XRFTB_EINVAL; }
    return check_launch("moments_kernel");
}

int xrftb_detrend_window(const void* in, void* out, const double* moments, int detrend, const void* w0, const void* w1,
                         const void* w2, int dtype, int64_t batch, int64_t n0, int64_t n1, int64_t n2, void* stream) {
    if (!in || !out || (detrend && !moments)) { set_error("detrend_window: bad arguments"); return XRFTB_EINVAL; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long total = batch * n0 * n1 * n2;
    // four elements of the last axis per thread when its rows keep the 16-byte alignment of the vector accesses
    const bool vec = n2 % 4 == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0;
    if (dtype == XRFTB_F32) {
        auto k = vec ? detrend_window_kernel<float, 4> : detrend_window_kernel<float, 1>;
        k<<<ew_grid(vec ? total / 4 : total), 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), moments, detrend,
            reinterpret_cast<const float*>(w0), reinterpret_cast<const float*>(w1), reinterpret_cast<const float*>(w2), n0, n1, n2, total);
    } else if (dtype == XRFTB_F64) {
        auto k = vec ? detrend_window_kernel<double, 4> : detrend_window_kernel<double, 1>;
        k<<<ew_grid(vec ? total / 4 : total), 256, 0, st>>>(reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), moments, detrend,
            reinterpret_cast<const double*>(w0), reinterpret_cast<const double*>(w1), reinterpret_cast<const double*>(w2), n0, n1, n2, total);
    } else { set_error("detrend_window: bad dtype"); return XRFTB_EINVAL; }
    return check_launch("detrend_window_kernel");
}

static int spectral_post_launch(const void* in1, const void* in2, void* out, int dtype, int mode, int64_t batch, int64_t k0, int64_t k1,
                                int64_t k2, int hermitian, int keep_half, const int* shift, const void* const* ramp, const void* weight,
                                double scale, int64_t seg_n, int64_t seg_inner, void* stream) {
    if (!in1 || !out || mode < 0 || mode > XRFTB_EPI_PHASE) { set_error("spectral_post: bad arguments"); return XRFTB_EINVAL; }
    if ((mode == XRFTB_EPI_CROSS || mode == XRFTB_EPI_PHASE) && !in2) { set_error("spectral_post: in2 required"); return XRFTB_EINVAL; }
    if (keep_half && !hermitian) { set_error("spectral_post: keep_half requires hermitian input"); return XRFTB_EINVAL; }
    if (dtype != XRFTB_F32 && dtype != XRFTB_F64) { set_error("spectral_post: bad dtype"); return XRFTB_EINVAL; }
    PostDesc d{};
    d.mode = mode; d.k0 = k0; d.k1 = k1; d.k2 = k2; d.k2in = hermitian ? k2 / 2 + 1 : k2; d.W = keep_half ? k2 / 2 + 1 : k2;
    d.hermitian = (hermitian && !keep_half) ? 1 : 0;
    for (int i = 0; i < 3; ++i) { d.shift[i] = shift ? shift[i] : 0; d.ramp[i] = ramp ? ramp[i] : nullptr; }
    if (keep_half && d.shift[2]) { set_error("spectral_post: cannot shift the half axis"); return XRFTB_EINVAL; }
    d.weight = weight; d.scale = scale;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (seg_n > 0) {
        if (mode == XRFTB_EPI_PHASE || seg_inner < 1 || batch % (seg_n * seg_inner) != 0) { set_error("spectral_post_segmean: bad segment layout or mode"); return XRFTB_EINVAL; }
        d.seg_n = seg_n; d.seg_inner = seg_inner;
        d.scale = scale / (double)seg_n;
        const size_t esz = (dtype == XRFTB_F32 ? sizeof(float) : sizeof(double)) * (mode == XRFTB_EPI_POWER ? 1 : 2);
        cudaError_t e = cudaMemsetAsync(out, 0, (size_t)(batch / seg_n) * k0 * k1 * d.W * esz, st);
        if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
    }
    const long total = batch * k0 * k1 * d.W;
    const bool vec = d.W % 4 == 0 && d.seg_n == 0 && ((uintptr_t)out % 16) == 0;   // four outputs of a row per thread
    if (dtype == XRFTB_F32) {
        auto k = vec ? spectral_post_kernel<float, 4> : spectral_post_kernel<float, 1>;
        k<<<ew_grid(vec ? total / 4 : total), 256, 0, st>>>(reinterpret_cast<const float2*>(in1), reinterpret_cast<const float2*>(in2), out, d, total);
    } else {
        auto k = vec ? spectral_post_kernel<double, 4> : spectral_post_kernel<double, 1>;
        k<<<ew_grid(vec ? total / 4 : total), 256, 0, st>>>(reinterpret_cast<const double2*>(in1), reinterpret_cast<const double2*>(in2), out, d, total);
    }
    return check_launch("spectral_post_kernel");
}

int xrftb_spectral_post(const void* in1, const void* in2, void* out, int dtype, int mode, int64_t batch, int64_t k0, int64_t k1,
                        int64_t k2, int hermitian, int keep_half, const int* shift, const void* const* ramp, const void* weight,
                        double scale, void* stream) {
    return spectral_post_launch(in1, in2, out, dtype, mode, batch, k0, k1, k2, hermitian, keep_half, shift, ramp, weight, scale, 0, 0, stream);
}

int xrftb_spectral_post_segmean(const void* in1, const void* in2, void* out, int dtype, int mode, int64_t batch, int64_t k0, int64_t k1,
                                int64_t k2, int hermitian, int keep_half, const int* shift, const void* const* ramp, const void* weight,
                                double scale, int64_t seg_n, int64_t seg_inner, void* stream) {
    if (seg_n < 1) { set_error("spectral_post_segmean: seg_n < 1"); return XRFTB_EINVAL; }
    return spectral_post_launch(in1, in2, out, dtype, mode, batch, k0, k1, k2, hermitian, keep_half, shift, ramp, weight, scale, seg_n, seg_inner, stream);
}

int xrftb_binned_sum(const void* array, const int32_t* lut, double* bins, int dtype, int is_complex, int64_t batch, int64_t ncell,
                     int nbins, void* stream) {
    if (!array || !lut || !bins || nbins < 1 || batch < 1 || ncell < 1) { set_error("binned_sum: bad arguments"); return XRFTB_EINVAL; }
    if (dtype != XRFTB_F32 && dtype != XRFTB_F64) { set_error("binned_sum: bad dtype"); return XRFTB_EINVAL; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int chunks = (int)((4L * sm_count() + batch - 1) / batch);
    if (chunks < 1) chunks = 1;
    if ((long)chunks * 1024 > ncell) chunks = (int)((ncell + 1023) / 1024);
    const unsigned gy = (unsigned)(batch < 65535 ? batch : 65535);
    dim3 grid(chunks, gy);
    const size_t smem = (size_t)nbins * (is_complex ? 2 : 1) * sizeof(double);
    if (nbins > 2048) {   // the histogram would not fit a CTA's default shared memory: global fp64 atomics
        long gx = (ncell + 255) / 256;
        if (gx > 4L * sm_count()) gx = 4L * sm_count();
        dim3 g2((unsigned)gx, gy);
        if (dtype == XRFTB_F32) {
            if (is_complex) binned_sum_global_kernel<float, true><<<g2, 256, 0, st>>>(reinterpret_cast<const float*>(array), lut, bins, ncell, nbins, batch);
            else binned_sum_global_kernel<float, false><<<g2, 256, 0, st>>>(reinterpret_cast<const float*>(array), lut, bins, ncell, nbins, batch);
        } else {
            if (is_complex) binned_sum_global_kernel<double, true><<<g2, 256, 0, st>>>(reinterpret_cast<const double*>(array), lut, bins, ncell, nbins, batch);
            else binned_sum_global_kernel<double, false><<<g2, 256, 0, st>>>(reinterpret_cast<const double*>(array), lut, bins, ncell, nbins, batch);
        }
        return check_launch("binned_sum_global_kernel");
    }
    if (dtype == XRFTB_F32) {
        if (is_complex) binned_sum_kernel<float, true><<<grid, 256, smem, st>>>(reinterpret_cast<const float*>(array), lut, bins, ncell, nbins, chunks, batch);
        else binned_sum_kernel<float, false><<<grid, 256, smem, st>>>(reinterpret_cast<const float*>(array), lut, bins, ncell, nbins, chunks, batch);
    } else if (dtype == XRFTB_F64) {
        if (is_complex) binned_sum_kernel<double, true><<<grid, 256, smem, st>>>(reinterpret_cast<const double*>(array), lut, bins, ncell, nbins, chunks, batch);
        else binned_sum_kernel<double, false><<<grid, 256, smem, st>>>(reinterpret_cast<const double*>(array), lut, bins, ncell, nbins, chunks, batch);
    } else { set_error("binned_sum: bad dtype"); return XRFTB_EINVAL; }
    return check_launch("binned_sum_kernel");
}

int xrftb_roll_scale(const void* in, void* out, int dtype, int is_complex, int64_t batch, int64_t n0, int64_t n1, int64_t n2,
                     int64_t s0, int64_t s1, int64_t s2, double scale, void* stream) {
    if (!in || !out || in == out) { set_error("roll_scale: bad arguments (in-place is not supported)"); return XRFTB_EINVAL; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long total = batch * n0 * n1 * n2;
    const int width = is_complex ? 2 : 1;
    if (dtype == XRFTB_F32)
        roll_scale_kernel<float><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), n0, n1, n2,
            ((s0 % n0) + n0) % n0, ((s1 % n1) + n1) % n1, ((s2 % n2) + n2) % n2, width, (float)scale, total);
    else if (dtype == XRFTB_F64)
        roll_scale_kernel<double><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), n0, n1, n2,
            ((s0 % n0) + n0) % n0, ((s1 % n1) + n1) % n1, ((s2 % n2) + n2) % n2, width, scale, total);
    else { set_error("roll_scale: bad dtype"); return XRFTB_EINVAL; }
    return check_launch("roll_scale_kernel");
}

size_t xrftb_fft2r_workspace(const xrftb_fft2r_desc* q) {
    if (!q || q->batch < 1 || q->ny < 2 || q->nx < 4) return 0;
    return q->dtype == XRFTB_F32 ? fft2r_workspace_impl<float>(*q) : q->dtype == XRFTB_F64 ? fft2r_workspace_impl<double>(*q) : 0;
}
int xrftb_fft2r(const xrftb_fft2r_desc* q, void* stream) {
    if (!q || !q->in || !q->out || q->batch < 1) { set_error("fft2r: bad descriptor"); return XRFTB_EINVAL; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (q->dtype == XRFTB_F32) return fft2r_impl<float>(*q, st);
    if (q->dtype == XRFTB_F64) return fft2r_impl<double>(*q, st);
    set_error("fft2r: bad dtype %d", q->dtype);
    return XRFTB_EINVAL;
}

int xrftb_permute(const void* in, void* out, int elem_bytes, int ndim, const int64_t* in_shape, const int* perm, const int* flip, void* stream) {
    if (!in || !out || in == out || ndim < 1 || ndim > 6 || !in_shape || !perm) { set_error("permute: bad arguments"); return XRFTB_EINVAL; }
    long istride[6], total = 1;
    { long sacc = 1; for (int a = ndim - 1; a >= 0; --a) { istride[a] = sacc; sacc *= in_shape[a]; total *= in_shape[a]; } }
    PermDesc d{};
    d.ndim = ndim; d.in_off = 0;
    int seen = 0;
    for (int a = 0; a < ndim; ++a) {
        const int p = perm[a];
        if (p < 0 || p >= ndim || (seen >> p) & 1) { set_error("permute: not a permutation"); return XRFTB_EINVAL; }
        seen |= 1 << p;
        d.out_n[a] = in_shape[p];
        d.in_stride[a] = istride[p];
        if (flip && flip[p]) { d.in_off += (in_shape[p] - 1) * istride[p]; d.in_stride[a] = -istride[p]; }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {
        // collapse output axes that are also adjacent in the input; [b][q][p] <- [b][p][q] without flips is a batched transpose
        long on[6], is_[6];
        int m = 0;
        for (int a = 0; a < ndim; ++a) {
            if (d.out_n[a] == 1) continue;
            if (m > 0 && is_[m - 1] == d.in_stride[a] * d.out_n[a] && d.in_stride[a] > 0) { on[m - 1] *= d.out_n[a]; is_[m - 1] = d.in_stride[a]; }
            else { on[m] = d.out_n[a]; is_[m] = d.in_stride[a]; ++m; }
        }
        long Bt = 1, Q = 0, P = 0;
        bool tr = false;
        if (d.in_off == 0 && m == 2 && is_[0] == 1 && is_[1] == on[0]) { Q = on[0]; P = on[1]; tr = true; }
        else if (d.in_off == 0 && m == 3 && is_[1] == 1 && is_[2] == on[1] && is_[0] == on[1] * on[2]) { Bt = on[0]; Q = on[1]; P = on[2]; tr = true; }
        if (tr && P >= 8 && Q >= 8) {
            const long tp = (P + 31) / 32, tq = (Q + 31) / 32, nt = Bt * tp * tq;
            const unsigned grid = (unsigned)(nt < (long)sm_count() * 32 ? nt : (long)sm_count() * 32);
            if (elem_bytes == 4) transpose_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), P, Q, tp, tq, nt);
            else if (elem_bytes == 8) transpose_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), P, Q, tp, tq, nt);
            else if (elem_bytes == 16) transpose_kernel<double2><<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), P, Q, tp, tq, nt);
            else { set_error("permute: element size %d unsupported (4, 8 or 16 bytes)", elem_bytes); return XRFTB_EINVAL; }
            return check_launch("transpose_kernel");
        }
    }
    if (elem_bytes == 4) permute_kernel<float><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), d, total);
    else if (elem_bytes == 8) permute_kernel<double><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), d, total);
    else if (elem_bytes == 16) permute_kernel<double2><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), d, total);
    else { set_error("permute: element size %d unsupported (4, 8 or 16 bytes)", elem_bytes); return XRFTB_EINVAL; }
    return check_launch("permute_kernel");
}

int xrftb_pad(const void* in, void* out, int elem_bytes, int ndim, const int64_t* in_shape, const int64_t* pad_before,
              const int64_t* pad_after, int mode, const void* fill, void* stream) {
    if (!in || !out || ndim < 1 || ndim > 4 || !in_shape || !pad_before || !pad_after || mode < 0 || mode > 4) { set_error("pad: bad arguments"); return XRFTB_EINVAL; }
    PadDesc d{};
    long total = 1;
    for (int a = 0; a < 4; ++a) { d.in_n[a] = 1; d.out_n[a] = 1; d.before[a] = 0; }
    for (int a = 0; a < ndim; ++a) {
        const int k = 4 - ndim + a;
        if (in_shape[a] < 1 || pad_before[a] < 0 || pad_after[a] < 0) { set_error("pad: bad shape or widths"); return XRFTB_EINVAL; }
        d.in_n[k] = in_shape[a]; d.before[k] = pad_before[a]; d.out_n[k] = in_shape[a] + pad_before[a] + pad_after[a];
        total *= d.out_n[k];
    }
    d.mode = mode;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (elem_bytes == 4) {
        float f = 0.f; if (fill) memcpy(&f, fill, 4);
        pad_kernel<float><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), d, f, total);
    } else if (elem_bytes == 8) {
        double f = 0.0; if (fill) memcpy(&f, fill, 8);
        pad_kernel<double><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), d, f, total);
    } else if (elem_bytes == 16) {
        double2 f = make_double2(0.0, 0.0); if (fill) memcpy(&f, fill, 16);
        pad_kernel<double2><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), d, f, total);
    } else { set_error("pad: element size %d unsupported (4, 8 or 16 bytes)", elem_bytes); return XRFTB_EINVAL; }
    return check_launch("pad_kernel");
}

size_t xrftb_spectrum2d_workspace(int dtype, int ny, int nx, int two_fields, int64_t batch_in_flight) {
    const int ly = ilog2_exact(ny);
    if (ly < 1 || nx < 4) return 0;
    const int C = dtype == XRFTB_F32 ? cols_tile_width<float>(ly, two_fields != 0) : cols_tile_width<double>(ly, two_fields != 0);
    if (C < 1) return 0;
    const size_t per_item = dtype == XRFTB_F32 ? interm_bytes_per_item<float>(ny, nx, C) : interm_bytes_per_item<double>(ny, nx, C);
    const int fields = two_fields ? 2 : 1;
    if (batch_in_flight < 1) batch_in_flight = 1;
    // moments region is sized for up to 65536 items x 2 fields; row-line detrend tables (float32, one field) ride along
    const size_t rl_fixed = (dtype == XRFTB_F32 && !two_fields) ? rowline_fixed_bytes<float>(nx, C) : 0;
    const size_t rl_item = (dtype == XRFTB_F32 && !two_fields) ? rowline_item_bytes<float>(ny) : 0;
    size_t item_total = per_item * fields + rl_item;
    if (!two_fields) {   // the columns-first order keeps [ny/2+1][nx] complex per item
        const size_t cf = dtype == XRFTB_F32 ? colsfirst_item_bytes<float>(ny, nx) : colsfirst_item_bytes<double>(ny, nx);
        const size_t cl = dtype == XRFTB_F32 ? cf + colline_item_bytes<float>(nx) : cf;
        if (cl > item_total) item_total = cl;
    }
    const size_t cl_fixed = (dtype == XRFTB_F32 && !two_fields) ? colline_fixed_bytes<float>(ny) : 0;
    return ((size_t)65536 * 2 * 4 * sizeof(double)) + (rl_fixed > cl_fixed ? rl_fixed : cl_fixed) + item_total * (size_t)batch_in_flight + 1024;
}

int xrftb_spectrum2d(const xrftb_spectrum2d_desc* q, void* stream) {
    if (!q || !q->in1 || q->batch < 1) { set_error("spectrum2d: bad descriptor"); return XRFTB_EINVAL; }
    const bool bins_mode = (q->mode == XRFTB_EPI_BINS_POWER || q->mode == XRFTB_EPI_BINS_CROSS);
    if (bins_mode && (!q->lut || !q->bins || q->nbins < 1)) { set_error("spectrum2d: bins mode needs lut/bins/nbins"); return XRFTB_EINVAL; }
    if (!bins_mode && !q->out) { set_error("spectrum2d: out is NULL"); return XRFTB_EINVAL; }
    if (q->out2 && q->mode != XRFTB_EPI_CROSS) { set_error("spectrum2d: out2 (phase) goes with mode CROSS only"); return XRFTB_EINVAL; }
    if (q->dtype != XRFTB_F32 && q->dtype != XRFTB_F64) { set_error("spectrum2d: bad dtype %d", q->dtype); return XRFTB_EINVAL; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the per-call tables (moments region of the workspace, grid.y of the reductions) are sized for 65535 items: longer
    // batches run as consecutive blocks of the same call
    const int64_t kBlock = 65535;
    const size_t esz = q->dtype == XRFTB_F32 ? sizeof(float) : sizeof(double);
    const size_t in_item = (size_t)q->ny * q->nx * esz;
    const size_t W = q->keep_half ? (size_t)q->nx / 2 + 1 : (size_t)q->nx;
    const bool cplx_out = (q->mode == XRFTB_EPI_COMPLEX || q->mode == XRFTB_EPI_CROSS);
    const size_t out_item = (size_t)q->ny * W * esz * (cplx_out ? 2 : 1);
    const size_t bins_item = (size_t)q->nbins * (q->mode == XRFTB_EPI_BINS_CROSS ? 2 : 1);
    for (int64_t b0 = 0; b0 < q->batch; b0 += kBlock) {
        xrftb_spectrum2d_desc d = *q;
        d.batch = q->batch - b0 < kBlock ? q->batch - b0 : kBlock;
        d.in1 = reinterpret_cast<const char*>(q->in1) + (size_t)b0 * in_item;
        if (q->in2) d.in2 = reinterpret_cast<const char*>(q->in2) + (size_t)b0 * in_item;
        if (q->out) d.out = reinterpret_cast<char*>(q->out) + (size_t)b0 * out_item;
        if (q->out2) d.out2 = reinterpret_cast<char*>(q->out2) + (size_t)b0 * (size_t)q->ny * W * esz;
        if (q->bins) d.bins = q->bins + (size_t)b0 * bins_item;
        const int rc = q->dtype == XRFTB_F32 ? spectrum2d_impl<float>(d, st) : spectrum2d_impl<double>(d, st);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
