// xrft_b200 -- mixed-radix Stockham transform for "smooth" lengths n = 2^a 3^b 5^c 7^d that are not powers of two.
//
// The reference hands every length to numpy's pocketfft (np.fft.fftn, xrft/xrft.py:439-447), which factors n into small
// primes; its own tests use 10, 15, 20, 30, 40, 100, 360-day years, 1000, 3650 ... (xrft/tests/test_xrft.py:51-54, 1152-1163).
// Powers of two go through the register-resident engine (fft_core.cuh); smooth lengths come here instead of paying Bluestein's
// two transforms of length >= 2n: one kernel, one read and one write of the data.
//
// One CTA transforms a tile of C sequences of the [A][n][B] view in shared memory, ping-ponging between two buffers: one
// pass per factor r in {16, 8, 4, 2, 9, 3, 25, 5, 7}, one thread per radix-r butterfly (Stockham auto-sort: no bit reversal,
// natural order out; 9 and 25 are 3 x 3 and 5 x 5 butterflies in registers).  Strided sequences (B > 1): the tile is C adjacent
// columns, stored [n][C] so that global and shared accesses run along the columns.  Contiguous sequences (B == 1): the tile is
// C consecutive sequences, stored [C][n + 1].  Tiles arrive by asynchronous global -> shared copies (cp.async); an inverse
// transform is the conjugate of the forward transform of the conjugate (conjugation in the first pass and in the stores).
// Real transforms of even length N along the contiguous axis run as the packed transform of length N / 2 with the split of the
// half spectrum in the stores (R2C) or its merge in the loads (C2R).  Lengths from 6 up to 12800 (float32) / 6400 (float64).
#include <math.h>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include "../../include/xrft_b200.h"
#include "internal.h"

namespace xrftb {

namespace {

struct SmoothPlan { int nfac; int radix[28]; };

// factors in the order of the passes: 16s (float64: 8s), one of 8 / 4 / 2, 9s, 3, 25s (float64: 5s), 5, 7s -- the larger radices
// halve the number of passes
bool factorize(long n, SmoothPlan* p, bool wide) {   // wide (float32): radix 16 and 25 butterflies fit 128 registers
    p->nfac = 0;
    if (n < 2) return false;
    auto take = [&](int r, bool repeat) {
        while (n % r == 0 && p->nfac < 28) {
            p->radix[p->nfac++] = r;
            n /= r;
            if (!repeat) break;
        }
    };
    if (wide) take(16, true); else take(8, true);
    take(8, false); take(4, false); take(2, false);
    take(9, true); take(3, false);
    if (wide) take(25, true); else take(5, true);
    take(5, false);
    take(7, true);
    return n == 1;
}

constexpr int kSmoothMaxDynSmem = 224 * 1024;   // dynamic part; the kernel also has a few hundred bytes of static shared memory
template <typename T> struct SmoothCfg;
// points per buffer: the cap (two buffers within the 227 KB of one CTA) and the tile size aimed at (several CTAs per SM)
template <> struct SmoothCfg<float> { static constexpr int kMaxPoints = 12800, kTilePoints = 4096, kColsWide = 16; };
template <> struct SmoothCfg<double> { static constexpr int kMaxPoints = 6400, kTilePoints = 2048, kColsWide = 8; };

// cos / sin of 2 pi m / R for the odd radices, m in [1, (R-1)/2]
template <int R> __host__ __device__ constexpr double odd_cos(int m) {
    return R == 3 ? -0.5
         : R == 5 ? (m == 1 ? 0.30901699437494742 : -0.80901699437494742)
                  : (m == 1 ? 0.62348980185873353 : m == 2 ? -0.22252093395631440 : -0.90096886790241913);
}
template <int R> __host__ __device__ constexpr double odd_sin(int m) {
    return R == 3 ? 0.86602540378443865
         : R == 5 ? (m == 1 ? 0.95105651629515357 : 0.58778525229247313)
                  : (m == 1 ? 0.78183148246802981 : m == 2 ? 0.97492791218182361 : 0.43388373911755812);
}

// forward DFT of R points in registers, natural order
template <typename T, int R> __device__ __forceinline__ void dft_r(cplx<T>* v) {
    if constexpr (R == 2) fft2(v[0], v[1]);
    else if constexpr (R == 4) fft4(v[0], v[1], v[2], v[3]);
    else {
        // X_k = x_0 + sum_m [cos(2 pi k m / R) (x_m + x_{R-m}) - i sin(2 pi k m / R) (x_m - x_{R-m})], X_{R-k} with + i
        constexpr int H = (R - 1) / 2;
        cplx<T> sm[H], df[H];
#pragma unroll
        for (int m = 1; m <= H; ++m) { sm[m - 1] = cadd(v[m], v[R - m]); df[m - 1] = csub(v[m], v[R - m]); }
        const cplx<T> x0 = v[0];
        cplx<T> tot = x0;
#pragma unroll
        for (int m = 0; m < H; ++m) tot = cadd(tot, sm[m]);
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            cplx<T> a = x0, b = mk<T>(0, 0);
#pragma unroll
            for (int m = 1; m <= H; ++m) {
                const int idx = (k * m) % R;
                const int mm = idx <= H ? idx : R - idx;
                const T c = (T)odd_cos<R>(mm);
                const T s = (T)(idx <= H ? odd_sin<R>(mm) : -odd_sin<R>(mm));
                a.x += c * sm[m - 1].x; a.y += c * sm[m - 1].y;
                b.x += s * df[m - 1].x; b.y += s * df[m - 1].y;
            }
            // -i b = (b.y, -b.x)
            v[k] = mk<T>(a.x + b.y, a.y - b.x);
            v[R - k] = mk<T>(a.x - b.y, a.y + b.x);
        }
        v[0] = tot;
    }
}

// composite radix R1 R2 in registers (Cooley-Tukey inside the butterfly); wR[m] = exp(-2 pi i m / (R1 R2))
template <typename T, int R1, int R2> __device__ __forceinline__ void dft_comp(cplx<T>* v, const cplx<T>* wR) {
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
        cplx<T> a[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) a[n1] = v[n1 * R2 + n2];
        dft_r<T, R1>(a);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) v[k1 * R2 + n2] = (k1 > 0 && n2 > 0) ? cmul(a[k1], wR[k1 * n2]) : a[k1];
    }
    cplx<T> out[R1 * R2];
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        cplx<T> b[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) b[n2] = v[k1 * R2 + n2];
        dft_r<T, R2>(b);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) out[k1 + R1 * k2] = b[k2];
    }
#pragma unroll
    for (int i = 0; i < R1 * R2; ++i) v[i] = out[i];
}
template <typename T, int R> __device__ __forceinline__ void dft_any(cplx<T>* v, const cplx<T>* w9, const cplx<T>* w25) {
    if constexpr (R == 8) fft8<T>(v);
    else if constexpr (R == 16) fft16<T>(v);
    else if constexpr (R == 9) dft_comp<T, 3, 3>(v, w9);
    else if constexpr (R == 25) dft_comp<T, 5, 5>(v, w25);
    else dft_r<T, R>(v);
}

// asynchronous global -> shared copy of one complex point (LDGSTS: no register staging, every copy of the tile in flight at once)
__device__ __forceinline__ void cp_async_point(float2* dst, const float2* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_point(double2* dst, const double2* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// floor(a / d) for 0 <= a < 2^20 through the float reciprocal inv = 1 / d: (a + 1/2) / d is at least 1 / (2 d) away from an
// integer and the rounding error (a / d) 2^-23 stays below that for every quotient that occurs here (< 12800)
__device__ __forceinline__ int fdiv(int a, float inv) { return __float2int_rd(((float)a + 0.5f) * inv); }

// one Stockham pass of radix R over the whole tile: src -> dst (point stride sp, sequence stride sc)
template <typename T, int R>
__device__ __forceinline__ void smooth_pass(const cplx<T>* __restrict__ src, cplx<T>* __restrict__ dst, const cplx<T>* __restrict__ tw,
                                            int n, int Ns, int C, bool contig, int sp, int sc, const cplx<T>* w9, const cplx<T>* w25, bool cj) {
    const int nb = n / R;              // butterflies per sequence
    const int step = nb / Ns;          // n / (Ns R): table stride of this pass
    const int work = nb * C;
    const float inv_nb = 1.0f / (float)nb, inv_c = 1.0f / (float)C, inv_ns = 1.0f / (float)Ns;
    for (int w = threadIdx.x; w < work; w += blockDim.x) {
        int c, j;
        if (contig) { c = fdiv(w, inv_nb); j = w - c * nb; } else { j = fdiv(w, inv_c); c = w - j * C; }
        const int jm = j - fdiv(j, inv_ns) * Ns;
        const cplx<T>* ps = src + c * sc + j * sp;
        cplx<T> v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = ps[t * nb * sp];
        if (cj) {   // first pass of an inverse transform whose tile was copied in unconjugated
#pragma unroll
            for (int t = 0; t < R; ++t) v[t].y = -v[t].y;
        }
        if (Ns > 1) {
            const int ti = jm * step;
#pragma unroll
            for (int t = 1; t < R; ++t) v[t] = cmul(v[t], tw[t * ti]);   // (tw: the CTA's shared-memory copy for short lengths, else global)
        }
        dft_any<T, R>(v, w9, w25);
        cplx<T>* pd = dst + c * sc + ((j - jm) * R + jm) * sp;
#pragma unroll
        for (int t = 0; t < R; ++t) pd[t * Ns * sp] = v[t];
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) smooth_c2c_kernel(const cplx<T>* __restrict__ in, cplx<T>* __restrict__ out, const cplx<T>* __restrict__ tw,
                                                         long A, int n, long B, int C, int contig, int inverse, T scale, const __grid_constant__ SmoothPlan plan,
                                                         long ntiles, long tiles_per_item, int real_mode, const cplx<T>* __restrict__ twN, int tw_smem) {
    constexpr int kLoadAhead = sizeof(T) == 4 ? 8 : 4;   // C2R merge loop: kLoadAhead / 2 row pairs + twiddles in flight per thread
    extern __shared__ __align__(16) unsigned char smooth_smem[];
    __shared__ cplx<T> w9[9], w25[25];   // internal factors of the composite radices, from the length-n table (9 | n, 25 | n)
    if (n % 9 == 0 && threadIdx.x < 9) w9[threadIdx.x] = __ldg(tw + threadIdx.x * (n / 9));
    if (n % 25 == 0 && threadIdx.x >= 32 && threadIdx.x < 57) w25[threadIdx.x - 32] = __ldg(tw + (threadIdx.x - 32) * (n / 25));
    cplx<T>* buf0 = reinterpret_cast<cplx<T>*>(smooth_smem);
    const int sp = contig ? 1 : C, sc = contig ? n + 1 : 1;
    const int tile_elems = contig ? C * (n + 1) : n * C;
    cplx<T>* buf1 = buf0 + tile_elems;
    // short lengths: the butterflies read their twiddles from a copy of the table in shared memory (a global load per factor
    // and butterfly is a full L2 round trip in a kernel with two butterflies per thread and pass)
    const cplx<T>* twp = tw;
    if (tw_smem) {
        cplx<T>* tws = buf1 + tile_elems;
        for (int i = threadIdx.x; i < n; i += blockDim.x) tws[i] = __ldg(tw + i);
        twp = tws;   // published by the barrier behind the first tile's loads
    }
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- load (the inverse transform is the conjugate of the forward transform of the conjugate)
        long a0 = 0, b0 = 0;
        int nc;
        if (contig) {
            a0 = tile * C;
            nc = (int)(A - a0 < C ? A - a0 : C);
            if (real_mode == 2) {
                // C2R of even length N = 2n: the half spectrum X[0..n] of a real row folds into the n-point spectrum Z = E + i O of
                // the packed sequence z[m] = x[2m] + i x[2m+1]:  E = (X[k] + conj X[n-k]) / 2,  O = conj(w_N^k) (X[k] - conj X[n-k]) / 2
                // (the imaginary parts of X[0] and X[n] are ignored, like numpy's irfft)
                const cplx<T>* p = in + a0 * (long)(n + 1);
                const float inv_n = 1.0f / (float)n;
                const int tot = nc * n;
                for (int e0 = threadIdx.x; e0 < tot; e0 += blockDim.x * (kLoadAhead / 2)) {
                    cplx<T> xa[kLoadAhead / 2], xb[kLoadAhead / 2], wv[kLoadAhead / 2];
#pragma unroll
                    for (int u = 0; u < kLoadAhead / 2; ++u) {
                        const int e = e0 + u * blockDim.x;
                        if (e < tot) {
                            const int c = fdiv(e, inv_n), k = e - c * n;
                            xa[u] = p[(long)c * (n + 1) + k]; xb[u] = p[(long)c * (n + 1) + (n - k)]; wv[u] = __ldg(twN + k);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kLoadAhead / 2; ++u) {
                        const int e = e0 + u * blockDim.x;
                        if (e < tot) {
                            const int c = fdiv(e, inv_n), k = e - c * n;
                            cplx<T> xk = xa[u], xm = xb[u];
                            if (k == 0) { xk.y = (T)0; xm.y = (T)0; }
                            const cplx<T> ev = mk<T>((T)0.5 * (xk.x + xm.x), (T)0.5 * (xk.y - xm.y));
                            const cplx<T> dv = mk<T>((T)0.5 * (xk.x - xm.x), (T)0.5 * (xk.y + xm.y));
                            const cplx<T> od = cmulc(dv, wv[u]);
                            buf0[c * sc + k] = mk<T>(ev.x - od.y, -(ev.y + od.x));   // conj(Z): the inverse runs as a conjugated forward transform
                        }
                    }
                }
            } else {
                // asynchronous copies straight into the tile: every point of the tile in flight at once, no register staging (with one
                // register load at a time ncu showed 52 % of the samples in long-scoreboard stalls of this loop and 1 TB/s); an
                // inverse transform conjugates in its first pass instead
                const cplx<T>* p = in + a0 * n;
                const int tot = nc * n;
                const float inv_n = 1.0f / (float)n;
                for (int e = threadIdx.x; e < tot; e += blockDim.x) {
                    const int c = fdiv(e, inv_n), q = e - c * n;
                    cp_async_point(buf0 + c * sc + q, p + e);
                }
                cp_async_wait_all();
            }
        } else {
            a0 = tile / tiles_per_item;
            b0 = (tile - a0 * tiles_per_item) * C;
            nc = (int)(B - b0 < C ? B - b0 : C);
            const cplx<T>* p = in + a0 * n * B + b0;
            const int tot = n * C;
            const float inv_c = 1.0f / (float)C;
            for (int e = threadIdx.x; e < tot; e += blockDim.x) {
                const int q = fdiv(e, inv_c), c = e - q * C;
                if (c < nc) cp_async_point(buf0 + e, p + (long)q * B + c);
                else buf0[e] = mk<T>(0, 0);
            }
            cp_async_wait_all();
        }
        __syncthreads();
        // ---- one pass per factor
        const bool cj0 = inverse && real_mode != 2;   // (the C2R merge stores the conjugate itself)
        cplx<T>* src = buf0;
        cplx<T>* dst = buf1;
        int Ns = 1;
        const int cc = contig ? nc : C;
        for (int f = 0; f < plan.nfac; ++f) {
            const int R = plan.radix[f];
            switch (R) {
                case 2: smooth_pass<T, 2>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0); break;
                case 3: smooth_pass<T, 3>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0); break;
                case 4: smooth_pass<T, 4>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0); break;
                case 5: smooth_pass<T, 5>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0); break;
                case 7: smooth_pass<T, 7>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0); break;
                case 8: smooth_pass<T, 8>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0); break;
                case 9: smooth_pass<T, 9>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0); break;
                default:
                    if constexpr (sizeof(T) == 4) {
                        if (R == 16) smooth_pass<T, 16>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0);
                        else smooth_pass<T, 25>(src, dst, twp, n, Ns, cc, contig, sp, sc, w9, w25, cj0 && f == 0);
                    }
                    break;
            }
            __syncthreads();
            cplx<T>* t = src; src = dst; dst = t;
            Ns *= R;
        }
        // ---- store
        if (contig && real_mode == 1) {
            // R2C of even length N = 2n: the real row was read as n packed points z[m] = x[2m] + i x[2m+1]; with Z its spectrum,
            // X[k] = E + w_N^k O,  E = (Z[k] + conj Z[n-k]) / 2,  O = -i (Z[k] - conj Z[n-k]) / 2,  k = 0 .. n (Z[n] = Z[0])
            cplx<T>* p = out + a0 * (long)(n + 1);
            const float inv_h = 1.0f / (float)(n + 1);
            for (int e = threadIdx.x; e < nc * (n + 1); e += blockDim.x) {
                const int c = fdiv(e, inv_h), k = e - c * (n + 1);
                const cplx<T> zk = src[c * sc + (k == n ? 0 : k)];
                const cplx<T> zm = src[c * sc + ((k == 0 || k == n) ? 0 : n - k)];
                const cplx<T> ev = mk<T>((T)0.5 * (zk.x + zm.x), (T)0.5 * (zk.y - zm.y));
                const cplx<T> dv = mk<T>((T)0.5 * (zk.x - zm.x), (T)0.5 * (zk.y + zm.y));
                const cplx<T> x = cadd(ev, cmul(mk<T>(dv.y, -dv.x), __ldg(twN + k)));
                p[e] = cscale(x, scale);
            }
        } else if (contig) {
            cplx<T>* p = out + a0 * n;
            for (int e = threadIdx.x; e < nc * n; e += blockDim.x) {
                const int c = e / n, q = e - c * n;
                cplx<T> x = src[c * sc + q];
                if (inverse) x.y = -x.y;
                p[e] = cscale(x, scale);
            }
        } else {
            cplx<T>* p = out + a0 * n * B + b0;
            for (int e = threadIdx.x; e < n * C; e += blockDim.x) {
                const int q = e / C, c = e - q * C;
                if (c < nc) {
                    cplx<T> x = src[e];
                    if (inverse) x.y = -x.y;
                    p[q * B + c] = cscale(x, scale);
                }
            }
        }
        __syncthreads();   // the buffers are loaded into again by the next tile
    }
}

std::mutex g_smooth_mu;
std::map<std::tuple<int, int, long>, void*> g_smooth_tw;   // (device, dtype, n) -> exp(-2 pi i m / n), m in [0, n)
bool g_smooth_attr[2][64];                                  // MaxDynamicSharedMemorySize raised on (dtype, device)

template <typename T> const cplx<T>* smooth_table(long n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return nullptr; }
    const int dt = sizeof(T) == 4 ? 0 : 1;
    std::lock_guard<std::mutex> lk(g_smooth_mu);
    auto key = std::make_tuple(dev, dt, n);
    auto it = g_smooth_tw.find(key);
    if (it != g_smooth_tw.end()) return reinterpret_cast<const cplx<T>*>(it->second);
    std::vector<cplx<T>> h(n);
    for (long m = 0; m < n; ++m) {
        const double a = -2.0 * M_PI * (double)m / (double)n;
        h[m].x = (T)cos(a);
        h[m].y = (T)sin(a);
    }
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, n * sizeof(cplx<T>));
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), n * sizeof(cplx<T>), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("smooth twiddle table upload failed: %s", cudaGetErrorString(e)); return nullptr; }
    g_smooth_tw[key] = d;
    return reinterpret_cast<const cplx<T>*>(d);
}

}  // namespace

template <typename T> bool smooth_len_ok(long n) {
    if (n < 2 || (n & (n - 1)) == 0 || n + 1 > SmoothCfg<T>::kMaxPoints) return false;
    SmoothPlan p;
    return factorize(n, &p, sizeof(T) == 4);
}

namespace {
template <typename T>
int smooth_launch(const cplx<T>* src, cplx<T>* dst, long A, long n, long B, int inverse, T scale, cudaStream_t st, int real_mode);
}
template <typename T>
int smooth_c2c(const cplx<T>* src, cplx<T>* dst, long A, long n, long B, int inverse, T scale, cudaStream_t st) {
    return smooth_launch<T>(src, dst, A, n, B, inverse, scale, st, 0);
}
// real transforms of even length N along the contiguous axis through the packed half-length transform (N / 2 smooth):
// one pass over the data, no promotion to complex
template <typename T> bool smooth_real_ok(long N) { return N >= 4 && N % 2 == 0 && smooth_len_ok<T>(N / 2); }
template <typename T> int smooth_r2c(const T* in, cplx<T>* out, long nseq, long N, cudaStream_t st) {
    if (!smooth_real_ok<T>(N)) { set_error("smooth_r2c: length %ld is not covered", N); return XRFTB_EUNSUPPORTED; }
    return smooth_launch<T>(reinterpret_cast<const cplx<T>*>(in), out, nseq, N / 2, 1, 0, (T)1, st, 1);
}
template <typename T> int smooth_c2r(const cplx<T>* in, T* out, long nseq, long N, T scale, cudaStream_t st) {
    if (!smooth_real_ok<T>(N)) { set_error("smooth_c2r: length %ld is not covered", N); return XRFTB_EUNSUPPORTED; }
    return smooth_launch<T>(in, reinterpret_cast<cplx<T>*>(out), nseq, N / 2, 1, 1, scale, st, 2);
}
namespace {
template <typename T>
int smooth_launch(const cplx<T>* src, cplx<T>* dst, long A, long n, long B, int inverse, T scale, cudaStream_t st, int real_mode) {
    SmoothPlan plan;
    if (!smooth_len_ok<T>(n) || !factorize(n, &plan, sizeof(T) == 4)) { set_error("smooth_c2c: length %ld is not covered", n); return XRFTB_EUNSUPPORTED; }
    if (A < 1 || B < 1) return 0;
    const bool contig = (B == 1);
    using Cfg = SmoothCfg<T>;
    long C;
    if (contig) {
        C = Cfg::kTilePoints / (n + 1);
        if (C < 1) C = 1;
        if (C > 64) C = 64;
        if (C > A) C = A;
    } else {
        C = Cfg::kTilePoints / n;
        if (C < 2) C = Cfg::kMaxPoints / n < 2 ? 1 : 2;
        if (C > Cfg::kColsWide) C = Cfg::kColsWide;
        if (C > B) C = B;
    }
    const long tiles_per_item = contig ? 1 : (B + C - 1) / C;
    const long ntiles = contig ? (A + C - 1) / C : A * tiles_per_item;
    const size_t tile_elems = contig ? (size_t)C * (n + 1) : (size_t)n * C;
    const int tw_smem = (size_t)n * sizeof(cplx<T>) <= 16384 ? 1 : 0;   // table copy in shared memory for short lengths
    const size_t smem = 2 * tile_elems * sizeof(cplx<T>) + (tw_smem ? (size_t)n * sizeof(cplx<T>) : 0);
    if (smem > (size_t)kSmoothMaxDynSmem) { set_error("smooth_c2c: tile of length %ld does not fit shared memory", n); return XRFTB_EUNSUPPORTED; }
    const cplx<T>* tw = smooth_table<T>(n);
    if (!tw) return XRFTB_ECUDA;
    const cplx<T>* twN = real_mode ? smooth_table<T>(2 * n) : nullptr;   // exp(-2 pi i k / N) of the real-transform split
    if (real_mode && !twN) return XRFTB_ECUDA;
    auto kern = smooth_c2c_kernel<T>;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_smooth_mu);
        bool& done = g_smooth_attr[sizeof(T) == 4 ? 0 : 1][dev & 63];
        if (!done) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmoothMaxDynSmem);
            if (e != cudaSuccess) { set_error("smooth_c2c: %s", cudaGetErrorString(e)); return XRFTB_ECUDA; }
            done = true;
        }
    }
    long grid = (long)sm_count() * 4;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, 256, smem, st>>>(src, dst, tw, A, (int)n, B, (int)C, contig ? 1 : 0, inverse ? 1 : 0, scale, plan, ntiles, tiles_per_item,
                                            real_mode, twN, tw_smem);
    return check_launch("smooth_c2c_kernel");
}
}  // namespace

template bool smooth_len_ok<float>(long);
template bool smooth_len_ok<double>(long);
template int smooth_c2c<float>(const cplx<float>*, cplx<float>*, long, long, long, int, float, cudaStream_t);
template int smooth_c2c<double>(const cplx<double>*, cplx<double>*, long, long, long, int, double, cudaStream_t);
template bool smooth_real_ok<float>(long);
template bool smooth_real_ok<double>(long);
template int smooth_r2c<float>(const float*, cplx<float>*, long, long, cudaStream_t);
template int smooth_r2c<double>(const double*, cplx<double>*, long, long, cudaStream_t);
template int smooth_c2r<float>(const cplx<float>*, float*, long, long, float, cudaStream_t);
template int smooth_c2r<double>(const cplx<double>*, double*, long, long, double, cudaStream_t);

}  // namespace xrftb
