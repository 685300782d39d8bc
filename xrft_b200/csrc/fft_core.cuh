// xrft_b200 -- in-CTA Stockham auto-sort FFT engine for sm_100a.
//
// Replaces the arithmetic that the reference delegates to numpy's pocketfft
// (np.fft.fftn / rfftn / ifftn / irfftn, call sites xrft/xrft.py:398-404,
// 439-447, 586-591, 612-621).  No cuFFT, no Triton.
//
// Model: a sequence of L = 2^k complex points is owned by NT = L/E threads, each
// holding E points in registers: v[q] = x[u + q*NT].  One stage = (twiddle,
// radix-R butterflies in registers, R <= E) -> scatter to shared memory at the
// Stockham auto-sort position -> gather v[q] = y[u + q*NT].  The register
// residency makes the exchange in-place (one smem buffer, two barriers/stage).
// Radices: E, E, ..., 2^rem (first stage has Ns = 1, i.e. no twiddles).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace xrftb {

template <typename T> struct cplx_of;
template <> struct cplx_of<float> { using type = float2; };
template <> struct cplx_of<double> { using type = double2; };
template <typename T> using cplx = typename cplx_of<T>::type;

template <typename T> __host__ __device__ __forceinline__ cplx<T> mk(T x, T y) { cplx<T> r; r.x = x; r.y = y; return r; }

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
    C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}
template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }
template <typename C> __device__ __forceinline__ C mul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }  // * (-i)
template <typename C> __device__ __forceinline__ C mul_pi(C a) { C r; r.x = -a.y; r.y = a.x; return r; }  // * (+i)
template <typename C, typename T> __device__ __forceinline__ C cscale(C a, T s) { a.x *= s; a.y *= s; return a; }
// a (any complex type) times a plain complex factor w (.x, .y scalars)
template <typename CA, typename CW> __device__ __forceinline__ CA cmulw(CA a, CW w) {
    CA r; r.x = a.x * w.x - a.y * w.y; r.y = a.x * w.y + a.y * w.x; return r;
}

// ---------------------------------------------------------------------------------------------
// register radix kernels, forward sign (exp(-i..)), natural-order output, in place
// ---------------------------------------------------------------------------------------------
template <typename C> __device__ __forceinline__ void fft2(C& a, C& b) { C t = a; a = cadd(t, b); b = csub(t, b); }

template <typename C> __device__ __forceinline__ void fft4(C& v0, C& v1, C& v2, C& v3) {
    C a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_mi(csub(v1, v3));
    v0 = cadd(a0, a2); v1 = cadd(a1, a3); v2 = csub(a0, a2); v3 = csub(a1, a3);
}

template <typename T, typename C> __device__ __forceinline__ void fft8(C* a) {
    const T s = (T)0.70710678118654752440;
    C e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    C o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    fft4(e0, e1, e2, e3);
    fft4(o0, o1, o2, o3);
    C t1; t1.x = (o1.x + o1.y) * s; t1.y = (o1.y - o1.x) * s;      // * w8^1
    C t2 = mul_mi(o2);                                              // * w8^2
    C t3; t3.x = (o3.y - o3.x) * s; t3.y = -(o3.x + o3.y) * s;     // * w8^3
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, t1); a[5] = csub(e1, t1);
    a[2] = cadd(e2, t2); a[6] = csub(e2, t2);
    a[3] = cadd(e3, t3); a[7] = csub(e3, t3);
}

template <typename T, typename C> __device__ __forceinline__ void fft16(C* a) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, s = (T)0.70710678118654752440;
    C e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    fft8<T>(e);
    fft8<T>(o);
    C t;
    // w16^k = (cos(pi k/8), -sin(pi k/8))
    t = o[1]; o[1].x = t.x * c1 + t.y * s1; o[1].y = t.y * c1 - t.x * s1;
    t = o[2]; o[2].x = (t.x + t.y) * s;     o[2].y = (t.y - t.x) * s;
    t = o[3]; o[3].x = t.x * s1 + t.y * c1; o[3].y = t.y * s1 - t.x * c1;
    o[4] = mul_mi(o[4]);
    t = o[5]; o[5].x = -t.x * s1 + t.y * c1; o[5].y = -t.y * s1 - t.x * c1;
    t = o[6]; o[6].x = (t.y - t.x) * s;      o[6].y = -(t.x + t.y) * s;
    t = o[7]; o[7].x = -t.x * c1 + t.y * s1; o[7].y = -t.y * c1 - t.x * s1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = cadd(e[i], o[i]); a[i + 8] = csub(e[i], o[i]); }
}

template <typename T, int R> struct Radix;
template <typename T> struct Radix<T, 1> { template <typename CA> static __device__ __forceinline__ void run(CA*) {} };
template <typename T> struct Radix<T, 2> { template <typename CA> static __device__ __forceinline__ void run(CA* a) { fft2(a[0], a[1]); } };
template <typename T> struct Radix<T, 4> { template <typename CA> static __device__ __forceinline__ void run(CA* a) { fft4(a[0], a[1], a[2], a[3]); } };
template <typename T> struct Radix<T, 8> { template <typename CA> static __device__ __forceinline__ void run(CA* a) { fft8<T>(a); } };
template <typename T> struct Radix<T, 16> { template <typename CA> static __device__ __forceinline__ void run(CA* a) { fft16<T>(a); } };

// ---------------------------------------------------------------------------------------------
// twiddles: table tw[m] = exp(-2 pi i m / L), m in [0, L), host-computed in double.
// Loads w^1, w^2, w^4, w^8 (exact table entries) and derives the rest by <= 2 products.
// ---------------------------------------------------------------------------------------------
template <typename T, int R, typename CA>
__device__ __forceinline__ void apply_twiddles(CA* a, const cplx<T>* __restrict__ tw, int idx) {
    using C = cplx<T>;
    if constexpr (R >= 2) {
        C w1 = __ldg(tw + idx);
        a[1] = cmulw(a[1], w1);
        if constexpr (R >= 4) {
            C w2 = __ldg(tw + 2 * idx);
            C w3 = cmul(w1, w2);
            a[2] = cmulw(a[2], w2);
            a[3] = cmulw(a[3], w3);
            if constexpr (R >= 8) {
                C w4 = __ldg(tw + 4 * idx);
                a[4] = cmulw(a[4], w4);
                a[5] = cmulw(a[5], cmul(w1, w4));
                a[6] = cmulw(a[6], cmul(w2, w4));
                C w7 = cmul(w3, w4);
                a[7] = cmulw(a[7], w7);
                if constexpr (R >= 16) {
                    C w8 = __ldg(tw + 8 * idx);
                    a[8] = cmulw(a[8], w8);
                    a[9] = cmulw(a[9], cmul(w1, w8));
                    a[10] = cmulw(a[10], cmul(w2, w8));
                    a[11] = cmulw(a[11], cmul(w3, w8));
                    C w12 = cmul(w4, w8);
                    a[12] = cmulw(a[12], w12);
                    a[13] = cmulw(a[13], cmul(w1, w12));
                    a[14] = cmulw(a[14], cmul(w2, w12));
                    a[15] = cmulw(a[15], cmul(w3, w12));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// shared-memory view.  Logical point o of sequence-slot `s` lives at
//     base + (pad(o) * SI + s_off)          pad(o) = o + (o >> LOGPAD)
// rows (kernel A):   SI = 1, s_off = seq * seq_stride
// columns (kernel B): SI = C (tile width), s_off = column within the tile
// The pad (one point per first-stage radix) removes the stride-R bank conflicts of the
// Ns = 1 scatter.
// ---------------------------------------------------------------------------------------------
template <int LOGPAD> __host__ __device__ __forceinline__ int padded(int o) { return o + (o >> LOGPAD); }

template <int LOG2L, int LOGE> struct Geometry {
    static constexpr int L = 1 << LOG2L;
    static constexpr int E = 1 << LOGE;
    static constexpr int NT = L / E;
    static constexpr int LOGPAD = LOGE;              // first radix == E
    static constexpr int LPAD = L + (L >> LOGPAD);   // padded length in points
    static constexpr int NSTAGES = (LOG2L + LOGE - 1) / LOGE;
    static constexpr int LOGR_LAST = LOG2L - (NSTAGES - 1) * LOGE;  // in (0, LOGE]
    static constexpr int LOGNS_LAST = (NSTAGES - 1) * LOGE;
};

// v[g + t*G] after the LAST stage is output point final_index(u, g, t)
template <int LOG2L, int LOGE>
__device__ __forceinline__ int final_index(int u, int g, int t) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int R = 1 << G_::LOGR_LAST;
    constexpr int Ns = 1 << G_::LOGNS_LAST;
    int j = u + g * G_::NT;
    int jm = j & (Ns - 1);
    return (j - jm) * R + jm + t * Ns;
}

// Shared-memory addressing is affine in the unrolled indices so every access is [base + immediate]:
//   gather  point u + q*NT          -> pad(u) + q*(NT + NT/PADW)            (NT multiple of PADW)
//   scatter point o0 + t*Ns, Ns>=PADW -> pad(o0) + t*(Ns + Ns/PADW)
//   scatter point j*R + t,   Ns==1   -> j*(R+1) + t                          (R == PADW)
template <typename T, int LOG2L, int LOGE, int LOGNS, int SI>
struct Stages {
    using G_ = Geometry<LOG2L, LOGE>;
    static constexpr int E = G_::E;
    static constexpr int REM = LOG2L - LOGNS;
    static constexpr int LOGR = REM >= LOGE ? LOGE : REM;
    static constexpr int R = 1 << LOGR;
    static constexpr int G = E / R;
    static constexpr int Ns = 1 << LOGNS;
    static constexpr bool LAST = (LOGNS + LOGR == LOG2L);
    static constexpr int PADW = 1 << G_::LOGPAD;

    // sm: smem base of this thread's slot (already offset by the column / sequence); vstride: distance
    // between the NSEQV register-resident sequences of one thread.
    template <int NSEQV>
    static __device__ __forceinline__ void run(cplx<T> (&v)[NSEQV][E], int u, cplx<T>* sm, int vstride,
                                               const cplx<T>* __restrict__ tw) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            int j = u + g * G_::NT;
            int jm = j & (Ns - 1);
#pragma unroll
            for (int s = 0; s < NSEQV; ++s) {
                cplx<T> a[R];
#pragma unroll
                for (int t = 0; t < R; ++t) a[t] = v[s][g + t * G];
                if constexpr (LOGNS > 0) apply_twiddles<T, R>(a, tw, jm * (G_::L / (Ns * R)));
                Radix<T, R>::run(a);
#pragma unroll
                for (int t = 0; t < R; ++t) v[s][g + t * G] = a[t];
            }
        }
        if constexpr (!LAST) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                int j = u + g * G_::NT;
                int jm = j & (Ns - 1);
                int o0 = (j - jm) * R + jm;
                cplx<T>* base;
                int tstep;
                if constexpr (Ns == 1 && R == PADW) { base = sm + (j * (R + 1)) * SI; tstep = SI; }
                else if constexpr (Ns >= PADW) { base = sm + padded<G_::LOGPAD>(o0) * SI; tstep = (Ns + Ns / PADW) * SI; }
                else { base = nullptr; tstep = 0; }
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    cplx<T>* p;
                    if constexpr ((Ns == 1 && R == PADW) || Ns >= PADW) p = base + t * tstep;
                    else p = sm + padded<G_::LOGPAD>(o0 + t * Ns) * SI;
#pragma unroll
                    for (int s = 0; s < NSEQV; ++s) p[s * vstride] = v[s][g + t * G];
                }
            }
            __syncthreads();
            {
                cplx<T>* base = sm + padded<G_::LOGPAD>(u) * SI;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    cplx<T>* p;
                    if constexpr (G_::NT % PADW == 0) p = base + q * ((G_::NT + G_::NT / PADW) * SI);
                    else p = sm + padded<G_::LOGPAD>(u + q * G_::NT) * SI;
#pragma unroll
                    for (int s = 0; s < NSEQV; ++s) v[s][q] = p[s * vstride];
                }
            }
            __syncthreads();
            Stages<T, LOG2L, LOGE, LOGNS + LOGR, SI>::template run<NSEQV>(v, u, sm, vstride, tw);
        }
    }
};

// whole forward FFT of the register-resident sequence(s); results stay in registers, mapped by
// final_index<LOG2L, LOGE>(u, g, t) with g in [0, E/R_last), t in [0, R_last).
template <typename T, int LOG2L, int LOGE, int NSEQV, int SI>
__device__ __forceinline__ void block_fft(cplx<T> (&v)[NSEQV][1 << LOGE], int u, cplx<T>* sm, int vstride,
                                          const cplx<T>* __restrict__ tw) {
    Stages<T, LOG2L, LOGE, 0, SI>::template run<NSEQV>(v, u, sm, vstride, tw);
}

// ---------------------------------------------------------------------------------------------
// asynchronous tile loads: 1-D bulk copies (TMA engine, SASS UBLKCP) completing on an mbarrier.
// One elected thread arms the barrier with the byte count and issues the copies; every thread
// waits on the phase parity before reading the landing buffer.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "XRFTB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra XRFTB_DONE_%=;\n\t"
        "bra XRFTB_WAIT_%=;\n\t"
        "XRFTB_DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 2-D tensor-map tile load (TMA, SASS UTMALDG): box {c0 .. , c1 ..} of the tensor described by `tmap` -> dense smem box
__device__ __forceinline__ void tensor_load_2d_g2s(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Stage driver of the asynchronous column kernel.  Two register-resident sequences per thread
// (columns 2*cg, 2*cg+1 of a C-column tile, CG = C/2 column groups).  Exchange #0 goes through
// the full-width landing buffer `smL` ([pad(o)][C], 128-bit accesses, both sequences at once);
// after its gather the buffer is dead and `hook()` re-arms it with the NEXT tile's bulk load.
// Every later exchange goes through the half-size buffer `smX` ([pad(o)][CG]), one sequence at a
// time, so that the load stays in flight for the rest of the tile.
// ---------------------------------------------------------------------------------------------
template <typename T, int LOG2L, int LOGE, int LOGNS, int C>
struct StagesAsync {
    using G_ = Geometry<LOG2L, LOGE>;
    static constexpr int E = G_::E;
    static constexpr int CG = C / 2;
    static constexpr int REM = LOG2L - LOGNS;
    static constexpr int LOGR = REM >= LOGE ? LOGE : REM;
    static constexpr int R = 1 << LOGR;
    static constexpr int G = E / R;
    static constexpr int Ns = 1 << LOGNS;
    static constexpr bool LAST = (LOGNS + LOGR == LOG2L);
    static constexpr int PADW = 1 << G_::LOGPAD;

    // scatter sequence(s) [s0, s0+ns) of v to `sm` (point stride SI, sequence stride 1)
    template <int SI, int S0, int NS>
    static __device__ __forceinline__ void scatter(cplx<T> (&v)[2][E], int u, cplx<T>* sm) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            int j = u + g * G_::NT;
            int jm = j & (Ns - 1);
            int o0 = (j - jm) * R + jm;
            cplx<T>* base;
            int tstep;
            if constexpr (Ns == 1 && R == PADW) { base = sm + (j * (R + 1)) * SI; tstep = SI; }
            else if constexpr (Ns >= PADW) { base = sm + padded<G_::LOGPAD>(o0) * SI; tstep = (Ns + Ns / PADW) * SI; }
            else { base = nullptr; tstep = 0; }
#pragma unroll
            for (int t = 0; t < R; ++t) {
                cplx<T>* p;
                if constexpr ((Ns == 1 && R == PADW) || Ns >= PADW) p = base + t * tstep;
                else p = sm + padded<G_::LOGPAD>(o0 + t * Ns) * SI;
#pragma unroll
                for (int s = 0; s < NS; ++s) p[s] = v[S0 + s][g + t * G];
            }
        }
    }
    template <int SI, int S0, int NS>
    static __device__ __forceinline__ void gather(cplx<T> (&v)[2][E], int u, cplx<T>* sm) {
        cplx<T>* base = sm + padded<G_::LOGPAD>(u) * SI;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            cplx<T>* p;
            if constexpr (G_::NT % PADW == 0) p = base + q * ((G_::NT + G_::NT / PADW) * SI);
            else p = sm + padded<G_::LOGPAD>(u + q * G_::NT) * SI;
#pragma unroll
            for (int s = 0; s < NS; ++s) v[S0 + s][q] = p[s];
        }
    }

    template <class Hook>
    static __device__ __forceinline__ void run(cplx<T> (&v)[2][E], int u, cplx<T>* smL, cplx<T>* smX,
                                               const cplx<T>* __restrict__ tw, Hook&& hook) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            int j = u + g * G_::NT;
            int jm = j & (Ns - 1);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                cplx<T> a[R];
#pragma unroll
                for (int t = 0; t < R; ++t) a[t] = v[s][g + t * G];
                if constexpr (LOGNS > 0) apply_twiddles<T, R>(a, tw, jm * (G_::L / (Ns * R)));
                Radix<T, R>::run(a);
#pragma unroll
                for (int t = 0; t < R; ++t) v[s][g + t * G] = a[t];
            }
        }
        if constexpr (!LAST) {
            if constexpr (LOGNS == 0) {
                __syncthreads();            // every thread has taken its points of the landed tile out of smL
                scatter<C, 0, 2>(v, u, smL);
                __syncthreads();
                gather<C, 0, 2>(v, u, smL);
                __syncthreads();
                hook();                     // (the issuing thread fences the proxies itself: see cols_async_kernel)
            } else {
                scatter<CG, 0, 1>(v, u, smX);
                __syncthreads();
                gather<CG, 0, 1>(v, u, smX);
                __syncthreads();
                scatter<CG, 1, 1>(v, u, smX);
                __syncthreads();
                gather<CG, 1, 1>(v, u, smX);
                __syncthreads();
            }
            StagesAsync<T, LOG2L, LOGE, LOGNS + LOGR, C>::run(v, u, smL, smX, tw, hook);
        }
    }
};

// hint the L2 to fetch a contiguous chunk (next tile / next rows) while the current one is transformed
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

}  // namespace xrftb
