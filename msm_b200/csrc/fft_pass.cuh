// fft_pass.cuh -- one axis pass of the batched multi-stream complex fp64 FFT, with the point-wise
// operators of the MSM step fused into its load / store.
//
// Replaces (reference, andillio/MSM):
//   utils/fft.rs:6-98          forward / inverse / *_inplace  (ArrayFire fft3 / ifft3 -> cuFFT inside AF)
//   simulation_object.rs:504-516,562-574 (static) / :699-708,:771-780 (expanding)   drift  psi_k *= exp(-i c k^2)
//   simulation_object.rs:535-545 (static) / :726-742 (expanding)                     kick   psi   *= exp(-i kappa phi)
//   simulation_object.rs:1031-1063   calculate_density    rho = A |psi|^2
//   simulation_object.rs:1076-1102   phi_k = c rho_k / k^2, k = 0 -> 0
//   simulation_object.rs:905 / :954  max_all(abs(phi))
//   simulation_object.rs:1249-1293   check_alias
//
// Algorithm: a d-dimensional transform is d passes of length-N 1-D transforms, one per axis.  One CTA owns a
// tile of T = 8 adjacent lines (adjacent along the fastest array dimension for the strided axes, so that every
// global access of a quarter warp is one full 128-byte line) and transforms them with a Stockham decimation
// N = r1*r2*..*rm (radices <= 8, butterflies in registers, E = 8 points per thread), exchanging data between
// stages through shared memory laid out [position][line]: the 8 lanes of a quarter warp always touch 8
// consecutive 16-byte words, so every exchange is bank-conflict free without swizzling.  The CTA then repeats the
// tile for the next stream of its group (two streams share one complex "pair buffer" for the real fields
// rho / phi: rho_a + i rho_b -- the Poisson operator is real and linear, so one complex solve serves two streams).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace msm {

constexpr int MAX_CHUNK = 16;   // streams per launch group

enum LoadOp {
    L_NONE = 0,
    L_DRIFT = 1,
    L_KICK = 2,
    L_KICK_IX = 3,  // L_KICK, but the pair buffer holds phi BEFORE its last inverse pass: that pass runs here, on the
                    // same lines, so phi never travels to HBM in real space
    L_C2R = 4       // half-spectrum line X[0..N-1] (+ X[N] from the Nyquist plane) of a REAL line of 2N values ->
                    // Z[k] = (X[k] + conj X[N-k]) + i conj(W_2N^k) (X[k] - conj X[N-k]); the inverse N-point transform of
                    // Z is then z[j] = x[2j] + i x[2j+1]: the real line, stored as it lies in memory
};
enum StoreOp {
    S_NONE = 0,         // plain store
    S_SCALE = 1,        // * scale
    S_DRIFT = 2,        // * dtab[stream][k]                     (forward passes of the kick->drift transform)
    S_DRIFT_ALIAS = 3,  // S_DRIFT + masked |psi_k|^2 partial sum  (last forward pass)
    S_RHO_KEEP = 4,     // store psi and rho = rho_coef |psi|^2    (last inverse pass of the drift transform)
    S_RHO_ONLY = 5,     // rho only, psi is not written            (potential at time t: only max|phi| is needed)
    S_POISSON = 6,      // * poisson_coef / k^2, DC -> 0           (last forward pass of the Poisson solve)
    S_MAX = 7,          // no store, max|re| and max|im|           (last inverse pass of the dt Poisson solve)
    S_RHO_KEEP_FX = 9,  // S_RHO_KEEP / S_RHO_ONLY, then the FORWARD transform of the pair rho_a + i rho_b on the same
    S_RHO_ONLY_FX = 10, // lines: the pair buffer receives rho after the first pass of the Poisson solve
    S_DRIFT_ALIAS_IZ = 11,  // S_DRIFT_ALIAS, then the INVERSE transform of the stored psi_k into dst2 (first pass of the
                            // next dt-potential)
    S_POISSON_INV = 8,  // S_POISSON, then the INVERSE transform of the same lines, then store: the last forward
                        // and first inverse pass of the Poisson solve share their tile, so the k-space potential
                        // never travels to HBM
    S_R2C = 12          // the line was a REAL line of 2N values read as z[j] = x[2j] + i x[2j+1]: turn Z = F_N[z] into the
                        // half spectrum X[k] = (Z[k] + conj Z[N-k]) / 2 - i W_2N^k (Z[k] - conj Z[N-k]) / 2, k < N, stored
                        // in place; X[N] (real, like X[0]) goes to the Nyquist plane
};

struct PassParams {
    const double2* src;
    double2* dst;
    double2* dst2;                        // S_DRIFT_ALIAS_IZ: scratch slot per local index
    long long src_sstride, dst_sstride;   // elements between stream slots
    int src_by_sid, dst_by_sid;           // slot = stream id (resident array) or local index (scratch)
    int ns, gsz;                          // streams in this launch, streams per CTA group
    int sid[MAX_CHUNK];
    // tile geometry:  offset(tile, l, e) = outer(tile / tiles_inner) + (tile % tiles_inner) * inner_stride
    //                                      + l * lstride + along(e)
    //   outer(o) = (o >> olb) * outer_stride + (o & lomask) * outer_lo ;  along(e) likewise with alb / astride(_lo)
    int axis;                             // 0 = contiguous (x), 1 = stride n (y), 2 = stride n^2 (z)
    int n;
    int tiles_inner;
    long long outer_stride, inner_stride, lstride, astride;
    // blocked slow axis: index o / e splits into (hi, lo) with lo = low `olb` / `alb` bits
    int olb, alb, row_lb;
    long long outer_lo, astride_lo;
    int lvalid;                           // valid lines per tile (1-D grids have a single line)
    // operators
    const double2* twiddle;               // per-stage tables, see plan_tw_offset
    const double2* dtab;                  // [n_streams][N]  per-axis drift factors (scale folded in), by stream id
    double kick[MAX_CHUNK];               // kappa per local index
    double2* pbuf;                        // pair buffers (rho / phi)
    long long p_gstride;
    int p_summed;                         // 1: all streams accumulate into buffer 0 component x
    int rho_accumulate;                   // summed mode: add to what is already in the buffer
    int dtab_shared;                      // all streams of the launch use the drift factors of the first one
    const double* ksq;                    // (k_m)^2 = (m_signed / (n dx))^2, n entries  (utils/fft.rs:100-120)
    double four_pi2, alias_k2_thresh, poisson_coef, rho_coef, scale;
    double* alias_partial;                // [n_streams][ntiles] by stream id
    int ntiles;
    unsigned long long* maxbits;          // [2 * buffers]: bit patterns of non-negative doubles
    // real-field Poisson solve of the summed-density mode (S_R2C / L_C2R and the passes over the half spectrum)
    int nx;                               // extent of the fastest dimension (n, or n / 2 on the half-spectrum grid)
    double k2_fixed;                      // (k_x)^2 of the Nyquist plane (a 2-D grid at fixed k_x), else 0
    double2* nyq;                         // Nyquist plane [row c2][row c1]: X[N] of every line
    const double2* wreal;                 // W_2N^k = exp(-2 pi i k / 2N), k < N
    int tile0, tile_end;                  // this launch covers tiles [tile0, tile_end) (slab-pipelined launches)
    int tiles_per_cta;                    // one-tile kernel: consecutive tiles walked by one CTA (L2 prefetch depth)
    int interleave;                       // W: CTAs b, b+1, .. b+W-1 interleave their tiles (1 = each walks its own run)
    int zero;                             // always 0; only the compiler does not know (see data_dependent)
    int l2_prefetch;                      // pull the next item's tile into L2 while the current one computes
};

// ----------------------------------------------------------------------------------------------------------
// decomposition plans
// ----------------------------------------------------------------------------------------------------------
template <int N> struct Plan;
#define MSM_PLAN(N_, E_, T_, MINB_, NS_, R0, R1, R2, R3)                                   \
    template <> struct Plan<N_> {                                                          \
        static constexpr int E = E_, T = T_, MINB = MINB_, NS = NS_;                       \
        static constexpr int NT = N_ / E_;                                                 \
        static constexpr int THREADS = NT * T_;                                            \
        static constexpr int R[4] = {R0, R1, R2, R3};                                      \
    };
//        N    E  T  minB stages radices
MSM_PLAN(2,    2, 2,  1, 1, 2, 1, 1, 1)
MSM_PLAN(4,    4, 4,  1, 1, 4, 1, 1, 1)
MSM_PLAN(8,    8, 8,  1, 1, 8, 1, 1, 1)
MSM_PLAN(16,   8, 8,  1, 2, 2, 8, 1, 1)
MSM_PLAN(32,   8, 8,  1, 2, 4, 8, 1, 1)
MSM_PLAN(64,   8, 8,  1, 2, 8, 8, 1, 1)
MSM_PLAN(128,  8, 8,  1, 3, 2, 8, 8, 1)
MSM_PLAN(256,  8, 8,  2, 3, 4, 8, 8, 1)
// 512: E = 8 / 2 CTAs per SM measured best (E = 16: -3 %, E = 32 with 3 CTAs: -27 %; profiles/README.md)
#ifndef MSM_MINB512
#define MSM_MINB512 2
#endif
#ifndef MSM_E512
#define MSM_E512 8
#endif
MSM_PLAN(512,  MSM_E512, 8,  MSM_MINB512, 3, 8, 8, 8, 1)
MSM_PLAN(1024, 8, 8,  1, 4, 2, 8, 8, 8)
#undef MSM_PLAN

// Tile height.  Strided axes need T = 8 lines (full 128-byte rows); on the contiguous axis lines are 16*N contiguous
// bytes anyway, so large transforms use T = 2 (measured: 4 -> +3 %, 2 -> +1 % more): a quarter of the shared memory and threads per CTA, 8 CTAs per SM --
// the load / compute / store phases of eight small CTAs interleave better than those of two big ones (matters most for the fused
// two-transform kernels, all of which run on the contiguous axis or tolerate it).
#ifndef MSM_TX
#define MSM_TX 2
#endif
template <int N, bool XL> constexpr int tile_T() { return (XL && N >= 256 && Plan<N>::T == 8) ? MSM_TX : Plan<N>::T; }
template <int N, bool XL, int EV = Plan<N>::E> constexpr int tile_threads() { return (N / EV) * tile_T<N, XL>(); }
// Resident CTAs the small-tile contiguous-axis kernels are compiled for.  Measured (profiles/README.md): 4 CTAs per SM
// with 128 registers beat 8 CTAs with 64 (spills, no room to overlap loads with butterflies): +8 % on the whole step.
#ifndef MSM_XL_MINB
#define MSM_XL_MINB 4
#endif
template <int N, bool XL, int EV = Plan<N>::E> constexpr int tile_minb() {
    if (EV != Plan<N>::E) return Plan<N>::MINB;   // wide variant: half the threads, the same CTAs, twice the registers
    // (a 256-point line has 32 threads, so a T = 2 tile is a 64-thread CTA: twice the CTAs for the same 16 resident warps
    //  at 128 registers -- ncu showed the n/2 = 256-point R2C / C2R kernels of the 512^3 real-field solve at 8 warps / SM)
    return (XL && tile_T<N, XL>() != Plan<N>::T) ? MSM_XL_MINB * (N == 256 ? 2 : 1) : Plan<N>::MINB * (Plan<N>::T / tile_T<N, XL>());
}
// Points per thread of one kernel instance.  E = 8 everywhere (Plan) except where a kernel measured faster with E = 16
// (half the threads, 128 registers, the same two CTAs per SM): the two-transform strided kernel `drift+alias+inv`, which
// at 64 registers cannot hold its 8 store addresses per output array next to the butterflies (+12 %, profiles/README.md).
#ifndef MSM_WIDE_DAI
#define MSM_WIDE_DAI 1
#endif
// `poisson+inv` (two transforms + the k-space multiplier per tile): MSM_WIDE_PI = 1 gives it the same wide plan,
// MSM_PI_NOSTASH = 1 evaluates c / k^2 where it is used instead of parking 8 multipliers per thread in shared memory
#ifndef MSM_WIDE_PI
#define MSM_WIDE_PI 0
#endif
#ifndef MSM_PI_NOSTASH
#define MSM_PI_NOSTASH 1
#endif
// check_alias partial sums: 1 = one per (stream, CTA), reduced after the tile loop; 0 = one per (stream, tile) with a
// CTA-wide barrier at the end of every item (round 1)
#ifndef MSM_ALIAS_PER_CTA
#define MSM_ALIAS_PER_CTA 1
#endif
template <int N, int LOP, int SOP, bool XL> constexpr int plan_E() {
    return (N == 512 && !XL && ((MSM_WIDE_DAI && SOP == 11 /* S_DRIFT_ALIAS_IZ */) || (MSM_WIDE_PI && SOP == 8 /* S_POISSON_INV */)))
               ? 16 : Plan<N>::E;
}

template <int N> constexpr int plan_L(int q) {   // product of radices of stages < q
    int l = 1;
    for (int i = 0; i < q; ++i) l *= Plan<N>::R[i];
    return l;
}
// Twiddles are stored per stage as [k - 1][nu] (nu fastest): entry = W_N^(k * L_q * nu), k = 1..r_q-1, nu < M_q.
// Lanes that differ in nu then read consecutive 16-byte words and lanes that share nu broadcast, for both thread
// mappings (a flat W_N^j table would make the contiguous-axis mapping gather with stride k).
template <int N> constexpr int plan_tw_offset(int q) {
    int off = 0, l = 1;
    for (int i = 0; i < q; ++i) {
        const int r = Plan<N>::R[i];
        off += (r - 1) * (N / (l * r));
        l *= r;
    }
    return off;
}

// ----------------------------------------------------------------------------------------------------------
// complex helpers
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ double2 rot90(double2 a) {
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// 1 / x for normal positive x: hardware seed + two Newton steps (<= 1 ulp; the full IEEE division costs ~3x more)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

static __device__ __noinline__ void sincos_slow(double x, double* s, double* c) { sincos(x, s, c); }

// sin and cos of the kick phase -kappa*phi: the fdlibm minimax kernels on [-pi/4, pi/4] (< 1 ulp).
//   |x| <= pi/4   no reduction, no quadrant logic.  This is the path the integrator takes: dt <= cfl * pi * hbar / max|phi|
//                 (simulation_object.rs:906-909) bounds the kick phase by cfl * pi of the dt-potential, so with cfl <= 0.25
//                 practically every cell of every step lands here; the branch is taken per warp, a warp whose 32 phases are
//                 all small never executes the rest (round 2: the reduction and the selects were 14 of the 40
//                 instructions of every sincos).  Bit-identical to the general path, which gives q = 0, r = x here.
//   |x| <= 1e5    three-term Cody-Waite reduction by pi/2 (exact products via FMA);  larger arguments: library path.
__device__ __forceinline__ void sincos_kernels(double r, double* sn, double* cs) {
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    *sn = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    *cs = fma(z * z, pc, fma(z, -0.5, 1.0));
}
#ifndef MSM_KICK_FASTPATH
#define MSM_KICK_FASTPATH 1
#endif
__device__ __forceinline__ void kick_sincos(double x, double* s, double* c) {
#if MSM_KICK_FASTPATH
    if (fabs(x) <= 0.78539816339744828) {
        sincos_kernels(x, s, c);
        return;
    }
#endif
    if (fabs(x) > 1.0e5) {
        sincos_slow(x, s, c);
        return;
    }
    const double q = rint(x * 0.6366197723675814);
    double r = fma(-q, 1.5707963267948966, x);
    r = fma(-q, 6.123233995736766e-17, r);
    r = fma(-q, -1.4973849048591698e-33, r);
    const int n = (int)q;
    double sn, cs;
    sincos_kernels(r, &sn, &cs);
    const double a = (n & 1) ? cs : sn, b = (n & 1) ? sn : cs;
    *s = (n & 2) ? -a : a;
    *c = ((n + 1) & 2) ? -b : b;
}

// ptr + (bits(x) & zero): the same pointer, but every load through it is data dependent on x for the assembler's
// scheduler as well.  Used to stop it from fetching the (loop-invariant, L1-resident) twiddles of a whole item ahead of
// the tile data and parking them in local memory when registers are capped at 64.
__device__ __forceinline__ const double2* data_dependent(const double2* ptr, double x, int zero) {
    return ptr + (__double2loint(x) & zero);
}

template <int R, bool INV> struct Dft;
template <bool INV> struct Dft<1, INV> {
    static __device__ __forceinline__ void run(double2*) {}
};
template <bool INV> struct Dft<2, INV> {
    static __device__ __forceinline__ void run(double2* v) {
        double2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <bool INV> struct Dft<4, INV> {
    static __device__ __forceinline__ void run(double2* v) {
        double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        double2 t2 = cadd(v[1], v[3]), t3 = rot90<INV>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[2] = csub(t0, t2);
        v[1] = cadd(t1, t3);
        v[3] = csub(t1, t3);
    }
};
template <bool INV> struct Dft<8, INV> {
    static __device__ __forceinline__ void run(double2* v) {
        constexpr double h = 0.70710678118654752440;
        // even / odd radix-4
        double2 e[4] = {v[0], v[2], v[4], v[6]};
        double2 o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4, INV>::run(e);
        Dft<4, INV>::run(o);
        // W8^1 = (1 -+ i)/sqrt2, W8^2 = -+ i, W8^3 = (-1 -+ i)/sqrt2   (upper sign forward)
        double2 o1, o3;
        if (INV) {
            o1 = make_double2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y));
            o3 = make_double2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y));
        } else {
            o1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
            o3 = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
        }
        double2 o2 = rot90<INV>(o[2]);
        v[0] = cadd(e[0], o[0]);
        v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], o1);
        v[5] = csub(e[1], o1);
        v[2] = cadd(e[2], o2);
        v[6] = csub(e[2], o2);
        v[3] = cadd(e[3], o3);
        v[7] = csub(e[3], o3);
    }
};

// ----------------------------------------------------------------------------------------------------------
// block reductions (deterministic order)
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
// max|.| is reduced on the BIT PATTERNS of |x| (non-negative doubles order like unsigned integers, and every NaN
// pattern lies above +Inf): unlike fmax, which drops NaNs, a NaN or Inf anywhere in phi reaches the host, which turns
// it into MSM_E_NAN (utils/grid.rs:66-105 check_complex_for_nans, utils/error.rs:10 RuntimeError::NanOrInf).
__device__ __forceinline__ unsigned long long abs_bits(double x) {
    return (unsigned long long)__double_as_longlong(fabs(x));
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned long long warp_umax(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = umax64(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// ----------------------------------------------------------------------------------------------------------
// the pass kernel
// ----------------------------------------------------------------------------------------------------------
// Stage q (0-based) transforms digit n_q into k_q.  With L = r_0..r_{q-1}, M = N / (L r_q):
//   butterfly b in [0, N / r_q):  kappa = b / M, nu = b % M
//   inputs   position kappa * (M r_q) + n * M + nu          (n = 0..r_q-1)
//   outputs  position (kappa + L k) * M + nu                 (k = 0..r_q-1), twiddle W_N^(k L nu)
// positions of stage 0 inputs are element indices of the line, positions of the last stage outputs are the
// output indices (natural order).
//
// Two thread <-> data mappings (template flag XL):
//   XL = false (strided axes)   tid = t * T + l, exchange buffer [position][line]
//   XL = true  (contiguous axis) tid = l * NT + t, exchange buffer [line][position + position / 8]: a warp then reads
//              32 consecutive elements of ONE line (512 contiguous bytes); the padding keeps the stride-8 gather of
//              the last stage conflict free.
template <int N, bool XL> __device__ __forceinline__ int sm_index(int pos, int l) {
    // contiguous axis: one 16-byte pad after every 8 positions.  The stride-8 gather of the last stage becomes
    // stride 9 (conflict free) and -- unlike an XOR swizzle -- every compile-time part of `pos` stays an additive
    // immediate, so a thread needs one base register per stage instead of one per access.
    if (XL) return l * (N + N / 8) + pos + (pos >> 3);
    return pos * tile_T<N, false>() + l;
}

// Exchange barrier.  Contiguous-axis mapping with NT a multiple of 32: the NT threads of one line are whole warps and
// exchange only among themselves, so each line gets its own named barrier (1 + l) and never waits for the other lines
// of the tile.  Otherwise the lines are interleaved across all warps: CTA-wide barrier.
template <int N, bool XL, int EV> __device__ __forceinline__ void exchange_barrier(int l) {
    if constexpr (XL && ((N / EV) % 32 == 0)) {
        // The barrier id must be an IMMEDIATE: with a register operand ptxas cannot tell which barriers the kernel
        // uses and reserves all 16 for every CTA ("used 16 barriers"), and an SM has 64 barrier slots -> 4 CTAs per SM
        // instead of 8 (ncu: 16 resident warps; profiles/README.md).  The branch is warp-uniform (a warp = one line).
        constexpr int T = tile_T<N, true>();
        static_assert(T <= 8, "one named barrier per line");
#define MSM_LINE_BAR(ID)                                                                      \
    if (T > ID && l == ID) asm volatile("bar.sync %0, %1;" ::"n"(1 + ID), "n"(N / EV) : "memory");
        MSM_LINE_BAR(0) MSM_LINE_BAR(1) MSM_LINE_BAR(2) MSM_LINE_BAR(3)
        MSM_LINE_BAR(4) MSM_LINE_BAR(5) MSM_LINE_BAR(6) MSM_LINE_BAR(7)
#undef MSM_LINE_BAR
    } else {
        __syncthreads();
    }
}

// Twiddle tables are copied to shared memory once per CTA (N * 16 bytes): an L1 hit costs a long-scoreboard wait after
// every butterfly stage, an LDS does not (measured +4.5 % on the whole step; MSM_TW_SMEM=0 reads them through L1)
#ifndef MSM_TW_SMEM
#define MSM_TW_SMEM 1
#endif
constexpr bool kTwSmem = MSM_TW_SMEM != 0;
// TEAM > 0: the transform is run by a TEAM of that many threads inside a larger CTA (fft_tma.cu); its exchanges then
// synchronise on the named barrier `team_bar` instead of the CTA-wide one.
template <int N, bool XL, int EV, int TEAM> __device__ __forceinline__ void stage_barrier(int l, int team_bar) {
    if constexpr (TEAM > 0) asm volatile("bar.sync %0, %1;" ::"r"(team_bar), "n"(TEAM) : "memory");
    else exchange_barrier<N, XL, EV>(l);
}
template <int N, bool INV, bool XL, int Q, int EV, int TEAM = 0>
__device__ __forceinline__ void run_stages(double2 (&v)[EV], double2* sm, int t, int l,
                                           const double2* __restrict__ tw, int team_bar = 0) {
    using PL = Plan<N>;
    constexpr int E = EV, NT = N / EV;
    constexpr int R = PL::R[Q];
    constexpr int L = plan_L<N>(Q);
    constexpr int M = N / (L * R);
    constexpr int NB = E / R;   // butterflies per thread in this stage
#pragma unroll
    for (int c = 0; c < NB; ++c) Dft<R, INV>::run(&v[c * R]);
    if constexpr (Q + 1 < PL::NS) {
        // twiddle + scatter to the exchange buffer
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const int b = t + NT * c;
            const int kappa = b / M, nu = b % M;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                double2 x = v[c * R + k];
                if (k > 0) {
                    double2 w = kTwSmem ? tw[plan_tw_offset<N>(Q) + (k - 1) * M + nu]
                                        : __ldg(&tw[plan_tw_offset<N>(Q) + (k - 1) * M + nu]);
                    if (INV) w.y = -w.y;
                    x = cmul(x, w);
                }
                sm[sm_index<N, XL>((kappa + L * k) * M + nu, l)] = x;
            }
        }
        stage_barrier<N, XL, EV, TEAM>(l, team_bar);
        constexpr int R2 = PL::R[Q + 1];
        constexpr int L2 = L * R;
        constexpr int M2 = N / (L2 * R2);
        constexpr int NB2 = E / R2;
#pragma unroll
        for (int c = 0; c < NB2; ++c) {
            const int b = t + NT * c;
            const int kappa = b / M2, nu = b % M2;
#pragma unroll
            for (int n = 0; n < R2; ++n) v[c * R2 + n] = sm[sm_index<N, XL>(kappa * (M2 * R2) + n * M2 + nu, l)];
        }
        stage_barrier<N, XL, EV, TEAM>(l, team_bar);
        run_stages<N, INV, XL, Q + 1, EV, TEAM>(v, sm, t, l, tw, team_bar);
    }
}

template <int LOP, int SOP> constexpr bool uses_stash() {
    return LOP == L_KICK || LOP == L_KICK_IX || SOP == S_RHO_KEEP || SOP == S_RHO_ONLY ||
           (SOP == S_POISSON_INV && !MSM_PI_NOSTASH) || SOP == S_RHO_KEEP_FX || SOP == S_RHO_ONLY_FX;
}
constexpr bool sop_is_rho(int sop) {
    return sop == S_RHO_KEEP || sop == S_RHO_ONLY || sop == S_RHO_KEEP_FX || sop == S_RHO_ONLY_FX;
}
constexpr bool sop_is_alias(int sop) { return sop == S_DRIFT_ALIAS || sop == S_DRIFT_ALIAS_IZ; }
constexpr bool sop_needs_k2(int sop) { return sop_is_alias(sop) || sop == S_POISSON || sop == S_POISSON_INV; }

// Each thread holds the elements e = t + NT * m (m = 0..E-1) of its line both as inputs of the first stage
// (slot c * R0 + n  <->  m = n * (E / R0) + c) and as outputs of the last stage (slot c * RL + k  <->
// m = c + (E / RL) * k).  Chaining a second transform on the same lines is therefore a compile-time register
// permutation, no data exchange.
template <int N, int EV> struct Slots {
    using PL = Plan<N>;
    static constexpr int E = EV, R0 = PL::R[0], RL = PL::R[PL::NS - 1];
    static constexpr int in_slot_m(int slot) { return (slot % R0) * (E / R0) + slot / R0; }
    static constexpr int out_slot_of_m(int m) { return (m % (E / RL)) * RL + m / (E / RL); }
    static constexpr bool identity = (R0 == E && RL == E);
};
template <int N, int EV> __device__ __forceinline__ void outputs_to_inputs(double2 (&v)[EV]) {
    if constexpr (!Slots<N, EV>::identity) {
        double2 w[EV];
#pragma unroll
        for (int i = 0; i < EV; ++i) w[i] = v[Slots<N, EV>::out_slot_of_m(Slots<N, EV>::in_slot_m(i))];
#pragma unroll
        for (int i = 0; i < EV; ++i) v[i] = w[i];
    }
}
// exchange region in double2 units; single-stage plans (N <= 8) have no exchange, but L_KICK_IX parks phi_a there
template <int N, int LOP, bool XL> constexpr int exchange_elems() {   // (E * threads = N * T for every E)
    // (L_C2R and its counterpart S_R2C pair every element with its mirror image through this buffer, even when the
    //  transform itself has a single stage; the launcher sizes S_R2C kernels with LOP = L_C2R)
    return (Plan<N>::NS > 1 || LOP == L_C2R) ? (XL ? (N + N / 8) : N) * tile_T<N, XL>()
                           : (LOP == L_KICK_IX ? (Plan<N>::E * tile_threads<N, XL>() + 1) / 2 : 0);
}
constexpr bool uses_wreal(int lop, int sop) { return lop == L_C2R || sop == S_R2C; }
template <int LOP, int SOP> constexpr int exch_lop() { return uses_wreal(LOP, SOP) ? (int)L_C2R : LOP; }
// Small read-only tables live in shared memory behind the exchange buffer: an L1 hit still costs a long-scoreboard
// wait at every use (the kernels have no registers to batch such loads), an LDS does not.
//   twiddles [N] double2 | (k_m)^2 [N] double (k^2 consumers) | drift factors [2][N] double2 (drift operators):
//   slot q of the drift table belongs to stream q of the CTA's group; groups of more than two streams (summed
//   coupling) reload slot 0 for every item.
constexpr bool uses_dtab(int lop, int sop) { return lop == L_DRIFT || sop == S_DRIFT || sop_is_alias(sop); }
// Resident CTAs a kernel instance is compiled for.  Contiguous-axis kernels without drift tables need 26-34 KiB of shared
// memory instead of 50-54: MSM_XL_MINB_NODT lets them run more, smaller-register CTAs (A/B: profiles/README.md).
#ifndef MSM_XL_MINB_NODT
#define MSM_XL_MINB_NODT MSM_XL_MINB
#endif
template <int N, int LOP, int SOP, bool XL> constexpr int kernel_minb() {
    constexpr int EV = plan_E<N, LOP, SOP, XL>();
    if (EV == Plan<N>::E && XL && tile_T<N, XL>() != Plan<N>::T && !uses_dtab(LOP, SOP)) return MSM_XL_MINB_NODT * (N == 256 ? 2 : 1);
    return tile_minb<N, XL, EV>();
}
// Register double buffering (contiguous-axis kernels with plain loads): the next item's elements are loaded into a
// second register set while the current item is transformed, so that a CTA's critical path per item is its butterflies
// and exchanges, not load latency + butterflies (these kernels run 4 CTAs of 128 threads per SM with 128 registers).
#ifndef MSM_RP
#define MSM_RP 0
#endif
template <int N, int LOP, int SOP, bool XL> constexpr bool reg_prefetch() {
    return MSM_RP && XL && tile_T<N, XL>() != Plan<N>::T && (LOP == L_NONE || LOP == L_DRIFT) && plan_E<N, LOP, SOP, XL>() == 8;
}
template <int N, int LOP, int SOP> constexpr int table_elems() {   // double2 units
    return (kTwSmem ? N : 0) + (sop_needs_k2(SOP) ? N / 2 + 1 : 0) + (uses_dtab(LOP, SOP) ? 2 * N : 0) +
           (uses_wreal(LOP, SOP) ? N : 0);
}
template <int N, int LOP, int SOP, bool XL> constexpr size_t pass_smem_bytes() {   // E * threads = N * T for every E
    return sizeof(double2) * (exchange_elems<N, exch_lop<LOP, SOP>(), XL>() + table_elems<N, LOP, SOP>()) +
           (uses_stash<LOP, SOP>() ? sizeof(double) * Plan<N>::E * tile_threads<N, XL>() : 0);
}

template <int N, bool INV, int LOP, int SOP, bool XL>
__global__ void __launch_bounds__(tile_threads<N, XL, plan_E<N, LOP, SOP, XL>()>(),
                                  kernel_minb<N, LOP, SOP, XL>()) fft_pass_kernel(const PassParams p) {
    using PL = Plan<N>;
    constexpr int E = plan_E<N, LOP, SOP, XL>(), T = tile_T<N, XL>(), NT = N / E, THREADS = NT * T;
    constexpr int R0 = PL::R[0];
    constexpr int M0 = N / R0;
    constexpr int NB0 = E / R0;
    constexpr int RL = PL::R[PL::NS - 1];
    constexpr int LL = plan_L<N>(PL::NS - 1);
    constexpr int NBL = E / RL;

    extern __shared__ double2 sm[];
    // per-thread stash [E][THREADS] behind the exchange buffer: holds the partner stream's rho / phi so that the
    // pair buffer is always accessed as full 16-byte words
    constexpr int EXCH = exchange_elems<N, exch_lop<LOP, SOP>(), XL>();
    double* stash = reinterpret_cast<double*>(sm + EXCH + table_elems<N, LOP, SOP>());
    __shared__ double red[2][32];
    __shared__ unsigned long long redm[2][32];
    double2* tws = sm + EXCH;
    double* ks = reinterpret_cast<double*>(tws + (kTwSmem ? N : 0));
    double2* dts = tws + (kTwSmem ? N : 0) + (sop_needs_k2(SOP) ? N / 2 + 1 : 0);
    double2* wrs = dts + (uses_dtab(LOP, SOP) ? 2 * N : 0);   // W_2N^k, k < N (real-field passes)
    const double2* tw_base = p.twiddle;
    if constexpr (kTwSmem) {
        for (int i = threadIdx.x; i < N; i += THREADS) tws[i] = p.twiddle[i];
        tw_base = tws;
    }
    if constexpr (sop_needs_k2(SOP)) {
        for (int i = threadIdx.x; i < N; i += THREADS) ks[i] = p.ksq[i];
    }
    if constexpr (uses_wreal(LOP, SOP)) {
        for (int i = threadIdx.x; i < N; i += THREADS) wrs[i] = p.wreal[i];
    }
    if constexpr (kTwSmem || sop_needs_k2(SOP) || uses_wreal(LOP, SOP)) __syncthreads();
    unsigned long long run_max = 0ull, run_max2 = 0ull;   // S_MAX with one buffer per CTA column: reduced once, after the tile loop
    [[maybe_unused]] int item_parity = 0;
    // check_alias partial sums of the (at most two) streams of this CTA's group, carried over all its tiles
    [[maybe_unused]] double alias0 = 0.0, alias1 = 0.0;

    const int tid = threadIdx.x;
    const int l = XL ? tid / NT : tid % T;
    const int t = XL ? tid % NT : tid / T;
    const bool lv = l < p.lvalid;
    const int g = blockIdx.y;

    // offsets inside one grid fit 32 bits (C <= 2^30 elements); only the stream slot needs 64
    auto tile_origin = [&](int tile_) -> int {
        const int o = tile_ / p.tiles_inner;
        return (int)((long long)(o >> p.olb) * p.outer_stride + (long long)(o & ((1 << p.olb) - 1)) * p.outer_lo +
                     (long long)(tile_ % p.tiles_inner) * p.inner_stride);
    };
    auto along = [&](int e) -> long long {
        return (long long)(e >> p.alb) * p.astride + (long long)(e & ((1 << p.alb) - 1)) * p.astride_lo;
    };
    // Every element this thread loads or stores has index e = t + NT * j (j < E) along the pass axis, and NT is a
    // multiple of the slow-axis block (msm_create clamps the block), so its offset is affine in j: a0 + j * astep,
    // 32-bit element units inside one grid (C <= 2^30).  One IMAD + one IMAD.WIDE per access instead of the general
    // (hi, lo) split -- the pass kernels are bound by issue slots and latency, not by HBM (profiles/README.md).
    const int a0 = XL ? t : (int)along(t);
    const int astep = XL ? NT : (int)((long long)(NT >> p.alb) * p.astride);
#ifdef MSM_ADDR_GENERAL   // A/B builds: the general (hi, lo) split, 64-bit, recomputed per access
    auto eoff = [&](int j) -> long long { return along(t + NT * j); };
    auto fresh_offsets = [&](double) {};
#else
    // The store sections re-derive their offsets from a value that depends on the butterfly results (`dep`; p.zero is
    // 0, only the compiler does not know): otherwise the store addresses are formed next to the load addresses, live
    // across the whole transform and get spilled in the register-tight kernels (8 x 64-bit per output array).
    int a0v = a0;
    auto eoff = [&](int j) -> int { return a0v + j * astep; };
#ifdef MSM_NO_FRESH
    auto fresh_offsets = [&](double) {};
#else
    auto fresh_offsets = [&](double dep) { a0v = a0 + (__double2loint(dep) & p.zero); };
#endif
#endif
    // Lines beyond lvalid (1-D grids only: one line per tile) are duplicates of line 0: they load the same elements,
    // compute the same values and store them to the same addresses, and stay out of the reductions (`lv`).  Plain
    // stores are therefore unconditional -- besides saving the predicate, this avoids a code-generation problem seen with
    // CUDA 12.9 ptxas for sm_100a: the one `@!P STG.128` it formed from `if (lv) *o = x` at the end of the item body lost
    // the low word of its data to the next tile's prologue (LDC into the same register).
    const int la = lv ? l : 0;
    // L2 prefetch of one tile: 128-byte lines, thread `tid` takes lines tid, tid + THREADS, ... (strided axes: one line per
    // position, N of them; contiguous axis: the tile is T * N * 16 contiguous bytes = N * T / 8 lines)
    constexpr int PF_LINES = XL ? (N * T) / 8 : N;
    constexpr int PF_ITERS = (PF_LINES + THREADS - 1) / THREADS;
    const int pf_off = XL ? tid * 8 : (int)along(tid < N ? tid : 0);
    const int pf_step = XL ? THREADS * 8 : (int)((long long)(THREADS >> p.alb) * p.astride);
    auto prefetch_tile = [&](const double2* arr_at_origin) {
#pragma unroll
        for (int i = 0; i < PF_ITERS; ++i)
            if (tid + i * THREADS < PF_LINES)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(arr_at_origin + pf_off + i * pf_step));
    };

    // the host makes tiles_per_cta a divisor of tiles_inner: consecutive tiles of one CTA differ by inner_stride
    // Tile walk.  W = p.interleave CTAs with consecutive block indices share a run of W * tiles_per_cta adjacent tiles and
    // take every W-th one: CTAs that were launched together read ADJACENT 128-byte segments of the same rows at about the
    // same time (DRAM row locality), instead of each CTA walking its own 4 adjacent tiles one after the other.
    const int W = p.interleave > 0 ? p.interleave : 1;
    const int first_tile = p.tile0 + (blockIdx.x / W) * (W * p.tiles_per_cta) + blockIdx.x % W;
    const int tstride = W * (int)p.inner_stride, loff = la * (int)p.lstride;
    int origin = tile_origin(first_tile) - tstride;
    constexpr bool RP = reg_prefetch<N, LOP, SOP, XL>();
    [[maybe_unused]] double2 nxt[RP ? E : 1];
    [[maybe_unused]] bool have_nxt = false;
    // the item after (ti_, q_) in this CTA's walk: partner stream of the same tile, else first stream of the next tile
    auto next_item = [&](int ti_, int q_, int origin_, int& ti2, int& q2, int& origin2) -> bool {
        if (q_ + 1 < p.gsz && g * p.gsz + q_ + 1 < p.ns) {
            ti2 = ti_, q2 = q_ + 1, origin2 = origin_;
            return true;
        }
        if (ti_ + 1 < p.tiles_per_cta && first_tile + (ti_ + 1) * W < p.tile_end) {
            ti2 = ti_ + 1, q2 = 0, origin2 = origin_ + tstride;
            return true;
        }
        return false;
    };
    auto item_src = [&](int q_, int origin_) -> const double2* {
        const int li_ = g * p.gsz + q_;
        return p.src + (long long)(p.src_by_sid ? p.sid[li_] : li_) * p.src_sstride + origin_;
    };
    for (int ti = 0; ti < p.tiles_per_cta; ++ti) {
    const int tile = first_tile + ti * W;
    if (tile >= p.tile_end) break;
    origin += tstride;
    const int base = origin + loff;

    // coordinates of this line along the two non-pass axes (only the k^2 consumers need them)
    double k_a = 0.0, k_b = 0.0;
    int c0 = 0, c1 = 0, c2 = 0;
    if constexpr (sop_needs_k2(SOP) || uses_wreal(LOP, SOP)) {
        const int n = p.n;
        if (p.axis == 0) {
            const int line = tile * T + l;   // row index in the (blocked) device layout, see core.cu blk_index
            c1 = (line >> p.row_lb) % n;
            c2 = ((line >> p.row_lb) / n << p.row_lb) + (line & ((1 << p.row_lb) - 1));
        } else if (p.axis == 1) {
            c2 = tile / p.tiles_inner;
            c0 = (tile % p.tiles_inner) * T + l;
        } else {
            const int line = tile * T + l;   // (nx = n except on the half-spectrum grid of the real-field solve)
            c1 = line / p.nx;
            c0 = line % p.nx;
        }
        if (!lv) c0 = c1 = c2 = 0;
    }
    if constexpr (sop_needs_k2(SOP)) {
        // spec_grid sums ((k0^2 + k1^2) + k2^2) * (2 pi)^2 with dim 0 the fastest axis (utils/fft.rs:141-160)
        // one expression for all axes: x + 0 == x exactly, and every term is >= 0.  k2_fixed is (k_x)^2 of the Nyquist
        // plane of the real-field solve (a 2-D grid over (k_y, k_z) at fixed k_x), 0 everywhere else.
        if (p.axis == 2) k_a = ks[c0] + ks[c1], k_b = 0.0;
        else if (p.axis == 1) k_a = p.k2_fixed + ks[c0], k_b = ks[c2];
        else k_a = ks[c1], k_b = ks[c2];
    }
    auto k2_of = [&](int e) -> double {
        // axis 0: (k[e] + k[c1]) + k[c2];  axis 1: (k[c0] + k[e]) + k[c2];  axis 2: (k[c0] + k[c1]) + k[e]
        return ((ks[e] + k_a) + k_b) * p.four_pi2;
    };

    // check_alias mask of this thread's outputs (k^2 > k2_cutoff * k2_max, simulation_object.rs:1259-1280): one bit per
    // output slot, once per tile while registers are free -- the same for every stream of the group
    unsigned amask = 0u;
    if constexpr (sop_is_alias(SOP)) {
#pragma unroll
        for (int c = 0; c < NBL; ++c) {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                if (lv && k2_of(t + NT * c + LL * k) > p.alias_k2_thresh) amask |= 1u << (c * RL + k);
            }
        }
    }
    if constexpr (SOP == S_POISSON_INV && !MSM_PI_NOSTASH) {
        // c / (k^2 n^d) of this thread's 8 outputs, once per tile and while registers are free; parked in the stash
#pragma unroll
        for (int c = 0; c < NBL; ++c) {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const double k2 = k2_of(t + NT * c + LL * k);
                stash[(c * RL + k) * THREADS + tid] = (k2 == 0.0) ? 0.0 : p.poisson_coef * fast_rcp(k2);
            }
        }
    }

    for (int q = 0; q < p.gsz; ++q) {
        const int li = g * p.gsz + q;
        if (li >= p.ns) break;
        const int s = p.sid[li];
        const bool last_of_group = (q + 1 == p.gsz) || (li + 1 >= p.ns);
        // pointers to this thread's line of the item; element j of the thread sits at [a0 + j * astep]
        const double2* __restrict__ src = p.src + (long long)(p.src_by_sid ? s : li) * p.src_sstride + base;
        double2* __restrict__ dst = p.dst + (long long)(p.dst_by_sid ? s : li) * p.dst_sstride + base;
        double2* __restrict__ pb = p.pbuf + (p.p_summed ? 0 : (long long)g * p.p_gstride);
        double2* __restrict__ pbl = pb + base;
        const double2* dt = dts;   // this item's drift factors in shared memory
        if constexpr (uses_dtab(LOP, SOP)) {
            // both streams of the group stay in their slots for all tiles; a group with one common drift coefficient
            // (summed density: one dt for all streams) needs a single table for all items
            const bool resident = p.gsz <= 2 || p.dtab_shared;
            if (resident && !p.dtab_shared) dt = dts + q * N;
            if (!resident || (ti == 0 && (q == 0 || !p.dtab_shared))) {
                if (!resident) __syncthreads();   // the previous item still reads slot 0
                for (int i = threadIdx.x; i < N; i += THREADS) const_cast<double2*>(dt)[i] = __ldg(&p.dtab[(long long)s * N + i]);
                __syncthreads();
            }
        }

        // pull the NEXT item (partner stream of this tile, else first stream of the next tile) into L2 now, so its
        // loads find the data on chip: DRAM stays busy while this item computes
        int ti1 = 0, q1 = 0, origin1 = 0, ti2 = 0, q2 = 0, origin2 = 0;
        const bool has1 = next_item(ti, q, origin, ti1, q1, origin1);
        if (p.l2_prefetch && (p.tiles_per_cta > 1 || p.gsz > 1)) {
            if constexpr (RP) {   // the next item goes to registers (below): L2 gets the one after it
                if (has1 && next_item(ti1, q1, origin1, ti2, q2, origin2)) prefetch_tile(item_src(q2, origin2));
            } else if (has1) {
                prefetch_tile(item_src(q1, origin1));
                if ((LOP == L_KICK || LOP == L_KICK_IX) && ti1 != ti) prefetch_tile(pb + origin1);
            }
        }

        double2 v[E];
        a0v = a0;
        if constexpr (LOP == L_KICK_IX) {
            // last inverse pass of the Poisson solve on this tile of the pair buffer, once per unit; phi_a is parked in
            // the (still idle) exchange buffer, phi_b in the stash, each thread in its own slots
            if (q == 0) {
#pragma unroll
                for (int c = 0; c < NB0; ++c) {
#pragma unroll
                    for (int n = 0; n < R0; ++n) {
                        v[c * R0 + n] = pbl[eoff(c + n * NB0)];
                    }
                }
                run_stages<N, true, XL, 0>(v, sm, t, l, data_dependent(tw_base, v[0].x, p.zero));
                outputs_to_inputs<N>(v);
                __syncthreads();   // phi_a slots span the whole exchange buffer: all lines must be done with it
                double* phia = reinterpret_cast<double*>(sm);
#pragma unroll
                for (int i = 0; i < E; ++i) {
                    phia[i * THREADS + tid] = v[i].x;
                    stash[i * THREADS + tid] = p.p_summed ? v[i].x : v[i].y;
                }
            }
        }
        // ---- load (stage-0 input order): all global loads first, operators afterwards ----
        if constexpr (RP) {
            if (have_nxt) {
#pragma unroll
                for (int i = 0; i < E; ++i) v[i] = nxt[i];
            } else {
#pragma unroll
                for (int c = 0; c < NB0; ++c) {
#pragma unroll
                    for (int n = 0; n < R0; ++n) v[c * R0 + n] = src[eoff(c + n * NB0)];
                }
            }
            have_nxt = has1;
            if (has1) {   // in flight while this item is transformed
                const double2* __restrict__ nsrc = item_src(q1, origin1) + loff;
#pragma unroll
                for (int c = 0; c < NB0; ++c) {
#pragma unroll
                    for (int n = 0; n < R0; ++n) nxt[c * R0 + n] = nsrc[a0 + (c + n * NB0) * astep];
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < NB0; ++c) {
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    v[c * R0 + n] = src[eoff(c + n * NB0)];
                }
            }
        }
        if constexpr (LOP == L_DRIFT) {
#pragma unroll
            for (int c = 0; c < NB0; ++c) {
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int e = n * M0 + t + NT * c;
                    const double2 w = dt[e];
                    v[c * R0 + n] = cmul(v[c * R0 + n], w);
                }
            }
        }
        if constexpr (LOP == L_C2R) {
            // half spectrum of a real line of 2N values -> spectrum of z[j] = x[2j] + i x[2j+1]; every element meets its
            // mirror image X[N - k] through the (idle) exchange buffer, X[N] comes from the Nyquist plane
            double2 xn = make_double2(0.0, 0.0);
            if (t == 0) xn = p.nyq[((long long)li * p.n + c2) * p.n + c1];
#pragma unroll
            for (int c = 0; c < NB0; ++c) {
#pragma unroll
                for (int n = 0; n < R0; ++n) sm[sm_index<N, XL>(n * M0 + t + NT * c, l)] = v[c * R0 + n];
            }
            exchange_barrier<N, XL, E>(l);
#pragma unroll
            for (int c = 0; c < NB0; ++c) {
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int e = n * M0 + t + NT * c;
                    double2 pr = sm[sm_index<N, XL>((N - e) & (N - 1), l)];
                    if (c == 0 && n == 0 && t == 0) pr = xn;
                    const double2 x = v[c * R0 + n], w = wrs[e];
                    const double2 a = make_double2(x.x + pr.x, x.y - pr.y), b = make_double2(x.x - pr.x, x.y + pr.y);
                    const double2 mm = cmul(make_double2(w.x, -w.y), b);
                    v[c * R0 + n] = make_double2(a.x - mm.y, a.y + mm.x);
                }
            }
            exchange_barrier<N, XL, E>(l);   // everyone has read its mirror image before the first stage scatters
        }
        if constexpr (LOP == L_KICK_IX) {
            const double* phia = reinterpret_cast<const double*>(sm);
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const double ph = (q == 0) ? phia[i * THREADS + tid] : stash[i * THREADS + tid];
                double sn, cs;
                kick_sincos(-p.kick[li] * ph, &sn, &cs);
                v[i] = cmul(v[i], make_double2(cs, sn));
            }
            if (q == 0) __syncthreads();   // phi_a lives in the exchange buffer: everyone reads before anyone scatters
        }
        if constexpr (LOP == L_KICK) {
            // psi *= exp(-i kappa phi)    (simulation_object.rs:535-545); phi_a + i phi_b is read once per pair
#pragma unroll
            for (int c = 0; c < NB0; ++c) {
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    double ph;
                    double* slot = &stash[(c * R0 + n) * THREADS + tid];
                    if (q == 0) {
                        if (p.p_summed == 2) {   // shared potential as a REAL plane (summed density, real-field solve)
                            ph = reinterpret_cast<const double*>(pb)[base + eoff(c + n * NB0)];
                            if (!last_of_group) *slot = ph;
                        } else {
                            const double2 pp = pbl[eoff(c + n * NB0)];
                            ph = pp.x;
                            if (!last_of_group) *slot = p.p_summed ? pp.x : pp.y;
                        }
                    } else {
                        ph = *slot;
                    }
                    double sn, cs;
                    kick_sincos(-p.kick[li] * ph, &sn, &cs);
                    v[c * R0 + n] = cmul(v[c * R0 + n], make_double2(cs, sn));
                }
            }
        }

        const double2* tw1 = data_dependent(tw_base, v[0].x, p.zero);
        run_stages<N, INV, XL, 0>(v, sm, t, l, tw1);

        if constexpr (SOP == S_POISSON_INV) {
            // phi_k = c rho_k / k^2 (DC -> 0) on the forward outputs, which each thread holds at e = t + NT * m:
            // exactly the element set the inverse transform's first stage wants, in a different register order
#pragma unroll
            for (int c = 0; c < NBL; ++c) {
#pragma unroll
                for (int k = 0; k < RL; ++k) {
#if MSM_PI_NOSTASH
                    const double k2 = k2_of(t + NT * c + LL * k);
                    const double m = (k2 == 0.0) ? 0.0 : p.poisson_coef * fast_rcp(k2);
#else
                    const double m = stash[(c * RL + k) * THREADS + tid];
#endif
                    v[c * RL + k].x *= m;
                    v[c * RL + k].y *= m;
                }
            }
            outputs_to_inputs<N>(v);
            run_stages<N, !INV, XL, 0>(v, sm, t, l, data_dependent(tw_base, v[0].x, p.zero));
        }

        if constexpr (SOP == S_R2C) {
            // Z[k] of this line into the (idle) exchange buffer: every output needs Z[N - k] as well
#pragma unroll
            for (int c = 0; c < NBL; ++c) {
#pragma unroll
                for (int k = 0; k < RL; ++k) sm[sm_index<N, XL>(t + NT * c + LL * k, l)] = v[c * RL + k];
            }
            exchange_barrier<N, XL, E>(l);
        }
        // ---- store (last-stage output order) ----
        fresh_offsets(v[0].x);
        double acc = 0.0;
        unsigned long long mx = 0ull, mx2 = 0ull;
#pragma unroll
        for (int c = 0; c < NBL; ++c) {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const int e = t + NT * c + LL * k;
                const auto off = eoff(c + k * NBL);
                double2 x = v[c * RL + k];
                if constexpr (SOP == S_SCALE) {
                    x.x *= p.scale;
                    x.y *= p.scale;
                }
                if constexpr (SOP == S_R2C) {
                    const double2 zp = sm[sm_index<N, XL>((N - e) & (N - 1), l)], w = wrs[e];
                    const double2 a = make_double2(x.x + zp.x, x.y - zp.y), b = make_double2(x.x - zp.x, x.y + zp.y);
                    const double2 mm = cmul(w, b);
                    if (c == 0 && k == 0 && t == 0) {   // k = 0: X[0] = Re Z + Im Z and X[N] = Re Z - Im Z, both real
                        if (lv) p.nyq[((long long)li * p.n + c2) * p.n + c1] = make_double2(x.x - x.y, 0.0);
                        x = make_double2(x.x + x.y, 0.0);
                    } else {
                        x = make_double2(0.5 * (a.x + mm.y), 0.5 * (a.y - mm.x));
                    }
                }
                if constexpr (SOP == S_DRIFT || sop_is_alias(SOP)) {
                    const double2 w = dt[e];
                    x = cmul(x, w);
                }
                if constexpr (sop_is_alias(SOP)) {
                    // check_alias: sum |psi_k|^2 where k^2 > k2_cutoff * k2_max  (simulation_object.rs:1259-1280)
#ifdef MSM_NO_AMASK
                    if (lv && k2_of(e) > p.alias_k2_thresh) acc += x.x * x.x + x.y * x.y;
#else
                    if (amask & (1u << (c * RL + k))) acc += x.x * x.x + x.y * x.y;
#endif
                }
                if constexpr (SOP == S_POISSON) {
                    // phi_k = c rho_k / k^2, 0/0 at k = 0 replaced by 0  (simulation_object.rs:1076-1102)
                    const double k2 = k2_of(e);
                    const double m = (k2 == 0.0) ? 0.0 : p.poisson_coef * fast_rcp(k2);
                    x.x *= m;
                    x.y *= m;
                }
                if constexpr (SOP == S_MAX) {
                    mx = umax64(mx, abs_bits(x.x));
                    mx2 = umax64(mx2, abs_bits(x.y));
                }
                if constexpr (sop_is_rho(SOP)) {
                    // rho = A real(psi conj(psi))   (simulation_object.rs:1051-1062)
                    constexpr bool FX = (SOP == S_RHO_KEEP_FX || SOP == S_RHO_ONLY_FX);
                    const double rho = p.rho_coef * (x.x * x.x + x.y * x.y);
                    double* slot = &stash[(c * RL + k) * THREADS + tid];
                    double2 pair = make_double2(0.0, 0.0);
                    if (!p.p_summed) {
                        // pair buffer rho_a + i rho_b: the first stream parks its value, the second completes the pair
                        if (!last_of_group) *slot = rho;
                        else pair = (q == 0) ? make_double2(rho, 0.0) : make_double2(*slot, rho);
                    } else {
                        const double sum = (q == 0) ? rho : *slot + rho;
                        if (!last_of_group) *slot = sum;
                        else if (p.p_summed == 2) {   // the summed density is kept as a REAL plane (8 B / cell)
                            double* rp = reinterpret_cast<double*>(pb) + base;
                            if (lv) rp[off] = p.rho_accumulate ? rp[off] + sum : sum;
                        }
                        else pair = make_double2(p.rho_accumulate ? pbl[off].x + sum : sum, 0.0);   // (stored by lv lines only)
                    }
                    if (SOP != S_RHO_ONLY && SOP != S_RHO_ONLY_FX) {
                        dst[off] = x;
                    }
                    if (FX) v[c * RL + k] = pair;               // transformed below, after psi has been stored
                    else if (last_of_group && lv && p.p_summed != 2) pbl[off] = pair;
                }
                if constexpr (!sop_is_rho(SOP) && SOP != S_MAX) {
                    dst[off] = x;
                }
                if constexpr (SOP == S_DRIFT_ALIAS_IZ) v[c * RL + k] = x;   // psi_k as stored; transformed back below
            }
        }

        if constexpr (SOP == S_R2C) exchange_barrier<N, XL, E>(l);   // mirror images read before the next item scatters
        if constexpr (SOP == S_RHO_KEEP_FX || SOP == S_RHO_ONLY_FX) {
            // first (x) pass of the Poisson solve on the finished pair rho_a + i rho_b, same lines: forward transform
            if (last_of_group) {
                outputs_to_inputs<N>(v);
                run_stages<N, false, XL, 0>(v, sm, t, l, data_dependent(tw_base, v[0].x, p.zero));
                fresh_offsets(v[0].x);
#pragma unroll
                for (int c = 0; c < NBL; ++c) {
#pragma unroll
                    for (int k = 0; k < RL; ++k) {
                        pbl[eoff(c + k * NBL)] = v[c * RL + k];
                    }
                }
            }
        }
        if constexpr (SOP == S_DRIFT_ALIAS_IZ) {
            // first pass of the next dt-potential: inverse transform of the psi_k just stored, into the scratch slot
            double2* __restrict__ d2 = p.dst2 + (long long)li * p.dst_sstride + base;
            outputs_to_inputs<N>(v);
            run_stages<N, true, XL, 0>(v, sm, t, l, data_dependent(tw_base, v[0].x, p.zero));
            fresh_offsets(v[0].x);
#pragma unroll
            for (int c = 0; c < NBL; ++c) {
#pragma unroll
                for (int k = 0; k < RL; ++k) {
                    d2[eoff(c + k * NBL)] = v[c * RL + k];
                }
            }
        }
#if MSM_ALIAS_PER_CTA
        if constexpr (sop_is_alias(SOP)) {
            // alias passes run in groups of at most two streams: a thread keeps one running sum per stream over all tiles
            // of its CTA; the block reduction happens once, after the tile loop (no CTA-wide barrier per item)
            if (q == 0) alias0 += acc;
            else alias1 += acc;
        }
#else
        if constexpr (sop_is_alias(SOP)) {
            // one partial per (stream, tile), summed in a fixed order -> deterministic.  `red` is double buffered by
            // item parity: by the time a parity is reused the stage barriers of the item in between have passed.
            acc = warp_sum(acc);
            if ((tid & 31) == 0) red[item_parity][tid >> 5] = acc;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < (THREADS + 31) / 32; ++w) tot += red[item_parity][w];
                p.alias_partial[(long long)s * p.ntiles + tile] = tot;
            }
            item_parity ^= 1;
        }
#endif
        if constexpr (SOP == S_MAX) {
            run_max = umax64(run_max, mx);
            run_max2 = umax64(run_max2, mx2);
            if (p.gsz > 1) {   // several buffers per CTA: flush per item (not used by the solver, kept for generality)
                const unsigned long long m1 = warp_umax(run_max), m2 = warp_umax(run_max2);
                if ((tid & 31) == 0) {
                    if (m1) atomicMax(&p.maxbits[2 * li], m1);
                    if (m2) atomicMax(&p.maxbits[2 * li + 1], m2);
                }
                run_max = run_max2 = 0ull;
            }
        }
    }
    // All stores of this tile must have left the register file before the next tile's prologue runs.  Without this
    // fence the LAST 128-bit stores of a tile were seen (CUDA 12.9 ptxas, sm_100a) to pick up register contents written
    // by the first instructions of the next prologue (LDC / multiplier set-up) -- on non-final tiles of a CTA only,
    // 1e-6-relative garbage in the low word at N = 1024, wrong Poisson outputs at N = 64; tiles_per_cta = 1 or this
    // fence make the results exact again (tests: test_fft_matches_pocketfft[1024-2], blocked-layout 64^3 trajectories).
#ifndef MSM_NO_TILE_FENCE
    asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
    }   // tiles of this CTA

#if MSM_ALIAS_PER_CTA
    if constexpr (sop_is_alias(SOP)) {
        // one partial per (stream, CTA), every sum in a fixed order -> deterministic (simulation_object.rs:1259-1280)
        const double a0 = warp_sum(alias0), a1 = warp_sum(alias1);
        if ((tid & 31) == 0) {
            red[0][tid >> 5] = a0;
            red[1][tid >> 5] = a1;
        }
        __syncthreads();
        if (tid < 2 && tid < p.gsz && g * p.gsz + tid < p.ns) {
            double tot = 0.0;
            for (int w = 0; w < (THREADS + 31) / 32; ++w) tot += red[tid][w];
            p.alias_partial[(long long)p.sid[g * p.gsz + tid] * p.ntiles + blockIdx.x] = tot;
        }
    }
#endif
    if constexpr (SOP == S_MAX) {
        if (p.gsz == 1 && g < p.ns) {
            // max|re|, max|im| of this CTA's tiles of buffer g; bit patterns of non-negative doubles order like integers
            const unsigned long long m1 = warp_umax(run_max), m2 = warp_umax(run_max2);
            if ((tid & 31) == 0) {
                redm[0][tid >> 5] = m1;
                redm[1][tid >> 5] = m2;
            }
            __syncthreads();
            if (tid == 0) {
                unsigned long long a = 0ull, b = 0ull;
                for (int w = 0; w < (THREADS + 31) / 32; ++w) {
                    a = umax64(a, redm[0][w]);
                    b = umax64(b, redm[1][w]);
                }
                if (a) atomicMax(&p.maxbits[2 * g], a);
                if (b) atomicMax(&p.maxbits[2 * g + 1], b);
            }
        }
    }
}

// host-side launcher, one translation unit per N (fft_inst.cu compiled with -DMSM_FFT_N=<N>)
typedef int (*pass_launcher_t)(bool inv, int lop, int sop, bool xl, const PassParams& p, int ntiles, int groups,
                               cudaStream_t st);
pass_launcher_t get_pass_launcher(int n);
// radices of the plan for length n (host side, for building the twiddle tables); returns the number of stages
int plan_radices(int n, int radices[4]);
// tile height of the contiguous-axis kernels for length n (host side)
int plan_tx(int n);
const char* pass_kernel_name(int n, bool inv, int lop, int sop);

}  // namespace msm
