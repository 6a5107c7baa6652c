// core.cu -- grid-level C ABI (msm_*): device state of a batch of streams and the fused pass sequences of one
// MSM time step.  See include/msm_b200.h for the contract and DESIGN.md for the data layout.
//
// Device layout (n = size, C = n^dims cells, S = local streams, G = chunk_streams):
//   X    [S][C] complex128   the ONE resident array per stream.  Between steps it holds psi_k (after the second
//                            drift, simulation_object.rs:574); inside msm_step it is transformed in place
//                            psi_k -> psi -> psi_k.  (The reference keeps psi, psi_k and phi: 3 arrays, :42-64.)
//   T    [G][C] complex128   scratch for the out-of-place transform psi_k -> rho that feeds the dt potential
//                            (the reference's `calculate_potential` at time t, :497) and for dumps.
//   P    [ceil(G/2)][C] complex128   pair buffers: rho_a + i rho_b  ->  phi_a + i phi_b.
//   dtab [S][n] complex128   per-axis drift factors n^(-1/2) exp(-i c_s (2 pi k_m)^2): exp(-i c k^2) is separable,
//                            so the reference's 2 GiB `k_evolution` temporary (:504-514) becomes a 8 KiB table.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/msm_b200.h"
#include "fft_pass.cuh"
#include "fft_tma.h"

namespace msm {
#define DECL(N) int launch_pass_##N(bool, int, int, bool, const PassParams&, int, int, cudaStream_t);
DECL(2) DECL(4) DECL(8) DECL(16) DECL(32) DECL(64) DECL(128) DECL(256) DECL(512) DECL(1024)
#undef DECL

pass_launcher_t get_pass_launcher(int n) {
    switch (n) {
#define C(N) case N: return launch_pass_##N;
        C(2) C(4) C(8) C(16) C(32) C(64) C(128) C(256) C(512) C(1024)
#undef C
    }
    return nullptr;
}

template <int N> static int radices_of(int r[4]) {
    for (int i = 0; i < 4; ++i) r[i] = Plan<N>::R[i];
    return Plan<N>::NS;
}
int plan_radices(int n, int r[4]) {
    switch (n) {
#define C(N) case N: return radices_of<N>(r);
        C(2) C(4) C(8) C(16) C(32) C(64) C(128) C(256) C(512) C(1024)
#undef C
    }
    return 0;
}

template <int N> static int tx_of() { return tile_T<N, true>(); }
int plan_tx(int n) {
    switch (n) {
#define C(N) case N: return tx_of<N>();
        C(2) C(4) C(8) C(16) C(32) C(64) C(128) C(256) C(512) C(1024)
#undef C
    }
    return 0;
}

static int plan_T(int n) { return n >= 8 ? 8 : n; }

// ---------------------------------------------------------------------------------------------------------
// device layout
// ---------------------------------------------------------------------------------------------------------
// Host buffers are linear: cell (i, j, k) at q = (i*n + j)*n + k (the reference's layout).  On the device 3-D grids
// with n >= 512 block the slowest axis: i = (i_hi, i_lo) with LO = 2^lb values of i_lo, rows of n contiguous k stay
// intact and are ordered [i_hi][j][i_lo].  A z-line then touches n/LO pages of 2 MiB instead of n (512^3: 32 vs
// 512 -- the linear layout thrashes the TLB in the z pass: 3.8 TB/s vs 5.4 TB/s for the y pass, profiles/README.md),
// and a y-line 2*LO pages.  lb == 0 is the identity.
__host__ __device__ __forceinline__ long long blk_index(long long q, int n, int lb) {
    if (lb == 0) return q;
    const long long k = q % n, r = q / n;
    const long long j = r % n, i = r / n;
    return (((((i >> lb) * n + j) << lb) + (i & ((1 << lb) - 1))) * n) + k;
}

// ---------------------------------------------------------------------------------------------------------
// small kernels (all iterate over the LINEAR cell index q; the device side of every access goes through blk_index)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_relayout(const double2* __restrict__ in, double2* __restrict__ out, long long cells, int n, int lb,
                           int to_device) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x) {
        const long long b = blk_index(q, n, lb);
        if (to_device) out[b] = in[q];
        else out[q] = in[b];
    }
}
__global__ void k_interleave(const double* __restrict__ re, const double* __restrict__ im, double2* __restrict__ out,
                             long long cells, int n, int lb) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x)
        out[blk_index(i, n, lb)] = make_double2(re[i], im[i]);
}
__global__ void k_deinterleave(const double2* __restrict__ in, double* __restrict__ re, double* __restrict__ im,
                               long long cells, int n, int lb) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells;
         i += (long long)gridDim.x * blockDim.x) {
        double2 v = in[blk_index(i, n, lb)];
        re[i] = v.x;
        im[i] = v.y;
    }
}
__global__ void k_extract(const double2* __restrict__ in, double* __restrict__ out, long long cells, int n, int lb,
                          int comp) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells;
         i += (long long)gridDim.x * blockDim.x) {
        double2 v = in[blk_index(i, n, lb)];
        out[i] = comp ? v.y : v.x;
    }
}
// sums over streams of v and |v|^2 (synthesizer/src/main.rs:63-93), streams added in list order -> deterministic.
// All grids share the device layout, so the cell index needs no mapping.
struct StreamList {
    int n;
    int id[msm::MAX_CHUNK];
};
__global__ void k_stream_sums(const double2* __restrict__ base, long long stride, StreamList sl, double2* __restrict__ sum,
                              double* __restrict__ sum2, long long cells, double f1, double f2, int accumulate) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x) {
        double2 a = accumulate ? sum[q] : make_double2(0.0, 0.0);
        double b = accumulate ? sum2[q] : 0.0;
        for (int i = 0; i < sl.n; ++i) {
            const double2 v = base[(long long)sl.id[i] * stride + q];
            a.x += f1 * v.x;
            a.y += f1 * v.y;
            b += f2 * (v.x * v.x + v.y * v.y);
        }
        sum[q] = a;
        sum2[q] = b;
    }
}
__global__ void k_real_to_planes(const double* __restrict__ in, double* __restrict__ re, long long cells, int n, int lb) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x)
        re[q] = in[blk_index(q, n, lb)];
}

// summed mode: only the REAL density travels over NVLink (the imaginary half of the pair buffer is zero)
__global__ void k_pack_real(const double2* __restrict__ in, double* __restrict__ out, long long cells) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x)
        out[q] = in[q].x;
}
__global__ void k_unpack_real(const double* __restrict__ in, double2* __restrict__ out, long long cells) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x)
        out[q] = make_double2(in[q], 0.0);
}

// Per-step scalars reach the host through MAPPED pinned memory written by a kernel, not through cudaMemcpy: a small
// D2H copy queues on the copy engine behind any bulk download in flight (a dump: 2 GiB at 512^3) and stalled the step
// loop for the length of that transfer.
__global__ void k_publish(const double* __restrict__ src, double* __restrict__ host_dst, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) host_dst[i] = src[i];
}
// per-axis drift factors dtab[s][m] = n^(-1/2) exp(-i c_s (2 pi k_m)^2) built on the device from the coefficients the
// host left in mapped pinned memory: no H2D copy behind the bulk uploads of other streams
// (simulation_object.rs:504-514 builds the full n^3 `k_evolution` array instead)
__global__ void k_build_dtab(const double* __restrict__ host_coef, const int* __restrict__ host_ids, const double* __restrict__ ksq,
                             double2* __restrict__ dtab, int n, double four_pi2, double sc) {
    const int s = host_ids[blockIdx.x];
    const double c = host_coef[s];
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        double sn, cs;
        sincos(-c * (ksq[m] * four_pi2), &sn, &cs);
        dtab[(long long)s * n + m] = make_double2(sc * cs, sc * sn);
    }
}
__global__ void k_extract_real(const double* __restrict__ in, double* __restrict__ out, long long cells, int n, int lb) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x)
        out[i] = in[blk_index(i, n, lb)];
}

// alias_out[s] = dv * sum_tiles partial[s][tile]    (fixed summation order: deterministic)
__global__ void k_alias_reduce(const double* __restrict__ partial, double* __restrict__ out, int ntiles, int pitch,
                               double dv) {
    __shared__ double sh[256];
    const int s = blockIdx.x;
    const double* p = partial + (long long)s * pitch;
    (void)pitch;
    double a = 0.0;
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x) a += p[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[s] = sh[0] * dv;
}
// k^2 from indices, same expression as the pass kernels (parity with spec_grid, utils/fft.rs:123-161)
__global__ void k_spec_grid(const double* __restrict__ ksq, double* __restrict__ out, int n, int dims, double four_pi2) {
    long long total = 1;
    for (int d = 0; d < dims; ++d) total *= n;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
         q += (long long)gridDim.x * blockDim.x) {
        int c0 = (int)(q % n);
        int c1 = dims >= 2 ? (int)((q / n) % n) : 0;
        int c2 = dims >= 3 ? (int)(q / ((long long)n * n)) : 0;
        double s = ksq[c0];
        if (dims >= 2) s = s + ksq[c1];
        if (dims >= 3) s = s + ksq[c2];
        out[q] = s * four_pi2;
    }
}
__global__ void k_scale_real(double2* __restrict__ a, long long n, double f) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double2 v = a[i];
        v.x *= f;
        v.y *= f;
        a[i] = v;
    }
}

// ---- on-device initial conditions (restates simulator/src/ics.rs) --------------------------
// psi(i,j,k) = gx[k] * gy[j] * gz[i] * norm          (cold_gauss, ics.rs:24-162: separable, already normalised)
__global__ void k_ic_separable(double2* __restrict__ psi, const double* __restrict__ g, int n, int dims, double norm,
                               int lb) {
    long long total = 1;
    for (int d = 0; d < dims; ++d) total *= n;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
         q += (long long)gridDim.x * blockDim.x) {
        int c0 = (int)(q % n);
        int c1 = dims >= 2 ? (int)((q / n) % n) : 0;
        int c2 = dims >= 3 ? (int)(q / ((long long)n * n)) : 0;
        double v = g[c0];
        if (dims >= 2) v = g[n + c1] * v;
        if (dims >= 3) v = g[2 * n + c2] * v;
        psi[blk_index(q, n, lb)] = make_double2(v * norm, 0.0);
    }
}
// spherical_tophat (ics.rs:165-280): sqrt(1 + delta / (1 + exp(slope (r/R - 1)))), un-normalised
__global__ void k_ic_tophat(double2* __restrict__ psi, int n, int dims, double dx, double half, double radius,
                            double delta, double slope, int lb) {
    long long total = 1;
    for (int d = 0; d < dims; ++d) total *= n;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
         q += (long long)gridDim.x * blockDim.x) {
        int c0 = (int)(q % n);
        int c1 = dims >= 2 ? (int)((q / n) % n) : -1;
        int c2 = dims >= 3 ? (int)(q / ((long long)n * n)) : -1;
        // reference loop order: x outermost (slowest NumPy axis), z innermost; the radius is symmetric
        double a = (2.0 * c0 + 1.0) * dx / 2.0 - half;
        double b = c1 >= 0 ? (2.0 * c1 + 1.0) * dx / 2.0 - half : 0.0;
        double c = c2 >= 0 ? (2.0 * c2 + 1.0) * dx / 2.0 - half : 0.0;
        double r;
        if (dims == 3) r = sqrt(c * c + b * b + a * a);
        else if (dims == 2) r = sqrt(b * b + a * a + 0.0);
        else r = sqrt(a * a + 0.0 + 0.0);
        double ramp = 1.0 / (1.0 + exp(slope * (r / radius - 1.0)));
        psi[blk_index(q, n, lb)] = make_double2(sqrt(1.0 + delta * ramp), 0.0);
    }
}
__global__ void k_norm2_partial(const double2* __restrict__ a, long long n, double* __restrict__ partial) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double2 v = a[i];
        acc += v.x * v.x + v.y * v.y;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
// psi *= exp(2 pi i u(q)), u from Philox draw slot `draw` (cold_gauss_kspace random phases, ics.rs:407-423)
__global__ void k_random_phase(double2* __restrict__ psi, long long cells, uint64_t seed, uint32_t draw, int n, int lb) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x) {
        uint32_t c0 = (uint32_t)q, c1 = (uint32_t)((uint64_t)q >> 32), c2 = draw, c3 = 0u;
        philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
        double sn, cs;
        sincos(2.0 * 3.14159265358979323846 * u53(c0, c1), &sn, &cs);
        const long long b = blk_index(q, n, lb);
        const double2 v = psi[b];
        psi[b] = make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    }
}
// sample_quantum_perturbation, Wigner / Husimi (ics.rs:560-646): psi = (psi sqrt(dV) + (N + iN) / div) / sqrt(dV)
__global__ void k_sample_gauss(double2* __restrict__ psi, long long cells, uint64_t seed, double sqrt_dv, double div,
                               int n, int lb) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x) {
        uint32_t c0 = (uint32_t)q, c1 = (uint32_t)((uint64_t)q >> 32), c2 = 0u, c3 = 0u;
        philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = u53(c0, c1), u2 = u53(c2, c3);
        const double r = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincos(2.0 * 3.14159265358979323846 * u2, &sn, &cs);
        const long long b = blk_index(q, n, lb);   // the Philox counter is the LINEAR cell index
        double2 v = psi[b];
        v.x = (v.x * sqrt_dv + (r * cs) / div) / sqrt_dv;
        v.y = (v.y * sqrt_dv + (r * sn) / div) / sqrt_dv;
        psi[b] = v;
    }
}
// sample_quantum_perturbation, Poisson scheme (ics.rs:495-558): |psi| <- sqrt(Pois(|psi|^2 dV n_tot) / n_tot) / sqrt(dV),
// phase kept.  The reference draws from the unseeded thread_rng (rand_distr 0.4.3), so only the distribution can be
// matched: counter-based Philox keyed (seed, cell), draw slots 16, 17, ...; Knuth's product method for lambda < 10,
// Hoermann's PTRS transformed rejection above.
__device__ __forceinline__ void philox_pair(long long q, uint64_t seed, uint32_t draw, double* u1, double* u2) {
    uint32_t c0 = (uint32_t)q, c1 = (uint32_t)((uint64_t)q >> 32), c2 = draw, c3 = 0u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    *u1 = u53(c0, c1);
    *u2 = u53(c2, c3);
}
__device__ double poisson_variate(double lam, long long q, uint64_t seed) {
    uint32_t draw = 16u;
    double u1, u2;
    if (!(lam > 0.0)) return 0.0;
    if (lam < 10.0) {
        const double limit = exp(-lam);
        double prod = 1.0, k = -1.0;
        for (;;) {
            philox_pair(q, seed, draw++, &u1, &u2);
            prod *= u1;
            k += 1.0;
            if (!(prod > limit)) return k;
            prod *= u2;
            k += 1.0;
            if (!(prod > limit)) return k;
        }
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        philox_pair(q, seed, draw++, &u1, &u2);
        const double U = u1 - 0.5, us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && u2 <= vr) return k;
        if (k < 0.0 || (us < 0.013 && u2 > us)) continue;
        if (log(u2) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
}
__global__ void k_sample_poisson(double2* __restrict__ psi, long long cells, uint64_t seed, double dv, double n_tot, int n,
                                 int lb) {
    const double sqrt_dv = sqrt(dv);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < cells; q += (long long)gridDim.x * blockDim.x) {
        const long long b = blk_index(q, n, lb);
        const double2 v = psi[b];
        const double norm2 = v.x * v.x + v.y * v.y;
        const double count = poisson_variate(norm2 * dv * n_tot, q, seed);     // ics.rs:509-519
        const double mag = sqrt(count / n_tot);                                 // :523
        // exp(i arg(psi)) (:535-544); arg(0) = 0
        const double r = sqrt(norm2);
        const double cs = r > 0.0 ? v.x / r : 1.0, sn = r > 0.0 ? v.y / r : 0.0;
        psi[b] = make_double2(mag * cs / sqrt_dv, mag * sn / sqrt_dv);          // :547-557
    }
}
}  // namespace msm

using namespace msm;

// ---------------------------------------------------------------------------------------------------------
// NCCL through dlopen (only the coupled mode with nranks > 1 needs it)
// ---------------------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_uid_t;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return false;
        GetUniqueId = (int (*)(nccl_uid_t*))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (int (*)(void**, int, nccl_uid_t, int))dlsym(h, "ncclCommInitRank");
        CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
        GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce;
    }
};
NcclApi g_nccl;
const int NCCL_DOUBLE = 8;  // ncclFloat64
const int NCCL_SUM = 0;

std::string g_create_error;
}  // namespace

struct ProfEvent {
    cudaEvent_t a, b;
    int key;
    double bytes;
};

struct msm_ctx {
    msm_config cfg;
    int n = 0, dims = 0, S = 0, chunk = 0, T = 0, TX = 0;   // tile heights: strided axes / contiguous axis
    int lb = 0;   // log2 of the slow-axis block (device layout, see blk_index)
    long long C = 0;
    cudaStream_t st = nullptr;
    double2 *X = nullptr, *Tscr = nullptr, *P = nullptr, *tw = nullptr, *dtab = nullptr;
    double* h_coef = nullptr;    // pinned + mapped: drift coefficient per stream, read by k_build_dtab
    int* h_ids = nullptr;        // pinned + mapped: streams whose tables are rebuilt
    cudaEvent_t dtab_done = nullptr;
    double* ksq = nullptr;
    double* alias_partial = nullptr;
    double* alias_out = nullptr;
    double* h_scal = nullptr;    // pinned + mapped, 2*(S+2) doubles: written by k_publish, read by the host after a sync
    unsigned long long* maxbits = nullptr;
    double* scratch_small = nullptr;  // 4096 doubles
    double2* ic_base = nullptr;                        // msm_ic_store / msm_ic_load: one saved wavefunction
    char ic_base_in_k = 0;
    double2 *ens_psi = nullptr, *ens_psik = nullptr;   // ensemble sums (row f-3), allocated on first use
    double *ens_psi2 = nullptr, *ens_psik2 = nullptr;
    std::vector<char> in_k, has_psi;
    // max|phi| of the CURRENT psi_k, computed eagerly at the end of msm_step (its first pass is fused into the step's
    // last pass); msm_potential_max returns it without touching the GPU while it is valid
    std::vector<char> pmax_valid;
    std::vector<double> pmax_cache;
    std::vector<int> pmax_pending;   // streams whose value still sits in `maxbits` (position = index in this list)
    std::vector<double> h_ksq;
    double four_pi2 = 0, k2_max = 0, dv = 0;
    int ntiles_last = 0, ntiles_used = 0;   // allocation pitch bound / tiles of the last forward pass as launched
    int alias_count = 0;                    // alias partials per stream written by the last alias pass (MSM_ALIAS_PER_CTA: CTAs)
    uint64_t bytes = 0, launches = 0;
    void* comm = nullptr;
    cudaEvent_t tm_a = nullptr, tm_b = nullptr;
    cudaStream_t copy_st = nullptr;                       // msm_get_psi_many: D2H overlapped with compute
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    // asynchronous transfers (msm_upload_begin / msm_download_begin): PCIe traffic of other streams overlaps the
    // step kernels.  Uploads run on up_st (staging + relayout when the device layout is blocked) and signal one
    // event per stream; downloads are transformed / plane-split on the compute stream into two dedicated staging
    // buffers and leave on copy_st.
    cudaStream_t up_st = nullptr;
    std::vector<cudaEvent_t> up_ev;
    std::vector<char> up_pending;
    cudaEvent_t ev_x_free = nullptr;
    double2* up_stage = nullptr;
    double* dl_stage[2] = {nullptr, nullptr};
    cudaEvent_t dl_ready[2] = {nullptr, nullptr}, dl_copied[2] = {nullptr, nullptr};
    char dl_used[2] = {0, 0};
    int dl_next = 0;
    // download tickets (msm_download_ticket / msm_download_wait): a ring of events on the copy stream; writer threads
    // of the host-logic level wait on them, hence the mutex
    std::mutex tk_mu;
    cudaEvent_t tk_ev[64] = {};
    uint64_t tk_next = 0;
    // profiling
    bool prof = false;
    std::vector<ProfEvent> prof_pending;
    std::vector<std::string> prof_names;
    std::map<std::string, int> prof_index;
    std::vector<msm_profile_record> prof_rec;
    std::string err;
    pass_launcher_t launcher = nullptr;
    // real-field Poisson solve of the summed-density mode (dims == 3, n >= 16): the real plane of n^3 doubles IS a grid
    // of n/2 x n x n complex pairs (x[2j], x[2j+1]); R2C / C2R run as n/2-point passes over it, the y / z passes as
    // n-point passes over the half spectrum, and X[n/2] of every x line lives in a small Nyquist plane (n x n complex)
    // TMA variant of the plain strided 512-point pass (fft_tma.cu), MSM_B200_TMA=1: tensor maps [array][axis - 1]
    bool tma = false;
    TmaMap tma_map[3][2];
    bool real_solve = false;
    // slab pipeline of the density all-reduce (nranks > 1): the last x pass that completes rho is launched in `ar_slabs`
    // row ranges; as soon as a range is final its ncclAllReduce runs on comm_st (high priority) while the next range is
    // still being computed, and the R2C pass of the Poisson solve starts range by range as the sums arrive
    int ar_slabs = 1;
    cudaStream_t comm_st = nullptr;
    cudaEvent_t ev_slab[8] = {}, ev_ar[8] = {};
    pass_launcher_t launcher_half = nullptr;
    double2 *tw_half = nullptr, *wreal = nullptr, *nyq = nullptr;
    int TXH = 0;           // contiguous-axis tile height of the n/2-point kernels
    bool xl = true;   // contiguous-axis thread mapping (MSM_B200_XL=0 selects the generic mapping, for A/B timing)
    bool fuse = true;  // fused passes (MSM_B200_FUSE=0 runs the plain 3+3 pass sequences, for A/B timing)
    int l2_prefetch = 1;   // MSM_B200_PREFETCH=0 switches the L2 prefetch of the next item off (A/B timing)
    int tiles_per_cta = 4;   // consecutive tiles per CTA of the one-tile kernel; next item is prefetched into L2
    int tiles_per_cta_x = 16;  // the same for the small tiles of the contiguous axis
    int tiles_per_cta_fused = 4;   // multiplier of tiles_per_cta for the strided two-transform kernels (MSM_B200_TPCF)
    int interleave = 1;            // MSM_B200_ILV: neighbouring CTAs interleave their tiles on the strided axes
    int num_sms = 148;
};

namespace {

int fail(msm_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, MSM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

struct Geom {
    int axis, tiles_inner, ntiles, lvalid, olb = 0, alb = 0;
    long long outer, inner, lstride, astride, outer_lo = 0, astride_lo = 0;
};
// A grid as the pass kernels see it: `dims` axes of n points, except that the fastest one has nx (= n for the
// wavefunctions, n / 2 for the half-spectrum grid of the real-field solve); slow axis blocked by 2^lb (blk_index).
struct GridShape {
    int dims, n, nx, lb;
    long long cells() const {
        long long c = nx;
        for (int d = 1; d < dims; ++d) c *= n;
        return c;
    }
};
GridShape main_shape(const msm_ctx* c) { return GridShape{c->dims, c->n, c->n, c->lb}; }
// what a pass runs on: the wavefunction grid, or one of the three views of the real-field solve
enum Target { TG_MAIN = 0, TG_HALF_X = 1, TG_HALF_YZ = 2, TG_NYQ = 3 };

Geom make_geom(const GridShape& s, int axis, int T) {
    Geom g{};
    const int n = s.n, nx = s.nx;
    g.axis = axis;
    if (axis == 0) {
        const long long nlines = s.cells() / nx;
        g.ntiles = (int)((nlines + T - 1) / T);
        g.tiles_inner = g.ntiles;
        g.inner = (long long)T * nx;
        g.outer = 0;
        g.lstride = nx;
        g.astride = 1;
        g.lvalid = (int)std::min<long long>(T, nlines);
    } else if (axis == 1) {
        // along j for fixed (i, k0..k0+T-1): tile = i * (nx/T) + m
        const long long LO = 1LL << s.lb;
        g.tiles_inner = nx / T;
        g.ntiles = g.tiles_inner * (s.dims == 3 ? n : 1);
        g.inner = T;
        g.olb = s.lb;                        // i = (i_hi, i_lo)
        g.outer = (long long)nx * n * LO;    // i_hi
        g.outer_lo = nx;                     // i_lo
        g.lstride = 1;
        g.astride = (long long)nx * LO;      // j
        g.lvalid = T;
    } else {
        // along i for fixed (j, k0..k0+T-1): tile = j * (nx/T) + m
        const long long LO = 1LL << s.lb;
        g.tiles_inner = nx / T;
        g.ntiles = (int)(((long long)nx * n) / T);
        g.inner = T;
        g.outer = (long long)nx * LO;        // j
        g.lstride = 1;
        g.alb = s.lb;
        g.astride = (long long)nx * n * LO;  // i_hi
        g.astride_lo = nx;                   // i_lo
        g.lvalid = T;
    }
    return g;
}
Geom make_geom(const msm_ctx* c, int axis, int T) { return make_geom(main_shape(c), axis, T); }

const char* lop_name(int l) {
    return l == L_NONE ? "none" : l == L_DRIFT ? "drift" : l == L_KICK ? "kick" : l == L_C2R ? "c2r" : "invx+kick";
}
const char* sop_name(int s) {
    switch (s) {
        case S_NONE: return "none";
        case S_SCALE: return "scale";
        case S_DRIFT: return "drift";
        case S_DRIFT_ALIAS: return "drift+alias";
        case S_RHO_KEEP: return "psi+rho";
        case S_RHO_ONLY: return "rho";
        case S_POISSON: return "poisson";
        case S_MAX: return "max";
        case S_POISSON_INV: return "poisson+inv";
        case S_RHO_KEEP_FX: return "psi+rho+fwdx";
        case S_RHO_ONLY_FX: return "rho+fwdx";
        case S_DRIFT_ALIAS_IZ: return "drift+alias+inv";
        case S_R2C: return "r2c";
    }
    return "?";
}

// algorithmic HBM bytes of one pass over `cells` complex elements per stream; share = streams that share one real
// rho / phi value (2 for pair buffers, the whole CTA group in the summed-density mode)
double pass_bytes(double cells, int lop, int sop, int ns, int share) {
    double per = 16.0;                                   // read the line
    if (lop == L_KICK || lop == L_KICK_IX) per += 16.0 / share;   // phi
    if (sop != S_RHO_ONLY && sop != S_RHO_ONLY_FX && sop != S_MAX) per += 16.0;   // write the line
    if (sop_is_rho(sop)) per += 16.0 / share;            // rho
    if (sop == S_DRIFT_ALIAS_IZ) per += 16.0;            // second output
    return per * cells * ns;
}

int prof_key(msm_ctx* c, const std::string& name) {
    auto it = c->prof_index.find(name);
    if (it != c->prof_index.end()) return it->second;
    int k = (int)c->prof_names.size();
    c->prof_names.push_back(name);
    c->prof_index[name] = k;
    c->prof_rec.push_back(msm_profile_record{nullptr, 0, 0.0, 0.0});
    return k;
}

struct ProfScope {
    msm_ctx* c;
    ProfEvent ev{};
    bool on;
    cudaStream_t s;
    ProfScope(msm_ctx* c_, const std::string& name, double bytes, cudaStream_t stream = nullptr)
        : c(c_), on(c_->prof), s(stream ? stream : c_->st) {
        if (!on) return;
        ev.key = prof_key(c, name);
        ev.bytes = bytes;
        cudaEventCreate(&ev.a);
        cudaEventCreate(&ev.b);
        cudaEventRecord(ev.a, s);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(ev.b, s);
        c->prof_pending.push_back(ev);
    }
};

void prof_drain(msm_ctx* c) {
    for (auto& e : c->prof_pending) {
        cudaEventSynchronize(e.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        c->prof_rec[e.key].launches += 1;
        c->prof_rec[e.key].ms_total += ms;
        c->prof_rec[e.key].algorithmic_bytes += e.bytes;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    c->prof_pending.clear();
}

// what the passes of one transform do besides the FFT
struct XformOps {
    int lop_first = L_NONE, lop_each = L_NONE, sop_each = S_NONE, sop_last = S_NONE;
    int gsz = 2;
    int p_summed = 0, rho_accumulate = 0;
    double rho_coef = 0, scale = 1, poisson_coef = 0;
    const double* kick = nullptr;   // per local index
    double2* pbuf = nullptr;
    unsigned long long* maxbits = nullptr;
    double2* dst2 = nullptr;
    bool dtab_shared = false;       // every stream of the launch has the same drift coefficient (summed density: one dt)
};

struct PassSpec {
    int axis;
    bool inv;
    int lop, sop;
    int target = TG_MAIN;
    int tile0 = 0, tile_end = -1;   // tile range of this launch (slab-pipelined launches); -1 = all
};

// a sequence of axis passes over the streams ids[0..ns): the first pass reads `src`, every pass writes `work`.
int run_passes(msm_ctx* ctx, const std::vector<PassSpec>& seq, const int* ids, int ns, const double2* src, int src_by_sid,
               double2* work, int work_by_sid, const XformOps& o) {
    PassParams p{};
    p.src_sstride = p.dst_sstride = ctx->C;
    p.ns = ns;
    p.gsz = o.gsz;
    for (int i = 0; i < ns; ++i) p.sid[i] = ids[i];
    p.n = ctx->n;
    p.twiddle = ctx->tw;
    p.dtab = ctx->dtab;
    if (o.kick) for (int i = 0; i < ns; ++i) p.kick[i] = o.kick[i];
    p.pbuf = o.pbuf ? o.pbuf : ctx->P;
    p.p_gstride = ctx->C;
    p.p_summed = o.p_summed;
    p.rho_accumulate = o.rho_accumulate;
    p.dtab_shared = o.dtab_shared ? 1 : 0;
    p.ksq = ctx->ksq;
    p.four_pi2 = ctx->four_pi2;
    p.alias_k2_thresh = ctx->k2_max * ctx->cfg.k2_cutoff;   // simulation_object.rs:1265
    p.poisson_coef = o.poisson_coef;
    p.rho_coef = o.rho_coef;
    p.scale = o.scale;
    p.alias_partial = ctx->alias_partial;
    p.maxbits = o.maxbits ? o.maxbits : ctx->maxbits;
    p.dst2 = o.dst2;
    p.nyq = ctx->nyq;
    p.wreal = ctx->wreal;
    for (size_t k = 0; k < seq.size(); ++k) {
        const int axis = seq[k].axis, lop = seq[k].lop, sop = seq[k].sop, tg = seq[k].target;
        const bool inv = seq[k].inv;
        // grid view, transform length, kernels and tile heights of this pass
        GridShape shape = main_shape(ctx);
        pass_launcher_t launcher = ctx->launcher;
        int N = ctx->n, TXL = ctx->TX;
        p.twiddle = ctx->tw;
        p.k2_fixed = 0.0;
        if (tg == TG_HALF_X) {          // n/2-point passes along x over the real plane read as complex pairs
            shape.nx = ctx->n / 2;
            launcher = ctx->launcher_half;
            N = ctx->n / 2;
            TXL = ctx->TXH;
            p.twiddle = ctx->tw_half;
        } else if (tg == TG_HALF_YZ) {  // n-point passes along y / z over the half spectrum
            shape.nx = ctx->n / 2;
        } else if (tg == TG_NYQ) {      // the plane k_x = Nyquist: a linear 2-D grid over (k_y, k_z)
            shape = GridShape{2, ctx->n, ctx->n, 0};
            p.k2_fixed = ctx->h_ksq[ctx->n / 2];
        }
        const bool xl = axis == 0 && ctx->xl;
        // Streams per CTA group.  Only passes that touch the pair buffer need the whole group in one CTA (summed
        // coupling: rho accumulates over / phi is shared by all streams of the group); every other pass runs in pairs, so
        // that the two drift tables of a CTA stay resident in shared memory (fft_pass.cuh).
        const bool pair_pass = lop == L_KICK || lop == L_KICK_IX || sop_is_rho(sop) || sop == S_POISSON ||
                               sop == S_POISSON_INV || sop == S_MAX;
        p.gsz = pair_pass ? o.gsz : std::min(o.gsz, 2);
        const int groups = (ns + p.gsz - 1) / p.gsz;
        const Geom g = make_geom(shape, axis, xl ? TXL : ctx->T);
        const bool first = (k == 0);
        p.src = first ? src : work;
        p.src_by_sid = first ? src_by_sid : work_by_sid;
        p.dst = work;
        p.dst_by_sid = work_by_sid;
        p.src_sstride = p.dst_sstride = tg == TG_MAIN ? ctx->C : shape.cells();
        p.axis = axis;
        p.nx = shape.nx;
        p.tiles_inner = g.tiles_inner;
        p.outer_stride = g.outer;
        p.inner_stride = g.inner;
        p.lstride = g.lstride;
        p.astride = g.astride;
        p.olb = g.olb;
        p.alb = g.alb;
        p.row_lb = shape.lb;
        p.outer_lo = g.outer_lo;
        p.astride_lo = g.astride_lo;
        p.lvalid = g.lvalid;
        p.ntiles = g.ntiles;
        p.tile0 = seq[k].tile0;
        p.tile_end = seq[k].tile_end < 0 ? g.ntiles : std::min(seq[k].tile_end, g.ntiles);
        char nm[112];
        // consecutive tiles of one CTA must differ by inner_stride only: tiles_per_cta divides tiles_inner
        // (the contiguous axis has small tiles: more of them per CTA amortise the table preamble)
        // (the strided two-transform kernels gain 3 % from longer walks, the plain passes lose: profiles/r02l_sweep.log)
        const int walk = xl ? ctx->tiles_per_cta_x
                            : ctx->tiles_per_cta * ((sop == S_POISSON_INV || sop == S_DRIFT_ALIAS_IZ) ? ctx->tiles_per_cta_fused : 1);
        p.tiles_per_cta = (int)std::__gcd((long long)walk, (long long)g.tiles_inner);
        // interleaved walk of W neighbouring CTAs (strided axes only; whole launches only; W * tiles_per_cta must tile a row)
        p.interleave = 1;
        if (!xl && ctx->interleave > 1 && seq[k].tile0 == 0 && seq[k].tile_end < 0 &&
            g.tiles_inner % (ctx->interleave * p.tiles_per_cta) == 0)
            p.interleave = ctx->interleave;
        if (p.tile0 % p.tiles_per_cta) p.tiles_per_cta = (int)std::__gcd((long long)p.tiles_per_cta, (long long)p.tile0);
        p.l2_prefetch = ctx->l2_prefetch;
        if (sop_is_alias(sop))
            ctx->alias_count = MSM_ALIAS_PER_CTA ? (g.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta : g.ntiles;
        snprintf(nm, sizeof nm, "fft_pass<%d,%s,%s,%s,%s%s>", N, inv ? "inv" : "fwd", lop_name(lop), sop_name(sop),
                 axis == 0 ? "x" : axis == 1 ? "y" : "z", tg == TG_MAIN ? "" : tg == TG_NYQ ? ",nyquist" : ",half");
        const double frac = (double)(p.tile_end - p.tile0) / (double)g.ntiles;
        int rc;
        const double2* arrays[3] = {ctx->X, ctx->Tscr, ctx->P};
        int which = -1;
        for (int a = 0; a < 3; ++a)
            if (p.src == arrays[a] && p.dst == arrays[a]) which = a;
        if (ctx->tma && tg == TG_MAIN && lop == L_NONE && sop == S_NONE && axis > 0 && which >= 0 && p.tile0 == 0 &&
            p.tile_end == g.ntiles) {
            // experiment: the same pass through cp.async.bulk.tensor (fft_tma.cu), one CTA column per stream
            TmaPassParams tp{};
            tp.twiddle = ctx->tw;
            for (int i = 0; i < ns; ++i) tp.slot[i] = p.src_by_sid ? ids[i] : i;
            tp.axis = axis;
            tp.ntiles = g.ntiles;
            tp.tiles_inner = g.tiles_inner;
            tp.tiles_per_cta = p.tiles_per_cta;
            tp.lb = ctx->lb;
            // cp.async.bulk.prefetch.tensor of later items measured SLOWER (1.29 x DRAM reads: prefetch and load both miss)
            tp.l2_prefetch = getenv("MSM_B200_TMA_PF") ? atoi(getenv("MSM_B200_TMA_PF")) : 0;
            strncat(nm, "[tma]", sizeof nm - strlen(nm) - 1);
            ProfScope ps(ctx, nm, pass_bytes((double)shape.cells(), lop, sop, ns, 2));
            tp.ns = ns;
            rc = tma_launch_pass(inv, &ctx->tma_map[which][axis - 1], tp, ctx->num_sms, ctx->st);
        } else {
            ProfScope ps(ctx, nm, frac * pass_bytes((double)shape.cells(), lop, sop, ns, o.p_summed ? std::max(1, p.gsz) : 2));
            rc = launcher(inv, lop, sop, xl, p, g.ntiles, groups, ctx->st);
        }
        ctx->launches++;
        if (rc == -1) return fail(ctx, MSM_E_ARG, std::string("no kernel instance for ") + nm);
        if (rc != 0) return fail(ctx, MSM_E_CUDA, std::string(nm) + ": " + cudaGetErrorString((cudaError_t)rc));
    }
    return MSM_OK;
}

// passes of a d-dimensional transform (forward x,y,z / inverse z,y,x) with the operators of `o`
std::vector<PassSpec> transform_seq(const msm_ctx* ctx, bool inv, const XformOps& o) {
    std::vector<PassSpec> seq;
    for (int k = 0; k < ctx->dims; ++k) {
        const bool first = (k == 0), last = (k == ctx->dims - 1);
        PassSpec s{};
        s.axis = inv ? ctx->dims - 1 - k : k;
        s.inv = inv;
        s.lop = (first && o.lop_first != L_NONE) ? o.lop_first : o.lop_each;
        s.sop = last ? o.sop_last : o.sop_each;
        s.target = TG_MAIN;
        s.tile0 = 0;
        s.tile_end = -1;
        seq.push_back(s);
    }
    return seq;
}
int run_transform(msm_ctx* ctx, bool inv, const int* ids, int ns, const double2* src, int src_by_sid, double2* work,
                  int work_by_sid, const XformOps& o) {
    return run_passes(ctx, transform_seq(ctx, inv, o), ids, ns, src, src_by_sid, work, work_by_sid, o);
}

int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 16); }

// the compute stream must not touch X[s] before an asynchronous upload of stream s has landed
int wait_upload(msm_ctx* ctx, int s) {
    if (!ctx->up_pending.empty() && ctx->up_pending[s]) {
        if (cudaStreamWaitEvent(ctx->st, ctx->up_ev[s], 0) != cudaSuccess)
            return fail(ctx, MSM_E_CUDA, "cudaStreamWaitEvent(upload) failed");
        ctx->up_pending[s] = 0;
    }
    return MSM_OK;
}

int ensure_kspace(msm_ctx* ctx, const std::vector<int>& ids) {
    std::vector<int> todo;
    for (int s : ids) {
        if (int rc = wait_upload(ctx, s)) return rc;
        if (!ctx->has_psi[s]) return fail(ctx, MSM_E_STATE, "stream has no wavefunction (call msm_set_psi first)");
        if (!ctx->in_k[s]) todo.push_back(s);
    }
    for (size_t i = 0; i < todo.size(); i += ctx->chunk) {
        const int ns = (int)std::min<size_t>(ctx->chunk, todo.size() - i);
        XformOps o;
        o.sop_each = o.sop_last = S_SCALE;
        o.scale = 1.0 / sqrt((double)ctx->n);   // n^(-dims/2) over dims passes (utils/fft.rs:17)
        int rc = run_transform(ctx, false, &todo[i], ns, ctx->X, 1, ctx->X, 1, o);
        if (rc) return rc;
    }
    for (int s : todo) ctx->in_k[s] = 1;
    return MSM_OK;
}

std::vector<int> active_list(const msm_ctx* ctx, const int32_t* active) {
    std::vector<int> ids;
    for (int s = 0; s < ctx->S; ++s)
        if (!active || active[s]) ids.push_back(s);
    return ids;
}

// plane k (0 / 1) of the summed-density mode: n^d doubles inside pair buffer 0.  Plane 0 holds the density / potential
// of the kick, plane 1 the density of the NEXT step's dt-potential (accumulated while later chunks still read phi)
double2* real_plane(msm_ctx* ctx, int k) {
    return reinterpret_cast<double2*>(reinterpret_cast<double*>(ctx->P) + (size_t)k * ctx->C);
}

// Sum the density over the ranks: `ncclAllReduce` of n^d doubles, in place.  With the real-field solve the REAL plane
// goes over NVLink as it lies in memory (1 GiB at 512^3); the complex fallback (dims < 3) packs / unpacks the real parts.
int allreduce_rho(msm_ctx* ctx, int plane) {
    if (ctx->cfg.nranks <= 1) return MSM_OK;
    int rc;
    if (ctx->real_solve) {
        double* r = reinterpret_cast<double*>(real_plane(ctx, plane));
        {
            ProfScope ps(ctx, "nccl_allreduce_rho", 0.0);
            rc = g_nccl.AllReduce(r, r, (size_t)ctx->C, NCCL_DOUBLE, NCCL_SUM, ctx->comm, ctx->st);
        }
        if (rc != 0)
            return fail(ctx, MSM_E_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
        return MSM_OK;
    }
    double* packed = reinterpret_cast<double*>(ctx->Tscr);
    {
        ProfScope ps(ctx, "pack_rho", 24.0 * ctx->C);
        k_pack_real<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->P, packed, ctx->C);
    }
    ctx->launches++;
    {
        ProfScope ps(ctx, "nccl_allreduce_rho", 0.0);
        rc = g_nccl.AllReduce(packed, packed, (size_t)ctx->C, NCCL_DOUBLE, NCCL_SUM, ctx->comm, ctx->st);
    }
    if (rc != 0)
        return fail(ctx, MSM_E_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    {
        ProfScope ps(ctx, "unpack_rho", 24.0 * ctx->C);
        k_unpack_real<<<grid_for(ctx->C), 256, 0, ctx->st>>>(packed, ctx->P, ctx->C);
    }
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, MSM_E_CUDA, "pack/unpack of rho failed");
    return MSM_OK;
}

int run_passes(msm_ctx* ctx, const std::vector<PassSpec>& seq, const int* ids, int ns, const double2* src, int src_by_sid,
               double2* work, int work_by_sid, const XformOps& o);

// rows [k, k+1) * rows / ar_slabs of `plane` are final on the compute stream: sum them over the ranks on the
// communication stream, behind the kernels that still compute the later ranges
int allreduce_slab(msm_ctx* ctx, int plane, int k) {
    const long long rows = ctx->C / ctx->n;
    const long long r0 = rows * k / ctx->ar_slabs, r1 = rows * (k + 1) / ctx->ar_slabs;
    double* r = reinterpret_cast<double*>(real_plane(ctx, plane)) + r0 * ctx->n;
    if (cudaEventRecord(ctx->ev_slab[k], ctx->st) != cudaSuccess || cudaStreamWaitEvent(ctx->comm_st, ctx->ev_slab[k], 0) != cudaSuccess)
        return fail(ctx, MSM_E_CUDA, "allreduce_slab: event handshake failed");
    int rc;
    {
        ProfScope ps(ctx, "nccl_allreduce_rho", 0.0, ctx->comm_st);
        rc = g_nccl.AllReduce(r, r, (size_t)((r1 - r0) * ctx->n), NCCL_DOUBLE, NCCL_SUM, ctx->comm, ctx->comm_st);
    }
    if (rc != 0)
        return fail(ctx, MSM_E_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    if (cudaEventRecord(ctx->ev_ar[k], ctx->comm_st) != cudaSuccess) return fail(ctx, MSM_E_CUDA, "allreduce_slab: event record failed");
    return MSM_OK;
}

// the LAST pass that completes the summed density (seq.back(), an x pass with a rho store operator), launched range by
// range with the all-reduce of every finished range started at once
int run_passes_slabbed(msm_ctx* ctx, std::vector<PassSpec> seq, const int* ids, int ns, const double2* src, int src_by_sid,
                       double2* work, int work_by_sid, const XformOps& o, int plane) {
    PassSpec last = seq.back();
    seq.pop_back();
    int rc;
    if (!seq.empty() && (rc = run_passes(ctx, seq, ids, ns, src, src_by_sid, work, work_by_sid, o))) return rc;
    const long long rows = ctx->C / ctx->n;
    for (int k = 0; k < ctx->ar_slabs; ++k) {
        last.tile0 = (int)(rows * k / ctx->ar_slabs / ctx->TX);
        last.tile_end = (int)(rows * (k + 1) / ctx->ar_slabs / ctx->TX);
        // (after the first pass of `seq` every pass reads the work array)
        if ((rc = run_passes(ctx, std::vector<PassSpec>{last}, ids, ns, seq.empty() ? src : work, seq.empty() ? src_by_sid : work_by_sid,
                             work, work_by_sid, o)))
            return rc;
        if ((rc = allreduce_slab(ctx, plane, k))) return rc;
    }
    return MSM_OK;
}

// Real-field Poisson solve (summed-density mode, dims == 3), in place on one real plane of n^3 doubles:
//   phi = F^-1[ c / (k^2 n^3) F[rho] ]          (simulation_object.rs:1066-1110; the reference transforms the real
//                                                density as a full complex array, :1071, :1105)
// The plane IS an array of n/2 x n x n complex pairs z[j] = rho[2j] + i rho[2j+1] along x.  R2C: an n/2-point pass +
// S_R2C gives the half spectrum k_x = 0 .. n/2-1 in place and X[n/2] in the Nyquist plane; y and z run as ordinary
// n-point passes over the half-spectrum grid (last forward + multiplier + first inverse fused: S_POISSON_INV), the
// Nyquist plane as a 2-D grid at fixed k_x; C2R: L_C2R + an n/2-point inverse pass puts phi back as the real plane.
// Half the bytes and flops of the complex solve: 5 x 16 = 80 B per cell instead of 160 (+ 48 of pack / unpack).
int poisson_real(msm_ctx* ctx, int plane, bool max_only, unsigned long long* maxbits, bool slabbed = false) {
    const int id0 = 0;
    XformOps o;
    o.gsz = 1;
    o.poisson_coef = ctx->cfg.poisson_coeff / pow((double)ctx->n, (double)ctx->dims);
    o.maxbits = maxbits;
    double2* r = real_plane(ctx, plane);
    int rc;
    std::vector<PassSpec> fwd{PassSpec{0, false, L_NONE, S_R2C, TG_HALF_X}, PassSpec{1, false, L_NONE, S_NONE, TG_HALF_YZ}};
    std::vector<PassSpec> nyq{PassSpec{0, false, L_NONE, S_NONE, TG_NYQ}};
    if (ctx->fuse) {
        fwd.push_back(PassSpec{2, false, L_NONE, S_POISSON_INV, TG_HALF_YZ});
        nyq.push_back(PassSpec{1, false, L_NONE, S_POISSON_INV, TG_NYQ});
    } else {
        fwd.push_back(PassSpec{2, false, L_NONE, S_POISSON, TG_HALF_YZ});
        fwd.push_back(PassSpec{2, true, L_NONE, S_NONE, TG_HALF_YZ});
        nyq.push_back(PassSpec{1, false, L_NONE, S_POISSON, TG_NYQ});
        nyq.push_back(PassSpec{1, true, L_NONE, S_NONE, TG_NYQ});
    }
    fwd.push_back(PassSpec{1, true, L_NONE, S_NONE, TG_HALF_YZ});
    nyq.push_back(PassSpec{0, true, L_NONE, S_NONE, TG_NYQ});
    if (slabbed) {
        // the density arrives range by range from the all-reduce stream (allreduce_slab): R2C follows it
        const long long rows = ctx->C / ctx->n;
        for (int k = 0; k < ctx->ar_slabs; ++k) {
            if (cudaStreamWaitEvent(ctx->st, ctx->ev_ar[k], 0) != cudaSuccess) return fail(ctx, MSM_E_CUDA, "cudaStreamWaitEvent failed");
            PassSpec ps = fwd[0];
            ps.tile0 = (int)(rows * k / ctx->ar_slabs / ctx->TXH);
            ps.tile_end = (int)(rows * (k + 1) / ctx->ar_slabs / ctx->TXH);
            if ((rc = run_passes(ctx, std::vector<PassSpec>{ps}, &id0, 1, r, 0, r, 0, o))) return rc;
        }
        fwd.erase(fwd.begin());
    }
    if ((rc = run_passes(ctx, fwd, &id0, 1, r, 0, r, 0, o))) return rc;
    if ((rc = run_passes(ctx, nyq, &id0, 1, ctx->nyq, 0, ctx->nyq, 0, o))) return rc;
    std::vector<PassSpec> back{PassSpec{0, true, L_C2R, max_only ? S_MAX : S_NONE, TG_HALF_X}};
    return run_passes(ctx, back, &id0, 1, r, 0, r, 0, o);
}

// Poisson solve in place on `nbuf` pair buffers starting at ctx->P:  phi = F^-1[ c/(k^2 n^d) F[rho] ]
// dims >= 2: the last forward pass, the multiply and the first inverse pass run as ONE kernel (S_POISSON_INV).
// x_fwd_done: the buffers already hold rho after the forward x pass (S_RHO_*_FX); x_inv_skip: leave out the inverse x
// pass (the kick kernel L_KICK_IX runs it on the fly).
int poisson(msm_ctx* ctx, int nbuf, bool max_only, unsigned long long* maxbits, bool x_fwd_done = false,
            bool x_inv_skip = false) {
    std::vector<int> ids(nbuf);
    for (int i = 0; i < nbuf; ++i) ids[i] = i;
    XformOps o;
    o.gsz = 1;
    o.poisson_coef = ctx->cfg.poisson_coeff / pow((double)ctx->n, (double)ctx->dims);
    o.maxbits = maxbits;
    const int d = ctx->dims;
    const bool fuse = d >= 2 && ctx->fuse;
    std::vector<PassSpec> seq;
    for (int a = x_fwd_done ? 1 : 0; a < d - 1; ++a) seq.push_back(PassSpec{a, false, L_NONE, S_NONE});
    if (fuse) {
        seq.push_back(PassSpec{d - 1, false, L_NONE, S_POISSON_INV});
    } else {
        seq.push_back(PassSpec{d - 1, false, L_NONE, S_POISSON});
        seq.push_back(PassSpec{d - 1, true, L_NONE, d == 1 ? (max_only ? S_MAX : S_NONE) : S_NONE});
    }
    for (int a = d - 2; a >= (x_inv_skip ? 1 : 0); --a)
        seq.push_back(PassSpec{a, true, L_NONE, (a == 0 && max_only) ? S_MAX : S_NONE});
    return run_passes(ctx, seq, ids.data(), nbuf, ctx->P, 0, ctx->P, 0, o);
}

bool full_fusion(const msm_ctx* ctx) {
    return ctx->fuse && ctx->dims >= 2 && ctx->cfg.coupling == MSM_COUPLING_INDEPENDENT;
}

// dt-potential of a chunk whose psi_k has ALREADY been taken through the inverse pass of the last axis into the
// scratch slots: remaining inverse passes -> rho (+ forward x) -> Poisson -> max|phi| into maxbits[pos...]
int dt_potential_tail(msm_ctx* ctx, const int* ids, int ns, unsigned long long* maxbits) {
    const int d = ctx->dims;
    XformOps o;
    o.gsz = 2;
    o.rho_coef = ctx->cfg.density_prefactor / pow((double)ctx->n, (double)d);   // un-normalised inverse passes
    std::vector<PassSpec> seq;
    for (int a = d - 2; a >= 1; --a) seq.push_back(PassSpec{a, true, L_NONE, S_NONE});
    seq.push_back(PassSpec{0, true, L_NONE, S_RHO_ONLY_FX});
    int rc = run_passes(ctx, seq, ids, ns, ctx->Tscr, 0, ctx->Tscr, 0, o);
    if (rc) return rc;
    return poisson(ctx, (ns + 1) / 2, true, maxbits, /*x_fwd_done=*/true, false);
}

// psi of stream s changes: its cached max|phi| is stale, and so is a value of it that still waits in `maxbits`
// after a non-blocking msm_step (the slot keeps its position in the list, marked -1)
void invalidate_stream(msm_ctx* ctx, int s) {
    ctx->pmax_valid[s] = 0;
    for (int& q : ctx->pmax_pending)
        if (q == s) q = -1;
}

int fetch_pending_pmax(msm_ctx* ctx) {
    if (ctx->pmax_pending.empty()) return MSM_OK;
    const size_t n = ctx->pmax_pending.size();
    const bool shared = ctx->cfg.coupling == MSM_COUPLING_SUMMED;   // one potential for all streams: max(|re|, |im|) slots 0, 1
    k_publish<<<1, 128, 0, ctx->st>>>(reinterpret_cast<const double*>(ctx->maxbits), ctx->h_scal + ctx->S + 2,
                                      (int)(shared ? 2 : n));
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->st) != cudaSuccess)
        return fail(ctx, MSM_E_CUDA, "fetching max|phi| failed");
    const double* h = ctx->h_scal + ctx->S + 2;
    for (size_t i = 0; i < n; ++i) {
        const int s = ctx->pmax_pending[i];
        if (s < 0) continue;   // invalidated since (invalidate_stream)
        ctx->pmax_cache[s] = shared ? (h[0] != h[0] || h[1] != h[1] ? NAN : fmax(h[0], h[1])) : h[i];
        ctx->pmax_valid[s] = 1;
    }
    ctx->pmax_pending.clear();
    return MSM_OK;
}

// RuntimeError::NanOrInf (utils/error.rs:10; the reference only checks in debug builds, utils/grid.rs:66-105): the max
// reduction works on bit patterns, so a NaN / Inf anywhere in phi arrives here as a non-finite max|phi|
int check_finite(msm_ctx* ctx, const std::vector<int>& ids, const double* v, const char* what) {
    for (int s : ids)
        if (!std::isfinite(v[s]))
            return fail(ctx, MSM_E_NAN, std::string("NaN or Inf in ") + what + " of stream " + std::to_string(s));
    return MSM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

const char* msm_version(void) { return "msm_b200 0.1 (sm_100a)"; }

const char* msm_strerror(int code) {
    switch (code) {
        case MSM_OK: return "ok";
        case MSM_E_ARG: return "bad argument";
        case MSM_E_CUDA: return "CUDA error";
        case MSM_E_NCCL: return "NCCL error";
        case MSM_E_ALIASING: return "Fourier aliasing above threshold";
        case MSM_E_NAN: return "NaN or Inf in grid";
        case MSM_E_STATE: return "invalid call sequence";
        case MSM_E_NOMEM: return "out of device memory";
        case MSM_E_IO: return "I/O error";
    }
    return "unknown error";
}

int msm_nccl_unique_id(void* out128) {
    if (!out128) return MSM_E_ARG;
    if (!g_nccl.load()) return fail(nullptr, MSM_E_NCCL, "cannot load libnccl.so.2");
    nccl_uid_t id;
    if (g_nccl.GetUniqueId(&id) != 0) return fail(nullptr, MSM_E_NCCL, "ncclGetUniqueId failed");
    memcpy(out128, &id, 128);
    return MSM_OK;
}

const char* msm_last_error(const msm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int msm_create(const msm_config* cfg, msm_ctx** out) {
    msm_ctx* ctx = nullptr;
    if (!cfg || !out) return fail(nullptr, MSM_E_ARG, "null argument");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(msm_config)) return fail(nullptr, MSM_E_ARG, "msm_config.struct_size mismatch");
    if (cfg->dims < 1 || cfg->dims > 3) return fail(nullptr, MSM_E_ARG, "dims must be 1, 2 or 3");
    const int n = cfg->size;
    if (n < 2 || n > 1024 || (n & (n - 1))) return fail(nullptr, MSM_E_ARG, "size must be a power of two in [2, 1024]");
    if (cfg->n_streams < 1) return fail(nullptr, MSM_E_ARG, "n_streams must be >= 1");
    if (!(cfg->dx > 0.0)) return fail(nullptr, MSM_E_ARG, "dx must be positive");
    if (cfg->coupling != MSM_COUPLING_INDEPENDENT && cfg->coupling != MSM_COUPLING_SUMMED)
        return fail(nullptr, MSM_E_ARG, "unknown coupling mode");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MSM_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                             " (msm_b200 has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MSM_E_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) return fail(nullptr, MSM_E_CUDA, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, MSM_E_CUDA, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, MSM_E_CUDA, cudaGetErrorString(e));

    ctx = new msm_ctx();
    ctx->cfg = *cfg;
    ctx->n = n;
    ctx->dims = cfg->dims;
    ctx->S = cfg->n_streams;
    ctx->T = plan_T(n);
    ctx->TX = plan_tx(n);
    ctx->C = 1;
    for (int d = 0; d < cfg->dims; ++d) ctx->C *= n;
    ctx->launcher = get_pass_launcher(n);
    if (const char* e = getenv("MSM_B200_XL")) ctx->xl = atoi(e) != 0;
    // 1-D grids have one line per tile; the other lines of the tile run as duplicates of it (fft_pass.cuh), which is
    // only race-free when they share a warp with it: generic thread mapping
    if (cfg->dims == 1) ctx->xl = false;
    if (const char* e = getenv("MSM_B200_FUSE")) ctx->fuse = atoi(e) != 0;
    if (const char* e = getenv("MSM_B200_PREFETCH")) ctx->l2_prefetch = atoi(e) != 0;
    if (const char* e = getenv("MSM_B200_TPC")) ctx->tiles_per_cta = std::max(1, atoi(e));
    if (const char* e = getenv("MSM_B200_TPCX")) ctx->tiles_per_cta_x = std::max(1, atoi(e));
    if (const char* e = getenv("MSM_B200_TPCF")) ctx->tiles_per_cta_fused = std::max(1, atoi(e));
    if (const char* e = getenv("MSM_B200_ILV")) ctx->interleave = std::max(1, atoi(e));
    ctx->lb = (cfg->dims == 3 && n >= 512) ? 4 : 0;
    if (const char* e = getenv("MSM_B200_LB")) ctx->lb = (cfg->dims == 3 && (1 << atoi(e)) <= n) ? std::max(0, atoi(e)) : 0;
    // the pass kernels address a thread's elements e = t + NT * j (NT = n / 8 threads per line) as a0 + j * step,
    // which needs NT to be a multiple of the block
    // (n / 16 at n = 512, where one kernel runs 16 points per thread: fft_pass.cuh plan_E)
    while (ctx->lb > 0 && (n < 16 || ((n == 512 ? n / 16 : n / 8) % (1 << ctx->lb)))) ctx->lb--;
    ctx->num_sms = prop.multiProcessorCount;
    int chunk = cfg->chunk_streams > 0 ? cfg->chunk_streams : 8;
    chunk = std::min(chunk, MAX_CHUNK);
    chunk = std::min(chunk, ctx->S + (ctx->S & 1));
    if (chunk > 1) chunk &= ~1;   // even, so that pairs never straddle chunks
    chunk = std::max(chunk, 1);
    ctx->chunk = chunk;
    ctx->pmax_valid.assign(ctx->S, 0);
    ctx->pmax_cache.assign(ctx->S, 0.0);
    ctx->in_k.assign(ctx->S, 0);
    ctx->has_psi.assign(ctx->S, 0);
    ctx->four_pi2 = (2.0 * M_PI) * (2.0 * M_PI);
    ctx->dv = pow(cfg->dx, (double)cfg->dims);   // dk = dx (simulation_object.rs:263), p_mass uses dk^dims (:1281-1285)

    // k grid (utils/fft.rs:100-120) and k2_max = max(spec_grid) (simulation_object.rs:274)
    ctx->h_ksq.resize(n);
    for (int i = 0; i < n; ++i) {
        const double ii = (i < n / 2) ? (double)i : (double)(i - n);
        const double k = ii / ((double)n * cfg->dx);
        ctx->h_ksq[i] = k * k;
    }
    {
        const double m = ctx->h_ksq[n / 2];
        double s = m;
        if (cfg->dims >= 2) s = s + m;
        if (cfg->dims >= 3) s = s + m;
        ctx->k2_max = s * ctx->four_pi2;
    }

    auto bail = [&](int code, const std::string& msg) {
        std::string m = msg;
        msm_destroy(ctx);
        return fail(nullptr, code, m);
    };
#define CUC(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess)                                                         \
            return bail(e_ == cudaErrorMemoryAllocation ? MSM_E_NOMEM : MSM_E_CUDA,    \
                        std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)
    CUC(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&ctx->dtab_done, cudaEventDisableTiming));
    const size_t cb = sizeof(double2) * (size_t)ctx->C;
    const int npair = (chunk + 1) / 2;
    CUC(cudaMalloc(&ctx->X, cb * ctx->S));
    CUC(cudaMalloc(&ctx->Tscr, cb * chunk));
    CUC(cudaMalloc(&ctx->P, cb * npair));
    CUC(cudaMalloc(&ctx->tw, sizeof(double2) * n));
    CUC(cudaMalloc(&ctx->dtab, sizeof(double2) * n * ctx->S));
    CUC(cudaMalloc(&ctx->ksq, sizeof(double) * n));
    const Geom glast = make_geom(ctx, ctx->dims - 1, std::min(ctx->T, 4));   // finest tiling any kernel uses
    ctx->ntiles_last = glast.ntiles;
    ctx->ntiles_used = make_geom(ctx, ctx->dims - 1, (ctx->dims == 1 && ctx->xl) ? ctx->TX : ctx->T).ntiles;
    CUC(cudaMalloc(&ctx->alias_partial, sizeof(double) * (size_t)ctx->S * glast.ntiles));
    CUC(cudaMalloc(&ctx->alias_out, sizeof(double) * ctx->S));
    CUC(cudaMalloc(&ctx->maxbits, sizeof(unsigned long long) * (ctx->S + 2)));
    CUC(cudaMalloc(&ctx->scratch_small, sizeof(double) * 4096));
    CUC(cudaHostAlloc(&ctx->h_coef, sizeof(double) * ctx->S, cudaHostAllocMapped));
    CUC(cudaHostAlloc(&ctx->h_ids, sizeof(int) * ctx->S, cudaHostAllocMapped));
    CUC(cudaHostAlloc(&ctx->h_scal, sizeof(double) * 2 * (ctx->S + 2), cudaHostAllocMapped));   // [0, S+2): alias / max ; [S+2, ..): eager max
    ctx->bytes = cb * (ctx->S + chunk + npair) + sizeof(double2) * n * (ctx->S + 1) + sizeof(double) * n +
                 sizeof(double) * (size_t)ctx->S * (glast.ntiles + 1) + 8 * (ctx->S + 2) + 8 * 4096;
    CUC(cudaMemsetAsync(ctx->alias_out, 0, sizeof(double) * ctx->S, ctx->st));
    CUC(cudaMemsetAsync(ctx->alias_partial, 0, sizeof(double) * (size_t)ctx->S * glast.ntiles, ctx->st));
    {
        // per-stage twiddle tables [k - 1][nu] = exp(-2 pi i k L nu / n)   (fft_pass.cuh: plan_tw_offset)
        auto stage_twiddles = [](int len) {
            std::vector<double2> tw(len, make_double2(1.0, 0.0));
            int rad[4];
            const int nst = plan_radices(len, rad);
            int off = 0, L = 1;
            for (int q = 0; q + 1 < nst; ++q) {
                const int M = len / (L * rad[q]);
                for (int k = 1; k < rad[q]; ++k)
                    for (int nu = 0; nu < M; ++nu) {
                        const long long j = (long long)k * L * nu;   // < len
                        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)len;
                        tw[off + (k - 1) * M + nu] = make_double2((double)cosl(a), (double)sinl(a));
                    }
                off += (rad[q] - 1) * M;
                L *= rad[q];
            }
            return tw;
        };
        const std::vector<double2> tw = stage_twiddles(n);
        CUC(cudaMemcpyAsync(ctx->tw, tw.data(), sizeof(double2) * n, cudaMemcpyHostToDevice, ctx->st));
        CUC(cudaMemcpyAsync(ctx->ksq, ctx->h_ksq.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->st));
        CUC(cudaStreamSynchronize(ctx->st));
        if (const char* e = getenv("MSM_B200_TMA")) {
            if (atoi(e) != 0 && n == 512 && cfg->dims == 3 && ctx->lb > 0) {
                void* bases[3] = {ctx->X, ctx->Tscr, ctx->P};
                const int slots[3] = {ctx->S, chunk, npair};
                ctx->tma = true;
                for (int a = 0; a < 3 && ctx->tma; ++a)
                    for (int ax = 1; ax <= 2 && ctx->tma; ++ax)
                        if (tma_make_map(&ctx->tma_map[a][ax - 1], bases[a], n, ctx->lb, slots[a], ax) != 0) ctx->tma = false;
                if (!ctx->tma) return bail(MSM_E_CUDA, "MSM_B200_TMA=1: cuTensorMapEncodeTiled failed");
            }
        }
        // Real-field Poisson solve for the summed density (one real field per rank; the independent mode gets the same
        // saving by packing two streams into one complex solve).  MSM_B200_REAL=0 keeps the complex solve (A/B timing).
        ctx->real_solve = cfg->coupling == MSM_COUPLING_SUMMED && cfg->dims == 3 && n >= 16 && ctx->xl;
        if (const char* e = getenv("MSM_B200_REAL")) ctx->real_solve = ctx->real_solve && atoi(e) != 0;
        if (ctx->real_solve) {
            const int nh = n / 2;
            ctx->launcher_half = get_pass_launcher(nh);
            ctx->TXH = ctx->xl ? plan_tx(nh) : plan_T(nh);
            const std::vector<double2> twh = stage_twiddles(nh);
            std::vector<double2> wr(nh);
            for (int k = 0; k < nh; ++k) {
                const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
                wr[k] = make_double2((double)cosl(a), (double)sinl(a));
            }
            CUC(cudaMalloc(&ctx->tw_half, sizeof(double2) * nh));
            CUC(cudaMalloc(&ctx->wreal, sizeof(double2) * nh));
            CUC(cudaMalloc(&ctx->nyq, sizeof(double2) * (size_t)n * n));
            ctx->bytes += sizeof(double2) * ((size_t)n * n + 2 * nh);
            CUC(cudaMemcpyAsync(ctx->tw_half, twh.data(), sizeof(double2) * nh, cudaMemcpyHostToDevice, ctx->st));
            CUC(cudaMemcpyAsync(ctx->wreal, wr.data(), sizeof(double2) * nh, cudaMemcpyHostToDevice, ctx->st));
            CUC(cudaStreamSynchronize(ctx->st));
        }
    }
    // the communicator: required by the summed-density mode, optional otherwise (msm_ensemble_allreduce)
    if (cfg->nranks > 1 && (cfg->coupling == MSM_COUPLING_SUMMED || cfg->nccl_unique_id)) {
        if (!cfg->nccl_unique_id) return bail(MSM_E_ARG, "nranks > 1 requires nccl_unique_id");
        if (!g_nccl.load()) return bail(MSM_E_NCCL, "cannot load libnccl.so.2");
        nccl_uid_t id;
        memcpy(&id, cfg->nccl_unique_id, 128);
        int rc = g_nccl.CommInitRank(&ctx->comm, cfg->nranks, id, cfg->rank);
        if (rc != 0)
            return bail(MSM_E_NCCL, std::string("ncclCommInitRank failed: ") +
                                        (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?") + " (code " +
                                        std::to_string(rc) + ")");
        if (ctx->real_solve) {
            // measured (profiles/README.md): 4 slabs are FASTER where NCCL reduces in the switch (NVLS; N = 4: 134.5 -> 132.9 ms,
            // N = 8: 71.6 -> 69.7 ms per step) and SLOWER on 2 ranks (ring over P2P: 249.6 -> 259.8 ms, the NCCL kernels compete
            // with the pass they hide behind).  MSM_B200_AR_SLABS overrides.
            ctx->ar_slabs = cfg->nranks >= 4 ? 4 : 1;
            if (const char* e = getenv("MSM_B200_AR_SLABS")) ctx->ar_slabs = std::max(1, std::min(8, atoi(e)));
            // a slab must be whole tiles of both x-pass kernels and whole CTAs: rows per slab multiple of 64
            while (ctx->ar_slabs > 1 && (((long long)n * n) % (64LL * ctx->ar_slabs))) ctx->ar_slabs /= 2;
            if (ctx->ar_slabs > 1) {
                int lo = 0, hi = 0;
                CUC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                CUC(cudaStreamCreateWithPriority(&ctx->comm_st, cudaStreamNonBlocking, hi));
                for (int i = 0; i < ctx->ar_slabs; ++i) {
                    CUC(cudaEventCreateWithFlags(&ctx->ev_slab[i], cudaEventDisableTiming));
                    CUC(cudaEventCreateWithFlags(&ctx->ev_ar[i], cudaEventDisableTiming));
                }
            }
        }
    }
#undef CUC
    *out = ctx;
    return MSM_OK;
}

void msm_destroy(msm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->st) cudaStreamSynchronize(ctx->st);
    if (ctx->copy_st) cudaStreamSynchronize(ctx->copy_st);
    prof_drain(ctx);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    cudaFree(ctx->X);
    cudaFree(ctx->Tscr);
    cudaFree(ctx->P);
    cudaFree(ctx->tw);
    cudaFree(ctx->tw_half);
    cudaFree(ctx->wreal);
    cudaFree(ctx->nyq);
    cudaFree(ctx->dtab);
    cudaFree(ctx->ksq);
    cudaFree(ctx->alias_partial);
    cudaFree(ctx->alias_out);
    cudaFree(ctx->maxbits);
    cudaFree(ctx->scratch_small);
    cudaFree(ctx->ic_base);
    cudaFree(ctx->ens_psi);
    cudaFree(ctx->ens_psik);
    cudaFree(ctx->ens_psi2);
    cudaFree(ctx->ens_psik2);
    if (ctx->h_coef) cudaFreeHost(ctx->h_coef);
    if (ctx->h_ids) cudaFreeHost(ctx->h_ids);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->dtab_done) cudaEventDestroy(ctx->dtab_done);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    }
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    if (ctx->up_st) {
        cudaStreamSynchronize(ctx->up_st);
        cudaStreamDestroy(ctx->up_st);
    }
    for (cudaEvent_t e : ctx->up_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_x_free) cudaEventDestroy(ctx->ev_x_free);
    cudaFree(ctx->up_stage);
    for (int i = 0; i < 2; ++i) {
        cudaFree(ctx->dl_stage[i]);
        if (ctx->dl_ready[i]) cudaEventDestroy(ctx->dl_ready[i]);
        if (ctx->dl_copied[i]) cudaEventDestroy(ctx->dl_copied[i]);
    }
    for (cudaEvent_t e : ctx->tk_ev)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 8; ++i) {
        if (ctx->ev_slab[i]) cudaEventDestroy(ctx->ev_slab[i]);
        if (ctx->ev_ar[i]) cudaEventDestroy(ctx->ev_ar[i]);
    }
    if (ctx->comm_st) {
        cudaStreamSynchronize(ctx->comm_st);
        cudaStreamDestroy(ctx->comm_st);
    }
    if (ctx->tm_a) cudaEventDestroy(ctx->tm_a);
    if (ctx->tm_b) cudaEventDestroy(ctx->tm_b);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}

int msm_device_bytes(const msm_ctx* ctx, uint64_t* bytes) {
    if (!ctx || !bytes) return MSM_E_ARG;
    *bytes = ctx->bytes;
    return MSM_OK;
}

int msm_synchronize(msm_ctx* ctx) {
    if (!ctx) return MSM_E_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaStreamSynchronize(ctx->st));
    return MSM_OK;
}

int msm_set_psi(msm_ctx* ctx, int32_t s, const double* psi) {
    if (!ctx || !psi || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_set_psi: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, s)) return rc;
    if (ctx->lb == 0) {
        CU(cudaMemcpyAsync(ctx->X + (size_t)s * ctx->C, psi, sizeof(double2) * (size_t)ctx->C, cudaMemcpyHostToDevice, ctx->st));
    } else {   // host layout is linear, device layout blocked: stage through scratch slot 0
        CU(cudaMemcpyAsync(ctx->Tscr, psi, sizeof(double2) * (size_t)ctx->C, cudaMemcpyHostToDevice, ctx->st));
        k_relayout<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->Tscr, ctx->X + (size_t)s * ctx->C, ctx->C, ctx->n, ctx->lb, 1);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(ctx->st));
    ctx->in_k[s] = 0;
    ctx->has_psi[s] = 1;
    invalidate_stream(ctx, s);
    return MSM_OK;
}

int msm_set_psi_planes(msm_ctx* ctx, int32_t s, const double* re, const double* im) {
    if (!ctx || !re || !im || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_set_psi_planes: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, s)) return rc;
    double* stage = reinterpret_cast<double*>(ctx->Tscr);   // 2*C doubles = one scratch slot
    CU(cudaMemcpyAsync(stage, re, sizeof(double) * (size_t)ctx->C, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(stage + ctx->C, im, sizeof(double) * (size_t)ctx->C, cudaMemcpyHostToDevice, ctx->st));
    k_interleave<<<grid_for(ctx->C), 256, 0, ctx->st>>>(stage, stage + ctx->C, ctx->X + (size_t)s * ctx->C, ctx->C, ctx->n, ctx->lb);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->st));
    ctx->in_k[s] = 0;
    ctx->has_psi[s] = 1;
    invalidate_stream(ctx, s);
    return MSM_OK;
}

// psi of stream s into scratch slot 0 (or X itself when no transform has happened yet); returns the device pointer
static int psi_on_device(msm_ctx* ctx, int s, const double2** out, int slot = 0) {
    if (!ctx->has_psi[s]) return fail(ctx, MSM_E_STATE, "stream has no wavefunction");
    if (int rc = wait_upload(ctx, s)) return rc;
    if (!ctx->in_k[s]) {
        *out = ctx->X + (size_t)s * ctx->C;
        return MSM_OK;
    }
    XformOps o;
    o.gsz = 1;
    o.sop_each = o.sop_last = S_SCALE;
    o.scale = 1.0 / sqrt((double)ctx->n);
    int id = s;
    // first pass reads X[s], every pass writes scratch slot 0 (work is indexed by local index)
    double2* work = ctx->Tscr + (size_t)slot * ctx->C;
    int rc = run_transform(ctx, true, &id, 1, ctx->X, 1, work, 0, o);
    if (rc) return rc;
    *out = work;
    return MSM_OK;
}

int msm_get_psi_interleaved(msm_ctx* ctx, int32_t s, double* out) {
    if (!ctx || !out || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_get_psi_interleaved: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    const double2* d = nullptr;
    int rc = psi_on_device(ctx, s, &d);
    if (rc) return rc;
    if (ctx->lb != 0) {   // back to the host's linear layout through the (idle) pair buffer
        k_relayout<<<grid_for(ctx->C), 256, 0, ctx->st>>>(d, ctx->P, ctx->C, ctx->n, ctx->lb, 0);
        ctx->launches++;
        CU(cudaGetLastError());
        d = ctx->P;
    }
    CU(cudaMemcpyAsync(out, d, sizeof(double2) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MSM_OK;
}

int msm_get_psi(msm_ctx* ctx, int32_t s, double* re, double* im) {
    if (!ctx || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_get_psi: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    const double2* d = nullptr;
    int rc = psi_on_device(ctx, s, &d);
    if (rc) return rc;
    double* planes = reinterpret_cast<double*>(ctx->P);   // 2*C doubles
    k_deinterleave<<<grid_for(ctx->C), 256, 0, ctx->st>>>(d, planes, planes + ctx->C, ctx->C, ctx->n, ctx->lb);
    ctx->launches++;
    CU(cudaGetLastError());
    if (re) CU(cudaMemcpyAsync(re, planes, sizeof(double) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st));
    if (im) CU(cudaMemcpyAsync(im, planes + ctx->C, sizeof(double) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MSM_OK;
}

int msm_get_psi_many(msm_ctx* ctx, int32_t n, const int32_t* streams, double* const* re, double* const* im) {
    if (!ctx || n < 0 || (n > 0 && (!streams || !re || !im))) return fail(ctx, MSM_E_ARG, "msm_get_psi_many: bad argument");
    for (int i = 0; i < n; ++i)
        if (streams[i] < 0 || streams[i] >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_get_psi_many: stream out of range");
    const int npair = (ctx->chunk + 1) / 2;
    if (npair < 2 || ctx->chunk < 2) {   // not enough staging for a pipeline: one stream after the other
        for (int i = 0; i < n; ++i) {
            int rc = msm_get_psi(ctx, streams[i], re[i], im[i]);
            if (rc) return rc;
        }
        return MSM_OK;
    }
    CU(cudaSetDevice(ctx->cfg.device));
    if (!ctx->copy_st) {
        CU(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
        }
    }
    const size_t pb = sizeof(double) * (size_t)ctx->C;
    for (int i = 0; i < n; ++i) {
        const int b = i & 1;
        if (i >= 2) CU(cudaStreamWaitEvent(ctx->st, ctx->ev_copied[b], 0));   // staging b has left the device
        const double2* d = nullptr;
        int rc = psi_on_device(ctx, streams[i], &d, b);
        if (rc) return rc;
        double* planes = reinterpret_cast<double*>(ctx->P + (size_t)b * ctx->C);
        k_deinterleave<<<grid_for(ctx->C), 256, 0, ctx->st>>>(d, planes, planes + ctx->C, ctx->C, ctx->n, ctx->lb);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev_ready[b], ctx->st));
        CU(cudaStreamWaitEvent(ctx->copy_st, ctx->ev_ready[b], 0));
        if (re[i]) CU(cudaMemcpyAsync(re[i], planes, pb, cudaMemcpyDeviceToHost, ctx->copy_st));
        if (im[i]) CU(cudaMemcpyAsync(im[i], planes + ctx->C, pb, cudaMemcpyDeviceToHost, ctx->copy_st));
        CU(cudaEventRecord(ctx->ev_copied[b], ctx->copy_st));
    }
    CU(cudaStreamSynchronize(ctx->copy_st));
    CU(cudaStreamSynchronize(ctx->st));
    return MSM_OK;
}

// ---- asynchronous transfers ------------------------------------------------------------------------------------
int msm_chunk_streams(const msm_ctx* ctx, int32_t* chunk) {
    if (!ctx || !chunk) return MSM_E_ARG;
    *chunk = ctx->chunk;
    return MSM_OK;
}

int msm_upload_begin(msm_ctx* ctx, int32_t s, const double* psi) {
    if (!ctx || !psi || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_upload_begin: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    const size_t cb = sizeof(double2) * (size_t)ctx->C;
    if (!ctx->up_st) {
        CU(cudaStreamCreateWithFlags(&ctx->up_st, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_x_free, cudaEventDisableTiming));
        ctx->up_ev.assign(ctx->S, nullptr);
        ctx->up_pending.assign(ctx->S, 0);
    }
    if (ctx->lb != 0 && !ctx->up_stage) {
        if (cudaMalloc(&ctx->up_stage, cb) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, MSM_E_NOMEM, "msm_upload_begin: no memory for the upload staging buffer");
        }
        ctx->bytes += cb;
    }
    if (!ctx->up_ev[s]) CU(cudaEventCreateWithFlags(&ctx->up_ev[s], cudaEventDisableTiming));
    // X[s] may still be read by work already enqueued on the compute stream: order the upload behind it
    CU(cudaEventRecord(ctx->ev_x_free, ctx->st));
    CU(cudaStreamWaitEvent(ctx->up_st, ctx->ev_x_free, 0));
    double2* dst = ctx->X + (size_t)s * ctx->C;
    if (ctx->lb == 0) {
        CU(cudaMemcpyAsync(dst, psi, cb, cudaMemcpyHostToDevice, ctx->up_st));
    } else {
        CU(cudaMemcpyAsync(ctx->up_stage, psi, cb, cudaMemcpyHostToDevice, ctx->up_st));
        k_relayout<<<grid_for(ctx->C), 256, 0, ctx->up_st>>>(ctx->up_stage, dst, ctx->C, ctx->n, ctx->lb, 1);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ctx->up_ev[s], ctx->up_st));
    ctx->up_pending[s] = 1;
    ctx->in_k[s] = 0;
    ctx->has_psi[s] = 1;
    invalidate_stream(ctx, s);
    return MSM_OK;
}

int msm_download_begin(msm_ctx* ctx, int32_t s, double* re, double* im) {
    if (!ctx || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_download_begin: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    const size_t pb = sizeof(double) * (size_t)ctx->C;
    if (!ctx->copy_st) {
        CU(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
        }
    }
    if (!ctx->dl_stage[0]) {
        for (int i = 0; i < 2; ++i) {
            if (cudaMalloc(&ctx->dl_stage[i], 2 * pb) != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, MSM_E_NOMEM, "msm_download_begin: no memory for the download staging buffers");
            }
            ctx->bytes += 2 * pb;
            CU(cudaEventCreateWithFlags(&ctx->dl_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->dl_copied[i], cudaEventDisableTiming));
        }
    }
    const int b = ctx->dl_next;
    ctx->dl_next ^= 1;
    if (ctx->dl_used[b]) CU(cudaStreamWaitEvent(ctx->st, ctx->dl_copied[b], 0));   // staging b has left the device
    const double2* d = nullptr;
    int rc = psi_on_device(ctx, s, &d, ctx->chunk >= 2 ? b : 0);
    if (rc) return rc;
    double* planes = ctx->dl_stage[b];
    k_deinterleave<<<grid_for(ctx->C), 256, 0, ctx->st>>>(d, planes, planes + ctx->C, ctx->C, ctx->n, ctx->lb);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->dl_ready[b], ctx->st));
    CU(cudaStreamWaitEvent(ctx->copy_st, ctx->dl_ready[b], 0));
    if (re) CU(cudaMemcpyAsync(re, planes, pb, cudaMemcpyDeviceToHost, ctx->copy_st));
    if (im) CU(cudaMemcpyAsync(im, planes + ctx->C, pb, cudaMemcpyDeviceToHost, ctx->copy_st));
    CU(cudaEventRecord(ctx->dl_copied[b], ctx->copy_st));
    ctx->dl_used[b] = 1;
    return MSM_OK;
}

int msm_download_ticket(msm_ctx* ctx, uint64_t* ticket) {
    if (!ctx || !ticket) return fail(ctx, MSM_E_ARG, "msm_download_ticket: bad argument");
    if (!ctx->copy_st) return fail(ctx, MSM_E_STATE, "msm_download_ticket: no download has been started");
    CU(cudaSetDevice(ctx->cfg.device));
    std::lock_guard<std::mutex> lock(ctx->tk_mu);
    cudaEvent_t& e = ctx->tk_ev[ctx->tk_next % 64];
    if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // (a waiter of ticket t - 64 whose event is re-recorded here waits for the newer record: later, never earlier)
    CU(cudaEventRecord(e, ctx->copy_st));
    *ticket = ctx->tk_next++;
    return MSM_OK;
}

int msm_download_wait(msm_ctx* ctx, uint64_t ticket) {
    if (!ctx) return MSM_E_ARG;
    cudaEvent_t e;
    {
        std::lock_guard<std::mutex> lock(ctx->tk_mu);
        if (ticket >= ctx->tk_next) return MSM_E_ARG;
        e = ctx->tk_ev[ticket % 64];
    }
    // no fail(): this runs on writer threads, ctx->err belongs to the thread that owns the context
    if (cudaSetDevice(ctx->cfg.device) != cudaSuccess || cudaEventSynchronize(e) != cudaSuccess) return MSM_E_CUDA;
    return MSM_OK;
}

int msm_host_alloc(msm_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out || bytes == 0) return fail(ctx, MSM_E_ARG, "msm_host_alloc: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    if (cudaMallocHost(out, bytes) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(ctx, MSM_E_NOMEM, "msm_host_alloc: cudaMallocHost failed");
    }
    return MSM_OK;
}

int msm_host_free(msm_ctx* ctx, void* p) {
    if (!ctx) return MSM_E_ARG;
    if (p) cudaFreeHost(p);
    return MSM_OK;
}

int msm_transfers_wait(msm_ctx* ctx) {
    if (!ctx) return MSM_E_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    if (ctx->up_st) CU(cudaStreamSynchronize(ctx->up_st));
    if (ctx->copy_st) {
        // downloads are produced on the compute stream: everything enqueued there has to reach the copy stream first
        CU(cudaStreamSynchronize(ctx->st));
        CU(cudaStreamSynchronize(ctx->copy_st));
    }
    return MSM_OK;
}

int msm_get_psik_interleaved(msm_ctx* ctx, int32_t s, double* out) {
    if (!ctx || !out || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_get_psik_interleaved: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_kspace(ctx, std::vector<int>{s});
    if (rc) return rc;
    const double2* d = ctx->X + (size_t)s * ctx->C;
    if (ctx->lb != 0) {
        k_relayout<<<grid_for(ctx->C), 256, 0, ctx->st>>>(d, ctx->Tscr, ctx->C, ctx->n, ctx->lb, 0);
        ctx->launches++;
        CU(cudaGetLastError());
        d = ctx->Tscr;
    }
    CU(cudaMemcpyAsync(out, d, sizeof(double2) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MSM_OK;
}

// rho of the listed streams from psi_k (out of place through scratch), into the pair buffers
static int density_from_psik(msm_ctx* ctx, const int* ids, int ns, bool summed, bool accumulate) {
    XformOps o;
    o.sop_last = S_RHO_ONLY;
    const double n3 = pow((double)ctx->n, (double)ctx->dims);
    const int sg = ctx->cfg.n_streams_global > 0 ? ctx->cfg.n_streams_global : ctx->S;
    // un-normalised inverse passes: |psi|^2 = |u|^2 / n^d
    o.rho_coef = ctx->cfg.density_prefactor / n3 / (summed ? (double)sg : 1.0);
    o.gsz = summed ? ns : 2;
    o.p_summed = summed ? (ctx->real_solve ? 2 : 1) : 0;
    if (summed && ctx->real_solve) o.pbuf = real_plane(ctx, 1);   // the dt-potential lives in plane 1
    o.rho_accumulate = accumulate ? 1 : 0;
    return run_transform(ctx, true, ids, ns, ctx->X, 1, ctx->Tscr, 0, o);
}

// summed mode: potential of the density accumulated in plane 1 (real-field solve) / pair buffer 0 (complex fallback)
static int summed_potential(msm_ctx* ctx, bool max_only, unsigned long long* maxbits) {
    int rc = allreduce_rho(ctx, 1);
    if (rc) return rc;
    return ctx->real_solve ? poisson_real(ctx, 1, max_only, maxbits) : poisson(ctx, 1, max_only, maxbits);
}

static int potential_max_of(msm_ctx* ctx, const std::vector<int>& ids, double* out) {
    int rc = fetch_pending_pmax(ctx);
    if (rc) return rc;
    {
        bool all = true;
        for (int s : ids) all = all && ctx->pmax_valid[s];
        if (all) {   // computed eagerly by the last msm_step
            for (int s : ids) out[s] = ctx->pmax_cache[s];
            return MSM_OK;
        }
    }
    rc = ensure_kspace(ctx, ids);
    if (rc) return rc;
    const bool summed = ctx->cfg.coupling == MSM_COUPLING_SUMMED;
    CU(cudaMemsetAsync(ctx->maxbits, 0, sizeof(unsigned long long) * (ctx->S + 2), ctx->st));
    if (full_fusion(ctx)) {
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            XformOps o;   // inverse pass of the last axis, out of place into the scratch slots
            std::vector<PassSpec> first{PassSpec{ctx->dims - 1, true, L_NONE, S_NONE}};
            rc = run_passes(ctx, first, &ids[i], ns, ctx->X, 1, ctx->Tscr, 0, o);
            if (rc) return rc;
            rc = dt_potential_tail(ctx, &ids[i], ns, ctx->maxbits + i);
            if (rc) return rc;
        }
    } else if (!summed) {
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            rc = density_from_psik(ctx, &ids[i], ns, false, false);
            if (rc) return rc;
            rc = poisson(ctx, (ns + 1) / 2, true, ctx->maxbits + i);   // maxbits[position in ids]
            if (rc) return rc;
        }
    } else {
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            rc = density_from_psik(ctx, &ids[i], ns, true, i > 0);
            if (rc) return rc;
        }
        rc = summed_potential(ctx, true, ctx->maxbits);
        if (rc) return rc;
    }
    k_publish<<<1, 128, 0, ctx->st>>>(reinterpret_cast<const double*>(ctx->maxbits), ctx->h_scal,
                                      (int)std::max<size_t>(2, ids.size()));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->st));
    const double* h = ctx->h_scal;
    const double shared_max = (h[0] != h[0] || h[1] != h[1]) ? NAN : fmax(h[0], h[1]);   // summed: max(|re|, |im|) slots
    for (size_t i = 0; i < ids.size(); ++i) {
        out[ids[i]] = summed ? shared_max : ctx->h_scal[i];
        ctx->pmax_cache[ids[i]] = out[ids[i]];
        ctx->pmax_valid[ids[i]] = 1;
    }
    return MSM_OK;
}

int msm_potential_max(msm_ctx* ctx, const int32_t* active, double* out) {
    if (!ctx || !out) return fail(ctx, MSM_E_ARG, "msm_potential_max: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    std::vector<int> ids = active_list(ctx, active);
    if (ids.empty()) return MSM_OK;
    int rc = potential_max_of(ctx, ids, out);
    if (rc) return rc;
    // NaN / Inf guard.  Two independent streams share one complex pair buffer (rho_a + i rho_b), so a non-finite stream
    // also poisons its partner's max|phi|: every suspect is solved again on its own before it is blamed.
    std::vector<int> bad;
    for (int s : ids)
        if (!std::isfinite(out[s])) bad.push_back(s);
    if (bad.size() > 1 && ctx->cfg.coupling == MSM_COUPLING_INDEPENDENT) {
        for (int s : bad) {
            ctx->pmax_valid[s] = 0;
            if ((rc = potential_max_of(ctx, std::vector<int>{s}, out))) return rc;
        }
    }
    return check_finite(ctx, ids, out, "the potential");
}

int msm_get_potential(msm_ctx* ctx, int32_t s, double* phi) {
    if (!ctx || !phi || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_get_potential: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    const bool summed = ctx->cfg.coupling == MSM_COUPLING_SUMMED;
    int rc;
    if (!summed) {
        rc = ensure_kspace(ctx, std::vector<int>{s});
        if (rc) return rc;
        int id = s;
        rc = density_from_psik(ctx, &id, 1, false, false);
        if (rc) return rc;
        rc = poisson(ctx, 1, false, nullptr);
    } else {
        std::vector<int> ids = active_list(ctx, nullptr);
        rc = ensure_kspace(ctx, ids);
        if (rc) return rc;
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            rc = density_from_psik(ctx, &ids[i], ns, true, i > 0);
            if (rc) return rc;
        }
        rc = summed_potential(ctx, false, nullptr);
    }
    if (rc) return rc;
    double* stage = reinterpret_cast<double*>(ctx->Tscr);
    if (summed && ctx->real_solve)   // phi is a real plane already
        k_extract_real<<<grid_for(ctx->C), 256, 0, ctx->st>>>(reinterpret_cast<const double*>(real_plane(ctx, 1)), stage, ctx->C,
                                                              ctx->n, ctx->lb);
    else
        k_extract<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->P, stage, ctx->C, ctx->n, ctx->lb, 0);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(phi, stage, sizeof(double) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MSM_OK;
}

int msm_read_alias(msm_ctx* ctx, double* alias_mass) {
    if (!ctx || !alias_mass) return fail(ctx, MSM_E_ARG, "msm_read_alias: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    k_publish<<<1, 128, 0, ctx->st>>>(ctx->alias_out, ctx->h_scal, ctx->S);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->st));
    for (int s = 0; s < ctx->S; ++s) alias_mass[s] = ctx->h_scal[s];
    return MSM_OK;
}

int msm_step(msm_ctx* ctx, const int32_t* active, const double* drift, const double* kick, double* alias_mass) {
    if (!ctx || !drift || !kick) return fail(ctx, MSM_E_ARG, "msm_step: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    std::vector<int> ids = active_list(ctx, active);
    if (ids.empty()) return MSM_OK;
    int rc = ensure_kspace(ctx, ids);
    if (rc) return rc;
    const int n = ctx->n;
    const bool summed = ctx->cfg.coupling == MSM_COUPLING_SUMMED;

    // per-axis drift tables: exp(-i c k^2) = prod_axes exp(-i c (2 pi k_m)^2); n^(-1/2) per pass = unitary scale
    CU(cudaEventSynchronize(ctx->dtab_done));   // the previous step's build kernel has read the coefficients
    bool same_drift = true;
    for (size_t i = 0; i < ids.size(); ++i) {
        ctx->h_ids[i] = ids[i];
        ctx->h_coef[ids[i]] = drift[ids[i]];
        same_drift = same_drift && drift[ids[i]] == drift[ids[0]];
    }
    k_build_dtab<<<(int)ids.size(), 128, 0, ctx->st>>>(ctx->h_coef, ctx->h_ids, ctx->ksq, ctx->dtab, n, ctx->four_pi2,
                                                     1.0 / sqrt((double)n));
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->dtab_done, ctx->st));

    const int sg = ctx->cfg.n_streams_global > 0 ? ctx->cfg.n_streams_global : ctx->S;
    const bool real = summed && ctx->real_solve;
    const bool slabbed = real && ctx->cfg.nranks > 1 && ctx->ar_slabs > 1;   // all-reduce pipelined behind the last x pass
    auto drift_inverse = [&](const int* cid, int ns, bool accumulate, bool completes_rho = false) {
        XformOps o;   // psi_k * drift -> psi (in place), rho out
        o.lop_each = L_DRIFT;
        o.sop_last = S_RHO_KEEP;
        o.rho_coef = ctx->cfg.density_prefactor / (summed ? (double)sg : 1.0);
        o.gsz = summed ? ns : 2;
        o.p_summed = summed ? (real ? 2 : 1) : 0;
        if (real) o.pbuf = real_plane(ctx, 0);
        o.rho_accumulate = accumulate ? 1 : 0;
        o.dtab_shared = summed && same_drift;
        if (slabbed && completes_rho)
            return run_passes_slabbed(ctx, transform_seq(ctx, true, o), cid, ns, ctx->X, 1, ctx->X, 1, o, 0);
        return run_transform(ctx, true, cid, ns, ctx->X, 1, ctx->X, 1, o);
    };
    auto kick_forward = [&](const int* cid, int ns, bool start_next_potential) {
        XformOps o;   // psi * kick -> psi_k * drift (in place), alias partials out
        double kk[MAX_CHUNK];
        for (int i = 0; i < ns; ++i) kk[i] = kick[cid[i]];
        o.lop_first = L_KICK;
        o.sop_each = S_DRIFT;
        // the last pass can also run the first inverse pass of the next dt-potential into the scratch slots
        o.sop_last = start_next_potential ? S_DRIFT_ALIAS_IZ : S_DRIFT_ALIAS;
        o.dst2 = ctx->Tscr;
        o.kick = kk;
        o.gsz = summed ? ns : 2;
        o.p_summed = summed ? (real ? 2 : 1) : 0;
        if (real) o.pbuf = real_plane(ctx, 0);
        o.dtab_shared = summed && same_drift;
        return run_transform(ctx, false, cid, ns, ctx->X, 1, ctx->X, 1, o);
    };
    for (int s : ids) ctx->pmax_valid[s] = 0;
    ctx->pmax_pending.clear();
    if (full_fusion(ctx)) {
        // 14 launches per chunk instead of 21, 384 instead of 480 algorithmic bytes per cell-update
        const int d = ctx->dims;
        CU(cudaMemsetAsync(ctx->maxbits, 0, sizeof(unsigned long long) * (ctx->S + 2), ctx->st));
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            const int* cid = &ids[i];
            {   // drift + inverse (in place); the x pass also forward-transforms rho_a + i rho_b into the pair buffer
                XformOps o;
                o.gsz = 2;
                o.rho_coef = ctx->cfg.density_prefactor;
                std::vector<PassSpec> seq;
                for (int a = d - 1; a >= 1; --a) seq.push_back(PassSpec{a, true, L_DRIFT, S_NONE});
                seq.push_back(PassSpec{0, true, L_DRIFT, S_RHO_KEEP_FX});
                if ((rc = run_passes(ctx, seq, cid, ns, ctx->X, 1, ctx->X, 1, o))) return rc;
            }
            if ((rc = poisson(ctx, (ns + 1) / 2, false, nullptr, /*x_fwd_done=*/true, /*x_inv_skip=*/true))) return rc;
            {   // (inverse x of phi) + kick + forward + drift + alias; the last pass also starts the next dt-potential
                XformOps o;
                double kk[MAX_CHUNK];
                for (int j = 0; j < ns; ++j) kk[j] = kick[cid[j]];
                o.gsz = 2;
                o.kick = kk;
                o.dst2 = ctx->Tscr;
                std::vector<PassSpec> seq;
                seq.push_back(PassSpec{0, false, L_KICK_IX, S_DRIFT});
                for (int a = 1; a < d - 1; ++a) seq.push_back(PassSpec{a, false, L_NONE, S_DRIFT});
                seq.push_back(PassSpec{d - 1, false, L_NONE, S_DRIFT_ALIAS_IZ});
                if ((rc = run_passes(ctx, seq, cid, ns, ctx->X, 1, ctx->X, 1, o))) return rc;
            }
            // potential of the NEW psi_k (the reference computes it at the start of the next update(), :497)
            if ((rc = dt_potential_tail(ctx, cid, ns, ctx->maxbits + i))) return rc;
        }
        ctx->pmax_pending = ids;
    } else if (!summed) {
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            if ((rc = drift_inverse(&ids[i], ns, false))) return rc;
            if ((rc = poisson(ctx, (ns + 1) / 2, false, nullptr))) return rc;
            if ((rc = kick_forward(&ids[i], ns, false))) return rc;
        }
    } else if (real) {
        // Summed density, real-field solve (dims == 3).  rho accumulates over the local streams into real plane 0, is
        // summed over the ranks in place, and one replicated half-spectrum solve leaves phi in the same plane.  While the
        // chunks are kicked (all reading plane 0), their last forward pass already starts the NEXT step's dt-potential
        // (first inverse pass into the scratch slots); its density accumulates into plane 1, which is reduced and solved
        // for max|phi| only.  Per step and rank: 2 x 8 B/cell over NVLink and 2 x 80 B/cell of replicated solve.
        const int d = ctx->dims;
        const bool all_streams = ids.size() == (size_t)ctx->S;
        const bool eager = ctx->fuse && all_streams;   // the shared potential needs every stream's density
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            if ((rc = drift_inverse(&ids[i], ns, i > 0, i + ctx->chunk >= ids.size()))) return rc;
        }
        if (!slabbed && (rc = allreduce_rho(ctx, 0))) return rc;
        if ((rc = poisson_real(ctx, 0, false, nullptr, slabbed))) return rc;
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            if ((rc = kick_forward(&ids[i], ns, eager))) return rc;
            if (eager) {   // rest of psi_k -> psi of this chunk (scratch slots), |psi|^2 into plane 1
                XformOps o;
                o.gsz = ns;
                o.p_summed = 2;
                o.pbuf = real_plane(ctx, 1);
                o.rho_accumulate = i > 0;
                o.rho_coef = ctx->cfg.density_prefactor / pow((double)n, (double)d) / (double)sg;   // un-normalised passes
                std::vector<PassSpec> seq;
                for (int a = d - 2; a >= 1; --a) seq.push_back(PassSpec{a, true, L_NONE, S_NONE});
                seq.push_back(PassSpec{0, true, L_NONE, S_RHO_ONLY});
                const bool completes = all_streams && slabbed && i + ctx->chunk >= ids.size();
                rc = completes ? run_passes_slabbed(ctx, seq, &ids[i], ns, ctx->Tscr, 0, ctx->Tscr, 0, o, 1)
                               : run_passes(ctx, seq, &ids[i], ns, ctx->Tscr, 0, ctx->Tscr, 0, o);
                if (rc) return rc;
            }
        }
        if (eager) {
            CU(cudaMemsetAsync(ctx->maxbits, 0, sizeof(unsigned long long) * (ctx->S + 2), ctx->st));
            if (slabbed) rc = poisson_real(ctx, 1, true, ctx->maxbits, true);
            else rc = summed_potential(ctx, true, ctx->maxbits);
            if (rc) return rc;
            ctx->pmax_pending = ids;
        }
    } else {
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            if ((rc = drift_inverse(&ids[i], ns, i > 0))) return rc;
        }
        if ((rc = allreduce_rho(ctx, 0))) return rc;
        if ((rc = poisson(ctx, 1, false, nullptr))) return rc;
        for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
            const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
            if ((rc = kick_forward(&ids[i], ns, false))) return rc;
        }
    }
    {
        ProfScope ps(ctx, "alias_reduce", 8.0 * ctx->S * ctx->ntiles_last);
        k_alias_reduce<<<ctx->S, 256, 0, ctx->st>>>(ctx->alias_partial, ctx->alias_out, ctx->alias_count, ctx->ntiles_used, ctx->dv);
    }
    ctx->launches++;
    CU(cudaGetLastError());
    if (alias_mass) {
        k_publish<<<1, 128, 0, ctx->st>>>(ctx->alias_out, ctx->h_scal, ctx->S);
        CU(cudaGetLastError());
        if ((rc = fetch_pending_pmax(ctx))) return rc;   // same sync
        CU(cudaStreamSynchronize(ctx->st));
        for (int s : ids) alias_mass[s] = ctx->h_scal[s];
        if ((rc = check_finite(ctx, ids, alias_mass, "psi_k (alias sum)"))) return rc;
    }
    return MSM_OK;
}

int msm_fft(int32_t device, int32_t dims, int32_t size, int32_t inverse, int32_t batch, double* data) {
    if (!data || batch < 1) return fail(nullptr, MSM_E_ARG, "msm_fft: bad argument");
    msm_config cfg{};
    cfg.struct_size = sizeof(msm_config);
    cfg.dims = dims;
    cfg.size = size;
    cfg.n_streams = batch;
    cfg.device = device;
    cfg.dx = 1.0;
    cfg.nranks = 1;
    msm_ctx* ctx = nullptr;
    int rc = msm_create(&cfg, &ctx);
    if (rc) return rc;
    auto done = [&](int code) {
        if (code) g_create_error = ctx->err;
        msm_destroy(ctx);
        return code;
    };
    const size_t cb = sizeof(double2) * (size_t)ctx->C;
    for (int b = 0; b < batch; ++b) {
        rc = msm_set_psi(ctx, b, data + 2 * (size_t)ctx->C * b);   // linear host layout -> device layout
        if (rc) return done(rc);
    }
    std::vector<int> ids(batch);
    for (int i = 0; i < batch; ++i) ids[i] = i;
    for (int i = 0; i < batch; i += ctx->chunk) {
        const int ns = std::min(ctx->chunk, batch - i);
        XformOps o;
        o.sop_each = o.sop_last = S_SCALE;
        o.scale = 1.0 / sqrt((double)size);
        rc = run_transform(ctx, inverse != 0, &ids[i], ns, ctx->X, 1, ctx->X, 1, o);
        if (rc) return done(rc);
    }
    for (int b = 0; b < batch; ++b) {
        const double2* d = ctx->X + (size_t)b * ctx->C;
        if (ctx->lb != 0) {
            k_relayout<<<grid_for(ctx->C), 256, 0, ctx->st>>>(d, ctx->Tscr, ctx->C, ctx->n, ctx->lb, 0);
            d = ctx->Tscr;
        }
        if (cudaMemcpyAsync(data + 2 * (size_t)ctx->C * b, d, cb, cudaMemcpyDeviceToHost, ctx->st) != cudaSuccess) return done(MSM_E_CUDA);
        if (cudaStreamSynchronize(ctx->st) != cudaSuccess) return done(MSM_E_CUDA);
    }
    if (cudaStreamSynchronize(ctx->st) != cudaSuccess) {
        ctx->err = cudaGetErrorString(cudaGetLastError());
        return done(MSM_E_CUDA);
    }
    return done(MSM_OK);
}

int msm_spec_grid(int32_t device, int32_t dims, int32_t size, double dx, double* out) {
    if (!out) return fail(nullptr, MSM_E_ARG, "msm_spec_grid: bad argument");
    msm_config cfg{};
    cfg.struct_size = sizeof(msm_config);
    cfg.dims = dims;
    cfg.size = size;
    cfg.n_streams = 1;
    cfg.device = device;
    cfg.dx = dx;
    cfg.nranks = 1;
    msm_ctx* ctx = nullptr;
    int rc = msm_create(&cfg, &ctx);
    if (rc) return rc;
    double* d = reinterpret_cast<double*>(ctx->Tscr);
    k_spec_grid<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->ksq, d, size, dims, ctx->four_pi2);
    cudaError_t e = cudaMemcpyAsync(out, d, sizeof(double) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
    if (e != cudaSuccess) g_create_error = cudaGetErrorString(e);
    msm_destroy(ctx);
    return e == cudaSuccess ? MSM_OK : MSM_E_CUDA;
}

// ---- on-device initial conditions -------------------------------------------------------------------------
static int normalize_on_device(msm_ctx* ctx, double2* psi) {
    // normalize (utils/grid.rs:11-33): psi *= sqrt(dx^-dims / sum|psi|^2)
    const int nb = 1024;
    k_norm2_partial<<<nb, 256, 0, ctx->st>>>(psi, ctx->C, ctx->scratch_small);
    ctx->launches++;
    CU(cudaGetLastError());
    std::vector<double> part(nb);
    CU(cudaMemcpyAsync(part.data(), ctx->scratch_small, sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    double tot = 0.0;
    for (double v : part) tot += v;
    const double f = sqrt(pow(ctx->cfg.dx, -(double)ctx->dims) / tot);
    k_scale_real<<<grid_for(ctx->C), 256, 0, ctx->st>>>(psi, ctx->C, f);
    ctx->launches++;
    CU(cudaGetLastError());
    return MSM_OK;
}

int msm_ic_cold_gauss(msm_ctx* ctx, int32_t s, const double* mean, const double* std) {
    if (!ctx || !mean || !std || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_ic_cold_gauss: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, s)) return rc;
    const int n = ctx->n, d = ctx->dims;
    const double dx = ctx->cfg.dx;
    // 1-D factors, each normalised with dx^dims as the reference does (ics.rs:91,110,132)
    std::vector<double> g(3 * n, 1.0);
    for (int a = 0; a < d; ++a) {
        double norm = 0.0;
        for (int i = 0; i < n; ++i) {
            const double x = (double)(2 * i + 1) * dx / 2.0;
            const double v = exp(-0.5 * pow((x - mean[a]) / std[a], 2.0));
            g[a * n + i] = v;
            norm += v * v;
        }
        const double f = sqrt(pow(dx, -(double)d) / norm);
        for (int i = 0; i < n; ++i) g[a * n + i] *= f;
    }
    if (3 * n > 4096) return fail(ctx, MSM_E_ARG, "size too large for the IC staging buffer");
    CU(cudaMemcpyAsync(ctx->scratch_small, g.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    double2* psi = ctx->X + (size_t)s * ctx->C;
    k_ic_separable<<<grid_for(ctx->C), 256, 0, ctx->st>>>(psi, ctx->scratch_small, n, d, 1.0, ctx->lb);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->st));
    int rc = normalize_on_device(ctx, psi);   // ics.rs:142
    if (rc) return rc;
    ctx->in_k[s] = 0;
    ctx->has_psi[s] = 1;
    invalidate_stream(ctx, s);
    return MSM_OK;
}

int msm_ic_cold_gauss_kspace(msm_ctx* ctx, int32_t s, const double* mean, const double* std, uint64_t phase_seed) {
    if (!ctx || !mean || !std || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_ic_cold_gauss_kspace: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, s)) return rc;
    const int n = ctx->n, d = ctx->dims;
    const double dk = ctx->cfg.dx;   // dk = dx in the reference (simulation_object.rs:263)
    if (3 * n > 4096) return fail(ctx, MSM_E_ARG, "size too large for the IC staging buffer");
    std::vector<double> g(3 * n, 1.0);
    for (int a = 0; a < d; ++a) {   // 1-D factors on the k grid, each normalised with dk^dims (ics.rs:336-390)
        double norm = 0.0;
        for (int i = 0; i < n; ++i) {
            const double k = ((i < n / 2) ? (double)i : (double)(i - n)) / ((double)n * ctx->cfg.dx);
            const double v = exp(-0.5 * pow((k - mean[a]) / std[a], 2.0));
            g[a * n + i] = v;
            norm += v * v;
        }
        const double f = sqrt(pow(dk, -(double)d) / norm);
        for (int i = 0; i < n; ++i) g[a * n + i] *= f;
    }
    CU(cudaMemcpyAsync(ctx->scratch_small, g.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    double2* psi = ctx->X + (size_t)s * ctx->C;
    k_ic_separable<<<grid_for(ctx->C), 256, 0, ctx->st>>>(psi, ctx->scratch_small, n, d, 1.0, ctx->lb);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->st));
    int rc = normalize_on_device(ctx, psi);   // ics.rs:395
    if (rc) return rc;
    k_random_phase<<<grid_for(ctx->C), 256, 0, ctx->st>>>(psi, ctx->C, phase_seed, 7u, n, ctx->lb);   // ics.rs:407-423
    ctx->launches++;
    CU(cudaGetLastError());
    XformOps o;   // forward_inplace (ics.rs:425), unitary
    o.gsz = 1;
    o.sop_each = o.sop_last = S_SCALE;
    o.scale = 1.0 / sqrt((double)n);
    int id = s;
    rc = run_transform(ctx, false, &id, 1, ctx->X, 1, ctx->X, 1, o);
    if (rc) return rc;
    ctx->in_k[s] = 0;   // what X holds now is the reference's spatial psi
    ctx->has_psi[s] = 1;
    invalidate_stream(ctx, s);
    return MSM_OK;
}

int msm_ic_spherical_tophat(msm_ctx* ctx, int32_t s, double axis_length, double radius, double delta, double slope) {
    if (!ctx || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_ic_spherical_tophat: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, s)) return rc;
    double2* psi = ctx->X + (size_t)s * ctx->C;
    const double dx = axis_length / (double)ctx->n;   // ics.rs:203 (axis length, not the comoving box)
    k_ic_tophat<<<grid_for(ctx->C), 256, 0, ctx->st>>>(psi, ctx->n, ctx->dims, dx, axis_length / 2.0, radius, delta, slope, ctx->lb);
    ctx->launches++;
    CU(cudaGetLastError());
    int rc = normalize_on_device(ctx, psi);   // ics.rs:261
    if (rc) return rc;
    ctx->in_k[s] = 0;
    ctx->has_psi[s] = 1;
    invalidate_stream(ctx, s);
    return MSM_OK;
}

int msm_ic_copy(msm_ctx* ctx, int32_t dst, int32_t src) {
    if (!ctx || dst < 0 || dst >= ctx->S || src < 0 || src >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_ic_copy: bad argument");
    if (!ctx->has_psi[src]) return fail(ctx, MSM_E_STATE, "msm_ic_copy: source stream is empty");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, src)) return rc;
    if (int rc = wait_upload(ctx, dst)) return rc;
    if (dst != src)
        CU(cudaMemcpyAsync(ctx->X + (size_t)dst * ctx->C, ctx->X + (size_t)src * ctx->C, sizeof(double2) * (size_t)ctx->C,
                           cudaMemcpyDeviceToDevice, ctx->st));
    ctx->in_k[dst] = ctx->in_k[src];
    ctx->has_psi[dst] = 1;
    invalidate_stream(ctx, dst);
    return MSM_OK;
}

int msm_ic_store(msm_ctx* ctx, int32_t s) {
    if (!ctx || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_ic_store: bad argument");
    if (!ctx->has_psi[s]) return fail(ctx, MSM_E_STATE, "msm_ic_store: stream is empty");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, s)) return rc;
    const size_t cb = sizeof(double2) * (size_t)ctx->C;
    if (!ctx->ic_base) {
        if (cudaMalloc(&ctx->ic_base, cb) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, MSM_E_NOMEM, "msm_ic_store: no memory for the saved wavefunction");
        }
        ctx->bytes += cb;
    }
    CU(cudaMemcpyAsync(ctx->ic_base, ctx->X + (size_t)s * ctx->C, cb, cudaMemcpyDeviceToDevice, ctx->st));
    ctx->ic_base_in_k = ctx->in_k[s];
    return MSM_OK;
}

int msm_ic_load(msm_ctx* ctx, int32_t s) {
    if (!ctx || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_ic_load: bad argument");
    if (!ctx->ic_base) return fail(ctx, MSM_E_STATE, "msm_ic_load: nothing stored (call msm_ic_store first)");
    CU(cudaSetDevice(ctx->cfg.device));
    if (int rc = wait_upload(ctx, s)) return rc;
    CU(cudaMemcpyAsync(ctx->X + (size_t)s * ctx->C, ctx->ic_base, sizeof(double2) * (size_t)ctx->C, cudaMemcpyDeviceToDevice, ctx->st));
    ctx->in_k[s] = ctx->ic_base_in_k;
    ctx->has_psi[s] = 1;
    invalidate_stream(ctx, s);
    return MSM_OK;
}

int msm_sample_perturbation(msm_ctx* ctx, int32_t s, int32_t scheme, uint64_t seed, double n_tot) {
    if (!ctx || s < 0 || s >= ctx->S) return fail(ctx, MSM_E_ARG, "msm_sample_perturbation: bad argument");
    if (!ctx->has_psi[s] || ctx->in_k[s]) return fail(ctx, MSM_E_STATE, "msm_sample_perturbation: psi must be freshly set");
    if (scheme == MSM_SCHEME_NONE) return MSM_OK;
    if (scheme != MSM_SCHEME_WIGNER && scheme != MSM_SCHEME_HUSIMI && scheme != MSM_SCHEME_POISSON)
        return fail(ctx, MSM_E_ARG, "unknown sampling scheme");
    CU(cudaSetDevice(ctx->cfg.device));
    if (scheme == MSM_SCHEME_POISSON) {   // ics.rs:495-558; statistically the reference's sampler (its rng is unseeded, :497)
        if (!(n_tot > 0.0)) return fail(ctx, MSM_E_ARG, "msm_sample_perturbation: n_tot must be positive");
        invalidate_stream(ctx, s);
        if (int rc = wait_upload(ctx, s)) return rc;
        k_sample_poisson<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->X + (size_t)s * ctx->C, ctx->C, seed,
                                                                pow(ctx->cfg.dx, (double)ctx->dims), n_tot, ctx->n, ctx->lb);
        ctx->launches++;
        CU(cudaGetLastError());
        return MSM_OK;
    }
    const double sqrt_dv = sqrt(pow(ctx->cfg.dx, (double)ctx->dims));
    const double div = sqrt(n_tot) * (scheme == MSM_SCHEME_WIGNER ? 2.0 : sqrt(2.0));   // ics.rs:581 / :625
    invalidate_stream(ctx, s);
    if (int rc = wait_upload(ctx, s)) return rc;
    k_sample_gauss<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->X + (size_t)s * ctx->C, ctx->C, seed, sqrt_dv, div, ctx->n, ctx->lb);
    ctx->launches++;
    CU(cudaGetLastError());
    return MSM_OK;
}

// ---- ensemble statistics (row f-3) ---------------------------------------------------------------------------
int msm_ensemble_accumulate(msm_ctx* ctx, const int32_t* active) {
    if (!ctx) return MSM_E_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    std::vector<int> ids = active_list(ctx, active);
    if (ids.empty()) return fail(ctx, MSM_E_ARG, "msm_ensemble_accumulate: no stream selected");
    int rc = ensure_kspace(ctx, ids);
    if (rc) return rc;
    if (!ctx->ens_psi) {
        const size_t cb = sizeof(double2) * (size_t)ctx->C;
        cudaError_t e = cudaMalloc(&ctx->ens_psi, cb);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->ens_psik, cb);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->ens_psi2, cb / 2);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->ens_psik2, cb / 2);
        if (e != cudaSuccess) return fail(ctx, MSM_E_NOMEM, std::string("ensemble grids: ") + cudaGetErrorString(e));
        ctx->bytes += 3 * cb;
    }
    const double nd = pow((double)ctx->n, (double)ctx->dims);
    for (size_t i = 0; i < ids.size(); i += ctx->chunk) {
        const int ns = (int)std::min<size_t>(ctx->chunk, ids.size() - i);
        StreamList byid{ns, {0}}, local{ns, {0}};
        for (int j = 0; j < ns; ++j) {
            byid.id[j] = ids[i + j];
            local.id[j] = j;
        }
        // psi_k as the synthesizer defines it: un-normalised DFT of psi = n^(d/2) * (resident unitary psi_k)
        {
            ProfScope ps(ctx, "ensemble_sums", 16.0 * ctx->C * ns);
            k_stream_sums<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->X, ctx->C, byid, ctx->ens_psik, ctx->ens_psik2, ctx->C,
                                                                   sqrt(nd), nd, i > 0);
        }
        ctx->launches++;
        CU(cudaGetLastError());
        // psi = unitary inverse transform of the resident psi_k, into the scratch slots
        XformOps o;
        o.sop_each = o.sop_last = S_SCALE;
        o.scale = 1.0 / sqrt((double)ctx->n);
        rc = run_transform(ctx, true, &ids[i], ns, ctx->X, 1, ctx->Tscr, 0, o);
        if (rc) return rc;
        {
            ProfScope ps(ctx, "ensemble_sums", 16.0 * ctx->C * ns);
            k_stream_sums<<<grid_for(ctx->C), 256, 0, ctx->st>>>(ctx->Tscr, ctx->C, local, ctx->ens_psi, ctx->ens_psi2, ctx->C, 1.0,
                                                                   1.0, i > 0);
        }
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return MSM_OK;
}

int msm_allreduce_max(msm_ctx* ctx, double* value) {
    if (!ctx || !value) return fail(ctx, MSM_E_ARG, "msm_allreduce_max: bad argument");
    if (ctx->cfg.nranks <= 1 || !ctx->comm) return MSM_OK;
    CU(cudaSetDevice(ctx->cfg.device));
    const int NCCL_MAX = 2;   // ncclMax
    ctx->h_scal[0] = *value;  // mapped pinned memory: the device reads it directly
    k_publish<<<1, 32, 0, ctx->st>>>(ctx->h_scal, ctx->scratch_small, 1);
    CU(cudaGetLastError());
    const int rc = g_nccl.AllReduce(ctx->scratch_small, ctx->scratch_small, 1, NCCL_DOUBLE, NCCL_MAX, ctx->comm, ctx->st);
    if (rc != 0)
        return fail(ctx, MSM_E_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    k_publish<<<1, 32, 0, ctx->st>>>(ctx->scratch_small, ctx->h_scal, 1);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->st));
    *value = ctx->h_scal[0];
    return MSM_OK;
}

int msm_ensemble_allreduce(msm_ctx* ctx) {
    if (!ctx) return MSM_E_ARG;
    if (!ctx->ens_psi) return fail(ctx, MSM_E_STATE, "msm_ensemble_allreduce: call msm_ensemble_accumulate first");
    if (ctx->cfg.nranks <= 1) return MSM_OK;
    if (!ctx->comm) return fail(ctx, MSM_E_STATE, "msm_ensemble_allreduce: the context was created without nccl_unique_id");
    CU(cudaSetDevice(ctx->cfg.device));
    // the four accumulators of synthesizer/src/lib.rs:217-240, summed over the ranks in place (complex grids as 2 C doubles)
    struct { void* p; size_t n; } grids[4] = {{ctx->ens_psi, 2 * (size_t)ctx->C}, {ctx->ens_psik, 2 * (size_t)ctx->C},
                                              {ctx->ens_psi2, (size_t)ctx->C}, {ctx->ens_psik2, (size_t)ctx->C}};
    for (auto& g : grids) {
        int rc;
        {
            ProfScope ps(ctx, "nccl_allreduce_ensemble", 0.0);
            rc = g_nccl.AllReduce(g.p, g.p, g.n, NCCL_DOUBLE, NCCL_SUM, ctx->comm, ctx->st);
        }
        if (rc != 0)
            return fail(ctx, MSM_E_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    }
    return MSM_OK;
}

int msm_ensemble_get(msm_ctx* ctx, int32_t field, double* re, double* im) {
    if (!ctx || field < 0 || field > 3) return fail(ctx, MSM_E_ARG, "msm_ensemble_get: bad argument");
    if (!ctx->ens_psi) return fail(ctx, MSM_E_STATE, "msm_ensemble_get: call msm_ensemble_accumulate first");
    CU(cudaSetDevice(ctx->cfg.device));
    double* planes = reinterpret_cast<double*>(ctx->P);   // 2*C doubles of staging
    if (field == 0 || field == 2) {
        k_deinterleave<<<grid_for(ctx->C), 256, 0, ctx->st>>>(field == 0 ? ctx->ens_psi : ctx->ens_psik, planes, planes + ctx->C,
                                                              ctx->C, ctx->n, ctx->lb);
    } else {
        k_real_to_planes<<<grid_for(ctx->C), 256, 0, ctx->st>>>(field == 1 ? ctx->ens_psi2 : ctx->ens_psik2, planes, ctx->C,
                                                                ctx->n, ctx->lb);
        CU(cudaMemsetAsync(planes + ctx->C, 0, sizeof(double) * (size_t)ctx->C, ctx->st));
    }
    ctx->launches++;
    CU(cudaGetLastError());
    if (re) CU(cudaMemcpyAsync(re, planes, sizeof(double) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st));
    if (im) CU(cudaMemcpyAsync(im, planes + ctx->C, sizeof(double) * (size_t)ctx->C, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MSM_OK;
}

// ---- profiling ---------------------------------------------------------------------------------------------
int msm_profile_enable(msm_ctx* ctx, int32_t on) {
    if (!ctx) return MSM_E_ARG;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->st);
    prof_drain(ctx);
    ctx->prof = on != 0;
    if (on) for (auto& r : ctx->prof_rec) r = msm_profile_record{nullptr, 0, 0.0, 0.0};
    return MSM_OK;
}

int msm_profile_read(msm_ctx* ctx, msm_profile_record* out, int32_t cap, int32_t* n_out) {
    if (!ctx || !n_out) return MSM_E_ARG;
    cudaSetDevice(ctx->cfg.device);
    prof_drain(ctx);
    int k = 0;
    for (size_t i = 0; i < ctx->prof_rec.size(); ++i) {
        if (ctx->prof_rec[i].launches == 0) continue;
        if (out && k < cap) {
            out[k] = ctx->prof_rec[i];
            out[k].name = ctx->prof_names[i].c_str();
        }
        ++k;
    }
    *n_out = k;
    return MSM_OK;
}

int msm_timer_start(msm_ctx* ctx) {
    if (!ctx) return MSM_E_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    if (!ctx->tm_a) {
        CU(cudaEventCreate(&ctx->tm_a));
        CU(cudaEventCreate(&ctx->tm_b));
    }
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaEventRecord(ctx->tm_a, ctx->st));
    return MSM_OK;
}

int msm_timer_stop(msm_ctx* ctx, double* elapsed_ms) {
    if (!ctx || !elapsed_ms || !ctx->tm_a) return fail(ctx, MSM_E_ARG, "msm_timer_stop: bad argument");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaEventRecord(ctx->tm_b, ctx->st));
    CU(cudaEventSynchronize(ctx->tm_b));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->tm_a, ctx->tm_b));
    *elapsed_ms = ms;
    return MSM_OK;
}

int msm_launch_count(const msm_ctx* ctx, uint64_t* launches) {
    if (!ctx || !launches) return MSM_E_ARG;
    *launches = ctx->launches;
    return MSM_OK;
}

}  // extern "C"
