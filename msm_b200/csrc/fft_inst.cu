// fft_inst.cu -- instantiates the pass kernels for ONE transform length (compile with -DMSM_FFT_N=<N>).
#include "fft_pass.cuh"

#include <stdlib.h>

#ifndef MSM_FFT_N
#error "compile with -DMSM_FFT_N=<power of two>"
#endif

namespace msm {

#define MSM_CAT2(a, b) a##b
#define MSM_CAT(a, b) MSM_CAT2(a, b)

namespace {
constexpr int N = MSM_FFT_N;

template <bool INV, int LOP, int SOP, bool XL>
int launch(const PassParams& p, int ntiles, int groups, cudaStream_t st) {
    using PL = Plan<N>;
    static bool configured = false;
    const size_t smem = pass_smem_bytes<N, LOP, SOP, XL>();
    auto kern = fft_pass_kernel<N, INV, LOP, SOP, XL>;
    if (!configured) {
        if (smem > 32 * 1024) {   // static __shared__ (reduction scratch) counts towards the 48 KiB default
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        // L1 / shared memory split: ask for exactly what the CTAs this kernel is compiled for need (tables live in shared
        // memory, L1 only serves spills).  The driver default sometimes leaves room for one CTA less; the maximum
        // carve-out cost 8 % while the twiddle / drift tables were still read through L1 (profiles/README.md).
        // MSM_B200_CARVEOUT (A/B timing): -1 = driver default, 0..100 = percent of the maximum.
        cudaFuncAttributes fa;
        int pct = (int)cudaSharedmemCarveoutDefault;
        if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {
            const size_t need = (smem + fa.sharedSizeBytes + 1024) * (size_t)kernel_minb<N, LOP, SOP, XL>();
            pct = (int)((need * 100 + 227 * 1024 - 1) / (227 * 1024)) + 1;
            if (pct > 100) pct = 100;
        }
        if (const char* co = getenv("MSM_B200_CARVEOUT")) pct = atoi(co);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        configured = true;
    }
    (void)ntiles;
    dim3 grid((p.tile_end - p.tile0 + p.tiles_per_cta - 1) / p.tiles_per_cta, groups, 1);
    kern<<<grid, tile_threads<N, XL, plan_E<N, LOP, SOP, XL>()>(), smem, st>>>(p);
    return (int)cudaPeekAtLastError();
}

}  // namespace

// the (direction, load, store) combinations the step actually uses
int MSM_CAT(launch_pass_, MSM_FFT_N)(bool inv, int lop, int sop, bool xl, const PassParams& p, int ntiles, int groups,
                                     cudaStream_t st) {
#define CASE(I, L, S)                     \
    if (inv == I && lop == L && sop == S) \
        return xl ? launch<I, L, S, true>(p, ntiles, groups, st) : launch<I, L, S, false>(p, ntiles, groups, st);
    // inverse transforms
    CASE(true, L_NONE, S_NONE)
    CASE(true, L_NONE, S_SCALE)
    CASE(true, L_NONE, S_RHO_ONLY)
    CASE(true, L_NONE, S_RHO_KEEP)
    CASE(true, L_NONE, S_MAX)
    CASE(true, L_DRIFT, S_NONE)
    CASE(true, L_DRIFT, S_RHO_KEEP)
    CASE(true, L_NONE, S_RHO_ONLY_FX)
    CASE(true, L_DRIFT, S_RHO_KEEP_FX)
    // forward transforms
    CASE(false, L_NONE, S_NONE)
    CASE(false, L_NONE, S_SCALE)
    CASE(false, L_NONE, S_DRIFT)
    CASE(false, L_NONE, S_DRIFT_ALIAS)
    CASE(false, L_NONE, S_POISSON)
    CASE(false, L_NONE, S_POISSON_INV)
    CASE(false, L_KICK, S_DRIFT)
    CASE(false, L_KICK, S_DRIFT_ALIAS)
    CASE(false, L_KICK_IX, S_DRIFT)
    CASE(false, L_NONE, S_DRIFT_ALIAS_IZ)
#undef CASE
    // real-field Poisson solve of the summed-density mode: R2C / C2R (run as n/2-point passes along x)
#define CASE_X(I, L, S) \
    if (inv == I && lop == L && sop == S) return launch<I, L, S, true>(p, ntiles, groups, st);
    if (xl) {
        CASE_X(false, L_NONE, S_R2C)
        CASE_X(true, L_C2R, S_NONE)
        CASE_X(true, L_C2R, S_MAX)
    }
#undef CASE_X
    return -1;
}

}  // namespace msm
