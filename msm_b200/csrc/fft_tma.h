// fft_tma.h -- host interface of the TMA variant of the plain strided 512-point pass (fft_tma.cu; MSM_B200_TMA=1)
#pragma once
#include <cuda_runtime.h>

#include "fft_pass.cuh"

namespace msm {

struct alignas(64) TmaMap {   // holds a CUtensorMap (128 bytes, 64-byte aligned) without pulling cuda.h into core.cu
    unsigned char bytes[128];
};

struct TmaPassParams {
    const double2* twiddle;   // per-stage tables of the 512-point plan (fft_pass.cuh: plan_tw_offset)
    int ns;                   // grids in this launch
    int slot[MAX_CHUNK];      // grid slot (stream id or scratch index) of each
    int axis;                 // 1 = y, 2 = z
    int ntiles, tiles_inner, tiles_per_cta;
    int lb;                   // log2 of the slow-axis block
    int l2_prefetch;          // cp.async.bulk.prefetch.tensor of the items six ahead (MSM_B200_PREFETCH)
};

// tensor map over `slots` grids of n^3 complex128 at `base` (blocked layout), box = one pass tile of `axis`; 0 = ok
int tma_make_map(TmaMap* out, void* base, int n, int lb, int slots, int axis);
int tma_launch_pass(bool inv, const TmaMap* map, const TmaPassParams& p, int num_sms, cudaStream_t st);

}  // namespace msm
