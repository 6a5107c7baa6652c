// fft_tma.cu -- TMA variant of the PLAIN strided 512-point pass (load none / store none, y and z axes of the 512^3
// blocked device layout).  An experiment kept behind MSM_B200_TMA=1 (profiles/README.md has the A/B numbers): the tile of
// 8 adjacent lines x 512 positions is described by a 5-D tensor map, lands DIRECTLY in the [position][line] exchange
// layout of the Stockham stages through cp.async.bulk.tensor (completion tracked by an mbarrier, issued three items
// ahead), later items are pulled into L2 by cp.async.bulk.prefetch.tensor, and the finished tile leaves shared memory
// through a bulk tensor store -- no per-thread global addressing, no LDG / STG, and no register-sourced stores in front of
// the next tile's prologue (the code-generation hazard of DESIGN.md section 6 cannot occur by construction).
//
// Same arithmetic as fft_pass_kernel<512, INV, L_NONE, S_NONE, false>: run_stages is shared, results are bit-identical.
// Replaces (reference): utils/fft.rs:6-98 forward / inverse, one axis of the transform.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_pass.cuh"
#include "fft_tma.h"

namespace msm {

namespace {

constexpr int TN = 512, TT = 8, TE = 8, TTHREADS = (TN / TE) * TT;   // 512 threads, 8 points each
constexpr uint32_t TILE_BYTES = TN * TT * sizeof(double2);           // 64 KiB

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_addr(src)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// Tile coordinates in the 5-D view (doubles along k, i_lo, j, i_hi, slot) of the blocked layout [slot][i_hi][j][i_lo][k].
//   y pass: tile = i * (n/8) + m  -> fixed i, k0 = 8 m, all j   : two boxes (16, 1, 256, 1, 1) at j = 0 and j = 256
//   z pass: tile = j * (n/8) + m  -> fixed j, k0 = 8 m, all i   : one box  (16, LO, 1, n/LO, 1)
struct TileCoord {
    int c0, c1, c2, c3;
};
__device__ __forceinline__ TileCoord tile_coord(const TmaPassParams& p, int tile) {
    const int o = tile / p.tiles_inner, m = tile % p.tiles_inner;
    TileCoord c;
    c.c0 = 16 * m;
    if (p.axis == 1) {
        c.c1 = o & ((1 << p.lb) - 1);
        c.c2 = 0;
        c.c3 = o >> p.lb;
    } else {
        c.c1 = 0;
        c.c2 = o;
        c.c3 = 0;
    }
    return c;
}

// Pipelined kernel: ONE persistent CTA per SM with two TEAMS of 512 threads and THREE 64 KiB tile buffers.  Item g of the
// CTA's contiguous range of (slot, tile) items is transformed by team g % 2 in buffer g % 3 -- the tile lands there by
// TMA, is exchanged there by the Stockham stages (the landing buffer IS the exchange buffer) and leaves from there by a
// TMA store.  When a team has stored item g, its first thread waits until the store has READ the buffer and issues the
// load of item g + 3 into it, which the other team will pick up: loads run 1.5 tile-times ahead of their consumer, and
// cp.async.bulk.prefetch.tensor pulls items g + 6 into L2.  192 KiB of tiles + 8 KiB of twiddles per SM.
constexpr int TEAMS = 2, NBUF = 3, PIPE_THREADS = TEAMS * TTHREADS;

template <bool INV>
__global__ void __launch_bounds__(PIPE_THREADS, 1) fft_tma_pass_kernel(const __grid_constant__ CUtensorMap map, const TmaPassParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2* bufs = reinterpret_cast<double2*>(smem_raw);                 // [NBUF][position][line]
    double2* tws = bufs + NBUF * TN * TT;
    uint64_t* full = reinterpret_cast<uint64_t*>(tws + TN);               // one mbarrier per buffer
    volatile int* issued = reinterpret_cast<volatile int*>(full + NBUF);   // loads issued into each buffer so far
    const int team = threadIdx.x / TTHREADS, tid = threadIdx.x % TTHREADS, l = tid % TT, t = tid / TT;
    for (int i = threadIdx.x; i < TN; i += PIPE_THREADS) tws[i] = p.twiddle[i];
    if (threadIdx.x == 0)
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&full[b], 1);
            issued[b] = 0;
        }
    __syncthreads();
    // This CTA's items of the flattened (slot, tile) sequence: chunks of CH consecutive tiles dealt round-robin to the
    // CTAs, so that at any moment the SMs work on neighbouring tiles (adjacent 128-byte segments of the same DRAM pages),
    // like the grid order of the register-staged kernel.  q = 0, 1, 2, ... indexes the CTA's own sequence.
    constexpr int CH = 4;
    const long long total = (long long)p.ns * p.ntiles;
    const long long nchunks = (total + CH - 1) / CH;
    const long long mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;   // chunks of this CTA
    const long long g0 = 0, g1 = mine * CH;                                       // range of q
    auto item_of = [&](long long q) { return ((q / CH) * gridDim.x + blockIdx.x) * CH + q % CH; };
    auto coords = [&](long long q, int half, int (&c)[5]) {
        const long long g = item_of(q);   // (the launcher requires total % CH == 0: no ragged chunk)
        const int tile = (int)(g % p.ntiles);
        const TileCoord tc = tile_coord(p, tile);
        c[0] = tc.c0;
        c[1] = tc.c1;
        c[2] = p.axis == 1 ? 256 * half : tc.c2;
        c[3] = tc.c3;
        c[4] = p.slot[g / p.ntiles];
    };
    const int nbox = p.axis == 1 ? 2 : 1;   // y tiles are two boxes of 256 positions (a box dimension is at most 256)
    auto issue_load = [&](long long g) {    // one thread
        double2* buf = bufs + (size_t)((g - g0) % NBUF) * TN * TT;
        uint64_t* bar = &full[(g - g0) % NBUF];
        mbar_expect_tx(bar, TILE_BYTES);
        for (int h = 0; h < nbox; ++h) {
            int c[5];
            coords(g, h, c);
            tma_load_5d(buf + h * 256 * TT, &map, bar, c[0], c[1], c[2], c[3], c[4]);
        }
        issued[(g - g0) % NBUF] = (int)((g - g0) / NBUF) + 1;
    };
    auto prefetch = [&](long long g) {
        for (int h = 0; h < nbox; ++h) {
            int c[5];
            coords(g, h, c);
            tma_prefetch_5d(&map, c[0], c[1], c[2], c[3], c[4]);
        }
    };
    if (threadIdx.x == 0) {
        for (long long g = g0; g < g1 && g < g0 + NBUF; ++g) issue_load(g);
        for (long long g = g0 + NBUF; p.l2_prefetch && g < g1 && g < g0 + 2 * NBUF; ++g) prefetch(g);
    }
    for (long long g = g0 + team; g < g1; g += TEAMS) {
        const int b = (int)((g - g0) % NBUF);
        double2* exch = bufs + (size_t)b * TN * TT;
        // Use k of a buffer completes phase k of its mbarrier.  A parity wait cannot tell phase k from phase k - 2 (on a
        // barrier still in phase k - 1 it would pass at once), so the team first makes sure that load k has been ISSUED --
        // which the other team does only after it has consumed use k - 1 -- and then waits for the bytes.
        const int use = (int)((g - g0) / NBUF);
        while (issued[b] <= use) {}
        mbar_wait(&full[b], (uint32_t)use & 1);
        // stage-0 inputs: slot n holds element n * 64 + t of line l
        double2 v[TE];
#pragma unroll
        for (int n = 0; n < TE; ++n) v[n] = exch[(n * (TN / TE) + t) * TT + l];
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(TTHREADS) : "memory");   // inputs taken before the first scatter
        run_stages<TN, INV, false, 0, TE, TTHREADS>(v, exch, t, l, tws, 1 + team);
        // last-stage outputs: slot k holds element t + 64 k; back into the tile layout, then one bulk store
#pragma unroll
        for (int k = 0; k < TE; ++k) exch[(t + (TN / TE) * k) * TT + l] = v[k];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(TTHREADS) : "memory");
        if (tid == 0) {
            for (int h = 0; h < nbox; ++h) {
                int c[5];
                coords(g, h, c);
                tma_store_5d(&map, exch + h * 256 * TT, c[0], c[1], c[2], c[3], c[4]);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (g + NBUF < g1) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile has left the buffer
                issue_load(g + NBUF);
                if (p.l2_prefetch && g + 2 * NBUF < g1) prefetch(g + 2 * NBUF);
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn_t encode_fn() {
    static encode_fn_t fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_fn_t>(p);
    }
    return fn;
}

}  // namespace

// 5-D tensor map over `slots` grids of n^3 complex128 in the blocked layout [slot][i_hi][j][i_lo][k] (fp64 elements).
int tma_make_map(TmaMap* out, void* base, int n, int lb, int slots, int axis) {
    encode_fn_t enc = encode_fn();
    if (!enc) return -1;
    const cuuint64_t LO = 1ull << lb;
    const cuuint64_t dims[5] = {2ull * n, LO, (cuuint64_t)n, (cuuint64_t)n / LO, (cuuint64_t)slots};
    const cuuint64_t row = 16ull * n;
    const cuuint64_t strides[4] = {row, row * LO, row * LO * n, row * (cuuint64_t)n * n};   // bytes, dims 1..4
    cuuint32_t box[5] = {16, 1, 256, 1, 1};
    if (axis == 2) {
        box[1] = (cuuint32_t)LO;
        box[2] = 1;
        box[3] = (cuuint32_t)(n / LO);
    }
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    static_assert(sizeof(TmaMap) == sizeof(CUtensorMap), "TmaMap must hold a CUtensorMap");
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

int tma_launch_pass(bool inv, const TmaMap* map, const TmaPassParams& p, int num_sms, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = (size_t)NBUF * TILE_BYTES + TN * sizeof(double2) + 64;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fft_tma_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fft_tma_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const long long total = (long long)p.ns * p.ntiles;
    if (total % 4) return -1;   // chunks of 4 tiles (CH in the kernel)
    dim3 grid((unsigned)(total / 4 < num_sms ? total / 4 : num_sms), 1, 1);   // persistent: one CTA per SM
    const CUtensorMap& m = *reinterpret_cast<const CUtensorMap*>(map);
    if (inv) fft_tma_pass_kernel<true><<<grid, PIPE_THREADS, smem, st>>>(m, p);
    else fft_tma_pass_kernel<false><<<grid, PIPE_THREADS, smem, st>>>(m, p);
    return (int)cudaPeekAtLastError();
}

}  // namespace msm
