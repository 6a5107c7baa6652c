// fft_pipe.cuh -- persistent, software-pipelined version of the axis-pass kernel (fft_pass.cuh) for N >= 128.
//
// Why: the ncu profile of the one-tile-per-CTA kernel (profiles/r01a_*) shows the pass is latency bound, not
// bandwidth bound: 64 KiB tiles held in registers + 64 regs/thread allow only 2 CTAs per SM, so the global-load
// latency of every tile is exposed (24 % of all stall samples sit on the first butterfly instruction) and barrier
// stalls are not covered by other work.
//
// How: one CTA per SM, made of G independent "compute groups" (named barriers).  Each group walks over its own
// sequence of work items (tile x stream) and owns two shared-memory buffers of one tile (T2 = 4 lines, 32 KiB at
// N = 512).  While a group transforms item i in buffer i&1 (the Stockham exchanges happen in place in that buffer),
// the tile of item i+1 streams into the other buffer with cp.async (LDGSTS, 16 B per thread, L1 bypass): registers
// hold a tile only while it is being computed, so G x 32 KiB of loads are always in flight per SM and the other
// groups fill every barrier bubble.  Every thread copies exactly the elements it later gathers, so the copy needs
// no barrier of its own -- only cp.async.wait_group.
//
//   prologue: prefetch item 0
//   loop:     wait_group 0 -> gather stage-0 inputs (smem -> registers) -> group barrier
//             -> prefetch item i+1 -> load operator -> stages (exchanges in place) -> store operator + st.global
#pragma once
#include "fft_pass.cuh"

namespace msm {

template <int N> struct PipePlan {
    static constexpr int T2 = 4;                                  // lines per tile
    static constexpr int GT = (N / Plan<N>::E) * T2;              // threads per compute group
    static constexpr int TILE_BYTES = N * T2 * (int)sizeof(double2);
    // groups per CTA: two tile buffers each (+ one phi buffer for the kick), <= 200 KiB, <= 768 threads
    static constexpr int groups(bool kick) {
        int per = TILE_BYTES * (kick ? 3 : 2);
        int g = (200 * 1024) / per;
        int gth = 768 / GT;
        g = g < gth ? g : gth;
        g = g > 8 ? 8 : g;
        return g < 1 ? 1 : g;
    }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// exchange-buffer index for T2 = 4 lines.  strided axes: [position][line] with the 64-byte half of every 128-byte
// row flipped by position bit 3 (keeps the stride-8 gather of the last stage conflict free);
// contiguous axis: [line][position ^ swizzle] as in fft_pass.cuh.
template <int N, bool XL> __device__ __forceinline__ int pipe_index(int pos, int l) {
    if (XL) return l * N + (pos ^ ((pos >> 3) & 7));
    return (pos * 4 + l) ^ (((pos >> 3) & 1) << 2);
}

template <int N, bool INV, bool XL, int Q>
__device__ __forceinline__ void pipe_stages(double2 (&v)[Plan<N>::E], double2* sm, int t, int l,
                                            const double2* __restrict__ tw, int bar_id) {
    using PL = Plan<N>;
    constexpr int E = PL::E, NT = PL::NT, GT = PipePlan<N>::GT;
    constexpr int R = PL::R[Q];
    constexpr int L = plan_L<N>(Q);
    constexpr int M = N / (L * R);
    constexpr int NB = E / R;
#pragma unroll
    for (int c = 0; c < NB; ++c) Dft<R, INV>::run(&v[c * R]);
    if constexpr (Q + 1 < PL::NS) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const int b = t + NT * c;
            const int kappa = b / M, nu = b % M;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                double2 x = v[c * R + k];
                if (k > 0) {
                    double2 w = __ldg(&tw[plan_tw_offset<N>(Q) + (k - 1) * M + nu]);
                    if (INV) w.y = -w.y;
                    x = cmul(x, w);
                }
                sm[pipe_index<N, XL>((kappa + L * k) * M + nu, l)] = x;
            }
        }
        group_barrier(bar_id, GT);
        constexpr int R2 = PL::R[Q + 1];
        constexpr int L2 = L * R;
        constexpr int M2 = N / (L2 * R2);
        constexpr int NB2 = E / R2;
#pragma unroll
        for (int c = 0; c < NB2; ++c) {
            const int b = t + NT * c;
            const int kappa = b / M2, nu = b % M2;
#pragma unroll
            for (int n = 0; n < R2; ++n) v[c * R2 + n] = sm[pipe_index<N, XL>(kappa * (M2 * R2) + n * M2 + nu, l)];
        }
        group_barrier(bar_id, GT);
        pipe_stages<N, INV, XL, Q + 1>(v, sm, t, l, tw, bar_id);
    }
}

template <int N, int LOP> constexpr size_t pipe_smem_bytes() {
    return (size_t)PipePlan<N>::groups(LOP == L_KICK) * PipePlan<N>::TILE_BYTES * (LOP == L_KICK ? 3 : 2);
}

template <int N, bool INV, int LOP, int SOP, bool XL>
__global__ void __launch_bounds__(PipePlan<N>::groups(LOP == L_KICK) * PipePlan<N>::GT, 1)
    fft_pipe_kernel(const PassParams p) {
    using PL = Plan<N>;
    using PP = PipePlan<N>;
    static_assert(PL::NS >= 2, "the pipelined kernel needs at least one exchange");
    constexpr int E = PL::E, NT = PL::NT, T2 = PP::T2, GT = PP::GT;
    constexpr int G = PP::groups(LOP == L_KICK);
    constexpr int R0 = PL::R[0];
    constexpr int M0 = N / R0;
    constexpr int NB0 = E / R0;
    constexpr int RL = PL::R[PL::NS - 1];
    constexpr int LL = plan_L<N>(PL::NS - 1);
    constexpr int NBL = E / RL;
    constexpr int TILE = N * T2;

    extern __shared__ double2 smem_all[];
    __shared__ double red[G][GT / 32], red2[G][GT / 32];

    const int grp = threadIdx.x / GT;
    const int tid = threadIdx.x % GT;
    const int bar_id = 1 + grp;
    double2* bufs = smem_all + (size_t)grp * TILE * (LOP == L_KICK ? 3 : 2);
    double2* phibuf = bufs + 2 * TILE;   // kick only
    const int l = XL ? tid / NT : tid % T2;
    const int t = XL ? tid % NT : tid / T2;
    const bool lv = l < p.lvalid;

    const int groups_s = (p.ns + p.gsz - 1) / p.gsz;                 // stream groups
    const long long units = (long long)p.ntiles * groups_s;          // (tile, stream group)
    const long long cg = (long long)blockIdx.x * G + grp;
    const long long ncg = (long long)gridDim.x * G;

    // item = (unit, q): stream q of the unit's stream group.  Items of a unit are consecutive (rho / phi pairing).
    auto decode = [&](long long it, int& tile, int& g, int& q, int& li) -> bool {
        const long long u = cg + (it / p.gsz) * ncg;
        if (u >= units) return false;
        tile = (int)(u % p.ntiles);
        g = (int)(u / p.ntiles);
        q = (int)(it % p.gsz);
        li = g * p.gsz + q;
        return true;
    };
    auto tile_base = [&](int tile) -> long long {
        const int o = tile / p.tiles_inner;
        return (long long)(o >> p.olb) * p.outer_stride + (long long)(o & ((1 << p.olb) - 1)) * p.outer_lo +
               (long long)(tile % p.tiles_inner) * p.inner_stride + (long long)l * p.lstride;
    };
    auto along = [&](int e) -> long long {
        return (long long)(e >> p.alb) * p.astride + (long long)(e & ((1 << p.alb) - 1)) * p.astride_lo;
    };
    auto prefetch = [&](long long it, int buf) {
        int tile, g, q, li;
        if (decode(it, tile, g, q, li) && li < p.ns) {
            const int s = p.sid[li];
            const double2* __restrict__ src = p.src + (long long)(p.src_by_sid ? s : li) * p.src_sstride;
            const long long base = tile_base(tile);
            double2* dstb = bufs + buf * TILE;
#pragma unroll
            for (int c = 0; c < NB0; ++c) {
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int e = n * M0 + t + NT * c;
                    cp_async16(&dstb[pipe_index<N, XL>(e, l)], lv ? (const void*)(src + base + along(e))
                                                                  : (const void*)p.src, lv);
                }
            }
            if constexpr (LOP == L_KICK) {
                if (q == 0) {   // phi_a + i phi_b of this tile, once per unit
                    const double2* __restrict__ pb = p.pbuf + (p.p_summed ? 0 : (long long)g * p.p_gstride);
#pragma unroll
                    for (int c = 0; c < NB0; ++c) {
#pragma unroll
                        for (int n = 0; n < R0; ++n) {
                            const int e = n * M0 + t + NT * c;
                            cp_async16(&phibuf[pipe_index<N, XL>(e, l)],
                                       lv ? (const void*)(pb + base + along(e)) : (const void*)p.pbuf, lv);
                        }
                    }
                }
            }
        }
        cp_async_commit();
    };

    double keep[(SOP == S_RHO_KEEP || SOP == S_RHO_ONLY) ? E : 1];   // the partner stream's rho (pair buffers)

    prefetch(0, 0);
    for (long long it = 0;; ++it) {
        int tile, g, q, li;
        if (!decode(it, tile, g, q, li)) break;
        const bool have = li < p.ns;             // an odd stream count leaves the last group half empty
        const int buf = (int)(it & 1);
        double2* sm = bufs + buf * TILE;
        const int s = have ? p.sid[li] : 0;
        const bool last_of_group = (q + 1 == p.gsz) || (li + 1 >= p.ns);
        const long long base = tile_base(tile);
        double2* __restrict__ dst = p.dst + (long long)(p.dst_by_sid ? s : li) * p.dst_sstride;
        double2* __restrict__ pb = p.pbuf + (p.p_summed ? 0 : (long long)g * p.p_gstride);

        // coordinates of this line along the two non-pass axes (k^2 consumers only)
        double kline = 0.0;
        int c0 = 0, c1 = 0, c2 = 0;
        if constexpr (SOP == S_DRIFT_ALIAS || SOP == S_POISSON) {
            const int n = p.n;
            if (p.axis == 0) {
                const int line = tile * T2 + l;
                c1 = (line >> p.row_lb) % n;
                c2 = ((line >> p.row_lb) / n << p.row_lb) + (line & ((1 << p.row_lb) - 1));
            } else if (p.axis == 1) {
                c2 = tile / p.tiles_inner;
                c0 = (tile % p.tiles_inner) * T2 + l;
            } else {
                const int line = tile * T2 + l;
                c1 = line / n;
                c0 = line % n;
            }
            if (!lv) c0 = c1 = c2 = 0;
            if (p.axis == 2) kline = p.ksq[c0] + p.ksq[c1];   // spec_grid order: ((k0^2 + k1^2) + k2^2) (2 pi)^2
        }
        auto k2_of = [&](int e) -> double {
            double sum;
            if (p.axis == 0) sum = (p.ksq[e] + p.ksq[c1]) + p.ksq[c2];
            else if (p.axis == 1) sum = (p.ksq[c0] + p.ksq[e]) + p.ksq[c2];
            else sum = kline + p.ksq[e];
            return sum * p.four_pi2;
        };

        cp_async_wait_all();
        double2 v[E];
        double ph[LOP == L_KICK ? E : 1];
#pragma unroll
        for (int c = 0; c < NB0; ++c) {
#pragma unroll
            for (int n = 0; n < R0; ++n) {
                const int e = n * M0 + t + NT * c;
                v[c * R0 + n] = sm[pipe_index<N, XL>(e, l)];
                if constexpr (LOP == L_KICK) {
                    const double2 pp = phibuf[pipe_index<N, XL>(e, l)];
                    ph[c * R0 + n] = (p.p_summed || q == 0) ? pp.x : pp.y;
                }
            }
        }
        group_barrier(bar_id, GT);   // everyone holds its inputs: the buffer is now exchange space
        prefetch(it + 1, buf ^ 1);
        if (!have) continue;

        if constexpr (LOP == L_DRIFT) {
#pragma unroll
            for (int c = 0; c < NB0; ++c) {
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int e = n * M0 + t + NT * c;
                    v[c * R0 + n] = cmul(v[c * R0 + n], __ldg(&p.dtab[(long long)s * N + e]));
                }
            }
        }
        if constexpr (LOP == L_KICK) {
            // psi *= exp(-i kappa phi)    (simulation_object.rs:535-545)
            const double kap = -p.kick[li];
#pragma unroll
            for (int j = 0; j < E; ++j) {
                double sn, cs;
                kick_sincos(kap * ph[j], &sn, &cs);
                v[j] = cmul(v[j], make_double2(cs, sn));
            }
        }

        pipe_stages<N, INV, XL, 0>(v, sm, t, l, p.twiddle, bar_id);

        double acc = 0.0, acc2 = 0.0;
#pragma unroll
        for (int c = 0; c < NBL; ++c) {
#pragma unroll
            for (int k = 0; k < RL; ++k) {
                const int e = t + NT * c + LL * k;
                const long long off = base + along(e);
                double2 x = v[c * RL + k];
                if constexpr (SOP == S_SCALE) {
                    x.x *= p.scale;
                    x.y *= p.scale;
                }
                if constexpr (SOP == S_DRIFT || SOP == S_DRIFT_ALIAS) x = cmul(x, __ldg(&p.dtab[(long long)s * N + e]));
                if constexpr (SOP == S_DRIFT_ALIAS) {
                    if (lv && k2_of(e) > p.alias_k2_thresh) acc += x.x * x.x + x.y * x.y;   // check_alias :1259-1280
                }
                if constexpr (SOP == S_POISSON) {
                    const double k2 = k2_of(e);                                            // :1076-1102, DC -> 0
                    const double m = (k2 == 0.0) ? 0.0 : p.poisson_coef * fast_rcp(k2);
                    x.x *= m;
                    x.y *= m;
                }
                if constexpr (SOP == S_MAX) {
                    acc = fmax(acc, fabs(x.x));
                    acc2 = fmax(acc2, fabs(x.y));
                }
                if constexpr (SOP == S_RHO_KEEP || SOP == S_RHO_ONLY) {
                    const double rho = p.rho_coef * (x.x * x.x + x.y * x.y);               // :1051-1062
                    const int j = c * RL + k;
                    if (!p.p_summed) {
                        if (!last_of_group) keep[j] = rho;
                        else if (lv) pb[off] = (q == 0) ? make_double2(rho, 0.0) : make_double2(keep[j], rho);
                    } else {
                        const double sum = (q == 0) ? rho : keep[j] + rho;
                        if (!last_of_group) keep[j] = sum;
                        else if (lv) pb[off] = make_double2(p.rho_accumulate ? pb[off].x + sum : sum, 0.0);
                    }
                }
                if constexpr (SOP != S_RHO_ONLY && SOP != S_MAX) {
                    if (lv) dst[off] = x;
                }
            }
        }
        if constexpr (SOP == S_DRIFT_ALIAS || SOP == S_MAX) {
            if (SOP == S_DRIFT_ALIAS) acc = warp_sum(acc);
            else {
                acc = warp_max(acc);
                acc2 = warp_max(acc2);
            }
            if ((tid & 31) == 0) {
                red[grp][tid >> 5] = acc;
                red2[grp][tid >> 5] = acc2;
            }
            group_barrier(bar_id, GT);
            if (tid == 0) {
                if (SOP == S_DRIFT_ALIAS) {
                    double tot = 0.0;
                    for (int w = 0; w < GT / 32; ++w) tot += red[grp][w];
                    p.alias_partial[(long long)s * p.ntiles + tile] = tot;
                } else {
                    double m1 = 0.0, m2 = 0.0;
                    for (int w = 0; w < GT / 32; ++w) {
                        m1 = fmax(m1, red[grp][w]);
                        m2 = fmax(m2, red2[grp][w]);
                    }
                    if (m1 > 0.0) atomicMax(&p.maxbits[2 * li], (unsigned long long)__double_as_longlong(m1));
                    if (m2 > 0.0) atomicMax(&p.maxbits[2 * li + 1], (unsigned long long)__double_as_longlong(m2));
                }
            }
            group_barrier(bar_id, GT);
        }
    }
    cp_async_wait_all();
}

}  // namespace msm
