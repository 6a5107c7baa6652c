// strided_copy_probe.cu -- what HBM bandwidth does a PURE COPY reach with the access pattern of the strided FFT passes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_copy_probe strided_copy_probe.cu && ./strided_copy_probe
// One CTA of 512 threads moves a tile of 8 adjacent complex128 (one 128-byte row segment) x 512 positions, exactly like
// fft_pass_kernel on the y / z axes of the blocked 512^3 layout [i_hi][j][i_lo][k] (LO = 16): thread (t, l) touches the
// elements e = t + 64 m (m < 8) of line l, 4 consecutive tiles per CTA, 2 CTAs per SM, in place over S = 8 grids (16 GiB).
// No butterflies, no shared memory: load 8 x 16 B, store them back.  The contiguous case is the usual streaming copy.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int N = 512, LO = 16, T = 8;

__global__ void __launch_bounds__(512, 2) copy_tiles(double2* __restrict__ a, int axis, int tiles_per_cta, long long sstride) {
    const int tid = threadIdx.x, l = tid % T, t = tid / T;
    double2* base = a + (long long)blockIdx.y * sstride;
    const int tiles_inner = N / T;
    for (int ti = 0; ti < tiles_per_cta; ++ti) {
        const int tile = blockIdx.x * tiles_per_cta + ti;
        const int o = tile / tiles_inner, m = tile % tiles_inner;
        long long origin, astep;
        if (axis == 1) {   // y: fixed i = o, positions along j (stride N * LO elements)
            origin = (long long)(o / LO) * N * N * LO + (long long)(o % LO) * N + m * T;
            astep = (long long)N * LO;
        } else {           // z: fixed j = o, positions along i = (i_hi, i_lo)
            origin = (long long)o * N * LO + m * T;
            astep = 0;
        }
        double2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int e = t + 64 * k;
            const long long off = axis == 1 ? origin + e * astep + l
                                            : origin + (long long)(e / LO) * N * N * LO + (long long)(e % LO) * N + l;
            v[k] = base[off];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int e = t + 64 * k;
            const long long off = axis == 1 ? origin + e * astep + l
                                            : origin + (long long)(e / LO) * N * N * LO + (long long)(e % LO) * N + l;
            v[k].x += 1.0;
            base[off] = v[k];
        }
    }
}

__global__ void copy_linear(double2* __restrict__ a, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double2 v = a[i];
        v.x += 1.0;
        a[i] = v;
    }
}

int main() {
    const int S = 8;
    const long long C = (long long)N * N * N;
    double2* a;
    if (cudaMalloc(&a, sizeof(double2) * C * S) != cudaSuccess) return 1;
    cudaMemset(a, 0, sizeof(double2) * C * S);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double gb = 2.0 * 16.0 * C * S / 1e9;
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) copy_linear<<<148 * 16, 512>>>(a, C * S);
            else copy_tiles<<<dim3(N * N / T / 4, S), 512>>>(a, mode, 4, C);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%-28s %7.3f ms  %7.1f GB/s (read + write, %d x 512^3 complex128 in place)\n",
               mode == 0 ? "contiguous copy" : mode == 1 ? "y-pass tiles (128 B rows)" : "z-pass tiles (128 B rows)", best, gb / (best * 1e-3), S);
    }
    if (cudaGetLastError() != cudaSuccess) return 2;
    return 0;
}
