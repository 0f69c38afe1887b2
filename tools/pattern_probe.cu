// Developer tool (not part of the product library): what HBM bandwidth does the ACCESS PATTERN of the score-only RTS
// smoother reach when nothing is computed?  Same structure-of-arrays layout ([component][step][trajectory]), same
// mapping (one thread per trajectory, 128-thread CTAs, reverse time loop), same 70 of the 90 component planes read per
// step (filtered mean 5, lower triangles of the filtered / predictive covariances 15 + 15, predictive mean 5,
// cross-covariance 25, truth 5) and 6 planes written (errors 5, quadratic form 1): 608 B per unit.  The loads are plain
// streaming loads summed into one value, so the instruction stream is ~150 instead of ~1 500 instructions per unit.
//   nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pattern_probe.so tools/pattern_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
constexpr int DX = 5, NP = 90, NR = 70, NW = 6;
// planes of one step: [fi_mean 0..4 | fi_cov 5..29 | pr_mean 30..34 | pr_cov 35..59 | pr_xx 60..84 | x 85..89]
__constant__ int c_planes[NR];

template <int MINB>
__global__ void __launch_bounds__(128, MINB) probe_kernel(const double *__restrict__ in, double *__restrict__ out, long long M, int N) {
    extern __shared__ double pad[];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    const long long cs = (long long)N * M;
    double acc = 0.0;
    for (int k = N - 1; k >= 0; --k) {
        const double *p = in + (long long)k * M + t;
        // the loads of a step go out in batches of NR / MINB planes when the register budget is small (MINB = 4: 128 registers)
        constexpr int NB = MINB > 2 ? 2 : 1, B = NR / NB;
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            double v[B];
#pragma unroll
            for (int j = 0; j < B; ++j) v[j] = __ldcs(p + (long long)c_planes[b * B + j] * cs);
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int j = 0; j + 1 < B; j += 2) { s0 += v[j]; s1 += v[j + 1]; }
            s += s0 + s1 + ((B & 1) ? v[B - 1] : 0.0);
            if (NB > 1) asm volatile("" ::: "memory");   // keep the batches apart
        }
        acc += s;
        double *q = out + (long long)k * M + t;
#pragma unroll
        for (int j = 0; j < NW; ++j) __stcs(q + (long long)j * cs, s + j);
    }
    if (acc == 12345.678 && pad) out[0] = acc;
}
// The same pattern through a DEPTH-stage ring of cp.async copies (lane pairs copy 16 bytes = one plane of both
// trajectories, L2 -> shared memory), DEPTH - 1 steps in flight per thread while one is summed: the pattern's ceiling
// when latency is covered.  One CTA of 128 threads per SM at DEPTH = 3 (3 x 70 KB of shared memory).
template <int DEPTH>
__global__ void __launch_bounds__(128, 1) probe_ring_kernel(const double *__restrict__ in, double *__restrict__ out, long long M, int N) {
    extern __shared__ __align__(16) double ring[];   // [DEPTH][NR][128]
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int par = threadIdx.x & 1;
    const bool pair_in = (t | 1) < M;
    const long long cs = (long long)N * M;
    auto issue = [&](int k, int slot) {
        if (pair_in && k >= 0) {
            const double *p = in + (long long)k * M + (t - par);
            double *d = ring + (size_t)slot * NR * 128 + (threadIdx.x & ~1);
#pragma unroll
            for (int j = 0; j < NR; ++j)
                if ((j & 1) == par) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(d + j * 128);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(p + (long long)c_planes[j] * cs) : "memory");
                }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int i = 0; i < DEPTH - 1; ++i) issue(N - 1 - i, i);
    double acc = 0.0;
    int slot = 0;
    for (int k = N - 1; k >= 0; --k) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 2) : "memory");
        __syncwarp();
        double s0 = 0.0, s1 = 0.0;
        if (t < M) {
            const double *d = ring + (size_t)slot * NR * 128 + threadIdx.x;
#pragma unroll
            for (int j = 0; j + 1 < NR; j += 2) { s0 += d[j * 128]; s1 += d[(j + 1) * 128]; }
        }
        __syncwarp();
        issue(k - (DEPTH - 1), (slot + DEPTH - 1) % DEPTH);
        const double s = s0 + s1;
        acc += s;
        if (t < M) {
            double *q = out + (long long)k * M + t;
#pragma unroll
            for (int j = 0; j < NW; ++j) __stcs(q + (long long)j * cs, s + j);
        }
        slot = (slot + 1) % DEPTH;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (acc == 12345.678) out[0] = acc;
}
}  // namespace

extern "C" int probe_ring_run(const double *in, double *out, long long M, int N, int depth, void *stream) {
    const unsigned grid = (unsigned)((M + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if ((M & 1) || ((uintptr_t)in & 15)) return -2;
    const int bytes = depth * NR * 128 * (int)sizeof(double);
    if (depth == 3) {
        cudaFuncSetAttribute(probe_ring_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        probe_ring_kernel<3><<<grid, 128, bytes, s>>>(in, out, M, N);
    } else if (depth == 2) {
        cudaFuncSetAttribute(probe_ring_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        probe_ring_kernel<2><<<grid, 128, bytes, s>>>(in, out, M, N);
    } else return -3;
    return (int)cudaGetLastError();
}

extern "C" int probe_init() {
    int h[NR], n = 0;
    for (int a = 0; a < DX; ++a) h[n++] = a;                                              // fi_mean
    for (int r = 0; r < DX; ++r) for (int c = 0; c <= r; ++c) h[n++] = 5 + r * DX + c;    // fi_cov, lower triangle
    for (int a = 0; a < DX; ++a) h[n++] = 30 + a;                                         // pr_mean
    for (int r = 0; r < DX; ++r) for (int c = 0; c <= r; ++c) h[n++] = 35 + r * DX + c;   // pr_cov, lower triangle
    for (int c = 0; c < DX * DX; ++c) h[n++] = 60 + c;                                    // pr_xx
    for (int a = 0; a < DX; ++a) h[n++] = 85 + a;                                         // truth
    if (n != NR) return -1;
    return (int)cudaMemcpyToSymbol(c_planes, h, sizeof(h));
}

// ctas_per_sm: 0 = four CTAs per SM (128 registers: the loads of a step go out in batches), 2 = the smoother's occupancy (8 warps per SM,
// enforced with a dynamic shared-memory allocation of 100 KB per CTA)
extern "C" int probe_run(const double *in, double *out, long long M, int N, int ctas_per_sm, void *stream) {
    const unsigned grid = (unsigned)((M + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if (ctas_per_sm == 2) {
        const int bytes = 100 * 1024;
        cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        probe_kernel<2><<<grid, 128, bytes, s>>>(in, out, M, N);
    } else {
        probe_kernel<4><<<grid, 128, 0, s>>>(in, out, M, N);
    }
    return (int)cudaGetLastError();
}
