// tmem_rmw.cu -- micro-benchmark for DESIGN.md section 6 ("tensor memory as a second accumulator store"):
// the RoIAlign backward's x pass is a read-modify-write of (f[2p][x], f[2p+1][x]) pairs with lane = channel,
// 7 independent columns per round.  Here the same loop runs (a) on shared memory (LDS.64 / FFMA2 / STS.64) and
// (b) on tensor memory (tcgen05.ld.32x32b.x2 / FFMA2 / tcgen05.st.32x32b.x2), 16 warps per CTA, one CTA per SM,
// and reports read-modify-written bytes per clock and SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_rmw tmem_rmw.cu && ./tmem_rmw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int WARPS = 16, ROUNDS = 4096, PWN = 7;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>       // 0 = shared memory, 1 = tensor memory, 2 = half the warps each
__global__ void __launch_bounds__(WARPS * 32, 1) rmw_kernel(float *out, long long *cycles, int stride) {
    extern __shared__ __align__(16) float2 tile[];                 // [32 lanes][pitch] float2, pitch odd
    __shared__ uint32_t tmem_slot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PITCH = 513;                                     // float2 per lane (16 row pairs x 32 columns + 1)
    float2 *mine = tile + lane * PITCH;
    for (int i = threadIdx.x; i < 32 * PITCH; i += blockDim.x) tile[i] = make_float2(0.f, 0.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w owns row pair w: 32 columns x 2 rows = 64 TMEM columns in its lane quarter (4 warps per quarter)
    const uint32_t tbase = tmem_slot + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * (warp >> 2));
    {   // zero the TMEM region of this warp
        for (int c = 0; c < 64; c += 2)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tbase + c), "r"(0u), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    const bool use_tmem = MODE == 1 || (MODE == 2 && (warp & 1));
    float2 *row = mine + warp * 32;                                // smem: row pair `warp`, 32 columns
    const float w = 1.0f + 1e-3f * lane;
    __syncthreads();
    const long long t0 = clock64();
    int xo = warp & 7;
    for (int r = 0; r < ROUNDS; ++r) {
        float2 v[PWN];
        if (use_tmem) {
#pragma unroll
            for (int pw = 0; pw < PWN; ++pw) {
                uint32_t a, b;
                const uint32_t col = tbase + 2u * (uint32_t)((xo + pw * stride) & 31);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(col) : "memory");
                v[pw] = make_float2(__uint_as_float(a), __uint_as_float(b));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int pw = 0; pw < PWN; ++pw) {
                v[pw].x = fmaf(w, 0.5f, v[pw].x); v[pw].y = fmaf(w, 0.25f, v[pw].y);
                const uint32_t col = tbase + 2u * (uint32_t)((xo + pw * stride) & 31);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(col), "r"(__float_as_uint(v[pw].x)), "r"(__float_as_uint(v[pw].y)) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int pw = 0; pw < PWN; ++pw) v[pw] = row[(xo + pw * stride) & 31];
#pragma unroll
            for (int pw = 0; pw < PWN; ++pw) {
                v[pw].x = fmaf(w, 0.5f, v[pw].x); v[pw].y = fmaf(w, 0.25f, v[pw].y);
                row[(xo + pw * stride) & 31] = v[pw];
            }
            asm volatile("" ::: "memory");
        }
        xo = (xo + 1) & 31;
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    // keep results alive
    float acc = 0.f;
    if (use_tmem) {
        uint32_t a, b;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(tbase) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc = __uint_as_float(a) + __uint_as_float(b);
    } else acc = row[0].x + row[0].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    }
}

template <int MODE>
static void run(const char *name, int stride) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * sms * WARPS * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const size_t smem = sizeof(float2) * 32 * 513;
    cudaFuncSetAttribute(rmw_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) rmw_kernel<MODE><<<sms, WARPS * 32, smem>>>(out, cyc, stride);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double bytes = (double)WARPS * ROUNDS * PWN * 32 * 8;     // read-modify-written bytes per SM
    printf("%-34s stride %d: %9.0f cycles, %6.1f B/clk/SM updated (read + write = %6.1f B/clk), %5.1f cycles per round and warp\n",
           name, stride, avg, bytes / avg, 2 * bytes / avg, avg / ROUNDS);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int stride : {2, 1}) {
        run<0>("shared memory (LDS.64/STS.64)", stride);
        run<1>("tensor memory (tcgen05.ld/st.x2)", stride);
        run<2>("half the warps each", stride);
    }
    return 0;
}
