// imma_peak.cu -- measured dense INT8 tensor-core peak of this chip for the roofline of the mask-overlap kernel
// (MEASURED_PEAKS.json only has bf16; round 1 ASSUMED int8 = 2 x bf16).  Every CTA issues the overlap kernel's MMA --
// tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32, S32 accumulators in tensor memory -- back to back from
// one elected thread, with NO operand traffic at all: A sits in tensor memory (mode ts, as in mask_overlap_tc.cu) or
// in shared memory (mode ss), B in shared memory (K-major, SWIZZLE_128B), both never refilled.  What comes out is the
// issue-limited tensor-pipe rate, the upper bound for any kernel built from this instruction.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o imma_peak imma_peak.cu && ./imma_peak [json-out]
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

constexpr int TM = 128, TN = 256, KB = 128;          // one "K-block" = 4 MMAs of K = 32, like the overlap kernel
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((TN >> 3) << 17) | ((TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    return leader != 0;
}

template <bool A_TMEM>
__global__ void __launch_bounds__(128, 1) imma_kernel(int kblocks, int *sink) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *b_tile = smem;                       // [256 rows][128 B]  one K-block of B
    unsigned char *a_tile = smem + TN * KB;             // [128 rows][128 B]  one K-block of A (mode ss)
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (TN + TM) * KB; i += blockDim.x) smem[i] = (unsigned char)(i & 1 ? 0xFF : 0x00);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (A_TMEM) {       // A operand: 32 TMEM columns (= 128 bytes of K per row) at column 256, every lane quarter
        const uint32_t ta = tmem + 256 + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < 32; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(ta + c), "r"(0xFF00FF00u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1 && elect_one()) {
        const uint64_t bd = smem_desc(smem_u32(b_tile)), ad = smem_desc(smem_u32(a_tile));
        for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
            for (int k = 0; k < KB / 32; ++k) {
                const uint32_t acc = (kb | k) != 0;
                if (A_TMEM)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem),
                                 "r"(tmem + 256 + 8 * k), "l"(bd + 2 * k), "r"(IDESC), "r"(acc), "r"(0u)
                                 : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem),
                                 "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(IDESC), "r"(acc), "r"(0u)
                                 : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
    __syncwarp();
    {   // everyone waits for the last MMA
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&done)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(32 * warp) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x12345678u) sink[0] = (int)v;             // keeps the accumulator alive
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <bool A_TMEM>
static double run(int sms, int kblocks, int *sink) {
    const size_t smem = 1024 + (size_t)(TN + TM) * KB;
    cudaFuncSetAttribute(imma_kernel<A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        imma_kernel<A_TMEM><<<sms, 128, smem>>>(kblocks, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tops = 2.0 * TM * TN * KB * (double)kblocks * sms / (ms * 1e-3) / 1e12;
        if (rep > 0 && tops > best) best = tops;
    }
    return best;
}

int main(int argc, char **argv) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int *sink;
    cudaMalloc(&sink, 4);
    const int kblocks = 1 << 16;                      // 65536 x 4 MMAs per CTA: ~0.3 s per launch at peak
    const double ts = run<true>(sms, kblocks, sink), ss = run<false>(sms, kblocks, sink);
    cudaError_t err = cudaDeviceSynchronize();
    printf("tcgen05.mma.kind::i8 M128 N256 K32, %d SMs, one CTA each, no operand traffic: A in TMEM %.1f TOP/s, A in smem %.1f TOP/s (%s)\n",
           sms, ts, ss, cudaGetErrorString(err));
    if (argc > 1 && err == cudaSuccess) {
        FILE *f = fopen(argv[1], "w");
        fprintf(f, "{\n \"int8_tops\": %.1f,\n \"int8_tops_a_in_tmem\": %.1f,\n \"int8_tops_a_in_smem\": %.1f,\n"
                   " \"how\": \"tools/micro/imma_peak.cu: tcgen05.mma.cta_group::1.kind::i8 M128 N256 K32 issued back to back on %d SMs, "
                   "operands resident (no loads), best of 5 launches of %d K-blocks per CTA, CUDA events\",\n \"sms\": %d\n}\n",
                ts > ss ? ts : ss, ts, ss, sms, kblocks, sms);
        fclose(f);
    }
    return err == cudaSuccess ? 0 : 1;
}
