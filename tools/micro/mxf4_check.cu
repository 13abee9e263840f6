// mxf4_check.cu -- (1) does tcgen05.mma.kind::mxf4 count mask intersections exactly?  (2) its measured dense peak.
//
// The mask-overlap kernel (cim_b200/csrc/mask_overlap_tc.cu) contracts 0/1 operands; as INT8 bytes the expanded B
// operand costs 32 KB of shared-memory stores + 32 KB of tensor-core reads per 128-pixel K-block, which is what
// bounds it (the shared-memory port, not the tensor pipe).  As E2M1 nibbles (kind::mxf4, K = 64 per MMA, block
// scales all 2^0) both halve and the MMA runs at twice the int8 rate.  A set pixel becomes a single-bit nibble --
// 0b0001 = 0.5, 0b0010 = 1.0, 0b0100 = 2.0 -- chosen per pixel position so that A's value times B's value is 1.0:
//   A: px 4j -> 0.5, 4j+1 -> 1.0, 4j+2 -> 2.0, 4j+3 -> 2.0        B: 2.0, 1.0, 0.5, 0.5
// The fp32 accumulator then holds the exact count (< 2^24).
// part 1: one CTA, 128 x 256 tile, random masks of several densities, A expanded into TENSOR MEMORY, B into the
//         K-major SWIZZLE_64B shared-memory layout, compared with AND + POPC on the host.
// part 2: every SM issues M128 N256 K64 MMAs back to back with resident operands (like imma_peak.cu).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mxf4_check mxf4_check.cu && ./mxf4_check [json-out]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int TM = 128, TN = 256, KB = 128;          // K-block = 128 pixels = 64 operand bytes per row = 2 MMAs of K = 64
constexpr int A_COLS = KB / 8;                       // 16 TMEM columns of A per K-block
constexpr int SF_COL = 448;                          // 64 columns of 0x7F (UE8M0 1.0): scale factors of A and B
// block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): a_format = b_format = 1 (MXF4Format::E2M1),
// K-major both, N >> 3 at bit 17, scale format UE8M0 (bit 23), M >> 4 at bit 24, sf ids 0, K = 64
constexpr uint32_t IDESC = (1u << 7) | (1u << 10) | ((TN >> 3) << 17) | (1u << 23) | ((TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major SWIZZLE_64B: 8-row x 64 B atoms, 512 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    return leader != 0;
}
__device__ __forceinline__ void mma_mxf4_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t sfa, uint32_t sfb, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}" ::"r"(d),
                 "r"(a), "l"(bd), "r"(IDESC), "r"(acc), "r"(sfa), "r"(sfb)
                 : "memory");
}
__device__ __forceinline__ void expand_a(uint32_t w, uint32_t (&o)[4]) {
    o[0] = w & 0x11111111u; o[1] = w & 0x22222222u; o[2] = w & 0x44444444u; o[3] = (w >> 1) & 0x44444444u;
}
__device__ __forceinline__ void expand_b(uint32_t w, uint32_t (&o)[4]) {
    o[0] = (w << 2) & 0x44444444u; o[1] = w & 0x22222222u; o[2] = (w >> 2) & 0x11111111u; o[3] = (w >> 3) & 0x11111111u;
}
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- part 1: correctness.  a_bits [128][nkb * 4] words, b_bits [256][nkb * 4] words, out [128][256] float
__global__ void __launch_bounds__(128, 1) check_kernel(const uint32_t *a_bits, const uint32_t *b_bits, int nkb, float *out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *b_tile = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // [256][64 B]
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)(32 * warp) << 16;
    for (int c = 0; c < 64; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_base + SF_COL + c), "r"(0x7F7F7F7Fu) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int kb = 0; kb < nkb; ++kb) {
        // A: row = tid -> 16 TMEM columns at column 256
        for (int q = 0; q < 4; ++q) {
            uint32_t o[4];
            expand_a(a_bits[(size_t)tid * nkb * 4 + kb * 4 + q], o);
            for (int g = 0; g < 4; ++g)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_base + 256 + q * 4 + g), "r"(o[g]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // B: rows tid and tid + 128 -> 4 swizzled 16-byte chunks each
        for (int h = 0; h < 2; ++h) {
            const int r = tid + 128 * h;
            for (int q = 0; q < 4; ++q) {
                uint32_t o[4];
                expand_b(b_bits[(size_t)r * nkb * 4 + kb * 4 + q], o);
                *reinterpret_cast<uint4 *>(b_tile + (r >> 3) * 512 + (r & 7) * 64 + ((q ^ ((r >> 1) & 3)) << 4)) =
                    make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 1 && elect_one()) {
            const uint64_t bd = smem_desc(smem_u32(b_tile));
            for (int k = 0; k < 2; ++k)
                mma_mxf4_ts(tmem, tmem + 256 + 8 * k, bd + 2 * k, tmem + SF_COL, tmem + SF_COL, (kb | k) != 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
        }
        __syncwarp();
        wait_bar(&done, kb & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    for (int c = 0; c < TN; ++c) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + lane_base + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[(size_t)tid * TN + c] = __uint_as_float(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---- part 2: peak
__global__ void __launch_bounds__(128, 1) peak_kernel(int kblocks, int *sink) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *b_tile = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < TN * 64; i += blockDim.x) b_tile[i] = (unsigned char)(i & 1 ? 0x22 : 0x00);
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
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)(32 * warp) << 16;
    for (int c = 0; c < 64; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_base + SF_COL + c), "r"(0x7F7F7F7Fu) : "memory");
    for (int c = 0; c < A_COLS; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_base + 256 + c), "r"(0x22002200u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1 && elect_one()) {
        const uint64_t bd = smem_desc(smem_u32(b_tile));
        for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
                mma_mxf4_ts(tmem, tmem + 256 + 8 * k, bd + 2 * k, tmem + SF_COL, tmem + SF_COL, (kb | k) != 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
    __syncwarp();
    wait_bar(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + lane_base) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x12345678u) sink[0] = (int)v;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main(int argc, char **argv) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = 1024 + (size_t)TN * 64;
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // ---- part 1
    long long bad_total = 0, cells = 0;
    int max_count = 0;
    for (int trial = 0; trial < 4; ++trial) {
        const int nkb = trial == 3 ? 2048 : 37;                     // 2048 K-blocks = a 512 x 512 mask
        const double dens[4] = {0.5, 0.05, 1.0, 0.9};
        std::vector<uint32_t> a((size_t)TM * nkb * 4), b((size_t)TN * nkb * 4);
        srand(1234 + trial);
        auto fill = [&](std::vector<uint32_t> &v) {
            for (auto &w : v) {
                w = 0;
                for (int i = 0; i < 32; ++i) w |= (uint32_t)((rand() / (double)RAND_MAX) < dens[trial]) << i;
            }
        };
        fill(a); fill(b);
        uint32_t *da, *db; float *dout;
        cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, (size_t)TM * TN * 4);
        cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
        check_kernel<<<1, 128, smem>>>(da, db, nkb, dout);
        std::vector<float> out((size_t)TM * TN);
        cudaError_t e = cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("check kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long bad = 0;
        for (int i = 0; i < TM; ++i)
            for (int j = 0; j < TN; ++j) {
                int want = 0;
                for (int w = 0; w < nkb * 4; ++w) want += __builtin_popcount(a[(size_t)i * nkb * 4 + w] & b[(size_t)j * nkb * 4 + w]);
                if (want > max_count) max_count = want;
                if (out[(size_t)i * TN + j] != (float)want) {
                    if (bad < 4) printf("  trial %d (%d,%d): got %.2f want %d\n", trial, i, j, out[(size_t)i * TN + j], want);
                    ++bad;
                }
            }
        printf("trial %d: density %.2f, %d K-blocks: %lld of %d cells differ\n", trial, dens[trial], nkb, bad, TM * TN);
        bad_total += bad; cells += TM * TN;
        cudaFree(da); cudaFree(db); cudaFree(dout);
    }
    // ---- part 2
    int *sink;
    cudaMalloc(&sink, 4);
    const int kblocks = 1 << 17;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        peak_kernel<<<sms, 128, smem>>>(kblocks, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tops = 2.0 * TM * TN * KB * (double)kblocks * sms / (ms * 1e-3) / 1e12;
        if (rep > 0 && tops > best) best = tops;
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("exact counts: %s (%lld of %lld cells differ, largest count %d)\n", bad_total ? "NO" : "yes", bad_total, cells, max_count);
    printf("tcgen05.mma.kind::mxf4 M128 N256 K64, %d SMs, A in TMEM, no operand traffic: %.1f TOP/s (%s)\n", sms, best,
           cudaGetErrorString(err));
    if (argc > 1 && err == cudaSuccess) {
        FILE *f = fopen(argv[1], "w");
        fprintf(f, "{\n \"mxf4_tops\": %.1f,\n \"exact_counts\": %s,\n \"largest_count_checked\": %d,\n"
                   " \"how\": \"tools/micro/mxf4_check.cu: tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 M128 N256 K64 (E2M1 operands, "
                   "UE8M0 scales all 1.0) issued back to back on %d SMs, A in tensor memory, B in shared memory, operands resident, "
                   "best of 5 launches of %d K-blocks (128 px) per CTA, CUDA events; counts compared with AND + POPC on the host\",\n \"sms\": %d\n}\n",
                best, bad_total ? "false" : "true", max_count, sms, kblocks, sms);
        fclose(f);
    }
    return (err == cudaSuccess && bad_total == 0) ? 0 : 1;
}
