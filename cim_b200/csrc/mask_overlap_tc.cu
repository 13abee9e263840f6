// mask_overlap_tc.cu -- pairwise mask intersection counts on the 5th-gen tensor cores (sm_100a).
//
// inter[i,j] = sum_p m_i[p] * m_j[p] is a dense R x HW x R contraction of 0/1 operands
// (2.1 TFLOP per image at R = 2000, 512x512 masks), far too much for the
// popcount pipe (mask_overlap.cu: 16 POPC lanes / clk / SM).  Here it runs as
// tcgen05.mma.kind::i8 with S32 accumulation in TMEM, which is exact.
//
// Operands never exist as bytes in HBM: masks stay bit-packed (32 px / word).  Per K-block of
// 128 pixels, 12 expander warps read 16 B of each of the tile's 128 + 256 mask rows and turn
// every bit into a byte 0x00 / 0xFF with PRMT's sign-replicate mode (one PRMT per 4 pixels) --
// 0xFF is -1 as INT8, so a pixel common to both masks contributes (-1)*(-1) = +1.
//   * The 256 B-operand rows are stored straight into the canonical K-major SWIZZLE_128B layout
//     the UMMA smem descriptor expects (8-row x 128 B atoms, 16 B chunk index XOR row % 8).
//   * The 128 A-operand rows never touch shared memory: each thread writes its row's 128 bytes
//     into tensor memory with one tcgen05.st.32x32b.x32 (lane = row, 32 columns per K-block) and
//     the MMA takes A from TMEM.  The first version staged A in smem as well and was bound by the
//     shared-memory port (L1 83 %, tensor pipe 45 %: profiles/r1_ncu_full_v1.txt).
//   * Each expanded B stage (128 rows) is used by TWO A tiles (2 x 128 rows, two accumulators):
//     at M = 128 the UMMA reads its B operand at 64 B/clk, expanding B at the same rate would
//     take the other half of the 128 B/clk shared-memory port (second version: L1 85 %, tensor
//     pipe 58 %: profiles/r1_ncu_full_v2.txt); reusing the stage halves the store traffic.
// A 4-stage mbarrier ring hands stages to the single MMA-issuing thread (M = 128, N = 128, K = 32
// per instruction, 2 x 4 per stage) and tcgen05.commit returns the stage.  The pixel -> K-slot order
// inside a K-block is a fixed permutation (the same for both operands), which a contraction does
// not care about.
//
// Only tiles touching the upper triangle are computed; the epilogue reads the accumulator from
// TMEM (tcgen05.ld), applies iou = I / (a_i + a_j - I), asy = I / a_j in fp32 -> fp16, and writes
// the mirror block through a shared-memory transpose (asy is not symmetric, I is).
#include "common.cuh"

namespace {

constexpr int TM = 256;                 // tile rows: two UMMA M = 128 sub-tiles sharing every B stage
constexpr int TN = 128;                 // tile cols   (UMMA N)
constexpr int UM = 128;                 // UMMA M
constexpr int KB = 128;                 // pixels (= operand bytes per row) per K-block: one SW128 atom
constexpr int STAGES = 4;
constexpr int B_BYTES = TN * KB;        // 16 KB of expanded B operand per stage (smem)
constexpr int STAGE_BYTES = B_BYTES;
constexpr int A_COLS = KB / 4;          // 32 TMEM columns of expanded A operand per stage and sub-tile
constexpr int EXP_WARPS = (TM + TN) / 32;          // 12 expander warps, one operand row per thread
constexpr int MMA_WARP = EXP_WARPS;                // warp 12 issues the MMAs and owns TMEM
constexpr int THREADS = (EXP_WARPS + 1) * 32;      // 416
constexpr int EPI_WARPS = 8;                       // warps 0..7 drain the two accumulators
constexpr int SPITCH = TN + 1;                     // int32 pitch of the transpose buffer
constexpr int TMEM_COLS = 512;                     // accumulators: 0..127 and 128..255; A stages: 256 + 64 s + 32 t
constexpr int TMEM_A0 = 2 * TN;
constexpr size_t SI_BYTES = (size_t)TM * SPITCH * 4;                 // 132 KB transpose buffer (epilogue)
constexpr size_t RING_BYTES = (size_t)STAGES * STAGE_BYTES;          // 64 KB
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (SI_BYTES > RING_BYTES ? SI_BYTES : RING_BYTES) + 128;

// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, INT8 x INT8,
// both operands K-major, N = 128, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((TN >> 3) << 17) | ((UM >> 4) << 24);

// shared memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major SWIZZLE_128B:
// start address >> 4, LBO (unused for swizzled K-major) = 1, SBO = 1024 B between 8-row groups,
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31,%32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 mask bits -> 32 operand bytes (8 words): word g, byte b = 0xFF iff bit 8b+g is set.
// (w << (7-g)) moves bit 8b+g to the top of byte b; PRMT selector 0xBA98 replicates each byte's
// sign bit over the byte.
// (inline PTX: the __byte_perm intrinsic masks the selector to 3 bits per byte and would drop the
// replicate flag.)
__device__ __forceinline__ uint32_t sign_bytes(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0u), "r"(0xBA98u));
    return r;
}
__device__ __forceinline__ void expand32(uint32_t w, uint32_t (&o)[8]) {
#pragma unroll
    for (int g = 0; g < 8; ++g) o[g] = sign_bytes(w << (7 - g));
}

__device__ __forceinline__ __half2 pack_ratio2(int i0, int d0, int i1, int d1) {
    return __halves2half2(__float2half_rn(__fdiv_rn((float)i0, (float)d0)),
                          __float2half_rn(__fdiv_rn((float)i1, (float)d1)));
}

__global__ void __launch_bounds__(THREADS, 1)
mask_overlap_tc_kernel(const uint32_t *__restrict__ packed, const int32_t *__restrict__ area_all, int n,
                       long long words, int tiles_per_img, int32_t *__restrict__ inter_all,
                       __half *__restrict__ iou_all, __half *__restrict__ asy_all) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *stages = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (SI_BYTES > RING_BYTES ? SI_BYTES : RING_BYTES));
    uint64_t *empty = full + STAGES;
    uint64_t *accum_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.x / tiles_per_img;
    // linear tile id -> (ti, tj): row block ti (256 rows) pairs with column blocks (128) tj >= 2 ti
    int ti = 0, rem = blockIdx.x % tiles_per_img;
    const int ncb = (n + TN - 1) / TN;
    while (rem >= ncb - 2 * ti) { rem -= ncb - 2 * ti; ++ti; }
    const int tj = 2 * ti + rem;
    const int row0 = ti * TM, col0 = tj * TN;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], EXP_WARPS); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        fence_mbar_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = (int)(words / 4);

    if (warp < EXP_WARPS) {
        // ------------------------------------------------------------------ expanders
        const bool is_a = tid < TM;                           // warps 0..7: sub-tile warp / 4, TMEM lanes 32 (warp % 4)..
        const int lr = is_a ? tid : tid - TM;                 // row inside the A / B tile
        const int grow = (is_a ? row0 : col0) + lr;           // mask index inside the image
        const bool valid = grow < n;
        const uint4 *src = reinterpret_cast<const uint4 *>(packed + ((size_t)img * n + (valid ? grow : 0)) * words);
        const uint32_t row_off = (lr >> 3) * 1024 + (lr & 7) * 128;
        const uint32_t sw = lr & 7;
        const uint32_t a_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + TMEM_A0 + (uint32_t)((warp >> 2) * A_COLS);
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        uint4 cur[2] = {zero4, zero4};                         // two K-blocks = one 32 B sector per row
        if (valid) { cur[0] = __ldg(src); if (nkb > 1) cur[1] = __ldg(src + 1); }
        for (int kb2 = 0; kb2 < nkb; kb2 += 2) {
            uint4 nxt[2] = {zero4, zero4};
            if (valid && kb2 + 2 < nkb) nxt[0] = __ldg(src + kb2 + 2);
            if (valid && kb2 + 3 < nkb) nxt[1] = __ldg(src + kb2 + 3);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kb = kb2 + h;
                if (kb >= nkb) break;
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
                const uint32_t pw[4] = {cur[h].x, cur[h].y, cur[h].z, cur[h].w};
                if (is_a) {
                    uint32_t o[32];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t t[8];
                        expand32(pw[q], t);
#pragma unroll
                        for (int g = 0; g < 8; ++g) o[q * 8 + g] = t[g];
                    }
                    tc_st32(a_lane + (uint32_t)(s * 2 * A_COLS), o);  // includes tcgen05.wait::st
                    tc_fence_before();
                } else {
                    unsigned char *rowp = stages + (size_t)s * STAGE_BYTES + row_off;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t o[8];
                        expand32(pw[q], o);
                        // 16 B chunks 2q and 2q+1 of the 128 B row, XOR-swizzled with row % 8
                        *reinterpret_cast<uint4 *>(rowp + (((2 * q) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(rowp + (((2 * q + 1) ^ sw) << 4)) =
                            make_uint4(o[4], o[5], o[6], o[7]);
                    }
                    fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
            cur[0] = nxt[0];
            cur[1] = nxt[1];
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&full[s], (kb / STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t bd = smem_desc(smem_u32(stages + (size_t)s * STAGE_BYTES));
                const uint32_t a_t = tmem_base + TMEM_A0 + (uint32_t)(s * 2 * A_COLS);
#pragma unroll
                for (int t = 0; t < 2; ++t)             // the two A sub-tiles share this B stage
#pragma unroll
                    for (int k = 0; k < KB / 32; ++k)   // K = 32 bytes: 8 TMEM columns of A, +2 (x16 B) of B
                        tc_mma_i8_ts(tmem_base + (uint32_t)(t * TN), a_t + (uint32_t)(t * A_COLS + 8 * k), bd + 2 * k,
                                     (kb | k) != 0);
                tc_commit(&empty[s]);                    // arrives when the MMAs above have read the stage
                if (kb == nkb - 1) tc_commit(accum_full);
            }
            __syncwarp();
        }
    }

    // ---------------------------------------------------------------------- epilogue
    int32_t *sI = reinterpret_cast<int32_t *>(stages);          // [TM][SPITCH], reuses the stage ring
    const int32_t *area = area_all + (size_t)img * n;
    int32_t *inter = inter_all ? inter_all + (size_t)img * n * n : nullptr;
    __half *iou = iou_all + (size_t)img * n * n;
    __half *asy = asy_all + (size_t)img * n * n;
    if (warp < EPI_WARPS) {
        mbar_wait(accum_full, 0);
        tc_fence_after();
        const int sub = warp >> 2;                                // accumulator / A sub-tile
        const int rl = UM * sub + 32 * (warp & 3) + lane;          // tile row
        const int r = row0 + rl;
        const int a_r = r < n ? area[r] : 0;
#pragma unroll 1
        for (int cc = 0; cc < TN; cc += 32) {
            int v[32];
            tc_ld32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(sub * TN + cc), v);
            const int cbase = col0 + cc;
#pragma unroll
            for (int j = 0; j < 32; ++j) sI[rl * SPITCH + cc + j] = v[j];
            if (r < n) {
                if (cbase + 32 <= n && (n & 7) == 0) {           // 16 B vector stores
                    uint4 *pi = reinterpret_cast<uint4 *>(iou + (size_t)r * n + cbase);
                    uint4 *pa = reinterpret_cast<uint4 *>(asy + (size_t)r * n + cbase);
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        __half2 hi[4], ha[4];
#pragma unroll
                        for (int j2 = 0; j2 < 4; ++j2) {
                            const int j = j8 * 8 + j2 * 2;
                            const int a0 = area[cbase + j], a1 = area[cbase + j + 1];
                            hi[j2] = pack_ratio2(v[j], a_r + a0 - v[j], v[j + 1], a_r + a1 - v[j + 1]);
                            ha[j2] = pack_ratio2(v[j], a0, v[j + 1], a1);
                        }
                        pi[j8] = *reinterpret_cast<uint4 *>(hi);
                        pa[j8] = *reinterpret_cast<uint4 *>(ha);
                    }
                } else {
                    for (int j = 0; j < 32; ++j) {
                        const int c = cbase + j;
                        if (c >= n) break;
                        const int a_c = area[c];
                        iou[(size_t)r * n + c] = __float2half_rn(__fdiv_rn((float)v[j], (float)(a_r + a_c - v[j])));
                        asy[(size_t)r * n + c] = __float2half_rn(__fdiv_rn((float)v[j], (float)a_c));
                    }
                }
                if (inter)
                    for (int j = 0; j < 32 && cbase + j < n; ++j) inter[(size_t)r * n + cbase + j] = v[j];
            }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        // mirror block out[c][r]: needed for the row sub-blocks (128 rows) whose own tile
        // (row block r / 256, column block c / 128) is not computed, i.e. c / 128 < 2 * (r / 256)
        for (int cl = warp; cl < TN; cl += EPI_WARPS) {
            const int c = col0 + cl;
            if (c >= n) break;
            const int a_c = area[c];
#pragma unroll
            for (int h = 0; h < TM / 32; ++h) {
                const int rl2 = h * 32 + lane, r2 = row0 + rl2;
                if (r2 >= n) continue;
                // element (c, r2) belongs to tile (c / 256, r2 / 128): computed iff r2 / 128 >= 2 * (c / 256)
                if ((r2 / TN) >= 2 * (c / TM)) continue;
                const int I = sI[rl2 * SPITCH + cl], a_r2 = area[r2];
                const size_t o = (size_t)c * n + r2;
                iou[o] = __float2half_rn(__fdiv_rn((float)I, (float)(a_c + a_r2 - I)));
                asy[o] = __float2half_rn(__fdiv_rn((float)I, (float)a_r2));
                if (inter) inter[o] = I;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

}  // namespace

// true when the tensor-core path can take this problem (else the popcount kernel runs)
bool cim_mask_overlap_tc_eligible(int n, long long words) {
    return n >= 64 && words >= 4 && (words % 4) == 0 && (size_t)cim_max_smem_optin() >= SMEM_BYTES;
}

int cim_mask_overlap_tc_launch(const uint32_t *packed, const int32_t *area, int n_img, int n, long long words,
                               int32_t *inter, __half *iou, __half *asy, cudaStream_t st) {
    const int nrb = (n + TM - 1) / TM, ncb = (n + TN - 1) / TN;
    int tiles = 0;
    for (int i = 0; i < nrb; ++i) tiles += ncb - 2 * i;
    cudaFuncSetAttribute(mask_overlap_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    mask_overlap_tc_kernel<<<(unsigned)(tiles * n_img), THREADS, SMEM_BYTES, st>>>(packed, area, n, words, tiles,
                                                                                  inter, iou, asy);
    return cim_launch_status();
}
