// mask_overlap_tc.cu -- pairwise mask intersection counts on the 5th-gen tensor cores (sm_100a).
//
// inter[i,j] = sum_p m_i[p] * m_j[p] is a dense R x HW x R contraction of 0/1 operands
// (2.1 TFLOP per image at R = 2000, 512x512 masks), far too much for the
// popcount pipe (mask_overlap.cu: 16 POPC lanes / clk / SM).  Here it runs as
// tcgen05.mma.kind::mxf4 (M = 128, N = 256, K = 64; E2M1 operands, UE8M0 block scales all 2^0) with fp32
// accumulation in TMEM -- every product is exactly 1.0 and the sums stay below 2^24, so the counts are exact.
//
// Operands never exist as bytes in HBM: masks stay bit-packed (32 px / word).  A set pixel becomes a single-bit
// E2M1 nibble -- 0b0001 = 0.5, 0b0010 = 1.0, 0b0100 = 2.0 -- with one AND (sometimes a shift + an AND) per 8
// pixels, the bit position chosen per pixel so that the A operand's value times the B operand's value is 1.0:
//     pixel 4j, 4j+1, 4j+2, 4j+3 of a word:   A = 0.5, 1.0, 2.0, 2.0     B = 2.0, 1.0, 0.5, 0.5
// The pixel -> K-slot order inside a K-block is a fixed permutation, the same for both operands, which a
// contraction ignores.  (Rounds 1-2 used kind::i8 with 0xFF bytes: twice the operand bytes.  What bounded that
// kernel was the SHARED-MEMORY PORT -- per 128-pixel K-block 32 KB of expanded B stored by the expanders plus the same
// 32 KB read by the tensor core, 512 of the 531 cycles the four MMAs take -- not the expanders' ALU work: halving
// the ALU work alone changed nothing.  Nibbles halve both streams and the MMA runs at twice the int8 rate,
// tools/micro/mxf4_check.cu: exactness check + measured peak.)
//
// Warp roles of one CTA (one 128 x 256 output tile, K = all pixels, 128 px per K-block):
//   expanders (24)      ONE operand row per thread; two groups that ALTERNATE K-blocks, so the expand phase of one group
//                       overlaps the store + release-fence phase of the other.  Per group:
//                         A rows (128, 4 warps): written to TENSOR MEMORY with one tcgen05.st.32x32b.x16
//                           (lane = row, 16 columns per K-block);
//                         B rows (256, 8 warps): stored into the canonical K-major SWIZZLE_64B smem layout
//                           (8-row x 64 B atoms, chunk ^= (row / 2) % 4) with four STS.128.
//                       Every thread prefetches the 16 packed bytes of ITS row PF K-blocks (of its parity) ahead with
//                       cp.async (LDGSTS, L1 bypass) into a 16-byte smem slot nobody else touches
//                       (cp.async.wait_group, no barrier); the tile's visited K-blocks come from a list in smem built
//                       once per tile.  (Round 2 had 16 expander warps with TWO B rows per thread: the kernel was
//                       bound by the serial latency of a B thread's iteration -- ~1560 cycles for ~160 dependent
//                       instructions at 4.5 warps per scheduler -- not by any pipe)
//   MMA issuer (1)      waits for a stage, issues 2 x tcgen05.mma (A from TMEM, B from smem),
//                       tcgen05.commit hands the stage back.
//   epilogue            expander warps 0-15: tcgen05.ld, fp32 div.rn, cvt.rn.f16 into smem tiles, then
//                       contiguous global stores: tile rows (direct block) and tile columns (mirror
//                       block; asy is not symmetric, so the mirror carries inter / area_row).
// Tiles work in SORTED index space (mask_sort_kernel) and visit only the K-blocks where both operand
// blocks are non-zero (AND of the blocks' union bitmaps; with the tiled 8 x 16 pixel layout of
// cim_mask_pack_tiled a K-block is an image patch, so this follows the 2-D footprint of the masks);
// the caller un-permutes the maps afterwards.
// History of this kernel, with the ncu evidence, is in DESIGN.md section 4.3.
#include "common.cuh"

namespace {

constexpr int TM = 128, TN = 256;       // output tile = UMMA M x N
constexpr int KB = 128;                 // pixels per K-block = 64 operand bytes per row: one SW64 atom row, 2 MMAs
constexpr int ROW_BYTES = KB / 2;       // expanded operand bytes per row and K-block (E2M1: 2 pixels per byte)
constexpr int A_COLS = ROW_BYTES / 4;   // 16 TMEM columns of expanded A operand per stage
constexpr int TMEM_COLS = 512;          // accumulator: columns 0..255; A stages: 256 + 16 s; scale factors: 448..511
constexpr int TMEM_A0 = TN;
constexpr int TMEM_SF = 448, SF_COLS = 64;   // UE8M0 1.0 (0x7F) in every byte: whatever layout the MMA reads, the scale is 1
constexpr int B_BYTES = TN * ROW_BYTES; // 16 KB of expanded B operand per stage
constexpr int NEXP = 24;                // expander warps: 0-3 A even, 4-7 A odd, 8-15 B even, 16-23 B odd K-blocks
constexpr int NEPI = 16;                // warps 0-15 run the epilogue
constexpr int MMA_WARP = NEXP;
constexpr int THREADS = (NEXP + 1) * 32;
constexpr int FULL_ARRIVALS = 4 + 8;    // warps that fill one stage
constexpr size_t SI_BYTES = (size_t)(TM + 2 * TN) * 4 + 3 * (size_t)TM * 130 * 2;   // epilogue staging (aliases the rings)

constexpr int KLIST = 16384;            // smem list of a tile's visited K-blocks: masks up to 2 Mpixel
constexpr int SLOT_BYTES = NEXP * 32 * 16;   // one private 16-byte cp.async slot per expander thread and prefetch depth

// STAGES_: expanded operand stages (B in smem, A in tensor memory).  PF_: K-blocks (of a thread's parity) in flight.
// (A register prefetch with ld.global.nc was measured in round 1: it needs L1 for its outstanding misses and crawls once
// the stages leave little of it; a loader warp + staging ring serialised ~850 dependent instructions per 4 K-blocks.)
template <int STAGES_, int PF_ = 4>
struct Cfg {
    static constexpr int STAGES = STAGES_;
    static constexpr int PF = PF_;
    static constexpr size_t RING_BYTES = (size_t)STAGES * B_BYTES + (size_t)KLIST * 2 + (size_t)PF_ * SLOT_BYTES;
    static constexpr size_t BODY_BYTES = SI_BYTES > RING_BYTES ? SI_BYTES : RING_BYTES;
    static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + BODY_BYTES + 256;
    static_assert(TN + STAGES * A_COLS <= TMEM_SF, "TMEM budget");
    static_assert(SMEM_BYTES <= 227 * 1024, "smem budget");
};

// tcgen05 block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): a_format = b_format = 1
// (MXF4Format::E2M1), both operands K-major, N >> 3 at bit 17, scale format UE8M0 (bit 23), M >> 4 at bit 24,
// scale-factor ids 0, K = 64
constexpr uint32_t IDESC = (1u << 7) | (1u << 10) | ((TN >> 3) << 17) | (1u << 23) | ((TM >> 4) << 24);

// shared memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major SWIZZLE_64B:
// start address >> 4, LBO (unused for swizzled K-major) = 1, SBO = 512 B between 8-row groups,
// version 1 (Blackwell), layout type 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
           (4ull << 61);
}
// byte offset of 16-byte chunk c (0..3) of operand row r inside a B stage
__device__ __forceinline__ uint32_t b_chunk_off(int r, int c) {
    return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + ((c ^ ((r >> 1) & 3)) << 4));
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
// D[tmem] (+)= A[tmem] * B[smem], block scales from TMEM (all 1.0)
__device__ __forceinline__ void tc_mma_mxf4_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t tmem_sf,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%5], p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(tmem_sf)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
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

// 32 mask bits -> 32 E2M1 nibbles (4 words): nibble j of word g is non-zero iff bit 4j+g is set.
// A operand: values 0.5, 1.0, 2.0, 2.0 for g = 0..3;  B operand: 2.0, 1.0, 0.5, 0.5 -- every product is 1.0.
__device__ __forceinline__ void expand32_a(uint32_t w, uint32_t *o) {
    o[0] = w & 0x11111111u; o[1] = w & 0x22222222u; o[2] = w & 0x44444444u; o[3] = (w >> 1) & 0x44444444u;
}
__device__ __forceinline__ void expand32_b(uint32_t w, uint32_t *o) {
    o[0] = (w << 2) & 0x44444444u; o[1] = w & 0x22222222u; o[2] = (w >> 2) & 0x11111111u; o[3] = (w >> 3) & 0x11111111u;
}
__device__ __forceinline__ void sts16(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// fp16(RN(fp32(i) / fp32(d))) -- the reference's float32 division cast to float16 (create_cob_iou.py:48) -- without
// an IEEE division per element: q = i * (1/d) with an approximate reciprocal is within ~2.5 fp32 ulps of the exact
// quotient, so it rounds to the same float16 unless it lies that close to a float16 rounding boundary (the 13 fp32
// mantissa bits below float16's: 0x1000 = the midpoint).  Those cases (~0.2 %), the float16 subnormal range and NaN
// (0 / 0 of empty masks) take the exact division.  32 K x 3 IEEE divisions per tile were a quarter of the kernel.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ __half ratio_f16(float fi, float fd, float rcp_d) {
    float q = fi * rcp_d;
    const int low = (int)(__float_as_uint(q) & 0x1FFFu) - 0x1000;
    const bool easy = (q == 0.f || q >= 6.2e-5f) && (low > 8 || low < -8);      // false for NaN
    if (!easy) q = __fdiv_rn(fi, fd);
    return __float2half_rn(q);
}

template <class K>
__global__ void __launch_bounds__(THREADS, 1)   // 25 warps: 65536 / 800 -> 80 registers
mask_overlap_tc_kernel(const uint32_t *__restrict__ packed, const int32_t *__restrict__ area_all,
                       const int32_t *__restrict__ perm_all, const uint32_t *__restrict__ umap_a,
                       const uint32_t *__restrict__ umap_b, int bw, unsigned long long *__restrict__ visited,
                       const int32_t *__restrict__ tile_order, int n, long long words, int n_img, int32_t *__restrict__ inter_all, __half *__restrict__ iou_all,
                       __half *__restrict__ asy_all) {
    constexpr int STAGES = K::STAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *stages = smem;                                         // [STAGES][16 KB] expanded B
    uint16_t *klist = reinterpret_cast<uint16_t *>(stages + (size_t)STAGES * B_BYTES);   // [KLIST] visited K-blocks
    unsigned char *slots = reinterpret_cast<unsigned char *>(klist + KLIST);             // [PF][768 threads][16 B]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + K::BODY_BYTES);  // [STAGES] stage expanded
    uint64_t *empty = full + STAGES;                                      // [STAGES] stage consumed by the MMAs
    uint64_t *accum_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile id -> (d, img, ti), tj = ti / 2 + d: tiles are ordered by their distance d from the diagonal,
    // over all images.  After the locality sort the K-range of a tile shrinks with d, so the longest
    // tiles start first and the short ones fill the tail (tiles_per_img is unused by this order).
    const int ncb = (n + TN - 1) / TN, nrb = (n + TM - 1) / TM;
    int img, ti, tj;
    if (tile_order) {                        // longest tile first (mask_rank_kernel)
        const int code = __ldg(tile_order + blockIdx.x);
        img = code >> 16; ti = (code >> 8) & 255; tj = code & 255;
    } else {
        int d = 0, rem = blockIdx.x;
        for (;;) {
            const int cnt = min(nrb, 2 * (ncb - d)) * n_img;      // row blocks ti with ti / 2 + d < ncb
            if (rem < cnt) break;
            rem -= cnt;
            ++d;
        }
        const int per_img = min(nrb, 2 * (ncb - d));
        img = rem / per_img; ti = rem - img * per_img; tj = (ti >> 1) + d;
    }
    const int row0 = ti * TM, col0 = tj * TN;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], FULL_ARRIVALS); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        fence_mbar_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // All indices below are SORTED positions (mask_sort_kernel); perm maps them to the stored masks.
    // K-blocks of this tile: those where BOTH operand blocks have a non-zero row (AND of the two union
    // bitmaps).  Every warp counts them.
    const int32_t *perm = perm_all + (size_t)img * n;
    const uint32_t *ua = umap_a + ((size_t)img * nrb + ti) * bw, *ub = umap_b + ((size_t)img * ncb + tj) * bw;
    int nkb = 0;
    for (int j = lane; j < bw; j += 32) nkb += __popc(__ldg(ua + j) & __ldg(ub + j));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nkb += __shfl_xor_sync(0xffffffffu, nkb, o);
    if (tid == 0 && visited) atomicAdd(visited, (unsigned long long)nkb);

    // list of the visited K-blocks, in ascending order: warp w takes the bitmap words 32 c .. 32 c + 31 for
    // c = w, w + 25, ...; the offset of a chunk is the popcount of everything before it
    for (int c = warp; c * 32 < bw; c += THREADS / 32) {
        int pre = 0;
        for (int j = lane; j < c * 32; j += 32) pre += __popc(__ldg(ua + j) & __ldg(ub + j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
        const int j = c * 32 + lane;
        uint32_t mj = j < bw ? (__ldg(ua + j) & __ldg(ub + j)) : 0u;
        int incl = __popc(mj);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int off = pre + incl - __popc(mj);
        while (mj) {
            const int b = __ffs(mj) - 1;
            mj &= mj - 1;
            if (off < KLIST) klist[off] = (uint16_t)(j * 32 + b);
            ++off;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {     // block scales: UE8M0 1.0 in every byte of the scale-factor columns, all 128 lanes
        uint32_t one[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) one[i] = 0x7F7F7F7Fu;
        for (int c = 0; c < SF_COLS; c += 16) tc_st16(tmem_base + ((uint32_t)(32 * warp) << 16) + TMEM_SF + c, one);
        tc_fence_before();
    }
    __syncthreads();    // the scale columns are written by warps 0-3 and read by the issuer's MMAs
    tc_fence_after();
    if (warp < NEXP) {
        // ------------------------------------------------------------------ expanders
        const bool is_a = warp < 8;
        const int grp = is_a ? (warp >> 2) : ((warp - 8) >> 3);          // K-blocks i == grp (mod 2)
        // A: tile row 32 (warp % 4) + lane (= TMEM lane; a warp reaches the lane quarter warp % 4).
        // B: tile row 32 ((warp - 8) % 8) + lane.
        const int r = is_a ? 32 * (warp & 3) + lane : 32 * ((warp - 8) & 7) + lane;
        const uint32_t a_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + TMEM_A0;
        uint32_t chunk[4];                                    // smem addresses of the B row's 4 swizzled chunks, stage 0
#pragma unroll
        for (int c = 0; c < 4; ++c) chunk[c] = smem_u32(stages) + b_chunk_off(r, c);
        const int g = is_a ? row0 + r : col0 + r;
        const bool valid = g < n;
        const uint32_t *rp = packed + ((size_t)img * n + (size_t)(valid ? __ldg(perm + g) : 0)) * words;
        const int nj = (nkb - grp + 1) >> 1;                  // this group's K-blocks: i = 2 j + grp < nkb
        constexpr int PF = K::PF;
        const uint32_t klist_s = smem_u32(klist);
        const uint32_t slot0 = smem_u32(slots) + (uint32_t)tid * 16u;
        // one K-block of this thread: expand (before the wait, so that less sits between the stage's release and its
        // next MMAs), wait for the stage, store, hand the stage to the MMA issuer
        auto process = [&](int j, const uint4 &p) {
            const int i = 2 * j + grp, u = i / STAGES, s = i - u * STAGES;
            uint32_t o[16];
            const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
            if (is_a) {                                       // warp-uniform
#pragma unroll
                for (int q = 0; q < 4; ++q) expand32_a(pw[q], o + 4 * q);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) expand32_b(pw[q], o + 4 * q);
            }
            if (u > 0) mbar_wait(&empty[s], (u - 1) & 1);
            if (is_a) {
                tc_st16(a_lane + (uint32_t)(s * A_COLS), o);          // includes tcgen05.wait::st
                tc_fence_before();
            } else {
                const uint32_t so = (uint32_t)s * B_BYTES;
#pragma unroll
                for (int q = 0; q < 4; ++q) sts16(chunk[q] + so, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
#ifndef CIM_OV_ABL_NOFENCE
                fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core
#endif
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        };
        auto fetch = [&](int j, int d) {       // one cp.async group per call, also when there is nothing to load
            if (j < nj) {
                uint32_t kbi;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(kbi) : "r"(klist_s + (uint32_t)(2 * j + grp) * 2u) : "memory");
#ifndef CIM_OV_ABL_NOLOAD
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot0 + (uint32_t)d * SLOT_BYTES),
                             "l"(rp + (size_t)kbi * 4), "r"(valid ? 16u : 0u) : "memory");
#else
                if (kbi == 0xFFFFFu) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot0 + (uint32_t)d * SLOT_BYTES),
                             "l"(rp + (size_t)kbi * 4), "r"(valid ? 16u : 0u) : "memory");
#endif
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int d = 0; d < PF; ++d) fetch(d, d);
        for (int j0 = 0; j0 < nj; j0 += PF) {
#pragma unroll
            for (int d = 0; d < PF; ++d) {
                const int j = j0 + d;
                if (j < nj) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
                    uint4 p;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(p.x), "=r"(p.y), "=r"(p.z), "=r"(p.w) : "r"(slot0 + (uint32_t)d * SLOT_BYTES) : "memory");
                    process(j, p);                       // consumes p: the slot may be refilled now
                    fetch(j + PF, d);
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        // ------------------------------------------------------------------ MMA issuer
        // ONE thread runs the whole loop (waits included): with the loop around an `if (lane == 0)` the compiler
        // wraps every tcgen05 instruction into ELECT / R2UR / vote sequences (~100 dependent instructions per
        // K-block of this single warp).  Stage index and phase are compile-time / incremental, descriptors are
        // one add away from a per-tile base.
        if (elect_one() && nkb > 0) {        // elect.sync, not lane == 0: see common.cuh
            const uint64_t bd0 = smem_desc(smem_u32(stages));
            const uint32_t a0 = tmem_base + TMEM_A0, sf = tmem_base + TMEM_SF;
            uint32_t ph = 0;
            for (int kb0 = 0; kb0 < nkb; kb0 += STAGES) {
#pragma unroll
                for (int s = 0; s < STAGES; ++s) {
                    const int kb = kb0 + s;
                    if (kb < nkb) {
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint64_t bd = bd0 + (uint64_t)(s * (B_BYTES >> 4));
                        const uint32_t a_t = a0 + (uint32_t)(s * A_COLS);
#ifndef CIM_OV_ABL_NOMMA
                        tc_mma_mxf4_ts(tmem_base, a_t, bd, sf, kb != 0);
#pragma unroll
                        for (int k = 1; k < KB / 64; ++k)       // K = 64 nibbles: 8 TMEM columns of A, +2 (x16 B) of B
                            tc_mma_mxf4_ts(tmem_base, a_t + 8 * k, bd + 2 * k, sf, 1u);
#endif
                        tc_commit(&empty[s]);                    // arrives when the MMAs above have read the stage
                    }
                }
                ph ^= 1;
            }
            tc_commit(accum_full);                               // ... when every MMA of the tile has completed
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------------- epilogue
    // Expander warps 0-15 drain the accumulator: warp w reads TMEM lanes 32 (w % 4).. and, per pass of
    // 128 columns, the 32-column slice w / 4.  Areas are cached in smem; the fp16 results (iou, asy and
    // the mirror's asy) go to smem tiles first so that every global store is a contiguous segment:
    // a row of the tile (direct block) or a column of it (mirror block, written transposed).
    constexpr int EP = 130;                                       // half pitch (65 words: odd -> conflict-free)
    int *areaR = reinterpret_cast<int *>(stages);                 // [128]
    int *areaC = areaR + TM;                                      // [256]
    float *rcpC = reinterpret_cast<float *>(areaC + TN);          // [256] approximate 1 / area of the tile's columns
    __half *s_iou = reinterpret_cast<__half *>(rcpC + TN);        // [128][EP]
    __half *s_asy = s_iou + TM * EP;                              // [128][EP]  inter / area_col
    __half *s_asyT = s_asy + TM * EP;                             // [128][EP]  inter / area_row (mirror block)
    const int32_t *area = area_all + (size_t)img * n;
    int32_t *inter = inter_all ? inter_all + (size_t)img * n * n : nullptr;
    __half *iou = iou_all + (size_t)img * n * n;
    __half *asy = asy_all + (size_t)img * n * n;
    if (warp < NEPI) {
        if (nkb > 0) mbar_wait(accum_full, 0);                   // empty K-range: nothing was accumulated
        tc_fence_after();
#ifdef CIM_OV_ABL_NOEPI
        if (n > 0) goto epi_done;
#endif
        for (int i = tid; i < TM + TN; i += NEPI * 32) {
            const int gidx = i < TM ? row0 + i : col0 + (i - TM);
            const int a = gidx < n ? area[perm[gidx]] : 0;
            areaR[i] = a;                                         // areaC follows areaR in memory
            if (i >= TM) rcpC[i - TM] = rcp_approx((float)a);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
        const int q4 = warp & 3, cs = warp >> 2;
        const int rl = 32 * q4 + lane;                            // TMEM lane = tile row
        const int r = row0 + rl;
        const int a_r = areaR[rl];
        const float rcp_r = rcp_approx((float)a_r);
        const bool vec_ok = (n & 3) == 0;
#pragma unroll 1
        for (int pass = 0; pass < TN / 128; ++pass) {
            const int cl0 = pass * 128 + cs * 32;                 // first tile column of this warp's slice
            int v[32];
            if (nkb > 0) {
                tc_ld32(tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)cl0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float2int_rn(__int_as_float(v[j]));     // exact: integers < 2^24
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) {                     // two columns per 32-bit smem store
                __half h_iou[2], h_asy[2], h_asyT[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int a_c = areaC[cl0 + j + e];
                    const float fi = (float)v[j + e], fu = (float)(a_r + a_c - v[j + e]);
                    h_iou[e] = ratio_f16(fi, fu, rcp_approx(fu));
                    h_asy[e] = ratio_f16(fi, (float)a_c, rcpC[cl0 + j + e]);
                    h_asyT[e] = ratio_f16(fi, (float)a_r, rcp_r);
                }
                const int so = rl * EP + cs * 32 + j;             // even: EP, cs * 32 and j are
                *reinterpret_cast<__half2 *>(s_iou + so) = __halves2half2(h_iou[0], h_iou[1]);
                *reinterpret_cast<__half2 *>(s_asy + so) = __halves2half2(h_asy[0], h_asy[1]);
                *reinterpret_cast<__half2 *>(s_asyT + so) = __halves2half2(h_asyT[0], h_asyT[1]);
            }
            // mirror needed?  (c, r) belongs to tile (c / 128, r / 256): computed itself iff
            // 256 (ti / 2 + 1) > 128 (c / 128); c / 128 is constant over a pass
            const int cpass = col0 + pass * 128;
            const bool mirror = !(TN * ((ti >> 1) + 1) > TM * (cpass / TM));
            if (inter && r < n) {                                 // integer counts: tests only, simple stores
                for (int j = 0; j < 32 && col0 + cl0 + j < n; ++j) {
                    inter[(size_t)r * n + col0 + cl0 + j] = v[j];
                    if (mirror) inter[(size_t)(col0 + cl0 + j) * n + r] = v[j];
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
            // direct block: warp w writes tile rows w, w + 16, ...; a lane covers 4 consecutive columns
            for (int rr = warp; rr < TM; rr += NEPI) {
                const int gr = row0 + rr, gc = cpass + 4 * lane;
                if (gr >= n || gc >= n) continue;
                const uint32_t *pi = reinterpret_cast<const uint32_t *>(s_iou + rr * EP + 4 * lane);
                const uint32_t *pa = reinterpret_cast<const uint32_t *>(s_asy + rr * EP + 4 * lane);
                const size_t o = (size_t)gr * n + gc;
                if (vec_ok && gc + 4 <= n) {
                    *reinterpret_cast<uint2 *>(iou + o) = make_uint2(pi[0], pi[1]);
                    *reinterpret_cast<uint2 *>(asy + o) = make_uint2(pa[0], pa[1]);
                } else {
                    for (int e = 0; e < 4 && gc + e < n; ++e) {
                        iou[o + e] = s_iou[rr * EP + 4 * lane + e];
                        asy[o + e] = s_asy[rr * EP + 4 * lane + e];
                    }
                }
            }
            // mirror block: warp w writes output rows (= tile columns) w, w + 16, ...; lanes cover the
            // 128 tile rows in 4 strides of 32 (conflict-free transposed smem reads, 64 B global segments)
            if (mirror) {
                for (int cc = warp; cc < 128; cc += NEPI) {
                    const int gc = cpass + cc;
                    if (gc >= n) break;
#pragma unroll
                    for (int h = 0; h < TM / 32; ++h) {
                        const int rr = h * 32 + lane, gr = row0 + rr;
                        if (gr >= n) continue;
                        const size_t o = (size_t)gc * n + gr;
                        iou[o] = s_iou[rr * EP + cc];
                        asy[o] = s_asyT[rr * EP + cc];
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
        }
    }
#ifdef CIM_OV_ABL_NOEPI
epi_done:
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

#ifndef CIM_OV_STAGES           // tuning knobs
#define CIM_OV_STAGES 8
#endif
#ifndef CIM_OV_PF
#define CIM_OV_PF 4
#endif

template <class K>
static int launch(const uint32_t *packed, const int32_t *area, const int32_t *perm, const uint32_t *umap_a,
                  const uint32_t *umap_b, int bw, unsigned long long *visited, const int32_t *tile_order, int n_img,
                  int n, long long words, int32_t *inter, __half *iou, __half *asy, cudaStream_t st) {
    const int nrb = (n + TM - 1) / TM, ncb = (n + TN - 1) / TN;
    int tiles = 0;
    for (int i = 0; i < nrb; ++i) tiles += ncb - (i >> 1);
    cudaFuncSetAttribute(mask_overlap_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    mask_overlap_tc_kernel<K><<<(unsigned)(tiles * n_img), THREADS, K::SMEM_BYTES, st>>>(
        packed, area, perm, umap_a, umap_b, bw, visited, tile_order, n, words, n_img, inter, iou, asy);
    return cim_launch_status();
}

}  // namespace

using CfgRun = Cfg<CIM_OV_STAGES, CIM_OV_PF>;

// true when the tensor-core path can take this problem (else the popcount kernel runs): the tile's K-block list must
// fit its smem array (masks up to 2 Mpixel)
bool cim_mask_overlap_tc_eligible(int n, long long words) {
    return n >= 64 && words >= 4 && (words % 4) == 0 && (words + 3) / 4 <= KLIST &&
           (size_t)cim_max_smem_optin() >= CfgRun::SMEM_BYTES;
}

int cim_mask_overlap_tc_launch(const uint32_t *packed, const int32_t *area, const int32_t *perm,
                               const uint32_t *umap_a, const uint32_t *umap_b, int bw, unsigned long long *visited,
                               const int32_t *tile_order, int n_img, int n, long long words, int32_t *inter,
                               __half *iou, __half *asy, cudaStream_t st) {
    return launch<CfgRun>(packed, area, perm, umap_a, umap_b, bw, visited, tile_order, n_img, n, words, inter, iou, asy, st);
}
