// mask_overlap_tc.cu -- pairwise mask intersection counts on the 5th-gen tensor cores (sm_100a).
//
// inter[i,j] = sum_p m_i[p] * m_j[p] is a dense R x HW x R contraction of 0/1 operands
// (2.1 TFLOP per image at R = 2000, 512x512 masks), far too much for the
// popcount pipe (mask_overlap.cu: 16 POPC lanes / clk / SM).  Here it runs as
// tcgen05.mma.kind::i8 (M = 128, N = 256, K = 32) with S32 accumulation in TMEM, which is exact.
//
// Operands never exist as bytes in HBM: masks stay bit-packed (32 px / word).  A bit becomes a
// byte 0x00 / 0xFF with PRMT's sign-replicate mode (one PRMT per 4 pixels); 0xFF is -1 as INT8, so
// a pixel common to both masks contributes (-1)*(-1) = +1.  The pixel -> K-slot order inside a
// K-block is a fixed permutation, the same for both operands, which a contraction ignores.
//
// Warp roles of one CTA (one 128 x 256 output tile, K = all pixels, 128 px per K-block):
//   loader (1 warp)     cp.async (LDGSTS, L1 bypass) of 16 B per operand row and K-block into a
//                       3-buffer smem staging ring, 4 K-blocks per group, completion through
//                       cp.async.mbarrier.arrive.noinc.
//   expanders (16)      two groups that ALTERNATE K-blocks, so the ALU phase (expand) of one group
//                       overlaps the store + release-fence phase of the other:
//                         A rows (128): one row per thread, written to TENSOR MEMORY with one
//                           tcgen05.st.32x32b.x32 (lane = row, 32 columns per K-block);
//                         B rows (256): two rows per thread, stored into the canonical K-major
//                           SWIZZLE_128B smem layout (8-row x 128 B atoms, chunk ^= row % 8).
//   MMA issuer (1)      waits for a stage, issues 4 x tcgen05.mma (A from TMEM, B from smem),
//                       tcgen05.commit hands the stage back.
//   epilogue            the 16 expander warps: tcgen05.ld, fp32 div.rn, cvt.rn.f16 into smem tiles, then
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
constexpr int KB = 128;                 // pixels (= operand bytes per row) per K-block: one SW128 atom
constexpr int A_COLS = KB / 4;          // 32 TMEM columns of expanded A operand per stage
constexpr int TMEM_COLS = 512;          // accumulator: columns 0..255; A stages: 256 + 32 s
constexpr int TMEM_A0 = TN;
constexpr int ROWS = TM + TN;           // operand rows per tile
constexpr int GK = 4;                   // K-blocks per load group
constexpr int NBUF = 3;                 // staging buffers [GK][ROWS][16 B]
constexpr int BUF_BYTES = GK * ROWS * 16;
constexpr int B_BYTES = TN * KB;        // 32 KB of expanded B operand per stage
constexpr int NEXP = 16;                // expander warps: 0-3 A even, 4-7 A odd, 8-11 B even, 12-15 B odd K-blocks
constexpr int MMA_WARP = NEXP, LOAD_WARP = NEXP + 1;
constexpr int THREADS = (NEXP + 2) * 32;
constexpr int KMAP_WORDS = 512;         // smem copy of a tile's K-block bitmap (16384 K-blocks = 2 Mpixel masks)
constexpr size_t SI_BYTES = (size_t)(TM + TN) * 4 + 3 * (size_t)TM * 130 * 2;   // epilogue staging (aliases the rings)

constexpr int KLIST = 4096;             // direct variants: smem list of a tile's visited K-blocks (512 Kpixel masks)
constexpr int SLOT_BYTES = (8 * 32 + 2 * 8 * 32) * 16;  // cp.async variant: 16 B per (A thread) / 2 x 16 B per (B thread), private

// MODE_ 0: a loader warp streams packed rows into a staging ring with cp.async; K-blocks come from the tile's bitmap
//          (any mask size).
// MODE_ 2: no loader warp and no staging ring -- every expander thread prefetches the 16 B of ITS row(s) PF_ K-blocks
//          (of its parity) ahead with cp.async (LDGSTS, L1 bypass) into 16-byte smem slots private to the thread
//          (cp.async.wait_group, no barrier); the visited K-blocks come from a list in smem built once per tile.
//          (A register prefetch with ld.global.nc was measured too: it needs L1 for its outstanding misses and
//          crawls once the stages leave little of it -- 1.99 ms with 6 stages against 1.22 ms with 4.)
template <int STAGES_, int MODE_ = 0, int PF_ = 4>
struct Cfg {
    static constexpr int STAGES = STAGES_;
    static constexpr int MODE = MODE_;
    static constexpr int PF = PF_;
    static constexpr bool DIRECT = MODE_ != 0;
    static constexpr size_t LOAD_BYTES = MODE_ == 0 ? (size_t)NBUF * BUF_BYTES : (size_t)KLIST * 2 + (size_t)PF_ * SLOT_BYTES;
    static constexpr size_t RING_BYTES = (size_t)STAGES * B_BYTES + LOAD_BYTES;
    static constexpr size_t BODY_BYTES = SI_BYTES > RING_BYTES ? SI_BYTES : RING_BYTES;
    static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + BODY_BYTES + 256 + (DIRECT ? 0 : KMAP_WORDS * 4);
    static_assert(MODE_ != 0 || STAGES % 2 == 0, "loader variant: the two expander groups own alternate stages");
    static_assert(TN + STAGES * A_COLS <= TMEM_COLS, "TMEM budget");
    static_assert(SMEM_BYTES <= 227 * 1024, "smem budget");
};

// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, INT8 x INT8, both
// operands K-major, N = 256, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((TN >> 3) << 17) | ((TM >> 4) << 24);

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
// sign bit over the byte.  Inline PTX: the __byte_perm intrinsic masks the selector to 3 bits per
// byte and would drop the replicate flag.
__device__ __forceinline__ uint32_t sign_bytes(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0u), "r"(0xBA98u));
    return r;
}
__device__ __forceinline__ void expand32(uint32_t w, uint32_t (&o)[8]) {
#pragma unroll
    for (int g = 0; g < 8; ++g) o[g] = sign_bytes(w << (7 - g));
}
// one operand row of a K-block (16 packed bytes) -> 128 operand bytes in the swizzled smem row
__device__ __forceinline__ void expand_row_to_smem(const uint4 &p, unsigned char *stage, const uint32_t (&choff)[8]) {
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t o[8];
        expand32(pw[q], o);
        *reinterpret_cast<uint4 *>(stage + choff[2 * q]) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(stage + choff[2 * q + 1]) = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

__device__ __forceinline__ void sts16(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts16_off(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0+%5], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d), "n"(OFF)
                 : "memory");
}

__device__ __forceinline__ __half2 pack_ratio2(int i0, int d0, int i1, int d1) {
    return __halves2half2(__float2half_rn(__fdiv_rn((float)i0, (float)d0)),
                          __float2half_rn(__fdiv_rn((float)i1, (float)d1)));
}

template <class K>
__global__ void __launch_bounds__(THREADS, 1)   // 18 warps = 5 on one SM sub-partition: 16384 / (5 * 32) -> 96 registers
mask_overlap_tc_kernel(const uint32_t *__restrict__ packed, const int32_t *__restrict__ area_all,
                       const int32_t *__restrict__ perm_all, const uint32_t *__restrict__ umap_a,
                       const uint32_t *__restrict__ umap_b, int bw, unsigned long long *__restrict__ visited,
                       const int32_t *__restrict__ tile_order, int n, long long words, int n_img, int32_t *__restrict__ inter_all, __half *__restrict__ iou_all,
                       __half *__restrict__ asy_all) {
    constexpr int STAGES = K::STAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *stages = smem;                                         // [STAGES][32 KB] expanded B
    unsigned char *staging = stages + (size_t)STAGES * B_BYTES;           // [NBUF][GK][ROWS][16 B] packed bits
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + K::BODY_BYTES);  // [STAGES] stage expanded
    uint64_t *empty = full + STAGES;                                      // [STAGES] stage consumed by the MMAs
    uint64_t *accum_full = empty + STAGES;
    uint64_t *loaded = accum_full + 1;             // [NBUF] staging buffer filled (32 loader lanes)
    uint64_t *consumed = loaded + NBUF;            // [NBUF] staging buffer read by every expander warp
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(consumed + NBUF);
    uint32_t *kmap = tmem_slot + 2;                // [KMAP_WORDS] AND of the two union bitmaps (loader)
    uint16_t *klist = reinterpret_cast<uint16_t *>(staging);   // direct variant: [KLIST] visited K-blocks

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile id -> (d, img, ti), tj = ti / 2 + d: tiles are ordered by their distance d from the diagonal,
    // over all images.  After the locality sort the K-range of a tile shrinks with d, so the longest
    // tiles start first and the short ones fill the tail (tiles_per_img is unused by this order).
    const int ncb = (n + TN - 1) / TN, nrb = (n + TM - 1) / TM;
    int img, ti, tj;
    if (tile_order) {                        // longest tile first (mask_tile_order_kernel)
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
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NEXP / 2); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        for (int i = 0; i < NBUF; ++i) { mbar_init(&loaded[i], 32); mbar_init(&consumed[i], NEXP); }
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
    // bitmaps).  Every warp counts them; only the loader needs their indices.
    const int32_t *perm = perm_all + (size_t)img * n;
    const uint32_t *ua = umap_a + ((size_t)img * nrb + ti) * bw, *ub = umap_b + ((size_t)img * ncb + tj) * bw;
    int nkb = 0;
    for (int j = lane; j < bw; j += 32) {
        const uint32_t mj = __ldg(ua + j) & __ldg(ub + j);
        nkb += __popc(mj);
        if (!K::DIRECT && warp == LOAD_WARP && j < KMAP_WORDS) kmap[j] = mj;      // the loader's private copy
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nkb += __shfl_xor_sync(0xffffffffu, nkb, o);
    const int ngroups = (nkb + GK - 1) / GK;
    if (tid == 0 && visited) atomicAdd(visited, (unsigned long long)nkb);

    if (K::DIRECT) {
        // list of the visited K-blocks, in ascending order: warp w takes the bitmap words 32 c .. 32 c + 31 for
        // c = w, w + 18, ...; the offset of a chunk is the popcount of everything before it
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
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < NEXP) {
        // ------------------------------------------------------------------ expanders
        const bool is_a = warp < 8;
        const int grp = (warp >> 2) & 1;                      // K-blocks kb == grp (mod 2)
        const int q4 = warp & 3;
        // A: tile row 32 q4 + lane (= TMEM lane).  B: tile rows 64 q4 + lane and + 32.
        const int ra = 32 * q4 + lane;
        const int rb0 = 64 * q4 + lane, rb1 = rb0 + 32;
        const uint32_t a_lane = tmem_base + ((uint32_t)(32 * q4) << 16) + TMEM_A0;
        uint32_t ch0[8], ch1[8];                              // swizzled chunk offsets of the two B rows
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            ch0[c] = (rb0 >> 3) * 1024 + (rb0 & 7) * 128 + ((c ^ (rb0 & 7)) << 4);
            ch1[c] = (rb1 >> 3) * 1024 + (rb1 & 7) * 128 + ((c ^ (rb1 & 7)) << 4);
        }
        if (K::DIRECT) {
            const uint32_t *img_base = packed + (size_t)img * n * words;
            const int g0 = is_a ? row0 + ra : col0 + rb0, g1 = col0 + rb1;
            const bool v0 = g0 < n, v1 = !is_a && g1 < n;
            const uint32_t *r0p = img_base + (size_t)(v0 ? __ldg(perm + g0) : 0) * words;
            const uint32_t *r1p = img_base + (size_t)(v1 ? __ldg(perm + g1) : 0) * words;
            const int nj = (nkb - grp + 1) >> 1;              // this group's K-blocks: i = 2 j + grp < nkb
            constexpr int PF = K::PF;
            // smem addresses of the first B row's 8 swizzled 16-byte chunks in stage 0; the second row (+ 32 rows,
            // same row % 8) sits 4096 B further on, a stage is B_BYTES further on
            uint32_t chunk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) chunk[c] = smem_u32(stages) + ch0[c];
            // one K-block of this thread: expand, wait for the stage, store, hand the stage to the MMA issuer.
            // The first row's ALU part (bits -> bytes in registers) happens BEFORE the wait, so that less sits
            // between the stage's release and its next MMAs.
            auto process = [&](int j, const uint4 &p0, const uint4 &p1) {
                const int i = 2 * j + grp, u = i / STAGES, s = i - u * STAGES;
                uint32_t o[32];
                {
                    const uint32_t pw[4] = {p0.x, p0.y, p0.z, p0.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t t[8];
                        expand32(pw[q], t);
#pragma unroll
                        for (int gg = 0; gg < 8; ++gg) o[q * 8 + gg] = t[gg];
                    }
                }
                if (u > 0) mbar_wait(&empty[s], (u - 1) & 1);
                if (is_a) {
                    tc_st32(a_lane + (uint32_t)(s * A_COLS), o);      // includes tcgen05.wait::st
                    tc_fence_before();
                } else {
                    const uint32_t so = (uint32_t)s * B_BYTES;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        sts16(chunk[q] + so, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                    const uint32_t pw[4] = {p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t t[8];
                        expand32(pw[q], t);
                        sts16_off<4096>(chunk[2 * q] + so, t[0], t[1], t[2], t[3]);
                        sts16_off<4096>(chunk[2 * q + 1] + so, t[4], t[5], t[6], t[7]);
                    }
                    fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            };
            // slot d of this thread: [d][A threads 256 x 16 B | B threads' first rows 256 x 16 B | second rows]
            const uint32_t slot0 = smem_u32(staging) + KLIST * 2 +
                                   (is_a ? (uint32_t)tid * 16u : 4096u + (uint32_t)(tid - 256) * 16u);
            auto fetch = [&](int j, int d) {       // one cp.async group per call, also when there is nothing to load
                if (j < nj) {
                    const int kbi = klist[2 * j + grp];
                    const uint32_t dst = slot0 + (uint32_t)d * SLOT_BYTES;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                                 "l"(r0p + (size_t)kbi * 4), "r"(v0 ? 16u : 0u) : "memory");
                    if (!is_a)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 4096u),
                                     "l"(r1p + (size_t)kbi * 4), "r"(v1 ? 16u : 0u) : "memory");
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
                        const uint32_t src = slot0 + (uint32_t)d * SLOT_BYTES;
                        uint4 p0, p1 = make_uint4(0u, 0u, 0u, 0u);
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(p0.x), "=r"(p0.y), "=r"(p0.z), "=r"(p0.w) : "r"(src) : "memory");
                        if (!is_a)
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(p1.x), "=r"(p1.y), "=r"(p1.z), "=r"(p1.w) : "r"(src + 4096u) : "memory");
                        process(j, p0, p1);                  // consumes p0 / p1: the slot may be refilled now
                        fetch(j + PF, d);
                    }
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
        const int srow0 = is_a ? ra : TM + rb0, srow1 = TM + rb1;     // rows inside a staging buffer
        for (int g = 0; g < ngroups; ++g) {
            const int buf = g % NBUF;
            mbar_wait(&loaded[buf], (g / NBUF) & 1);
            // this thread's packed bits for its two K-blocks of the group: kk = grp and grp + 2
            const unsigned char *sb = staging + (size_t)buf * BUF_BYTES;
            uint4 p0[2], p1[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = grp + 2 * h;
                p0[h] = *reinterpret_cast<const uint4 *>(sb + (kk * ROWS + srow0) * 16);
                p1[h] = p0[h];
                if (!is_a) p1[h] = *reinterpret_cast<const uint4 *>(sb + (kk * ROWS + srow1) * 16);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&consumed[buf]);     // the loader may refill this buffer
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kb = g * GK + grp + 2 * h;
                if (kb >= nkb) break;
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
                if (is_a) {
                    const uint32_t pw[4] = {p0[h].x, p0[h].y, p0[h].z, p0[h].w};
                    uint32_t o[32];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t t[8];
                        expand32(pw[q], t);
#pragma unroll
                        for (int gg = 0; gg < 8; ++gg) o[q * 8 + gg] = t[gg];
                    }
                    tc_st32(a_lane + (uint32_t)(s * A_COLS), o);      // includes tcgen05.wait::st
                    tc_fence_before();
                } else {
                    unsigned char *stg = stages + (size_t)s * B_BYTES;
                    expand_row_to_smem(p0[h], stg, ch0);
                    expand_row_to_smem(p1[h], stg, ch1);
                    fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
        }
    } else if (warp == LOAD_WARP) {
        // ------------------------------------------------------------------ loader
        // lane l streams rows l, l + 32, ...; staging layout [kk][row][16 B] keeps both the LDGSTS
        // writes and the expanders' LDS.128 reads conflict-free.  Rows past n and K-blocks past the
        // end are zero-filled (src-size 0).
        const uint32_t *img_base = packed + (size_t)img * n * words;
        int bj = -1;                       // bitmap word being scanned and its bits not yet taken
        uint32_t bm = 0;
        for (int g = 0; g < (K::DIRECT ? 0 : ngroups); ++g) {
            const int buf = g % NBUF;
            if (g >= NBUF) mbar_wait(&consumed[buf], ((g / NBUF) - 1) & 1);
            unsigned char *dst = staging + (size_t)buf * BUF_BYTES;
            int kbs[GK];                   // the next GK visited K-blocks (warp-uniform)
#pragma unroll
            for (int kk = 0; kk < GK; ++kk) {
                kbs[kk] = 0;
                if (g * GK + kk < nkb) {
                    while (bm == 0u) { ++bj; bm = bj < KMAP_WORDS ? kmap[bj] : (__ldg(ua + bj) & __ldg(ub + bj)); }
                    kbs[kk] = 32 * bj + __ffs(bm) - 1;
                    bm &= bm - 1;
                }
            }
#pragma unroll 4
            for (int t = lane; t < ROWS; t += 32) {
                const int grow = t < TM ? row0 + t : col0 + (t - TM);
                const bool rv = grow < n;
                const uint32_t *rsrc = img_base + (size_t)(rv ? __ldg(perm + grow) : 0) * words;
#pragma unroll
                for (int kk = 0; kk < GK; ++kk) {
                    const int kb = g * GK + kk;
                    const bool ok = rv && kb < nkb;
                    const uint32_t nbytes = ok ? 16u : 0u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                                     smem_u32(dst + (kk * ROWS + t) * 16)),
                                 "l"(rsrc + (ok ? kbs[kk] * 4 : 0)), "r"(nbytes)
                                 : "memory");
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&loaded[buf]))
                         : "memory");
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer
        // ONE thread runs the whole loop (waits included): with the loop around an `if (lane == 0)` the compiler
        // wraps every tcgen05 instruction into ELECT / R2UR / vote sequences and the ~100 dependent instructions
        // per K-block of this single warp were what bounded the kernel (ncu: the issuer never waits on `full`,
        // the 16 expander warps wait on `empty` 40 % of all samples, tensor pipe 50 %).  Stage index and phase
        // are compile-time / incremental, descriptors are one add away from a per-tile base.
        if (elect_one() && nkb > 0) {        // elect.sync, not lane == 0: see common.cuh
            const uint64_t bd0 = smem_desc(smem_u32(stages));
            const uint32_t a0 = tmem_base + TMEM_A0;
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
                        tc_mma_i8_ts(tmem_base, a_t, bd, kb != 0);
#pragma unroll
                        for (int k = 1; k < KB / 32; ++k)       // K = 32 bytes: 8 TMEM columns of A, +2 (x16 B) of B
                            tc_mma_i8_ts(tmem_base, a_t + 8 * k, bd + 2 * k, 1u);
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
    // The 16 expander warps drain the accumulator: warp w reads TMEM lanes 32 (w % 4).. and, per pass of
    // 128 columns, the 32-column slice w / 4.  Areas are cached in smem; the fp16 results (iou, asy and
    // the mirror's asy) go to smem tiles first so that every global store is a contiguous segment:
    // a row of the tile (direct block) or a column of it (mirror block, written transposed).
    constexpr int EP = 130;                                       // half pitch (65 words: odd -> conflict-free)
    int *areaR = reinterpret_cast<int *>(stages);                 // [128]
    int *areaC = areaR + TM;                                      // [256]
    __half *s_iou = reinterpret_cast<__half *>(areaC + TN);       // [128][EP]
    __half *s_asy = s_iou + TM * EP;                              // [128][EP]  inter / area_col
    __half *s_asyT = s_asy + TM * EP;                             // [128][EP]  inter / area_row (mirror block)
    const int32_t *area = area_all + (size_t)img * n;
    int32_t *inter = inter_all ? inter_all + (size_t)img * n * n : nullptr;
    __half *iou = iou_all + (size_t)img * n * n;
    __half *asy = asy_all + (size_t)img * n * n;
    if (warp < NEXP) {
        if (nkb > 0) mbar_wait(accum_full, 0);                   // empty K-range: nothing was accumulated
        tc_fence_after();
        for (int i = tid; i < TM + TN; i += NEXP * 32) {
            const int gidx = i < TM ? row0 + i : col0 + (i - TM);
            areaR[i] = gidx < n ? area[perm[gidx]] : 0;           // areaC follows areaR in memory
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NEXP * 32) : "memory");
        const int q4 = warp & 3, cs = warp >> 2;
        const int rl = 32 * q4 + lane;                            // TMEM lane = tile row
        const int r = row0 + rl;
        const int a_r = areaR[rl];
        const bool vec_ok = (n & 3) == 0;
#pragma unroll 1
        for (int pass = 0; pass < TN / 128; ++pass) {
            const int cl0 = pass * 128 + cs * 32;                 // first tile column of this warp's slice
            int v[32];
            if (nkb > 0) {
                tc_ld32(tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)cl0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int a_c = areaC[cl0 + j];
                const float fi = (float)v[j];
                const int so = rl * EP + cs * 32 + j;
                s_iou[so] = __float2half_rn(__fdiv_rn(fi, (float)(a_r + a_c - v[j])));
                s_asy[so] = __float2half_rn(__fdiv_rn(fi, (float)a_c));
                s_asyT[so] = __float2half_rn(__fdiv_rn(fi, (float)a_r));
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
            asm volatile("bar.sync 1, %0;" ::"n"(NEXP * 32) : "memory");
            // direct block: warp w writes tile rows w, w + 16, ...; a lane covers 4 consecutive columns
            for (int rr = warp; rr < TM; rr += NEXP) {
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
                for (int cc = warp; cc < 128; cc += NEXP) {
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
            asm volatile("bar.sync 1, %0;" ::"n"(NEXP * 32) : "memory");
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

using CfgDefault = Cfg<4>;

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

// true when the tensor-core path can take this problem (else the popcount kernel runs)
bool cim_mask_overlap_tc_eligible(int n, long long words) {
    return n >= 64 && words >= 4 && (words % 4) == 0 && (size_t)cim_max_smem_optin() >= CfgDefault::SMEM_BYTES;
}

int cim_mask_overlap_tc_launch(const uint32_t *packed, const int32_t *area, const int32_t *perm,
                               const uint32_t *umap_a, const uint32_t *umap_b, int bw, unsigned long long *visited,
                               const int32_t *tile_order, int n_img, int n, long long words, int32_t *inter,
                               __half *iou, __half *asy, cudaStream_t st) {
    // The direct kernel needs the tile's K-block list to fit its smem array (masks up to 512 Kpixel); larger masks
    // take the loader-warp kernel.  CIM_DBG_OVERLAP_LOADER_WARP forces the loader-warp kernel (tests, A/B timing).
#define CIM_OV_ARGS packed, area, perm, umap_a, umap_b, bw, visited, tile_order, n_img, n, words, inter, iou, asy, st
    const bool direct_ok = (words + 3) / 4 <= KLIST;
    if (!direct_ok || (cim_get_debug_flags() & CIM_DBG_OVERLAP_LOADER_WARP)) return launch<CfgDefault>(CIM_OV_ARGS);
    return launch<Cfg<4, 2, 4>>(CIM_OV_ARGS);
#undef CIM_OV_ARGS
}
