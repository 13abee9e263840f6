// score_heads_tc.cu -- the scoring GEMM of cls_iou_model on the 5th-gen tensor cores (sm_100a).
//
// logits[m][n] = sum_k x[m][k] * W[n][k] + b[n]  with  M = n_img * R (16000), K = 4096 and
// N = (2 + 2K_ref) * (C + 1) = 168 (VOC): 22 GFLOP per step that the fp32 FFMA kernel
// (score_heads.cu) needs 0.83 ms for.  Parity demands fp32-like accuracy (1e-5 relative), which a
// plain TF32 GEMM misses, so every operand is split into two TF32 numbers
//     x = x_hi + x_lo,  x_hi = x with the low 13 mantissa bits cleared,  x_lo = x - x_hi (exact)
// and   x * w  ~=  x_hi w_hi + x_hi w_lo + x_lo w_hi      (the dropped x_lo w_lo term is ~2^-22),
// three tcgen05.mma.kind::tf32 products accumulated in fp32 in tensor memory.
//
// One CTA per 128 rows of x and per group of heads (all 8 heads at once when 8 (C+1) <= 256).
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) of the x tile [128 x 32 fp32] and
//               of the W_hi / W_lo tiles [NT x 32] per k-block; mbarrier complete_tx.
//   warps 2..5  split: read the raw x tile, write x_hi in place and x_lo next to it (generic
//               proxy -> fence.proxy.async); later the epilogue (TMEM lane quarter = warp % 4).
//   warp 1      MMA issuer: 3 products x 4 k-steps per k-block, tcgen05.commit frees the stage.
// Epilogue: tcgen05.ld per head, bias, softmax over classes / sigmoid in registers, one contiguous
// store per (head, row).  The detector head keeps its logits (softmax over the proposals of an
// image runs in score_col_softmax_kernel).
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int BM = 128;                  // rows of x per CTA (UMMA M)
constexpr int BK = 32;                   // fp32 elements per k-block = 128 B = one SW128 atom row
constexpr int NSTAGE = 2;
constexpr int A_TILE = BM * BK * 4;      // 16 KB
constexpr int THREADS_TC = 6 * 32;

__device__ __forceinline__ uint64_t smem_desc128(uint32_t saddr) {       // K-major SWIZZLE_128B, SBO 1024 B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__device__ __forceinline__ void mbar_arrive1(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// W -> (W_hi, W_lo), both exactly representable... W_hi in TF32, W_lo = W - W_hi (the MMA drops its
// low mantissa bits, an error of 2^-21 |W|)
__global__ void score_split_w_kernel(const float *__restrict__ w, float *__restrict__ w_hi, float *__restrict__ w_lo,
                                     long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = w[i], h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        w_hi[i] = h;
        w_lo[i] = v - h;
    }
}

// NCH = ceil(C1 / 32): 32-column TMEM loads per head
template <int NCH>
__global__ void __launch_bounds__(THREADS_TC, 1)
score_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_whi,
                     const __grid_constant__ CUtensorMap tm_wlo, const float *__restrict__ bias,
                     float *__restrict__ scores, int M, int D, int C1, int n_ref, int heads_per_tile, int NT) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int W_TILE = NT * BK * 4;
    const int STAGE = 2 * A_TILE + 2 * W_TILE;                 // x_hi | x_lo | W_hi | W_lo
    uint64_t *tma_full = reinterpret_cast<uint64_t *>(smem + (size_t)NSTAGE * STAGE);
    uint64_t *lo_ready = tma_full + NSTAGE;
    uint64_t *empty = lo_ready + NSTAGE;
    uint64_t *accum_full = empty + NSTAGE;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);
    float *s_bias = reinterpret_cast<float *>(tmem_slot + 2);   // [heads_per_tile * C1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * BM;
    const int h0 = blockIdx.y * heads_per_tile;                 // first head of this CTA
    const int nheads = 2 + 2 * n_ref;
    const int nh = min(heads_per_tile, nheads - h0);
    const int n0 = h0 * C1;                                     // first row of W / first logit column
    const int nkb = D / BK;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&tma_full[s], 1); mbar_init(&lo_ready[s], 4); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < nh * C1; i += THREADS_TC) s_bias[i] = bias[n0 + i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // instruction descriptor: F32 accumulate, TF32 x TF32, K-major, N = NT, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NSTAGE;
                if (kb >= NSTAGE) mbar_wait(&empty[s], ((kb / NSTAGE) - 1) & 1);
                unsigned char *st = smem + (size_t)s * STAGE;
                mbar_expect_tx(&tma_full[s], (uint32_t)(A_TILE + 2 * W_TILE));
                tma_load_2d(st, &tm_x, kb * BK, m0, &tma_full[s]);                       // raw x -> the x_hi slot
                tma_load_2d(st + 2 * A_TILE, &tm_whi, kb * BK, n0, &tma_full[s]);
                tma_load_2d(st + 2 * A_TILE + W_TILE, &tm_wlo, kb * BK, n0, &tma_full[s]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % NSTAGE;
            mbar_wait(&tma_full[s], (kb / NSTAGE) & 1);
            mbar_wait(&lo_ready[s], (kb / NSTAGE) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t base = smem_u32(smem + (size_t)s * STAGE);
                const uint64_t xh = smem_desc128(base), xl = smem_desc128(base + A_TILE);
                const uint64_t wh = smem_desc128(base + 2 * A_TILE), wl = smem_desc128(base + 2 * A_TILE + W_TILE);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {        // K = 8 tf32 = 32 B per instruction: +2 in 16 B units
                    mma_tf32(tmem_base, xh + 2 * k, wh + 2 * k, idesc, (kb | k) != 0);
                    mma_tf32(tmem_base, xh + 2 * k, wl + 2 * k, idesc, 1u);
                    mma_tf32(tmem_base, xl + 2 * k, wh + 2 * k, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 smem_u32(&empty[s]))
                             : "memory");
                if (kb == nkb - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(accum_full))
                                 : "memory");
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ hi / lo split of the x tile
        const int row = tid - 64;                              // 0..127: one x row per thread
        const uint32_t row_off = (row >> 3) * 1024 + (row & 7) * 128, sw = row & 7;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % NSTAGE;
            mbar_wait(&tma_full[s], (kb / NSTAGE) & 1);
            unsigned char *xh = smem + (size_t)s * STAGE + row_off, *xl = xh + A_TILE;
#pragma unroll
            for (int c = 0; c < 8; ++c) {                      // the row's eight 16 B chunks (swizzled position)
                const uint32_t o = (uint32_t)((c ^ sw) << 4);
                const uint4 v = *reinterpret_cast<const uint4 *>(xh + o);
                const uint4 h = make_uint4(v.x & 0xFFFFE000u, v.y & 0xFFFFE000u, v.z & 0xFFFFE000u, v.w & 0xFFFFE000u);
                float4 l;
                l.x = __uint_as_float(v.x) - __uint_as_float(h.x);
                l.y = __uint_as_float(v.y) - __uint_as_float(h.y);
                l.z = __uint_as_float(v.z) - __uint_as_float(h.z);
                l.w = __uint_as_float(v.w) - __uint_as_float(h.w);
                *reinterpret_cast<uint4 *>(xh + o) = h;
                *reinterpret_cast<float4 *>(xl + o) = l;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive1(&lo_ready[s]);
        }

        // ------------------------------------------------------------------ epilogue
        mbar_wait(accum_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q4 = warp & 3;                               // TMEM lane quarter this warp may read
        const int m = m0 + 32 * q4 + lane;
        for (int hh = 0; hh < nh; ++hh) {
            const int h = h0 + hh;
            float z[NCH * 32];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)(hh * C1 + 32 * c), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) z[c * 32 + j] = v[j];
            }
            if (m >= M) continue;
            const float *bh = s_bias + hh * C1;
            float *dst = scores + ((size_t)h * M + m) * C1;
            const bool softmax = (h == 0) || (h >= 2 && h < 2 + n_ref);
            if (softmax) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < NCH * 32; ++j)
                    if (j < C1) { z[j] += bh[j]; mx = fmaxf(mx, z[j]); }
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < NCH * 32; ++j)
                    if (j < C1) { z[j] = expf(z[j] - mx); sum += z[j]; }
#pragma unroll
                for (int j = 0; j < NCH * 32; ++j)
                    if (j < C1) dst[j] = z[j] / sum;
            } else if (h == 1) {                               // detector: logits, column softmax follows
#pragma unroll
                for (int j = 0; j < NCH * 32; ++j)
                    if (j < C1) dst[j] = z[j] + bh[j];
            } else {                                           // refine_iou: sigmoid
#pragma unroll
                for (int j = 0; j < NCH * 32; ++j)
                    if (j < C1) dst[j] = 1.f / (1.f + expf(-(z[j] + bh[j])));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// row-major fp32 matrix [rows][cols], box = [box_rows][32 columns], 128 B swizzle, OOB -> 0
bool make_map(CUtensorMap *map, const float *base, long long rows, long long cols, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// can the tensor-core path take this shape?  (else score_heads.cu's FFMA kernel runs)
bool cim_score_tc_eligible(long long M, int D, int C1, int n_ref) {
    (void)n_ref;
    return M >= BM && (D % BK) == 0 && D >= BK && C1 >= 1 && C1 <= 96 && encode_tiled() != nullptr &&
           cim_max_smem_optin() >= 200 * 1024;
}
size_t cim_score_tc_workspace_bytes(int D, int C1, int n_ref) {
    return 2 * sizeof(float) * (size_t)(2 + 2 * n_ref) * C1 * D + 512;
}

// logits + row activations for every head into `scores`; the detector head (1) is left as logits
int cim_score_tc_launch(const float *x, const float *weight, const float *bias, float *scores, long long M, int D,
                        int C1, int n_ref, void *workspace, cudaStream_t st) {
    const int nheads = 2 + 2 * n_ref, N = nheads * C1;
    float *w_hi = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float *w_lo = w_hi + (size_t)N * D;
    score_split_w_kernel<<<256, 256, 0, st>>>(weight, w_hi, w_lo, (long long)N * D);
    int rc = cim_launch_status();
    if (rc) return rc;
    const int heads_per_tile = min(nheads, 256 / C1);
    const int ntiles = (nheads + heads_per_tile - 1) / heads_per_tile;
    const int NT = (heads_per_tile * C1 + 15) / 16 * 16;        // UMMA N: multiple of 16, <= 256
    CUtensorMap tx, twh, twl;
    if (!make_map(&tx, x, M, D, BM) || !make_map(&twh, w_hi, N, D, NT) || !make_map(&twl, w_lo, N, D, NT))
        return CIM_ERR_ARG;
    const size_t stage = 2 * (size_t)A_TILE + 2 * (size_t)NT * BK * 4;
    const size_t smem = 1024 + NSTAGE * stage + 256 + sizeof(float) * (size_t)heads_per_tile * C1;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)ntiles);
    const int nch = (C1 + 31) / 32;
#define LAUNCH_TC(NCH)                                                                                             \
    cudaFuncSetAttribute(score_gemm_tc_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    score_gemm_tc_kernel<NCH><<<grid, THREADS_TC, smem, st>>>(tx, twh, twl, bias, scores, (int)M, D, C1, n_ref,    \
                                                               heads_per_tile, NT)
    if (nch == 1) { LAUNCH_TC(1); } else if (nch == 2) { LAUNCH_TC(2); } else { LAUNCH_TC(3); }
#undef LAUNCH_TC
    return cim_launch_status();
}
