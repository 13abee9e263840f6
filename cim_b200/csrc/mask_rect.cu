// mask_rect.cu -- rectangular mask-overlap ratios: the offline callers of lib/utils/mask_utils.py.
//
// The training path needs the square N x N maps of one proposal set (mask_overlap*.cu).  The reference's offline
// tools call the same functions on two DIFFERENT sets, usually N proposals x 1 averaged peak mask:
//   tools/pre/AGPL_label_assign.py:84,165, tools/pre/point_level_label_assign.py:79   mask_iou(N, 1)
//   tools/generate_mask_for_MaskRCNN.py                                                mask_iou / mask_inside / mask_outside
//   tools/pre/create_cob_iou.py:45, create_cob_asy_iou.py:46                            (N, 1) column by column
// Ratios (lib/utils/mask_utils.py), all float32 = float32(int / int) like the reference's float32 result array:
//   mode 0  mask_iou            I / |a or b|                         (:6-18)
//   mode 1  mask_asymmetric_iou I / mask_b.sum()  -- the sum over ALL masks of b, as the reference writes it (:27)
//   mode 2  mask_inside         I / |b_k|                            (:35-47)
//   mode 3  mask_outside        I / |a_n|                            (:50-62)
// Bit-packed operands, AND + POPC, 64 x 64 output tile per CTA (HBM/L2-bound integer work; no tensor cores: with one
// or a few columns there is no contraction worth the name).
#include "common.cuh"

namespace {

constexpr int TS = 64, KC = 32, KP = KC + 1;

__global__ void __launch_bounds__(256)
mask_row_popc_kernel(const uint32_t *__restrict__ packed, long long words, int32_t *__restrict__ area) {
    __shared__ int red[8];
    const uint32_t *row = packed + (size_t)blockIdx.x * words;
    int s = 0;
    for (long long k = threadIdx.x; k < words; k += 256) s += __popc(__ldg(row + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += red[i];
        area[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
mask_pair_ratio_kernel(const uint32_t *__restrict__ pa, const uint32_t *__restrict__ pb, int na, int nb, long long words,
                       int mode, const int32_t *__restrict__ area_a, const int32_t *__restrict__ area_b,
                       float *__restrict__ ratio, int32_t *__restrict__ inter) {
    __shared__ uint32_t As[TS][KP];
    __shared__ uint32_t Bs[TS][KP];
    __shared__ long long s_total_b;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
    if (mode == 1 && tid == 0) {                        // mask_b.sum() over the whole second set
        long long t = 0;
        for (int j = 0; j < nb; ++j) t += area_b[j];
        s_total_b = t;
    }
    int acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0;
    for (long long k0 = 0; k0 < words; k0 += KC) {
#pragma unroll
        for (int it = 0; it < (TS * KC) / 256; ++it) {
            const int e = it * 256 + tid, row = e / KC, kk = e % KC;
            const long long kw = k0 + kk;
            const int ra = i0 + row, rb = j0 + row;
            As[row][kk] = (ra < na && kw < words) ? __ldg(pa + (size_t)ra * words + kw) : 0u;
            Bs[row][kk] = (rb < nb && kw < words) ? __ldg(pb + (size_t)rb * words + kw) : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < KC; ++kk) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[ty + 16 * r][kk];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = Bs[tx + 16 * c][kk];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] += __popc(a[r] & b[c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 16 * r;
        if (i >= na) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx + 16 * c;
            if (j >= nb) continue;
            const int I = acc[r][c];
            float den;
            if (mode == 0) den = (float)(area_a[i] + area_b[j] - I);
            else if (mode == 1) den = (float)s_total_b;
            else if (mode == 2) den = (float)area_b[j];
            else den = (float)area_a[i];
            const size_t o = (size_t)i * nb + j;
            ratio[o] = __fdiv_rn((float)I, den);          // 0 / 0 -> NaN, as numpy
            if (inter) inter[o] = I;
        }
    }
}

}  // namespace

CIM_API int cim_mask_pair_ratio(const uint32_t *packed_a, const uint32_t *packed_b, int na, int nb, int64_t words,
                                int mode, float *ratio, int32_t *inter, int32_t *area_a, int32_t *area_b,
                                cim_stream_t stream) {
    if (!packed_a || !packed_b || !ratio || !area_a || !area_b) return CIM_ERR_ARG;
    if (na < 0 || nb < 0 || words <= 0 || mode < 0 || mode > 3) return CIM_ERR_ARG;
    if (na == 0 || nb == 0) return CIM_OK;
    if (words * 32 > (1LL << 24)) return CIM_ERR_SHAPE;         // counts must stay exact in float32
    cudaStream_t st = (cudaStream_t)stream;
    mask_row_popc_kernel<<<(unsigned)na, 256, 0, st>>>(packed_a, words, area_a);
    mask_row_popc_kernel<<<(unsigned)nb, 256, 0, st>>>(packed_b, words, area_b);
    int rc = cim_launch_status();
    if (rc) return rc;
    dim3 grid((unsigned)((nb + TS - 1) / TS), (unsigned)((na + TS - 1) / TS));
    if (grid.y > 65535) return CIM_ERR_SHAPE;
    mask_pair_ratio_kernel<<<grid, 256, 0, st>>>(packed_a, packed_b, na, nb, words, mode, area_a, area_b, ratio, inter);
    return cim_launch_status();
}
