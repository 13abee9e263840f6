// mask_overlap.cu -- pairwise mask-proposal IoU and containment maps for sm_100a.
//
// Replaces lib/utils/mask_utils.py:6-18 (mask_iou) and :20-32 (mask_asymmetric_iou) of the
// reference as driven by tools/pre/create_cob_iou.py:43-48 / create_cob_asy_iou.py:43-51
// (N^2 pairs, two cupy reductions each, float16 cast, pickled, re-loaded every training step
// at lib/modeling/model_builder.py:148-156).
//
//   inter[i,j] = |m_i & m_j|         exact integer
//   iou[i,j]   = fp16( fp32(inter) / fp32(area_i + area_j - inter) )
//   asy[i,j]   = fp16( fp32(inter) / fp32(area_j) )                  0/0 -> NaN like numpy
//
// The reference divides in float64 and stores into float32 before the float16 cast; for integer
// operands below 2^24 the float64 quotient can never sit close enough to a float32 rounding
// boundary for the double rounding to matter, so one IEEE fp32 division gives the same float32
// (tests/test_oracle_masks.py checks this exhaustively for small counts).  The fp32 -> fp16 step
// is kept as a second rounding exactly as in the reference.
//
// Kernels in this file:
//   mask_pack_kernel      byte masks -> bit masks (32 pixels per word), HBM-bound streaming
//   mask_area_kernel      per-mask popcount
//   mask_overlap_popc     64x64 output tile per CTA over bit-packed rows: AND + POPC, only the
//                         upper triangle of tiles is computed, the mirror tile is written from
//                         the same counts (inter is symmetric, asy is not)
#include "common.cuh"

// mask_overlap_tc.cu: the tcgen05 (tensor core) path for large problems
bool cim_mask_overlap_tc_eligible(int n, long long words);
int cim_mask_overlap_tc_launch(const uint32_t *packed, const int32_t *area, const int32_t *perm, const int2 *range_a,
                               const int2 *range_b, int n_img, int n, long long words, int32_t *inter, __half *iou,
                               __half *asy, cudaStream_t st);

namespace {

// ---------------------------------------------------------------------------------------- pack
__device__ __forceinline__ uint32_t nibble_of(uint32_t four_bytes) {
    // 0xFF per non-zero byte, keep bit i of byte i, add the bytes up in the top byte
    uint32_t m = __vcmpne4(four_bytes, 0u) & 0x08040201u;
    return (m * 0x01010101u) >> 24;
}

__global__ void mask_pack_kernel(const uint8_t *__restrict__ masks, uint32_t *__restrict__ packed,
                                 long long n_masks, long long hw, long long words) {
    const long long total = n_masks * words;
    const bool vec = (hw % 16) == 0 && (((uintptr_t)masks) % 16) == 0;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long m = idx / words, w = idx - m * words;
        const long long p0 = w * 32;
        const uint8_t *src = masks + m * hw + p0;
        uint32_t bits = 0;
        if (vec && p0 + 32 <= hw) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src));
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
            bits = nibble_of(a.x) | (nibble_of(a.y) << 4) | (nibble_of(a.z) << 8) | (nibble_of(a.w) << 12) |
                   (nibble_of(b.x) << 16) | (nibble_of(b.y) << 20) | (nibble_of(b.z) << 24) |
                   (nibble_of(b.w) << 28);
        } else {
            for (int i = 0; i < 32; ++i)
                if (p0 + i < hw && src[i] != 0) bits |= 1u << i;
        }
        packed[idx] = bits;
    }
}

// ------------------------------------------------------------------------------- unpack crops
// Wire format for proposal masks (host -> device): only the bounding box of every mask is sent.
// crop row r, word k holds pixels (y0 + r, 32 * wx0 + 32 k .. + 31); it is OR-ed into the flat
// bit-packed mask (pixel p = y * W + x -> bit p & 31 of word p >> 5) with a funnel shift when the
// image width is not a multiple of 32.  One CTA per mask; `packed` is zeroed beforehand.
__global__ void mask_unpack_crops_kernel(const uint32_t *__restrict__ crop_words, const int32_t *__restrict__ meta,
                                         const long long *__restrict__ off, uint32_t *__restrict__ packed,
                                         int H, int W, long long words) {
    const long long m = blockIdx.x;
    const int wx0 = meta[4 * m], y0 = meta[4 * m + 1], ww = meta[4 * m + 2], h = meta[4 * m + 3];
    const uint32_t *src = crop_words + off[m];
    uint32_t *dst = packed + m * words;
    const bool aligned = (W & 31) == 0;
    for (int e = threadIdx.x; e < ww * h; e += blockDim.x) {
        const int r = e / ww, k = e - r * ww;
        const int y = y0 + r, x = 32 * (wx0 + k);
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        uint32_t w = __ldg(src + e);
        if (x + 32 > W) w &= (1u << (W - x)) - 1u;             // bits beyond the image row are dropped
        if (w == 0u) continue;
        const long long p = (long long)y * W + x;
        const long long wi = p >> 5;
        const int sh = (int)(p & 31);
        if (aligned) {
            dst[wi] = w;                                        // every output word has one source word
        } else {
            atomicOr(dst + wi, w << sh);
            if (sh != 0 && (w >> (32 - sh)) != 0u) atomicOr(dst + wi + 1, w >> (32 - sh));
        }
    }
}

// ---------------------------------------------------------------------------------------- area
// per mask (one warp): popcount and the range [lo, hi) of words that hold any set bit
__global__ void mask_area_kernel(const uint32_t *__restrict__ packed, int32_t *__restrict__ area,
                                 int2 *__restrict__ krange, long long n_masks, long long words) {
    const long long m = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= n_masks) return;
    const int lane = threadIdx.x & 31;
    const uint32_t *row = packed + m * words;
    int s = 0, lo = 0x7fffffff, hi = 0;
    for (long long w = lane; w < words; w += 32) {
        const uint32_t v = __ldg(row + w);
        s += __popc(v);
        if (v) { lo = min(lo, (int)w); hi = max(hi, (int)w + 1); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
        area[m] = s;
        if (krange) krange[m] = hi ? make_int2(lo, hi) : make_int2(0, 0);
    }
}

// ------------------------------------------------------------------------- locality sort + ranges
// One CTA per image.  Masks are sorted by the centre of their non-zero word range (= vertical
// position, pixels are row-major), so the 128 / 256 masks of a tile block share a narrow range.
// perm[k] = original index of the k-th mask in sorted order, inv = inverse.  For every block of
// 128 (A operand) and 256 (B operand) sorted masks the union range is stored in K-blocks of 4 words;
// a tile only has to visit the intersection of its two ranges: everywhere else one operand is all
// zero.  (tools/ estimate: 0.36-0.43 of the K-blocks on the synthetic proposals.)
__global__ void __launch_bounds__(1024)
mask_sort_kernel(const int2 *__restrict__ krange_all, int n, int npad, int32_t *__restrict__ perm_all,
                 int32_t *__restrict__ inv_all, int2 *__restrict__ range_a, int2 *__restrict__ range_b, int nrb,
                 int ncb) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    int *key = reinterpret_cast<int *>(sort_smem);        // [npad]
    int *idx = key + npad;                                // [npad]
    const int img = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int2 *kr = krange_all + (size_t)img * n;
    for (int i = tid; i < npad; i += nthr) {
        key[i] = i < n ? kr[i].x + kr[i].y : 0x7fffffff;
        idx[i] = i < n ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int size = 2; size <= npad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (npad >> 1); t += nthr) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const int ka = key[lo], kb = key[hi], ia = idx[lo], ib = idx[hi];
                const bool a_first = ka < kb || (ka == kb && ia < ib);
                if (asc ? !a_first : a_first) { key[lo] = kb; key[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
            }
            __syncthreads();
        }
    int32_t *perm = perm_all + (size_t)img * n, *inv = inv_all + (size_t)img * n;
    for (int i = tid; i < n; i += nthr) { perm[i] = idx[i]; inv[idx[i]] = i; }
    // block ranges (in K-blocks of 4 words); empty masks (hi == 0) do not count
    for (int blk = tid; blk < nrb + ncb; blk += nthr) {
        const bool is_a = blk < nrb;
        const int bs = is_a ? 128 : 256, b0 = (is_a ? blk : blk - nrb) * bs;
        int lo = 0x7fffffff, hi = 0;
        for (int i = b0; i < min(n, b0 + bs); ++i) {
            const int2 r = kr[idx[i]];
            if (r.y) { lo = min(lo, r.x); hi = max(hi, r.y); }
        }
        const int2 out = hi ? make_int2(lo >> 2, (hi + 3) >> 2) : make_int2(0, 0);
        if (is_a) range_a[(size_t)img * nrb + blk] = out;
        else range_b[(size_t)img * ncb + (blk - nrb)] = out;
    }
}

// out[i][j] = tmp[inv[i]][inv[j]]: one CTA per output row, the source row goes through smem
template <typename T>
__global__ void __launch_bounds__(256)
mask_unpermute_kernel(const T *__restrict__ tmp_all, const int32_t *__restrict__ inv_all, T *__restrict__ out_all,
                      int n) {
    extern __shared__ __align__(16) unsigned char unp_smem[];
    T *row = reinterpret_cast<T *>(unp_smem);
    const int i = blockIdx.x, img = blockIdx.y;
    const int32_t *inv = inv_all + (size_t)img * n;
    const T *src = tmp_all + ((size_t)img * n + inv[i]) * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) row[j] = src[j];
    __syncthreads();
    T *dst = out_all + ((size_t)img * n + i) * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = row[inv[j]];
}

// --------------------------------------------------------------------------------- popc overlap
constexpr int TS = 64;      // output tile side
constexpr int KC = 32;      // words per k-chunk
constexpr int KP = KC + 1;  // padded smem pitch

__device__ __forceinline__ void write_pair(int32_t *inter, __half *iou, __half *asy, size_t idx, int I,
                                           int a_row, int a_col) {
    const float fi = (float)I;
    if (inter) inter[idx] = I;
    iou[idx] = __float2half_rn(__fdiv_rn(fi, (float)(a_row + a_col - I)));
    asy[idx] = __float2half_rn(__fdiv_rn(fi, (float)a_col));
}

__global__ void __launch_bounds__(256)
mask_overlap_popc_kernel(const uint32_t *__restrict__ packed, const int32_t *__restrict__ area_all,
                         int n, long long words, int32_t *__restrict__ inter_all, __half *__restrict__ iou_all,
                         __half *__restrict__ asy_all) {
    __shared__ uint32_t As[TS][KP];
    __shared__ uint32_t Bs[TS][KP];
    const int img = blockIdx.y;
    // linear id -> (ti, tj) with ti <= tj
    const int nt = (n + TS - 1) / TS;
    int ti = 0, rem = blockIdx.x;
    while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
    const int tj = ti + rem;

    const uint32_t *base = packed + (size_t)img * n * words;
    const int32_t *area = area_all + (size_t)img * n;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0;

    for (long long k0 = 0; k0 < words; k0 += KC) {
        // 64 rows x 32 words per operand, 256 threads: 8 words each, coalesced along k
#pragma unroll
        for (int it = 0; it < (TS * KC) / 256; ++it) {
            const int e = it * 256 + tid, row = e / KC, kk = e % KC;
            const long long kw = k0 + kk;
            const int ra = ti * TS + row, rb = tj * TS + row;
            As[row][kk] = (ra < n && kw < words) ? __ldg(base + (size_t)ra * words + kw) : 0u;
            Bs[row][kk] = (rb < n && kw < words) ? __ldg(base + (size_t)rb * words + kw) : 0u;
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

    int32_t *inter = inter_all ? inter_all + (size_t)img * n * n : nullptr;
    __half *iou = iou_all + (size_t)img * n * n;
    __half *asy = asy_all + (size_t)img * n * n;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = ti * TS + ty + 16 * r;
        if (i >= n) continue;
        const int ai = area[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = tj * TS + tx + 16 * c;
            if (j >= n) continue;
            const int aj = area[j];
            write_pair(inter, iou, asy, (size_t)i * n + j, acc[r][c], ai, aj);
            if (ti != tj) write_pair(inter, iou, asy, (size_t)j * n + i, acc[r][c], aj, ai);
        }
    }
}

}  // namespace

CIM_API int cim_mask_pack(const uint8_t *masks, uint32_t *packed, int64_t n_masks, int64_t hw, int64_t words,
                          cim_stream_t stream) {
    if (!masks || !packed || n_masks < 0 || hw <= 0 || words * 32 < hw) return CIM_ERR_ARG;
    if (n_masks == 0) return CIM_OK;
    const long long total = n_masks * words;
    const int blocks = (int)min((long long)cim_num_sms() * 16, (total + 255) / 256);
    mask_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(masks, packed, n_masks, hw, words);
    return cim_launch_status();
}

CIM_API int cim_mask_unpack_crops(const uint32_t *crop_words, const int32_t *crop_meta, const int64_t *crop_off,
                                  uint32_t *packed, int64_t n_masks, int H, int W, int64_t words,
                                  cim_stream_t stream) {
    if (!crop_words || !crop_meta || !crop_off || !packed || n_masks < 0 || H <= 0 || W <= 0) return CIM_ERR_ARG;
    if (words * 32 < (int64_t)H * W) return CIM_ERR_ARG;
    if (n_masks == 0) return CIM_OK;
    if (n_masks > 0x7fffffffLL) return CIM_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(packed, 0, sizeof(uint32_t) * (size_t)n_masks * words, st);
    mask_unpack_crops_kernel<<<(unsigned)n_masks, 128, 0, st>>>(crop_words, crop_meta,
                                                                reinterpret_cast<const long long *>(crop_off), packed,
                                                                H, W, words);
    return cim_launch_status();
}

namespace {
struct OverlapWs {
    int32_t *area;
    int2 *krange, *range_a, *range_b;
    int32_t *perm, *inv, *tmp_inter;
    __half *tmp_iou, *tmp_asy;
    size_t bytes;
};
inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
OverlapWs carve_overlap_ws(void *base, int n_img, int n, int want_inter) {
    OverlapWs w{};
    const size_t nm = (size_t)(n_img > 0 ? n_img : 0) * (size_t)(n > 0 ? n : 0), nn = nm * (size_t)(n > 0 ? n : 0);
    const size_t nrb = (size_t)n_img * ((n + 127) / 128), ncb = (size_t)n_img * ((n + 255) / 256);
    char *p = (char *)base;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *q = p ? p + o : nullptr; o += up256(bytes); return q; };
    w.area = (int32_t *)take(nm * 4);
    w.krange = (int2 *)take(nm * 8);
    w.perm = (int32_t *)take(nm * 4);
    w.inv = (int32_t *)take(nm * 4);
    w.range_a = (int2 *)take(nrb * 8);
    w.range_b = (int2 *)take(ncb * 8);
    w.tmp_iou = (__half *)take(nn * 2);
    w.tmp_asy = (__half *)take(nn * 2);
    w.tmp_inter = want_inter ? (int32_t *)take(nn * 4) : nullptr;
    w.bytes = o + 256;
    return w;
}
}  // namespace

CIM_API size_t cim_mask_overlap_workspace_bytes(int n_img, int n, int64_t words, int want_inter) {
    (void)words;
    return carve_overlap_ws(nullptr, n_img, n, want_inter).bytes;
}

CIM_API int cim_mask_overlap(const uint32_t *packed, int n_img, int n, int64_t words, int32_t *inter,
                             int32_t *area, void *iou_f16, void *asy_f16, void *workspace, size_t ws_bytes,
                             cim_stream_t stream) {
    return cim_mask_overlap_algo(packed, n_img, n, words, inter, area, iou_f16, asy_f16, workspace, ws_bytes,
                                 CIM_OVERLAP_AUTO, stream);
}

CIM_API int cim_mask_overlap_algo(const uint32_t *packed, int n_img, int n, int64_t words, int32_t *inter,
                                  int32_t *area, void *iou_f16, void *asy_f16, void *workspace, size_t ws_bytes,
                                  int algo, cim_stream_t stream) {
    if (algo != CIM_OVERLAP_AUTO && algo != CIM_OVERLAP_POPC && algo != CIM_OVERLAP_TENSOR) return CIM_ERR_ARG;
    if (!packed || !iou_f16 || !asy_f16 || n_img < 0 || n < 0 || words <= 0) return CIM_ERR_ARG;
    if (words * 32 >= (1LL << 24)) return CIM_ERR_SHAPE;      // counts must stay exact in fp32
    if (n_img == 0 || n == 0) return CIM_OK;
    if (n_img > 65535 || n > 16384) return CIM_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc_ok = cim_mask_overlap_tc_eligible(n, words) && cim_aligned(packed, 16);
    if (algo == CIM_OVERLAP_TENSOR && !tc_ok) return CIM_ERR_SHAPE;
    // the tensor path pays off once a 128 x 256 tile is reasonably full and K is long
    const bool use_tc = algo == CIM_OVERLAP_TENSOR || (algo == CIM_OVERLAP_AUTO && tc_ok && n >= 256 && words >= 128);
    const size_t need = use_tc ? cim_mask_overlap_workspace_bytes(n_img, n, words, inter != nullptr)
                               : (area ? 0 : up256(sizeof(int32_t) * (size_t)n_img * n) + 256);
    if (need && (!workspace || ws_bytes < need || !cim_aligned(workspace, 256))) return CIM_ERR_WORKSPACE;
    const OverlapWs w = carve_overlap_ws(workspace, n_img, n, inter != nullptr);
    if (!area) area = w.area;
    const long long n_masks = (long long)n_img * n;
    mask_area_kernel<<<(unsigned)((n_masks + 7) / 8), 256, 0, st>>>(packed, area, use_tc ? w.krange : nullptr, n_masks,
                                                                    words);
    int rc = cim_launch_status();
    if (rc) return rc;
    __half *iou = reinterpret_cast<__half *>(iou_f16), *asy = reinterpret_cast<__half *>(asy_f16);
    if (use_tc) {
        int npad = 1;
        while (npad < n) npad <<= 1;
        const int nrb = (n + 127) / 128, ncb = (n + 255) / 256;
        const size_t smem_sort = (size_t)npad * 8;
        cudaFuncSetAttribute(mask_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort);
        mask_sort_kernel<<<n_img, 1024, smem_sort, st>>>(w.krange, n, npad, w.perm, w.inv, w.range_a, w.range_b, nrb,
                                                         ncb);
        if ((rc = cim_launch_status())) return rc;
        // the tensor kernel works in sorted index space and writes sorted-order maps
        rc = cim_mask_overlap_tc_launch(packed, area, w.perm, w.range_a, w.range_b, n_img, n, words, w.tmp_inter,
                                        w.tmp_iou, w.tmp_asy, st);
        if (rc) return rc;
        dim3 g((unsigned)n, (unsigned)n_img);
        mask_unpermute_kernel<__half><<<g, 256, (size_t)n * 2, st>>>(w.tmp_iou, w.inv, iou, n);
        mask_unpermute_kernel<__half><<<g, 256, (size_t)n * 2, st>>>(w.tmp_asy, w.inv, asy, n);
        if (inter) {
            cudaFuncSetAttribute(mask_unpermute_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4);
            mask_unpermute_kernel<int32_t><<<g, 256, (size_t)n * 4, st>>>(w.tmp_inter, w.inv, inter, n);
        }
        return cim_launch_status();
    }
    const int nt = (n + TS - 1) / TS;
    dim3 grid((unsigned)(nt * (nt + 1) / 2), (unsigned)n_img);
    mask_overlap_popc_kernel<<<grid, 256, 0, st>>>(packed, area, n, words, inter, iou, asy);
    return cim_launch_status();
}
