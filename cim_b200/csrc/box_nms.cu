// box_nms.cu -- test-time post-processing of the CIM heads on the GPU.
//
//  cim_test_scores : lib/core/test.py:130-133 + lib/modeling/model_builder.py:60-68 -- the refinement heads'
//                    (cls * iou)[:, 1:] averaged over the K heads: out[r][c] = mean_k cls_k[r][c+1] * iou_k[r][c+1]
//                    (summed in head order then divided by K, as `scores += ...; scores /= K` does).
//  cim_box_nms     : the per-class loop of lib/utils/mask_eval_utils.py:57-79 (and box_results_with_nms_and_limit):
//                    candidates = scores[:, c] > score_thresh, greedy NMS of lib/utils/cython_nms.pyx:37-87:
//                    areas (x2-x1+1)(y2-y1+1), boxes visited by descending score, a later box is suppressed when
//                    inter / (area_i + area_j - inter) >= nms_thresh, all in float32.  One CTA per class:
//                    bitonic sort of the candidates in shared memory, then rounds over a compacted list of the
//                    candidates that are still alive (64-box chunk -> hit matrix -> kept members -> applied to
//                    the later live candidates -> compaction).
//                    Equal scores are visited by descending proposal index (what a stable ascending argsort,
//                    reversed, yields; numpy's default argsort leaves the order of ties unspecified).
#include "common.cuh"

namespace {

__global__ void cim_test_scores_kernel(const float *__restrict__ scores, float *__restrict__ out, long long M, int C1,
                                       int K) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int C = C1 - 1;
    if (idx >= M * C) return;
    const long long r = idx / C;
    const int c = (int)(idx - r * C) + 1;
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
        const float cls = scores[((size_t)(2 + k) * M + r) * C1 + c];
        const float iou = scores[((size_t)(2 + K + k) * M + r) * C1 + c];
        s = __fadd_rn(s, __fmul_rn(cls, iou));
    }
    out[idx] = __fdiv_rn(s, (float)K);
}

constexpr int NMS_THREADS = 1024;
constexpr int CHUNK = 64;

// float32 arithmetic of cython_nms.pyx:76-84, no contraction
__device__ __forceinline__ float box_area(float4 b) {
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}
__device__ __forceinline__ bool nms_hit(float4 a, float aarea, float4 b, float barea, float thr) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f)), h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    // fl(inter / uni) >= thr decided without the division when inter is not within 2^-20 (relative) of thr * uni:
    // the rounding errors of the product and of the quotient are < 2^-22, so both shortcuts agree with the division
    if (thr > 0.f && uni > 1e-30f && uni < 3.0e38f) {
        const float t = thr * uni;
        if (inter > t * 1.00000095367431640625f) return true;
        if (inter < t * 0.99999904632568359375f) return false;
    }
    return __fdiv_rn(inter, uni) >= thr;
}

// One CTA per (class, image).  After the sort the kernel works on a LIVE list (positions in score order of the
// candidates no kept box has suppressed yet) and proceeds in rounds:
//   A  the first <= 64 live candidates form the chunk; all threads fill its 64 x 64 hit matrix (row i = the later
//      chunk members box i would suppress), one 64-bit word per row
//   B  one thread walks the rows in score order: a member still alive is KEPT and clears the members it hits
//   C  every later live candidate is tested against the kept members only, and the live list is compacted (order
//      preserved by a block-wide scan) for the next round
// A chunk member is alive w.r.t. all earlier kept boxes by construction, so it is either kept or suppressed by a
// kept member of its own chunk: the number of rounds is ~ (kept + suppressed-inside-chunks) / 64 instead of m / 64,
// and the tests are those of the sequential algorithm (box i of cython_nms.pyx:60-87 only meets boxes that are
// still alive) plus the <= 2016 pairs of each chunk.  Round 2's first kernel walked all m / 64 chunks and scanned
// 64 (mostly dead) members per later box: 2.1 ms for 640 lists of 4000 candidates; this one: see DESIGN 4.10.
__global__ void __launch_bounds__(NMS_THREADS)
cim_box_nms_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, int n, int ncls, int score_stride,
                   float score_thresh, float nms_thresh, int npad, uint8_t *__restrict__ keep) {
    extern __shared__ unsigned char smem_raw[];
    unsigned long long *key = reinterpret_cast<unsigned long long *>(smem_raw);      // [npad] (score bits, index)
    float4 *box = reinterpret_cast<float4 *>(key + npad);                            // [npad] sorted order
    uint16_t *live0 = reinterpret_cast<uint16_t *>(box + npad);                      // [npad] live list, ping
    uint16_t *live1 = live0 + npad;                                                  // [npad] pong
    __shared__ unsigned long long s_row[CHUNK];
    __shared__ unsigned long long s_kept;
    __shared__ int s_count, s_wtot[NMS_THREADS / 32];
    const int c = blockIdx.x, tid = threadIdx.x;
    // blockIdx.y = image of a batch: boxes [n_img][n][4], scores [n_img][n][score_stride], keep [n_img][ncls][n]
    boxes += (size_t)blockIdx.y * n * 4;
    scores += (size_t)blockIdx.y * n * score_stride;
    uint8_t *kout = keep + ((size_t)blockIdx.y * ncls + c) * n;

    // candidates: score > thresh (mask_eval_utils.py:64).  Key = (orderable score bits << 32) | index, sorted
    // descending; non-candidates get key 0 and sink to the end.
    int local = 0;
    for (int i = tid; i < npad; i += NMS_THREADS) {
        unsigned long long k = 0ull;
        if (i < n) {
            const float s = scores[(size_t)i * score_stride + c];
            if (s > score_thresh) {
                uint32_t b = __float_as_uint(s);
                b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);                      // total order on floats
                k = ((unsigned long long)b << 32) | (uint32_t)(i + 1);               // + 1: 0 is "none"
                ++local;
            }
            kout[i] = 0;
        }
        key[i] = k;
    }
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (local) atomicAdd(&s_count, local);
    // bitonic sort, descending; thread t owns compare-exchange t of the npad / 2 of a pass (i = t with a zero bit
    // inserted at bit log2(j), partner i | j), so every thread of a pass does useful work
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int t = tid; t < (npad >> 1); t += NMS_THREADS) {
                const int lo = t & (j - 1), i = ((t - lo) << 1) | lo, p = i | j;
                const unsigned long long a = key[i], b = key[p];
                const bool desc = (i & k) == 0;
                if (desc ? (a < b) : (a > b)) { key[i] = b; key[p] = a; }
            }
        }
    __syncthreads();
    int m = s_count;                                       // live candidates of the current round
    for (int i = tid; i < m; i += NMS_THREADS) {
        const int src = (int)(uint32_t)(key[i] & 0xFFFFFFFFull) - 1;
        box[i] = *reinterpret_cast<const float4 *>(boxes + (size_t)src * 4);
        live0[i] = (uint16_t)i;
    }
    const int lane = tid & 31, warp = tid >> 5;
    uint16_t *cur = live0, *nxt = live1;
    while (m > 0) {
        if (tid < CHUNK) s_row[tid] = 0ull;
        __syncthreads();                                   // live list / boxes of this round, s_row zeroed
        const int nc = min(m, CHUNK);
        // A: hit matrix of the chunk, 16 threads per row, 4 columns each
        {
            const int i = tid >> 4, j0 = (tid & 15) * 4;
            if (i < nc && j0 + 3 > i) {
                const float4 bi = box[cur[i]];
                const float ai = box_area(bi);
                unsigned long long bits = 0ull;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const int j = j0 + d;
                    if (j > i && j < nc) {
                        const float4 bj = box[cur[j]];
                        if (nms_hit(bi, ai, bj, box_area(bj), nms_thresh)) bits |= 1ull << j;
                    }
                }
                if (bits) atomicOr(&s_row[i], bits);
            }
        }
        __syncthreads();
        // B: sequential resolution in score order
        if (tid == 0) {
            unsigned long long alive = nc == 64 ? ~0ull : ((1ull << nc) - 1ull), kept = 0ull;
            while (alive) {
                const int i = __ffsll((long long)alive) - 1;
                kept |= 1ull << i;
                alive &= ~(s_row[i] | (1ull << i));
            }
            s_kept = kept;
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        if (tid < nc && ((kept >> tid) & 1ull)) kout[(int)(uint32_t)(key[cur[tid]] & 0xFFFFFFFFull) - 1] = 1;
        // C: later live candidates against the kept members; thread t owns a contiguous run so that one scan of the
        // per-thread counts keeps the score order
        const int rest = m - nc;
        const int per = (rest + NMS_THREADS - 1) / NMS_THREADS;             // <= npad / 1024
        const int b0 = nc + tid * per, b1 = min(b0 + per, m);
        uint32_t alive_bits = 0;
        int cnt = 0;
        for (int j = b0; j < b1; ++j) {
            const float4 bj = box[cur[j]];
            const float aj = box_area(bj);
            bool dead = false;
            unsigned long long km = kept;
            while (km && !dead) {
                const int i = __ffsll((long long)km) - 1;
                km &= km - 1ull;
                const float4 bi = box[cur[i]];
                dead = nms_hit(bi, box_area(bi), bj, aj, nms_thresh);
            }
            if (!dead) { alive_bits |= 1u << (j - b0); ++cnt; }
        }
        int incl = cnt;                                    // inclusive scan over the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        int wsum = s_wtot[lane];                           // every warp scans the 32 warp totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wsum, d);
            if (lane >= d) wsum += v;
        }
        const int total = __shfl_sync(0xffffffffu, wsum, 31);
        int pos = incl - cnt + (warp ? __shfl_sync(0xffffffffu, wsum, (warp + 31) & 31) : 0);
        for (int j = b0; j < b1; ++j)
            if ((alive_bits >> (j - b0)) & 1u) nxt[pos++] = cur[j];
        m = total;
        uint16_t *t = cur; cur = nxt; nxt = t;
        // the __syncthreads at the top of the next round orders these writes (and the reads of s_wtot / s_kept)
        // before their next use
    }
}

inline int next_pow2(int v) {
    int p = 64;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

CIM_API int cim_test_scores(const float *scores, float *out, int64_t M, int C1, int K, cim_stream_t stream) {
    if (!scores || !out) return CIM_ERR_ARG;
    if (M < 0 || C1 < 2 || K < 1 || K > 8) return CIM_ERR_ARG;
    if (M == 0) return CIM_OK;
    const long long n = M * (C1 - 1);
    cim_test_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(scores, out, M, C1, K);
    return cim_launch_status();
}

CIM_API int cim_box_nms(const float *boxes, const float *scores, int n, int n_classes, int score_stride,
                        float score_thresh, float nms_thresh, uint8_t *keep, cim_stream_t stream) {
    return cim_box_nms_batched(boxes, scores, 1, n, n_classes, score_stride, score_thresh, nms_thresh, keep, stream);
}

CIM_API int cim_box_nms_batched(const float *boxes, const float *scores, int n_img, int n, int n_classes,
                                int score_stride, float score_thresh, float nms_thresh, uint8_t *keep,
                                cim_stream_t stream) {
    if (!boxes || !scores || !keep) return CIM_ERR_ARG;
    if (n_img < 0 || n < 0 || n_classes < 0 || score_stride < n_classes) return CIM_ERR_ARG;
    if (n_img == 0 || n == 0 || n_classes == 0) return CIM_OK;
    if (n > 8192 || n_classes > 65535 || n_img > 65535) return CIM_ERR_SHAPE;
    if (!cim_aligned(boxes, 16)) return CIM_ERR_ALIGN;
    const int npad = next_pow2(n);
    const size_t smem = (size_t)npad * (8 + 16 + 2 + 2);
    if ((int)smem > cim_max_smem_optin()) return CIM_ERR_SHAPE;
    cudaFuncSetAttribute(cim_box_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cim_box_nms_kernel<<<dim3((unsigned)n_classes, (unsigned)n_img), NMS_THREADS, smem, (cudaStream_t)stream>>>(
        boxes, scores, n, n_classes, score_stride, score_thresh, nms_thresh, npad, keep);
    return cim_launch_status();
}
