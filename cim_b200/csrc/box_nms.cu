// box_nms.cu -- test-time post-processing of the CIM heads on the GPU.
//
//  cim_test_scores : lib/core/test.py:130-133 + lib/modeling/model_builder.py:60-68 -- the refinement heads'
//                    (cls * iou)[:, 1:] averaged over the K heads: out[r][c] = mean_k cls_k[r][c+1] * iou_k[r][c+1]
//                    (summed in head order then divided by K, as `scores += ...; scores /= K` does).
//  cim_box_nms     : the per-class loop of lib/utils/mask_eval_utils.py:57-79 (and box_results_with_nms_and_limit):
//                    candidates = scores[:, c] > score_thresh, greedy NMS of lib/utils/cython_nms.pyx:37-87:
//                    areas (x2-x1+1)(y2-y1+1), boxes visited by descending score, a later box is suppressed when
//                    inter / (area_i + area_j - inter) >= nms_thresh, all in float32.  One CTA per class:
//                    bitonic sort of the candidates in shared memory, then 64-box chunks -- the chunk is resolved
//                    sequentially by one warp, its survivors are applied to all later boxes by the whole CTA.
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
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
    return ovr >= thr;
}

__global__ void __launch_bounds__(NMS_THREADS)
cim_box_nms_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, int n, int ncls, int score_stride,
                   float score_thresh, float nms_thresh, int npad, uint8_t *__restrict__ keep) {
    extern __shared__ unsigned char smem_raw[];
    unsigned long long *key = reinterpret_cast<unsigned long long *>(smem_raw);      // [npad] (score bits, index)
    float4 *box = reinterpret_cast<float4 *>(key + npad);                            // [npad] sorted order
    unsigned char *supp = reinterpret_cast<unsigned char *>(box + npad);             // [npad]
    __shared__ int s_count;
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
    // bitonic sort, descending
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = tid; i < npad; i += NMS_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = key[i], b = key[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { key[i] = b; key[p] = a; }
                }
            }
        }
    __syncthreads();
    const int m = s_count;
    for (int i = tid; i < m; i += NMS_THREADS) {
        const int src = (int)(uint32_t)(key[i] & 0xFFFFFFFFull) - 1;
        const float4 b = *reinterpret_cast<const float4 *>(boxes + (size_t)src * 4);
        box[i] = b;
        supp[i] = 0;
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int s0 = 0; s0 < m; s0 += CHUNK) {
        const int s1 = min(s0 + CHUNK, m);
        if (warp == 0) {                                  // the chunk against itself, in score order
            for (int i = s0; i < s1; ++i) {
                if (!supp[i]) {                           // warp-uniform: supp[] is only written under __syncwarp
                    const float4 bi = box[i];
                    const float ai = box_area(bi);
                    for (int j = i + 1 + lane; j < s1; j += 32)
                        if (!supp[j] && nms_hit(bi, ai, box[j], box_area(box[j]), nms_thresh)) supp[j] = 1;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int j = s1 + tid; j < m; j += NMS_THREADS) {  // survivors of the chunk against everything later
            const float4 bj = box[j];
            const float aj = box_area(bj);
            bool dead = supp[j];
            for (int i = s0; i < s1 && !dead; ++i)
                if (!supp[i] && nms_hit(box[i], box_area(box[i]), bj, aj, nms_thresh)) dead = true;
            if (dead) supp[j] = 1;
        }
        __syncthreads();
    }
    for (int i = tid; i < m; i += NMS_THREADS)
        if (!supp[i]) kout[(int)(uint32_t)(key[i] & 0xFFFFFFFFull) - 1] = 1;
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
    const size_t smem = (size_t)npad * (8 + 16 + 1);
    if ((int)smem > cim_max_smem_optin()) return CIM_ERR_SHAPE;
    cudaFuncSetAttribute(cim_box_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cim_box_nms_kernel<<<dim3((unsigned)n_classes, (unsigned)n_img), NMS_THREADS, smem, (cudaStream_t)stream>>>(
        boxes, scores, n, n_classes, score_stride, score_thresh, nms_thresh, npad, keep);
    return cim_launch_status();
}
