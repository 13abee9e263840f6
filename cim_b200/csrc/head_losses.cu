// head_losses.cu -- the CIM loss block, forward + backward in one launch.
//
// Replaces, per image and refinement layer l, heads.cls_iou_loss (lib/modeling/heads.py:78-138) with its
// loss_weight_bag_loss (heads.py:43-74), plus heads.mil_bag_loss (heads.py:149-166) on the MIL heads, as wired
// at lib/modeling/model_builder.py:170-202, together with the gradient autograd would hand back for them:
//     cls_loss  = sum_{r in ind, c} -P[r,c] log(cls[r,c]) w[r] / sum_{r in ind, c} P[r,c]
//     iou_loss  = sum_{r in fg} smooth_l1(sum_c P[r,c] iou[r,c] - pseudo_iou[r]) w[r] / sum_{r in fg, c} P[r,c]
//     bag_loss  = mean_c BCE(clamp(agg[c]), label[c]) * lw[c],   agg[c] = max_r of ind P cls iou (present classes,
//                 background included) or max_r cls iou (absent classes); lw = w[argmax] resp. 1
//     mil_bag   = mean_c BCE(clamp(sum_r predict_cls predict_det), [1, labels])
// with P = (pseudo_labels != 0), ind = rows with any P, fg = rows with a foreground P, w = lmda * loss_weights,
// every score clamped to [1e-6, 1 - 1e-6] first (gradient 0 outside, as torch.clamp).  The reference runs ~60
// small launches and several host syncs per image for this; here one thread-block cluster per (head pair, image)
// makes three passes over its [R, C+1] slices (the rows split over the cluster's CTAs).  The gradient w.r.t. the
// score tensor [2+2K, n_img*R, C+1] is written completely (zeros where nothing flows), ready for
// cim_score_heads_bwd.  PCL_loss (heads.py:10-41, needs the dataset's cluster matrix) is not part of this kernel.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

// One thread-block CLUSTER per (head pair, image): the R rows are split evenly over the CL CTAs of the cluster,
// which exchange their partial sums and per-class maxima through distributed shared memory (two cluster barriers
// per refinement layer, one for the MIL heads).  Round 1 ran one CTA per (head pair, image) -- 32 CTAs on 148 SMs,
// 0.47 ms at C = 80 of pure latency; see DESIGN 4.7.  Partials are combined in cluster-rank order by every CTA, so
// the results do not depend on scheduling.
constexpr int LT = 1024;
constexpr int MAX_CL = 8;
constexpr float LO = 1e-6f, HI = 1.f - 1e-6f;

__device__ __forceinline__ float clampf(float v) { return fminf(fmaxf(v, LO), HI); }
__device__ __forceinline__ bool in_clamp(float v) { return v >= LO && v <= HI; }

// deterministic block sum (fixed tree); all threads get the result
__device__ float block_sum(float v, float *red) {
    const int tid = threadIdx.x;
    __syncthreads();
    red[tid] = v;
    __syncthreads();
    for (int o = LT / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    return red[0];
}

__device__ __forceinline__ unsigned long long max_key(float v, int r) {       // v >= 0; ties: lowest r
    return ((unsigned long long)__float_as_uint(v) << 32) | (uint32_t)(~r);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct LossArgs {
    const float *scores, *pseudo_labels, *loss_weights, *labels;
    const __half *pseudo_iou;
    const uint8_t *valid;
    float *losses, *grad;
    int n_img, R, C1, K, n_layers, rows_per;
    float lmda0, lmda_rest, iou_weight, grad_scale;
};

__global__ void __launch_bounds__(LT)
cim_head_losses_kernel(LossArgs a) {
    extern __shared__ unsigned char dyn[];
    float *dsl = reinterpret_cast<float *>(dyn);                                  // [rows_per] smooth-l1' * w per fg row
    unsigned char *rowinfo = reinterpret_cast<unsigned char *>(dsl + a.rows_per); // [rows_per] bit0 ind, bit1 fg
    __shared__ float red[LT];
    __shared__ unsigned long long mx_fg[1024], mx_un[1024];                       // this CTA's maxima, read by peers
    __shared__ float col_p[1024];                                                 // this CTA's column sums (MIL)
    __shared__ float cl_part[4];                                                  // this CTA's scalar partials
    __shared__ float col_g[1024];
    __shared__ int col_idx[1024];
    __shared__ float col_l[1024];

    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int slot = blockIdx.x / CL, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = a.R, C1 = a.C1, K = a.K;
    const int r0 = min(R, rank * a.rows_per), r1 = min(R, r0 + a.rows_per);      // this CTA's rows
    const long long M = (long long)a.n_img * R;
    const int G = LT / C1, rg = tid / C1, c = tid - rg * C1;                      // element passes: thread = (rg, c)
    const bool elem = tid < G * C1;
    float *lout = a.losses + ((size_t)b * (K + 1) + slot) * 3;
    const float *lab = a.labels + (size_t)b * (C1 - 1);
    auto label_tmp = [&](int cc) { return cc == 0 ? 1.f : lab[cc - 1]; };          // heads.py:84-85 / :157-158

    if (slot == K) {
        // ---------------------------------------------------------------- mil_bag_loss (heads.py:149-166)
        const float *pc = a.scores + ((size_t)0 * M + (size_t)b * R) * C1;
        const float *pd = a.scores + ((size_t)1 * M + (size_t)b * R) * C1;
        float part = 0.f;
        if (elem)
            for (int r = r0 + rg; r < r1; r += G) part = fmaf(pc[(size_t)r * C1 + c], pd[(size_t)r * C1 + c], part);
        red[tid] = part;
        __syncthreads();
        if (tid < C1) {
            float s = 0.f;
            for (int g = 0; g < G; ++g) s += red[g * C1 + tid];
            col_p[tid] = s;
        }
        cluster.sync();
        if (tid < C1) {
            float s = 0.f;
            for (int q = 0; q < CL; ++q) s += *cluster.map_shared_rank(&col_p[tid], q);
            const float l = label_tmp(tid), p = clampf(s);
            col_g[tid] = in_clamp(s) ? -(l / p - (1.f - l) / (1.f - p)) / C1 : 0.f;
            col_l[tid] = -(l * logf(p) + (1.f - l) * logf(1.f - p));
        }
        __syncthreads();
        if (tid == 0 && rank == 0) {                                              // per-class losses, class order
            float loss = 0.f;
            for (int cc = 0; cc < C1; ++cc) loss += col_l[cc];
            lout[0] = 0.f; lout[1] = 0.f; lout[2] = loss / C1;
        }
        if (a.grad && elem) {
            float *gc = a.grad + ((size_t)0 * M + (size_t)b * R) * C1, *gd = a.grad + ((size_t)1 * M + (size_t)b * R) * C1;
            const float g = col_g[c] * a.grad_scale;
            for (int r = r0 + rg; r < r1; r += G) {
                const size_t e = (size_t)r * C1 + c;
                gc[e] = g * pd[e];
                gd[e] = g * pc[e];
            }
        }
        cluster.sync();                                    // peers may still be reading col_p
        return;
    }

    // -------------------------------------------------------------------- refinement layer `slot`
    const int l = slot;
    const float *cls = a.scores + ((size_t)(2 + l) * M + (size_t)b * R) * C1;
    const float *iou = a.scores + ((size_t)(2 + K + l) * M + (size_t)b * R) * C1;
    float *g_cls = a.grad ? a.grad + ((size_t)(2 + l) * M + (size_t)b * R) * C1 : nullptr;
    float *g_iou = a.grad ? a.grad + ((size_t)(2 + K + l) * M + (size_t)b * R) * C1 : nullptr;
    const bool live = l < a.n_layers && a.valid[(size_t)l * a.n_img + b] != 0;     // model_builder.py:189-190
    if (!live) {                                           // the same for every CTA of the cluster
        if (tid == 0 && rank == 0) { lout[0] = 0.f; lout[1] = 0.f; lout[2] = 0.f; }
        if (a.grad)
            for (int e = r0 * C1 + tid; e < r1 * C1; e += LT) { g_cls[e] = 0.f; g_iou[e] = 0.f; }
        return;
    }
    const float *P = a.pseudo_labels + (((size_t)l * a.n_img + b) * R) * C1;
    const float *lw = a.loss_weights + ((size_t)l * a.n_img + b) * R;
    const __half *pi = a.pseudo_iou + ((size_t)l * a.n_img + b) * R;
    const float lmda = l == 0 ? a.lmda0 : a.lmda_rest;                             // model_builder.py:172,194

    // pass A, warp per row: ind / fg flags, denominators, the IoU regression term (heads.py:88,102-132)
    float cls_den = 0.f, iou_den = 0.f, iou_part = 0.f;
    for (int r = r0 + warp; r < r1; r += LT / 32) {
        const float *p = P + (size_t)r * C1, *io = iou + (size_t)r * C1;
        int nz = 0, nz_fg = 0;
        float s = 0.f;
        for (int cc = lane; cc < C1; cc += 32)
            if (p[cc] != 0.f) {
                ++nz;
                nz_fg += cc > 0;
                s += clampf(io[cc]);
            }
        nz = warp_sum(nz);
        nz_fg = warp_sum(nz_fg);
        s = warp_sum(s);
        if (lane == 0) {
            rowinfo[r - r0] = (unsigned char)((nz > 0) | ((nz_fg > 0) << 1));
            float d1 = 0.f;
            if (nz > 0) cls_den += (float)nz;
            if (nz_fg > 0) {
                iou_den += (float)nz;
                const float w = lmda * lw[r], d = s - __half2float(pi[r]);
                const float ad = fabsf(d);
                iou_part += (ad < 1.f ? 0.5f * d * d : ad - 0.5f) * w;             // smooth_l1, beta = 1
                d1 = (d < -1.f ? -1.f : (d > 1.f ? 1.f : d)) * w;                  // NaN labels stay NaN
            }
            dsl[r - r0] = d1;
        }
    }
    cls_den = block_sum(cls_den, red);
    iou_den = block_sum(iou_den, red);
    iou_part = block_sum(iou_part, red);
    if (tid < C1) { mx_fg[tid] = 0ull; mx_un[tid] = 0ull; }
    __syncthreads();

    // pass B, thread = (row group, class): weighted CE and the per-class maxima of the bag loss
    float cls_part = 0.f;
    if (elem) {
        unsigned long long best_fg = 0ull, best_un = 0ull;
        for (int r = r0 + rg; r < r1; r += G) {
            const size_t e = (size_t)r * C1 + c;
            const float cl = clampf(cls[e]), io = clampf(iou[e]);
            const bool p = P[e] != 0.f, ind = rowinfo[r - r0] & 1;
            if (ind && p) cls_part += -logf(cl) * (lmda * lw[r]);
            const float pred = cl * io;
            const unsigned long long kf = max_key((ind && p) ? pred : 0.f, r), ku = max_key(pred, r);
            best_fg = kf > best_fg ? kf : best_fg;
            best_un = ku > best_un ? ku : best_un;
        }
        atomicMax(&mx_fg[c], best_fg);
        atomicMax(&mx_un[c], best_un);
    }
    cls_part = block_sum(cls_part, red);
    if (tid == 0) { cl_part[0] = cls_den; cl_part[1] = iou_den; cl_part[2] = iou_part; cl_part[3] = cls_part; }
    cluster.sync();
    // every CTA combines the partials of the cluster in rank order
    cls_den = iou_den = iou_part = cls_part = 0.f;
    for (int q = 0; q < CL; ++q) {
        const float *pp = cluster.map_shared_rank(&cl_part[0], q);
        cls_den += pp[0]; iou_den += pp[1]; iou_part += pp[2]; cls_part += pp[3];
    }
    if (tid < C1) {                                                                // heads.py:55-72
        const float lbl = label_tmp(tid);
        unsigned long long k = 0ull;                                               // a CTA without rows holds 0
        for (int q = 0; q < CL; ++q) {
            const unsigned long long kq = *cluster.map_shared_rank(lbl == 1.f ? &mx_fg[tid] : &mx_un[tid], q);
            k = kq > k ? kq : k;
        }
        const float raw = __uint_as_float((uint32_t)(k >> 32));
        const int idx = k ? (int)(~(uint32_t)(k & 0xFFFFFFFFull)) : 0;
        const float agg = clampf(raw);
        const float wgt = lbl == 1.f ? lmda * lw[idx] : 1.f;
        col_l[tid] = -(lbl * logf(agg) + (1.f - lbl) * logf(1.f - agg)) * wgt;
        float g = in_clamp(raw) ? -(lbl / agg - (1.f - lbl) / (1.f - agg)) * wgt / C1 : 0.f;
        if (lbl == 1.f && P[(size_t)idx * C1 + tid] == 0.f) g = 0.f;             // ind * P factor (P != 0 implies ind)
        col_g[tid] = g;
        col_idx[tid] = idx;
    }
    __syncthreads();
    if (tid == 0 && rank == 0) {
        float bag = 0.f;
        for (int cc = 0; cc < C1; ++cc) bag += col_l[cc];
        lout[0] = cls_den > 0.f ? cls_part / cls_den : 0.f;                        // heads.py:105,115-116
        lout[1] = iou_den > 0.f ? iou_part / iou_den : 0.f;                        // heads.py:131-132
        lout[2] = bag / C1;
    }
    // pass C: gradients of (cls_loss + iou_weight * iou_loss + bag_loss) * grad_scale
    if (a.grad && elem) {
        const float gi = col_g[c];
        const int idx = col_idx[c];
        const float inv_cls = cls_den > 0.f ? 1.f / cls_den : 0.f, inv_iou = iou_den > 0.f ? a.iou_weight / iou_den : 0.f;
        for (int r = r0 + rg; r < r1; r += G) {
            const size_t e = (size_t)r * C1 + c;
            const float cr = cls[e], ir = iou[e], cl = clampf(cr), io = clampf(ir);
            const bool p = P[e] != 0.f;
            const unsigned char info = rowinfo[r - r0];
            float gc = 0.f, gio = 0.f;
            if ((info & 1) && p) gc = -(lmda * lw[r]) * inv_cls / cl;
            if (info & 2) gio = (p ? 1.f : 0.f) * (dsl[r - r0] * inv_iou);       // 0 * NaN = NaN, as autograd
            if (r == idx) { gc += gi * io; gio += gi * cl; }
            g_cls[e] = in_clamp(cr) ? gc * a.grad_scale : 0.f;
            g_iou[e] = in_clamp(ir) ? gio * a.grad_scale : 0.f;
        }
    }
    cluster.sync();                                        // peers may still be reading cl_part / mx_*
}

}  // namespace

CIM_API int cim_head_losses(const float *scores, const float *pseudo_labels, const void *pseudo_iou_f16,
                            const float *loss_weights, const uint8_t *valid, const float *labels, float *losses,
                            float *grad_scores, int n_img, int R, int C, int K, int n_layers, float lmda0,
                            float lmda_rest, float iou_weight, float grad_scale, cim_stream_t stream) {
    if (!scores || !pseudo_labels || !pseudo_iou_f16 || !loss_weights || !valid || !labels || !losses) return CIM_ERR_ARG;
    if (n_img < 0 || R < 0 || C < 1 || K < 1 || K > 8 || n_layers < 0 || n_layers > K) return CIM_ERR_ARG;
    if (n_img == 0) return CIM_OK;
    if (R == 0 || R > 10240 || C + 1 > 1024 || n_img > 65535) return CIM_ERR_SHAPE;
    // cluster size: enough CTAs to cover the SMs, at least ~128 rows each
    int cl = 1;
    while (cl < MAX_CL && (long long)(K + 1) * n_img * cl < 2 * cim_num_sms() && R / (cl * 2) >= 128) cl *= 2;
    LossArgs a;
    a.scores = scores; a.pseudo_labels = pseudo_labels; a.loss_weights = loss_weights; a.labels = labels;
    a.pseudo_iou = reinterpret_cast<const __half *>(pseudo_iou_f16);
    a.valid = valid; a.losses = losses; a.grad = grad_scores;
    a.n_img = n_img; a.R = R; a.C1 = C + 1; a.K = K; a.n_layers = n_layers;
    a.rows_per = (R + cl - 1) / cl;
    a.lmda0 = lmda0; a.lmda_rest = lmda_rest; a.iou_weight = iou_weight; a.grad_scale = grad_scale;
    const size_t smem = (size_t)a.rows_per * 5 + 16;
    cudaFuncSetAttribute(cim_head_losses_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((K + 1) * cl), (unsigned)n_img);
    cfg.blockDim = dim3(LT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, cim_head_losses_kernel, a);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return cim_launch_status();
}
