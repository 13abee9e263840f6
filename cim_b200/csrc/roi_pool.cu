// roi_pool.cu -- RoIPool forward / backward for sm_100a.
//
// Replaces mmcv.ops.RoIPool at lib/modeling/model_builder.py:227-228 of the reference (the
// default ROI_XFORM_METHOD 'RoIPoolF', lib/core/config.py:366; no shipped yaml selects it).
// Two bin conventions (cim_roi_pool_fwd_ex `variant`):
//   LEGACY  the vendored lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93 (rounded integer ROI, +1 extent,
//           floor/ceil bin edges clipped to the map), which is also what torchvision.ops.roi_pool computes; pinned
//           against both.
//   MMCV    what the reference's import actually resolves to (lib/ops/__init__.py:6, mmcv-full 1.x, an un-vendored,
//           un-pinned dependency): float ROI corners x1 * s .. (x2 + 1) * s, bin edges floor(p * bin + start) ..
//           ceil((p + 1) * bin + start), ROIs with w <= 0 or h <= 0 pool nothing.  Restated from mmcv 1.x's published
//           roi_pool_cuda_kernel.cuh; mmcv itself is absent here, so this variant is "parity unpinned".
// Both: max with the first maximum in row-major scan order winning, empty bin -> 0 / argmax -1; argmax is the index
// inside the H*W plane.
//
// Forward: one warp per (roi, channel, 7 output bins of a row) would waste lanes on tiny bins,
// so the mapping is one thread per output element with the channel fastest across the warp's
// outputs of one ROI kept contiguous (coalesced stores); bins are a handful of pixels and the
// feature planes are L2 resident.  Backward: grad_feat[argmax] += grad_out.  The reference
// gathers per input pixel over ALL ROIs (O(B*C*H*W*K)); here each output element scatters once
// with red.global.add.f32 into a zeroed buffer (order of the fp32 sums is not fixed).
#include "common.cuh"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// MMCV = false: the vendored kernel's bins (rounded integer ROI corners, extent x2 - x1 + 1 >= 1, integer bin edges
// floor(ph * bin) + y1 .. ceil((ph + 1) * bin) + y1).  MMCV = true: mmcv 1.x's (float corners x1 * s .. (x2 + 1) * s,
// bin edges floor(ph * bin + y1) .. ceil((ph + 1) * bin + y1), a ROI with w <= 0 or h <= 0 pools nothing).
template <bool MMCV>
__global__ void roi_pool_fwd_kernel(const float *__restrict__ feat, const float *__restrict__ rois,
                                    float *__restrict__ out, int32_t *__restrict__ argmax, int B, int C,
                                    int H, int W, int K, int oh, int ow, float scale) {
    const int k = blockIdx.x;
    const float *r = rois + 5 * (size_t)k;
    const int b = (int)r[0];
    const int x1 = (int)roundf(r[1] * scale), y1 = (int)roundf(r[2] * scale);
    const int x2 = (int)roundf(r[3] * scale), y2 = (int)roundf(r[4] * scale);
    const int rw = max(x2 - x1 + 1, 1), rh = max(y2 - y1 + 1, 1);
    float bh = (float)rh / (float)oh, bw = (float)rw / (float)ow;
    const int per_roi = C * oh * ow;
    bool valid_b = b >= 0 && b < B;
    float fx1 = 0.f, fy1 = 0.f;
    if (MMCV) {
        fx1 = __fmul_rn(r[1], scale);
        fy1 = __fmul_rn(r[2], scale);
        const float fw = __fsub_rn(__fmul_rn(__fadd_rn(r[3], 1.f), scale), fx1);
        const float fh = __fsub_rn(__fmul_rn(__fadd_rn(r[4], 1.f), scale), fy1);
        if (fw <= 0.f || fh <= 0.f) valid_b = false;
        bw = __fdiv_rn(fw, (float)ow);
        bh = __fdiv_rn(fh, (float)oh);
    }
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < per_roi; e += gridDim.y * blockDim.x) {
        const int pw = e % ow, ph = (e / ow) % oh, c = e / (ow * oh);
        const size_t oidx = (size_t)k * per_roi + e;
        float best = 0.f;
        int besti = -1;
        if (valid_b) {
            int hs, he, ws, we;
            if (MMCV) {
                hs = clampi((int)floorf(__fadd_rn(__fmul_rn((float)ph, bh), fy1)), 0, H);
                he = clampi((int)ceilf(__fadd_rn(__fmul_rn((float)(ph + 1), bh), fy1)), 0, H);
                ws = clampi((int)floorf(__fadd_rn(__fmul_rn((float)pw, bw), fx1)), 0, W);
                we = clampi((int)ceilf(__fadd_rn(__fmul_rn((float)(pw + 1), bw), fx1)), 0, W);
            } else {
                hs = clampi((int)floorf((float)ph * bh) + y1, 0, H);
                he = clampi((int)ceilf((float)(ph + 1) * bh) + y1, 0, H);
                ws = clampi((int)floorf((float)pw * bw) + x1, 0, W);
                we = clampi((int)ceilf((float)(pw + 1) * bw) + x1, 0, W);
            }
            const float *plane = feat + ((size_t)b * C + c) * H * W;
            if (he > hs && we > ws) best = -3.402823466e+38f;
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    const float v = __ldg(plane + h * W + w);
                    if (v > best) { best = v; besti = h * W + w; }
                }
        }
        out[oidx] = best;
        if (argmax) argmax[oidx] = besti;
    }
}

__global__ void roi_pool_bwd_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ argmax,
                                    const float *__restrict__ rois, float *__restrict__ grad_feat, int B,
                                    int C, int H, int W, int K, int oh, int ow) {
    const int k = blockIdx.x;
    const int b = (int)rois[5 * (size_t)k];
    if (b < 0 || b >= B) return;
    const int per_roi = C * oh * ow, bins = oh * ow;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < per_roi; e += gridDim.y * blockDim.x) {
        const size_t oidx = (size_t)k * per_roi + e;
        const int a = argmax[oidx];
        if (a >= 0) atomicAdd(grad_feat + ((size_t)b * C + e / bins) * H * W + a, grad_out[oidx]);
    }
}

}  // namespace

CIM_API int cim_roi_pool_fwd(const float *feat, const float *rois, float *out, int32_t *argmax, int B, int C,
                             int H, int W, int K, int oh, int ow, float scale, cim_stream_t stream) {
    return cim_roi_pool_fwd_ex(feat, rois, out, argmax, B, C, H, W, K, oh, ow, scale, CIM_ROI_POOL_LEGACY, stream);
}

CIM_API int cim_roi_pool_fwd_ex(const float *feat, const float *rois, float *out, int32_t *argmax, int B, int C,
                                int H, int W, int K, int oh, int ow, float scale, int variant, cim_stream_t stream) {
    if (variant != CIM_ROI_POOL_LEGACY && variant != CIM_ROI_POOL_MMCV) return CIM_ERR_ARG;
    if (!feat || !out || (K > 0 && !rois)) return CIM_ERR_ARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K < 0 || oh <= 0 || ow <= 0) return CIM_ERR_ARG;
    if ((long long)C * oh * ow > (1LL << 30)) return CIM_ERR_SHAPE;
    if (K == 0) return CIM_OK;
    const int per_roi = C * oh * ow;
    dim3 grid((unsigned)K, (unsigned)min(64, (per_roi + 255) / 256));
    if (variant == CIM_ROI_POOL_MMCV)
        roi_pool_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, rois, out, argmax, B, C, H, W, K, oh,
                                                                           ow, scale);
    else
        roi_pool_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, rois, out, argmax, B, C, H, W, K, oh,
                                                                            ow, scale);
    return cim_launch_status();
}

CIM_API int cim_roi_pool_bwd(const float *grad_out, const int32_t *argmax, const float *rois, float *grad_feat,
                             int B, int C, int H, int W, int K, int oh, int ow, cim_stream_t stream) {
    if (!grad_out || !argmax || !grad_feat || (K > 0 && !rois)) return CIM_ERR_ARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K < 0 || oh <= 0 || ow <= 0) return CIM_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W, st);
    if (K == 0) return cim_launch_status();
    const int per_roi = C * oh * ow;
    dim3 grid((unsigned)K, (unsigned)min(64, (per_roi + 255) / 256));
    roi_pool_bwd_kernel<<<grid, 256, 0, st>>>(grad_out, argmax, rois, grad_feat, B, C, H, W, K, oh, ow);
    return cim_launch_status();
}
