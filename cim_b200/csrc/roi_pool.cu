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
// Two implementations behind the same entry points (cim_set_debug_flags(CIM_DBG_ROI_POOL_SIMPLE) forces the second):
//  * TILE kernels (7 x 7 bins, C a multiple of 8, maps of up to ~6400 cells): a CTA keeps CH = 32 / 16 / 8 channels
//    of one image's map in shared memory, channel-interleaved (tile[cell][c]: the 32 lanes of a warp = 32 channels
//    read one cell without bank conflicts and share every bin bound, so there is no divergence), and its warps take
//    the image's ROIs one at a time.
//      forward   49 bins per (roi, chunk); a bin is scanned in row-major order (first maximum wins, as the
//                reference's `>`), 4 columns per step with the columns past the bin clamped to its last cell (a
//                duplicate never beats the running maximum, so no predicates); values and argmax are staged in
//                shared memory as the [CH][49] block they are in global memory and leave with ONE bulk copy
//                (cp.async.bulk) / coalesced int32 stores.
//      backward  grad tile in shared memory (swizzled so that both lane = channel and lane = cell are conflict
//                free), red.shared.add per pooled element straight from coalesced reads of grad_out / argmax, tile
//                added to grad_feat once per CTA.  fp32 sums in unspecified order, as the reference's atomicAdd.
//    Units = (image, channel chunk, ROI split); a CTA compacts the ROIs of its image and split into a shared-memory
//    list (the rois need not be grouped by image) and its warps pop them dynamically.
//  * SIMPLE kernels (any geometry): one thread per output element, red.global.add backward -- round 1's kernels.
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


// ------------------------------------------------------------------------------------------ tile kernels
constexpr int PO = 7, PBINS = PO * PO;                  // bins of the tile path
constexpr int LIST = 1024;                              // ROI ids compacted per round
constexpr int BWD_WARPS = 12;
// warps of a forward CTA: the bin scan is a chain of dependent shared-memory loads and selects (~0.15 IPC per warp),
// so the kernel lives on warps per SM; 16 / 8-channel tiles leave room for 24 staging buffers, a 32-channel tile for 10
__host__ __device__ constexpr int fwd_warps(int ch) { return ch == 32 ? 10 : 24; }

struct PoolBins {                                       // bin bounds of one ROI, identical in every lane
    int ws[PO], we[PO];
    int y1;
    float bh, fy1;
    bool valid;
};

template <bool MMCV>
__device__ __forceinline__ void pool_bins(const float *r, float scale, int B, int W, PoolBins &pb) {
    const int b = (int)r[0];
    pb.valid = b >= 0 && b < B;
    if (MMCV) {
        const float fx1 = __fmul_rn(r[1], scale);
        pb.fy1 = __fmul_rn(r[2], scale);
        const float fw = __fsub_rn(__fmul_rn(__fadd_rn(r[3], 1.f), scale), fx1);
        const float fh = __fsub_rn(__fmul_rn(__fadd_rn(r[4], 1.f), scale), pb.fy1);
        if (fw <= 0.f || fh <= 0.f) pb.valid = false;
        const float bw = __fdiv_rn(fw, (float)PO);
        pb.bh = __fdiv_rn(fh, (float)PO);
        pb.y1 = 0;
#pragma unroll
        for (int pw = 0; pw < PO; ++pw) {
            pb.ws[pw] = clampi((int)floorf(__fadd_rn(__fmul_rn((float)pw, bw), fx1)), 0, W);
            pb.we[pw] = clampi((int)ceilf(__fadd_rn(__fmul_rn((float)(pw + 1), bw), fx1)), 0, W);
        }
    } else {
        const int x1 = (int)roundf(r[1] * scale), x2 = (int)roundf(r[3] * scale);
        const int y2 = (int)roundf(r[4] * scale);
        pb.y1 = (int)roundf(r[2] * scale);
        const int rw = max(x2 - x1 + 1, 1), rh = max(y2 - pb.y1 + 1, 1);
        const float bw = (float)rw / (float)PO;
        pb.bh = (float)rh / (float)PO;
        pb.fy1 = 0.f;
#pragma unroll
        for (int pw = 0; pw < PO; ++pw) {
            pb.ws[pw] = clampi((int)floorf((float)pw * bw) + x1, 0, W);
            pb.we[pw] = clampi((int)ceilf((float)(pw + 1) * bw) + x1, 0, W);
        }
    }
}

// ROIs k = split + S * j (j in [j0, j0 + LIST)) that belong to image b -> list[]; image 0 also takes the ROIs whose
// batch index is out of range (they pool nothing but their outputs must be written).  Returns the count.
__device__ __forceinline__ int pool_collect(const float *__restrict__ rois, int K, int B, int b, int split, int S,
                                            int j0, int *list, int *s_n, int nthreads) {
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    for (int j = j0 + threadIdx.x; j < j0 + LIST; j += nthreads) {
        const long long k = split + (long long)S * j;
        if (k >= K) break;
        const int rb = (int)__ldg(rois + 5 * k);
        const bool bad = rb < 0 || rb >= B;
        if (rb == b || (bad && b == 0)) list[atomicAdd(s_n, 1)] = (int)k;
    }
    __syncthreads();
    return *s_n;
}

template <int CH, bool MMCV>
__global__ void __launch_bounds__(fwd_warps(CH) * 32, 1)
roi_pool_fwd_tile_kernel(const float *__restrict__ feat, const float *__restrict__ rois, float *__restrict__ out,
                         int32_t *__restrict__ argmax, int B, int C, int H, int W, int K, float scale, int S) {
    constexpr int SUB = 32 / CH;                        // bin rows handled side by side in a warp
    constexpr int FWD_WARPS = fwd_warps(CH), NT = FWD_WARPS * 32;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int HW = H * W;
    float *tile = reinterpret_cast<float *>(dyn);                                     // [HW][CH]
    unsigned char *stage0 = dyn + (((size_t)HW * CH * 4 + 15) & ~(size_t)15);
    constexpr int STAGE_BYTES = CH * PBINS * 4 + ((CH * PBINS * 2 + 15) & ~15);
    int *list = reinterpret_cast<int *>(stage0 + (size_t)FWD_WARPS * STAGE_BYTES);    // [LIST]
    __shared__ int s_n, s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunks = C / CH;
    const int split = blockIdx.x % S, chunk = (blockIdx.x / S) % chunks, b = blockIdx.x / (S * chunks);
    const int c0 = chunk * CH, c = lane % CH, sub = lane / CH;
    float *sv = reinterpret_cast<float *>(stage0 + (size_t)warp * STAGE_BYTES);       // [CH][49] values
    int16_t *sa = reinterpret_cast<int16_t *>(sv + CH * PBINS);                       // [CH][49] argmax

    // the map: lane = channel, so the shared-memory stores are conflict free (global reads: one 16-byte piece per
    // lane and plane, the planes are L2 resident)
    {
        const float *plane = feat + ((size_t)b * C + c0 + c) * HW;
        if ((HW & 3) == 0) {
            for (int q = warp * SUB + sub; q < (HW >> 2); q += FWD_WARPS * SUB) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(plane) + q);
                float *t = tile + (size_t)(4 * q) * CH + c;
                t[0] = v.x; t[CH] = v.y; t[2 * CH] = v.z; t[3 * CH] = v.w;
            }
        } else {
            for (int q = warp * SUB + sub; q < HW; q += FWD_WARPS * SUB) tile[(size_t)q * CH + c] = __ldg(plane + q);
        }
    }
    const float *tp = tile + c;
    bool pending = false;                               // lane 0: a bulk store still reads sv
    const int rounds = (K + S * LIST - 1) / (S * LIST);
    for (int round = 0; round < rounds; ++round) {
        const int n = pool_collect(rois, K, B, b, split, S, round * LIST, list, &s_n, NT);
        if (tid == 0) s_next = 0;
        __syncthreads();                                // (also orders the tile stores before the first ROI)
        while (true) {
            int it = 0;
            if (lane == 0) it = atomicAdd(&s_next, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
            if (it >= n) break;
            const int k = list[it];
            PoolBins pb;
            pool_bins<MMCV>(rois + 5 * (size_t)k, scale, B, W, pb);
            if (lane == 0 && pending) { bulk_wait_read<0>(); pending = false; }
            __syncwarp();                               // the previous ROI's staging has been read
            for (int ph = sub; ph < PO; ph += SUB) {
                int hs, he;
                if (MMCV) {
                    hs = clampi((int)floorf(__fadd_rn(__fmul_rn((float)ph, pb.bh), pb.fy1)), 0, H);
                    he = clampi((int)ceilf(__fadd_rn(__fmul_rn((float)(ph + 1), pb.bh), pb.fy1)), 0, H);
                } else {
                    hs = clampi((int)floorf((float)ph * pb.bh) + pb.y1, 0, H);
                    he = clampi((int)ceilf((float)(ph + 1) * pb.bh) + pb.y1, 0, H);
                }
#pragma unroll
                for (int pw = 0; pw < PO; ++pw) {
                    const int ws = pb.ws[pw], we = pb.we[pw];
                    float best = 0.f;
                    int besti = -1;
                    if (pb.valid && he > hs && we > ws) {
                        best = -3.402823466e+38f;
                        const int last = we - 1;
                        for (int h = hs; h < he; ++h) {
                            const int rowi = h * W;
                            const float *rowp = tp + (size_t)rowi * CH;
                            for (int w0 = ws; w0 < we; w0 += 4) {
                                const float v0 = rowp[w0 * CH], v1 = rowp[min(w0 + 1, last) * CH];
                                const float v2 = rowp[min(w0 + 2, last) * CH], v3 = rowp[min(w0 + 3, last) * CH];
                                if (v0 > best) { best = v0; besti = rowi + w0; }
                                if (v1 > best) { best = v1; besti = rowi + w0 + 1; }
                                if (v2 > best) { best = v2; besti = rowi + w0 + 2; }
                                if (v3 > best) { best = v3; besti = rowi + w0 + 3; }
                            }
                        }
                    }
                    sv[c * PBINS + ph * PO + pw] = best;
                    sa[c * PBINS + ph * PO + pw] = (int16_t)besti;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            const size_t obase = ((size_t)k * C + c0) * PBINS;
            if (lane == 0) {
                bulk_s2g(out + obase, sv, CH * PBINS * 4);
                bulk_commit();
                pending = true;
            }
            if (argmax)
                for (int i = lane; i < CH * PBINS; i += 32) argmax[obase + i] = (int)sa[i];
        }
        // list[] / s_next are rewritten by the next round: every warp must have left the loop
        __syncthreads();
    }
    if (lane == 0 && pending) bulk_wait_read<0>();
}

// swizzled grad tile: element (cell, c) lives at cell * CH + ((c + cell / SUB) & (CH - 1)), SUB = 32 / CH: 32 lanes
// holding 32 channels of one cell (or CH channels of SUB cells) and 32 lanes holding 32 consecutive cells of one
// channel both touch 32 different banks
template <int CH>
__device__ __forceinline__ int pool_sw(int cell, int c) {
    return cell * CH + ((c + cell / (32 / CH)) & (CH - 1));
}

template <int CH>
__global__ void __launch_bounds__(BWD_WARPS * 32, 1)
roi_pool_bwd_tile_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ argmax,
                         const float *__restrict__ rois, float *__restrict__ grad_feat, int B, int C, int H, int W,
                         int K, int S) {
    constexpr int SUB = 32 / CH, NT = BWD_WARPS * 32, UNIT = CH * PBINS;             // elements per (roi, chunk)
    extern __shared__ __align__(16) unsigned char dyn[];
    const int HW = H * W;
    float *tile = reinterpret_cast<float *>(dyn);                                     // [HW][CH], swizzled
    unsigned char *stage0 = dyn + (((size_t)HW * CH * 4 + 15) & ~(size_t)15);         // [warps][2][grad | argmax]
    int *list = reinterpret_cast<int *>(stage0 + (size_t)BWD_WARPS * 2 * UNIT * 8);   // [LIST]
    uint64_t *bars = reinterpret_cast<uint64_t *>(list + LIST);                       // [warps][2]
    __shared__ int s_n, s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunks = C / CH;
    const int split = blockIdx.x % S, chunk = (blockIdx.x / S) % chunks, b = blockIdx.x / (S * chunks);
    const int c0 = chunk * CH, c = lane % CH, sub = lane / CH;
    unsigned char *stage_w = stage0 + (size_t)warp * 2 * UNIT * 8;
    uint64_t *bar = bars + warp * 2;
    for (int i = tid; i < HW * CH; i += NT) tile[i] = 0.f;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    // a warp streams its (roi, chunk) units -- [CH][49] gradients and argmax, contiguous in global memory -- through
    // two staging buffers with bulk copies, one unit ahead; lanes = channels (x SUB bin phases), so the lanes of one
    // shared-memory atomic never meet in one address (fp32 atomics on shared memory are CAS loops)
    auto pop = [&]() {
        int it = 0;
        if (lane == 0) it = atomicAdd(&s_next, 1);
        return __shfl_sync(0xffffffffu, it, 0);
    };
    auto issue = [&](int it, int buf) {
        if (lane == 0) {
            const size_t obase = ((size_t)list[it] * C + c0) * PBINS;
            unsigned char *dst = stage_w + (size_t)buf * UNIT * 8;
            mbar_expect_tx(&bar[buf], UNIT * 8);
            bulk_g2s(dst, grad_out + obase, UNIT * 4, &bar[buf]);
            bulk_g2s(dst + UNIT * 4, argmax + obase, UNIT * 4, &bar[buf]);
        }
    };
    uint32_t phase0 = 0, phase1 = 0;
    int buf = 0;
    const int rounds = (K + S * LIST - 1) / (S * LIST);
    for (int round = 0; round < rounds; ++round) {
        const int n = pool_collect(rois, K, B, b, split, S, round * LIST, list, &s_n, NT);
        if (tid == 0) s_next = 0;
        __syncthreads();                                // (first round: tile zeroed, barriers initialised)
        int it = pop();
        if (it < n) issue(it, buf);
        while (it < n) {
            const int nxt = pop();
            if (nxt < n) issue(nxt, buf ^ 1);
            if (buf == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
            else { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
            const int k = list[it];
            if ((int)__ldg(rois + 5 * (size_t)k) == b) {                  // out-of-range batch index: no gradient
                const float *sg = reinterpret_cast<const float *>(stage_w + (size_t)buf * UNIT * 8);
                const int *sa = reinterpret_cast<const int *>(sg + UNIT);
#pragma unroll 5
                for (int bin = sub; bin < PBINS; bin += SUB) {
                    const int a = sa[c * PBINS + bin];
                    const float g = sg[c * PBINS + bin];
                    if (a >= 0 && a < HW) atomicAdd(&tile[pool_sw<CH>(a, c)], g);
                }
            }
            __syncwarp();                               // the buffer may be refilled by the next issue
            buf ^= 1;
            it = nxt;
        }
        __syncthreads();
    }
    // the tile -> grad_feat (zeroed by the launcher when S > 1: several CTAs add into the same planes)
    for (int cc = 0; cc < CH; ++cc) {
        float *plane = grad_feat + ((size_t)b * C + c0 + cc) * HW;
        for (int cell = tid; cell < HW; cell += NT) {
            const float v = tile[pool_sw<CH>(cell, cc)];
            if (S == 1) plane[cell] = v;
            else if (v != 0.f) atomicAdd(plane + cell, v);
        }
    }
}

inline size_t pool_bwd_other_bytes(int ch) { return (size_t)BWD_WARPS * 2 * ch * PBINS * 8 + LIST * 4 + BWD_WARPS * 16 + 32; }
inline size_t pool_fwd_stage_bytes(int ch) {
    return (size_t)fwd_warps(ch) * ((size_t)ch * PBINS * 4 + (((size_t)ch * PBINS * 2 + 15) & ~(size_t)15)) + LIST * 4 + 16;
}
// ROI splits per (image, channel chunk): enough units for ~6 waves of one CTA per SM
inline int pool_splits(int B, int chunks) {
    const int want = (6 * cim_num_sms() + B * chunks - 1) / (B * chunks);
    return max(1, min(8, want));
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
    cudaStream_t st = (cudaStream_t)stream;
    const bool simple = (cim_get_debug_flags() & CIM_DBG_ROI_POOL_SIMPLE) != 0;
    const int HW = H * W;
    if (!simple && oh == PO && ow == PO && HW < 32768 && cim_aligned(out, 16) && cim_aligned(feat, 16)) {
        int ch = 0;                                     // 16 first: 24 warps per SM (see fwd_warps)
        for (int t : {16, 8, 32})
            if (!ch && C % t == 0 && (size_t)HW * t * 4 + pool_fwd_stage_bytes(t) + 1024 <= (size_t)cim_max_smem_optin())
                ch = t;
        if (ch) {
            const int chunks = C / ch, S = pool_splits(B, chunks);
            const size_t smem = (((size_t)HW * ch * 4 + 15) & ~(size_t)15) + pool_fwd_stage_bytes(ch);
            const unsigned grid = (unsigned)(B * chunks * S);
            const bool mm = variant == CIM_ROI_POOL_MMCV;
#define POOL_FWD(CHV, MM)                                                                                          \
    do {                                                                                                           \
        cudaFuncSetAttribute(roi_pool_fwd_tile_kernel<CHV, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                             (int)smem);                                                                           \
        roi_pool_fwd_tile_kernel<CHV, MM><<<grid, fwd_warps(CHV) * 32, smem, st>>>(feat, rois, out, argmax, B, C, H, W, \
                                                                              K, scale, S);                        \
    } while (0)
            if (ch == 32) { if (mm) POOL_FWD(32, true); else POOL_FWD(32, false); }
            else if (ch == 16) { if (mm) POOL_FWD(16, true); else POOL_FWD(16, false); }
            else { if (mm) POOL_FWD(8, true); else POOL_FWD(8, false); }
#undef POOL_FWD
            return cim_launch_status();
        }
    }
    const int per_roi = C * oh * ow;
    dim3 grid((unsigned)K, (unsigned)min(64, (per_roi + 255) / 256));
    if (variant == CIM_ROI_POOL_MMCV)
        roi_pool_fwd_kernel<true><<<grid, 256, 0, st>>>(feat, rois, out, argmax, B, C, H, W, K, oh, ow, scale);
    else
        roi_pool_fwd_kernel<false><<<grid, 256, 0, st>>>(feat, rois, out, argmax, B, C, H, W, K, oh, ow, scale);
    return cim_launch_status();
}

CIM_API int cim_roi_pool_bwd(const float *grad_out, const int32_t *argmax, const float *rois, float *grad_feat,
                             int B, int C, int H, int W, int K, int oh, int ow, cim_stream_t stream) {
    if (!grad_out || !argmax || !grad_feat || (K > 0 && !rois)) return CIM_ERR_ARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K < 0 || oh <= 0 || ow <= 0) return CIM_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool simple = (cim_get_debug_flags() & CIM_DBG_ROI_POOL_SIMPLE) != 0;
    const int HW = H * W;
    if (!simple && K > 0 && oh == PO && ow == PO && HW < 32768 && cim_aligned(grad_out, 16) && cim_aligned(argmax, 16)) {
        int ch = 0;
        for (int t : {16, 8})
            if (!ch && C % t == 0 && (size_t)HW * t * 4 + pool_bwd_other_bytes(t) + 1024 <= (size_t)cim_max_smem_optin())
                ch = t;
        if (ch) {
            const int chunks = C / ch, S = pool_splits(B, chunks);
            const size_t smem = (((size_t)HW * ch * 4 + 15) & ~(size_t)15) + pool_bwd_other_bytes(ch);
            const unsigned grid = (unsigned)(B * chunks * S);
            if (S > 1) cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * HW, st);
#define POOL_BWD(CHV)                                                                                              \
    do {                                                                                                           \
        cudaFuncSetAttribute(roi_pool_bwd_tile_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        roi_pool_bwd_tile_kernel<CHV><<<grid, BWD_WARPS * 32, smem, st>>>(grad_out, argmax, rois, grad_feat, B, C, \
                                                                          H, W, K, S);                             \
    } while (0)
            if (ch == 16) POOL_BWD(16);
            else POOL_BWD(8);
#undef POOL_BWD
            return cim_launch_status();
        }
    }
    cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W, st);
    if (K == 0) return cim_launch_status();
    const int per_roi = C * oh * ow;
    dim3 grid((unsigned)K, (unsigned)min(64, (per_roi + 255) / 256));
    roi_pool_bwd_kernel<<<grid, 256, 0, st>>>(grad_out, argmax, rois, grad_feat, B, C, H, W, K, oh, ow);
    return cim_launch_status();
}
