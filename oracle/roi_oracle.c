/*
 * oracle/roi_oracle.c -- CPU restatement of the reference's ROI operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in cim_b200/ may link, load or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker / the timed CPU baseline.
 *
 * What it restates (scalar fp32 arithmetic, same operation order per sample):
 *   - RoIAlign forward   : lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:16-63
 *                          (bilinear + border rules) and :65-121 (bins, adaptive
 *                          sampling grid, averaging)
 *   - RoIAlign backward  : same file :150-193 (gradient weights), :195-270 (scatter)
 *   - RoIPool forward    : lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93
 *   - RoIPool backward   : same file :128-203 (here as a scatter from argmax, which
 *                          sums the same terms)
 * plus the one thing the vendored kernel does not have and the live op
 * (mmcv.ops.RoIAlign, lib/ops/__init__.py:6, called at
 * lib/modeling/model_builder.py:230-231) does: `aligned` -- subtract half a pixel
 * after scaling and do not clamp the ROI extent to >= 1.  mmcv itself is not in
 * /root/reference (un-vendored, un-pinned dependency), so for aligned=1 parity is
 * pinned against torchvision.ops.roi_align(aligned=True) (same Detectron2-derived
 * algorithm) in tests/test_oracle_roi.py; for aligned=0 it is additionally pinned
 * against the vendored CUDA kernel itself compiled into oracle/_ref (GPU tests).
 *
 * Conventions: feat NCHW fp32 contiguous, rois [K,5] = (batch_idx,x1,y1,x2,y2),
 * out [K,C,oh,ow].  RoIPool argmax is the index inside the (H*W) plane, -1 for an
 * empty bin (the vendored kernel stores a global index; only fwd->bwd consistency
 * matters).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>

/* weights + corner indices of one bilinear sample; returns 0 when the sample is
 * outside [-1,H] x [-1,W] and contributes nothing (kernel.cu:19-22, :158-163) */
static int sample_corners(int H, int W, float y, float x,
                          int *y0, int *y1, int *x0, int *x1,
                          float *w00, float *w01, float *w10, float *w11)
{
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
    float ly = y - (float)yl, lx = x - (float)xl;
    float hy = 1.f - ly, hx = 1.f - lx;
    *y0 = yl; *y1 = yh; *x0 = xl; *x1 = xh;
    *w00 = hy * hx; *w01 = hy * lx; *w10 = ly * hx; *w11 = ly * lx;
    return 1;
}

typedef struct {
    int b, gh, gw;
    float y_start, x_start, bin_h, bin_w, count;
} roi_geom;

static void roi_geometry(const float *roi, float scale, int oh, int ow,
                         int sampling_ratio, int aligned, roi_geom *g)
{
    float off = aligned ? 0.5f : 0.f;
    g->b = (int)roi[0];
    float x1 = roi[1] * scale - off, y1 = roi[2] * scale - off;
    float x2 = roi[3] * scale - off, y2 = roi[4] * scale - off;
    float rw = x2 - x1, rh = y2 - y1;
    if (!aligned) {                      /* kernel.cu:85-86: malformed ROIs -> 1x1 */
        rw = fmaxf(rw, 1.f);
        rh = fmaxf(rh, 1.f);
    }
    g->x_start = x1; g->y_start = y1;
    g->bin_h = rh / (float)oh;
    g->bin_w = rw / (float)ow;
    g->gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)oh);
    g->gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)ow);
    float c = (float)(g->gh * g->gw);
    g->count = c > 1.f ? c : 1.f;        /* mmcv/torchvision: max(gh*gw,1) */
}

void oracle_roi_align_fwd(const float *feat, const float *rois, float *out,
                          int B, int C, int H, int W, int K, int oh, int ow,
                          float scale, int sampling_ratio, int aligned)
{
    (void)B;
#pragma omp parallel for schedule(dynamic, 4)
    for (int k = 0; k < K; ++k) {
        roi_geom g;
        roi_geometry(rois + 5 * k, scale, oh, ow, sampling_ratio, aligned, &g);
        for (int c = 0; c < C; ++c) {
            const float *plane = feat + ((size_t)g.b * C + c) * H * W;
            float *o = out + ((size_t)k * C + c) * oh * ow;
            for (int ph = 0; ph < oh; ++ph)
                for (int pw = 0; pw < ow; ++pw) {
                    float acc = 0.f;
                    for (int iy = 0; iy < g.gh; ++iy) {
                        float y = g.y_start + ph * g.bin_h +
                                  ((float)iy + .5f) * g.bin_h / (float)g.gh;
                        for (int ix = 0; ix < g.gw; ++ix) {
                            float x = g.x_start + pw * g.bin_w +
                                      ((float)ix + .5f) * g.bin_w / (float)g.gw;
                            int y0, y1, x0, x1; float a, b2, c2, d;
                            if (!sample_corners(H, W, y, x, &y0, &y1, &x0, &x1,
                                                &a, &b2, &c2, &d))
                                continue;
                            acc += a * plane[y0 * W + x0] + b2 * plane[y0 * W + x1] +
                                   c2 * plane[y1 * W + x0] + d * plane[y1 * W + x1];
                        }
                    }
                    o[ph * ow + pw] = acc / g.count;
                }
        }
    }
}

void oracle_roi_align_bwd(const float *grad_out, const float *rois, float *grad_feat,
                          int B, int C, int H, int W, int K, int oh, int ow,
                          float scale, int sampling_ratio, int aligned)
{
    memset(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W);
    /* channel-outer: every thread owns whole planes, ROIs still accumulate in index order */
#pragma omp parallel for schedule(dynamic, 4)
    for (int c = 0; c < C; ++c) {
        for (int k = 0; k < K; ++k) {
            roi_geom g;
            roi_geometry(rois + 5 * k, scale, oh, ow, sampling_ratio, aligned, &g);
            float *plane = grad_feat + ((size_t)g.b * C + c) * H * W;
            const float *go = grad_out + ((size_t)k * C + c) * oh * ow;
            for (int ph = 0; ph < oh; ++ph)
                for (int pw = 0; pw < ow; ++pw) {
                    float top = go[ph * ow + pw];
                    for (int iy = 0; iy < g.gh; ++iy) {
                        float y = g.y_start + ph * g.bin_h +
                                  ((float)iy + .5f) * g.bin_h / (float)g.gh;
                        for (int ix = 0; ix < g.gw; ++ix) {
                            float x = g.x_start + pw * g.bin_w +
                                      ((float)ix + .5f) * g.bin_w / (float)g.gw;
                            int y0, y1, x0, x1; float a, b2, c2, d;
                            if (!sample_corners(H, W, y, x, &y0, &y1, &x0, &x1,
                                                &a, &b2, &c2, &d))
                                continue;
                            plane[y0 * W + x0] += top * a / g.count;
                            plane[y0 * W + x1] += top * b2 / g.count;
                            plane[y1 * W + x0] += top * c2 / g.count;
                            plane[y1 * W + x1] += top * d / g.count;
                        }
                    }
                }
        }
    }
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void oracle_roi_pool_fwd(const float *feat, const float *rois, float *out, int32_t *argmax,
                         int B, int C, int H, int W, int K, int oh, int ow, float scale)
{
    (void)B;
    for (int k = 0; k < K; ++k) {
        const float *r = rois + 5 * k;
        int b = (int)r[0];
        int x1 = (int)roundf(r[1] * scale), y1 = (int)roundf(r[2] * scale);
        int x2 = (int)roundf(r[3] * scale), y2 = (int)roundf(r[4] * scale);
        int rw = x2 - x1 + 1 > 1 ? x2 - x1 + 1 : 1;
        int rh = y2 - y1 + 1 > 1 ? y2 - y1 + 1 : 1;
        float bh = (float)rh / (float)oh, bw = (float)rw / (float)ow;
        for (int c = 0; c < C; ++c) {
            const float *plane = feat + ((size_t)b * C + c) * H * W;
            size_t ob = ((size_t)k * C + c) * oh * ow;
            for (int ph = 0; ph < oh; ++ph)
                for (int pw = 0; pw < ow; ++pw) {
                    int hs = clampi((int)floorf((float)ph * bh) + y1, 0, H);
                    int he = clampi((int)ceilf((float)(ph + 1) * bh) + y1, 0, H);
                    int ws = clampi((int)floorf((float)pw * bw) + x1, 0, W);
                    int we = clampi((int)ceilf((float)(pw + 1) * bw) + x1, 0, W);
                    int empty = (he <= hs) || (we <= ws);
                    float best = empty ? 0.f : -FLT_MAX;
                    int besti = -1;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w)
                            if (plane[h * W + w] > best) { best = plane[h * W + w]; besti = h * W + w; }
                    out[ob + ph * ow + pw] = best;
                    if (argmax) argmax[ob + ph * ow + pw] = besti;
                }
        }
    }
}

/* RoIPool forward with the bins of mmcv 1.x (the module lib/ops/__init__.py:6 imports; mmcv-full is an un-vendored,
 * un-pinned dependency that is absent here, so this restates its PUBLISHED kernel -- mmcv/ops/csrc/common/cuda/
 * roi_pool_cuda_kernel.cuh, roi_pool_forward_cuda_kernel -- and is "parity unpinned"): float ROI corners
 * x1 * s, y1 * s, (x2 + 1) * s, (y2 + 1) * s; a ROI with w <= 0 or h <= 0 is skipped (output stays 0; argmax is
 * reported as -1 here, i.e. no gradient); bin edges floor(p * bin + start) .. ceil((p + 1) * bin + start) clipped to
 * the map; empty bin -> 0 / -1; strict > keeps the first maximum in row-major order. */
void oracle_roi_pool_fwd_mmcv(const float *feat, const float *rois, float *out, int32_t *argmax,
                              int B, int C, int H, int W, int K, int oh, int ow, float scale)
{
    (void)B;
    for (int k = 0; k < K; ++k) {
        const float *r = rois + 5 * k;
        int b = (int)r[0];
        float x1 = r[1] * scale, y1 = r[2] * scale;
        float x2 = (r[3] + 1.f) * scale, y2 = (r[4] + 1.f) * scale;
        float rw = x2 - x1, rh = y2 - y1;
        int skip = rw <= 0.f || rh <= 0.f;
        float bw = rw / (float)ow, bh = rh / (float)oh;
        for (int c = 0; c < C; ++c) {
            const float *plane = feat + ((size_t)b * C + c) * H * W;
            size_t ob = ((size_t)k * C + c) * oh * ow;
            for (int ph = 0; ph < oh; ++ph)
                for (int pw = 0; pw < ow; ++pw) {
                    float best = 0.f;
                    int besti = -1;
                    if (!skip) {
                        int ws = clampi((int)floorf((float)pw * bw + x1), 0, W);
                        int hs = clampi((int)floorf((float)ph * bh + y1), 0, H);
                        int we = clampi((int)ceilf((float)(pw + 1) * bw + x1), 0, W);
                        int he = clampi((int)ceilf((float)(ph + 1) * bh + y1), 0, H);
                        int empty = (he <= hs) || (we <= ws);
                        best = empty ? 0.f : -FLT_MAX;
                        for (int h = hs; h < he; ++h)
                            for (int w = ws; w < we; ++w)
                                if (plane[h * W + w] > best) { best = plane[h * W + w]; besti = h * W + w; }
                    }
                    out[ob + ph * ow + pw] = best;
                    if (argmax) argmax[ob + ph * ow + pw] = besti;
                }
        }
    }
}

void oracle_roi_pool_bwd(const float *grad_out, const int32_t *argmax, const float *rois,
                         float *grad_feat, int B, int C, int H, int W, int K, int oh, int ow)
{
    memset(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W);
    for (int k = 0; k < K; ++k) {
        int b = (int)rois[5 * k];
        for (int c = 0; c < C; ++c) {
            float *plane = grad_feat + ((size_t)b * C + c) * H * W;
            size_t ob = ((size_t)k * C + c) * oh * ow;
            for (int p = 0; p < oh * ow; ++p) {
                int a = argmax[ob + p];
                if (a >= 0) plane[a] += grad_out[ob + p];
            }
        }
    }
}

/* Pairwise mask overlap on byte masks (0 / non-zero), integer counts only:
 * inter[i,j] = |m_i & m_j|, area[i] = |m_i|.  lib/utils/mask_utils.py:13-17,27-31
 * take `bitwise_and(...).sum()` / `bitwise_or(...).sum()` per pair; the union is
 * area_i + area_j - inter.  Used to time the CPU baseline at sizes where the
 * literal Python double loop of the reference would take hours. */
void oracle_mask_counts(const uint8_t *masks, int N, int64_t HW, int32_t *inter, int32_t *area)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < N; ++i) {
        const uint8_t *a = masks + (size_t)i * HW;
        for (int j = i; j < N; ++j) {
            const uint8_t *b = masks + (size_t)j * HW;
            int32_t s = 0;
            for (int64_t p = 0; p < HW; ++p) s += (a[p] != 0) & (b[p] != 0);
            inter[(size_t)i * N + j] = s;
            inter[(size_t)j * N + i] = s;
        }
        area[i] = inter[(size_t)i * N + i];
    }
}
