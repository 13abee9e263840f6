// roi_align.cu -- RoIAlign forward / backward for sm_100a.
//
// Replaces mmcv.ops.RoIAlign at lib/modeling/model_builder.py:230-231 of the reference
// (arithmetic: lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:16-121 fwd,
// :150-270 bwd, plus mmcv's `aligned` half-pixel shift).
//
// Design (DESIGN.md "RoIAlign"):
//   1. roi_prep_kernel turns every ROI into a descriptor: per output bin and axis the first
//      feature row/column touched, the number touched and the COLLAPSED bilinear weights
//      (all samples of the bin folded into <= 8 taps per axis; the 2-D weight of a tap is the
//      product of its row and column weight, already divided by the sample count).  It also
//      checks that rois are grouped by image and records each image's ROI range.
//   2. Tile kernels: a CTA keeps the [32 channels x H x W] feature (or gradient) tile of one
//      image in shared memory (pitch odd => lane = channel is bank-conflict free) and sweeps
//      the image's ROIs.
//        forward : persistent, one CTA per SM, (roi, chunk) units split evenly; a warp owns one
//                  ROI at a time, its descriptor is prefetched into smem with cp.async (LDGSTS)
//                  one ROI ahead; ROW SWEEP: each feature value of the ROI window is loaded once,
//                  the 7 column sums of a row are added into the output rows whose window holds
//                  that row (49 accumulators in registers); the 32 x 49 outputs are staged in
//                  smem and leave with one 6272-byte cp.async.bulk (UBLKCP) per (roi, 32 ch).
//        backward: gradients + descriptors stream through a 3-slot mbarrier ring of 4 ROIs each
//                  (cp.async.bulk, no CTA-wide barrier); warp w of 16 owns the tile row pairs
//                  (y >> 1) % 16 == w, so every tile address has exactly one writer and ROIs are
//                  applied in index order: no atomics, bit-reproducible.
//   3. Generic kernels (one thread per output element, sample by sample) take whatever the
//      tile kernels cannot: other output sizes, feature maps too large for shared memory,
//      bins wider than 8 taps, ROIs not grouped by image.
#include "common.cuh"

namespace {

constexpr int PH = 7, PW = 7, NBIN = 49;
constexpr int MAXT = 8;             // max collapsed taps per bin and axis in a descriptor
constexpr int DESC_WORDS = 160;     // 640 B per ROI
enum {
    D_B = 0, D_FLAGY = 1, D_FLAGX = 2, D_TX = 3, D_Y0 = 4, D_Y1 = 5, D_XINC = 6, D_OWN = 7,
    D_YLO = 8, D_YN = 16, D_XLO = 24, D_XN = 32, D_WY = 40, D_WX = 96
};
constexpr int WS_HDR_BYTES = 256;   // word 0: "rois not grouped by image" flag
constexpr int CH = 32;              // channels per tile (= lanes)
constexpr int STAGE_FLOATS = CH * NBIN;          // 1568 floats = 6272 B

struct RoiWs {
    int *hdr;          // [64]
    int *img_start;    // [B+1]
    int *desc;         // [K][DESC_WORDS]
};

__host__ __device__ inline size_t ws_img_off() { return WS_HDR_BYTES; }
__host__ __device__ inline size_t ws_desc_off(int B) {
    size_t o = WS_HDR_BYTES + sizeof(int) * (size_t)(B + 1);
    return (o + 15) & ~(size_t)15;
}

// ------------------------------------------------------------------------------------ prep
// One sample coordinate, evaluated exactly like the reference expression
//   start + p * bin + (s + .5f) * bin / grid          (roi_align_kernel.cu:103-108)
// with every operation individually rounded (no FMA contraction), so that the integer
// decisions (floor, border tests) agree with a plain C evaluation.
__device__ __forceinline__ float sample_coord(float start, float bin, int p, int s, int grid) {
    float a = __fadd_rn(start, __fmul_rn((float)p, bin));
    float b = __fdiv_rn(__fmul_rn(__fadd_rn((float)s, .5f), bin), (float)grid);
    return __fadd_rn(a, b);
}

__global__ void roi_prep_kernel(const float *__restrict__ rois, int K, int B, int H, int W,
                                int oh, int ow, float scale, int sampling_ratio, int aligned,
                                int *__restrict__ hdr, int *__restrict__ img_start,
                                int *__restrict__ desc) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int k = gid >> 1, axis = gid & 1;          // axis 0 = y (rows), 1 = x (columns)
    if (k >= K) return;
    const float *r = rois + 5 * (size_t)k;
    int *d = desc + (size_t)k * DESC_WORDS;
    int b = (int)r[0];

    if (axis == 0) {
        d[D_B] = b;
        // image ranges + grouping check (rois of one image must be contiguous, ascending)
        int bprev = k > 0 ? (int)rois[5 * (size_t)(k - 1)] : -1;
        if (k > 0 && b < bprev) atomicOr(hdr, 1);
        int lo = max(bprev + 1, 0), hi = min(b, B);
        if (k == 0) lo = 0;
        for (int j = lo; j <= hi; ++j) img_start[j] = k;
        if (k == K - 1)
            for (int j = max(b + 1, 0); j <= B; ++j) img_start[j] = K;
    }

    const int nb = axis == 0 ? oh : ow;       // bins along this axis (7 on the tile path)
    const int L = axis == 0 ? H : W;
    const float off = aligned ? .5f : 0.f;
    float c1 = __fsub_rn(__fmul_rn(r[1 + (axis == 0 ? 1 : 0)], scale), off);
    float c2 = __fsub_rn(__fmul_rn(r[3 + (axis == 0 ? 1 : 0)], scale), off);
    float ext = __fsub_rn(c2, c1);
    if (!aligned) ext = fmaxf(ext, 1.f);
    float bin = __fdiv_rn(ext, (float)nb);
    int grid = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(ext, (float)nb));

    int flag = 0;
    if (nb != 7 || grid > 64 || b < 0 || b >= B) flag = 1;

    int lo_[7], n_[7];
    float w_[7][MAXT];
    int nmax = 0;
    if (!flag) {
        for (int p = 0; p < 7; ++p) {
            int lo0 = -1, n = 0;
            float w[MAXT];
#pragma unroll
            for (int i = 0; i < MAXT; ++i) w[i] = 0.f;
            for (int s = 0; s < grid; ++s) {
                float t = sample_coord(c1, bin, p, s, grid);
                if (t < -1.f || t > (float)L) continue;
                if (t <= 0.f) t = 0.f;
                int l0 = (int)t, h0;
                if (l0 >= L - 1) { h0 = l0 = L - 1; t = (float)l0; } else { h0 = l0 + 1; }
                float fl = t - (float)l0, fh = 1.f - fl;
                if (lo0 < 0) lo0 = l0;
                int i0 = l0 - lo0, i1 = h0 - lo0;
                if (i1 >= MAXT) { flag = 1; break; }
                w[i0] += fh;
                w[i1] += fl;
                n = max(n, i1 + 1);
            }
            if (flag) break;
            lo_[p] = lo0 < 0 ? 0 : lo0;
            n_[p] = n;
            nmax = max(nmax, n);
            float g = (float)(grid > 0 ? grid : 1);
#pragma unroll
            for (int i = 0; i < MAXT; ++i) w_[p][i] = w[i] / g;
        }
    }

    if (axis == 0) {
        d[D_FLAGY] = flag;
        int y0 = 1 << 30, y1 = 0;
        for (int p = 0; p < 7; ++p) {
            int lo = flag ? 0 : lo_[p], n = flag ? 0 : n_[p];
            d[D_YLO + p] = lo;
            d[D_YN + p] = n;
            if (n > 0) { y0 = min(y0, lo); y1 = max(y1, lo + n); }
            for (int i = 0; i < MAXT; ++i)
                reinterpret_cast<float *>(d)[D_WY + p * MAXT + i] = flag ? 0.f : w_[p][i];
        }
        if (y1 == 0) y0 = 0;
        d[D_Y0] = y0;
        d[D_Y1] = y1;
        // backward: which of the 16 warps own a row of [y0, y1)?  warp w owns the row pairs (y >> 1) % 16 == w
        int own = 0;
        for (int y = y0; y < y1 && own != 0xffff; ++y) own |= 1 << ((y >> 1) & 15);
        d[D_OWN] = flag ? 0 : own;
        d[D_YLO + 7] = 0;
        d[D_YN + 7] = 0;
    } else {
        // tap class T in {2,3,4,6,8}: the x loops of the tile kernels are unrolled T times.
        int T = nmax <= 2 ? 2 : nmax <= 3 ? 3 : nmax <= 4 ? 4 : nmax <= 6 ? 6 : 8;
        if (W < T) flag = 1;
        d[D_FLAGX] = flag;
        d[D_TX] = T;
        int inc = 1, prev = -1;
        for (int p = 0; p < 7; ++p) {
            int lo = flag ? 0 : lo_[p];
            // shift the window left so that lo + T - 1 stays inside the row; the padding
            // taps get weight 0 but must still address valid shared memory
            int lo2 = min(lo, W - T);
            if (lo2 < 0) lo2 = 0;
            int sh = lo - lo2;
            d[D_XLO + p] = lo2;
            d[D_XN + p] = flag ? 0 : n_[p];
            for (int i = 0; i < MAXT; ++i) {
                int src = i - sh;
                float v = (!flag && src >= 0 && src < MAXT) ? w_[p][src] : 0.f;
                reinterpret_cast<float *>(d)[D_WX + p * MAXT + i] = v;
            }
            if (lo2 <= prev) inc = 0;
            prev = lo2;
        }
        d[D_XINC] = inc;      // 1: window starts strictly increase -> taps of one l never alias
        d[D_XLO + 7] = 0;
        d[D_XN + 7] = 0;
    }
}

// ------------------------------------------------------------------------------ tile: forward
__device__ __forceinline__ void load_tile(float *tile, const float *__restrict__ src, int HW,
                                          int pitch, int tid, int nthr) {
    if ((HW & 3) == 0) {
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        int n4 = CH * HW / 4;
        for (int e = tid; e < n4; e += nthr) {
            float4 v = __ldg(s4 + e);
            int idx = e * 4, c = idx / HW, p = idx - c * HW;
            float *dst = tile + c * pitch + p;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    } else {
        for (int e = tid; e < CH * HW; e += nthr) {
            int c = e / HW, p = e - c * HW;
            tile[c * pitch + p] = __ldg(src + e);
        }
    }
}

// 16-byte async copy global -> shared (LDGSTS), per-thread completion groups
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One ROI x 32 channels (lane = channel).  Row sweep: every feature value of the ROI's window is
// loaded from the smem tile once; the 7 column sums of a row are formed with the register-resident
// column weights and then added into the (at most 2-3) output rows whose window contains this
// feature row.  All control flow depends only on the descriptor, i.e. is warp-uniform.
template <int T>
__device__ __forceinline__ void fwd_rows(const float *__restrict__ tile_c, int W, const int *__restrict__ d,
                                         float *__restrict__ stage_c, float *stage_w, int lane) {
    float wx[PW][T];
    int xo[PW], ylo[PH], yhi[PH];
    const float *dwx = reinterpret_cast<const float *>(d + D_WX);
    const float *dwy = reinterpret_cast<const float *>(d + D_WY);
    {
        int4 a = *reinterpret_cast<const int4 *>(d + D_XLO), b = *reinterpret_cast<const int4 *>(d + D_XLO + 4);
        xo[0] = a.x; xo[1] = a.y; xo[2] = a.z; xo[3] = a.w; xo[4] = b.x; xo[5] = b.y; xo[6] = b.z;
        a = *reinterpret_cast<const int4 *>(d + D_YLO); b = *reinterpret_cast<const int4 *>(d + D_YLO + 4);
        ylo[0] = a.x; ylo[1] = a.y; ylo[2] = a.z; ylo[3] = a.w; ylo[4] = b.x; ylo[5] = b.y; ylo[6] = b.z;
        a = *reinterpret_cast<const int4 *>(d + D_YN); b = *reinterpret_cast<const int4 *>(d + D_YN + 4);
        yhi[0] = ylo[0] + a.x; yhi[1] = ylo[1] + a.y; yhi[2] = ylo[2] + a.z; yhi[3] = ylo[3] + a.w;
        yhi[4] = ylo[4] + b.x; yhi[5] = ylo[5] + b.y; yhi[6] = ylo[6] + b.z;
    }
#pragma unroll
    for (int pw = 0; pw < PW; ++pw) {
        float4 a = *reinterpret_cast<const float4 *>(dwx + pw * MAXT);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (T > 4) b = *reinterpret_cast<const float4 *>(dwx + pw * MAXT + 4);
        float t8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int l = 0; l < T; ++l) wx[pw][l] = t8[l];
    }
    float acc[PH][PW];
#pragma unroll
    for (int ph = 0; ph < PH; ++ph)
#pragma unroll
        for (int pw = 0; pw < PW; ++pw) acc[ph][pw] = 0.f;

    const int y0 = d[D_Y0], y1 = d[D_Y1];
#pragma unroll 1
    for (int y = y0; y < y1; ++y) {
        const float *row = tile_c + y * W;
        float r[PW];
#pragma unroll
        for (int pw = 0; pw < PW; ++pw) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < T; ++l) t = fmaf(wx[pw][l], row[xo[pw] + l], t);
            r[pw] = t;
        }
#pragma unroll
        for (int ph = 0; ph < PH; ++ph) {
            if (y >= ylo[ph] && y < yhi[ph]) {
                const float w = dwy[ph * MAXT + (y - ylo[ph])];
#pragma unroll
                for (int pw = 0; pw < PW; ++pw) acc[ph][pw] = fmaf(w, r[pw], acc[ph][pw]);
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();            // the previous bulk store has drained this stage
    __syncwarp();
#pragma unroll
    for (int ph = 0; ph < PH; ++ph)
#pragma unroll
        for (int pw = 0; pw < PW; ++pw) stage_c[ph * PW + pw] = acc[ph][pw];
    (void)stage_w;
}

__global__ void __launch_bounds__(384, 1)
roi_align_fwd_tile_kernel(const float *__restrict__ feat, const int *__restrict__ hdr,
                          const int *__restrict__ img_start, const int *__restrict__ descs,
                          float *__restrict__ out, int B, int C, int H, int W, int pitch) {
    extern __shared__ __align__(128) float smem[];
    const int nw = blockDim.x >> 5;
    float *tile = smem;
    float *stage = smem + (size_t)CH * pitch;                                    // [nw][1568]
    int *dslots = reinterpret_cast<int *>(stage + (size_t)nw * STAGE_FLOATS);   // [nw][2][160]
    if (__ldg(hdr) != 0) return;              // rois not grouped by image: generic kernel runs

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = H * W, nchunks = C / CH;
    const int s0 = __ldg(img_start), sB = __ldg(img_start + B);
    const long long U = (long long)nchunks * (sB - s0);     // (roi, chunk) units
    long long u = U * blockIdx.x / gridDim.x;
    const long long u_end = U * (blockIdx.x + 1) / gridDim.x;
    float *stage_w = stage + warp * STAGE_FLOATS;
    float *stage_c = stage_w + lane * NBIN;
    int *my_slots = dslots + warp * 2 * DESC_WORDS;

    // descriptor prefetch: 640 B = 40 x 16 B per ROI, lanes 0..31 + lanes 0..7 again
    auto prefetch = [&](int roi, int slot) {
        const int *src = descs + (size_t)roi * DESC_WORDS;
        int *dst = my_slots + slot * DESC_WORDS;
        cp_async16(dst + lane * 4, src + lane * 4);
        if (lane < 8) cp_async16(dst + (32 + lane) * 4, src + (32 + lane) * 4);
        cp_async_commit();
    };

    int b = 0;
    while (u < u_end) {
        // locate the tile (image b, channel chunk ch) that unit u falls into
        while (b < B && (long long)nchunks * (__ldg(img_start + b + 1) - s0) <= u) ++b;
        const int is = __ldg(img_start + b), nroi = __ldg(img_start + b + 1) - is;
        const long long base = (long long)nchunks * (is - s0);
        const int ch = (int)((u - base) / nroi);
        const int r0 = (int)((u - base) - (long long)ch * nroi);
        const int r1 = (int)min((long long)nroi, r0 + (u_end - u));
        const int c0 = ch * CH;

        int r = r0 + warp, slot = 0;
        if (r < r1) prefetch(is + r, 0);       // overlaps the tile load below
        __syncthreads();                       // everyone is done with the previous tile
        load_tile(tile, feat + ((size_t)b * C + c0) * HW, HW, pitch, tid, blockDim.x);
        __syncthreads();

        const float *tile_c = tile + lane * pitch;
        while (r < r1) {
            const int rn = r + nw;
            if (rn < r1) { prefetch(is + rn, slot ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            __syncwarp();
            const int *d = my_slots + slot * DESC_WORDS;
            const int roi = is + r;
            if ((d[D_FLAGY] | d[D_FLAGX]) == 0) {
                switch (d[D_TX]) {
                    case 2: fwd_rows<2>(tile_c, W, d, stage_c, stage_w, lane); break;
                    case 3: fwd_rows<3>(tile_c, W, d, stage_c, stage_w, lane); break;
                    case 4: fwd_rows<4>(tile_c, W, d, stage_c, stage_w, lane); break;
                    case 6: fwd_rows<6>(tile_c, W, d, stage_c, stage_w, lane); break;
                    default: fwd_rows<8>(tile_c, W, d, stage_c, stage_w, lane); break;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    bulk_s2g(out + ((size_t)roi * C + c0) * NBIN, stage_w, STAGE_FLOATS * 4);
                    bulk_commit();
                }
            }
            __syncwarp();                      // slot is re-filled two iterations from now
            slot ^= 1;
            r = rn;
        }
        u += r1 - r0;
    }
    if (lane == 0) bulk_wait_read<0>();
}

// ----------------------------------------------------------------------------- tile: backward
constexpr int NWB = 16;             // consumer warps of the backward CTA; warp w owns row pairs
                                    // {y : (y >> 1) % 16 == w}  (D_OWN is computed for exactly this map)
constexpr int NBR = 4;              // ROIs per ring slot: one full/empty barrier round per 4 ROIs
constexpr int NS = 3;               // ring slots
constexpr int SLOT_FLOATS = NBR * (STAGE_FLOATS + DESC_WORDS);

template <int T, bool XINC>
__device__ __forceinline__ void bwd_rows(float *__restrict__ tile_c, int W, const int *d,
                                         const float *__restrict__ g, int warp, int y0, int y1) {
    float wx[PW][T];
    int xo[PW], ylo[PH], yn[PH];
    const float *dwx = reinterpret_cast<const float *>(d + D_WX);
    const float *dwy = reinterpret_cast<const float *>(d + D_WY);
    {
        int4 a = *reinterpret_cast<const int4 *>(d + D_XLO), b = *reinterpret_cast<const int4 *>(d + D_XLO + 4);
        xo[0] = a.x; xo[1] = a.y; xo[2] = a.z; xo[3] = a.w; xo[4] = b.x; xo[5] = b.y; xo[6] = b.z;
        a = *reinterpret_cast<const int4 *>(d + D_YLO); b = *reinterpret_cast<const int4 *>(d + D_YLO + 4);
        ylo[0] = a.x; ylo[1] = a.y; ylo[2] = a.z; ylo[3] = a.w; ylo[4] = b.x; ylo[5] = b.y; ylo[6] = b.z;
        a = *reinterpret_cast<const int4 *>(d + D_YN); b = *reinterpret_cast<const int4 *>(d + D_YN + 4);
        yn[0] = a.x; yn[1] = a.y; yn[2] = a.z; yn[3] = a.w; yn[4] = b.x; yn[5] = b.y; yn[6] = b.z;
    }
#pragma unroll
    for (int pw = 0; pw < PW; ++pw) {
        float4 a = *reinterpret_cast<const float4 *>(dwx + pw * MAXT);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (T > 4) b = *reinterpret_cast<const float4 *>(dwx + pw * MAXT + 4);
        float t8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int l = 0; l < T; ++l) wx[pw][l] = t8[l];
    }
    // the row pairs this warp owns: 2*warp + 32*m, clipped to the ROI's rows [y0, y1)
    for (int yb = 2 * warp; yb < y1; yb += 2 * NWB) {
#pragma unroll 1
        for (int y = max(yb, y0); y < min(yb + 2, y1); ++y) {
            float r[PW];
#pragma unroll
            for (int pw = 0; pw < PW; ++pw) r[pw] = 0.f;
#pragma unroll
            for (int ph = 0; ph < PH; ++ph) {
                int dd = y - ylo[ph];
                if (dd >= 0 && dd < yn[ph]) {
                    float w = dwy[ph * MAXT + dd];
#pragma unroll
                    for (int pw = 0; pw < PW; ++pw) r[pw] = fmaf(w, g[ph * PW + pw], r[pw]);
                }
            }
            float *row = tile_c + y * W;
            if (XINC) {
                // window starts strictly increase: the 7 taps of one l hit 7 distinct columns,
                // so they can be loaded, updated and stored as a group
#pragma unroll
                for (int l = 0; l < T; ++l) {
                    float v[PW];
#pragma unroll
                    for (int pw = 0; pw < PW; ++pw) v[pw] = row[xo[pw] + l];
#pragma unroll
                    for (int pw = 0; pw < PW; ++pw) row[xo[pw] + l] = fmaf(wx[pw][l], r[pw], v[pw]);
                }
            } else {
#pragma unroll
                for (int pw = 0; pw < PW; ++pw)
#pragma unroll
                    for (int l = 0; l < T; ++l) {
                        volatile float *p = row + xo[pw] + l;
                        *p = fmaf(wx[pw][l], r[pw], *p);
                    }
            }
        }
    }
}

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 4 floats added to global memory with one reduction (RED.E.ADD.F32x4)
__device__ __forceinline__ void red_add4(float *gptr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Persistent: one CTA per SM.  The (roi, channel-chunk) units of the whole batch are split evenly and
// contiguously over the CTAs (like the forward), so a CTA works through 1-3 tile SEGMENTS: zero the
// smem tile, apply the ROIs of the segment, add the tile to grad_feat with red.global.add.  grad_feat
// is zeroed beforehand and every tile is touched by at most TWO CTAs (a CTA's share is larger than one
// tile), i.e. each element is 0 + a (+ b): exact and independent of the order -> still bit-reproducible.
__global__ void __launch_bounds__(NWB * 32, 1)
roi_align_bwd_tile_kernel(const float *__restrict__ grad_out, const int *__restrict__ hdr,
                          const int *__restrict__ img_start, const int *__restrict__ descs,
                          float *__restrict__ grad_feat, int B, int C, int H, int W, int pitch) {
    extern __shared__ __align__(128) float smem[];
    float *tile = smem;
    float *ring = smem + (size_t)CH * pitch;                 // [NS][NBR x 1568 grads | NBR x 160 desc words]
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)NS * SLOT_FLOATS);
    uint64_t *empty = full + NS;
    if (__ldg(hdr) != 0) return;              // rois not grouped by image: the generic kernel does it all

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HW = H * W, nchunks = C / CH;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWB); }
        fence_mbar_init();
    }
    __syncthreads();

    const int s0 = __ldg(img_start), sB = __ldg(img_start + B);
    const long long U = (long long)nchunks * (sB - s0);
    // Even split of the units only if a CTA's share covers the largest tile (then a tile is shared by
    // at most two CTAs and the merge is order-independent); otherwise whole tiles, round-robin.
    int maxroi = 0;
    for (int i = 0; i < B; ++i) maxroi = max(maxroi, __ldg(img_start + i + 1) - __ldg(img_start + i));
    const bool split = U / gridDim.x >= maxroi;
    long long u = split ? U * blockIdx.x / gridDim.x : 0;
    const long long u_end = split ? U * (blockIdx.x + 1) / gridDim.x : 0;
    int tile_id = blockIdx.x;                 // whole-tile mode: tiles blockIdx.x, + gridDim.x, ...
    float *tile_c = tile + lane * pitch;
    int gb = 0;                               // batches consumed so far (all segments): slot / phase bookkeeping
    int issued = 0;                           // batches issued so far (producer thread only)
    int b = 0;
    while (true) {
        int is, nroi, ch, r0, r1;
        if (split) {
            if (u >= u_end) break;
            while (b < B && (long long)nchunks * (__ldg(img_start + b + 1) - s0) <= u) ++b;
            is = __ldg(img_start + b);
            nroi = __ldg(img_start + b + 1) - is;
            const long long base = (long long)nchunks * (is - s0);
            ch = (int)((u - base) / nroi);
            r0 = (int)((u - base) - (long long)ch * nroi);
            r1 = (int)min((long long)nroi, r0 + (u_end - u));
            u += r1 - r0;
        } else {
            if (tile_id >= B * nchunks) break;
            b = tile_id / nchunks;
            ch = tile_id - b * nchunks;
            tile_id += gridDim.x;
            is = __ldg(img_start + b);
            nroi = __ldg(img_start + b + 1) - is;
            r0 = 0;
            r1 = nroi;
            if (nroi == 0) continue;          // its slab stays zero
        }
        const int c0 = ch * CH;
        const int first_roi = is + r0, n = r1 - r0;
        const int nbatch = (n + NBR - 1) / NBR, gb0 = gb;

        for (int e = tid; e < CH * pitch; e += NWB * 32) tile[e] = 0.f;
        __syncthreads();

        // producer duty (lane 0 of warp 0): batch j of this segment (NBR consecutive ROIs: gradients +
        // descriptors) goes to slot (gb0 + j) % NS once every warp has released the batch that used
        // the slot before.  `must` = the batch warp 0 itself is about to read: blocking wait.
        auto produce = [&](int want, int must) {
            while (issued < want && issued < gb0 + nbatch) {
                const int s = issued % NS;
                if (issued >= NS) {
                    const uint32_t par = ((issued / NS) - 1) & 1;
                    if (issued <= must) mbar_wait(&empty[s], par);
                    else if (!mbar_test(&empty[s], par)) break;
                }
                float *slot = ring + (size_t)s * SLOT_FLOATS;
                const int first = (issued - gb0) * NBR, cnt = min(NBR, n - first);
                mbar_expect_tx(&full[s], (uint32_t)cnt * (STAGE_FLOATS + DESC_WORDS) * 4);
                for (int j = 0; j < cnt; ++j)
                    bulk_g2s(slot + j * STAGE_FLOATS, grad_out + ((size_t)(first_roi + first + j) * C + c0) * NBIN,
                             STAGE_FLOATS * 4, &full[s]);
                bulk_g2s(slot + NBR * STAGE_FLOATS, descs + (size_t)(first_roi + first) * DESC_WORDS,
                         (uint32_t)cnt * DESC_WORDS * 4, &full[s]);
                ++issued;
            }
        };
        if (tid == 0) produce(gb0 + NS, -1);

        for (int bi = 0; bi < nbatch; ++bi, ++gb) {
            if (tid == 0) produce(gb + NS, gb);
            const int s = gb % NS;
            mbar_wait(&full[s], (gb / NS) & 1);
            const float *slot = ring + (size_t)s * SLOT_FLOATS;
            const int cnt = min(NBR, n - bi * NBR);
            for (int j = 0; j < cnt; ++j) {
                const int *d = reinterpret_cast<const int *>(slot + NBR * STAGE_FLOATS) + j * DESC_WORDS;
                if (((d[D_OWN] >> warp) & 1) == 0 || d[D_FLAGX] != 0) continue;   // not my rows / generic-path ROI
                const int2 yr = *reinterpret_cast<const int2 *>(d + D_Y0);
                const int y0 = yr.x, y1 = yr.y;
                const float *g = slot + j * STAGE_FLOATS + lane * NBIN;
                const int T = d[D_TX];
                if (d[D_XINC]) {
                    switch (T) {
                        case 2: bwd_rows<2, true>(tile_c, W, d, g, warp, y0, y1); break;
                        case 3: bwd_rows<3, true>(tile_c, W, d, g, warp, y0, y1); break;
                        case 4: bwd_rows<4, true>(tile_c, W, d, g, warp, y0, y1); break;
                        case 6: bwd_rows<6, true>(tile_c, W, d, g, warp, y0, y1); break;
                        default: bwd_rows<8, true>(tile_c, W, d, g, warp, y0, y1); break;
                    }
                } else {
                    switch (T) {
                        case 2: bwd_rows<2, false>(tile_c, W, d, g, warp, y0, y1); break;
                        case 3: bwd_rows<3, false>(tile_c, W, d, g, warp, y0, y1); break;
                        case 4: bwd_rows<4, false>(tile_c, W, d, g, warp, y0, y1); break;
                        case 6: bwd_rows<6, false>(tile_c, W, d, g, warp, y0, y1); break;
                        default: bwd_rows<8, false>(tile_c, W, d, g, warp, y0, y1); break;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&empty[s]);       // this warp is done reading slot s
        }
        __syncthreads();

        // tile -> grad_feat (+=): the slab is zero or holds the other CTA's partial sum
        float *dst = grad_feat + ((size_t)b * C + c0) * HW;
        if ((HW & 3) == 0) {
            for (int e = tid * 4; e < CH * HW; e += NWB * 32 * 4) {
                const int c = e / HW, p = e - c * HW;
                const float *t = tile + c * pitch + p;
                red_add4(dst + e, t[0], t[1], t[2], t[3]);
            }
        } else {
            for (int e = tid; e < CH * HW; e += NWB * 32) {
                const int c = e / HW, p = e - c * HW;
                atomicAdd(dst + e, tile[c * pitch + p]);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------- generic path
struct Geom {
    int b, gh, gw;
    float ys, xs, bh, bw, count;
};
__device__ __forceinline__ Geom roi_geom(const float *__restrict__ r, float scale, int oh, int ow,
                                         int sr, int aligned) {
    Geom g;
    const float off = aligned ? .5f : 0.f;
    g.b = (int)r[0];
    float x1 = __fsub_rn(__fmul_rn(r[1], scale), off), y1 = __fsub_rn(__fmul_rn(r[2], scale), off);
    float x2 = __fsub_rn(__fmul_rn(r[3], scale), off), y2 = __fsub_rn(__fmul_rn(r[4], scale), off);
    float rw = __fsub_rn(x2, x1), rh = __fsub_rn(y2, y1);
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    g.xs = x1; g.ys = y1;
    g.bh = __fdiv_rn(rh, (float)oh);
    g.bw = __fdiv_rn(rw, (float)ow);
    g.gh = sr > 0 ? sr : (int)ceilf(__fdiv_rn(rh, (float)oh));
    g.gw = sr > 0 ? sr : (int)ceilf(__fdiv_rn(rw, (float)ow));
    float cnt = fmaxf((float)(g.gh * g.gw), 1.f);
    g.count = cnt;
    return g;
}
__device__ __forceinline__ bool corners(int H, int W, float y, float x, int &y0, int &y1, int &x0,
                                        int &x1, float &w00, float &w01, float &w10, float &w11) {
    if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) return false;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    y0 = (int)y; x0 = (int)x;
    if (y0 >= H - 1) { y1 = y0 = H - 1; y = (float)y0; } else { y1 = y0 + 1; }
    if (x0 >= W - 1) { x1 = x0 = W - 1; x = (float)x0; } else { x1 = x0 + 1; }
    float ly = y - (float)y0, lx = x - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    w00 = hy * hx; w01 = hy * lx; w10 = ly * hx; w11 = ly * lx;
    return true;
}

// mode: 0 = every ROI; 1 = only ROIs the tile kernel skipped (flagged, or all if ungrouped)
template <bool BWD>
__global__ void roi_align_generic_kernel(const float *__restrict__ in, const float *__restrict__ rois,
                                         float *__restrict__ outp, const int *__restrict__ hdr,
                                         const int *__restrict__ descs, int mode, int B, int C, int H,
                                         int W, int K, int oh, int ow, float scale, int sr, int aligned) {
    const int k = blockIdx.x;
    if (mode == 1 && __ldg(hdr) == 0) {
        const int *d = descs + (size_t)k * DESC_WORDS;
        if ((__ldg(d + D_FLAGY) | __ldg(d + D_FLAGX)) == 0) return;
    }
    const Geom g = roi_geom(rois + 5 * (size_t)k, scale, oh, ow, sr, aligned);
    const int per_roi = C * oh * ow;
    const bool valid_b = g.b >= 0 && g.b < B;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < per_roi; e += gridDim.y * blockDim.x) {
        const int pw = e % ow, ph = (e / ow) % oh, c = e / (ow * oh);
        const size_t oidx = (size_t)k * per_roi + e;
        if (!valid_b) { if (!BWD) outp[oidx] = 0.f; continue; }
        const size_t plane = ((size_t)g.b * C + c) * H * W;
        float acc = 0.f;
        const float top = BWD ? in[oidx] : 0.f;
        for (int iy = 0; iy < g.gh; ++iy) {
            const float y = sample_coord(g.ys, g.bh, ph, iy, g.gh);
            for (int ix = 0; ix < g.gw; ++ix) {
                const float x = sample_coord(g.xs, g.bw, pw, ix, g.gw);
                int y0, y1, x0, x1; float a, b2, c2, d2;
                if (!corners(H, W, y, x, y0, y1, x0, x1, a, b2, c2, d2)) continue;
                if (BWD) {
                    float *p = outp + plane;
                    atomicAdd(p + y0 * W + x0, top * a / g.count);
                    atomicAdd(p + y0 * W + x1, top * b2 / g.count);
                    atomicAdd(p + y1 * W + x0, top * c2 / g.count);
                    atomicAdd(p + y1 * W + x1, top * d2 / g.count);
                } else {
                    const float *p = in + plane;
                    acc += a * __ldg(p + y0 * W + x0) + b2 * __ldg(p + y0 * W + x1) +
                           c2 * __ldg(p + y1 * W + x0) + d2 * __ldg(p + y1 * W + x1);
                }
            }
        }
        if (!BWD) outp[oidx] = acc / g.count;
    }
}

// ------------------------------------------------------------------------------------- host
struct Plan {
    bool tile;
    int pitch, fwd_warps;
    size_t smem_fwd, smem_bwd;
};
static Plan make_plan(int C, int H, int W, int oh, int ow) {
    Plan p{};
    const int HW = H * W;
    p.pitch = (HW & 1) ? HW : HW + 1;
    const size_t cap = (size_t)cim_max_smem_optin();
    // forward: as many warps (12 ... 4) as fit next to the tile: per warp one output stage and two
    // descriptor slots
    const size_t per_warp = (size_t)STAGE_FLOATS * 4 + 2 * DESC_WORDS * 4;
    p.fwd_warps = 12;
    while (p.fwd_warps > 4 && (size_t)CH * p.pitch * 4 + p.fwd_warps * per_warp > cap) p.fwd_warps -= 2;
    p.smem_fwd = (size_t)CH * p.pitch * 4 + p.fwd_warps * per_warp;
    p.smem_bwd = (size_t)CH * p.pitch * 4 + (size_t)NS * SLOT_FLOATS * 4 + 2 * NS * 8;
    p.tile = oh == PH && ow == PW && (C % CH) == 0 && W >= MAXT && p.smem_fwd <= cap &&
             p.smem_bwd <= cap && (((size_t)CH * p.pitch * 4) % 16 == 0);
    return p;
}

static int check_args(const void *a, const void *rois, const void *c, int B, int C, int H, int W, int K,
                      int oh, int ow, const void *ws, size_t ws_bytes) {
    if (!a || !c || (K > 0 && !rois)) return CIM_ERR_ARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K < 0 || oh <= 0 || ow <= 0) return CIM_ERR_ARG;
    if ((long long)H * W > (1 << 24) || (long long)C * oh * ow > (1LL << 30)) return CIM_ERR_SHAPE;
    if (!ws || ws_bytes < cim_roi_align_workspace_bytes(K)) return CIM_ERR_WORKSPACE;
    if (!cim_aligned(ws, 16) || !cim_aligned(a, 16) || !cim_aligned(c, 16)) return CIM_ERR_ALIGN;
    return CIM_OK;
}

static RoiWs carve(void *ws, int B) {
    RoiWs w;
    char *p = (char *)ws;
    w.hdr = (int *)p;
    w.img_start = (int *)(p + ws_img_off());
    w.desc = (int *)(p + ws_desc_off(B));
    return w;
}

static int run_prep(const float *rois, int B, int H, int W, int K, int oh, int ow, float scale, int sr,
                    int aligned, const RoiWs &w, cudaStream_t st) {
    cudaMemsetAsync(w.hdr, 0, WS_HDR_BYTES + sizeof(int) * (size_t)(B + 1), st);
    if (K > 0) {
        int threads = 128, blocks = (2 * K + threads - 1) / threads;
        roi_prep_kernel<<<blocks, threads, 0, st>>>(rois, K, B, H, W, oh, ow, scale, sr, aligned, w.hdr,
                                                    w.img_start, w.desc);
    }
    return cim_launch_status();
}

}  // namespace

CIM_API size_t cim_roi_align_workspace_bytes(int K) {
    // header + image ranges (up to 4096 images) + descriptors
    return WS_HDR_BYTES + 4097 * sizeof(int) + 64 + (size_t)(K > 0 ? K : 0) * DESC_WORDS * 4;
}

CIM_API int cim_roi_align_fwd(const float *feat, const float *rois, float *out, int B, int C, int H, int W,
                              int K, int oh, int ow, float scale, int sr, int aligned, void *ws,
                              size_t ws_bytes, cim_stream_t stream) {
    if (B > 4096) return CIM_ERR_SHAPE;
    if (K == 0 && feat && B > 0 && C > 0 && H > 0 && W > 0 && oh > 0 && ow > 0) return CIM_OK;   // empty output
    int rc = check_args(feat, rois, out, B, C, H, W, K, oh, ow, ws, ws_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const Plan p = make_plan(C, H, W, oh, ow);
    const RoiWs w = carve(ws, B);
    const int per_roi = C * oh * ow;
    dim3 ggrid((unsigned)K, (unsigned)min(64, (per_roi + 255) / 256));
    if (!p.tile) {
        roi_align_generic_kernel<false><<<ggrid, 256, 0, st>>>(feat, rois, out, nullptr, nullptr, 0, B, C, H, W,
                                                               K, oh, ow, scale, sr, aligned);
        return cim_launch_status();
    }
    if ((rc = run_prep(rois, B, H, W, K, oh, ow, scale, sr, aligned, w, st))) return rc;
    cudaFuncSetAttribute(roi_align_fwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_fwd);
    const long long units = (long long)(C / CH) * K;
    const int grid = (int)min((long long)cim_num_sms(), units);
    roi_align_fwd_tile_kernel<<<grid, p.fwd_warps * 32, p.smem_fwd, st>>>(feat, w.hdr, w.img_start, w.desc, out, B,
                                                                           C, H, W, p.pitch);
    if ((rc = cim_launch_status())) return rc;
    // leftover pass: one CTA per ROI, which exits at once unless the tile kernel skipped that ROI
    roi_align_generic_kernel<false><<<dim3((unsigned)K, 1), 256, 0, st>>>(feat, rois, out, w.hdr, w.desc, 1, B, C, H,
                                                                          W, K, oh, ow, scale, sr, aligned);
    return cim_launch_status();
}

CIM_API int cim_roi_align_bwd(const float *grad_out, const float *rois, float *grad_feat, int B, int C, int H,
                              int W, int K, int oh, int ow, float scale, int sr, int aligned, void *ws,
                              size_t ws_bytes, cim_stream_t stream) {
    if (B > 4096) return CIM_ERR_SHAPE;
    if (K == 0 && grad_feat && B > 0 && C > 0 && H > 0 && W > 0) {                // no ROI: zero gradient
        cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W, (cudaStream_t)stream);
        return cim_launch_status();
    }
    int rc = check_args(grad_out, rois, grad_feat, B, C, H, W, K, oh, ow, ws, ws_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const Plan p = make_plan(C, H, W, oh, ow);
    const RoiWs w = carve(ws, B);
    const int per_roi = C * oh * ow;
    dim3 ggrid((unsigned)max(K, 1), (unsigned)min(64, (per_roi + 255) / 256));
    if (!p.tile || K == 0) {
        cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W, st);
        if (K > 0)
            roi_align_generic_kernel<true><<<ggrid, 256, 0, st>>>(grad_out, rois, grad_feat, nullptr, nullptr, 0,
                                                                  B, C, H, W, K, oh, ow, scale, sr, aligned);
        return cim_launch_status();
    }
    if ((rc = run_prep(rois, B, H, W, K, oh, ow, scale, sr, aligned, w, st))) return rc;
    cudaFuncSetAttribute(roi_align_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bwd);
    cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)B * C * H * W, st);
    const long long units = (long long)(C / CH) * K;
    const int grid = (int)min((long long)cim_num_sms(), units);
    roi_align_bwd_tile_kernel<<<grid, NWB * 32, p.smem_bwd, st>>>(grad_out, w.hdr, w.img_start, w.desc, grad_feat, B,
                                                                   C, H, W, p.pitch);
    if ((rc = cim_launch_status())) return rc;
    roi_align_generic_kernel<true><<<dim3((unsigned)K, 1), 256, 0, st>>>(grad_out, rois, grad_feat, w.hdr, w.desc, 1,
                                                                         B, C, H, W, K, oh, ow, scale, sr, aligned);
    return cim_launch_status();
}
