// le_lsd.cu -- LSD line segment detector (a7), refine 0/1, for sm_100a.  Compiled with -fmad=false:
// cv2's scalar code is unfused, and region growing feeds rounding differences back into decisions.
//
// Pipeline per frame (all intermediates stay in HBM scratch owned by the context):
//   1  Q8 Gaussian (taps from the host, REFLECT_101) + exact-bilinear down-scale  -> scaled u8
//   2  2x2 gradient: level-line angle (degrees, f32, cv2's fastAtan2 polynomial) packed in a u32
//      with bit 31 = USED and 0xFFFFFFFF = NOTDEF; frame max of gx^2+gy^2
//   3  stable counting sort of the DEFINED pixels by descending gradient bin (row-major in a bin):
//      per-chunk u16 histograms -> column scan -> warp-stable scatter (match_any ranks)
//   4  greedy region growing in that order, one WARP per frame: lanes fetch the 8 neighbours of a
//      region point in one memory round, the accept chain (running mean angle) is warp-uniform;
//      then the inertia rectangle (f64, sequential sums in region order like cv2).
// Region growing is sequential inside a frame by definition (each accepted pixel moves the
// region angle used by the next test); the data parallelism is across frames.
#include "le_common.cuh"
#include <vector>
#include <stdlib.h>

#define LSD_NOTDEF 0xFFFFFFFFu
#define LSD_USED 0x80000000u
#define D2R 0.017453292519943295  // CV_PI / 180

// ------------------------------------------------------------------ 1a: generic Q8 separable blur
#define SB_TW 64
#define SB_TH 16
#define SB_RMAX 8
__global__ void __launch_bounds__(256) k_sepblur_q8(const u8 *__restrict__ src, u8 *__restrict__ dst, int h, int w,
                                                    const int *__restrict__ taps, int r)
{
    __shared__ u8 g[SB_TH + 2 * SB_RMAX][SB_TW + 2 * SB_RMAX];
    __shared__ u16 hb[SB_TH + 2 * SB_RMAX][SB_TW];
    __shared__ int tp[2 * SB_RMAX + 1];
    const u8 *s = src + (size_t)blockIdx.z * h * w;
    u8 *d = dst + (size_t)blockIdx.z * h * w;
    const int x0 = blockIdx.x * SB_TW, y0 = blockIdx.y * SB_TH;
    const int gw = SB_TW + 2 * r, gh = SB_TH + 2 * r;
    if (threadIdx.x < 2 * r + 1) tp[threadIdx.x] = taps[threadIdx.x];
    for (int i = threadIdx.x; i < gh * gw; i += 256) {
        int ty = i / gw, tx = i % gw;
        g[ty][tx] = s[(size_t)d_reflect101(y0 + ty - r, h) * w + d_reflect101(x0 + tx - r, w)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < gh * SB_TW; i += 256) {
        int ty = i / SB_TW, tx = i % SB_TW;
        u32 acc = 0;
        for (int t = 0; t <= 2 * r; t++) acc += (u32)tp[t] * g[ty][tx + t];
        hb[ty][tx] = (u16)acc;  // <= 255*256
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SB_TH * SB_TW; i += 256) {
        int ty = i / SB_TW, tx = i % SB_TW;
        int x = x0 + tx, y = y0 + ty;
        if (x < w && y < h) {
            u32 acc = 0;
            for (int t = 0; t <= 2 * r; t++) acc += (u32)tp[t] * hb[ty + t][tx];
            d[(size_t)y * w + x] = (u8)((acc + 32768u) >> 16);
        }
    }
}

// ------------------------------------------------------------------ 1a': the same blur, word-wide
// Rows that are multiples of 4 bytes (1080p, 4K, 600x400 ...): the tile is loaded as aligned 32-bit words; the
// VERTICAL pass runs first on packed pixel pairs (a tap times a byte stays below 2^16 and so does the whole
// column sum, so one IMAD on (p0 | p1 << 16) serves two pixels), the horizontal pass then reads the u16 column
// sums back.  Both passes are exact integer sums, so their order does not change the result
// ((sum_ij t_i t_j p + 32768) >> 16 either way).
#define SV_TW 128                        // output pixels per tile row (32 words)
#define SV_TH 32
#define SV_CW (SV_TW / 4 + 4)            // words per staged row: 8 halo pixels on each side (R <= 8)
template <int R>
__global__ void __launch_bounds__(256) k_sepblur_q8_w(const u8 *__restrict__ src, u8 *__restrict__ dst, int h, int w,
                                                      const int *__restrict__ taps)
{
    __shared__ u32 raw[SV_TH + 2 * R][SV_CW];
    __shared__ __align__(8) u32 vb[SV_TH][SV_CW][2];  // column sums, u16 x 4 per staged word
    __shared__ int tp[2 * R + 1];
    const u8 *s = src + (size_t)blockIdx.z * h * w;
    u8 *d = dst + (size_t)blockIdx.z * h * w;
    const int x0 = blockIdx.x * SV_TW, y0 = blockIdx.y * SV_TH;
    if (threadIdx.x < 2 * R + 1) tp[threadIdx.x] = taps[threadIdx.x];
    const bool interior = x0 >= 8 && x0 + SV_TW + 8 <= w;
    for (int i = threadIdx.x; i < (SV_TH + 2 * R) * SV_CW; i += 256) {
        const int ty = i / SV_CW, c = i - ty * SV_CW;
        const u8 *row = s + (size_t)d_reflect101(y0 + ty - R, h) * w;
        const int xb = x0 - 8 + 4 * c;
        u32 v;
        if (interior) v = *(const u32 *)(row + xb);
        else {
            v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) v |= (u32)row[d_reflect101(min(xb + k, 2 * w - 2), w)] << (8 * k);
        }
        raw[ty][c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SV_TH * SV_CW; i += 256) {
        const int ty = i / SV_CW, c = i - ty * SV_CW;
        u32 a01 = 0, a23 = 0;
#pragma unroll
        for (int t = 0; t <= 2 * R; t++) {
            const u32 wd = raw[ty + t][c];
            a01 += (u32)tp[t] * __byte_perm(wd, 0, 0x4140);
            a23 += (u32)tp[t] * __byte_perm(wd, 0, 0x4342);
        }
        *(uint2 *)vb[ty][c] = make_uint2(a01, a23);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SV_TH * (SV_TW / 4); i += 256) {
        const int ty = i / (SV_TW / 4), oc = i - ty * (SV_TW / 4);
        const int x = x0 + 4 * oc, y = y0 + ty;
        if (x >= w || y >= h) continue;
        const u16 *v16 = (const u16 *)vb[ty] + 8 + 4 * oc - R;  // column sum of pixel x - R
        u32 vv[4 + 2 * R];
#pragma unroll
        for (int j = 0; j < 4 + 2 * R; j++) vv[j] = v16[j];
        u32 out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u32 acc = 0;
#pragma unroll
            for (int t = 0; t <= 2 * R; t++) acc += (u32)tp[t] * vv[k + t];
            out |= ((acc + 32768u) >> 16) << (8 * k);
        }
        *(u32 *)(d + (size_t)y * w + x) = out;
    }
}

// ------------------------------------------------------------------ 1b: INTER_LINEAR_EXACT resize
// xtab/ytab: (index, q8 weight of the next sample) per destination column / row, from the host
__global__ void __launch_bounds__(256) k_resize_exact(const u8 *__restrict__ src, int h, int w, u8 *__restrict__ dst,
                                                      int dh, int dw, const int2 *__restrict__ xtab,
                                                      const int2 *__restrict__ ytab)
{
    const int f = blockIdx.z;
    const u8 *s = src + (size_t)f * h * w;
    u8 *d = dst + (size_t)f * dh * dw;
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    int2 xt = xtab[x], yt = ytab[y];
    int x0 = xt.x, x1 = min(x0 + 1, w - 1), y0 = yt.x, y1 = min(y0 + 1, h - 1);
    const u8 *r0 = s + (size_t)y0 * w, *r1 = s + (size_t)y1 * w;
    u32 a = (u32)(256 - xt.y) * r0[x0] + (u32)xt.y * r0[x1];
    u32 b = (u32)(256 - xt.y) * r1[x0] + (u32)xt.y * r1[x1];
    u32 v = (u32)(256 - yt.y) * a + (u32)yt.y * b;
    d[(size_t)y * dw + x] = (u8)((v + 32768u) >> 16);
}

// ------------------------------------------------------------------ 1a' + 1b fused: blur tile -> resize
// One CTA produces a dtw x dth tile of the down-scaled image: it stages the source words its samples need
// (+ the blur halo), runs the two blur passes of k_sepblur_q8_w into a shared 128 x 32 tile of blurred bytes
// and samples that tile with the exact bilinear weights -- the full-resolution blurred image (one write and one
// read of W*H bytes per frame) never exists.  The host picks dtw / dth so that every tile's samples fit the
// shared tile (checked against the tables).  Same integer arithmetic as the two kernels: same bytes.
template <int R>
__global__ void __launch_bounds__(256) k_blur_resize_w(const u8 *__restrict__ src, int h, int w, u8 *__restrict__ dst,
                                                       int dh, int dw, int dth, int dtw, const int *__restrict__ taps,
                                                       const int2 *__restrict__ xtab, const int2 *__restrict__ ytab)
{
    __shared__ u32 raw[SV_TH + 2 * R][SV_CW];
    __shared__ __align__(8) u32 vb[SV_TH][SV_CW][2];  // column sums, u16 x 4 per staged word
    __shared__ u32 bl[SV_TH][SV_TW / 4];              // blurred bytes
    __shared__ int tp[2 * R + 1];
    const u8 *s = src + (size_t)blockIdx.z * h * w;
    u8 *d = dst + (size_t)blockIdx.z * dh * dw;
    const int dx0 = blockIdx.x * dtw, dy0 = blockIdx.y * dth;
    const int tw = min(dtw, dw - dx0), th = min(dth, dh - dy0);
    const int x0 = xtab[dx0].x & ~3, y0 = ytab[dy0].x;  // source origin of the shared tile
    if (threadIdx.x < 2 * R + 1) tp[threadIdx.x] = taps[threadIdx.x];
    const bool interior = x0 >= 8 && x0 + SV_TW + 8 <= w;
    for (int i = threadIdx.x; i < (SV_TH + 2 * R) * SV_CW; i += 256) {
        const int ty = i / SV_CW, c = i - ty * SV_CW;
        const u8 *row = s + (size_t)d_reflect101(y0 + ty - R, h) * w;
        const int xb = x0 - 8 + 4 * c;
        u32 v;
        if (interior) v = *(const u32 *)(row + xb);
        else {
            v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) v |= (u32)row[d_reflect101(min(xb + k, 2 * w - 2), w)] << (8 * k);
        }
        raw[ty][c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SV_TH * SV_CW; i += 256) {
        const int ty = i / SV_CW, c = i - ty * SV_CW;
        u32 a01 = 0, a23 = 0;
#pragma unroll
        for (int t = 0; t <= 2 * R; t++) {
            const u32 wd = raw[ty + t][c];
            a01 += (u32)tp[t] * __byte_perm(wd, 0, 0x4140);
            a23 += (u32)tp[t] * __byte_perm(wd, 0, 0x4342);
        }
        *(uint2 *)vb[ty][c] = make_uint2(a01, a23);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SV_TH * (SV_TW / 4); i += 256) {
        const int ty = i / (SV_TW / 4), oc = i - ty * (SV_TW / 4);
        const u16 *v16 = (const u16 *)vb[ty] + 8 + 4 * oc - R;  // column sum of pixel x0 + 4 oc - R
        u32 vv[4 + 2 * R];
#pragma unroll
        for (int j = 0; j < 4 + 2 * R; j++) vv[j] = v16[j];
        u32 out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u32 acc = 0;
#pragma unroll
            for (int t = 0; t <= 2 * R; t++) acc += (u32)tp[t] * vv[k + t];
            out |= ((acc + 32768u) >> 16) << (8 * k);
        }
        bl[ty][oc] = out;
    }
    __syncthreads();
    // resize: the tile's column table (source column relative to the tile, its right neighbour, Q8 weight) is staged
    // once; a thread produces FOUR neighbouring destination pixels and stores one word (dtw and dx0 are multiples
    // of four: the host picks dtw that way) -- the per-pixel index arithmetic, table loads and byte stores of the
    // one-pixel-per-thread loop were 47 % of the kernel's instructions (profiles/r2_ncu_digests.txt)
    const u8 *b8 = (const u8 *)bl;
    __shared__ u32 s_xt[SV_TW];
    for (int xx = threadIdx.x; xx < tw; xx += 256) {
        const int2 xt = xtab[dx0 + xx];
        s_xt[xx] = (u32)(xt.x - x0) | ((u32)(min(xt.x + 1, w - 1) - x0) << 8) | ((u32)xt.y << 16);
    }
    __syncthreads();
    const int ngrp = (tw + 3) >> 2;
    for (int i = threadIdx.x; i < th * ngrp; i += 256) {
        const int yy = i / ngrp, xg = (i - yy * ngrp) * 4;
        const int2 yt = ytab[dy0 + yy];
        const int r0 = (yt.x - y0) * SV_TW, r1 = (min(yt.x + 1, h - 1) - y0) * SV_TW;
        const u32 wy1 = (u32)yt.y, wy0 = 256u - wy1;
        u32 out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 e = s_xt[min(xg + k, tw - 1)];
            const int c0 = (int)(e & 0xff), c1 = (int)((e >> 8) & 0xff);
            const u32 wx1 = e >> 16, wx0 = 256u - wx1;
            const u32 a = wx0 * b8[r0 + c0] + wx1 * b8[r0 + c1];
            const u32 b = wx0 * b8[r1 + c0] + wx1 * b8[r1 + c1];
            out |= ((wy0 * a + wy1 * b + 32768u) >> 16) << (8 * k);
        }
        u8 *o = d + (size_t)(dy0 + yy) * dw + dx0 + xg;
        if (xg + 4 <= tw && ((dw | dx0) & 3) == 0) *(u32 *)o = out;
        else
            for (int k = 0; k < 4 && xg + k < tw; k++) o[k] = (u8)(out >> (8 * k));
    }
}

// ------------------------------------------------------------------ 2: gradient / level-line angle
static __device__ __forceinline__ float d_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = 2.220446049250313e-16f;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

static __device__ __forceinline__ int grad_s(const u8 *__restrict__ img, int W, int x, int y, int *gx, int *gy)
{
    const u8 *r0 = img + (size_t)y * W + x, *r1 = r0 + W;
    int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
    *gx = DA + BC;
    *gy = DA - BC;
    return *gx * *gx + *gy * *gy;
}

__global__ void __launch_bounds__(256) k_lsd_grad(const u8 *__restrict__ img, u32 *__restrict__ ang,
                                                  int *__restrict__ counters, int H, int W, int s_thr)
{
    const int f = blockIdx.z;
    const u8 *im = img + (size_t)f * H * W;
    u32 *an = ang + (size_t)f * H * W;
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    int smax = 0;
    if (x < W && y < H) {
        u32 out = LSD_NOTDEF;
        if (x < W - 1 && y < H - 1) {
            int gx, gy;
            int s = grad_s(im, W, x, y, &gx, &gy);
            if (s > s_thr) {  // sqrt(s/4.0) > rho, threshold resolved on the host in f64
                float deg = d_fast_atan2((float)gx, (float)-gy);
                out = __float_as_uint(deg);
                smax = s;
            }
        }
        an[(size_t)y * W + x] = out;
    }
    for (int o = 16; o; o >>= 1) smax = max(smax, __shfl_xor_sync(~0u, smax, o));
    if ((threadIdx.x & 31) == 0 && smax > 0) atomicMax(&counters[f * LE_C_STRIDE + LE_C_MAX], smax);
}

// four pixels per thread for rows that are multiples of 4 bytes: two aligned words + one byte per row in,
// one 16-byte store out
__global__ void __launch_bounds__(256) k_lsd_grad4(const u8 *__restrict__ img, u32 *__restrict__ ang,
                                                   int *__restrict__ counters, int H, int W, int s_thr)
{
    const int f = blockIdx.z;
    const u8 *im = img + (size_t)f * H * W;
    u32 *an = ang + (size_t)f * H * W;
    const int x = 4 * (blockIdx.x * 32 + (threadIdx.x & 31)), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    int smax = 0;
    if (x < W && y < H) {
        u32 out[4] = {LSD_NOTDEF, LSD_NOTDEF, LSD_NOTDEF, LSD_NOTDEF};
        if (y < H - 1) {
            const u8 *r0 = im + (size_t)y * W + x, *r1 = r0 + W;
            const u32 w0 = *(const u32 *)r0, w1 = *(const u32 *)r1;
            const bool last = x + 4 >= W;  // the row's last word: its fourth pixel has no right neighbour
            int a[5], b[5];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                a[k] = (int)((w0 >> (8 * k)) & 0xff);
                b[k] = (int)((w1 >> (8 * k)) & 0xff);
            }
            a[4] = last ? 0 : (int)r0[4];
            b[4] = last ? 0 : (int)r1[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k == 3 && last) break;
                const int DA = b[k + 1] - a[k], BC = a[k + 1] - b[k];
                const int gx = DA + BC, gy = DA - BC;
                const int s = gx * gx + gy * gy;
                if (s > s_thr) {
                    out[k] = __float_as_uint(d_fast_atan2((float)gx, (float)-gy));
                    smax = max(smax, s);
                }
            }
        }
        *(uint4 *)(an + (size_t)y * W + x) = make_uint4(out[0], out[1], out[2], out[3]);
    }
    for (int o = 16; o; o >>= 1) smax = max(smax, __shfl_xor_sync(~0u, smax, o));
    if ((threadIdx.x & 31) == 0 && smax > 0) atomicMax(&counters[f * LE_C_STRIDE + LE_C_MAX], smax);
}

// ------------------------------------------------------------------ 3: stable counting sort
#define LS_CHUNK 8192  // pixels per sort chunk (multiple of 128; a u16 chunk histogram holds up to 65535)
static __device__ __forceinline__ int lsd_bin(int s, double bin_coef, int n_bins)
{
    int b = (int)(sqrt((double)s / 4.0) * bin_coef);
    return min(b, n_bins - 1);
}
static __device__ __forceinline__ double lsd_bin_coef(const int *counters, int f, int n_bins)
{
    int smax = counters[f * LE_C_STRIDE + LE_C_MAX];
    return smax > 0 ? (double)(n_bins - 1) / sqrt((double)smax / 4.0) : 0.0;
}

// per-chunk histogram (u16) of defined pixels; chunk = LS_CHUNK consecutive pixels (row-major)
__global__ void __launch_bounds__(256) k_lsd_hist(const u8 *__restrict__ img, const u32 *__restrict__ ang,
                                                  const int *__restrict__ counters, u16 *__restrict__ chist, int H,
                                                  int W, int nchunk, int n_bins)
{
    extern __shared__ u32 sh[];  // n_bins counters
    const int f = blockIdx.y, c = blockIdx.x;
    const size_t npx = (size_t)H * W;
    const u8 *im = img + (size_t)f * npx;
    const u32 *an = ang + (size_t)f * npx;
    for (int i = threadIdx.x; i < n_bins; i += 256) sh[i] = 0;
    __syncthreads();
    const double coef = lsd_bin_coef(counters, f, n_bins);
    size_t b0 = (size_t)c * LS_CHUNK, b1 = min(npx, b0 + LS_CHUNK);
    auto vote = [&](size_t i) {
        int y = (int)(i / W), x = (int)(i - (size_t)y * W), gx, gy;
        atomicAdd(&sh[lsd_bin(grad_s(im, W, x, y, &gx, &gy), coef, n_bins)], 1u);
    };
    if ((npx & 3) == 0 && (((size_t)an) & 15) == 0) {
        // 16 bytes of angle words per thread and load, two loads in flight: the kernel was bound by the latency of
        // one 4-byte load per thread and trip (39 long-scoreboard stall cycles per issue, 25 % of the DRAM rate)
        const uint4 *a4 = (const uint4 *)an;
        for (size_t v = b0 / 4 + threadIdx.x; v < b1 / 4; v += 512) {
            const uint4 q0 = __ldg(a4 + v);
            const bool two = v + 256 < b1 / 4;
            const uint4 q1 = two ? __ldg(a4 + v + 256) : make_uint4(LSD_NOTDEF, LSD_NOTDEF, LSD_NOTDEF, LSD_NOTDEF);
            if (q0.x != LSD_NOTDEF) vote(4 * v);
            if (q0.y != LSD_NOTDEF) vote(4 * v + 1);
            if (q0.z != LSD_NOTDEF) vote(4 * v + 2);
            if (q0.w != LSD_NOTDEF) vote(4 * v + 3);
            if (q1.x != LSD_NOTDEF) vote(4 * (v + 256));
            if (q1.y != LSD_NOTDEF) vote(4 * (v + 256) + 1);
            if (q1.z != LSD_NOTDEF) vote(4 * (v + 256) + 2);
            if (q1.w != LSD_NOTDEF) vote(4 * (v + 256) + 3);
        }
    } else {
        for (size_t i = b0 + threadIdx.x; i < b1; i += 256)
            if (an[i] != LSD_NOTDEF) vote(i);
    }
    __syncthreads();
    u16 *o = chist + ((size_t)f * nchunk + c) * n_bins;
    for (int i = threadIdx.x; i < n_bins; i += 256) o[i] = (u16)sh[i];
}

// offsets: for bin b (descending order) and chunk c (ascending): start position in the ordered list.
// One block per frame, thread per bin for the chunk-direction scan, then a block scan over bins.
__global__ void __launch_bounds__(1024) k_lsd_scan(u16 *__restrict__ chist, u32 *__restrict__ coff,
                                                   int *__restrict__ counters, int nchunk, int n_bins)
{
    extern __shared__ u32 tot[];  // n_bins totals, then bin starts
    const int f = blockIdx.x;
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
        u32 acc = 0;
        for (int c = 0; c < nchunk; c++) {
            size_t k = ((size_t)f * nchunk + c) * n_bins + b;
            u32 v = chist[k];
            coff[k] = acc;  // offset of this chunk inside the bin
            acc += v;
        }
        tot[b] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 acc = 0;
        for (int b = n_bins - 1; b >= 0; b--) {
            u32 v = tot[b];
            tot[b] = acc;
            acc += v;
        }
        counters[f * LE_C_STRIDE + LE_C_AUX0] = (int)acc;  // number of defined pixels
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) {
        u32 st = tot[b];
        for (int c = 0; c < nchunk; c++) coff[((size_t)f * nchunk + c) * n_bins + b] += st;
    }
}

// stable scatter: one warp per chunk walks its pixels in row-major order, 32 at a time
__global__ void __launch_bounds__(32) k_lsd_scatter(const u8 *__restrict__ img, const u32 *__restrict__ ang,
                                                    const int *__restrict__ counters, const u32 *__restrict__ coff,
                                                    u32 *__restrict__ order, int H, int W, int nchunk, int n_bins)
{
    extern __shared__ u32 off[];  // n_bins running offsets of this chunk
    const int f = blockIdx.y, c = blockIdx.x, lane = threadIdx.x;
    const size_t npx = (size_t)H * W;
    const u8 *im = img + (size_t)f * npx;
    const u32 *an = ang + (size_t)f * npx;
    u32 *ord = order + (size_t)f * npx;
    const u32 *co = coff + ((size_t)f * nchunk + c) * n_bins;
    for (int i = lane; i < n_bins; i += 32) off[i] = co[i];
    __syncwarp();
    const double coef = lsd_bin_coef(counters, f, n_bins);
    size_t b0 = (size_t)c * LS_CHUNK, b1 = min(npx, b0 + LS_CHUNK);
    // 128 pixels per trip: the four angle words of a lane are fetched together (one memory round per trip
    // instead of four), then ranked 32 at a time in row-major order
    u32 nx[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const size_t i = b0 + 32 * q + lane;
        nx[q] = i < b1 ? __ldg(an + i) : LSD_NOTDEF;
    }
    for (size_t j0 = b0; j0 < b1; j0 += 128) {
        u32 av[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            av[q] = nx[q];
            const size_t i = j0 + 128 + 32 * q + lane;  // the next trip's words are in flight while this one is ranked
            nx[q] = i < b1 ? __ldg(an + i) : LSD_NOTDEF;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const size_t i = j0 + 32 * q + lane;
            const bool def = av[q] != LSD_NOTDEF;
            if (!__any_sync(~0u, def)) continue;
            int bin = -1 - lane;  // unique negative keys for undefined lanes
            if (def) {
                int y = (int)(i / W), x = (int)(i - (size_t)y * W), gx, gy;
                bin = lsd_bin(grad_s(im, W, x, y, &gx, &gy), coef, n_bins);
            }
            u32 peers = __match_any_sync(~0u, bin);
            if (def) {
                int rank = __popc(peers & ((1u << lane) - 1));
                u32 base = off[bin];
                ord[base + rank] = (u32)i;
                __syncwarp(peers);
                if (rank == __popc(peers) - 1) off[bin] = base + rank + 1;
            }
            __syncwarp();
        }
    }
}

// cos/sin of the level-line angle exactly as region growing uses them (cv2 adds cos(float(angle)),
// sin(float(angle)) to its f32 running sums), for the DEFINED pixels only: dense over the ordered list, so the
// double-precision cos/sin run on full warps instead of on the few defined lanes of an image-order warp
__global__ void __launch_bounds__(256) k_lsd_cs(const u32 *__restrict__ ang, const u32 *__restrict__ order,
                                                const int *__restrict__ counters, float2 *__restrict__ cs, size_t npx)
{
    const int f = blockIdx.y;
    const int ndef = counters[f * LE_C_STRIDE + LE_C_AUX0];
    const u32 *an = ang + (size_t)f * npx, *ord = order + (size_t)f * npx;
    float2 *c = cs + (size_t)f * npx;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < ndef; i += gridDim.x * 256) {
        const u32 p = ord[i];
        const float af = (float)((double)__uint_as_float(an[p]) * D2R);
        c[p] = make_float2((float)cos((double)af), (float)sin((double)af));
    }
}

// ------------------------------------------------------------------ 4: region growing + rectangle
struct LsdParams {
    int H, W, n_bins, refine, max_segs;
    unsigned min_reg_size;
    double prec, p, scale, density_th, log_eps, log_nt;
};

static __device__ __forceinline__ bool aligned_to(double theta, double a, double prec)
{
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > 4.71238898038468985769) {  // 3*pi/2
        n_theta -= 6.28318530717958647692;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}
static __device__ __forceinline__ double angle_diff_signed(double a, double b)
{
    double diff = a - b;
    while (diff <= -3.14159265358979323846) diff += 6.28318530717958647692;
    while (diff > 3.14159265358979323846) diff -= 6.28318530717958647692;
    return diff;
}

struct LsdRect {
    double x1, y1, x2, y2, width, x, y, theta, dx, dy;
    double prec, p;  // angle tolerance and its probability (LSD_REFINE_ADV may halve them)
};

// grow from seed (sx, sy); returns region size.  Warp-uniform; lane k<9 prefetches neighbour k.
// One frame is owned by exactly one warp, so its angle plane may live in L1: these loads only have to
// defeat register caching by the compiler (asm volatile), not the cache (the SM's own stores keep L1 coherent).
static __device__ __forceinline__ u32 ld_ca(const volatile u32 *p)
{
    u32 v;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#define LSD_RING 256  // most recent region points, kept in shared memory for the BFS front

// BFS trips take up to LSD_FRONT queued pixels at once: lanes 0-8 hold the neighbours of queue entry i, lanes 9-17
// those of entry i + 1, lanes 18-26 those of entry i + 2, all fetched in ONE memory round trip (the growth is a
// chain of dependent round trips, one per queue entry before).  The entries are then processed one after the
// other exactly as the sequential scan would: a neighbour accepted while an earlier entry of the trip was
// processed is struck from the later entries' candidates by comparing coordinates (its USED flag was loaded
// before it was set), so the order of acceptance and the running region angle are unchanged.
#define LSD_FRONT 3
static __device__ int lsd_grow(volatile u32 *an, const float2 *__restrict__ cs, u32 *reg, u32 *ring, int W, int H, int sx, int sy, double tol,
                               double *reg_angle_out, int lane)
{
    int n = 0;
    u32 sv = ld_ca(an + (size_t)sy * W + sx) & ~LSD_USED;
    double reg_angle = (double)__uint_as_float(sv) * D2R;
    if (lane == 0) {
        u32 e = ((u32)sy << 16) | (u32)sx;
        reg[0] = e;
        ring[0] = e;
        an[(size_t)sy * W + sx] = sv | LSD_USED;
    }
    n = 1;
    float sumdx = (float)cos(reg_angle), sumdy = (float)sin(reg_angle);
    __syncwarp();
    const int g = lane / 9, kk = lane - 9 * g;  // queue entry of the trip, neighbour of that entry
    const int ky = kk / 3 - 1, kx = kk % 3 - 1;
    for (int i = 0; i < n;) {
        const int m = min(LSD_FRONT, n - i);
        // BFS front: the queue tail is almost always within the shared ring
        int px = 0, py = 0;
        u32 v = LSD_NOTDEF;
        float2 csn = make_float2(0.f, 0.f);
        if (g < m) {
            const int qi = i + g;
            const u32 pt = (n - qi <= LSD_RING) ? ((volatile u32 *)ring)[qi & (LSD_RING - 1)] : ((volatile u32 *)reg)[qi];
            px = (int)(pt & 0xffff);
            py = (int)(pt >> 16);
        }
        const int xx = px + kx, yy = py + ky;
        if (g < m && xx >= 0 && xx < W && yy >= 0 && yy < H) {
            v = ld_ca(an + (size_t)yy * W + xx);
            csn = __ldg(cs + (size_t)yy * W + xx);  // read-only plane; garbage where the angle is undefined (never used)
        }
        bool avail = v != LSD_NOTDEF && !(v & LSD_USED);
        const double a_l = (double)__uint_as_float(v) * D2R;  // this lane's neighbour's level-line angle
        for (int gg = 0; gg < m; gg++) {
            // The neighbours that are defined and free are visited in (yy, xx) scan order.  Every lane tests ITS
            // neighbour against the current region angle at once; the first aligned one in scan order is accepted
            // (which moves the angle), the ones before it were rejected at that angle and stay rejected for this
            // entry, the ones behind it are tested again at the new angle -- the sequential scan, with rejected
            // candidates costing nothing.
            u32 remaining = (__ballot_sync(~0u, avail) >> (9 * gg)) & 0x1ffu;
            const int cpx = __shfl_sync(~0u, px, 9 * gg), cpy = __shfl_sync(~0u, py, 9 * gg);
            while (remaining) {
                const bool al = aligned_to(reg_angle, a_l, tol) && avail;
                const u32 mm = (__ballot_sync(~0u, al) >> (9 * gg)) & remaining;
                if (!mm) break;
                const int k = __ffs(mm) - 1;
                remaining &= ~((2u << k) - 1u);
                const int src = 9 * gg + k;
                u32 vk = __shfl_sync(~0u, v, src);
                int ax = cpx + (k % 3) - 1, ay = cpy + (k / 3) - 1;
                if (lane == 0) {
                    u32 e = ((u32)ay << 16) | (u32)ax;
                    an[(size_t)ay * W + ax] = vk | LSD_USED;
                    reg[n] = e;
                    ring[n & (LSD_RING - 1)] = e;
                }
                n++;
                sumdx += __shfl_sync(~0u, csn.x, src);
                sumdy += __shfl_sync(~0u, csn.y, src);
                reg_angle = (double)d_fast_atan2(sumdy, sumdx) * D2R;
                if (xx == ax && yy == ay) avail = false;  // also a neighbour of a later entry of this trip
            }
        }
        i += m;
        __syncwarp();
    }
    *reg_angle_out = reg_angle;
    return n;
}

// cos / sin of the rectangle angle from a double-double evaluation (~100 bits), rounded once.  cv2 takes them from
// glibc, whose results are correctly rounded for 99.7 percent of the angles LSD produces (4 M sampled); CUDA's
// cos / sin are 1-2 ulp functions and differ far more often, and the last bit of dx / dy decides whether a pixel
// that lies exactly on a rectangle edge (the extreme pixels of the region always do) is inside.
struct dd_t { double h, l; };
static __device__ __forceinline__ dd_t dd_two_sum(double a, double b) { double s = a + b, bb = s - a; return dd_t{s, (a - (s - bb)) + (b - bb)}; }
static __device__ __forceinline__ dd_t dd_fast_two_sum(double a, double b) { double s = a + b; return dd_t{s, b - (s - a)}; }
static __device__ __forceinline__ dd_t dd_two_prod(double a, double b) { double p = a * b; return dd_t{p, __fma_rn(a, b, -p)}; }
static __device__ __forceinline__ dd_t dd_add(dd_t a, dd_t b) { dd_t s = dd_two_sum(a.h, b.h); return dd_fast_two_sum(s.h, s.l + (a.l + b.l)); }
static __device__ __forceinline__ dd_t dd_mul(dd_t a, dd_t b) { dd_t p = dd_two_prod(a.h, b.h); return dd_fast_two_sum(p.h, p.l + (a.h * b.l + a.l * b.h)); }
__device__ const double LSD_SINC[14][2] = {{1.0, 0.0}, {-0.16666666666666666, -9.25185853854297e-18}, {0.008333333333333333, 1.1564823173178714e-19}, {-0.0001984126984126984, -1.7209558293420705e-22}, {2.7557319223985893e-06, -1.858393274046472e-22}, {-2.505210838544172e-08, 1.448814070935912e-24}, {1.6059043836821613e-10, 1.2585294588752098e-26}, {-7.647163731819816e-13, -7.03872877733453e-30}, {2.8114572543455206e-15, 1.6508842730861433e-31}, {-8.22063524662433e-18, -2.2141894119604265e-34}, {1.9572941063391263e-20, -1.3643503830087908e-36}, {-3.868170170630684e-23, 8.843177655482344e-40}, {6.446950284384474e-26, -1.9330404233703465e-42}, {-9.183689863795546e-29, -1.4303150396787322e-45}};
__device__ const double LSD_COSC[14][2] = {{1.0, 0.0}, {-0.5, 0.0}, {0.041666666666666664, 2.3129646346357427e-18}, {-0.001388888888888889, 5.300543954373577e-20}, {2.48015873015873e-05, 2.1511947866775882e-23}, {-2.755731922398589e-07, -2.3767714622250297e-23}, {2.08767569878681e-09, -1.20734505911326e-25}, {-1.1470745597729725e-11, -2.0655512752830745e-28}, {4.779477332387385e-14, 4.399205485834081e-31}, {-1.5619206968586225e-16, -1.1910679660273754e-32}, {4.110317623312165e-19, 1.4412973378659527e-36}, {-8.896791392450574e-22, 7.911402614872376e-38}, {1.6117375710961184e-24, -3.6846573564509766e-41}, {-2.4795962632247976e-27, 1.2953730964765229e-43}};
static __device__ __noinline__ void d_sincos_cr(double x, double *s, double *c)
{
    const double PIO2_1 = 1.5707963267948966, PIO2_2 = 6.123233995736766e-17, PIO2_3 = -1.4973849048591698e-33;
    const double kf = rint(x * 0.63661977236758134308);
    const int k = (int)kf;
    const dd_t t1 = dd_two_prod(kf, PIO2_1), t2 = dd_two_prod(kf, PIO2_2);
    dd_t r = dd_two_sum(x, -t1.h);
    r = dd_add(r, dd_t{-t1.l, 0});
    r = dd_add(r, dd_t{-t2.h, -t2.l});
    r = dd_add(r, dd_t{-(kf * PIO2_3), 0});
    const dd_t r2 = dd_mul(r, r);
    dd_t ps = {LSD_SINC[13][0], LSD_SINC[13][1]}, pc = {LSD_COSC[13][0], LSD_COSC[13][1]};
#pragma unroll 1
    for (int i = 12; i >= 0; i--) {
        ps = dd_add(dd_mul(ps, r2), dd_t{LSD_SINC[i][0], LSD_SINC[i][1]});
        pc = dd_add(dd_mul(pc, r2), dd_t{LSD_COSC[i][0], LSD_COSC[i][1]});
    }
    ps = dd_mul(ps, r);
    const double sv = ps.h + ps.l, cv = pc.h + pc.l;
    switch (k & 3) {
    case 0: *s = sv; *c = cv; break;
    case 1: *s = cv; *c = -sv; break;
    case 2: *s = -sv; *c = -cv; break;
    default: *s = -cv; *c = sv; break;
    }
}

// rectangle from a region (cv2 region2rect + get_theta); sums run in region order like cv2.
static __device__ void lsd_region2rect(const u8 *__restrict__ im, const u32 *reg, int n, int W, double reg_angle,
                                       double prec, LsdRect *rec, int lane)
{
    double x = 0, y = 0, sum = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + lane;
        u32 pt = 0;
        double wgt = 0;
        if (i < n) {
            pt = ((volatile const u32 *)reg)[i];
            int gx, gy;
            wgt = sqrt((double)grad_s(im, W, (int)(pt & 0xffff), (int)(pt >> 16), &gx, &gy) / 4.0);
        }
        int cnt = min(32, n - i0);
        for (int j = 0; j < cnt; j++) {
            u32 pj = __shfl_sync(~0u, pt, j);
            double wj = __shfl_sync(~0u, wgt, j);
            x += (double)(int)(pj & 0xffff) * wj;
            y += (double)(int)(pj >> 16) * wj;
            sum += wj;
        }
    }
    x /= sum;
    y /= sum;
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + lane;
        u32 pt = 0;
        double wgt = 0;
        if (i < n) {
            pt = ((volatile const u32 *)reg)[i];
            int gx, gy;
            wgt = sqrt((double)grad_s(im, W, (int)(pt & 0xffff), (int)(pt >> 16), &gx, &gy) / 4.0);
        }
        int cnt = min(32, n - i0);
        for (int j = 0; j < cnt; j++) {
            u32 pj = __shfl_sync(~0u, pt, j);
            double wj = __shfl_sync(~0u, wgt, j);
            double dx = (double)(int)(pj & 0xffff) - x, dy = (double)(int)(pj >> 16) - y;
            Ixx += dy * dy * wj;
            Iyy += dx * dx * wj;
            Ixy -= dx * dy * wj;
        }
    }
    double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)d_fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                           : (double)d_fast_atan2((float)Ixy, (float)(lambda - Iyy));
    theta *= D2R;
    if (fabs(angle_diff_signed(theta, reg_angle)) > prec) theta += 3.14159265358979323846;
    double dx, dy;
    d_sincos_cr(theta, &dy, &dx);
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (int i = lane; i < n; i += 32) {
        u32 pt = ((volatile const u32 *)reg)[i];
        double rdx = (double)(int)(pt & 0xffff) - x, rdy = (double)(int)(pt >> 16) - y;
        double l = rdx * dx + rdy * dy, wv = -rdx * dy + rdy * dx;
        l_max = fmax(l_max, l);
        l_min = fmin(l_min, l);
        w_max = fmax(w_max, wv);
        w_min = fmin(w_min, wv);
    }
    for (int o = 16; o; o >>= 1) {
        l_max = fmax(l_max, __shfl_xor_sync(~0u, l_max, o));
        l_min = fmin(l_min, __shfl_xor_sync(~0u, l_min, o));
        w_max = fmax(w_max, __shfl_xor_sync(~0u, w_max, o));
        w_min = fmin(w_min, __shfl_xor_sync(~0u, w_min, o));
    }
    rec->x1 = x + l_min * dx;
    rec->y1 = y + l_min * dy;
    rec->x2 = x + l_max * dx;
    rec->y2 = y + l_max * dy;
    rec->width = w_max - w_min;
    rec->x = x;
    rec->y = y;
    rec->theta = theta;
    rec->dx = dx;
    rec->dy = dy;
    if (rec->width < 1.0) rec->width = 1.0;
    rec->prec = prec;
    rec->p = prec / 3.14159265358979323846;  // overwritten by the caller with the exact ang_th / 180
}

static __device__ __forceinline__ double dist2d(double x1, double y1, double x2, double y2)
{
    return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
}

// ---- LSD_REFINE_ADV: number of false alarms of a rectangle (cv2 rect_nfa / nfa / rect_improve)
static __device__ double d_log_gamma(double x)
{
    if (x > 15.0)  // Windschitl
        return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424,
                         2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5), b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= log(x + (double)n);
        b += q[n] * pow(x, (double)n);
    }
    return a + log(b);
}

static __device__ double d_nfa(int n, int k, double p, double log_nt)
{
    if (n == 0 || k == 0) return -log_nt;
    if (n == k) return -log_nt - (double)n * log10(p);
    const double p_term = p / (1 - p);
    const double log1term = d_log_gamma((double)n + 1) - d_log_gamma((double)k + 1) - d_log_gamma((double)(n - k) + 1) +
                            (double)k * log(p) + (double)(n - k) * log(1.0 - p);
    double term = exp(log1term);
    // double_equal(term, 0): relative difference against max(|term|, DBL_MIN) within 100 eps
    if (term == 0 || fabs(term) / fmax(fabs(term), 2.2250738585072014e-308) <= 100.0 * 2.220446049250313e-16) {
        if (k > n * p) return -log1term / 2.30258509299404568402 - log_nt;
        return -log_nt;
    }
    double bin_tail = term;
    for (int i = k + 1; i <= n; ++i) {
        const double bin_term = (double)(n - i + 1) / (double)i;
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1) {
            const double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
            if (err < 0.1 * fabs(-log10(bin_tail) - log_nt) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - log_nt;
}

// Pixels of the rectangle as cv2 4.13 counts them (identified from the (points, aligned) pairs its NFA values
// invert to): the topmost vertex (ties: the left one) heads a left and a right chain of two
// edges; an edge crossing no integer row has slope 0; rows ceil(min vertex y) .. ceil(max vertex y); in a row
// ceil(left) .. floor(right), both limits evaluated from a vertex and a slope; the right limit changes edge at
// the first row at or below its vertex, the left limit one row later.  Warp-parallel over the pixels of a row.
static __device__ __noinline__ double lsd_rect_nfa(volatile u32 *an, int W, int H, const LsdRect &rec, double log_nt, int lane)
{
    const double hw = rec.width / 2.0, dyhw = rec.dy * hw, dxhw = rec.dx * hw;
    double vx[4] = {rec.x1 - dyhw, rec.x2 - dyhw, rec.x2 + dyhw, rec.x1 + dyhw};
    double vy[4] = {rec.y1 + dxhw, rec.y2 + dxhw, rec.y2 - dxhw, rec.y1 - dxhw};
    int m = 0;  // the topmost vertex; of two at the same height the left one
    for (int i = 1; i < 4; i++)
        if (vy[i] < vy[m] || (vy[i] == vy[m] && vx[i] < vx[m])) m = i;
    const int a1 = (m + 1) & 3, b1 = (m + 3) & 3, op = (m + 2) & 3;
    const bool ha = vy[m] == vy[a1], hb = vy[m] == vy[b1];
    const double ta = ha ? 0.0 : (vx[m] - vx[a1]) / (vy[m] - vy[a1]), tb = hb ? 0.0 : (vx[m] - vx[b1]) / (vy[m] - vy[b1]);
    const bool a_left = (ha || hb) ? vx[a1] < vx[b1] : ta < tb;
    const int L1 = a_left ? a1 : b1, R1 = a_left ? b1 : a1;
    // an edge that crosses no integer row counts as horizontal: its slope is taken as 0
    auto slope = [&](int p, int q) { return ceil(vy[p]) == ceil(vy[q]) ? 0.0 : (vx[p] - vx[q]) / (vy[p] - vy[q]); };
    const double fl = slope(m, L1), sl = slope(L1, op), fr = slope(m, R1), sr = slope(R1, op);
    const double ya_d = ceil(vy[m]), yb_d = ceil(vy[op]);
    const int ya = (int)fmax(fmin(ya_d, 1e9), -1e9);
    const int yb = min((int)fmax(fmin(yb_d, 1e9), -1e9), H - 1);
    int total = 0, alg = 0;
    for (int y = max(ya, 0); y <= yb; ++y) {
        const bool l2 = (double)(y - 1) >= vy[L1];
        const bool r2 = (double)y >= vy[R1];
        const double lx = l2 ? vx[L1] + ((double)y - vy[L1]) * sl : vx[m] + ((double)y - vy[m]) * fl;
        const double rx = r2 ? vx[R1] + ((double)y - vy[R1]) * sr : vx[m] + ((double)y - vy[m]) * fr;
        const double a_d = ceil(lx), b_d = floor(rx);
        const int a = !(a_d > 0) ? 0 : (int)fmin(a_d, 1e9);
        const int b = !(b_d < W - 1) ? W - 1 : (int)fmax(b_d, -1e9);
        if (b < a) continue;
        total += b - a + 1;
        for (int x = a + lane; x <= b; x += 32) {
            const u32 v = ld_ca(an + (size_t)y * W + x);
            if (v != LSD_NOTDEF && aligned_to(rec.theta, (double)__uint_as_float(v & ~LSD_USED) * D2R, rec.prec)) alg++;
        }
    }
    alg = __reduce_add_sync(~0u, alg);
    return d_nfa(total, alg, rec.p, log_nt);
}

// (kept out of line: inlined, its registers would be charged to the refine 0 / 1 paths of the growing kernels)
static __device__ __noinline__ double lsd_rect_improve(volatile u32 *an, int W, int H, LsdRect &rec, double log_nt, double log_eps,
                                          int lane)
{
    const double delta = 0.5, delta_2 = delta / 2.0;
    double log_nfa = lsd_rect_nfa(an, W, H, rec, log_nt, lane);
    if (log_nfa > log_eps) return log_nfa;
    LsdRect r = rec;  // finer precision
    for (int n = 0; n < 5; ++n) {
        r.p /= 2;
        r.prec = r.p * 3.14159265358979323846;
        const double v = lsd_rect_nfa(an, W, H, r, log_nt, lane);
        if (v > log_nfa) { log_nfa = v; rec = r; }
    }
    if (log_nfa > log_eps) return log_nfa;
    r = rec;  // reduce width
    for (int n = 0; n < 5; ++n) {
        if ((r.width - delta) >= 0.5) {
            r.width -= delta;
            const double v = lsd_rect_nfa(an, W, H, r, log_nt, lane);
            if (v > log_nfa) { rec = r; log_nfa = v; }
        }
    }
    if (log_nfa > log_eps) return log_nfa;
#pragma unroll 1
    for (int side = 0; side < 2; side++) {  // reduce one side, then the other
        const double sg = side == 0 ? 1.0 : -1.0;
        r = rec;
        for (int n = 0; n < 5; ++n) {
            if ((r.width - delta) >= 0.5) {
                r.x1 += sg * (-r.dy * delta_2); r.y1 += sg * (r.dx * delta_2);
                r.x2 += sg * (-r.dy * delta_2); r.y2 += sg * (r.dx * delta_2);
                r.width -= delta;
                const double v = lsd_rect_nfa(an, W, H, r, log_nt, lane);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        }
        if (log_nfa > log_eps) return log_nfa;
    }
    r = rec;  // even finer precision
    for (int n = 0; n < 5; ++n) {
        if ((r.width - delta) >= 0.5) {
            r.p /= 2;
            r.prec = r.p * 3.14159265358979323846;
            const double v = lsd_rect_nfa(an, W, H, r, log_nt, lane);
            if (v > log_nfa) { log_nfa = v; rec = r; }
        }
    }
    return log_nfa;
}

// LSD_REFINE_STD (cv2 refine + reduce_region_radius); returns false if the region is rejected
static __device__ bool lsd_refine(const u8 *__restrict__ im, volatile u32 *an, const float2 *__restrict__ cs, u32 *reg, u32 *ring, int *pn, int W, int H,
                                  double *reg_angle, double prec, LsdRect *rec, double density_th, int lane)
{
    int n = *pn;
    double density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
    if (density >= density_th) return true;
    u32 p0 = ((volatile u32 *)reg)[0];
    const int sx = (int)(p0 & 0xffff), sy = (int)(p0 >> 16);
    const double xc = (double)sx, yc = (double)sy;
    const double ang_c = (double)__uint_as_float(an[(size_t)sy * W + sx] & ~LSD_USED) * D2R;
    double sum = 0, s_sum = 0;
    int cnt = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + lane;
        double ang_d = 0;
        bool in = false;
        if (i < n) {
            u32 pt = ((volatile u32 *)reg)[i];
            int px = (int)(pt & 0xffff), py = (int)(pt >> 16);
            u32 v = an[(size_t)py * W + px] & ~LSD_USED;
            an[(size_t)py * W + px] = v;  // NOTUSED
            if (dist2d(xc, yc, (double)px, (double)py) < rec->width) {
                in = true;
                ang_d = angle_diff_signed((double)__uint_as_float(v) * D2R, ang_c);
            }
        }
        u32 bal = __ballot_sync(~0u, in);
        while (bal) {  // sequential in region order
            int j = __ffs(bal) - 1;
            bal &= bal - 1;
            double dj = __shfl_sync(~0u, ang_d, j);
            sum += dj;
            s_sum += dj * dj;
            ++cnt;
        }
    }
    __syncwarp();
    double mean_angle = sum / (double)cnt;
    double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
    n = lsd_grow(an, cs, reg, ring, W, H, sx, sy, tau, reg_angle, lane);
    *pn = n;
    if (n < 2) return false;
    lsd_region2rect(im, reg, n, W, *reg_angle, prec, rec, lane);
    density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
    if (density >= density_th) return true;
    // reduce_region_radius
    double radSq1 = (xc - rec->x1) * (xc - rec->x1) + (yc - rec->y1) * (yc - rec->y1);
    double radSq2 = (xc - rec->x2) * (xc - rec->x2) + (yc - rec->y2) * (yc - rec->y2);
    double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
    while (density < density_th) {
        radSq *= 0.75 * 0.75;
        // swap-with-last removal in cv2's exact order (sequential; regions here are small)
        if (lane == 0) {
            for (int i = 0; i < n; i++) {
                u32 pt = reg[i];
                int px = (int)(pt & 0xffff), py = (int)(pt >> 16);
                double ddx = (double)px - xc, ddy = (double)py - yc;
                if (ddx * ddx + ddy * ddy > radSq) {
                    an[(size_t)py * W + px] &= ~LSD_USED;
                    reg[i] = reg[n - 1];
                    reg[n - 1] = pt;
                    n--;
                    i--;
                }
            }
        }
        n = __shfl_sync(~0u, n, 0);
        __syncwarp();
        *pn = n;
        if (n < 2) return false;
        lsd_region2rect(im, reg, n, W, *reg_angle, prec, rec, lane);
        density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
    }
    return true;
}

__global__ void __launch_bounds__(32) k_lsd_grow(const u8 *__restrict__ img, u32 *__restrict__ ang,
                                                 const float2 *__restrict__ csbuf,
                                                 const u32 *__restrict__ order, u32 *__restrict__ regbuf,
                                                 const int *__restrict__ counters, LsdParams P,
                                                 float *__restrict__ segs, double *__restrict__ widths,
                                                 double *__restrict__ precs, double *__restrict__ nfas,
                                                 int32_t *__restrict__ counts)
{
    const int f = blockIdx.x, lane = threadIdx.x;
    const int W = P.W, H = P.H;
    const size_t npx = (size_t)H * W;
    const u8 *im = img + (size_t)f * npx;
    volatile u32 *an = ang + (size_t)f * npx;
    const float2 *cs = csbuf + (size_t)f * npx;
    const u32 *ord = order + (size_t)f * npx;
    u32 *reg = regbuf + (size_t)f * npx;
    const int ndef = counters[f * LE_C_STRIDE + LE_C_AUX0];
    int nseg = 0;
    __shared__ u32 s_ring[LSD_RING];
    for (int i0 = 0; i0 < ndef; i0 += 32) {
        // 32 seeds per round: their USED flags are fetched together.  A flag can only flip to USED, so a
        // seed seen USED is skipped for good; one seen free is re-read only if a region grew since.
        u32 mine = (i0 + lane < ndef) ? ord[i0 + lane] : 0;
        u32 myv = (i0 + lane < ndef) ? ld_ca(an + mine) : LSD_USED;
        bool dirty = false;
        int cnt = min(32, ndef - i0);
        for (int j = 0; j < cnt; j++) {
            u32 pidx = __shfl_sync(~0u, mine, j);
            u32 v = __shfl_sync(~0u, myv, j);
            if (v & LSD_USED) continue;  // (NOTDEF has bit 31 set too, but only defined pixels are listed)
            if (dirty) {
                v = ld_ca(an + pidx);
                if (v & LSD_USED) continue;
            }
            dirty = true;
            int sy = (int)(pidx / (u32)W), sx = (int)(pidx - (u32)sy * W);
            double reg_angle;
            int n = lsd_grow(an, cs, reg, s_ring, W, H, sx, sy, P.prec, &reg_angle, lane);
            if ((unsigned)n < P.min_reg_size) continue;
            LsdRect rec;
            lsd_region2rect(im, reg, n, W, reg_angle, P.prec, &rec, lane);
            if (P.refine > 0) {
                if (!lsd_refine(im, an, cs, reg, s_ring, &n, W, H, &reg_angle, P.prec, &rec, P.density_th, lane)) continue;
            }
            rec.prec = P.prec;
            rec.p = P.p;
            double log_nfa = -1;
            if (P.refine >= 2) {
                log_nfa = lsd_rect_improve(an, W, H, rec, P.log_nt, P.log_eps, lane);
                if (log_nfa <= P.log_eps) continue;
            }
            rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
            if (P.scale != 1) {
                rec.x1 /= P.scale; rec.y1 /= P.scale; rec.x2 /= P.scale; rec.y2 /= P.scale;
                rec.width /= P.scale;
            }
            if (lane == 0 && nseg < P.max_segs) {
                float *o = segs + ((size_t)f * P.max_segs + nseg) * 4;
                o[0] = (float)rec.x1; o[1] = (float)rec.y1; o[2] = (float)rec.x2; o[3] = (float)rec.y2;
                if (widths) widths[(size_t)f * P.max_segs + nseg] = rec.width;
                if (precs) precs[(size_t)f * P.max_segs + nseg] = rec.p;
                if (nfas) nfas[(size_t)f * P.max_segs + nseg] = log_nfa;
            }
            nseg++;
        }
    }
    if (lane == 0) counts[f] = nseg;
}

// ------------------------------------------------------------------ 4b: speculative region growing
// K warps per frame grow the next K free seeds of the ordered list AT THE SAME TIME, then warp 0 commits them in
// seed order, so the result is the sequential one bit for bit:
//  * nothing is marked USED while growing.  A warp tags the pixels it accepts in a separate plane
//    (own[p] = round, warp id); plain stores, the last writer wins.
//  * a pixel tagged in this round by a LOWER id is treated as taken (in seed order that region comes first) and
//    the id is remembered (relied mask); a pixel tagged by a HIGHER id is taken over.
//  * commit, in id order: a seed that an earlier region of the round swallowed is void (exactly what the
//    sequential scan does with a USED seed).  Region w is valid iff none of its pixels is USED by then (all
//    earlier regions of the round are committed) and no id it relied on turned out void / invalid / refined.  The first region of a round relies on
//    nobody and is never validated: it IS the sequential next region, so every round makes progress.  At the
//    first invalid region the commit stops and the next round starts from that seed.
//  * density refinement (refine >= 1) re-grows with another tolerance: only the first region of a round may do
//    it (through the sequential code path); any other region that needs it waits to be first.
#define SG_KMAX 16
#ifdef LSD_STATS
__device__ int g_lsd_stats[16];
__device__ long long g_lsd_cyc[8];
#endif
#define SG_ENT 96  // free seeds examined per round (slots + deferred)
#define SG_ROUND_LIM ((1u << 20) - 1)  // rounds that still fit the tag
struct SgSlot {
    u32 seed;     // pixel index of the seed
    int ord_idx;  // its position in the ordered list
    int n;        // region size (speculative)
    u32 relied;   // ids of this round whose tags made this region skip a pixel
    int serial;   // 1: must be re-run as first of a round (region list overflow)
    int need_refine;
    double reg_angle, log_nfa;
    LsdRect rec;
};

static __device__ __forceinline__ u32 ld_cg(const u32 *p)
{
    u32 v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

#ifndef SG_HASH
#define SG_HASH 1024                 // per-warp exact set of the region's pixels (open addressing, shared memory)
#define SG_HASH_CAP 768              // a larger region waits to be the first of a round
#endif
static __device__ __forceinline__ u32 sg_hash(u32 e) { return ((e * 2654435761u) >> 16) & (SG_HASH - 1); }

static __device__ int lsd_grow_spec(volatile u32 *an, volatile u32 *own, const float2 *__restrict__ cs, u32 *reg,
                                    int reg_cap, u32 *ring, u32 *hset, int W, int H, int sx, int sy, double tol,
                                    double *reg_angle_out, int lane, u32 round, int wid, u32 *relied_out,
                                    int *overflow)
{
    const u32 tag = ((round << 5) | (u32)wid) + 1u;
    int n = 0;
    u32 relied = 0;
    u32 sv = ld_ca(an + (size_t)sy * W + sx) & ~LSD_USED;
    double reg_angle = (double)__uint_as_float(sv) * D2R;
    for (int j = lane; j < SG_HASH; j += 32) hset[j] = 0xFFFFFFFFu;
    __syncwarp();
    if (lane == 0) {
        u32 e = ((u32)sy << 16) | (u32)sx;
        reg[0] = e;
        ring[0] = e;
        own[(size_t)sy * W + sx] = tag;
        hset[sg_hash(e)] = e;
    }
    n = 1;
    float sumdx = (float)cos(reg_angle), sumdy = (float)sin(reg_angle);
    __syncwarp();
    const int g = lane / 9, kk = lane - 9 * g;  // queue entry of the trip, neighbour of that entry (see lsd_grow)
    const int ky = kk / 3 - 1, kx = kk % 3 - 1;
    bool full = false;
    for (int i = 0; i < n && !full;) {
        const int m = min(LSD_FRONT, n - i);
        int px = 0, py = 0;
        if (g < m) {
            const int qi = i + g;
            const u32 pt = (n - qi <= LSD_RING) ? ((volatile u32 *)ring)[qi & (LSD_RING - 1)] : ((volatile u32 *)reg)[qi];
            px = (int)(pt & 0xffff);
            py = (int)(pt >> 16);
        }
        const int xx = px + kx, yy = py + ky;
        u32 v = LSD_NOTDEF, o = 0;
        float2 csn = make_float2(0.f, 0.f);
        if (g < m && xx >= 0 && xx < W && yy >= 0 && yy < H) {
            v = ld_ca(an + (size_t)yy * W + xx);
            o = ld_ca(own + (size_t)yy * W + xx);
            csn = __ldg(cs + (size_t)yy * W + xx);
        }
        bool avail = v != LSD_NOTDEF && !(v & LSD_USED);
        const bool cur = o != 0 && ((o - 1u) >> 5) == round;
        const int oid = (int)((o - 1u) & 31u);
        const bool taken = cur && oid <= wid;  // mine already, or an earlier region of this round
        relied |= __reduce_or_sync(~0u, (avail && cur && oid < wid) ? (1u << oid) : 0u);
        // A tag of a HIGHER id says nothing about this region: that warp may have overwritten this region's own
        // tag (plain stores race).  The region's exact pixel set (shared-memory hash) settles it.
        bool mine = false;
        if (avail && cur && oid > wid) {
            const u32 e = ((u32)yy << 16) | (u32)xx;
            for (u32 hx = sg_hash(e);; hx = (hx + 1) & (SG_HASH - 1)) {
                const u32 t = ((volatile u32 *)hset)[hx];
                if (t == e) { mine = true; break; }
                if (t == 0xFFFFFFFFu) break;
            }
        }
        avail = avail && !taken && !mine;
        const double a_l = (double)__uint_as_float(v) * D2R;  // this lane's neighbour's level-line angle
        for (int gg = 0; gg < m && !full; gg++) {
            u32 remaining = (__ballot_sync(~0u, avail) >> (9 * gg)) & 0x1ffu;  // (see lsd_grow)
            const int cpx = __shfl_sync(~0u, px, 9 * gg), cpy = __shfl_sync(~0u, py, 9 * gg);
            while (remaining) {
                const bool al = aligned_to(reg_angle, a_l, tol) && avail;
                const u32 mm = (__ballot_sync(~0u, al) >> (9 * gg)) & remaining;
                if (!mm) break;
                if (n >= reg_cap || n >= SG_HASH_CAP) { full = true; break; }
                const int k = __ffs(mm) - 1;
                remaining &= ~((2u << k) - 1u);
                const int src = 9 * gg + k;
                int ax = cpx + (k % 3) - 1, ay = cpy + (k / 3) - 1;
                if (lane == 0) {
                    u32 e = ((u32)ay << 16) | (u32)ax;
                    own[(size_t)ay * W + ax] = tag;
                    reg[n] = e;
                    ring[n & (LSD_RING - 1)] = e;
                    u32 hx = sg_hash(e);
                    while (hset[hx] != 0xFFFFFFFFu) hx = (hx + 1) & (SG_HASH - 1);
                    hset[hx] = e;
                }
                n++;
                sumdx += __shfl_sync(~0u, csn.x, src);
                sumdy += __shfl_sync(~0u, csn.y, src);
                reg_angle = (double)d_fast_atan2(sumdy, sumdx) * D2R;
                if (xx == ax && yy == ay) avail = false;  // also a neighbour of a later entry of this trip
            }
        }
        i += m;
        __syncwarp();
    }
    *reg_angle_out = reg_angle;
    *relied_out = relied;
    *overflow = full ? 1 : 0;
    return n;
}

// ADV: LSD_REFINE_ADV code compiled in (kept out of the refine 0 / 1 instantiations: registers)
template <int SG_K, bool ADV>
__global__ void __launch_bounds__(SG_K * 32, SG_K == 8 ? 4 : 1) k_lsd_grow_spec(const u8 *__restrict__ img, u32 *__restrict__ ang,
                                                             u32 *__restrict__ ownbuf,
                                                             const float2 *__restrict__ csbuf,
                                                             const u32 *__restrict__ order, u32 *__restrict__ regbuf,
                                                             const int *__restrict__ counters, LsdParams P,
                                                             float *__restrict__ segs, double *__restrict__ widths,
                                                             double *__restrict__ precs, double *__restrict__ nfas,
                                                             int32_t *__restrict__ counts, u32 epoch)
{
    const int f = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int W = P.W, H = P.H;
    const size_t npx = (size_t)H * W;
    const u8 *im = img + (size_t)f * npx;
    volatile u32 *an = ang + (size_t)f * npx;
    volatile u32 *own = ownbuf + (size_t)f * npx;
    const float2 *cs = csbuf + (size_t)f * npx;
    const u32 *ord = order + (size_t)f * npx;
    // region lists: warp 0 owns a full-size list (a region can be the whole frame), the others share a second one
    u32 *reg0 = regbuf + (size_t)f * 2 * npx;
    const int cap_w = (int)(npx / (SG_K - 1));
    u32 *reg = wid == 0 ? reg0 : reg0 + npx + (size_t)(wid - 1) * cap_w;
    const int reg_cap = wid == 0 ? (int)npx : cap_w;
    const int ndef = counters[f * LE_C_STRIDE + LE_C_AUX0];
    extern __shared__ __align__(16) u32 s_dyn[];  // [SG_K][LSD_RING] rings, then [SG_K][SG_HASH] pixel sets
    u32(*s_ring)[LSD_RING] = (u32(*)[LSD_RING])s_dyn;
    u32(*s_hset)[SG_HASH] = (u32(*)[SG_HASH])(s_dyn + SG_K * LSD_RING);
    __shared__ SgSlot s_slot[SG_K];
    // the round's free seeds in list order: a slot id, or -1 = deferred (it touches an earlier seed of the round,
    // so that seed's region almost surely swallows it; it is only looked at again at commit time)
    __shared__ u32 s_ent_seed[SG_ENT], s_ent_xy[SG_ENT];
    __shared__ int s_ent_ord[SG_ENT], s_ent_kind[SG_ENT];
    __shared__ int s_pos, s_k, s_ne, s_nseg;
    // pre-commit results: per slot (bit 0: a pixel is USED, bit 1: the seed is), the earlier regions of the round
    // that share a pixel with it / hold its seed; per deferred seed the same two facts; and who marks below
    __shared__ int s_pre[SG_K], s_act[SG_K], s_def_used[SG_ENT];
    __shared__ u32 s_ovl[SG_K], s_sw[SG_K], s_def_sw[SG_ENT];
    if (threadIdx.x == 0) { s_pos = 0; s_nseg = 0; }
    if (threadIdx.x < SG_K) s_act[threadIdx.x] = 0;
    __syncthreads();
#ifdef LSD_STATS
    long long t_sel = 0, t_grow = 0, t_com = 0, t0 = 0;
#endif
    for (u32 round = 0;; round++) {
        // tags carry (call epoch, round, warp): the tag plane is cleared once per 126 calls instead of per call.
        // 20 bits of round; a frame that needs more rounds (more than a million seeds taken one at a time) goes on
        // with one region per round, which uses no tags
        const int kmax = round < SG_ROUND_LIM ? SG_K : 1;
#ifdef LSD_STATS
        t0 = clock64();
#endif
        // ---- the next free seeds in order (warp 0): SG_K of them get a warp, the ones touching an earlier seed
        // of the round are deferred
        if (wid == 0) {
            int k = 0, ne = 0, pos = s_pos;
            bool stop = false;
            while (!stop && pos < ndef) {
                // 128 list entries per trip: four seeds per lane, their USED flags fetched together
                u32 pidx4[4];
                bool fr4[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = pos + 32 * q + lane;
                    pidx4[q] = i < ndef ? __ldg(ord + i) : 0;
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = pos + 32 * q + lane;
                    fr4[q] = i < ndef && !(ld_cg((const u32 *)an + pidx4[q]) & LSD_USED);
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    u32 m = __ballot_sync(~0u, fr4[q]);
                    while (m && !stop) {
                        if (k == kmax || ne == SG_ENT) { stop = true; break; }
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        const u32 pj = __shfl_sync(~0u, pidx4[q], j);
                        const int y = (int)(pj / (u32)W), x = (int)(pj - (u32)y * W);
                        bool close = false;
                        for (int e = lane; e < ne; e += 32) {
                            const u32 qq = s_ent_xy[e];
                            close = close || (abs((int)(qq & 0xffff) - x) <= 1 && abs((int)(qq >> 16) - y) <= 1);
                        }
                        close = __any_sync(~0u, close);
                        if (lane == 0) {
                            s_ent_seed[ne] = pj;
                            s_ent_xy[ne] = ((u32)y << 16) | (u32)x;
                            s_ent_ord[ne] = pos + 32 * q + j;
                            s_ent_kind[ne] = close ? -1 : k;
                            if (!close) {
                                s_slot[k].seed = pj;
                                s_slot[k].ord_idx = pos + 32 * q + j;
                            }
                        }
                        __syncwarp();
                        if (!close) k++;
                        ne++;
                    }
                }
                pos += 128;
            }
            if (lane == 0) { s_k = k; s_ne = ne; }
        }
        __syncthreads();
#ifdef LSD_STATS
        t_sel += clock64() - t0; t0 = clock64();
#endif
        const int kc = s_k;
        if (kc == 0) break;
        // ---- speculative growth, one warp per seed
        if (wid < kc) {
            SgSlot *S = &s_slot[wid];
            const u32 pidx = S->seed;
            const int sy = (int)(pidx / (u32)W), sx = (int)(pidx - (u32)sy * W);
            double reg_angle;
            u32 relied = 0;
            int overflow = 0;
            // the first region of a round is the sequential next region: it marks USED as it grows (any size, no
            // tags); the others see its pixels as used -- late at worst, which the commit catches
            int n = wid == 0 ? lsd_grow(an, cs, reg, s_ring[0], W, H, sx, sy, P.prec, &reg_angle, lane)
                             : lsd_grow_spec(an, own, cs, reg, reg_cap, s_ring[wid], s_hset[wid], W, H, sx, sy, P.prec,
                                             &reg_angle, lane, (epoch << 20) | round, wid, &relied, &overflow);
            int need_refine = 0;
            LsdRect rec;
            if (!overflow && (unsigned)n >= P.min_reg_size) {
                lsd_region2rect(im, reg, n, W, reg_angle, P.prec, &rec, lane);
                if (P.refine > 0) {
                    double density = (double)n / (dist2d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
                    need_refine = density < P.density_th;
                }
                rec.prec = P.prec;
                rec.p = P.p;
                double log_nfa = -1;
                // the NFA only reads angles, never USED flags: it can be evaluated before the commit (unless the
                // region is going to be refined first)
                if (ADV && !need_refine) log_nfa = lsd_rect_improve(an, W, H, rec, P.log_nt, P.log_eps, lane);
                if (lane == 0) { S->rec = rec; S->log_nfa = log_nfa; }
            }
            if (lane == 0) {
                S->n = n;
                S->relied = relied;
                S->serial = overflow;
                S->need_refine = need_refine;
                S->reg_angle = reg_angle;
            }
        }
        __syncthreads();
#ifdef LSD_STATS
        t_grow += clock64() - t0; t0 = clock64();
#endif
        // ---- pre-commit, every warp for its own region (round 1 did all of this on warp 0, region after region,
        // three dependent L2 round trips each: 29 % of the kernel).  After the barrier above every region of the
        // round is grown, region 0 has marked its pixels and every pixel set (hash) is complete, so a region can
        // find out by itself (a) whether any of its pixels -- the seed first -- is USED already, and (b) which
        // earlier regions of the round share a pixel with it or hold its seed.  What is still unknown is which of
        // those regions will COMMIT; that is decided in list order below from these bit masks alone.
        auto in_set = [&](int v, u32 e) -> bool {
            for (u32 hx = sg_hash(e);; hx = (hx + 1) & (SG_HASH - 1)) {
                const u32 t = ((volatile u32 *)s_hset[v])[hx];
                if (t == e) return true;
                if (t == 0xFFFFFFFFu) return false;
            }
        };
        if (wid > 0 && wid < kc && !s_slot[wid].serial) {
            const int n = s_slot[wid].n;
            bool used_any = false, seed_used = false;
            u32 ovl = 0, sw = 0;
            for (int base = 0; base < n; base += 128) {
                u32 pt[4], av[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = base + 32 * q + lane;
                    pt[q] = i < n ? ((n - i <= LSD_RING) ? ((volatile u32 *)s_ring[wid])[i & (LSD_RING - 1)] : ld_cg(reg + i))
                                  : 0xFFFFFFFFu;
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    av[q] = pt[q] != 0xFFFFFFFFu ? ld_cg((const u32 *)an + (size_t)(pt[q] >> 16) * W + (pt[q] & 0xffff)) : 0u;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (pt[q] == 0xFFFFFFFFu) continue;
                    const bool first = base + 32 * q + lane == 0;
                    if (av[q] & LSD_USED) {
                        used_any = true;
                        seed_used = seed_used || first;
                    }
                    for (int v = 1; v < wid; v++) {
                        if (s_slot[v].serial || !in_set(v, pt[q])) continue;
                        ovl |= 1u << v;
                        if (first) sw |= 1u << v;
                    }
                }
            }
            used_any = __any_sync(~0u, used_any);
            seed_used = __any_sync(~0u, seed_used);
            ovl = __reduce_or_sync(~0u, ovl);
            sw = __reduce_or_sync(~0u, sw);
            if (lane == 0) {
                s_pre[wid] = (used_any ? 1 : 0) | (seed_used ? 2 : 0);
                s_ovl[wid] = ovl;
                s_sw[wid] = sw;
            }
        }
        if (wid == 0) {  // the deferred seeds: USED already, or inside one of the round's regions
            const int ne = s_ne;
            for (int e = lane; e < ne; e += 32) {
                if (s_ent_kind[e] >= 0) continue;
                u32 m = 0;
                for (int v = 1; v < kc; v++)
                    if (!s_slot[v].serial && in_set(v, s_ent_xy[e])) m |= 1u << v;
                s_def_sw[e] = m;
                s_def_used[e] = (ld_cg((const u32 *)an + s_ent_seed[e]) & LSD_USED) ? 1 : 0;
            }
        }
        __syncthreads();
        // ---- ordered decisions (warp 0; no memory round trips except for a refinement)
        if (wid == 0) {
            u32 bad = 0, committed = 0;
            int nseg = s_nseg;
            const int ne = s_ne;
            int e = 0;
            for (; e < ne; e++) {
                const int w = s_ent_kind[e];
                if (w < 0) {  // deferred seed: swallowed by now (the usual case), or the round ends here
                    if (s_def_used[e] || (s_def_sw[e] & committed)) continue;
                    break;
                }
                const SgSlot *S = &s_slot[w];
                if (S->serial && w > 0) break;
                if (w > 0) {
                    if ((s_pre[w] & 2) || (s_sw[w] & committed)) {  // swallowed by an earlier region of this round
                        bad |= 1u << w;
                        continue;
                    }
                    if (S->relied & bad) break;
                    if (S->need_refine) break;
                    // every pixel must still be free: not USED when the round's growth ended, and in no region
                    // that committed before this one
                    if ((s_pre[w] & 1) || (s_ovl[w] & committed)) break;
                    committed |= 1u << w;
                    if (lane == 0) s_act[w] = 1;  // the owning warp marks the pixels below
                }
                int n = S->n;
                if ((unsigned)n < P.min_reg_size) continue;
                LsdRect rec = S->rec;
                double reg_angle = S->reg_angle;
                bool refined = false;
                if (S->need_refine) {  // only the first region of a round: the sequential refinement path.  Its
                    // pixel set changes, so the masks above are stale: the round ends behind this region (what the
                    // entries behind it are -- swallowed, void or free -- the next round's seed scan finds out)
                    refined = true;
                    if (!lsd_refine(im, an, cs, reg0, s_ring[0], &n, W, H, &reg_angle, P.prec, &rec, P.density_th, lane)) {
                        e++;
                        break;
                    }
                    rec.prec = P.prec;
                    rec.p = P.p;
                }
                double log_nfa = S->log_nfa;
                bool emit = true;
                if (ADV) {
                    if (S->need_refine) log_nfa = lsd_rect_improve(an, W, H, rec, P.log_nt, P.log_eps, lane);
                    emit = log_nfa > P.log_eps;
                }
                if (emit) {
                    rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
                    if (P.scale != 1) {
                        rec.x1 /= P.scale; rec.y1 /= P.scale; rec.x2 /= P.scale; rec.y2 /= P.scale;
                        rec.width /= P.scale;
                    }
                    if (lane == 0 && nseg < P.max_segs) {
                        float *o = segs + ((size_t)f * P.max_segs + nseg) * 4;
                        o[0] = (float)rec.x1; o[1] = (float)rec.y1; o[2] = (float)rec.x2; o[3] = (float)rec.y2;
                        if (widths) widths[(size_t)f * P.max_segs + nseg] = rec.width;
                        if (precs) precs[(size_t)f * P.max_segs + nseg] = rec.p;
                        if (nfas) nfas[(size_t)f * P.max_segs + nseg] = log_nfa;
                    }
                    nseg++;
                }
                if (refined) {
                    e++;
                    break;
                }
            }
            if (lane == 0) {
                s_nseg = nseg;
                s_pos = e < ne ? s_ent_ord[e] : s_ent_ord[ne - 1] + 1;
            }
        }
        __syncthreads();
        // ---- the committed regions mark their pixels USED, each warp its own (plain stores of the angle word, like
        // the sequential path: the frame's angle plane is read through L1 by this SM only)
        if (wid > 0 && wid < kc && s_act[wid]) {
            const int n = s_slot[wid].n;
            for (int base = 0; base < n; base += 128) {
                u32 pt[4], av[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = base + 32 * q + lane;
                    pt[q] = i < n ? ((n - i <= LSD_RING) ? ((volatile u32 *)s_ring[wid])[i & (LSD_RING - 1)] : ld_cg(reg + i))
                                  : 0xFFFFFFFFu;
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    av[q] = pt[q] != 0xFFFFFFFFu ? ld_cg((const u32 *)an + (size_t)(pt[q] >> 16) * W + (pt[q] & 0xffff)) : 0u;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (pt[q] != 0xFFFFFFFFu) an[(size_t)(pt[q] >> 16) * W + (pt[q] & 0xffff)] = av[q] | LSD_USED;
            }
            __syncwarp();
            if (lane == 0) s_act[wid] = 0;
        }
        __syncthreads();
#ifdef LSD_STATS
        t_com += clock64() - t0;
#endif
    }
    if (threadIdx.x == 0) counts[f] = s_nseg;
#ifdef LSD_STATS
    if (threadIdx.x == 0 && f == 0) printf("LSDCYCLES select %lld grow %lld commit %lld (deferred %lld seedchk %lld validate %lld mark %lld out %lld) w0-before-sync %lld w1-before-sync %lld\n", t_sel, t_grow, t_com, g_lsd_cyc[0], g_lsd_cyc[1], g_lsd_cyc[2], g_lsd_cyc[3], g_lsd_cyc[4], g_lsd_cyc[5], g_lsd_cyc[6]);
    if (threadIdx.x == 0 && f == 0)
        printf("LSDSTATS rounds %d slots %d entries %d commits(w>0) %d voids %d | end: all-resolved %d deferred-free %d serial %d relied %d refine %d used %d\n",
               g_lsd_stats[0], g_lsd_stats[1], g_lsd_stats[2], g_lsd_stats[3], g_lsd_stats[4], g_lsd_stats[5], g_lsd_stats[6],
               g_lsd_stats[7], g_lsd_stats[8], g_lsd_stats[9], g_lsd_stats[10]);
#endif
}

// ------------------------------------------------------------------ host side
extern "C" int le_lsd_scaled_dims(int h, int w, double scale, int *sh, int *sw)
{
    if (!sh || !sw || h <= 0 || w <= 0 || !(scale > 0)) return LE_ERR_INVALID;
    *sw = scale != 1 ? (int)lrint(w * scale) : w;
    *sh = scale != 1 ? (int)lrint(h * scale) : h;
    return LE_OK;
}

static void gauss_taps_q8(double sigma, int k, std::vector<int> &taps)
{
    std::vector<double> ker(k);
    double sum = 0, scale2x = -0.5 / (sigma * sigma);
    for (int i = 0; i < k; i++) {
        double x = i - (k - 1) * 0.5;
        ker[i] = exp(scale2x * x * x);
        sum += ker[i];
    }
    for (int i = 0; i < k; i++) ker[i] /= sum;
    taps.assign(k, 0);
    double err = 0;
    long tot = 0;
    for (int i = 0; i < k / 2; i++) {
        double adj = ker[i] * 256.0 + err;
        long v0 = lrint(adj);
        err = adj - (double)v0;
        taps[i] = taps[k - 1 - i] = (int)v0;
        tot += v0;
    }
    taps[k / 2] = (int)(256 - 2 * tot);
}

static void resize_tab(int src, int dst, double inv_scale, std::vector<int2> &tab)
{
    tab.resize(dst);
    double sc = 1.0 / inv_scale;
    for (int d = 0; d < dst; d++) {
        double pos = (d + 0.5) * sc - 0.5;
        int i = (int)floor(pos);
        double fr = pos - i;
        if (i < 0) { i = 0; fr = 0; }
        if (i >= src - 1) { i = src - 1; fr = 0; }
        tab[d] = make_int2(i, (int)lrint(fr * 256.0));
    }
}

__global__ void k_lsd_zero(int *counters, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        counters[i * LE_C_STRIDE + LE_C_MAX] = 0;
        counters[i * LE_C_STRIDE + LE_C_AUX0] = 0;
    }
}

extern "C" int le_lsd(le_ctx *ctx, const uint8_t *gray, int n, int h, int w, int refine, double scale,
                      double sigma_scale, double quant, double ang_th, double log_eps, double density_th, int n_bins,
                      float *segs, double *width, double *prec_out, double *nfa_out, int32_t *counts, int max_segs,
                      void *stream)
{
    LE_ON_DEVICE(ctx);
    if (!ctx) return LE_ERR_INVALID;
    LE_REQUIRE(ctx, gray && segs && counts, "null pointer");
    LE_REQUIRE(ctx, n > 0 && h > 1 && w > 1 && max_segs > 0, "n, h, w, max_segs must be positive");
    LE_REQUIRE(ctx, h < 65536 && w < 65536, "frame too large");
    LE_REQUIRE(ctx, scale > 0 && sigma_scale > 0 && quant >= 0 && ang_th > 0 && ang_th < 180, "bad LSD parameters");
    LE_REQUIRE(ctx, n_bins > 0 && n_bins <= 8192, "n_bins must be in 1..8192");
    if (refine < 0 || refine > 2) return le_fail(ctx, LE_ERR_UNSUPPORTED, "LSD refine=%d is not a cv2 mode", refine);
    cudaStream_t st = (cudaStream_t)stream;
    int H = h, W = w;
    le_lsd_scaled_dims(h, w, scale, &H, &W);
    LE_REQUIRE(ctx, H > 1 && W > 1, "scaled image too small");
    const size_t npx = (size_t)H * W;
    const int nchunk = (int)((npx + LS_CHUNK - 1) / LS_CHUNK);
    int rc;
    if ((rc = le_ensure(ctx, &ctx->counters, sizeof(int) * LE_C_STRIDE * (size_t)n))) return rc;
    if ((rc = le_ensure(ctx, &ctx->lsd1, npx * n))) return rc;                                 // scaled image
    if ((rc = le_ensure(ctx, &ctx->lsd2, sizeof(u32) * npx * n))) return rc;                   // angles
    if ((rc = le_ensure(ctx, &ctx->lsd3, sizeof(u32) * npx * n))) return rc;                   // order
    const int spec = LE_KNOB_SET("LE_LSD_SERIAL") ? 0 : 1;  // development knob: the one-warp-per-frame kernel
    if ((rc = le_ensure(ctx, &ctx->lsd4, sizeof(u32) * npx * n * (spec ? 2 : 1)))) return rc;  // region lists
    if (spec) {                                                                                // speculative tags
        const void *before = ctx->lsd7.p;
        if ((rc = le_ensure(ctx, &ctx->lsd7, sizeof(u32) * npx * n))) return rc;
        if (ctx->lsd7.p != before) ctx->lsd_epoch = 0;  // a new buffer comes zeroed
    }
    if ((rc = le_ensure(ctx, &ctx->lsd6, sizeof(float2) * npx * n))) return rc;                // cos/sin of the angles
    if ((rc = le_ensure(ctx, &ctx->lsd5, (sizeof(u16) + sizeof(u32)) * (size_t)nchunk * n_bins * n + 16))) return rc;
    ctx->edge_list_for = nullptr;
    const u8 *img = gray;
    LE_T0(ctx, LE_K_LSD_PREP, st);
    if (scale != 1) {
        const double sigma = (scale < 1) ? (sigma_scale / scale) : sigma_scale;
        const unsigned hh = (unsigned)ceil(sigma * sqrt(2 * 3.0 * log(10.0)));
        const int k = 1 + 2 * (int)hh;
        LE_REQUIRE(ctx, hh <= SB_RMAX, "Gaussian kernel too large (sigma_scale / scale)");
        std::vector<int> taps;
        gauss_taps_q8(sigma, k, taps);
        std::vector<int2> xt, yt;
        resize_tab(w, W, scale, xt);
        resize_tab(h, H, scale, yt);
        size_t tab_bytes = 128 + sizeof(int2) * ((size_t)W + H);
        if ((rc = le_ensure(ctx, &ctx->lsd_tab, tab_bytes))) return rc;
        const bool wordwise = (w & 3) == 0 && (((size_t)gray) & 3) == 0 && (hh == 3 || hh == 5);
        // fused blur + resize: the largest destination tile whose samples fit the 128 x 32 shared tile
        int dtw = 0, dth = 0;
        const int unfused = LE_KNOB_SET("LE_LSD_UNFUSED") ? 1 : 0;  // development knob: the two-kernel path
        if (wordwise && scale < 1 && !unfused) {
            auto fits = [](const std::vector<int2> &tab, int src, int dt, int cap, int align) {
                for (int d0 = 0; d0 < (int)tab.size(); d0 += dt) {
                    const int last = std::min(d0 + dt, (int)tab.size()) - 1;
                    if (std::min(tab[last].x + 1, src - 1) - (tab[d0].x & ~align) >= cap) return false;
                }
                return true;
            };
            for (dtw = std::min(W, (int)((SV_TW - 5) * scale)) & ~3; dtw >= 16 && !fits(xt, w, dtw, SV_TW, 3); dtw -= 4) {}
            for (dth = std::min(H, (int)((SV_TH - 2) * scale)); dth >= 4 && !fits(yt, h, dth, SV_TH, 0); dth--) {}
            if (dtw < 16 || dth < 4) dtw = dth = 0;
        }
        if (!dtw && (rc = le_ensure(ctx, &ctx->lsd0, (size_t)h * w * n))) return rc;  // blurred full-res
        char *tb = (char *)ctx->lsd_tab.p;
        if (!(ctx->lsd_key[0] == h && ctx->lsd_key[1] == w && ctx->lsd_key[2] == scale && ctx->lsd_key[3] == sigma_scale)) {
            // pageable-source copies are staged before cudaMemcpyAsync returns
            LE_CUDA(ctx, cudaMemcpyAsync(tb, taps.data(), sizeof(int) * k, cudaMemcpyHostToDevice, st));
            LE_CUDA(ctx, cudaMemcpyAsync(tb + 128, xt.data(), sizeof(int2) * W, cudaMemcpyHostToDevice, st));
            LE_CUDA(ctx, cudaMemcpyAsync(tb + 128 + sizeof(int2) * W, yt.data(), sizeof(int2) * H, cudaMemcpyHostToDevice, st));
            ctx->lsd_key[0] = h; ctx->lsd_key[1] = w; ctx->lsd_key[2] = scale; ctx->lsd_key[3] = sigma_scale;
        }
        const dim3 gw((w + SV_TW - 1) / SV_TW, (h + SV_TH - 1) / SV_TH, n);
        const int2 *xtd = (const int2 *)(tb + 128), *ytd = (const int2 *)(tb + 128 + sizeof(int2) * W);
        if (dtw) {
            const dim3 gf((W + dtw - 1) / dtw, (H + dth - 1) / dth, n);
            if (hh == 3)
                k_blur_resize_w<3><<<gf, 256, 0, st>>>(gray, h, w, (u8 *)ctx->lsd1.p, H, W, dth, dtw, (const int *)tb, xtd, ytd);
            else
                k_blur_resize_w<5><<<gf, 256, 0, st>>>(gray, h, w, (u8 *)ctx->lsd1.p, H, W, dth, dtw, (const int *)tb, xtd, ytd);
        } else if (wordwise && hh == 3)
            k_sepblur_q8_w<3><<<gw, 256, 0, st>>>(gray, (u8 *)ctx->lsd0.p, h, w, (const int *)tb);
        else if (wordwise)
            k_sepblur_q8_w<5><<<gw, 256, 0, st>>>(gray, (u8 *)ctx->lsd0.p, h, w, (const int *)tb);
        else
            k_sepblur_q8<<<dim3((w + SB_TW - 1) / SB_TW, (h + SB_TH - 1) / SB_TH, n), 256, 0, st>>>(
                gray, (u8 *)ctx->lsd0.p, h, w, (const int *)tb, (int)hh);
        LE_LAUNCH_CHECK(ctx);
        if (!dtw) {
            k_resize_exact<<<dim3((W + 31) / 32, (H + 7) / 8, n), 256, 0, st>>>((const u8 *)ctx->lsd0.p, h, w,
                                                                                (u8 *)ctx->lsd1.p, H, W, xtd, ytd);
            LE_LAUNCH_CHECK(ctx);
        }
        img = (const u8 *)ctx->lsd1.p;
    }
    LsdParams P;
    P.H = H; P.W = W; P.n_bins = n_bins; P.refine = refine; P.max_segs = max_segs;
    P.prec = M_PI * ang_th / 180;
    P.p = ang_th / 180;
    P.scale = scale;
    P.density_th = density_th;
    const double rho = quant / sin(P.prec);
    // largest integer s with sqrt(s/4.0) <= rho (so "defined" is s > s_thr)
    int s_thr = (int)floor(4.0 * rho * rho);
    while (s_thr >= 0 && !(sqrt((double)s_thr / 4.0) <= rho)) s_thr--;
    while (sqrt((double)(s_thr + 1) / 4.0) <= rho) s_thr++;
    const double LOG_NT = 5 * (log10((double)W) + log10((double)H)) / 2 + log10(11.0);
    P.min_reg_size = (unsigned)(size_t)(-LOG_NT / log10(P.p));
    P.log_nt = LOG_NT;
    P.log_eps = log_eps;
    int *cnt = (int *)ctx->counters.p;
    k_lsd_zero<<<(n + 255) / 256, 256, 0, st>>>(cnt, n);
    LE_LAUNCH_CHECK(ctx);
    if ((W & 3) == 0 && (((size_t)img) & 3) == 0)
        k_lsd_grad4<<<dim3((W / 4 + 31) / 32, (H + 7) / 8, n), 256, 0, st>>>(img, (u32 *)ctx->lsd2.p, cnt, H, W, s_thr);
    else
        k_lsd_grad<<<dim3((W + 31) / 32, (H + 7) / 8, n), 256, 0, st>>>(img, (u32 *)ctx->lsd2.p, cnt, H, W, s_thr);
    LE_LAUNCH_CHECK(ctx);
    u16 *chist = (u16 *)ctx->lsd5.p;
    u32 *coff = (u32 *)((char *)ctx->lsd5.p + sizeof(u16) * (size_t)nchunk * n_bins * n);
    // coff must be 4-byte aligned: nchunk*n_bins*n*2 is even*... force alignment
    coff = (u32 *)(((size_t)coff + 3) & ~(size_t)3);
    k_lsd_hist<<<dim3(nchunk, n), 256, sizeof(u32) * n_bins, st>>>(img, (const u32 *)ctx->lsd2.p, cnt, chist, H, W,
                                                                    nchunk, n_bins);
    LE_LAUNCH_CHECK(ctx);
    k_lsd_scan<<<n, 1024, sizeof(u32) * n_bins, st>>>(chist, coff, cnt, nchunk, n_bins);
    LE_LAUNCH_CHECK(ctx);
    k_lsd_scatter<<<dim3(nchunk, n), 32, sizeof(u32) * n_bins, st>>>(img, (const u32 *)ctx->lsd2.p, cnt, coff,
                                                                     (u32 *)ctx->lsd3.p, H, W, nchunk, n_bins);
    LE_LAUNCH_CHECK(ctx);
    k_lsd_cs<<<dim3(64, n), 256, 0, st>>>((const u32 *)ctx->lsd2.p, (const u32 *)ctx->lsd3.p, cnt, (float2 *)ctx->lsd6.p, npx);
    LE_LAUNCH_CHECK(ctx);
    LE_T1(ctx, LE_K_LSD_PREP, st);
    LE_T0(ctx, LE_K_LSD_GROW, st);
    if (spec) {
        if (ctx->lsd_epoch >= 126) {  // the epoch field of the tags is exhausted: clear the plane, start over
            LE_CUDA(ctx, cudaMemsetAsync(ctx->lsd7.p, 0, ctx->lsd7.bytes, st));
            ctx->lsd_epoch = 0;
        }
        const u32 epoch = (u32)++ctx->lsd_epoch;
        // warps per frame: 16 while every frame can have an SM of its own, else 8 (four CTAs per SM, 64 registers):
        // a second wave of CTAs costs more than the narrower rounds
        int K = n <= ctx->sm_count ? 16 : 8;
        if (const int e = LE_KNOB_INT("LE_LSD_K", 0)) K = e == 8 ? 8 : (e == 16 ? 16 : K);  // development knob
        const int sg_smem = K * (LSD_RING + SG_HASH) * (int)sizeof(u32);
        static le_dev_once sg_attr;
        if (sg_attr.needs(ctx->device, 1)) {
            const int big = 16 * (LSD_RING + SG_HASH) * (int)sizeof(u32);
            LE_CUDA(ctx, cudaFuncSetAttribute(k_lsd_grow_spec<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            LE_CUDA(ctx, cudaFuncSetAttribute(k_lsd_grow_spec<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            sg_attr.done(ctx->device, 1);
        }
#define LE_GROW(KK, AA)                                                                                          \
    k_lsd_grow_spec<KK, AA><<<n, KK * 32, sg_smem, st>>>(img, (u32 *)ctx->lsd2.p, (u32 *)ctx->lsd7.p,               \
                                                         (const float2 *)ctx->lsd6.p, (const u32 *)ctx->lsd3.p,      \
                                                         (u32 *)ctx->lsd4.p, cnt, P, segs, width, prec_out,          \
                                                         refine >= 2 ? nfa_out : nullptr, counts, epoch)
        if (K == 16 && refine >= 2) LE_GROW(16, true);
        else if (K == 16) LE_GROW(16, false);
        else if (refine >= 2) LE_GROW(8, true);
        else LE_GROW(8, false);
#undef LE_GROW
    } else
        k_lsd_grow<<<n, 32, 0, st>>>(img, (u32 *)ctx->lsd2.p, (const float2 *)ctx->lsd6.p, (const u32 *)ctx->lsd3.p,
                                     (u32 *)ctx->lsd4.p, cnt, P, segs, width, prec_out, refine >= 2 ? nfa_out : nullptr, counts);
    LE_T1(ctx, LE_K_LSD_GROW, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
