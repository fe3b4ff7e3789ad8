// le_front.cu -- gray / blur / normalize / Canny (a1-a4) for sm_100a.
//
// Data layout in HBM: src u8 [n][h][w][3] (BGR) or [n][h][w]; edge map u8 [n][h][w] in {0,255};
// per-frame work queue u32 [n][h*w] holding the flat indices of edge pixels (front) and of weak
// candidates (back).  One pass over the input produces the NMS classification, hysteresis then
// runs on the sparse queue only.
#include "le_common.cuh"
#include <stdlib.h>

int le_front_rows_launch(le_ctx *ctx, const u8 *src, int ch, int blur, int lut, u8 *edges, int n, int h, int w, int lo,
                         int hi, cudaStream_t st);

// ------------------------------------------------------------------ standalone stages
__global__ void k_bgr2gray(const u8 *__restrict__ bgr, u8 *__restrict__ gray, size_t npx)
{
    // 4 pixels (12 B in, 4 B out) per thread when aligned
    size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= npx) return;
    if (i4 + 4 <= npx) {
        const u32 *p = (const u32 *)(bgr + 3 * i4);  // 3*i4 is a multiple of 12 -> 4-aligned
        u32 a = p[0], b = p[1], c = p[2];
        u32 g0 = (3735u * (a & 255) + 19235u * ((a >> 8) & 255) + 9798u * ((a >> 16) & 255) + 16384u) >> 15;
        u32 g1 = (3735u * (a >> 24) + 19235u * (b & 255) + 9798u * ((b >> 8) & 255) + 16384u) >> 15;
        u32 g2 = (3735u * ((b >> 16) & 255) + 19235u * (b >> 24) + 9798u * (c & 255) + 16384u) >> 15;
        u32 g3 = (3735u * ((c >> 8) & 255) + 19235u * ((c >> 16) & 255) + 9798u * (c >> 24) + 16384u) >> 15;
        *(u32 *)(gray + i4) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
        for (size_t i = i4; i < npx; i++)
            gray[i] = (u8)((3735u * bgr[3 * i] + 19235u * bgr[3 * i + 1] + 9798u * bgr[3 * i + 2] + 16384u) >> 15);
    }
}

// generic 5x5 blur, one thread per pixel, smem tile 32x8 + halo 2
#define GB_TW 64
#define GB_TH 16
__global__ void __launch_bounds__(256) k_gauss5(const u8 *__restrict__ src, u8 *__restrict__ dst, int h, int w)
{
    __shared__ u8 g[GB_TH + 4][GB_TW + 4];
    __shared__ u16 hb[GB_TH + 4][GB_TW];
    const u8 *s = src + (size_t)blockIdx.z * h * w;
    u8 *d = dst + (size_t)blockIdx.z * h * w;
    int x0 = blockIdx.x * GB_TW, y0 = blockIdx.y * GB_TH;
    for (int i = threadIdx.x; i < (GB_TH + 4) * (GB_TW + 4); i += 256) {
        int ty = i / (GB_TW + 4), tx = i % (GB_TW + 4);
        g[ty][tx] = s[(size_t)d_reflect101(y0 + ty - 2, h) * w + d_reflect101(x0 + tx - 2, w)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (GB_TH + 4) * GB_TW; i += 256) {
        int ty = i / GB_TW, tx = i % GB_TW;
        hb[ty][tx] = g[ty][tx] + 4 * g[ty][tx + 1] + 6 * g[ty][tx + 2] + 4 * g[ty][tx + 3] + g[ty][tx + 4];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < GB_TH * GB_TW; i += 256) {
        int ty = i / GB_TW, tx = i % GB_TW;
        int x = x0 + tx, y = y0 + ty;
        if (x < w && y < h) {
            u32 v = hb[ty][tx] + 4 * hb[ty + 1][tx] + 6 * hb[ty + 2][tx] + 4 * hb[ty + 3][tx] + hb[ty + 4][tx];
            d[(size_t)y * w + x] = (u8)((v + 128) >> 8);
        }
    }
}

// per-frame min/max (atomics on the frame's counters), then LUT build, then apply
__global__ void k_minmax_init(int *counters, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        counters[i * LE_C_STRIDE + LE_C_MIN] = 255;
        counters[i * LE_C_STRIDE + LE_C_MAX] = 0;
    }
}
__global__ void k_minmax(const u8 *__restrict__ src, int *counters, size_t npx)
{
    const u8 *s = src + (size_t)blockIdx.y * npx;
    int mn = 255, mx = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x * 16;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < npx; i += stride) {
        if (i + 16 <= npx && ((size_t)(s + i) & 15) == 0) {
            uint4 v = *(const uint4 *)(s + i);
            u32 wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 a = wv[k];
                u32 hi = __vmaxu4(a, __byte_perm(a, 0, 0x1032));
                hi = __vmaxu4(hi, __byte_perm(hi, 0, 0x2301));
                u32 lo = __vminu4(a, __byte_perm(a, 0, 0x1032));
                lo = __vminu4(lo, __byte_perm(lo, 0, 0x2301));
                mx = max(mx, (int)(hi & 255));
                mn = min(mn, (int)(lo & 255));
            }
        } else {
            for (size_t j = i; j < npx && j < i + 16; j++) {
                mn = min(mn, (int)s[j]);
                mx = max(mx, (int)s[j]);
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(~0u, mn, o));
        mx = max(mx, __shfl_xor_sync(~0u, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&counters[blockIdx.y * LE_C_STRIDE + LE_C_MIN], mn);
        atomicMax(&counters[blockIdx.y * LE_C_STRIDE + LE_C_MAX], mx);
    }
}
// LUT exactly as cv2: scale/shift in f64, cast to f32, fused multiply-add, round-half-even, saturate
__global__ void k_build_lut(const int *counters, u8 *lut)
{
    int f = blockIdx.x, v = threadIdx.x;
    double smin = counters[f * LE_C_STRIDE + LE_C_MIN], smax = counters[f * LE_C_STRIDE + LE_C_MAX];
    double scale = 255.0 * ((smax - smin) > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
    double shift = 0.0 - smin * scale;
    float r = __fmaf_rn((float)v, (float)scale, (float)shift);
    int iv = __float2int_rn(r);
    lut[f * 256 + v] = (u8)min(max(iv, 0), 255);
}
__global__ void k_apply_lut(const u8 *__restrict__ src, u8 *__restrict__ dst, const u8 *__restrict__ lut, size_t npx)
{
    __shared__ u8 sl[256];
    sl[threadIdx.x] = lut[blockIdx.y * 256 + threadIdx.x];
    __syncthreads();
    const u8 *s = src + (size_t)blockIdx.y * npx;
    u8 *d = dst + (size_t)blockIdx.y * npx;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) d[i] = sl[s[i]];
}

// ------------------------------------------------------------------ fused front + NMS
#define FT_TW 128
#define FT_TH 32
#define FT_THREADS 256
#define FT_GW (FT_TW + 8)
#define FT_GH (FT_TH + 8)
#define FT_PW (FT_TW + 4)
#define FT_PH (FT_TH + 4)
#define FT_MW (FT_TW + 2)
#define FT_MH (FT_TH + 2)

struct FrontSmem {
    u8 G[FT_GH][FT_GW];       // gray with reflect halo 4 (blur variants only)
    u16 Hb[FT_GH][FT_PW];     // horizontal blur, Q4
    u8 P[FT_PH][FT_PW];       // pre-Sobel image at clamped coordinates, halo 2
    u16 mag[FT_MH][FT_MW];    // |dx|+|dy|, 0 outside the image, halo 1
    short2 dxy[FT_TH][FT_TW]; // gradient at the tile's own pixels
    u16 lstrong[FT_TW * FT_TH];
    u16 lweak[FT_TW * FT_TH];
    u8 lut[256];
    int nstrong, nweak, base_strong, base_weak;
};

static __device__ __forceinline__ u32 gray_of(u32 b, u32 g, u32 r)
{
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

template <bool BGR>
static __device__ __forceinline__ u8 load_gray(const u8 *__restrict__ s, int y, int x, int w)
{
    if (BGR) {
        const u8 *p = s + ((size_t)y * w + x) * 3;
        return (u8)gray_of(p[0], p[1], p[2]);
    }
    return s[(size_t)y * w + x];
}

// Map values written: 0 none, 1 weak candidate, 255 strong.  Queue: strong pixels appended at the
// front (counter QTAIL), weak candidates at the back (counter WEAK).
template <bool BGR, bool BLUR, bool LUT>
__global__ void __launch_bounds__(FT_THREADS)
k_front_nms(const u8 *__restrict__ src, u8 *__restrict__ map, u32 *__restrict__ queue, int *__restrict__ counters,
            const u8 *__restrict__ lut, int h, int w, int lo, int hi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontSmem &S = *reinterpret_cast<FrontSmem *>(smem_raw);
    const int f = blockIdx.z;
    const int tid = threadIdx.x;
    const size_t npx = (size_t)h * w;
    const u8 *s = src + (size_t)f * npx * (BGR ? 3 : 1);
    u8 *m = map + (size_t)f * npx;
    const int x0 = blockIdx.x * FT_TW, y0 = blockIdx.y * FT_TH;

    if (tid == 0) { S.nstrong = 0; S.nweak = 0; }
    if (LUT) S.lut[tid] = lut[f * 256 + tid];

    if (BLUR) {
        // S0: gray tile with REFLECT_101 halo of 4
        const bool fast = BGR && x0 >= 4 && x0 + FT_TW + 4 <= w && ((w * 3) & 3) == 0;
        if (fast) {
            // 4 pixels = 3 aligned words per thread-iteration
            for (int i = tid; i < FT_GH * (FT_GW / 4); i += FT_THREADS) {
                int ty = i / (FT_GW / 4), q = i % (FT_GW / 4);
                int sy = d_reflect101(y0 + ty - 4, h);
                const u32 *p = (const u32 *)(s + ((size_t)sy * w + (x0 - 4 + q * 4)) * 3);
                u32 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
                u32 g0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
                u32 g1 = gray_of(a >> 24, b & 255, (b >> 8) & 255);
                u32 g2 = gray_of((b >> 16) & 255, b >> 24, c & 255);
                u32 g3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
                *(u32 *)&S.G[ty][q * 4] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
            }
        } else {
            for (int i = tid; i < FT_GH * FT_GW; i += FT_THREADS) {
                int ty = i / FT_GW, tx = i % FT_GW;
                S.G[ty][tx] = load_gray<BGR>(s, d_reflect101(y0 + ty - 4, h), d_reflect101(x0 + tx - 4, w), w);
            }
        }
        __syncthreads();
        // S1: horizontal [1 4 6 4 1] at clamped x
        for (int i = tid; i < FT_GH * FT_PW; i += FT_THREADS) {
            int ty = i / FT_PW, tx = i % FT_PW;
            int c = d_clamp(x0 - 2 + tx, 0, w - 1) - (x0 - 4);
            const u8 *g = &S.G[ty][c];
            S.Hb[ty][tx] = g[-2] + 4 * g[-1] + 6 * g[0] + 4 * g[1] + g[2];
        }
        __syncthreads();
        // S2: vertical at clamped y, round, optional LUT
        for (int i = tid; i < FT_PH * FT_PW; i += FT_THREADS) {
            int ty = i / FT_PW, tx = i % FT_PW;
            int c = d_clamp(y0 - 2 + ty, 0, h - 1) - (y0 - 4);
            u32 v = S.Hb[c - 2][tx] + 4 * S.Hb[c - 1][tx] + 6 * S.Hb[c][tx] + 4 * S.Hb[c + 1][tx] + S.Hb[c + 2][tx];
            u8 o = (u8)((v + 128) >> 8);
            S.P[ty][tx] = LUT ? S.lut[o] : o;
        }
    } else {
        __syncthreads();
        for (int i = tid; i < FT_PH * FT_PW; i += FT_THREADS) {
            int ty = i / FT_PW, tx = i % FT_PW;
            u8 o = load_gray<BGR>(s, d_clamp(y0 - 2 + ty, 0, h - 1), d_clamp(x0 - 2 + tx, 0, w - 1), w);
            S.P[ty][tx] = LUT ? S.lut[o] : o;
        }
    }
    __syncthreads();
    // S3: Sobel magnitude with halo 1 (0 outside the image); gradient kept for own pixels
    for (int i = tid; i < FT_MH * FT_MW; i += FT_THREADS) {
        int ty = i / FT_MW, tx = i % FT_MW;
        int x = x0 - 1 + tx, y = y0 - 1 + ty;
        u16 mg = 0;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            const u8 *r0 = &S.P[ty][tx], *r1 = &S.P[ty + 1][tx], *r2 = &S.P[ty + 2][tx];
            int gx = (r0[2] + 2 * r1[2] + r2[2]) - (r0[0] + 2 * r1[0] + r2[0]);
            int gy = (r2[0] + 2 * r2[1] + r2[2]) - (r0[0] + 2 * r0[1] + r0[2]);
            mg = (u16)(abs(gx) + abs(gy));
            if (tx >= 1 && tx <= FT_TW && ty >= 1 && ty <= FT_TH) S.dxy[ty - 1][tx - 1] = make_short2((short)gx, (short)gy);
        }
        S.mag[ty][tx] = mg;
    }
    __syncthreads();
    // S4: non-maximum suppression + double threshold, 4 pixels per thread-iteration
    const bool vec_store = (w & 3) == 0;
    for (int i = tid; i < FT_TH * (FT_TW / 4); i += FT_THREADS) {
        int ty = i / (FT_TW / 4), q = i % (FT_TW / 4);
        int y = y0 + ty;
        if (y >= h) continue;
        u32 packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int tx = q * 4 + k, x = x0 + tx;
            if (x >= w) break;
            int mv = S.mag[ty + 1][tx + 1];
            u32 cls = 0;
            if (mv > lo) {
                short2 g = S.dxy[ty][tx];
                int ax = abs((int)g.x), ay = abs((int)g.y) << 15;
                int tg22x = ax * 13573;
                bool keep;
                if (ay < tg22x) keep = mv > S.mag[ty + 1][tx] && mv >= S.mag[ty + 1][tx + 2];
                else {
                    int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) keep = mv > S.mag[ty][tx + 1] && mv >= S.mag[ty + 2][tx + 1];
                    else {
                        int sg = ((g.x ^ g.y) < 0) ? -1 : 1;
                        keep = mv > S.mag[ty][tx + 1 - sg] && mv > S.mag[ty + 2][tx + 1 + sg];
                    }
                }
                if (keep) {
                    if (mv > hi) { cls = 255; S.lstrong[atomicAdd(&S.nstrong, 1)] = (u16)(ty * FT_TW + tx); }
                    else { cls = 1; S.lweak[atomicAdd(&S.nweak, 1)] = (u16)(ty * FT_TW + tx); }
                }
            }
            if (vec_store) packed |= cls << (8 * k);
            else m[(size_t)y * w + x] = (u8)cls;
        }
        if (vec_store && x0 + q * 4 < w) *(u32 *)(m + (size_t)y * w + x0 + q * 4) = packed;
    }
    __syncthreads();
    // flush the tile's lists to the frame queue
    int *cnt = counters + f * LE_C_STRIDE;
    if (tid == 0) {
        S.base_strong = S.nstrong ? atomicAdd(&cnt[LE_C_QTAIL], S.nstrong) : 0;
        S.base_weak = S.nweak ? atomicAdd(&cnt[LE_C_WEAK], S.nweak) : 0;
    }
    __syncthreads();
    u32 *q = queue + (size_t)f * npx;
    // strong + weak never exceed npx in total (each pixel is in at most one list)
    for (int i = tid; i < S.nstrong; i += FT_THREADS) {
        int t = S.lstrong[i];
        q[S.base_strong + i] = (u32)((y0 + t / FT_TW) * w + x0 + t % FT_TW);
    }
    for (int i = tid; i < S.nweak; i += FT_THREADS) {
        int t = S.lweak[i];
        q[npx - 1 - (S.base_weak + i)] = (u32)((y0 + t / FT_TW) * w + x0 + t % FT_TW);
    }
}

// Hysteresis on the sparse queue: one CTA per frame, level-synchronous BFS.  A weak pixel (1)
// becomes 255 exactly once (atomicOr on its word decides the winner, who appends it).
// Promoted pixels are appended after the strong seeds; they can overwrite weak-list slots only
// after those slots are no longer needed?  No: front and back never meet because every pixel is
// in the weak list OR a strong seed, and each weak pixel is appended to the front at most once,
// so front <= nstrong + nweak_promoted and the back holds nweak; a promoted weak pixel occupies
// two slots.  Capacity is therefore checked and LE_C_OVERFLOW raised if they would collide.
__global__ void __launch_bounds__(1024) k_hysteresis(u8 *__restrict__ map, u32 *__restrict__ queue,
                                                     int *__restrict__ counters, int h, int w)
{
    const int f = blockIdx.x;
    const size_t npx = (size_t)h * w;
    u8 *m = map + (size_t)f * npx;
    u32 *q = queue + (size_t)f * npx;
    int *cnt = counters + f * LE_C_STRIDE;
    __shared__ int s_tail, s_over;
    int head = 0, tail = cnt[LE_C_QTAIL];
    const int nweak = cnt[LE_C_WEAK];
    const long long cap = (long long)npx - nweak;  // front may use [0, cap)
    if (nweak == 0) return;  // nothing to promote, nothing to clear (block-uniform)
    if (nweak <= 4096 && tail + nweak <= cap) {
        // Few weak candidates (the usual case when hi is low): sweep the WEAK list instead of flooding
        // from every strong pixel.  A weak pixel with a 255 neighbour is promoted; repeat to a fixpoint.
        __shared__ int s_changed, s_t2;
        if (threadIdx.x == 0) s_t2 = tail;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < nweak; i += blockDim.x) {
                u32 p = q[npx - 1 - i];
                if (__ldcg(m + p) != 1) continue;
                int y = (int)(p / (u32)w), x = (int)(p - (u32)y * w);
                // all 8 neighbour reads are issued together (one L2 round trip, no early exit)
                int nb[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int dy = (k < 3) ? -1 : (k < 5 ? 0 : 1);
                    const int dx = (k < 3) ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6));
                    int yy = y + dy, xx = x + dx;
                    bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
                    nb[k] = in ? (int)__ldcg(m + (size_t)(in ? yy : y) * w + (in ? xx : x)) : 0;
                }
                bool hit = false;
#pragma unroll
                for (int k = 0; k < 8; k++) hit |= nb[k] == 255;
                if (hit) {
                    __stcg(m + p, (u8)255);  // single owner: this pixel appears once in the weak list
                    q[atomicAdd(&s_t2, 1)] = p;
                    s_changed = 1;
                }
            }
            __syncthreads();
            if (!s_changed) break;
        }
        for (int i = threadIdx.x; i < nweak; i += blockDim.x) {
            u32 p = q[npx - 1 - i];
            if (__ldcg(m + p) == 1) __stcg(m + p, (u8)0);
        }
        if (threadIdx.x == 0) cnt[LE_C_QTAIL] = s_t2;
        return;
    }
    if (threadIdx.x == 0) { s_tail = tail; s_over = 0; }
    __syncthreads();
    while (head < tail) {
        for (int i = head + threadIdx.x; i < tail; i += blockDim.x) {
            int p = (int)q[i];
            int y = p / w, x = p - y * w;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                int yy = y + dy;
                if (yy < 0 || yy >= h) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    int xx = x + dx;
                    if ((dx | dy) == 0 || xx < 0 || xx >= w) continue;
                    size_t o = (size_t)yy * w + xx;
                    if (__ldcg(m + o) == 1) {
                        size_t ga = (size_t)f * npx + o;  // byte offset in the whole batch (base is 256-aligned)
                        u32 *word = (u32 *)(map + (ga & ~(size_t)3));
                        u32 sh = (u32)(ga & 3) * 8;
                        u32 old = atomicOr(word, 0xFEu << sh);
                        if (((old >> sh) & 255) == 1) {
                            int pos = atomicAdd(&s_tail, 1);
                            if (pos < cap) q[pos] = (u32)o;
                            else s_over = 1;
                        }
                    }
                }
            }
        }
        __syncthreads();
        head = tail;
        tail = min(s_tail, (int)min(cap, (long long)0x7fffffff));
        __syncthreads();
    }
    // clear weak candidates that were never reached
    for (int i = threadIdx.x; i < nweak; i += blockDim.x) {
        u32 p = q[npx - 1 - i];
        if (__ldcg(m + p) == 1) m[p] = 0;  // L2 read: the promotions were L2 atomics
    }
    if (threadIdx.x == 0) {
        cnt[LE_C_QTAIL] = tail;
        if (s_over) cnt[LE_C_OVERFLOW] = 1;
    }
}

// Edge list from an arbitrary edge map (non-zero == edge), unordered.  Used when le_hough_lines
// is given a map that did not come from le_front_canny_u8.
__global__ void __launch_bounds__(256) k_edge_list(const u8 *__restrict__ map, u32 *__restrict__ queue,
                                                   int *__restrict__ counters, size_t npx)
{
    const int f = blockIdx.y;
    const u8 *m = map + (size_t)f * npx;
    u32 *q = queue + (size_t)f * npx;
    int *cnt = counters + f * LE_C_STRIDE;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    size_t nit = (npx + stride - 1) / stride;
    for (size_t it = 0; it < nit; it++) {
        size_t i = it * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool e = i < npx && m[i] != 0;
        u32 b = __ballot_sync(~0u, e);
        if (b) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&cnt[LE_C_QTAIL], __popc(b));
            base = __shfl_sync(~0u, base, 0);
            if (e) q[base + __popc(b & ((1u << lane) - 1))] = (u32)i;
        }
    }
}

__global__ void k_zero_counters(int *counters, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (n + 1) * LE_C_STRIDE) counters[i] = 0;  // group n: launch-wide counters (front work items)
}

// ------------------------------------------------------------------ host side
static int front_scratch(le_ctx *ctx, int n, int h, int w)
{
    int rc = le_ensure(ctx, &ctx->counters, sizeof(int) * LE_C_STRIDE * ((size_t)n + 1));
    if (rc) return rc;
    return le_ensure(ctx, &ctx->queue, sizeof(u32) * (size_t)n * h * w);
}

static int check_img(le_ctx *ctx, const void *a, const void *b, int n, int h, int w)
{
    if (!ctx) return LE_ERR_INVALID;
    LE_REQUIRE(ctx, a && b, "null image pointer");
    LE_REQUIRE(ctx, n > 0 && h > 0 && w > 0, "n, h, w must be positive");
    LE_REQUIRE(ctx, (size_t)h * w < ((size_t)1 << 30), "frame too large");
    LE_REQUIRE(ctx, (((size_t)a | (size_t)b) & 3) == 0, "image pointers must be 4-byte aligned");
    cudaSetDevice(ctx->device);
    return LE_OK;
}

extern "C" int le_bgr2gray_u8(le_ctx *ctx, const uint8_t *bgr, uint8_t *gray, int n, int h, int w, void *stream)
{
    int rc = check_img(ctx, bgr, gray, n, h, w);
    if (rc) return rc;
    size_t npx = (size_t)n * h * w;
    size_t threads = (npx + 3) / 4;
    k_bgr2gray<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bgr, gray, npx);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_gauss5_u8(le_ctx *ctx, const uint8_t *src, uint8_t *dst, int n, int h, int w, void *stream)
{
    int rc = check_img(ctx, src, dst, n, h, w);
    if (rc) return rc;
    LE_REQUIRE(ctx, src != dst, "le_gauss5_u8 is not in-place");
    dim3 grid((w + GB_TW - 1) / GB_TW, (h + GB_TH - 1) / GB_TH, n);
    k_gauss5<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, h, w);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

static int minmax_lut(le_ctx *ctx, const u8 *src, int n, int h, int w, cudaStream_t st)
{
    int rc = le_ensure(ctx, &ctx->counters, sizeof(int) * LE_C_STRIDE * ((size_t)n + 1));
    if (rc) return rc;
    rc = le_ensure(ctx, &ctx->lut, 256 * (size_t)n);
    if (rc) return rc;
    int *cnt = (int *)ctx->counters.p;
    k_minmax_init<<<(n + 255) / 256, 256, 0, st>>>(cnt, n);
    LE_LAUNCH_CHECK(ctx);
    size_t npx = (size_t)h * w;
    int bx = (int)min((size_t)64, (npx + 256 * 16 - 1) / (256 * 16));
    k_minmax<<<dim3(bx, n), 256, 0, st>>>(src, cnt, npx);
    LE_LAUNCH_CHECK(ctx);
    k_build_lut<<<n, 256, 0, st>>>(cnt, (u8 *)ctx->lut.p);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_minmax_normalize_u8(le_ctx *ctx, const uint8_t *src, uint8_t *dst, int n, int h, int w,
                                      void *stream)
{
    int rc = check_img(ctx, src, dst, n, h, w);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = minmax_lut(ctx, src, n, h, w, st);
    if (rc) return rc;
    size_t npx = (size_t)h * w;
    int bx = (int)min((size_t)256, (npx + 255) / 256);
    k_apply_lut<<<dim3(bx, n), 256, 0, st>>>(src, dst, (const u8 *)ctx->lut.p, npx);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

template <bool BGR, bool BLUR, bool LUT>
static int launch_front_t(le_ctx *ctx, const u8 *src, u8 *edges, int n, int h, int w, int lo, int hi, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_front_nms<BGR, BLUR, LUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(FrontSmem)));
        attr_set = true;
    }
    dim3 grid((w + FT_TW - 1) / FT_TW, (h + FT_TH - 1) / FT_TH, n);
    LE_T0(ctx, LE_K_FRONT, st);
    k_front_nms<BGR, BLUR, LUT><<<grid, FT_THREADS, sizeof(FrontSmem), st>>>(
        src, edges, (u32 *)ctx->queue.p, (int *)ctx->counters.p, (const u8 *)ctx->lut.p, h, w, lo, hi);
    LE_T1(ctx, LE_K_FRONT, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

int le_front_launch(le_ctx *ctx, const u8 *src, int ch, int do_blur, int do_norm, u8 *edges, int n, int h, int w,
                    int lo, int hi, cudaStream_t st)
{
    int rc = front_scratch(ctx, n, h, w);
    if (rc) return rc;
    ctx->edge_list_for = nullptr;
    const u8 *in = src;
    int in_ch = ch;
    int blur = do_blur;
    if (do_norm) {
        // min/max of the (blurred) gray image is a global reduction: materialise that image once.
        size_t npx = (size_t)n * h * w;
        rc = le_ensure(ctx, &ctx->tmp0, npx);
        if (rc) return rc;
        const u8 *g = src;
        if (ch == 3) {
            rc = le_ensure(ctx, &ctx->tmp1, npx);
            if (rc) return rc;
            rc = le_bgr2gray_u8(ctx, src, (u8 *)ctx->tmp1.p, n, h, w, st);
            if (rc) return rc;
            g = (const u8 *)ctx->tmp1.p;
        }
        if (do_blur) {
            rc = le_gauss5_u8(ctx, g, (u8 *)ctx->tmp0.p, n, h, w, st);
            if (rc) return rc;
            g = (const u8 *)ctx->tmp0.p;
        }
        rc = minmax_lut(ctx, g, n, h, w, st);
        if (rc) return rc;
        in = g;
        in_ch = 1;
        blur = 0;
    }
    k_zero_counters<<<((n + 1) * LE_C_STRIDE + 255) / 256, 256, 0, st>>>((int *)ctx->counters.p, n);
    LE_LAUNCH_CHECK(ctx);
    if (getenv("LE_FRONT_V1")) {  // previous shared-memory tile kernel, kept for A/B measurements only
        if (do_norm) rc = launch_front_t<false, false, true>(ctx, in, edges, n, h, w, lo, hi, st);
        else if (in_ch == 3 && blur) rc = launch_front_t<true, true, false>(ctx, in, edges, n, h, w, lo, hi, st);
        else if (in_ch == 3) rc = launch_front_t<true, false, false>(ctx, in, edges, n, h, w, lo, hi, st);
        else if (blur) rc = launch_front_t<false, true, false>(ctx, in, edges, n, h, w, lo, hi, st);
        else rc = launch_front_t<false, false, false>(ctx, in, edges, n, h, w, lo, hi, st);
    } else {
        rc = le_front_rows_launch(ctx, in, in_ch, blur, do_norm, edges, n, h, w, lo, hi, st);
    }
    if (rc) return rc;
    LE_T0(ctx, LE_K_HYST, st);
    k_hysteresis<<<n, 1024, 0, st>>>(edges, (u32 *)ctx->queue.p, (int *)ctx->counters.p, h, w);
    LE_T1(ctx, LE_K_HYST, st);
    LE_LAUNCH_CHECK(ctx);
    ctx->edge_list_for = edges;
    ctx->edge_list_n = n;
    ctx->edge_list_h = h;
    ctx->edge_list_w = w;
    return LE_OK;
}

static int canny_thresholds(le_ctx *ctx, double lo, double hi, int *ilo, int *ihi)
{
    LE_REQUIRE(ctx, lo == lo && hi == hi, "NaN threshold");
    if (lo > hi) { double t = lo; lo = hi; hi = t; }
    *ilo = (int)floor(lo);
    *ihi = (int)floor(hi);
    return LE_OK;
}

extern "C" int le_canny_u8(le_ctx *ctx, const uint8_t *src, uint8_t *edges, int n, int h, int w, double lo,
                           double hi, void *stream)
{
    return le_front_canny_u8(ctx, src, 1, 0, 0, edges, n, h, w, lo, hi, stream);
}

extern "C" int le_front_canny_u8(le_ctx *ctx, const uint8_t *src, int src_channels, int do_blur, int do_normalize,
                                 uint8_t *edges, int n, int h, int w, double lo, double hi, void *stream)
{
    int rc = check_img(ctx, src, edges, n, h, w);
    if (rc) return rc;
    LE_REQUIRE(ctx, src_channels == 1 || src_channels == 3, "src_channels must be 1 or 3");
    LE_REQUIRE(ctx, src != edges, "in-place Canny is not supported");
    int ilo, ihi;
    rc = canny_thresholds(ctx, lo, hi, &ilo, &ihi);
    if (rc) return rc;
    return le_front_launch(ctx, src, src_channels, do_blur, do_normalize, edges, n, h, w, ilo, ihi,
                           (cudaStream_t)stream);
}

int le_edge_list_from_map(le_ctx *ctx, const u8 *edges, int n, int h, int w, cudaStream_t st)
{
    if (!edges) {
        // explicit reuse of the lists left by the last le_front_canny_u8 on this context
        if (ctx->edge_list_for && ctx->edge_list_n == n && ctx->edge_list_h == h && ctx->edge_list_w == w) return LE_OK;
        return le_fail(ctx, LE_ERR_INVALID, "edges == NULL but the context holds no edge lists for n=%d h=%d w=%d", n, h, w);
    }
    int rc = front_scratch(ctx, n, h, w);
    if (rc) return rc;
    k_zero_counters<<<((n + 1) * LE_C_STRIDE + 255) / 256, 256, 0, st>>>((int *)ctx->counters.p, n);
    LE_LAUNCH_CHECK(ctx);
    size_t npx = (size_t)h * w;
    int bx = (int)min((size_t)128, (npx + 255) / 256);
    LE_T0(ctx, LE_K_EDGELIST, st);
    k_edge_list<<<dim3(bx, n), 256, 0, st>>>(edges, (u32 *)ctx->queue.p, (int *)ctx->counters.p, npx);
    LE_T1(ctx, LE_K_EDGELIST, st);
    LE_LAUNCH_CHECK(ctx);
    ctx->edge_list_for = nullptr;  // a foreign map may change between calls: never cache it
    ctx->edge_list_n = n;
    ctx->edge_list_h = h;
    ctx->edge_list_w = w;
    return LE_OK;
}
