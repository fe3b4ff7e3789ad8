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

// Hysteresis on the sparse queue: one CTA per frame, level-synchronous BFS.  A weak pixel (1)
// becomes 255 exactly once (atomicOr on its word decides the winner, who appends it).
// Promoted pixels are appended after the strong seeds; they can overwrite weak-list slots only
// after those slots are no longer needed?  No: front and back never meet because every pixel is
// in the weak list OR a strong seed, and each weak pixel is appended to the front at most once,
// so front <= nstrong + nweak_promoted and the back holds nweak; a promoted weak pixel occupies
// two slots.  Capacity is therefore checked and LE_C_OVERFLOW raised if they would collide.
__global__ void __launch_bounds__(1024, 2) k_hysteresis(u8 *__restrict__ map, u32 *__restrict__ queue,
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
        // Few weak candidates (the usual case when hi is low): work on the WEAK list instead of flooding from every
        // strong pixel, and keep the propagation out of global memory.  ONE round of loads fetches the eight
        // neighbour bytes of every weak pixel (is a strong pixel adjacent?  which neighbours are weak?).  The weak
        // pixels sit in a shared-memory hash set; their 8-connected components are formed by a union-find over the
        // slots (atomicMin hooking, path halving), a component is promoted iff one of its pixels touches a strong
        // one.  A constant number of barriers instead of one L2 round trip (or one shared sweep) per hop of the
        // weak chains along a soft edge, which are hundreds of pixels long.
        constexpr int PER = 4;          // nweak <= 4096 and the CTA has 1024 threads
        constexpr int HS = 8192;        // hash slots (load factor <= 0.5)
        constexpr u32 EMPTY = 0x7fffffffu, DONE = 0x80000000u;
        extern __shared__ u32 s_dyn[];
        u32 *s_slot = s_dyn, *s_par = s_dyn + HS;  // pixel index (bit 31: component promoted) / union-find parent
        __shared__ int s_t2;
        for (int i = threadIdx.x; i < HS; i += 1024) s_slot[i] = EMPTY;
        if (threadIdx.x == 0) s_t2 = tail;
        u32 pw[PER];
        u32 wm = 0;   // per pixel one byte: which of the 8 neighbours are weak
        u32 st = 0;   // per pixel: bit k = in the list, bit 8 + k = touches a strong pixel
        int own[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = threadIdx.x + k * 1024;
            pw[k] = i < nweak ? q[npx - 1 - i] : 0u;
            if (i < nweak) st |= 1u << k;
            own[k] = 0;
        }
        __syncthreads();
        auto hsh = [](u32 p) { return (p * 2654435761u) >> 19; };  // 13 bits
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (!(st >> k & 1)) continue;
            u32 hx = hsh(pw[k]);
            while (atomicCAS(&s_slot[hx], EMPTY, pw[k]) != EMPTY) hx = (hx + 1) & (HS - 1);
            own[k] = (int)hx;
            s_par[hx] = hx;
        }
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (!(st >> k & 1)) continue;
            const u32 p = pw[k];
            const int y = (int)(p / (u32)w), x = (int)(p - (u32)y * w);
            int nb[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int dy = (j < 3) ? -1 : (j < 5 ? 0 : 1);
                const int dx = (j < 3) ? j - 1 : (j == 3 ? -1 : (j == 4 ? 1 : j - 6));
                const int yy = y + dy, xx = x + dx;
                const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
                nb[j] = in ? (int)__ldcg(m + (size_t)(in ? yy : y) * w + (in ? xx : x)) : 0;
            }
            bool hit = false;
            u32 mask = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                hit |= nb[j] == 255;
                mask |= (nb[j] == 1 ? 1u : 0u) << j;
            }
            wm |= mask << (8 * k);
            if (hit) st |= 1u << (8 + k);
        }
        __syncthreads();  // every weak pixel is in the set, every parent initialised
        volatile u32 *par = s_par;
        auto find = [&](u32 a) {
            for (;;) {
                const u32 pa = par[a];
                if (pa == a) return a;
                const u32 ga = par[pa];
                par[a] = ga;  // path halving (benign race: parents only ever move towards the root)
                a = ga;
            }
        };
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (!(st >> k & 1)) continue;
            // the four "forward" neighbours are enough: every adjacency is seen from one of its two ends
            const u32 mask = (wm >> (8 * k)) & 0xf0u;
            if (!mask) continue;
            const u32 p = pw[k];
            const int y = (int)(p / (u32)w), x = (int)(p - (u32)y * w);
            for (u32 mm = mask; mm; mm &= mm - 1) {
                const int j = __ffs(mm) - 1;
                const int dy = (j < 5) ? 0 : 1;
                const int dx = (j == 4) ? 1 : j - 6;
                const u32 np_ = (u32)((y + dy) * w + (x + dx));
                u32 hx = hsh(np_);
                while ((((volatile u32 *)s_slot)[hx] & ~DONE) != np_) hx = (hx + 1) & (HS - 1);  // a weak neighbour is in the set
                u32 a = (u32)own[k], b = hx;
                for (;;) {  // union: hook the larger root under the smaller
                    a = find(a);
                    b = find(b);
                    if (a == b) break;
                    if (a < b) { const u32 t = a; a = b; b = t; }
                    const u32 old = atomicMin(&s_par[a], b);
                    if (old == a) break;
                    a = old;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; k++)
            if ((st >> k & 1) && (st >> (8 + k) & 1)) atomicOr(&s_slot[find((u32)own[k])], DONE);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (!(st >> k & 1)) continue;
            if (((volatile u32 *)s_slot)[find((u32)own[k])] & DONE) {
                __stcg(m + pw[k], (u8)255);
                q[atomicAdd(&s_t2, 1)] = pw[k];
            } else
                __stcg(m + pw[k], (u8)0);  // weak candidates that were never reached
        }
        __syncthreads();
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
    return LE_OK;
}

extern "C" int le_bgr2gray_u8(le_ctx *ctx, const uint8_t *bgr, uint8_t *gray, int n, int h, int w, void *stream)
{
    LE_ON_DEVICE(ctx);
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
    LE_ON_DEVICE(ctx);
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
    LE_ON_DEVICE(ctx);
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

static int launch_hysteresis(le_ctx *ctx, u8 *edges, int n, int h, int w, cudaStream_t st);

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
    rc = le_front_rows_launch(ctx, in, in_ch, blur, do_norm, edges, n, h, w, lo, hi, st);
    if (rc) return rc;
    LE_T0(ctx, LE_K_HYST, st);
    static le_dev_once hyst_attr;
    if (hyst_attr.needs(ctx->device, 65536)) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_hysteresis, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        hyst_attr.done(ctx->device, 65536);
    }
    k_hysteresis<<<n, 1024, 65536, st>>>(edges, (u32 *)ctx->queue.p, (int *)ctx->counters.p, h, w);
    LE_T1(ctx, LE_K_HYST, st);
    LE_LAUNCH_CHECK(ctx);
    ctx->edge_list_for = edges;
    ctx->edge_list_n = n;
    ctx->edge_list_h = h;
    ctx->edge_list_w = w;
    return LE_OK;
}

// ---- cv2.Canny on a 3-channel image (SURVEY Appendix B convention; no call site of the reference uses it, so
// this is a plain one-thread-per-pixel kernel, not a tuned one).  Sobel per channel, the channel with the largest
// |dx| + |dy| wins (the first on ties); NMS on that (dx, dy, magnitude); strong / weak lists as the fused front
// kernel leaves them, so the usual hysteresis kernel finishes the job.
static __device__ __forceinline__ int sobel_c3(const u8 *__restrict__ im, int h, int w, int x, int y, int *gx_out,
                                               int *gy_out)
{
    if (x < 0 || x >= w || y < 0 || y >= h) return 0;  // cv2 pads the magnitude plane with zeros
    const int xm = max(x - 1, 0), xp = min(x + 1, w - 1), ym = max(y - 1, 0), yp = min(y + 1, h - 1);
    int best = -1;
#pragma unroll
    for (int c = 0; c < 3; c++) {
#define P3(yy, xx) ((int)im[((size_t)(yy) * w + (xx)) * 3 + c])
        const int gx = (P3(ym, xp) + 2 * P3(y, xp) + P3(yp, xp)) - (P3(ym, xm) + 2 * P3(y, xm) + P3(yp, xm));
        const int gy = (P3(yp, xm) + 2 * P3(yp, x) + P3(yp, xp)) - (P3(ym, xm) + 2 * P3(ym, x) + P3(ym, xp));
#undef P3
        const int m = abs(gx) + abs(gy);
        if (m > best) {
            best = m;
            *gx_out = gx;
            *gy_out = gy;
        }
    }
    return best;
}

__global__ void __launch_bounds__(256) k_canny_c3(const u8 *__restrict__ src, u8 *__restrict__ map,
                                                  u32 *__restrict__ queue, int *__restrict__ counters, int h, int w,
                                                  int lo, int hi)
{
    const int f = blockIdx.z;
    const size_t npx = (size_t)h * w;
    const u8 *im = src + (size_t)f * npx * 3;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    int gx = 0, gy = 0, t0, t1;
    const int mv = sobel_c3(im, h, w, x, y, &gx, &gy);
    u8 out = 0;
    if (mv > lo) {
        const int ax = abs(gx), ay = abs(gy) << 15, tg22x = ax * 13573;
        bool keep;
        if (ay < tg22x) keep = mv > sobel_c3(im, h, w, x - 1, y, &t0, &t1) && mv >= sobel_c3(im, h, w, x + 1, y, &t0, &t1);
        else if (ay > tg22x + (ax << 16))
            keep = mv > sobel_c3(im, h, w, x, y - 1, &t0, &t1) && mv >= sobel_c3(im, h, w, x, y + 1, &t0, &t1);
        else {
            const int s = ((gx ^ gy) < 0) ? -1 : 1;
            keep = mv > sobel_c3(im, h, w, x - s, y - 1, &t0, &t1) && mv > sobel_c3(im, h, w, x + s, y + 1, &t0, &t1);
        }
        if (keep) {
            int *cnt = counters + f * LE_C_STRIDE;
            u32 *q = queue + (size_t)f * npx;
            const u32 idx = (u32)((size_t)y * w + x);
            if (mv > hi) {
                out = 255;
                q[atomicAdd(&cnt[LE_C_QTAIL], 1)] = idx;
            } else {
                out = 1;
                q[npx - 1 - (size_t)atomicAdd(&cnt[LE_C_WEAK], 1)] = idx;
            }
        }
    }
    map[(size_t)f * npx + (size_t)y * w + x] = out;
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
    LE_ON_DEVICE(ctx);
    return le_front_canny_u8(ctx, src, 1, 0, 0, edges, n, h, w, lo, hi, stream);
}

static int launch_hysteresis(le_ctx *ctx, u8 *edges, int n, int h, int w, cudaStream_t st)
{
    LE_T0(ctx, LE_K_HYST, st);
    static le_dev_once hyst_attr;
    if (hyst_attr.needs(ctx->device, 65536)) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_hysteresis, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        hyst_attr.done(ctx->device, 65536);
    }
    k_hysteresis<<<n, 1024, 65536, st>>>(edges, (u32 *)ctx->queue.p, (int *)ctx->counters.p, h, w);
    LE_T1(ctx, LE_K_HYST, st);
    LE_LAUNCH_CHECK(ctx);
    ctx->edge_list_for = edges;
    ctx->edge_list_n = n;
    ctx->edge_list_h = h;
    ctx->edge_list_w = w;
    return LE_OK;
}

extern "C" int le_canny_u8c3(le_ctx *ctx, const uint8_t *src, uint8_t *edges, int n, int h, int w, double lo,
                             double hi, void *stream)
{
    LE_ON_DEVICE(ctx);
    int rc = check_img(ctx, src, edges, n, h, w);
    if (rc) return rc;
    LE_REQUIRE(ctx, src != edges, "in-place Canny is not supported");
    int ilo, ihi;
    if ((rc = canny_thresholds(ctx, lo, hi, &ilo, &ihi))) return rc;
    if ((rc = front_scratch(ctx, n, h, w))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ctx->edge_list_for = nullptr;
    k_zero_counters<<<((n + 1) * LE_C_STRIDE + 255) / 256, 256, 0, st>>>((int *)ctx->counters.p, n);
    LE_LAUNCH_CHECK(ctx);
    LE_T0(ctx, LE_K_FRONT, st);
    k_canny_c3<<<dim3((w + 31) / 32, (h + 7) / 8, n), 256, 0, st>>>(src, edges, (u32 *)ctx->queue.p,
                                                                    (int *)ctx->counters.p, h, w, ilo, ihi);
    LE_T1(ctx, LE_K_FRONT, st);
    LE_LAUNCH_CHECK(ctx);
    return launch_hysteresis(ctx, edges, n, h, w, st);
}

extern "C" int le_front_canny_u8(le_ctx *ctx, const uint8_t *src, int src_channels, int do_blur, int do_normalize,
                                 uint8_t *edges, int n, int h, int w, double lo, double hi, void *stream)
{
    LE_ON_DEVICE(ctx);
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
