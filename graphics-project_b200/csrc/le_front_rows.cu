// le_front_rows.cu -- fused [BGR->gray] -> [blur5] -> Sobel -> NMS front kernel, register-resident.
//
// One WARP owns a column strip of 128 pixels (lane = 4 consecutive pixels = one packed word) and
// walks DOWN the rows of a segment, keeping every vertical window (5 gray rows, 3 blurred rows,
// 3 magnitude rows) in registers as transposed-FIR partial sums of packed 16-bit pairs.  Horizontal
// neighbours come from the adjacent lanes with warp shuffles, so there is no shared memory, no
// block barrier and every input byte is loaded from HBM exactly once per strip (the 8-pixel strip
// overlap and the 8 halo rows per segment hit L2).  The arithmetic is packed: IDP.2A for the Q15
// gray conversion, PRMT byte windows, 2x16-bit IMAD/IADD3 for the [1 4 6 4 1] and Sobel sums,
// VIMNMX.U16x2 for |gx|+|gy|.  The direction-dependent NMS test is scalar but only runs for warp
// rows that contain a pixel with magnitude > low threshold (edges are sparse: <3 % of warp rows).
//
// Lane l of a strip covers x in [xs + 4l, xs + 4l + 3], xs = 120 * strip - 4; lanes 1..30 produce
// output, lanes 0 and 31 only provide the +-4 pixel halo (blur 2 + Sobel 1 + NMS 1).
#include "le_common.cuh"

#define FR_STRIP 120
#define FR_WARPS 4

static __device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 s) { return __byte_perm(a, b, s); }

// 4 gray bytes from 12 BGR bytes: byte 2 of 2*(3735 B + 19235 G + 9798 R + 16384) is the Q15 result
static __device__ __forceinline__ u32 gray4(u32 a, u32 b, u32 c)
{
    const u32 cBG = 7470u | (38470u << 16), cR0 = 19596u, c0B = 7470u << 16, cGR = 38470u | (19596u << 16);
    u32 s0 = __dp2a_lo(cBG, a, 32768u);
    s0 = __dp2a_hi(cR0, a, s0);
    u32 s1 = __dp2a_hi(c0B, a, 32768u);
    s1 = __dp2a_lo(cGR, b, s1);
    u32 s2 = __dp2a_hi(cBG, b, 32768u);
    s2 = __dp2a_lo(cR0, c, s2);
    u32 s3 = __dp2a_lo(c0B, c, 32768u);
    s3 = __dp2a_hi(cGR, c, s3);
    return prmt(prmt(s0, s1, 0x0062), prmt(s2, s3, 0x0062), 0x5410);
}

template <bool BGR>
static __device__ __forceinline__ u32 load_px(const u8 *__restrict__ s, int y, int x, int w)
{
    if (BGR) {
        const u8 *p = s + ((size_t)y * w + x) * 3;
        return (3735u * p[0] + 19235u * p[1] + 9798u * p[2] + 16384u) >> 15;
    }
    return s[(size_t)y * w + x];
}

struct FrontArgs {
    const u8 *src;
    u8 *map;
    u32 *queue;
    int *counters;
    const u8 *lut;
    int h, w, lo, hi;
    int rows_per_seg, nstrips, nsegs;
};

// push helpers: strong pixels at the front of the frame queue, weak candidates at the back
static __device__ __forceinline__ void push_list(bool pred, u32 value, int *counter, u32 *q, long long stride_sign,
                                                 size_t base_index, int lane)
{
    u32 b = __ballot_sync(~0u, pred);
    if (b) {
        int base = 0;
        if (lane == (__ffs(b) - 1)) base = atomicAdd(counter, __popc(b));
        base = __shfl_sync(~0u, base, __ffs(b) - 1);
        if (pred) {
            long long pos = (long long)base + __popc(b & ((1u << lane) - 1));
            q[(long long)base_index + stride_sign * pos] = value;
        }
    }
}

template <bool BGR, bool BLUR, bool LUT>
__global__ void __launch_bounds__(FR_WARPS * 32) k_front_rows(FrontArgs A)
{
    __shared__ u8 s_lut[LUT ? 256 : 4];
    const int lane = threadIdx.x & 31;
    if (LUT) {
        for (int i = threadIdx.x; i < 256; i += FR_WARPS * 32) s_lut[i] = A.lut[blockIdx.y * 256 + i];
        __syncthreads();
    }
    // work item of this warp: (frame f = blockIdx.y, segment, strip)
    const int f = blockIdx.y;
    const int wi = blockIdx.x * FR_WARPS + (threadIdx.x >> 5);
    if (wi >= A.nstrips * A.nsegs) return;
    const int strip = wi % A.nstrips, seg = wi / A.nstrips;
    const int h = A.h, w = A.w, lo = A.lo, hi = A.hi;
    const size_t npx = (size_t)h * w;
    const u8 *__restrict__ src = A.src + (size_t)f * npx * (BGR ? 3 : 1);
    u8 *__restrict__ map = A.map + (size_t)f * npx;
    u32 *q = A.queue + (size_t)f * npx;
    int *cnt = A.counters + f * LE_C_STRIDE;

    const int ys = seg * A.rows_per_seg, ye = min(h, ys + A.rows_per_seg);
    const int x_lane = strip * FR_STRIP - 4 + 4 * lane;         // first pixel of this lane's word
    const bool inside = x_lane >= 0 && x_lane + 3 < w;          // whole word inside the image
    const bool fast_rows = BGR ? ((w * 3) & 3) == 0 : (w & 3) == 0;
    const bool out_lane = lane >= 1 && lane <= 30 && x_lane < w;
    constexpr int LAGB = BLUR ? 2 : 0;

    // replicate fix-up of the blurred word at the left/right image border (BLUR only)
    u32 fix_sel = 0x3210;
    int fix_src = lane;
    bool need_fix = false;
    if (BLUR) {
        if (x_lane + 3 < 0) { fix_sel = 0x4444; fix_src = lane + 1; }                    // all bytes := B(0)
        else if (x_lane < 0) { fix_sel = 0x3210; }                                        // (cannot happen: x_lane % 4 == 0)
        else if (x_lane >= w) { fix_sel = ((w & 3) == 0) ? 0x7777 : 0x3210; fix_src = lane - 1; }
        else if (x_lane + 3 >= w) {                                                       // word contains W-1 at byte k
            int k = w - 1 - x_lane;
            fix_sel = k == 0 ? 0x0000 : (k == 1 ? 0x1110 : 0x2210);
        }
        fix_src = min(max(fix_src, 0), 31);
        need_fix = __any_sync(~0u, fix_sel != 0x3210);
    }
    // magnitude is 0 outside the image (per 16-bit half)
    u32 mmask01 = ((x_lane >= 0 && x_lane < w) ? 0xffffu : 0u) | ((x_lane + 1 >= 0 && x_lane + 1 < w) ? 0xffff0000u : 0u);
    u32 mmask23 = ((x_lane + 2 >= 0 && x_lane + 2 < w) ? 0xffffu : 0u) | ((x_lane + 3 >= 0 && x_lane + 3 < w) ? 0xffff0000u : 0u);

    // ---- loader of one packed gray (or pre-Sobel) word for row r
    auto load_row = [&](int r, u32 &wa, u32 &wb, u32 &wc) {
        const int rr = BLUR ? d_reflect101(r, h) : d_clamp(r, 0, h - 1);
        if (inside && fast_rows) {
            if (BGR) {
                const u32 *p = (const u32 *)(src + ((size_t)rr * w + x_lane) * 3);
                wa = __ldg(p); wb = __ldg(p + 1); wc = __ldg(p + 2);
            } else {
                wa = __ldg((const u32 *)(src + (size_t)rr * w + x_lane));
            }
        } else {
            // border lanes / unaligned rows: per-pixel with the border rule, result already gray
            u32 g = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int x = x_lane + k;
                int xx = BLUR ? d_reflect101(x, w) : d_clamp(x, 0, w - 1);
                g |= load_px<BGR>(src, rr, xx, w) << (8 * k);
            }
            wa = g; wb = 0; wc = 0xffffffffu;  // marker: wa is gray already
        }
    };
    auto to_gray = [&](u32 wa, u32 wb, u32 wc) -> u32 {
        if (BGR) return (inside && fast_rows) ? gray4(wa, wb, wc) : wa;
        return wa;
    };

    // ---- pipeline state (packed 16-bit pairs: lo = pixels 0,1; hi = pixels 2,3)
    u32 a1l = 0, a2l = 0, a3l = 0, a4l = 0, a1h = 0, a2h = 0, a3h = 0, a4h = 0;  // vertical blur (transposed FIR)
    u32 s1l = 0, s2l = 0, ppl = 0, s1h = 0, s2h = 0, pph = 0;                    // vertical Sobel
    u32 gxl_p = 0, gxh_p = 0, gyl_p = 0, gyh_p = 0;                              // gradient of the NMS row
    u32 m_up_l = 0, m_up_h = 0, m_mid_l = 0, m_mid_h = 0;                        // magnitude rows yo-1, yo
    bool has_mid = false;

    const int r0 = ys - 2 - LAGB, r1 = ye + 1 + LAGB;  // inclusive input row range
    u32 na, nb, nc;
    load_row(r0, na, nb, nc);
    for (int r = r0; r <= r1; r++) {
        u32 ca = na, cb = nb, cc = nc;
        if (r < r1) load_row(r + 1, na, nb, nc);  // prefetch
        u32 g = to_gray(ca, cb, cc);
        if (LUT) {
            g = (u32)s_lut[g & 255] | ((u32)s_lut[(g >> 8) & 255] << 8) | ((u32)s_lut[(g >> 16) & 255] << 16) |
                ((u32)s_lut[g >> 24] << 24);
        }
        u32 B;  // pre-Sobel word of row yb
        const int yb = r - LAGB;
        if (BLUR) {
            // vertical [1 4 6 4 1] on pairs, transposed form: output row yb = r - 2
            u32 pl = prmt(g, 0, 0x4140), ph = prmt(g, 0, 0x4342);
            u32 vl = a4l + pl, vh = a4h + ph;
            a4l = a3l + 4 * pl; a4h = a3h + 4 * ph;
            a3l = a2l + 6 * pl; a3h = a2h + 6 * ph;
            a2l = a1l + 4 * pl; a2h = a1h + 4 * ph;
            a1l = pl; a1h = ph;
            // horizontal on the vertical sums (<= 4080 per half)
            u32 L23 = __shfl_up_sync(~0u, vh, 1), R01 = __shfl_down_sync(~0u, vl, 1);
            u32 Pm10 = prmt(L23, vl, 0x5432), P12 = prmt(vl, vh, 0x5432), P34 = prmt(vh, R01, 0x5432);
            u32 o01 = L23 + vh + 0x00800080u + 4 * (Pm10 + P12) + 6 * vl;
            u32 o23 = vl + R01 + 0x00800080u + 4 * (P12 + P34) + 6 * vh;
            B = prmt(o01, o23, 0x7531);
            if (need_fix) {
                u32 nbw = __shfl_sync(~0u, B, fix_src);
                B = prmt(B, nbw, fix_sel);
            }
        } else {
            B = g;
        }
        // vertical Sobel on pairs: centre row ysob = yb - 1
        u32 bl = prmt(B, 0, 0x4140), bh = prmt(B, 0, 0x4342);
        if (BLUR) {
            if (yb == 0) { s1l = bl; s1h = bh; }                 // B(-1) := B(0)
            if (yb >= h) { bl = s1l; bh = s1h; }                 // B(H) := B(H-1)
        }
        const int ysob = yb - 1;
        u32 vsl = s2l + bl, vsh = s2h + bh;                      // B(y-1) + 2 B(y) + B(y+1)
        u32 vdl = bl + 0x01000100u - ppl, vdh = bh + 0x01000100u - pph;  // B(y+1) - B(y-1) + 256
        s2l = s1l + 2 * bl; s2h = s1h + 2 * bh;
        ppl = s1l; pph = s1h;
        s1l = bl; s1h = bh;
        // horizontal Sobel: neighbours' (vs, vd) of pixels -1 and 4
        u32 LN = __shfl_up_sync(~0u, prmt(vsh, vdh, 0x7632), 1);
        u32 RN = __shfl_down_sync(~0u, prmt(vsl, vdl, 0x5410), 1);
        u32 Sm10 = prmt(LN, vsl, 0x5410), S12 = prmt(vsl, vsh, 0x5432), S34 = prmt(vsh, RN, 0x5432);
        u32 Dm10 = prmt(LN, vdl, 0x5432), D12 = prmt(vdl, vdh, 0x5432), D34 = prmt(vdh, RN, 0x7632);
        u32 gxl = S12 + 0x04000400u - Sm10, gxh = S34 + 0x04000400u - S12;   // gx + 1024
        u32 gyl = Dm10 + 2 * vdl + D12, gyh = D12 + 2 * vdh + D34;           // gy + 1024
        // |gx| + |gy|
        u32 ml = __vmaxu2(gxl, 0x08000800u - gxl) + __vmaxu2(gyl, 0x08000800u - gyl) - 0x08000800u;
        u32 mh = __vmaxu2(gxh, 0x08000800u - gxh) + __vmaxu2(gyh, 0x08000800u - gyh) - 0x08000800u;
        const bool row_in = ysob >= 0 && ysob < h;
        ml = row_in ? (ml & mmask01) : 0u;
        mh = row_in ? (mh & mmask23) : 0u;
        const u32 kcmp = (0x7fffu - (u32)lo) * 0x00010001u;
        const bool has_new = ((__vmaxu2(ml, mh) + kcmp) & 0x80008000u) != 0;

        // ---- NMS + threshold for row yo = ysob - 1 (magnitudes: up, mid, new)
        const int yo = ysob - 1;
        if (yo >= ys && yo < ye) {
            u32 outw = 0;
            if (__any_sync(~0u, has_mid)) {
                // neighbours' magnitudes of pixels -1 (from lane-1, its pixel 3) and 4 (from lane+1, its pixel 0)
                u32 l_um = __shfl_up_sync(~0u, prmt(m_up_h, m_mid_h, 0x7632), 1);     // lo16 = up(-1), hi16 = mid(-1)
                u32 l_dn = __shfl_up_sync(~0u, mh >> 16, 1);                           // new(-1)
                u32 r_um = __shfl_down_sync(~0u, prmt(m_up_l, m_mid_l, 0x5410), 1);   // lo16 = up(4), hi16 = mid(4)
                u32 r_dn = __shfl_down_sync(~0u, ml & 0xffffu, 1);                     // new(4)
                if (has_mid) {
                    int up[6], mid[6], dn[6];
                    up[0] = l_um & 0xffff; mid[0] = l_um >> 16; dn[0] = l_dn;
                    up[1] = m_up_l & 0xffff; up[2] = m_up_l >> 16; up[3] = m_up_h & 0xffff; up[4] = m_up_h >> 16;
                    mid[1] = m_mid_l & 0xffff; mid[2] = m_mid_l >> 16; mid[3] = m_mid_h & 0xffff; mid[4] = m_mid_h >> 16;
                    dn[1] = ml & 0xffff; dn[2] = ml >> 16; dn[3] = mh & 0xffff; dn[4] = mh >> 16;
                    up[5] = r_um & 0xffff; mid[5] = r_um >> 16; dn[5] = r_dn;
                    int gx[4] = {(int)(gxl_p & 0xffff) - 1024, (int)(gxl_p >> 16) - 1024, (int)(gxh_p & 0xffff) - 1024,
                                 (int)(gxh_p >> 16) - 1024};
                    int gy[4] = {(int)(gyl_p & 0xffff) - 1024, (int)(gyl_p >> 16) - 1024, (int)(gyh_p & 0xffff) - 1024,
                                 (int)(gyh_p >> 16) - 1024};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int mv = mid[k + 1];
                        if (mv > lo) {
                            int ax = abs(gx[k]), ay = abs(gy[k]) << 15;
                            int tg22x = ax * 13573;
                            bool keep;
                            if (ay < tg22x) keep = mv > mid[k] && mv >= mid[k + 2];
                            else {
                                int tg67x = tg22x + (ax << 16);
                                if (ay > tg67x) keep = mv > up[k + 1] && mv >= dn[k + 1];
                                else if ((gx[k] ^ gy[k]) < 0) keep = mv > up[k + 2] && mv > dn[k];
                                else keep = mv > up[k] && mv > dn[k + 2];
                            }
                            if (keep) outw |= (mv > hi ? 255u : 1u) << (8 * k);
                        }
                    }
                }
                if (!out_lane) outw = 0;
                // strong -> front of the queue, weak -> back; pixel index = yo * w + x
                if (__any_sync(~0u, outw != 0)) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        u32 c = (outw >> (8 * k)) & 255u;
                        u32 pidx = (u32)(yo * w + x_lane + k);
                        push_list(c == 255u, pidx, &cnt[LE_C_QTAIL], q, 1, 0, lane);
                        push_list(c == 1u, pidx, &cnt[LE_C_WEAK], q, -1, npx - 1, lane);
                    }
                }
            }
            if (out_lane) {
                u8 *o = map + (size_t)yo * w + x_lane;
                if ((w & 3) == 0) *(u32 *)o = outw;
                else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (x_lane + k < w) o[k] = (u8)(outw >> (8 * k));
                }
            }
        }
        // rotate the NMS windows
        m_up_l = m_mid_l; m_up_h = m_mid_h;
        m_mid_l = ml; m_mid_h = mh;
        gxl_p = gxl; gxh_p = gxh; gyl_p = gyl; gyh_p = gyh;
        has_mid = has_new;
    }
}

template <bool BGR, bool BLUR, bool LUT>
static int launch_rows_t(le_ctx *ctx, const u8 *src, u8 *edges, int n, int h, int w, int lo, int hi, cudaStream_t st)
{
    FrontArgs A;
    A.src = src; A.map = edges; A.queue = (u32 *)ctx->queue.p; A.counters = (int *)ctx->counters.p;
    A.lut = (const u8 *)ctx->lut.p;
    A.h = h; A.w = w; A.lo = min(lo, 4000); A.hi = hi;
    A.nstrips = (w + FR_STRIP - 1) / FR_STRIP;
    // enough warps to fill the machine even for one small frame, long segments for big batches
    long long target_warps = (long long)ctx->sm_count * 32;
    int rows = 72;
    while (rows > 8 && (long long)n * A.nstrips * ((h + rows - 1) / rows) < target_warps) rows /= 2;
    A.rows_per_seg = rows;
    A.nsegs = (h + rows - 1) / rows;
    int items = A.nstrips * A.nsegs;
    dim3 grid((items + FR_WARPS - 1) / FR_WARPS, n);
    LE_T0(ctx, LE_K_FRONT, st);
    k_front_rows<BGR, BLUR, LUT><<<grid, FR_WARPS * 32, 0, st>>>(A);
    LE_T1(ctx, LE_K_FRONT, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

int le_front_rows_launch(le_ctx *ctx, const u8 *src, int ch, int blur, int lut, u8 *edges, int n, int h, int w, int lo,
                         int hi, cudaStream_t st)
{
    if (lut) return launch_rows_t<false, false, true>(ctx, src, edges, n, h, w, lo, hi, st);
    if (ch == 3 && blur) return launch_rows_t<true, true, false>(ctx, src, edges, n, h, w, lo, hi, st);
    if (ch == 3) return launch_rows_t<true, false, false>(ctx, src, edges, n, h, w, lo, hi, st);
    if (blur) return launch_rows_t<false, true, false>(ctx, src, edges, n, h, w, lo, hi, st);
    return launch_rows_t<false, false, false>(ctx, src, edges, n, h, w, lo, hi, st);
}
