// le_front_rows.cu -- fused [BGR->gray] -> [blur5] -> Sobel -> NMS front kernel, register-resident.
//
// One WARP owns a column strip of 128 pixels (lane = 4 consecutive pixels = one packed word) and
// walks DOWN the rows of a segment, keeping every vertical window (5 gray rows, 3 blurred rows,
// 3 magnitude rows) in registers as transposed-FIR partial sums of packed 16-bit pairs.  Horizontal
// neighbours come from the adjacent lanes with warp shuffles: no block barrier, and every input
// byte is fetched from HBM once per strip (the 8-pixel strip overlap and the 8 halo rows per
// segment hit L2).  Input rows are streamed through a per-warp shared-memory ring with cp.async
// (LDGSTS), 7 rows ahead, so the bytes in flight per SM do not depend on registers or occupancy.
// The arithmetic is packed: IDP.2A for the Q15 gray conversion, PRMT byte windows, 2x16-bit
// IMAD/IADD3 for the [1 4 6 4 1] and Sobel sums, VIMNMX.U16x2 for |gx|+|gy|.  The direction-
// dependent NMS test is scalar and out of line; it only runs for warp rows that contain a pixel
// with magnitude > low threshold (edges are sparse).
//
// Lane l of a strip covers x in [xs + 4l, xs + 4l + 3], xs = 120 * strip - 4; lanes 1..30 produce
// output, lanes 0 and 31 only provide the +-4 pixel halo (blur 2 + Sobel 1 + NMS 1).
#include "le_common.cuh"
#include <cuda.h>

#define FR_STRIP 120
#define FR_WARPS 1
#ifndef FR_DEPTH
#define FR_DEPTH 8
#endif
// FR_DEPTH: // rows in the cp.async ring (prefetch distance FR_DEPTH - 1)
#define FR_FLAT 9  // identical constant rows after which the pipeline state is at its fixed point
#ifndef FR_NS
#define FR_NS 4
#endif
// FR_NS: stages of 4 rows in the bulk-copy (TMA) ring
#define FR_LEVELS 4
#define FR_STAGE 224  // entries of a warp's strong / weak staging list (a row adds at most 120)
#ifndef FR_MINBLK
#define FR_MINBLK 24
#endif
#ifndef FR_UNROLL
#define FR_UNROLL 2
#endif

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

static __device__ __forceinline__ void cp_async4(u32 smem_addr, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}
static __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
static __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- bulk async copies (TMA engine, SASS UBLKCP) completing on an mbarrier
static __device__ __forceinline__ void mbar_init(u32 bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void tma_load_2d(u32 dst, const CUtensorMap *tm, int x, int y, u32 bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
static __device__ __forceinline__ void mbar_wait(u32 bar, u32 parity)
{
    u32 done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}

struct FrontArgs {
    const u8 *src;
    u8 *map;
    u32 *queue;
    int *counters;
    const u8 *lut;
    int h, w, lo, hi;
    int rows_per_seg, nstrips, nsegs;
    int *work;    // work-item counter of this launch (zeroed before it)
    int nframes;
    // guided schedule: the batch is cut into FR_LEVELS groups of frames with shorter and shorter row segments,
    // so that the work items handed out last are small (level l: frames [lvl_frame0[l], lvl_frame0[l+1]))
    int lvl_frame0[FR_LEVELS + 1], lvl_rows[FR_LEVELS], lvl_nsegs[FR_LEVELS], lvl_item0[FR_LEVELS + 1];
};

// Pipeline state of one lane (packed 16-bit pairs: l = pixels 0,1; h = pixels 2,3)
struct FrontState {
    u32 a1l, a2l, a3l, a4l, a1h, a2h, a3h, a4h;  // vertical blur, transposed FIR partial sums
    u32 s1l, s2l, ppl, s1h, s2h, pph;            // vertical Sobel
    u32 gxl, gxh, gyl, gyh;                      // biased gradient (+1024) of the NMS row
    u32 mul, muh, mml, mmh;                      // magnitude rows yo-1 (up) and yo (mid)
    bool has_mid;
};

// Out-of-line NMS + double threshold for the 4 pixels of a lane.  Called warp-convergently.
// Returns the map word (0 / 1 weak / 255 strong per byte).
static __device__ __forceinline__ u32 nms_slow(u32 mul, u32 muh, u32 mml, u32 mmh, u32 mdl, u32 mdh, u32 gxl, u32 gxh,
                                            u32 gyl, u32 gyh, int has_mid, int lo, int hi)
{
    // neighbours' magnitudes of pixel -1 (lane-1, its pixel 3) and pixel 4 (lane+1, its pixel 0)
    u32 l_um = __shfl_up_sync(~0u, prmt(muh, mmh, 0x7632), 1);    // lo16 = up(-1), hi16 = mid(-1)
    u32 l_dn = __shfl_up_sync(~0u, mdh >> 16, 1);                  // down(-1)
    u32 r_um = __shfl_down_sync(~0u, prmt(mul, mml, 0x5410), 1);  // lo16 = up(4), hi16 = mid(4)
    u32 r_dn = __shfl_down_sync(~0u, mdl & 0xffffu, 1);            // down(4)
    u32 outw = 0;
    if (has_mid) {
        int up[6], mid[6], dn[6];
        up[0] = l_um & 0xffff; mid[0] = l_um >> 16; dn[0] = l_dn;
        up[1] = mul & 0xffff; up[2] = mul >> 16; up[3] = muh & 0xffff; up[4] = muh >> 16;
        mid[1] = mml & 0xffff; mid[2] = mml >> 16; mid[3] = mmh & 0xffff; mid[4] = mmh >> 16;
        dn[1] = mdl & 0xffff; dn[2] = mdl >> 16; dn[3] = mdh & 0xffff; dn[4] = mdh >> 16;
        up[5] = r_um & 0xffff; mid[5] = r_um >> 16; dn[5] = r_dn;
        int gx[4] = {(int)(gxl & 0xffff) - 1024, (int)(gxl >> 16) - 1024, (int)(gxh & 0xffff) - 1024,
                     (int)(gxh >> 16) - 1024};
        int gy[4] = {(int)(gyl & 0xffff) - 1024, (int)(gyl >> 16) - 1024, (int)(gyh & 0xffff) - 1024,
                     (int)(gyh >> 16) - 1024};
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
    return outw;
}

// Strong pixels go to the front of the frame queue, weak candidates to the back.  A warp first collects
// them in its shared staging lists (ballot-compacted, no global traffic) and moves a list to the queue
// with ONE global atomic once it might not hold another row: the atomic's round trip to L2 was the
// largest stall of the kernel when it was paid per pixel column.
struct Stage {
    u32 *s;       // shared staging lists of this warp: strong at s, weak at s + FR_STAGE
    int ns, nw;   // entries held (warp-uniform)
};

static __device__ __forceinline__ void stage_flush(u32 *buf, int &n, int *counter, u32 *q, long long npx, bool back)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, n);
    base = __shfl_sync(~0u, base, 0);
    for (int i = lane; i < n; i += 32) {
        long long pos = (long long)base + i;
        q[back ? npx - 1 - pos : pos] = buf[i];
    }
    __syncwarp();
    n = 0;
}

static __device__ __forceinline__ void stage_push(u32 outw, u32 pidx0, Stage &T)
{
    u32 lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    const u32 sflag = outw & 0x80808080u;                          // byte == 255
    const u32 wflag = outw & ~(outw >> 7) & 0x01010101u;           // byte == 1
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool pred = (sflag >> (8 * k)) & 0x80u;
        const u32 b = __ballot_sync(~0u, pred);
        if (pred) T.s[T.ns + __popc(b & lt)] = pidx0 + k;
        T.ns += __popc(b);
    }
    if (__any_sync(~0u, wflag != 0)) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool pred = (wflag >> (8 * k)) & 1u;
            const u32 b = __ballot_sync(~0u, pred);
            if (pred) T.s[FR_STAGE + T.nw + __popc(b & lt)] = pidx0 + k;
            T.nw += __popc(b);
        }
    }
}

template <bool BGR, bool BLUR, bool LUT, bool TMA>
__global__ void __launch_bounds__(FR_WARPS * 32, FR_MINBLK) k_front_rows(FrontArgs A, const __grid_constant__ CUtensorMap tmap)
{
    // ring row: 128 pixels; the TMA variant copies from the 16-byte boundary below the strip start
    constexpr int SLOTB = TMA ? (BGR ? 400 : 144) : (BGR ? 384 : 128);
    constexpr int STAGEB = (4 * SLOTB + 127) / 128 * 128;  // TMA stage: four rows, 128-byte aligned
    constexpr int RINGW = (TMA ? FR_NS * STAGEB : FR_DEPTH * SLOTB) / 4;
    __shared__ u8 s_lut[LUT ? 256 : 4];
    __shared__ __align__(128) u32 s_ring[FR_WARPS][RINGW];
    __shared__ __align__(8) u64 s_bar[FR_WARPS][TMA ? FR_NS : 1];
    __shared__ u32 s_stage[FR_WARPS][2][FR_STAGE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int h = A.h, w = A.w, lo = A.lo, hi = A.hi;
    const long long npx = (long long)h * w;
    const u32 bar_s = (u32)__cvta_generic_to_shared(&s_bar[wib][0]);
    if (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < FR_NS; i++) mbar_init(bar_s + 8 * i, 1);
            mbar_init_fence();
        }
        __syncwarp();
    }
    u32 gbase = 0;  // TMA ring: four-row groups this warp has requested (and consumed) so far
    int lut_frame = -1;
    // Persistent warps: work items (frame, segment, strip) are handed out by an atomic counter, so a warp
    // that drew flat strips simply takes more of them and no SM waits for block launches.
#pragma unroll 1
    for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(A.work, 1);
    item = __shfl_sync(~0u, item, 0);
    if (item >= A.lvl_item0[FR_LEVELS]) break;
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < FR_LEVELS; l++) lvl += item >= A.lvl_item0[l];
    const int rows_per_seg = A.lvl_rows[lvl];
    const int items_per_frame = A.nstrips * A.lvl_nsegs[lvl];
    item -= A.lvl_item0[lvl];
    const int f = item / items_per_frame + A.lvl_frame0[lvl];
    const int wi = item % items_per_frame;
    if (LUT && f != lut_frame) {  // FR_WARPS == 1: the block is this warp
        __syncwarp();
        for (int i = lane; i < 256; i += 32) s_lut[i] = A.lut[f * 256 + i];
        __syncwarp();
        lut_frame = f;
    }
    const int strip = wi % A.nstrips, seg = wi / A.nstrips;
    const u8 *__restrict__ src = A.src + (size_t)f * npx * (BGR ? 3 : 1);
    u8 *__restrict__ map = A.map + (size_t)f * npx;
    u32 *q = A.queue + (size_t)f * npx;
    int *cnt = A.counters + f * LE_C_STRIDE;

    const int ys = seg * rows_per_seg, ye = min(h, ys + rows_per_seg);
    const int x_lane = strip * FR_STRIP - 4 + 4 * lane;  // first pixel of this lane's word
    const bool inside = x_lane >= 0 && x_lane + 3 < w;   // whole word inside the image
    const bool fast_rows = (w & 3) == 0;                 // rows are word aligned (BGR: 3w % 4 == 0 too)
    const bool out_lane = lane >= 1 && lane <= 30 && x_lane < w;
    constexpr int LAGB = BLUR ? 2 : 0;

    // replicate fix-up of the blurred word at the left/right image border (BLUR only)
    u32 fix_sel = 0x3210;
    int fix_src = lane;
    bool need_fix = false;
    if (BLUR) {
        if (x_lane + 3 < 0) { fix_sel = 0x4444; fix_src = lane + 1; }  // all bytes := B(0)
        else if (x_lane >= w) { fix_sel = ((w & 3) == 0) ? 0x7777 : 0x3210; fix_src = lane - 1; }
        else if (x_lane + 3 >= w) {  // word contains W-1 at byte k < 3
            int k = w - 1 - x_lane;
            fix_sel = k == 0 ? 0x0000 : (k == 1 ? 0x1110 : 0x2210);
        }
        fix_src = min(max(fix_src, 0), 31);
        need_fix = __any_sync(~0u, fix_sel != 0x3210);
    }
    // magnitude is 0 outside the image (per 16-bit half)
    u32 mbits = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) mbits |= (x_lane + k >= 0 && x_lane + k < w) ? (1u << k) : 0u;
    const u32 kcmp = (0x7fffu - (u32)lo) * 0x00010001u;
    const bool vec_store = (w & 3) == 0;

    // lean path at the left/right image border: out-of-image words are rebuilt from in-image lanes
    // (REFLECT_101 before a blur, replicate otherwise); their own loads use a clamped address
    const bool xborder = __any_sync(~0u, !inside);
    int gsx = lane, gsy = lane;
    u32 gsel = 0x3210;
    if (x_lane < 0) { gsx = lane + 1; gsy = lane + 2; gsel = BLUR ? 0x1234 : 0x0000; }
    else if (x_lane >= w) {
        int j = (x_lane - w) >> 2;
        if (BLUR) { gsx = lane - 2 - 2 * j; gsy = lane - 1 - 2 * j; gsel = 0x3456; }
        else { gsx = gsy = lane - 1 - j; gsel = 0x7777; }
    }
    gsx = min(max(gsx, 0), 31); gsy = min(max(gsy, 0), 31);
    // everything only border strips need lives in two packed registers (unpacked where used)
    const u32 bord0 = fix_sel | ((u32)fix_src << 16) | (mbits << 24);
    const u32 bord1 = gsel | ((u32)gsx << 16) | ((u32)gsy << 24);
    const int x_ld = min(max(x_lane, 0), max(w - 4, 0));  // load position (always inside the row)

    // ---- general loader of the gray (or pre-Sobel) word of row r: any lane, any row
    auto load_row = [&](int r) -> u32 {
        const int rr = BLUR ? d_reflect101(r, h) : d_clamp(r, 0, h - 1);
        if (inside && fast_rows) {
            if (BGR) {
                const u32 *p = (const u32 *)(src + ((size_t)rr * w + x_lane) * 3);
                return gray4(__ldg(p), __ldg(p + 1), __ldg(p + 2));
            }
            return __ldg((const u32 *)(src + (size_t)rr * w + x_lane));
        }
        u32 g = 0;  // border lanes / unaligned rows: per pixel with the border rule
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int x = x_lane + k;
            int xx = BLUR ? d_reflect101(x, w) : d_clamp(x, 0, w - 1);
            g |= load_px<BGR>(src, rr, xx, w) << (8 * k);
        }
        return g;
    };

    FrontState S;
    memset(&S, 0, sizeof(S));
    Stage T;
    T.s = &s_stage[wib][0][0];
    T.ns = T.nw = 0;

    // One pipeline step: consumes the gray word g of input row r, emits the map word of row r - LAGB - 2.
    // lean == true drops every border rule (valid for interior strips and rows away from the image
    // top/bottom and from the segment warm-up), which is the case for most steps.
    auto step = [&](u32 g, int r, u8 *orow, const bool lean, const bool emit) {
        if (LUT) {
            g = (u32)s_lut[g & 255] | ((u32)s_lut[(g >> 8) & 255] << 8) | ((u32)s_lut[(g >> 16) & 255] << 16) |
                ((u32)s_lut[g >> 24] << 24);
        }
        u32 B;  // pre-Sobel word of row yb
        const int yb = r - LAGB;
        if (BLUR) {
            // vertical [1 4 6 4 1] on pairs, transposed form: output row yb = r - 2
            u32 pl = prmt(g, 0, 0x4140), ph = prmt(g, 0, 0x4342);
            u32 vl = S.a4l + pl, vh = S.a4h + ph;
            S.a4l = S.a3l + 4 * pl; S.a4h = S.a3h + 4 * ph;
            S.a3l = S.a2l + 6 * pl; S.a3h = S.a2h + 6 * ph;
            S.a2l = S.a1l + 4 * pl; S.a2h = S.a1h + 4 * ph;
            S.a1l = pl; S.a1h = ph;
            // horizontal on the vertical sums (<= 4080 per half)
            u32 L23 = __shfl_up_sync(~0u, vh, 1), R01 = __shfl_down_sync(~0u, vl, 1);
            u32 Pm10 = prmt(L23, vl, 0x5432), P12 = prmt(vl, vh, 0x5432), P34 = prmt(vh, R01, 0x5432);
            u32 o01 = L23 + vh + 0x00800080u + 4 * (Pm10 + P12) + 6 * vl;
            u32 o23 = vl + R01 + 0x00800080u + 4 * (P12 + P34) + 6 * vh;
            B = prmt(o01, o23, 0x7531);
            if (need_fix) {  // warp-uniform: only strips that touch the left/right image border
                u32 nbw = __shfl_sync(~0u, B, (bord0 >> 16) & 31);
                B = prmt(B, nbw, bord0 & 0xffffu);
            }
        } else {
            B = g;
        }
        // vertical Sobel on pairs: centre row ysob = yb - 1
        u32 bl = prmt(B, 0, 0x4140), bh = prmt(B, 0, 0x4342);
        if (!lean && BLUR) {
            if (yb == 0) { S.s1l = bl; S.s1h = bh; }  // B(-1) := B(0)
            if (yb >= h) { bl = S.s1l; bh = S.s1h; }  // B(H) := B(H-1)
        }
        const int ysob = yb - 1;
        u32 vsl = S.s2l + bl, vsh = S.s2h + bh;                              // B(y-1) + 2 B(y) + B(y+1)
        u32 vdl = bl + 0x01000100u - S.ppl, vdh = bh + 0x01000100u - S.pph;  // B(y+1) - B(y-1) + 256
        S.s2l = S.s1l + 2 * bl; S.s2h = S.s1h + 2 * bh;
        S.ppl = S.s1l; S.pph = S.s1h;
        S.s1l = bl; S.s1h = bh;
        // horizontal Sobel: neighbours' (vs, vd) of pixels -1 and 4
        u32 LN = __shfl_up_sync(~0u, prmt(vsh, vdh, 0x7632), 1);
        u32 RN = __shfl_down_sync(~0u, prmt(vsl, vdl, 0x5410), 1);
        u32 Sm10 = prmt(LN, vsl, 0x5410), S12 = prmt(vsl, vsh, 0x5432), S34 = prmt(vsh, RN, 0x5432);
        u32 Dm10 = prmt(LN, vdl, 0x5432), D12 = prmt(vdl, vdh, 0x5432), D34 = prmt(vdh, RN, 0x7632);
        u32 gxl = S12 + 0x04000400u - Sm10, gxh = S34 + 0x04000400u - S12;  // gx + 1024
        u32 gyl = Dm10 + 2 * vdl + D12, gyh = D12 + 2 * vdh + D34;          // gy + 1024
        // |gx| + |gy|
        u32 ml = __vmaxu2(gxl, 0x08000800u - gxl) + __vmaxu2(gyl, 0x08000800u - gyl) - 0x08000800u;
        u32 mh = __vmaxu2(gxh, 0x08000800u - gxh) + __vmaxu2(gyh, 0x08000800u - gyh) - 0x08000800u;
        if (!lean) {
            const bool row_in = ysob >= 0 && ysob < h;
            ml = row_in ? ml : 0u;
            mh = row_in ? mh : 0u;
        }
        if (xborder) {  // warp-uniform
            const u32 mb = bord0 >> 24;
            ml &= ((mb & 1u) ? 0xffffu : 0u) | ((mb & 2u) ? 0xffff0000u : 0u);
            mh &= ((mb & 4u) ? 0xffffu : 0u) | ((mb & 8u) ? 0xffff0000u : 0u);
        }
        const bool has_new = ((__vmaxu2(ml, mh) + kcmp) & 0x80008000u) != 0;

        // ---- NMS + threshold for row yo = ysob - 1 (magnitude rows: up, mid, new)
        const int yo = ysob - 1;
        if (emit) {
            u32 outw = 0;
            if (__any_sync(~0u, S.has_mid)) {
                outw = nms_slow(S.mul, S.muh, S.mml, S.mmh, ml, mh, S.gxl, S.gxh, S.gyl, S.gyh, S.has_mid, lo, hi);
                if (!out_lane) outw = 0;
                if (__any_sync(~0u, outw != 0)) {
                    stage_push(outw, (u32)(yo * w + x_lane), T);
                    if (T.ns > FR_STAGE - FR_STRIP) stage_flush(T.s, T.ns, &cnt[LE_C_QTAIL], q, npx, false);
                    if (T.nw > FR_STAGE - FR_STRIP) stage_flush(T.s + FR_STAGE, T.nw, &cnt[LE_C_WEAK], q, npx, true);
                }
            }
            if (out_lane) {
                if (vec_store) *(u32 *)orow = outw;
                else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (x_lane + k < w) orow[k] = (u8)(outw >> (8 * k));
                }
            }
        }
        // rotate the NMS windows
        S.mul = S.mml; S.muh = S.mmh;
        S.mml = ml; S.mmh = mh;
        S.gxl = gxl; S.gxh = gxh; S.gyl = gyl; S.gyh = gyh;
        S.has_mid = has_new;
    };

    const int r0 = ys - 2 - LAGB, r1 = ye + 1 + LAGB;  // inclusive input row range
    // lean steps: interior strip with word-aligned rows; output enabled (r >= ys + LAGB + 2), no top
    // rule (r >= LAGB + 1) and no bottom rule / all ring rows readable (r <= h - 3)
    const bool strip_lean = fast_rows && w >= 16;
    int ra = max(r0, LAGB + 1), rb = min(r1, h - 3);
    if (!strip_lean || rb < ra) { ra = r1 + 1; rb = r1; }
    auto emit_row = [&](int rr) { int yo = rr - LAGB - 2; return yo >= ys && yo < ye; };
    u8 *orow = map + ((long long)(r0 - LAGB - 2) * w + x_lane);  // output row of the step (dereferenced only when valid)
    int r = r0;
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        // phase 0: rows before the lean range; phase 1: rows after it (one code instance for both)
        const int pe = phase == 0 ? min(ra - 1, r1) : r1;
#pragma unroll 1
        for (; r <= pe; r++, orow += w) step(load_row(r), r, orow, false, emit_row(r));
        if (phase == 0 && r <= rb) {
            // steady state: rows stream through a per-warp shared-memory ring
            constexpr int WPL = BGR ? 3 : 1;  // words per lane per row
            const u32 pitch = (u32)w * WPL;   // bytes per row
            u32 *ring = &s_ring[wib][0];
            // Flat-row skip: once FR_FLAT consecutive rows of this strip were one constant, every
            // pipeline register sits at its fixed point (partial sums of a constant, zero magnitudes),
            // so a further identical row leaves the state untouched and its output row is zero.
            u32 pa = 0, pb = 0, pc = 0, gprev = 0;
            int flat = 0;
            const int r_emit = ys + LAGB + 2;  // first input row whose step emits an output row
            auto gray_of = [&](u32 ca, u32 cb, u32 cc) -> u32 {
                u32 g = BGR ? gray4(ca, cb, cc) : ca;
                if (xborder) {  // warp-uniform
                    u32 gx_ = __shfl_sync(~0u, g, (bord1 >> 16) & 31), gy_ = __shfl_sync(~0u, g, bord1 >> 24);
                    g = prmt(gx_, gy_, bord1 & 0xffffu);
                }
                return g;
            };
            // The segment starts at r0, so the pipeline state there is "don't care": the first emitted
            // row (r0 + 8 with a blur) depends on rows >= r0 only.  If the first row is one constant,
            // start AT the fixed point of that constant instead of paying nine full warm-up steps to
            // reach it -- flat strips (most of a cube-edge frame) then never run a full step.
            auto set_fixed = [&](u32 g) {  // pipeline state after >= FR_FLAT rows of the constant word g
                u32 c = g & 255u;
                if (LUT) c = s_lut[c];
                const u32 P = c * 0x00010001u;
                if (BLUR) {
                    S.a1l = S.a1h = P; S.a2l = S.a2h = 5 * P; S.a3l = S.a3h = 11 * P; S.a4l = S.a4h = 15 * P;
                }
                S.s1l = S.s1h = P; S.ppl = S.pph = P; S.s2l = S.s2h = 3 * P;
                S.gxl = S.gxh = S.gyl = S.gyh = 0x04000400u;
                S.mul = S.muh = S.mml = S.mmh = 0u;
                S.has_mid = false;
            };
            auto fixed_start = [&](u32 ca, u32 cb, u32 cc) {
                const u32 g = gray_of(ca, cb, cc);
                int all_eq;
                __match_all_sync(~0u, g, &all_eq);
                if (all_eq && g == prmt(g, 0, 0x0000)) {
                    pa = ca; pb = cb; pc = cc;
                    gprev = g;
                    flat = FR_FLAT;
                    if (!TMA) set_fixed(g);  // (the TMA loop rebuilds the state when it leaves flat mode)
                }
            };
            auto full_row = [&](u32 ca, u32 cb, u32 cc) {  // one full pipeline step on row r
                pa = ca; pb = cb; pc = cc;
                const u32 g = gray_of(ca, cb, cc);
                int all_eq;
                __match_all_sync(~0u, g, &all_eq);
                const bool row_const = all_eq && g == prmt(g, 0, 0x0000);  // warp-uniform
                flat = row_const ? ((flat > 0 && g == gprev) ? flat + 1 : 1) : 0;
                gprev = g;
                step(g, r, orow, true, emit_row(r));
                r++;
                orow += w;
            };
            auto zero4 = [&]() {  // four flat output rows (inside the segment; lean implies w % 4 == 0)
                if (out_lane) {
                    u8 *o = orow;
#pragma unroll
                    for (int k = 0; k < 4; k++, o += w) *(u32 *)o = 0u;
                }
                r += 4;
                orow += 4 * (size_t)w;
            };
            if constexpr (TMA) {
                // TMA ring: FR_NS stages of four rows.  Lane 0 arms the stage's mbarrier with the box size
                // and issues ONE 2-D tensor copy (four rows x 100 words, starting at the 16-byte boundary
                // below the strip's first byte; columns/rows outside the batch are zero-filled); the
                // copy engine moves the bytes -- no LSU work and nothing to wait for per lane -- and
                // every lane waits on the stage parity.  Stage and parity follow from the row: ring row
                // k = r - rs belongs to group gbase + k / 4, which lives in stage (group % FR_NS) on its
                // (group / FR_NS)-th use -- no loop-carried ring state besides gbase.
                const int xb = (strip * FR_STRIP - 4) * WPL;  // first byte of the strip in the row (may be < 0)
                const int a0 = xb & ~15;
                const u32 ring_s = (u32)__cvta_generic_to_shared(ring);
                const u32 *rd = ring + ((xb - a0) >> 2) + lane * WPL;  // this lane's words of ring row 0
                const int rs = r;
                const int ybase = f * h + rs, ylast = f * h + rb;  // batch rows of the first / last ring row
                auto issue = [&](u32 k4) {  // request ring rows 4 * k4 .. 4 * k4 + 3
                    const int y = ybase + 4 * (int)k4;
                    if (lane == 0 && y <= ylast) {
                        const u32 stg = (gbase + k4) & (FR_NS - 1);
                        mbar_expect_tx(bar_s + 8 * stg, 4 * SLOTB);
                        tma_load_2d(ring_s + stg * STAGEB, &tmap, a0 >> 2, y, bar_s + 8 * stg);
                    }
                };
                auto acquire = [&](u32 k4) -> const u32 * {  // wait for group k4, return this lane's words
                    const u32 G = gbase + k4, stg = G & (FR_NS - 1);
                    mbar_wait(bar_s + 8 * stg, (G / FR_NS) & 1u);
                    return rd + stg * (STAGEB / 4);
                };
#pragma unroll
                for (int i = 0; i < FR_NS; i++) issue(i);
                if (r == r0) {
                    const u32 *sp = acquire(0);
                    fixed_start(sp[0], BGR ? sp[1] : 0, BGR ? sp[2] : 0);
                }
#pragma unroll 1
                while (r <= rb) {
                    if (flat >= FR_FLAT) {
                        // ---- flat mode: the pipeline state is the fixed point of the constant gprev and is
                        // not touched (nor kept in registers) here; it is rebuilt when the mode ends.
#pragma unroll 1
                        while (r <= rb) {
                            const u32 k = (u32)(r - rs);
                            const u32 *sp = acquire(k >> 2);
                            if ((k & 3u) == 0 && r >= r_emit && r + 3 <= rb) {
                                u32 diff = 0;
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    diff |= sp[j * (SLOTB / 4)] ^ pa;
                                    if (BGR) diff |= (sp[j * (SLOTB / 4) + 1] ^ pb) | (sp[j * (SLOTB / 4) + 2] ^ pc);
                                }
                                if (__all_sync(~0u, diff == 0)) {  // four flat rows: the whole stage at once
                                    zero4();
                                    __syncwarp();
                                    issue((k >> 2) + FR_NS);
                                    continue;
                                }
                            }
                            sp += (k & 3u) * (SLOTB / 4);
                            const u32 ca = sp[0], cb = BGR ? sp[1] : 0, cc = BGR ? sp[2] : 0;
                            if (!__all_sync(~0u, ((ca ^ pa) | (cb ^ pb) | (cc ^ pc)) == 0)) break;
                            if (out_lane && r >= r_emit) *(u32 *)orow = 0u;
                            r++;
                            orow += w;
                            if ((k & 3u) == 3u) {
                                __syncwarp();
                                issue((k >> 2) + FR_NS);
                            }
                        }
                        set_fixed(gprev);  // (also when the ring rows are used up: the rows after them run full steps)
                        if (r > rb) break;
                    }
                    // ---- full steps until the strip has been flat for FR_FLAT rows again
#pragma unroll 1
                    do {
                        const u32 k = (u32)(r - rs);
                        const u32 *sp = acquire(k >> 2) + (k & 3u) * (SLOTB / 4);
                        full_row(sp[0], BGR ? sp[1] : 0, BGR ? sp[2] : 0);
                        if ((k & 3u) == 3u) {
                            __syncwarp();
                            issue((k >> 2) + FR_NS);
                        }
                    } while (r <= rb && flat < FR_FLAT);
                }
                gbase += (u32)(rb - rs + 4) >> 2;  // groups of this item (all requested, all consumed)
            } else {
                // cp.async (LDGSTS) ring, one row per commit group, FR_DEPTH - 1 rows ahead
                const u8 *gp = src + ((size_t)r * w + x_ld) * WPL;
                const u32 ring_s = (u32)__cvta_generic_to_shared(ring) + lane * WPL * 4;
                constexpr u32 ROWB = SLOTB;
                int issue_r = r;  // next row to request
                auto refill = [&]() {  // request row issue_r into the slot consumed one iteration ago
                    if (issue_r <= rb) {
                        u32 sa = ring_s + (u32)(issue_r & (FR_DEPTH - 1)) * ROWB;
#pragma unroll
                        for (int k = 0; k < WPL; k++) cp_async4(sa + 4 * k, gp + 4 * k);
                        gp += pitch;
                    }
                    issue_r++;
                    cp_async_commit();
                };
#pragma unroll
                for (int d = 0; d < FR_DEPTH - 1; d++) refill();
                if (r == r0) {
                    cp_async_wait<FR_DEPTH - 2>();
                    const u32 *slot = ring + (r & (FR_DEPTH - 1)) * (ROWB / 4) + lane * WPL;
                    fixed_start(slot[0], BGR ? slot[1] : 0, BGR ? slot[2] : 0);
                }
#pragma unroll 1
                while (r <= rb) {
                    if (flat >= FR_FLAT && (r & 3) == 0 && r >= r_emit) {
                        // Four flat rows per trip: the ring slots of rows r..r+3 are consecutive, one vote
                        // decides for all four, the four freed slots (rows r-1..r+2) are refilled together.
#pragma unroll 1
                        while (issue_r + 3 <= rb) {
                            cp_async_wait<FR_DEPTH - 5>();  // rows r .. r+3 have landed
                            const u32 *s4 = ring + (r & (FR_DEPTH - 1)) * (ROWB / 4) + lane * WPL;
                            u32 diff = 0;
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                diff |= s4[k * (ROWB / 4)] ^ pa;
                                if (BGR) diff |= (s4[k * (ROWB / 4) + 1] ^ pb) | (s4[k * (ROWB / 4) + 2] ^ pc);
                            }
                            if (!__all_sync(~0u, diff == 0)) break;
                            {   // rows issue_r .. issue_r+3 go to slots (r-1), r, r+1, r+2
                                const u32 sb = ring_s + (u32)(r & (FR_DEPTH - 1)) * ROWB;
                                const u32 s0 = ring_s + (u32)((r - 1) & (FR_DEPTH - 1)) * ROWB;
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const u32 sa = j == 0 ? s0 : sb + (u32)(j - 1) * ROWB;
#pragma unroll
                                    for (int k = 0; k < WPL; k++) cp_async4(sa + 4 * k, gp + 4 * k);
                                    gp += pitch;
                                    cp_async_commit();
                                }
                                issue_r += 4;
                            }
                            zero4();
                        }
                    }
                    cp_async_wait<FR_DEPTH - 2>();  // the group of row r has landed
                    const u32 *slot = ring + (r & (FR_DEPTH - 1)) * (ROWB / 4) + lane * WPL;
                    u32 ca = slot[0], cb = BGR ? slot[1] : 0, cc = BGR ? slot[2] : 0;
                    refill();
                    if (flat >= FR_FLAT && __all_sync(~0u, ((ca ^ pa) | (cb ^ pb) | (cc ^ pc)) == 0)) {
                        // one flat row (alignment to a multiple of four, and the rows left over at the end)
                        if (out_lane && r >= r_emit) *(u32 *)orow = 0u;
                        r++;
                        orow += w;
                        continue;
                    }
                    full_row(ca, cb, cc);
                }
                cp_async_wait<0>();
            }
        }
    }
    if (T.ns) stage_flush(T.s, T.ns, &cnt[LE_C_QTAIL], q, npx, false);
    if (T.nw) stage_flush(T.s + FR_STAGE, T.nw, &cnt[LE_C_WEAK], q, npx, true);
    }  // work items
}

// Tensor map of the input batch for the TMA ring: 2-D, 32-bit elements, [n*h rows][row bytes / 4], box =
// 4 rows x (SLOTB / 4) words.  cuTensorMapEncodeTiled is taken from the driver at run time (no -lcuda).
static int make_input_tmap(le_ctx *ctx, CUtensorMap *tm, const u8 *src, int n, int h, int row_bytes, int slot_bytes)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LE_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return le_fail(ctx, LE_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = (encode_fn)fn;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)row_bytes / 4, (cuuint64_t)n * h};
    cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)slot_bytes / 4, 4};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)src, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return le_fail(ctx, LE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return LE_OK;
}

template <bool BGR, bool BLUR, bool LUT, bool TMA>
static int launch_rows_t(le_ctx *ctx, const u8 *src, u8 *edges, int n, int h, int w, int lo, int hi, cudaStream_t st)
{
    FrontArgs A;
    A.src = src; A.map = edges; A.queue = (u32 *)ctx->queue.p; A.counters = (int *)ctx->counters.p;
    A.lut = (const u8 *)ctx->lut.p;
    A.h = h; A.w = w; A.lo = max(0, min(lo, 4000)); A.hi = hi;
    A.nstrips = (w + FR_STRIP - 1) / FR_STRIP;
    // enough warps to fill the machine even for one small frame, long segments for big batches
    long long target_warps = (long long)ctx->sm_count * 32;
    int rows = 160;
    if (const int e = LE_KNOB_INT("LE_FRONT_ROWS", 0)) rows = max(8, e);  // A/B measurements
    while (rows > 8 && (long long)n * A.nstrips * ((h + rows - 1) / rows) < target_warps) rows /= 2;
    A.rows_per_seg = rows;
    A.nsegs = (h + rows - 1) / rows;
    A.nframes = n;
    A.work = (int *)ctx->counters.p + (size_t)n * LE_C_STRIDE;  // zeroed with the frame counters
    // Guided schedule.  A big batch: the first half of the frames in `rows`-row segments, then a quarter
    // in rows/2, the rest in rows/4 (each segment re-runs 8 warm-up rows, so short segments are only worth
    // it where they shorten the tail; measured best of 1..4 levels x 90..360 rows).  A small batch: one level.
    int nlv = (n >= 16 && rows >= 64) ? FR_LEVELS - 1 : 1;
    if (const int e = LE_KNOB_INT("LE_FRONT_LEVELS", 0)) nlv = max(1, min(FR_LEVELS, e));  // A/B measurements
    long long items = 0;
    int f0 = 0;
    for (int l = 0; l < FR_LEVELS; l++) {
        int f1 = (l >= nlv - 1) ? n : n - (n >> (l + 1));
        if (l >= nlv) f1 = n;
        A.lvl_frame0[l] = f0;
        A.lvl_rows[l] = max(8, rows >> min(l, nlv - 1));
        A.lvl_nsegs[l] = (h + A.lvl_rows[l] - 1) / A.lvl_rows[l];
        A.lvl_item0[l] = (int)items;
        items += (long long)A.nstrips * A.lvl_nsegs[l] * (f1 - f0);
        f0 = f1;
    }
    A.lvl_frame0[FR_LEVELS] = n;
    A.lvl_item0[FR_LEVELS] = (int)items;
    int per_sm = FR_MINBLK;
    if (const int e = LE_KNOB_INT("LE_FRONT_BLOCKS_PER_SM", 0)) per_sm = max(1, min(FR_MINBLK, e));  // A/B measurements
    dim3 grid((unsigned)min(items, (long long)ctx->sm_count * per_sm));
    LE_T0(ctx, LE_K_FRONT, st);
    static le_dev_once carveout_set;  // all of the SM's shared memory: 24 one-warp blocks with their rings
    if (carveout_set.needs(ctx->device, 1)) {
        cudaFuncSetAttribute(k_front_rows<BGR, BLUR, LUT, TMA>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        carveout_set.done(ctx->device, 1);
    }
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (TMA) {
        int rc = make_input_tmap(ctx, &tm, src, n, h, w * (BGR ? 3 : 1), BGR ? 400 : 144);
        if (rc != LE_OK) return rc;
    }
    k_front_rows<BGR, BLUR, LUT, TMA><<<grid, FR_WARPS * 32, 0, st>>>(A, tm);
    LE_T1(ctx, LE_K_FRONT, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

template <bool TMA>
static int launch_rows_v(le_ctx *ctx, const u8 *src, int ch, int blur, int lut, u8 *edges, int n, int h, int w, int lo,
                         int hi, cudaStream_t st)
{
    if (lut) return launch_rows_t<false, false, true, TMA>(ctx, src, edges, n, h, w, lo, hi, st);
    if (ch == 3 && blur) return launch_rows_t<true, true, false, TMA>(ctx, src, edges, n, h, w, lo, hi, st);
    if (ch == 3) return launch_rows_t<true, false, false, TMA>(ctx, src, edges, n, h, w, lo, hi, st);
    if (blur) return launch_rows_t<false, true, false, TMA>(ctx, src, edges, n, h, w, lo, hi, st);
    return launch_rows_t<false, false, false, TMA>(ctx, src, edges, n, h, w, lo, hi, st);
}

int le_front_rows_launch(le_ctx *ctx, const u8 *src, int ch, int blur, int lut, u8 *edges, int n, int h, int w, int lo,
                         int hi, cudaStream_t st)
{
    // Bulk (TMA) copies need 16-byte aligned sources and sizes: rows that are 16-byte multiples on a
    // 16-byte aligned base.  A ring row starts up to 12 bytes before the strip and ends up to 400 bytes
    // after its first byte, which must stay inside the frame for rows <= h - 3: w >= 160 covers it.
    const bool tma = (w % 16) == 0 && w >= 160 && ((uintptr_t)src % 16) == 0 && !LE_KNOB_SET("LE_FRONT_NO_TMA");
    if (tma) return launch_rows_v<true>(ctx, src, ch, blur, lut, edges, n, h, w, lo, hi, st);
    return launch_rows_v<false>(ctx, src, ch, blur, lut, edges, n, h, w, lo, hi, st);
}
