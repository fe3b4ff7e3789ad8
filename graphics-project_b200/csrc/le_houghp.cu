// le_houghp.cu -- progressive probabilistic Hough transform (a6), exact replay of cv2's order.
//
// PPHT is sequential by definition: every accepted segment erases pixels and votes that later
// random picks see.  The parallelism used here is (1) frames: one CTA per frame, several CTAs per
// SM so their memory latencies overlap; (2) angles: thread n owns accumulator row n, so the
// 180-angle vote of a pick is one parallel step with NO atomics (a row has a single writer) and
// the arg-max is a block reduction; (3) the line walk: a warp evaluates 32 steps of the Q16
// fixed-point walk at once and resolves the gap rule with ballots.
// The accumulator keeps only the rho support of each angle (s16 [numangle][stride], ~0.8 MB at
// 1080p instead of cv2's 4.3 MB int32) so the working set of all resident frames stays in L2.
#include "le_common.cuh"
#include <stdlib.h>

int le_ppht_fast_launch(le_ctx *ctx, int n, int h, int w, int numangle, int stride, int threshold, int min_len,
                        int max_gap, int32_t *segs, int32_t *counts, int max_segs, int32_t *trace, int trace_cap,
                        cudaStream_t st);
int le_ppht_rounds_launch(le_ctx *ctx, int n, int h, int w, int numangle, int stride, int threshold, int min_len,
                          int max_gap, int32_t *segs, int32_t *counts, int max_segs, cudaStream_t st);
int le_hough_upload_trig(le_ctx *ctx, int kind, int h, int w, double rho, double theta, int numangle, int *stride_out,
                         cudaStream_t st);

#define NZ_SEG 8192
// pass 1: count non-zeros per segment of NZ_SEG pixels
__global__ void __launch_bounds__(256) k_nz_count(const u8 *__restrict__ edges, int *__restrict__ segcnt, size_t npx,
                                                  int nseg)
{
    const int f = blockIdx.y, s = blockIdx.x;
    const u8 *e = edges + (size_t)f * npx;
    size_t b0 = (size_t)s * NZ_SEG, b1 = min(npx, b0 + NZ_SEG);
    int c = 0;
    for (size_t i = b0 + threadIdx.x; i < b1; i += 256) c += e[i] != 0;
    __shared__ int red[8];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(~0u, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; k++) t += red[k];
        segcnt[f * nseg + s] = t;
    }
}
// pass 2: ordered (row-major) write of packed (y<<16|x) + mask bytes
// The same two passes with 16 pixels per thread (one 16-byte load, __vsetne4 gives the mask bytes and, by popc,
// the count): for frames whose size is a multiple of 16 bytes (1080p, 4K, 600 x 400).  Same list, same mask.
__global__ void __launch_bounds__(256) k_nz_count16(const u8 *__restrict__ edges, int *__restrict__ segcnt, size_t npx,
                                                    int nseg)
{
    const int f = blockIdx.y, s = blockIdx.x;
    const uint4 *e = (const uint4 *)(edges + (size_t)f * npx);
    const size_t v0 = (size_t)s * (NZ_SEG / 16), v1 = min(npx / 16, v0 + NZ_SEG / 16);
    int c = 0;
    for (size_t i = v0 + threadIdx.x; i < v1; i += 256) {
        const uint4 v = e[i];
        c += __popc(__vsetne4(v.x, 0)) + __popc(__vsetne4(v.y, 0)) + __popc(__vsetne4(v.z, 0)) + __popc(__vsetne4(v.w, 0));
    }
    __shared__ int red[8];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(~0u, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; k++) t += red[k];
        segcnt[f * nseg + s] = t;
    }
}
__global__ void __launch_bounds__(256) k_nz_write16(const u8 *__restrict__ edges, const int *__restrict__ segcnt,
                                                    u32 *__restrict__ nz, u8 *__restrict__ mask,
                                                    int *__restrict__ counters, size_t npx, int nseg, int w)
{
    const int f = blockIdx.y, s = blockIdx.x;
    const uint4 *e = (const uint4 *)(edges + (size_t)f * npx);
    uint4 *mk = (uint4 *)(mask + (size_t)f * npx);
    u32 *out = nz + (size_t)f * npx;
    __shared__ int s_base, warp_off[8];
    int part = 0;  // exclusive prefix of the segment counts
    for (int k = threadIdx.x; k < s; k += 256) part += segcnt[f * nseg + k];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(~0u, part, o);
    if ((threadIdx.x & 31) == 0) warp_off[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; k++) t += warp_off[k];
        s_base = t;
        if (s == nseg - 1) counters[f * LE_C_STRIDE + LE_C_AUX0] = t + segcnt[f * nseg + s];
    }
    __syncthreads();
    int base = s_base;
    const size_t v0 = (size_t)s * (NZ_SEG / 16), v1 = min(npx / 16, v0 + NZ_SEG / 16);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (size_t c0 = v0; c0 < v1; c0 += 256) {
        const size_t i = c0 + threadIdx.x;
        u32 m[4] = {0, 0, 0, 0};
        if (i < v1) {
            const uint4 v = e[i];
            m[0] = __vsetne4(v.x, 0); m[1] = __vsetne4(v.y, 0); m[2] = __vsetne4(v.z, 0); m[3] = __vsetne4(v.w, 0);
            mk[i] = make_uint4(m[0], m[1], m[2], m[3]);
        }
        const int cnt = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
        int inc = cnt;  // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(~0u, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();
        if (lane == 31) warp_off[wid] = inc;
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < wid) off += warp_off[k];
            tot += warp_off[k];
        }
        if (cnt) {
            int pos = base + off + inc - cnt;
            const size_t p0 = i * 16;
            const int y = (int)(p0 / w), x = (int)(p0 - (size_t)y * w);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!m[j]) continue;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (m[j] >> (8 * k) & 1) {
                        int xx = x + 4 * j + k, yy = y;
                        while (xx >= w) { xx -= w; yy++; }
                        out[pos++] = ((u32)yy << 16) | (u32)xx;
                    }
            }
        }
        base += tot;
    }
}
__global__ void __launch_bounds__(256) k_nz_write(const u8 *__restrict__ edges, const int *__restrict__ segcnt,
                                                  u32 *__restrict__ nz, u8 *__restrict__ mask,
                                                  int *__restrict__ counters, size_t npx, int nseg, int w)
{
    const int f = blockIdx.y, s = blockIdx.x;
    const u8 *e = edges + (size_t)f * npx;
    u8 *mk = mask + (size_t)f * npx;
    u32 *out = nz + (size_t)f * npx;
    __shared__ int s_base, warp_off[8];
    // exclusive prefix of the segment counts
    int part = 0;
    for (int k = threadIdx.x; k < s; k += 256) part += segcnt[f * nseg + k];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(~0u, part, o);
    if ((threadIdx.x & 31) == 0) warp_off[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; k++) t += warp_off[k];
        s_base = t;
        if (s == nseg - 1) counters[f * LE_C_STRIDE + LE_C_AUX0] = t + segcnt[f * nseg + s];
    }
    __syncthreads();
    int base = s_base;
    size_t b0 = (size_t)s * NZ_SEG, b1 = min(npx, b0 + NZ_SEG);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (size_t c0 = b0; c0 < b1; c0 += 256) {
        size_t i = c0 + threadIdx.x;
        bool on = i < b1 && e[i] != 0;
        if (i < b1) mk[i] = on ? 1 : 0;
        u32 bal = __ballot_sync(~0u, on);
        __syncthreads();
        if (lane == 0) warp_off[wid] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int k = 0; k < 8; k++) {
            if (k < wid) off += warp_off[k];
            tot += warp_off[k];
        }
        if (on) {
            int y = (int)(i / w), x = (int)(i - (size_t)y * w);
            out[base + off + __popc(bal & ((1u << lane) - 1))] = ((u32)y << 16) | (u32)x;
        }
        base += tot;
    }
}

#define PP_THREADS 192
#define PP_NZ_SMEM 12288
#define PP_MAXT_WORDS 512  // up to 16384 steps per direction

struct PPShared {
    u32 nz[PP_NZ_SMEM];
    u32 walk_bits[2][PP_MAXT_WORDS];
    unsigned long long red[PP_THREADS / 32];
    int t_end[2];
    int end_xy[2][2];
    int pick[2];
    int nlines;
    int ntrace;
};

static __device__ __forceinline__ u32 rng_next(u64 &st)
{
    st = (u64)(u32)st * 4164903690ULL + (st >> 32);
    return (u32)st;
}

__global__ void __launch_bounds__(PP_THREADS)
k_ppht(u8 *__restrict__ mask, u32 *__restrict__ nz_g, short *__restrict__ accum, const float *__restrict__ trig,
       int *__restrict__ counters, int32_t *__restrict__ segs, int32_t *__restrict__ counts, int max_segs, int h,
       int w, int numangle, int stride, int threshold, int line_length, int line_gap, int32_t *trace, int trace_cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PPShared &S = *reinterpret_cast<PPShared *>(smem_raw);
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t npx = (size_t)h * w;
    u8 *mk = mask + (size_t)f * npx;
    short *acc = accum + (size_t)f * numangle * stride;  // signed: un-voting never-voted pixels goes negative (as in cv2)
    int count = counters[f * LE_C_STRIDE + LE_C_AUX0];
    const int *rlo_g = (const int *)(trig + 2 * numangle);
    // nz list: shared when it fits
    u32 *nz = nz_g + (size_t)f * npx;
    if (count <= PP_NZ_SMEM) {
        for (int i = tid; i < count; i += PP_THREADS) S.nz[i] = nz[i];
        nz = S.nz;
    }
    if (tid == 0) { S.nlines = 0; S.ntrace = 0; }
    // angle owned by this thread (generic numangle handled by a strided loop; 180 -> one each)
    const bool single = numangle <= PP_THREADS;
    float tc = 0.f, ts = 0.f;
    int rlo = 0;
    if (tid < numangle) {
        tc = trig[tid];
        ts = trig[numangle + tid];
        rlo = rlo_g[tid];
    }
    u64 rng = ~0ULL;
    int par = 0;
    long long dbg_t[4] = {0, 0, 0, 0}, dbg_picks = 0, dbg_votes = 0;  // only touched when tracing
    __syncthreads();
    const int shift = 16;
    for (;;) {
        // thread 0 is the producer: it replays cv::RNG, maintains the swap-remove list and skips
        // picks whose pixel was already consumed (no barrier needed for those)
        long long t_a = 0;
        if (trace && tid == 0) t_a = clock64();
        if (tid == 0) {
            int pick = -1;
            while (count > 0) {
                if (trace) dbg_picks++;
                int idx = (int)(rng_next(rng) % (u32)count);
                u32 pt = nz[idx];
                nz[idx] = nz[count - 1];
                count--;
                if (mk[(size_t)(pt >> 16) * w + (pt & 0xffff)]) { pick = (int)pt; break; }
            }
            S.pick[par] = pick;
        }
        __syncthreads();
        if (trace && tid == 0) { long long t = clock64(); dbg_t[0] += t - t_a; t_a = t; }
        const int pk = S.pick[par];
        par ^= 1;
        if (pk < 0) break;
        const int j = pk & 0xffff, i = (int)((u32)pk >> 16);
        // ---- vote
        unsigned long long best = 0;
        if (single) {
            if (tid < numangle) {
                int r = __float2int_rn(__fadd_rn(__fmul_rn((float)j, tc), __fmul_rn((float)i, ts)));
                short *cell = acc + (size_t)tid * stride + (r - rlo);
                int val = *cell + 1;
                *cell = (short)val;
                best = ((unsigned long long)(u32)(val + 0x8000) << 32) | (u32)(0x7fffffff - tid);
            }
        } else {
            for (int n = tid; n < numangle; n += PP_THREADS) {
                int r = __float2int_rn(__fadd_rn(__fmul_rn((float)j, trig[n]), __fmul_rn((float)i, trig[numangle + n])));
                short *cell = acc + (size_t)n * stride + (r - rlo_g[n]);
                int val = *cell + 1;
                *cell = (short)val;
                unsigned long long key = ((unsigned long long)(u32)(val + 0x8000) << 32) | (u32)(0x7fffffff - n);
                best = key > best ? key : best;
            }
        }
        for (int o = 16; o; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(~0u, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) S.red[wid] = best;
        __syncthreads();
        best = S.red[0];
#pragma unroll
        for (int k = 1; k < PP_THREADS / 32; k++) best = S.red[k] > best ? S.red[k] : best;
        const int max_val = (int)(best >> 32) - 0x8000;  // biased so that negative cells order correctly
        const int max_n = 0x7fffffff - (int)(u32)best;
        if (trace && tid == 0) { long long t = clock64(); dbg_t[1] += t - t_a; t_a = t; dbg_votes++; }
        if (max_val < threshold) continue;  // uniform (needs no barrier: S.red is rewritten only after the next sync)
        // ---- walk set-up (all threads, identical)
        const float a = -trig[numangle + max_n], b = trig[max_n];
        int x0 = j, y0 = i, dx0, dy0;
        bool xflag;
        if (fabsf(a) > fabsf(b)) {
            xflag = true;
            dx0 = a > 0 ? 1 : -1;
            dy0 = __float2int_rn(__fdiv_rn(__fmul_rn(b, 65536.f), fabsf(a)));
            y0 = (y0 << shift) + (1 << (shift - 1));
        } else {
            xflag = false;
            dy0 = b > 0 ? 1 : -1;
            dx0 = __float2int_rn(__fdiv_rn(__fmul_rn(a, 65536.f), fabsf(b)));
            x0 = (x0 << shift) + (1 << (shift - 1));
        }
        // ---- first walk: warp k handles direction k, 32 steps at a time
        if (wid < 2) {
            const int k = wid;
            const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
            int gap = 0, t_end = 0, ex = j, ey = i;
            bool stop = false;
            for (int base = 0; !stop; base += 32) {
                int t = base + lane;
                int x = x0 + t * dx, y = y0 + t * dy;
                int j1 = xflag ? x : (x >> shift), i1 = xflag ? (y >> shift) : y;
                bool inb = j1 >= 0 && j1 < w && i1 >= 0 && i1 < h;
                bool on = inb && mk[(size_t)i1 * w + j1] != 0;
                u32 bm = __ballot_sync(~0u, on);
                u32 below = bm & ((1u << lane) - 1);
                int run = below ? lane - (31 - __clz(below)) : gap + lane + 1;
                bool st = !inb || (!on && run > line_gap);
                u32 bs = __ballot_sync(~0u, st);
                u32 valid = bs ? ((1u << (__ffs(bs) - 1)) - 1) : ~0u;  // steps strictly before the stop
                u32 bmv = bm & valid;
                if ((base >> 5) < PP_MAXT_WORDS && lane == 0) S.walk_bits[k][base >> 5] = bmv;
                if (bmv) {
                    int hb = 31 - __clz(bmv);
                    t_end = base + hb;
                    ex = __shfl_sync(~0u, j1, hb);
                    ey = __shfl_sync(~0u, i1, hb);
                    gap = 31 - hb;
                } else gap += 32;
                stop = bs != 0;
            }
            if (lane == 0) {
                S.t_end[k] = t_end;
                S.end_xy[k][0] = ex;
                S.end_xy[k][1] = ey;
            }
        }
        __syncthreads();
        if (trace && tid == 0) { long long t = clock64(); dbg_t[2] += t - t_a; t_a = t; }
        const int e0x = S.end_xy[0][0], e0y = S.end_xy[0][1], e1x = S.end_xy[1][0], e1y = S.end_xy[1][1];
        const bool good = abs(e1x - e0x) >= line_length || abs(e1y - e0y) >= line_length;
        if (trace && f == 0 && tid == 0) {
            int tp = S.ntrace++;
            if (tp < trace_cap) {
                int32_t *o = trace + tp * 10;
                o[0] = j; o[1] = i; o[2] = max_val; o[3] = max_n; o[4] = e0x; o[5] = e0y; o[6] = e1x; o[7] = e1y;
                o[8] = good; o[9] = dx0 * 4 + dy0 % 4;
            }
        }
        // ---- second walk: clear the mask; un-vote if the segment is accepted
        for (int k = 0; k < 2; k++) {
            const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
            const int t_end = S.t_end[k];
            const int nwords = (t_end >> 5) + 1;
            for (int wd = 0; wd < nwords; wd++) {
                u32 bits = S.walk_bits[k][wd];
                if (k == 1 && wd == 0) bits &= ~1u;  // the start pixel was handled by direction 0
                while (bits) {
                    int bpos = __ffs(bits) - 1;
                    bits &= bits - 1;
                    int t = wd * 32 + bpos;
                    int x = x0 + t * dx, y = y0 + t * dy;
                    int j1 = xflag ? x : (x >> shift), i1 = xflag ? (y >> shift) : y;
                    if (good) {
                        if (single) {
                            if (tid < numangle) {
                                int r = __float2int_rn(__fadd_rn(__fmul_rn((float)j1, tc), __fmul_rn((float)i1, ts)));
                                acc[(size_t)tid * stride + (r - rlo)]--;
                            }
                        } else {
                            for (int n = tid; n < numangle; n += PP_THREADS) {
                                int r = __float2int_rn(
                                    __fadd_rn(__fmul_rn((float)j1, trig[n]), __fmul_rn((float)i1, trig[numangle + n])));
                                acc[(size_t)n * stride + (r - rlo_g[n])]--;
                            }
                        }
                    }
                    if (tid == 0) mk[(size_t)i1 * w + j1] = 0;
                }
            }
        }
        if (trace && tid == 0) { long long t = clock64(); dbg_t[3] += t - t_a; t_a = t; }
        if (good && tid == 0) {
            int pos = S.nlines++;
            if (pos < max_segs) {
                int32_t *o = segs + ((size_t)f * max_segs + pos) * 4;
                o[0] = e0x; o[1] = e0y; o[2] = e1x; o[3] = e1y;
            }
        }
    }
    __syncthreads();
    if (tid == 0) counts[f] = S.nlines;
    if (trace && tid == 0 && f == 0 && trace_cap >= 2) {  // debug: phase cycles in the last trace record
        long long *o = (long long *)(trace + (size_t)(trace_cap - 2) * 10);
        o[0] = dbg_t[0]; o[1] = dbg_t[1]; o[2] = dbg_t[2]; o[3] = dbg_t[3]; o[4] = dbg_picks; o[5] = dbg_votes;
    }
}

// debug hook (tests only): record the voting picks of frame 0 -> 10 ints each
static int32_t *g_ppht_trace = nullptr;
static int g_ppht_trace_cap = 0;
extern "C" void le_debug_ppht_trace(int32_t *dev_buf, int cap)
{
    g_ppht_trace = dev_buf;
    g_ppht_trace_cap = cap;
}

extern "C" int le_hough_lines_p(le_ctx *ctx, const uint8_t *edges, int n, int h, int w, double rho, double theta,
                                int threshold, int min_line_length, int max_line_gap, int32_t *segs,
                                int32_t *counts, int max_segs, int exact_order, void *stream)
{
    LE_ON_DEVICE(ctx);
    if (!ctx) return LE_ERR_INVALID;
    if (!exact_order)
        return le_fail(ctx, LE_ERR_UNSUPPORTED, "HoughLinesP: only cv2's exact visiting order is implemented -- cv2 "
                       "itself recalls under 40 %% of its segments at 1 px under another order, so an order-free "
                       "mode cannot meet the 99 %% criterion (profiles/r2_ppht_order_sensitivity.txt)");
    LE_REQUIRE(ctx, edges && segs && counts, "null pointer");
    LE_REQUIRE(ctx, n > 0 && h > 0 && w > 0 && max_segs > 0, "n, h, w, max_segs must be positive");
    LE_REQUIRE(ctx, h < 65536 && w < 65536, "frame too large");
    LE_REQUIRE(ctx, rho > 0 && theta > 0, "rho and theta must be positive");
    LE_REQUIRE(ctx, max(h, w) + 64 <= PP_MAXT_WORDS * 32, "frame too large for the walk bitmap");
    {
        // 16-bit accumulator cells (biased u16 / short) and magic-number rounding: see le_hough_lines
        const double diag = sqrt((double)h * h + (double)w * w);
        if ((ceil(rho) + 1) * (diag + 2) >= 32767.0 || (diag + 2) / (double)(float)rho >= 4194304.0)
            return le_fail(ctx, LE_ERR_UNSUPPORTED, "rho = %g on a %d x %d frame exceeds the 16-bit vote counters "
                           "or the exact rounding range of the Hough kernels", rho, w, h);
    }
    cudaStream_t st = (cudaStream_t)stream;
    int numangle, numrho, stride;
    le_hough_dims(h, w, rho, theta, &numangle, &numrho);
    int rc = le_hough_upload_trig(ctx, 1, h, w, rho, theta, numangle, &stride, st);
    if (rc) return rc;
    const size_t npx = (size_t)h * w;
    const int nseg = (int)((npx + NZ_SEG - 1) / NZ_SEG);
    if ((rc = le_ensure(ctx, &ctx->counters, sizeof(int) * LE_C_STRIDE * (size_t)n))) return rc;
    if ((rc = le_ensure(ctx, &ctx->ppht_mask, npx * n))) return rc;
    if ((rc = le_ensure(ctx, &ctx->ppht_nz, sizeof(u32) * npx * n))) return rc;
    if ((rc = le_ensure(ctx, &ctx->ppht_accum, sizeof(u16) * (size_t)numangle * stride * n))) return rc;
    if ((rc = le_ensure(ctx, &ctx->misc, sizeof(int) * (size_t)nseg * n + 256))) return rc;
    ctx->edge_list_for = nullptr;  // AUX counters are ours now; the edge lists stay valid but keep it simple
    int *segcnt = (int *)((char *)ctx->misc.p + 256);
    LE_T0(ctx, LE_K_PPHT_PREP, st);
    if ((npx & 15) == 0 && ((size_t)edges & 15) == 0) {  // 16 pixels per thread
        k_nz_count16<<<dim3(nseg, n), 256, 0, st>>>(edges, segcnt, npx, nseg);
        LE_LAUNCH_CHECK(ctx);
        k_nz_write16<<<dim3(nseg, n), 256, 0, st>>>(edges, segcnt, (u32 *)ctx->ppht_nz.p, (u8 *)ctx->ppht_mask.p,
                                                    (int *)ctx->counters.p, npx, nseg, w);
    } else {
        k_nz_count<<<dim3(nseg, n), 256, 0, st>>>(edges, segcnt, npx, nseg);
        LE_LAUNCH_CHECK(ctx);
        k_nz_write<<<dim3(nseg, n), 256, 0, st>>>(edges, segcnt, (u32 *)ctx->ppht_nz.p, (u8 *)ctx->ppht_mask.p,
                                                  (int *)ctx->counters.p, npx, nseg, w);
    }
    LE_LAUNCH_CHECK(ctx);
    // rounds of 64 picks over a precomputed visiting order (le_houghp_rounds.cu); LE_PPHT_ROUNDS=0 keeps the
    // round-1 kernel (also used when a trace is requested); same results from all three
    if (numangle <= 192 && !LE_KNOB_SET("LE_PPHT_GENERIC") && LE_KNOB_INT("LE_PPHT_ROUNDS", 1) && !g_ppht_trace)
        return le_ppht_rounds_launch(ctx, n, h, w, numangle, stride, threshold, min_line_length, max_line_gap, segs,
                                     counts, max_segs, st);
    if (numangle <= 192 && !LE_KNOB_SET("LE_PPHT_GENERIC"))  // batched kernel (le_houghp_fast.cu); same results
        return le_ppht_fast_launch(ctx, n, h, w, numangle, stride, threshold, min_line_length, max_line_gap, segs,
                                   counts, max_segs, g_ppht_trace, g_ppht_trace_cap, st);
    LE_CUDA(ctx, cudaMemsetAsync(ctx->ppht_accum.p, 0, sizeof(u16) * (size_t)numangle * stride * n, st));
    LE_T1(ctx, LE_K_PPHT_PREP, st);
    static le_dev_once attr;
    if (attr.needs(ctx->device, sizeof(PPShared))) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_ppht, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PPShared)));
        attr.done(ctx->device, sizeof(PPShared));
    }
    LE_T0(ctx, LE_K_PPHT, st);
    k_ppht<<<n, PP_THREADS, sizeof(PPShared), st>>>((u8 *)ctx->ppht_mask.p, (u32 *)ctx->ppht_nz.p,
                                                    (short *)ctx->ppht_accum.p, (const float *)ctx->trig.p,
                                                    (int *)ctx->counters.p, segs, counts, max_segs, h, w, numangle,
                                                    stride, threshold, min_line_length, max_line_gap, g_ppht_trace, g_ppht_trace_cap);
    LE_T1(ctx, LE_K_PPHT, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
