// le_hough.cu -- standard Hough transform (a5): vote, accumulator write, peaks, ordered output.
//
// Design (B200): the accumulator row of one angle only has support on the rho interval the image
// rectangle projects to (<= image diagonal + a few bins, although cv2 allocates 2(w+h)+1 bins).  A CTA
// owns one frame and a stripe of A angles (A-2 output rows + 1 halo row each side for the peak
// test) and keeps that support as packed u16 counters in shared memory (A x stride x 2 B, two angle
// rows per word, see hist_get).  Voting maps LANES to ANGLES: a pixel's coordinates are warp-uniform
// (one LDS.128 brings four half2-staged pixels, two float2 ones above 2048 x 2048), a lane's increment is a constant and its address one
// multiply-add on the bits of the magic-number-rounded float.  The kernel is bound by the SM's LSU data
// pipe (shared atomics + gathers + stores), so the 63 % of the int32 accumulator that lies outside
// the supports is written by the copy engine instead: bulk stores (cp.async.bulk, UBLKCP) from a
// zeroed shared buffer, issued before the vote loop.  Inside the support a lane gathers four consecutive
// words of its column (= four bins of two rows) and stores 16 aligned bytes to each of the two rows (the
// origin of every shared row is shifted by up to 3 bins so that bin 4k lands on a 16-byte boundary of the
// global row).  Peaks are tested straight from shared memory, so the accumulator is never re-read.
// CTAs have 512 threads and three of them share an SM: one CTA's votes run beside the write-out of the others.
#include "le_common.cuh"
#include <vector>
#include <stdlib.h>
#include <cuda_fp16.h>

static void hough_dims_host(int h, int w, double rho, double theta, int *numangle, int *numrho)
{
    // cv2 takes rho/theta as double but HoughLinesStandard receives them as float
    float thetaf = (float)theta, rhof = (float)rho;
    double min_theta = 0.0, max_theta = M_PI;
    int na = (int)floor((max_theta - min_theta) / (double)thetaf) + 1;
    if (na > 1 && fabs(M_PI - (na - 1) * (double)thetaf) < (double)thetaf / 2) --na;
    *numangle = na;
    *numrho = (int)lrint(((w + h) * 2 + 1) / (double)rhof);
}

extern "C" int le_hough_dims(int h, int w, double rho, double theta, int *numangle, int *numrho)
{
    if (!numangle || !numrho || h <= 0 || w <= 0 || !(rho > 0) || !(theta > 0)) return LE_ERR_INVALID;
    hough_dims_host(h, w, rho, theta, numangle, numrho);
    return LE_OK;
}

// trig layout in ctx->trig: float cos[na], float sin[na], int rlo[na]
// kind 0: HoughLinesStandard table (angle accumulates in float32)
// kind 1: HoughLinesProbabilistic table (cos(double(n)*theta))
// stride_out: bins per row = support + 4 (room for the alignment shift), multiple of 4
int le_hough_upload_trig(le_ctx *ctx, int kind, int h, int w, double rho, double theta, int numangle, int *stride_out,
                         cudaStream_t st)
{
    float rhof = (float)rho, thetaf = (float)theta;
    float irho = 1 / rhof;
    std::vector<float> tc(numangle), ts(numangle);
    std::vector<int> rlo(numangle);
    if (kind == 0) {
        float ang = 0.f;
        for (int n = 0; n < numangle; ang += thetaf, n++) {
            ts[n] = (float)(sin((double)ang) * irho);
            tc[n] = (float)(cos((double)ang) * irho);
        }
    } else {
        for (int n = 0; n < numangle; n++) {
            tc[n] = (float)(cos((double)n * thetaf) * irho);
            ts[n] = (float)(sin((double)n * thetaf) * irho);
        }
    }
    int span = 0;
    for (int n = 0; n < numangle; n++) {
        double xc = (double)(w - 1) * tc[n], ys = (double)(h - 1) * ts[n];
        double lo = fmin(0.0, xc) + fmin(0.0, ys), hi = fmax(0.0, xc) + fmax(0.0, ys);
        rlo[n] = (int)floor(lo) - 1;
        int rhi = (int)ceil(hi) + 1;
        span = rhi - rlo[n] + 1 > span ? rhi - rlo[n] + 1 : span;
    }
    *stride_out = (span + 4 + 3) & ~3;
    if (ctx->trig_kind == kind && ctx->trig_rho == rho && ctx->trig_theta == theta && ctx->trig_h == h &&
        ctx->trig_w == w)
        return LE_OK;
    int rc = le_ensure(ctx, &ctx->trig, sizeof(float) * 3 * (size_t)numangle);
    if (rc) return rc;
    // pageable-source async copies are staged by the runtime before returning
    LE_CUDA(ctx, cudaMemcpyAsync(ctx->trig.p, tc.data(), sizeof(float) * numangle, cudaMemcpyHostToDevice, st));
    LE_CUDA(ctx, cudaMemcpyAsync((float *)ctx->trig.p + numangle, ts.data(), sizeof(float) * numangle,
                                 cudaMemcpyHostToDevice, st));
    LE_CUDA(ctx, cudaMemcpyAsync((float *)ctx->trig.p + 2 * numangle, rlo.data(), sizeof(int) * numangle,
                                 cudaMemcpyHostToDevice, st));
    ctx->trig_kind = kind;
    ctx->trig_rho = rho;
    ctx->trig_theta = theta;
    ctx->trig_h = h;
    ctx->trig_w = w;
    return LE_OK;
}

// CTA shape of the vote kernel (overridable for development builds, scripts/vote_shape.sh)
// Measured at 256 x 1080p (stripe histogram 70.7 KB): 1024 threads x 2 CTAs per SM 0.396 ms, 512 x 2 0.412,
// 768 x 2 0.392, 384 x 3 0.399, 448 x 3 0.389, 512 x 3 0.373, 576 x 3 0.376, 608 x 3 0.369, 640 x 3 0.372, 512 x 4
// with 8-angle stripes 0.383: what counts is the third resident CTA (its vote phase overlaps the write-out of the
// other two); 512 x 3 leaves 2.8 KB of shared-memory slack.
#ifndef HV_THREADS
#define HV_THREADS 512
#endif
#ifndef HV_ZERO_WORDS
#define HV_ZERO_WORDS 256
#endif
#ifndef HV_MINB
#define HV_MINB 3  // resident CTAs per SM the stripe height is chosen for
#endif
#ifndef HV_PAIR_WRITE
#define HV_PAIR_WRITE 1  // write-out in which a lane owns both rows of its histogram word (0: one row per lane)
#endif
#define HV_CAND_CAP 8192   // per-frame list of peak candidates on stripe-boundary rows

// Shared histogram of a stripe: u16 counters, two ANGLE ROWS per 32-bit word, bin-major:
//     bin b (relative to the row's origin) of local row a  ->  half (a & 1) of hist[b * (A/2) + (a >> 1)].
// A vote instruction has lane = angle row, so its increment (1 or 0x10000) is a per-lane constant and its
// address is ONE multiply-add on the rounded float's bits -- no parity extraction, no shifts.  The two rows
// of a word never carry into each other (a bin holds at most a few thousand votes).  Lanes of one
// instruction spread over banks (b % (64/A)) * (A/2) + (a >> 1): measured 3.0 wavefronts per shared atomic
// (A = 16: 8 word columns x 4 lanes into 4 banks each; the maximum over the 8 groups is ~3).
template <int A>
static __device__ __forceinline__ int hist_get(const u32 *hist, int a, int b, int stride)
{
    return (b >= 0 && b < stride) ? (int)((hist[b * (A / 2) + (a >> 1)] >> ((a & 1) * 16)) & 0xffff) : 0;
}

// Order the peaks of a frame by (votes desc, index asc) == key desc and emit (rho, theta): rank sort (keys
// are unique, rank = number of larger keys).  Runs in the LAST vote CTA of the frame to finish, with the
// stripe histogram's shared memory as the key tile.
static __device__ void sort_frame(const u64 *pk, int total, int peak_cap, int numrho, float rho, float theta,
                                  float *__restrict__ lines, int32_t *__restrict__ votes, int32_t *__restrict__ counts,
                                  int max_lines, int f, u64 *tile, int tile_cap)
{
    const int n = min(total, peak_cap);
    if (threadIdx.x == 0) counts[f] = total;
    const int rowlen = numrho + 2;
    auto emit = [&](u64 key, int rank) {
        u32 idx = 0xffffffffu - (u32)(key & 0xffffffffu);
        int v = (int)(key >> 32);
        int nn = (int)(idx / (u32)rowlen) - 1;
        int r = (int)idx - (nn + 1) * rowlen - 1;
        float lr = __fmul_rn(__fsub_rn((float)r, __fmul_rn((float)(numrho - 1), 0.5f)), rho);
        float lt = __fadd_rn(0.f, __fmul_rn((float)nn, theta));
        float *o = lines + ((size_t)f * max_lines + rank) * 2;
        o[0] = lr;
        o[1] = lt;
        if (votes) votes[(size_t)f * max_lines + rank] = v;
    };
    // Many peaks (textured frames: thousands per frame, the rank sort below is quadratic): only ranks below
    // max_lines are emitted, and every key that can have such a rank has at least T votes, T = the largest vote
    // count reached by max_lines keys or more.  A shared histogram of the vote counts finds T, the keys at or above
    // it are compacted into the shared tile and ranked among themselves (all other keys are smaller).
    constexpr int SF_BINS = 4096, SF_SMALL = 2048;
    __shared__ int s_thr, s_m;
    if (n > SF_SMALL && n > max_lines && tile_cap >= SF_BINS / 2) {
        int *hist = (int *)tile;  // SF_BINS ints = SF_BINS / 2 tile entries
        __syncthreads();
        for (int i = threadIdx.x; i < SF_BINS; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[min((int)(__ldcg(pk + i) >> 32), SF_BINS - 1)], 1);
        __syncthreads();
        if (threadIdx.x < 32) {  // lane l owns bins [128 l, 128 l + 128): suffix sums over the lanes, then its own bins
            const int lane = threadIdx.x;
            int mine = 0;
            for (int b = 0; b < SF_BINS / 32; b++) mine += hist[lane * (SF_BINS / 32) + b];
            int above = 0;  // keys in the bins of higher lanes
            for (int l = 31; l >= 0; l--) {  // (every lane takes part in every shuffle)
                const int vl = __shfl_sync(~0u, mine, l);
                if (l > lane) above += vl;
            }
            const bool here = above < max_lines && above + mine >= max_lines;
            const u32 bal = __ballot_sync(~0u, here);
            if (bal == 0 && lane == 0) s_thr = 0;  // fewer than max_lines keys in total: cannot happen here (n > max_lines)
            if (here) {
                int acc = above, T = lane * (SF_BINS / 32);
                for (int b = SF_BINS / 32 - 1; b >= 0; b--) {
                    acc += hist[lane * (SF_BINS / 32) + b];
                    if (acc >= max_lines) { T = lane * (SF_BINS / 32) + b; break; }
                }
                s_thr = T;
            }
            if (lane == 0) s_m = 0;
        }
        __syncthreads();
        const int T = s_thr;
        if (T < SF_BINS - 1) {  // (the last bin holds every larger count: fall through to the full sort)
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const u64 key = __ldcg(pk + i);
                if ((int)(key >> 32) >= T) {
                    const int pos = atomicAdd(&s_m, 1);
                    if (pos < tile_cap) tile[pos] = key;
                }
            }
            __syncthreads();
            const int m = s_m;
            if (m <= tile_cap) {
                for (int i = threadIdx.x; i < m; i += blockDim.x) {
                    const u64 key = tile[i];
                    int rank = 0;
                    for (int j = 0; j < m; j++) rank += tile[j] > key;
                    if (rank < max_lines) emit(key, rank);
                }
                return;
            }
        }
        __syncthreads();
    }
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        int i = i0 + threadIdx.x;
        u64 key = i < n ? __ldcg(pk + i) : 0;
        int rank = 0;
        for (int t0 = 0; t0 < n; t0 += tile_cap) {
            int tn = min(tile_cap, n - t0);
            __syncthreads();
            for (int j = threadIdx.x; j < tn; j += blockDim.x) tile[j] = __ldcg(pk + t0 + j);
            __syncthreads();
            if (i < n)
                for (int j = 0; j < tn; j++) rank += tile[j] > key;
        }
        if (i < n && rank < max_lines) emit(key, rank);
    }
}

// A = angle rows per CTA, 32 % A == 0.  HALO: the stripe carries one extra row on each side for the peak test
// (A - 2 output rows).  Without HALO all A rows are output rows -- no pixel is voted twice.  The first and
// last row of a stripe then miss one neighbour row: their bins that pass every test that can be made from
// shared memory go to a per-frame candidate list, and the last CTA of the frame to finish (which also orders
// the peaks) makes the one missing comparison from the global accumulator (L2) -- a few hundred loads per
// frame.  That needs the accumulator output, so HALO is kept for accum == NULL.
// H16: frames up to 2048 x 2048 stage a pixel as a half2 (integers below 2049 are exact in fp16): one LDS.128
// then carries FOUR pixels to a half-warp instead of two -- half the staging wavefronts on the LSU data pipe.
template <int A, bool HALO, bool H16>
__global__ void __launch_bounds__(HV_THREADS, (A <= 16 ? HV_MINB : 1))
k_hough_vote(const u32 *__restrict__ queue, int *__restrict__ counters, const float *__restrict__ trig,
             int32_t *__restrict__ accum, u64 *__restrict__ peaks, int peak_cap, int h, int w, int numangle,
             int numrho, int stride, int threshold, float rho, float theta, float *__restrict__ lines,
             int32_t *__restrict__ votes, int32_t *__restrict__ counts, int max_lines, u64 *__restrict__ cand)
{
    extern __shared__ __align__(16) u32 hist[];  // [stride][A/2]
    __shared__ int s_rlo[A];
    __shared__ __align__(16) float2 s_stage[HV_THREADS / 32][32];
    __shared__ __align__(16) u32 s_zero[HV_ZERO_WORDS];  // source of the bulk zero stores
    const int f = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = HV_THREADS / 32;
    const int hw = stride >> 1;  // words per row
    const size_t npx = (size_t)h * w;
    const u32 *q = queue + (size_t)f * npx;
    int *cnt = counters + f * LE_C_STRIDE;
    const int nedge = cnt[LE_C_QTAIL];
    constexpr int OUT = HALO ? A - 2 : A;    // output rows per stripe
    constexpr int A0 = HALO ? 1 : 0;         // local index of the first output row
    const int n_first = blockIdx.x * OUT;    // first output angle of this stripe
    const int *rlo_g = (const int *)(trig + 2 * numangle);
    const int rowlen = numrho + 2;
    const int roff = (numrho - 1) / 2 + 1;  // accumulator column of r == 0
    const size_t asz = (size_t)(numangle + 2) * rowlen;
    // element index (mod 4) of the frame's accumulator relative to a 16-byte boundary
    const u32 acc_mis = accum ? (u32)((((size_t)accum >> 2) + (size_t)f * asz) & 3) : 0u;

    {
        uint4 *hz = (uint4 *)hist;
        for (int i = tid; i < A * hw / 4; i += HV_THREADS) hz[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < HV_ZERO_WORDS; i += HV_THREADS) s_zero[i] = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async-proxy reads
    }
    // shared row origin: rlo shifted down so that bin 0 sits on a 16-byte boundary of the global row
    auto row_origin = [&](int n) -> int {
        int rl = rlo_g[n];
        u32 e = acc_mis + (u32)(((size_t)(n + 1) * rowlen + roff + rl) & 3);  // misalignment of bin 0
        return rl - (int)(e & 3);
    };
    if (tid < A) {
        int n = n_first - A0 + tid;
        s_rlo[tid] = (n >= 0 && n < numangle) ? row_origin(n) : 0;
    }
    // lane -> (angle row a, pixel sub-slot)
    const int a = lane % A, sub = lane / A;
    constexpr int PPW = 32 / A;  // pixels per warp step
    const int n_mine = n_first - A0 + a;
    const bool row_ok = n_mine >= 0 && n_mine < numangle;
    float tcos = 0.f, tsin = 0.f;
    int rlo = 0;
    if (row_ok) {
        tcos = trig[n_mine];
        tsin = trig[numangle + n_mine];
        rlo = row_origin(n_mine);
    }
    // Lanes of rows outside [0, numangle) (halo of the first / last stripe) have cos = sin = 0 and vote into
    // bin 0 of their own scratch row, which nobody reads.
    // round-half-even by the magic-number add: bits(v + 1.5 * 2^23) - 0x4B400000 == rint(v) for |v| < 2^22,
    // the same integer __float2int_rn gives.  Byte address of the bin's word = col + (rint(v) - rlo) * 2A:
    // everything but the float's bits is folded into one per-lane constant.
    const float MAGIC = 12582912.f;
    const u32 kaddr = (u32)__cvta_generic_to_shared(hist + (a >> 1)) - (u32)(0x4B400000 + rlo) * (2u * A);
    const u32 inc = (a & 1) ? 0x10000u : 1u;
    float4 *stage = (float4 *)s_stage[warp];  // two staged pixels (x0, y0, x1, y1) per entry
    const float rw = 1.0f / (float)w;
    const bool small = npx <= ((size_t)1 << 24);
    auto vote = [&](float x, float y, bool on) {
        const float t = __fadd_rn(__fadd_rn(__fmul_rn(x, tcos), __fmul_rn(y, tsin)), MAGIC);
        const u32 sa = (u32)__float_as_int(t) * (2u * A) + kaddr;
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(sa), "r"(on ? inc : 0u) : "memory");
    };
    __syncthreads();
    // ---- accumulator rows of the stripe; the zeros outside the supports do not depend on the votes: they are
    // handed to the copy engine BEFORE the vote loop, so 63 % of the stripe's HBM writes run beside the voting
    int32_t *acc_f = accum ? accum + (size_t)f * asz : nullptr;
    u64 *pk = peaks + (size_t)f * peak_cap;
    const int n_last = min(n_first + OUT, numangle);  // exclusive
    // rows to write: accumulator rows n+1 for n in [n_first, n_last); plus pad rows 0 and numangle+1
    const int row_lo = (blockIdx.x == 0) ? -1 : n_first;              // as angle index n (row n+1)
    const int row_hi = (n_last == numangle) ? numangle + 1 : n_last;  // exclusive
    // (1) zeros left and right of each row's support, and the pad rows: one warp per row; the 16-byte
    // aligned body of each span goes out as bulk stores issued by lane 0, the few unaligned ints as STG
    if (acc_f) {
        const u32 zero_sa = (u32)__cvta_generic_to_shared(s_zero);
        bool issued = false;
        for (int n = row_lo + warp; n < row_hi; n += nwarps) {
            const int la = n - (n_first - A0);  // local row index
            const bool real = n >= 0 && n < numangle;
            int32_t *out = acc_f + (size_t)(n + 1) * rowlen;
            const int c_lo = real ? roff + s_rlo[la] : rowlen, c_hi = real ? c_lo + stride : rowlen;
            int z0 = 0, z1 = min(max(c_lo, 0), rowlen);
#pragma unroll 1
            for (int part = 0; part < 2; part++) {
                if (z1 > z0) {
                    const int head = (int)((4 - ((((size_t)(out + z0)) >> 2) & 3)) & 3);
                    const int a0 = min(z0 + head, z1), a1 = a0 + ((z1 - a0) & ~3);
                    if (z0 + lane < a0) out[z0 + lane] = 0;
                    if (lane == 0) {
                        for (int c = a0; c < a1; c += HV_ZERO_WORDS)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c),
                                         "r"(zero_sa), "r"(4 * min(HV_ZERO_WORDS, a1 - c))
                                         : "memory");
                        issued = true;
                    }
                    if (a1 + lane < z1) out[a1 + lane] = 0;
                }
                z0 = min(max(c_hi, 0), rowlen);
                z1 = rowlen;
            }
        }
        if (issued) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    int p_next = warp * 32 + lane < nedge ? (int)q[warp * 32 + lane] : 0;
    for (int base = warp * 32; base < nedge; base += nwarps * 32) {
        // each lane converts one queue entry to float (x, y) and stages it; the warp then votes the
        // 32 pixels, four (half2) or two (float2) per LDS.128 (lanes = angles, so no two lanes share a row).  The entry of the
        // warp's next batch is fetched now, so its L2 round trip runs beside the votes of this one.
        const int p = p_next;
        {
            const int nidx = base + nwarps * 32 + lane;
            p_next = nidx < nedge ? (int)q[nidx] : 0;
        }
        int py = (int)((float)p * rw);  // estimate of p / w, then exact correction
        int px = p - py * w;
        if (small) {  // p < 2^24 is exact in float and the estimate is off by one at most: two selects, no loop
            if (px < 0) { px += w; py--; }
            if (px >= w) { px -= w; py++; }
        } else {
            while (px < 0) { px += w; py--; }
            while (px >= w) { px -= w; py++; }
        }
        const int cntw = min(32, nedge - base);
        if (H16) {
            __syncwarp();
            ((__half2 *)s_stage[warp])[lane] = __floats2half2_rn((float)px, (float)py);
            __syncwarp();
            const uint4 *st4 = (const uint4 *)s_stage[warp];  // four staged pixels per entry
            if (cntw == 32) {
#pragma unroll
                for (int i = 0; i < 8; i += PPW) {
                    const uint4 v = st4[i + sub];
                    const float2 p0 = __half22float2(*(const __half2 *)&v.x), p1 = __half22float2(*(const __half2 *)&v.y);
                    const float2 p2 = __half22float2(*(const __half2 *)&v.z), p3 = __half22float2(*(const __half2 *)&v.w);
                    vote(p0.x, p0.y, true);
                    vote(p1.x, p1.y, true);
                    vote(p2.x, p2.y, true);
                    vote(p3.x, p3.y, true);
                }
            } else {
                for (int i = 0; i < 8; i += PPW) {
                    const uint4 v = st4[i + sub];
                    const float2 p0 = __half22float2(*(const __half2 *)&v.x), p1 = __half22float2(*(const __half2 *)&v.y);
                    const float2 p2 = __half22float2(*(const __half2 *)&v.z), p3 = __half22float2(*(const __half2 *)&v.w);
                    const int k0 = 4 * (i + sub);
                    vote(p0.x, p0.y, k0 < cntw);
                    vote(p1.x, p1.y, k0 + 1 < cntw);
                    vote(p2.x, p2.y, k0 + 2 < cntw);
                    vote(p3.x, p3.y, k0 + 3 < cntw);
                }
            }
            continue;
        }
        __syncwarp();
        s_stage[warp][lane] = make_float2((float)px, (float)py);
        __syncwarp();
        if (cntw == 32) {
#pragma unroll 4
            for (int i = 0; i < 16; i += PPW) {
                const float4 xy = stage[i + sub];
                vote(xy.x, xy.y, true);
                vote(xy.z, xy.w, true);
            }
        } else {
            for (int i = 0; i < 16; i += PPW) {
                const float4 xy = stage[i + sub];
                vote(xy.x, xy.y, 2 * (i + sub) < cntw);
                vote(xy.z, xy.w, 2 * (i + sub) + 1 < cntw);
            }
        }
    }
    __syncthreads();

    // ---- write the supports of the stripe's accumulator rows + peak test
#if HV_PAIR_WRITE
    // (2) the supports: lane = (word column kc, bin group g).  A histogram word holds the same bin of TWO rows, so
    // a lane gathers four consecutive bins of its column once (bank-rotated: the groups of a column read the four
    // bins in different order) and owns both rows: two 16-byte stores, each to its own global row (every row's
    // origin was shifted so that the store is aligned).  Against one row per lane: half the gathers, and a store
    // instruction covers A/2 rows in runs of 64 bytes instead of A rows in runs of 32.  Peaks are tested from
    // shared memory.
    {
        constexpr int NC = A / 2, G = 32 / NC;
        const int kc = lane % NC, gq = lane / NC;
        const u32 *col = hist + kc;
        int ar[2], nr[2], cl[2];
        bool orow[2], there[2];
        int32_t *outp[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            ar[r] = 2 * kc + r;
            nr[r] = n_first - A0 + ar[r];
            const bool ok = nr[r] >= 0 && nr[r] < numangle;
            orow[r] = ar[r] >= A0 && ar[r] < A0 + OUT && nr[r] < n_last && ok;
            // rows whose two neighbour rows are in this stripe (or do not exist) are tested here
            there[r] = (ar[r] >= 1 || nr[r] == 0) && (ar[r] <= A - 2 || nr[r] == numangle - 1);
            cl[r] = roff + s_rlo[ar[r]];
            outp[r] = (acc_f && orow[r]) ? acc_f + (size_t)(nr[r] + 1) * rowlen + cl[r] : nullptr;
        }
        if (orow[0] || orow[1]) {
            for (int b0 = 4 * (warp * G + gq); b0 < stride; b0 += 4 * G * nwarps) {
                u32 wq[4], wd[4];
#pragma unroll
                for (int k = 0; k < 4; k++) wq[k] = col[(b0 + ((k + gq) & 3)) * NC];  // wq[k] = bin b0 + ((k + gq) & 3)
#pragma unroll
                for (int m = 0; m < 4; m++) {  // wd[m] = bin b0 + m = wq[(m - gq) & 3]
                    const u32 t0 = (gq & 1) ? wq[(m + 3) & 3] : wq[m];
                    const u32 t2 = (gq & 1) ? wq[(m + 1) & 3] : wq[(m + 2) & 3];
                    wd[m] = (gq & 2) ? t2 : t0;
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    if (!orow[r]) continue;
                    int v[4];
#pragma unroll
                    for (int m = 0; m < 4; m++) v[m] = (int)((wd[m] >> (16 * r)) & 0xffff);
                    const int c = cl[r] + b0;
                    if (outp[r]) {
                        if (c >= 0 && c + 3 < rowlen)
                            *(int4 *)(outp[r] + b0) = make_int4(v[0], v[1], v[2], v[3]);
                        else {
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                if (c + k >= 0 && c + k < rowlen) outp[r][b0 + k] = v[k];
                        }
                    }
                    if (max(max(v[0], v[1]), max(v[2], v[3])) <= threshold) continue;  // nothing above the threshold
                    const int a_r = ar[r], n_r = nr[r];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int cc = c + k, b = b0 + k;
                        if (v[k] > threshold && cc >= 1 && cc <= numrho) {
                            int left = hist_get<A>(hist, a_r, b - 1, stride), right = hist_get<A>(hist, a_r, b + 1, stride);
                            int up = 0, down = 0;  // a neighbour row of another stripe counts as 0 here
                            if (n_r - 1 >= 0 && a_r >= 1)
                                up = hist_get<A>(hist, a_r - 1, cc - roff - s_rlo[a_r - 1], stride);
                            if (n_r + 1 < numangle && a_r <= A - 2)
                                down = hist_get<A>(hist, a_r + 1, cc - roff - s_rlo[a_r + 1], stride);
                            if (v[k] > left && v[k] >= right && v[k] > up && v[k] >= down) {
                                const u32 idx = (u32)((n_r + 1) * rowlen + cc);
                                const u64 key = ((u64)(u32)v[k] << 32) | (u64)(0xffffffffu - idx);
                                if (there[r]) {
                                    int pos = atomicAdd(&cnt[LE_C_PEAKS], 1);
                                    if (pos < peak_cap) pk[pos] = key;
                                    else cnt[LE_C_OVERFLOW] = 4;  // peak list full: le_check reports it
                                } else {
                                    int pos = atomicAdd(&cnt[LE_C_AUX1], 1);
                                    if (pos < HV_CAND_CAP) cand[(size_t)f * HV_CAND_CAP + pos] = key;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
#else
    // (2) the supports: lane = (row a, slot sub); a warp step covers 4 * PPW bins of every row of the
    // stripe -- each lane gathers four consecutive bins of ITS row and stores 16 bytes to ITS global row
    // (the row origin was shifted so that this store is aligned); peaks are tested from shared memory
    {
        const bool out_row = a >= A0 && a < A0 + OUT && n_mine < n_last && row_ok;
        // rows whose two neighbour rows are in this stripe (or do not exist) are tested here
        const bool test_here = (a >= 1 || n_mine == 0) && (a <= A - 2 || n_mine == numangle - 1);
        const int c_lo = roff + rlo;
        int32_t *out = (acc_f && out_row) ? acc_f + (size_t)(n_mine + 1) * rowlen + c_lo : nullptr;
        const int rlo_up = s_rlo[max(a - 1, 0)], rlo_dn = s_rlo[min(a + 1, A - 1)];
        const u32 *col = hist + (a >> 1);
        const int sh = (a & 1) * 16;
        if (out_row) {
            for (int b0 = 4 * (warp * PPW + sub); b0 < stride; b0 += 4 * PPW * nwarps) {
                int g[4], v[4];  // (the slots of a row read the four bins in rotated order: distinct banks)
#pragma unroll
                for (int k = 0; k < 4; k++) g[k] = (int)((col[(b0 + ((k + 2 * sub) & 3)) * (A / 2)] >> sh) & 0xffff);
#pragma unroll
                for (int m = 0; m < 4; m++) v[m] = (sub & 1) ? g[(m + 2) & 3] : g[m];
                const int c = c_lo + b0;
                if (out) {
                    if (c >= 0 && c + 3 < rowlen)
                        *(int4 *)(out + b0) = make_int4(v[0], v[1], v[2], v[3]);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (c + k >= 0 && c + k < rowlen) out[b0 + k] = v[k];
                    }
                }
                if (max(max(v[0], v[1]), max(v[2], v[3])) <= threshold) continue;  // nothing above the threshold
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int cc = c + k, b = b0 + k;
                    if (v[k] > threshold && cc >= 1 && cc <= numrho) {
                        int left = hist_get<A>(hist, a, b - 1, stride), right = hist_get<A>(hist, a, b + 1, stride);
                        int up = 0, down = 0;  // a neighbour row of another stripe counts as 0 here
                        if (n_mine - 1 >= 0 && a >= 1) up = hist_get<A>(hist, a - 1, cc - roff - rlo_up, stride);
                        if (n_mine + 1 < numangle && a <= A - 2) down = hist_get<A>(hist, a + 1, cc - roff - rlo_dn, stride);
                        if (v[k] > left && v[k] >= right && v[k] > up && v[k] >= down) {
                            const u32 idx = (u32)((n_mine + 1) * rowlen + cc);
                            const u64 key = ((u64)(u32)v[k] << 32) | (u64)(0xffffffffu - idx);
                            if (test_here) {
                                int pos = atomicAdd(&cnt[LE_C_PEAKS], 1);
                                if (pos < peak_cap) pk[pos] = key;
                                else cnt[LE_C_OVERFLOW] = 4;
                            } else {
                                int pos = atomicAdd(&cnt[LE_C_AUX1], 1);
                                if (pos < HV_CAND_CAP) cand[(size_t)f * HV_CAND_CAP + pos] = key;
                            }
                        }
                    }
                }
            }
        }
    }
#endif
    __shared__ int s_last;
    // ---- the last CTA of the frame to get here orders the frame's peaks (its histogram is free by then)
    // The CTA's rows and peak keys must be visible before it signals: the barrier orders every thread's writes
    // before thread 0's fence, which is cumulative (the grid-sync pattern).  One fence per CTA instead of one per
    // warp: MEMBAR.SC + CCTL.IVALL in all 32 warps held 12 % of the kernel's stall samples.
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&cnt[LE_C_AUX0], 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (!HALO) {
            // the one comparison the first / last row of every stripe could not make: its neighbour row in
            // the adjacent stripe, read from the (now complete) global accumulator; zero outside that row's
            // support, which is never read (those zeros may still be in flight from the copy engine)
            auto at = [&](int r, int cc) -> int {
                const int b = cc - roff - row_origin(r);
                return (b >= 0 && b < stride) ? __ldcg(acc_f + (size_t)(r + 1) * rowlen + cc) : 0;
            };
            auto settle = [&](int r, int cc, int v) {
                const bool first_row = r % OUT == 0;  // misses the row above; otherwise the row below
                const bool keep = first_row ? v > at(r - 1, cc) : v >= at(r + 1, cc);
                if (keep) {
                    int pos = atomicAdd(&cnt[LE_C_PEAKS], 1);
                    u32 idx = (u32)((r + 1) * rowlen + cc);
                    if (pos < peak_cap) pk[pos] = ((u64)(u32)v << 32) | (u64)(0xffffffffu - idx);
                    else cnt[LE_C_OVERFLOW] = 4;
                }
            };
            const int ncand = *(volatile int *)&cnt[LE_C_AUX1];
            if (ncand <= HV_CAND_CAP) {
                for (int i = tid; i < ncand; i += HV_THREADS) {
                    const u64 key = __ldcg(cand + (size_t)f * HV_CAND_CAP + i);
                    const u32 idx = 0xffffffffu - (u32)(key & 0xffffffffu);
                    const int r = (int)(idx / (u32)rowlen) - 1;
                    settle(r, (int)(idx - (u32)(r + 1) * rowlen), (int)(key >> 32));
                }
            } else {
                // list overflow (never on real frames): redo the boundary rows from the global accumulator
                for (int r = 0; r < numangle; r++) {
                    const bool first_row = r % OUT == 0 && r > 0, last_row = r % OUT == OUT - 1 && r + 1 < numangle;
                    if (!first_row && !last_row) continue;
                    const int o_r = row_origin(r);
                    for (int b = tid; b < stride; b += HV_THREADS) {
                        const int cc = roff + o_r + b;
                        if (cc < 1 || cc > numrho) continue;
                        const int v = at(r, cc);
                        if (v <= threshold) continue;
                        const int left = b > 0 ? at(r, cc - 1) : 0, right = b + 1 < stride ? at(r, cc + 1) : 0;
                        const int other = first_row ? (r + 1 < numangle ? at(r + 1, cc) : 0) : at(r - 1, cc);
                        if (v > left && v >= right && (first_row ? v >= other : v > other)) settle(r, cc, v);
                    }
                }
            }
            __threadfence();
            __syncthreads();
        }
        const int total = *(volatile int *)&cnt[LE_C_PEAKS];
        sort_frame(pk, total, peak_cap, numrho, rho, theta, lines, votes, counts, max_lines, f, (u64 *)hist,
                   A * stride / 4);
    }
    // the bulk zero stores read s_zero: they must be complete before the CTA (and its shared memory) goes
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int A, bool HALO, bool H16>
static int launch_vote(le_ctx *ctx, int n, int h, int w, int numangle, int numrho, int stride, int threshold,
                       int32_t *accum, int peak_cap, float rho, float theta, float *lines, int32_t *votes,
                       int32_t *counts, int max_lines, cudaStream_t st)
{
    size_t smem = (size_t)A * stride * 2;
    static le_dev_once attr;
    if (attr.needs(ctx->device, smem)) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_hough_vote<A, HALO, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.done(ctx->device, smem);
    }
    constexpr int OUT = HALO ? A - 2 : A;
    int stripes = (numangle + OUT - 1) / OUT;
    LE_T0(ctx, LE_K_VOTE, st);
    k_hough_vote<A, HALO, H16><<<dim3(stripes, n), HV_THREADS, smem, st>>>(
        (const u32 *)ctx->queue.p, (int *)ctx->counters.p, (const float *)ctx->trig.p, accum, (u64 *)ctx->peaks.p,
        peak_cap, h, w, numangle, numrho, stride, threshold, rho, theta, lines, votes, counts, max_lines,
        (u64 *)ctx->misc.p);
    LE_T1(ctx, LE_K_VOTE, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

__global__ void k_zero_peaks(int *counters, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        counters[i * LE_C_STRIDE + LE_C_PEAKS] = 0;
        counters[i * LE_C_STRIDE + LE_C_AUX0] = 0;  // vote CTAs of the frame that have finished
        counters[i * LE_C_STRIDE + LE_C_AUX1] = 0;  // peak candidates on stripe-boundary rows
        counters[i * LE_C_STRIDE + LE_C_OVERFLOW] &= ~4;  // "peak list full" belongs to one call
    }
}

extern "C" int le_hough_lines(le_ctx *ctx, const uint8_t *edges, int n, int h, int w, double rho, double theta,
                              int threshold, int32_t *accum, float *lines, int32_t *votes, int32_t *counts,
                              int max_lines, void *stream)
{
    LE_ON_DEVICE(ctx);
    if (!ctx) return LE_ERR_INVALID;
    LE_REQUIRE(ctx, n > 0 && h > 0 && w > 0, "n, h, w must be positive");
    LE_REQUIRE(ctx, lines && counts && max_lines > 0, "lines/counts must be non-null, max_lines positive");
    LE_REQUIRE(ctx, rho > 0 && theta > 0, "rho and theta must be positive");
    LE_REQUIRE(ctx, (size_t)h * w < ((size_t)1 << 30), "frame too large");
    LE_REQUIRE(ctx, !accum || (((size_t)accum) & 3) == 0, "accum must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int numangle, numrho, stride;
    hough_dims_host(h, w, rho, theta, &numangle, &numrho);
    LE_REQUIRE(ctx, numangle >= 1 && numrho >= 1, "degenerate accumulator");
    LE_REQUIRE(ctx, (size_t)(numangle + 2) * (numrho + 2) < ((size_t)1 << 31), "accumulator too large");
    {
        // The stripe histogram counts in u16 and rounds with a magic-number add (exact for |v| < 2^22).  A bin
        // collects at most the pixels of a strip (rho + 1) wide across the frame's diagonal (cv2 counts in
        // int32 and has no such limit): parameter combinations beyond that are refused, not wrapped.
        const double diag = sqrt((double)h * h + (double)w * w);
        if ((ceil(rho) + 1) * (diag + 2) >= 65535.0 || (diag + 2) / (double)(float)rho >= 4194304.0)
            return le_fail(ctx, LE_ERR_UNSUPPORTED, "rho = %g on a %d x %d frame exceeds the 16-bit vote counters "
                           "or the exact rounding range of the Hough kernels", rho, w, h);
    }
    int rc = le_edge_list_from_map(ctx, edges, n, h, w, st);
    if (rc) return rc;
    rc = le_hough_upload_trig(ctx, 0, h, w, rho, theta, numangle, &stride, st);
    if (rc) return rc;
    size_t cap64 = (size_t)max_lines * 4;
    if (cap64 < 65536) cap64 = 65536;
    if (cap64 > (size_t)numangle * numrho) cap64 = (size_t)numangle * numrho;
    int peak_cap = (int)cap64;
    rc = le_ensure(ctx, &ctx->peaks, sizeof(u64) * (size_t)n * peak_cap);
    if (rc) return rc;
    rc = le_ensure(ctx, &ctx->misc, sizeof(u64) * (size_t)n * HV_CAND_CAP);
    if (rc) return rc;
    k_zero_peaks<<<(n + 255) / 256, 256, 0, st>>>((int *)ctx->counters.p, n);
    LE_LAUNCH_CHECK(ctx);
    size_t budget = (size_t)ctx->smem_optin - 1024 - sizeof(float2) * HV_THREADS - 4 * HV_ZERO_WORDS;
    // angles per CTA: the largest stripe whose counters leave room for HV_MINB resident CTAs per SM (their
    // zero / vote / write phases then overlap); fall back to whatever fits one CTA
    size_t two = budget / HV_MINB - 4096;
    int A = 0;
    for (int cand = 32; cand >= 8 && !A; cand >>= 1)
        if ((size_t)cand * stride * 2 <= two) A = cand;
    for (int cand = 32; cand >= 4 && !A; cand >>= 1)
        if ((size_t)cand * stride * 2 <= budget) A = cand;
    if (const int v = LE_KNOB_INT("LE_VOTE_A", 0)) {  // development knob: force the stripe height
        if ((v == 4 || v == 8 || v == 16 || v == 32) && (size_t)v * stride * 2 <= budget) A = v;
    }
    // without halo rows when the accumulator is written (the deferred boundary test reads it back)
    const bool nohalo = accum != nullptr && A >= 4;
    const bool no_h16 = LE_KNOB_SET("LE_VOTE_F32");  // development knob: float2 staging everywhere
    const bool h16 = h <= 2048 && w <= 2048 && !no_h16;
#define LE_VOTE(AA, HH) (h16 ? launch_vote<AA, HH, true> : launch_vote<AA, HH, false>)(ctx, n, h, w, numangle, numrho, stride, threshold, accum, peak_cap, (float)rho, \
                                            (float)theta, lines, votes, counts, max_lines, st)
    if (A == 32) rc = nohalo ? LE_VOTE(32, false) : LE_VOTE(32, true);
    else if (A == 16) rc = nohalo ? LE_VOTE(16, false) : LE_VOTE(16, true);
    else if (A == 8) rc = nohalo ? LE_VOTE(8, false) : LE_VOTE(8, true);
    else if (A == 4) rc = nohalo ? LE_VOTE(4, false) : LE_VOTE(4, true);
#undef LE_VOTE
    else return le_fail(ctx, LE_ERR_UNSUPPORTED, "image too large for the shared-memory Hough stripes (stride %d)", stride);
    if (rc) return rc;
    return LE_OK;
}
