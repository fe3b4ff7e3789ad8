// le_hough.cu -- standard Hough transform (a5): vote, accumulator write, peaks, ordered output.
//
// Design (B200): the accumulator row of one angle only has support on the rho interval the image
// rectangle projects to (<= image diagonal + 3 bins, although cv2 allocates 2(w+h)+1 bins).  A CTA
// owns one frame and a stripe of A angles (A-2 output rows + 1 halo row each side for the peak
// test) and keeps that support as packed u16 counters in shared memory (A x span x 2 B).  Voting
// maps LANES to ANGLES, so a warp-wide shared atomic never has an intra-warp address collision, and
// a pixel's coordinates are warp-uniform (shuffle broadcast).  Each CTA then streams its rows of
// the int32 accumulator to HBM exactly once (zeros outside the support) with 128-bit stores and
// tests peaks straight from shared memory -- the accumulator is never re-read.
#include "le_common.cuh"
#include <vector>

struct HoughGeom {
    int numangle, numrho, stride;  // stride: u16 bins per smem row (even)
};

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
    *stride_out = (span + 1) & ~1;
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

#define HV_THREADS 512

static __device__ __forceinline__ int hist_get(const u32 *row, int b) { return (row[b >> 1] >> ((b & 1) * 16)) & 0xffff; }

// A = angles per CTA including the two halo rows; 32 % A == 0.
template <int A>
__global__ void __launch_bounds__(HV_THREADS)
k_hough_vote(const u32 *__restrict__ queue, int *__restrict__ counters, const float *__restrict__ trig,
             int32_t *__restrict__ accum, u64 *__restrict__ peaks, int peak_cap, int h, int w, int numangle,
             int numrho, int stride, int threshold)
{
    extern __shared__ __align__(16) u32 hist[];  // [A][stride/2]
    __shared__ int s_rlo[A];
    const int f = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = HV_THREADS / 32;
    const int hw = stride >> 1;  // words per row
    const size_t npx = (size_t)h * w;
    const u32 *q = queue + (size_t)f * npx;
    int *cnt = counters + f * LE_C_STRIDE;
    const int nedge = cnt[LE_C_QTAIL];
    const int n_first = blockIdx.x * (A - 2);  // first output angle of this stripe
    const int *rlo_g = (const int *)(trig + 2 * numangle);

    for (int i = tid; i < A * hw; i += HV_THREADS) hist[i] = 0;
    if (tid < A) {
        int n = n_first - 1 + tid;
        s_rlo[tid] = (n >= 0 && n < numangle) ? rlo_g[n] : 0;
    }
    // lane -> (angle row a, pixel sub-slot)
    const int a = lane % A, sub = lane / A;
    constexpr int PPW = 32 / A;  // pixels per warp step
    const int n_mine = n_first - 1 + a;
    const bool row_ok = n_mine >= 0 && n_mine < numangle;
    float tcos = 0.f, tsin = 0.f;
    int rlo = 0;
    if (row_ok) {
        tcos = trig[n_mine];
        tsin = trig[numangle + n_mine];
        rlo = rlo_g[n_mine];
    }
    u32 *myrow = hist + a * hw;
    __syncthreads();
    for (int base = warp * 32; base < nedge; base += nwarps * 32) {
        int idx = base + lane;
        int p = idx < nedge ? (int)q[idx] : -1;
        int py = p / w, px = p - py * w;
        int cntw = min(32, nedge - base);
#pragma unroll 4
        for (int j = 0; j < cntw; j += PPW) {
            int src = j + sub;
            int x = __shfl_sync(~0u, px, src & 31), y = __shfl_sync(~0u, py, src & 31);
            if (row_ok && src < cntw) {
                float fr = __fadd_rn(__fmul_rn((float)x, tcos), __fmul_rn((float)y, tsin));
                int b = __float2int_rn(fr) - rlo;
                atomicAdd(&myrow[b >> 1], (b & 1) ? 0x10000u : 1u);
            }
        }
    }
    __syncthreads();

    // ---- write the stripe's accumulator rows + peak test
    const int rowlen = numrho + 2;
    const int roff = (numrho - 1) / 2 + 1;  // column of r == 0
    const size_t asz = (size_t)(numangle + 2) * rowlen;
    int32_t *acc_f = accum ? accum + (size_t)f * asz : nullptr;
    u64 *pk = peaks + (size_t)f * peak_cap;
    const int n_last = min(n_first + (A - 2), numangle);  // exclusive
    // rows to write: accumulator rows n+1 for n in [n_first, n_last); plus pad rows 0 and numangle+1
    const int row_lo = (blockIdx.x == 0) ? -1 : n_first;                    // as angle index n (row n+1)
    const int row_hi = (n_last == numangle) ? numangle + 1 : n_last;        // exclusive
    for (int n = row_lo + warp; n < row_hi; n += nwarps) {
        const int la = n - (n_first - 1);  // local row index
        const bool real = n >= 0 && n < numangle;
        const u32 *row = hist + la * hw;
        const int rl = real ? s_rlo[la] : 0;
        int32_t *out = acc_f ? acc_f + (size_t)(n + 1) * rowlen : nullptr;
        // align to 16 B
        int head = 0;
        if (out) head = (int)((4 - (((size_t)out >> 2) & 3)) & 3);
        // columns [0, rowlen): value = hist(col - roff - rl) if real and in [0, stride)
        for (int c0 = -((4 - head) & 3) + lane * 4; c0 < rowlen; c0 += 128) {
            // chunk covers columns c0 .. c0+3 (c0 may be negative for the head chunk)
            int v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int c = c0 + k;
                int b = c - roff - rl;
                v[k] = (real && c >= 0 && c < rowlen && b >= 0 && b < stride) ? hist_get(row, b) : 0;
            }
            if (out) {
                if (c0 >= 0 && c0 + 3 < rowlen) *(int4 *)(out + c0) = make_int4(v[0], v[1], v[2], v[3]);
                else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (c0 + k >= 0 && c0 + k < rowlen) out[c0 + k] = v[k];
                }
            }
            if (real) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int c = c0 + k;
                    if (v[k] > threshold && c >= 1 && c <= numrho) {
                        int b = c - roff - rl;
                        int left = (b - 1 >= 0) ? hist_get(row, b - 1) : 0;
                        int right = (b + 1 < stride) ? hist_get(row, b + 1) : 0;
                        int up = 0, down = 0;
                        if (n - 1 >= 0) {
                            int bu = c - roff - s_rlo[la - 1];
                            if (bu >= 0 && bu < stride) up = hist_get(row - hw, bu);
                        }
                        if (n + 1 < numangle) {
                            int bd = c - roff - s_rlo[la + 1];
                            if (bd >= 0 && bd < stride) down = hist_get(row + hw, bd);
                        }
                        if (v[k] > left && v[k] >= right && v[k] > up && v[k] >= down) {
                            int pos = atomicAdd(&cnt[LE_C_PEAKS], 1);
                            u32 idx = (u32)((n + 1) * rowlen + c);
                            if (pos < peak_cap) pk[pos] = ((u64)(u32)v[k] << 32) | (u64)(0xffffffffu - idx);
                        }
                    }
                }
            }
        }
    }
}

// Order the peaks of each frame by (votes desc, index asc) == key desc, emit (rho, theta).
// Rank sort: keys are unique, rank = number of larger keys.
#define HS_THREADS 512
#define HS_TILE 2048
__global__ void __launch_bounds__(HS_THREADS)
k_hough_sort(const u64 *__restrict__ peaks, const int *__restrict__ counters, int peak_cap, int numrho, float rho,
             float theta, float *__restrict__ lines, int32_t *__restrict__ votes, int32_t *__restrict__ counts,
             int max_lines)
{
    __shared__ u64 tile[HS_TILE];
    const int f = blockIdx.x;
    const u64 *pk = peaks + (size_t)f * peak_cap;
    const int total = counters[f * LE_C_STRIDE + LE_C_PEAKS];
    const int n = min(total, peak_cap);
    if (threadIdx.x == 0) counts[f] = total;
    const int rowlen = numrho + 2;
    for (int i0 = 0; i0 < n; i0 += HS_THREADS) {
        int i = i0 + threadIdx.x;
        u64 key = i < n ? pk[i] : 0;
        int rank = 0;
        for (int t0 = 0; t0 < n; t0 += HS_TILE) {
            int tn = min(HS_TILE, n - t0);
            __syncthreads();
            for (int j = threadIdx.x; j < tn; j += HS_THREADS) tile[j] = pk[t0 + j];
            __syncthreads();
            if (i < n)
                for (int j = 0; j < tn; j++) rank += tile[j] > key;
        }
        if (i < n && rank < max_lines) {
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
        }
    }
}

template <int A>
static int launch_vote(le_ctx *ctx, int n, int h, int w, int numangle, int numrho, int stride, int threshold,
                       int32_t *accum, int peak_cap, cudaStream_t st)
{
    size_t smem = (size_t)A * stride * 2;
    static size_t attr = 0;
    if (smem > attr) {
        LE_CUDA(ctx, cudaFuncSetAttribute(k_hough_vote<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    int stripes = (numangle + (A - 2) - 1) / (A - 2);
    LE_T0(ctx, LE_K_VOTE, st);
    k_hough_vote<A><<<dim3(stripes, n), HV_THREADS, smem, st>>>(
        (const u32 *)ctx->queue.p, (int *)ctx->counters.p, (const float *)ctx->trig.p, accum, (u64 *)ctx->peaks.p,
        peak_cap, h, w, numangle, numrho, stride, threshold);
    LE_T1(ctx, LE_K_VOTE, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

__global__ void k_zero_peaks(int *counters, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counters[i * LE_C_STRIDE + LE_C_PEAKS] = 0;
}

extern "C" int le_hough_lines(le_ctx *ctx, const uint8_t *edges, int n, int h, int w, double rho, double theta,
                              int threshold, int32_t *accum, float *lines, int32_t *votes, int32_t *counts,
                              int max_lines, void *stream)
{
    if (!ctx) return LE_ERR_INVALID;
    LE_REQUIRE(ctx, n > 0 && h > 0 && w > 0, "n, h, w must be positive");
    LE_REQUIRE(ctx, lines && counts && max_lines > 0, "lines/counts must be non-null, max_lines positive");
    LE_REQUIRE(ctx, rho > 0 && theta > 0, "rho and theta must be positive");
    LE_REQUIRE(ctx, (size_t)h * w < ((size_t)1 << 30), "frame too large");
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    int numangle, numrho, stride;
    hough_dims_host(h, w, rho, theta, &numangle, &numrho);
    LE_REQUIRE(ctx, numangle >= 1 && numrho >= 1, "degenerate accumulator");
    LE_REQUIRE(ctx, (size_t)(numangle + 2) * (numrho + 2) < ((size_t)1 << 31), "accumulator too large");
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
    k_zero_peaks<<<(n + 255) / 256, 256, 0, st>>>((int *)ctx->counters.p, n);
    LE_LAUNCH_CHECK(ctx);
    size_t budget = (size_t)ctx->smem_optin - 1024;
    if ((size_t)32 * stride * 2 <= budget) rc = launch_vote<32>(ctx, n, h, w, numangle, numrho, stride, threshold, accum, peak_cap, st);
    else if ((size_t)16 * stride * 2 <= budget) rc = launch_vote<16>(ctx, n, h, w, numangle, numrho, stride, threshold, accum, peak_cap, st);
    else if ((size_t)8 * stride * 2 <= budget) rc = launch_vote<8>(ctx, n, h, w, numangle, numrho, stride, threshold, accum, peak_cap, st);
    else if ((size_t)4 * stride * 2 <= budget) rc = launch_vote<4>(ctx, n, h, w, numangle, numrho, stride, threshold, accum, peak_cap, st);
    else return le_fail(ctx, LE_ERR_UNSUPPORTED, "image too large for the shared-memory Hough stripes (stride %d)", stride);
    if (rc) return rc;
    LE_T0(ctx, LE_K_SORT, st);
    k_hough_sort<<<n, HS_THREADS, 0, st>>>((const u64 *)ctx->peaks.p, (const int *)ctx->counters.p, peak_cap, numrho,
                                            (float)rho, (float)theta, lines, votes, counts, max_lines);
    LE_T1(ctx, LE_K_SORT, st);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}
