// le_f32.cu -- float32 frame-buffer ingest: the CV_32F variants of cvtColor, GaussianBlur and
// normalize that the reference's doCanny (opencv.py:21-28) runs when postProcessFbo (opencv.py:110-113) hands it
// the float32 view made by _fboToImage (opencv.py:101-108: RGB, bottom-up -> [::-1, :, ::-1]).
//
// The arithmetic follows cv2's FMA3 SIMD kernels operation for operation (found by differential testing against
// the cv2 wheel):
//   gray  = fma(R, .299f, fma(B, .114f, G * .587f))
//   blur  rows: fma(x0, .375f, (x-1 + x1) * .25f) + (x-2 + x2) * .0625f        (BORDER_REFLECT_101)
//         cols: (x0 * .375f + (x-1 + x1) * .25f) + (x-2 + x2) * .0625f
//   norm  scale = float(255 * (1 / (max - min))), shift = -float(min * scale), out = fma(v, scale, shift);
//         the reference then truncates with .astype(np.uint8).
// The input is addressed with element strides, so the flipped view is consumed in place (no host copy).
// One tile kernel produces the blurred plane and the frame's min/max; one more pass normalises to u8, which
// then enters the same Canny front kernel as every other chain.
#include "le_common.cuh"

#define F32_TW 64
#define F32_TH 16

static __device__ __forceinline__ int f32_key(float f)
{
    int k = __float_as_int(f);
    return k >= 0 ? k : k ^ 0x7fffffff;  // monotonic in f; its own inverse
}

__global__ void k_f32_minmax_init(int *counters, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        counters[i * LE_C_STRIDE + LE_C_MIN] = 0x7fffffff;
        counters[i * LE_C_STRIDE + LE_C_MAX] = (int)0x80000000;
    }
}

static __device__ __forceinline__ float f32_gray_at(const float *p, long long sc)
{
    const float b = p[0], g = p[sc], r = p[2 * sc];
    return __fmaf_rn(r, 0.299f, __fmaf_rn(b, 0.114f, __fmul_rn(g, 0.587f)));
}

__global__ void k_f32_gray(const float *__restrict__ src, long long sn, long long sy, long long sx, long long sc,
                           float *__restrict__ gray, int h, int w)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    gray[((size_t)f * h + y) * w + x] = f32_gray_at(src + f * sn + y * sy + x * sx, sc);
}

// BGR: src is the strided 3-channel frame (gray formed on the fly); else src is a dense gray plane.
// MINMAX: also reduce the frame's min / max into counters (ordered-int keys).
template <bool BGR, bool MINMAX>
__global__ void __launch_bounds__(256)
k_f32_blur(const float *__restrict__ src, long long sn, long long sy, long long sx, long long sc,
           float *__restrict__ dst, int *__restrict__ counters, int h, int w)
{
    __shared__ float s_g[F32_TH + 4][F32_TW + 4];  // gray tile with a 2-px halo (reflected)
    __shared__ float s_r[F32_TH + 4][F32_TW];      // row-pass results
    __shared__ int s_mn, s_mx;
    const int f = blockIdx.z, x0 = blockIdx.x * F32_TW, y0 = blockIdx.y * F32_TH;
    const int tid = threadIdx.x;
    const float *fr = src + f * sn;
    if (MINMAX && tid == 0) {
        s_mn = 0x7fffffff;
        s_mx = (int)0x80000000;
    }
    for (int i = tid; i < (F32_TH + 4) * (F32_TW + 4); i += 256) {
        const int ty = i / (F32_TW + 4), tx = i - ty * (F32_TW + 4);
        const int yy = d_reflect101(y0 + ty - 2, h), xx = d_reflect101(min(x0 + tx - 2, w + 1), w);
        const float *p = fr + yy * sy + xx * sx;
        s_g[ty][tx] = BGR ? f32_gray_at(p, sc) : p[0];
    }
    __syncthreads();
    for (int i = tid; i < (F32_TH + 4) * F32_TW; i += 256) {
        const int ty = i / F32_TW, tx = i - ty * F32_TW;
        const float *g = &s_g[ty][tx];
        s_r[ty][tx] = __fadd_rn(__fmaf_rn(g[2], 0.375f, __fmul_rn(__fadd_rn(g[1], g[3]), 0.25f)),
                                __fmul_rn(__fadd_rn(g[0], g[4]), 0.0625f));
    }
    __syncthreads();
    int kmn = 0x7fffffff, kmx = (int)0x80000000;
    for (int i = tid; i < F32_TH * F32_TW; i += 256) {
        const int ty = i / F32_TW, tx = i - ty * F32_TW;
        const int x = x0 + tx, y = y0 + ty;
        if (x < w && y < h) {
            const float v = __fadd_rn(__fadd_rn(__fmul_rn(s_r[ty + 2][tx], 0.375f),
                                                __fmul_rn(__fadd_rn(s_r[ty + 1][tx], s_r[ty + 3][tx]), 0.25f)),
                                      __fmul_rn(__fadd_rn(s_r[ty][tx], s_r[ty + 4][tx]), 0.0625f));
            dst[((size_t)f * h + y) * w + x] = v;
            if (MINMAX) {
                const int k = f32_key(v);
                kmn = min(kmn, k);
                kmx = max(kmx, k);
            }
        }
    }
    if (MINMAX) {
        kmn = __reduce_min_sync(0xffffffffu, kmn);
        kmx = __reduce_max_sync(0xffffffffu, kmx);
        if ((tid & 31) == 0) {
            atomicMin(&s_mn, kmn);
            atomicMax(&s_mx, kmx);
        }
        __syncthreads();
        if (tid == 0) {
            atomicMin(&counters[f * LE_C_STRIDE + LE_C_MIN], s_mn);
            atomicMax(&counters[f * LE_C_STRIDE + LE_C_MAX], s_mx);
        }
    }
}

__global__ void k_f32_minmax(const float *__restrict__ src, int *__restrict__ counters, size_t npx)
{
    const int f = blockIdx.y;
    const float *p = src + (size_t)f * npx;
    int kmn = 0x7fffffff, kmx = (int)0x80000000;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        const int k = f32_key(p[i]);
        kmn = min(kmn, k);
        kmx = max(kmx, k);
    }
    kmn = __reduce_min_sync(0xffffffffu, kmn);
    kmx = __reduce_max_sync(0xffffffffu, kmx);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&counters[f * LE_C_STRIDE + LE_C_MIN], kmn);
        atomicMax(&counters[f * LE_C_STRIDE + LE_C_MAX], kmx);
    }
}

__global__ void k_f32_normalize(const float *__restrict__ src, const int *__restrict__ counters,
                                float *__restrict__ dst_f32, u8 *__restrict__ dst_u8, size_t npx)
{
    const int f = blockIdx.y;
    // f32_key is its own inverse: recover the frame's min / max floats from the ordered keys
    const double smin = (double)__int_as_float(f32_key(__int_as_float(counters[f * LE_C_STRIDE + LE_C_MIN])));
    const double smax = (double)__int_as_float(f32_key(__int_as_float(counters[f * LE_C_STRIDE + LE_C_MAX])));
    // cv::normalize, rtype CV_32F: scale rounded to float first, shift = (float)dmin - (float)(smin * scale)
    double scale = 255.0 * ((smax - smin) > 2.220446049250313e-16 ? __ddiv_rn(1.0, smax - smin) : 0.0);
    scale = (double)(float)scale;
    const float fs = (float)scale, fb = (float)(0.0 - (double)(float)__dmul_rn(smin, scale));
    const float *p = src + (size_t)f * npx;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        const float v = __fmaf_rn(p[i], fs, fb);
        if (dst_f32) dst_f32[(size_t)f * npx + i] = v;
        if (dst_u8) dst_u8[(size_t)f * npx + i] = (u8)min(max((int)v, 0), 255);  // astype(np.uint8): truncation
    }
}

static int f32_check(le_ctx *ctx, const void *a, const void *b, int n, int h, int w)
{
    if (!ctx) return LE_ERR_INVALID;
    LE_REQUIRE(ctx, a && b, "null image pointer");
    LE_REQUIRE(ctx, n > 0 && h > 0 && w > 0, "n, h, w must be positive");
    LE_REQUIRE(ctx, (size_t)h * w < ((size_t)1 << 30), "frame too large");
    LE_REQUIRE(ctx, (((size_t)a | (size_t)b) & 3) == 0, "image pointers must be 4-byte aligned");
    return LE_OK;
}

static int f32_counters(le_ctx *ctx, int n, cudaStream_t st)
{
    int rc = le_ensure(ctx, &ctx->counters, sizeof(int) * LE_C_STRIDE * ((size_t)n + 1));
    if (rc) return rc;
    k_f32_minmax_init<<<(n + 255) / 256, 256, 0, st>>>((int *)ctx->counters.p, n);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_f32_bgr2gray(le_ctx *ctx, const float *src, long long stride_n, long long stride_y,
                               long long stride_x, long long stride_c, float *gray, int n, int h, int w, void *stream)
{
    LE_ON_DEVICE(ctx);
    int rc = f32_check(ctx, src, gray, n, h, w);
    if (rc) return rc;
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, n);
    k_f32_gray<<<grid, block, 0, (cudaStream_t)stream>>>(src, stride_n, stride_y, stride_x, stride_c, gray, h, w);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_f32_gauss5(le_ctx *ctx, const float *src, float *dst, int n, int h, int w, void *stream)
{
    LE_ON_DEVICE(ctx);
    int rc = f32_check(ctx, src, dst, n, h, w);
    if (rc) return rc;
    LE_REQUIRE(ctx, src != dst, "le_f32_gauss5 is not in-place");
    dim3 grid((w + F32_TW - 1) / F32_TW, (h + F32_TH - 1) / F32_TH, n);
    k_f32_blur<false, false><<<grid, 256, 0, (cudaStream_t)stream>>>(src, (long long)h * w, w, 1, 0, dst, nullptr, h, w);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

static int f32_normalize_launch(le_ctx *ctx, const float *src, float *dst_f32, u8 *dst_u8, int n, int h, int w,
                                cudaStream_t st)
{
    const size_t npx = (size_t)h * w;
    const int bx = (int)min((size_t)128, (npx + 255) / 256);
    k_f32_normalize<<<dim3(bx, n), 256, 0, st>>>(src, (const int *)ctx->counters.p, dst_f32, dst_u8, npx);
    LE_LAUNCH_CHECK(ctx);
    return LE_OK;
}

extern "C" int le_f32_minmax_normalize(le_ctx *ctx, const float *src, float *dst_f32, uint8_t *dst_u8, int n, int h,
                                       int w, void *stream)
{
    LE_ON_DEVICE(ctx);
    int rc = f32_check(ctx, src, src, n, h, w);
    if (rc) return rc;
    LE_REQUIRE(ctx, dst_f32 || dst_u8, "le_f32_minmax_normalize: no output");
    cudaStream_t st = (cudaStream_t)stream;
    rc = f32_counters(ctx, n, st);
    if (rc) return rc;
    const size_t npx = (size_t)h * w;
    const int bx = (int)min((size_t)128, (npx + 255) / 256);
    k_f32_minmax<<<dim3(bx, n), 256, 0, st>>>(src, (int *)ctx->counters.p, npx);
    LE_LAUNCH_CHECK(ctx);
    ctx->edge_list_for = nullptr;  // the per-frame counters were reused
    return f32_normalize_launch(ctx, src, dst_f32, dst_u8, n, h, w, st);
}

extern "C" int le_f32_front_canny(le_ctx *ctx, const float *src, long long stride_n, long long stride_y,
                                  long long stride_x, long long stride_c, uint8_t *edges, int n, int h, int w,
                                  double lo, double hi, void *stream)
{
    LE_ON_DEVICE(ctx);
    int rc = f32_check(ctx, src, edges, n, h, w);
    if (rc) return rc;
    LE_REQUIRE(ctx, lo == lo && hi == hi, "NaN threshold");
    if (lo > hi) { double t = lo; lo = hi; hi = t; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npx = (size_t)n * h * w;
    rc = le_ensure(ctx, &ctx->lsd0, sizeof(float) * npx);  // blurred f32 plane (tmp0 / tmp1 belong to the u8 front)
    if (rc) return rc;
    rc = le_ensure(ctx, &ctx->lsd1, npx);                  // normalised u8 plane
    if (rc) return rc;
    rc = f32_counters(ctx, n, st);
    if (rc) return rc;
    dim3 grid((w + F32_TW - 1) / F32_TW, (h + F32_TH - 1) / F32_TH, n);
    k_f32_blur<true, true><<<grid, 256, 0, st>>>(src, stride_n, stride_y, stride_x, stride_c, (float *)ctx->lsd0.p,
                                                  (int *)ctx->counters.p, h, w);
    LE_LAUNCH_CHECK(ctx);
    rc = f32_normalize_launch(ctx, (const float *)ctx->lsd0.p, nullptr, (u8 *)ctx->lsd1.p, n, h, w, st);
    if (rc) return rc;
    return le_front_launch(ctx, (const u8 *)ctx->lsd1.p, 1, 0, 0, edges, n, h, w, (int)floor(lo), (int)floor(hi), st);
}
