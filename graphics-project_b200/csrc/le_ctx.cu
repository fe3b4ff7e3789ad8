// le_ctx.cu -- context life cycle, scratch management and error reporting of the C ABI.
#include "le_common.cuh"
#include <stdarg.h>
#include <stdlib.h>

static char g_create_err[512] = "";

int le_fail(le_ctx *ctx, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

int le_ensure(le_ctx *ctx, le_buf *b, size_t bytes)
{
    if (b->bytes >= bytes && b->p) return LE_OK;
    if (b->p) {
        // a previous async user of the old buffer may still be in flight
        cudaDeviceSynchronize();
        cudaFree(b->p);
        ctx->ws_bytes -= b->bytes;
        b->p = nullptr;
        b->bytes = 0;
    }
    size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&b->p, want);
    if (e != cudaSuccess) {
        b->p = nullptr;
        return le_fail(ctx, LE_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    // The clear runs on the legacy default stream, which does not order against the non-blocking streams the
    // callers launch on: without this wait a kernel of the same call could start while the new buffer is still
    // being zeroed under it (seen as an illegal address in the first call of a freshly created context when
    // several contexts were warming up at once).  Growth is rare, so the wait costs nothing.
    cudaMemset(b->p, 0, want);
    cudaDeviceSynchronize();
    b->bytes = want;
    ctx->ws_bytes += want;
    return LE_OK;
}

extern "C" const char *le_version(void) { return "lineengine 0.2 sm_100a"; }

extern "C" int le_create(le_ctx **out, int device, int max_h, int max_w, int max_batch)
{
    if (!out) return le_fail(nullptr, LE_ERR_INVALID, "le_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return le_fail(nullptr, LE_ERR_CUDA, "le_create: no CUDA device available (%s)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return le_fail(nullptr, LE_ERR_INVALID, "le_create: bad device %d", device);
    le_dev_guard on_device(device);  // the caller's current device is restored on return
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return le_fail(nullptr, LE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    le_ctx *ctx = (le_ctx *)calloc(1, sizeof(le_ctx));
    if (!ctx) return le_fail(nullptr, LE_ERR_NOMEM, "le_create: host allocation failed");
    ctx->device = device;
    ctx->max_h = max_h;
    ctx->max_w = max_w;
    ctx->max_batch = max_batch;
    ctx->trig_kind = -1;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (max_h > 0 && max_w > 0 && max_batch > 0) {
        int rc = le_ensure(ctx, &ctx->counters, sizeof(int) * LE_C_STRIDE * ((size_t)max_batch + 1));
        if (rc == LE_OK) rc = le_ensure(ctx, &ctx->queue, sizeof(u32) * (size_t)max_batch * max_h * max_w);
        if (rc != LE_OK) {
            strncpy(g_create_err, ctx->err, 511);
            le_destroy(ctx);
            return rc;
        }
    }
    *out = ctx;
    return LE_OK;
}

extern "C" int le_destroy(le_ctx *ctx)
{
    if (!ctx) return LE_OK;
    LE_ON_DEVICE(ctx);
    cudaDeviceSynchronize();
    le_buf *bufs[] = {&ctx->counters, &ctx->queue, &ctx->lut, &ctx->peaks, &ctx->trig, &ctx->tmp0, &ctx->tmp1,
                      &ctx->ppht_accum, &ctx->ppht_mask, &ctx->ppht_nz, &ctx->ppht_rng, &ctx->lsd0, &ctx->lsd1, &ctx->lsd2,
                      &ctx->lsd3, &ctx->lsd4, &ctx->lsd5, &ctx->lsd6, &ctx->lsd7, &ctx->lsd_tab, &ctx->misc};
    for (le_buf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (int id = 0; id < LE_K_COUNT; id++) {
        for (int i = 0; i < ctx->tcap[id]; i++) cudaEventDestroy(ctx->tev[id][i]);
        free(ctx->tev[id]);
    }
    free(ctx);
    return LE_OK;
}

extern "C" const char *le_last_error(const le_ctx *ctx) { return ctx ? ctx->err : g_create_err; }
extern "C" size_t le_workspace_bytes(const le_ctx *ctx) { return ctx ? ctx->ws_bytes : 0; }
extern "C" long long le_launch_count(const le_ctx *ctx) { return ctx ? ctx->launches : 0; }

static const char *g_kernel_names[LE_K_COUNT] = {"front_nms", "hysteresis", "hough_vote", "hough_sort", "edge_list",
                                                 "ppht_prep", "ppht", "lsd_prep", "lsd_grow", "misc"};
extern "C" const char *le_kernel_name(int id) { return (id >= 0 && id < LE_K_COUNT) ? g_kernel_names[id] : ""; }

void le_time_mark(le_ctx *ctx, int id, int end, cudaStream_t st)
{
    if (id < 0 || id >= LE_K_COUNT) return;
    if (!end) {
        if (ctx->tn[id] + 2 > ctx->tcap[id]) {
            int ncap = ctx->tcap[id] ? ctx->tcap[id] * 2 : 256;
            cudaEvent_t *nv = (cudaEvent_t *)realloc(ctx->tev[id], sizeof(cudaEvent_t) * ncap);
            if (!nv) return;
            for (int i = ctx->tcap[id]; i < ncap; i++) cudaEventCreate(&nv[i]);
            ctx->tev[id] = nv;
            ctx->tcap[id] = ncap;
        }
        cudaEventRecord(ctx->tev[id][ctx->tn[id]], st);
    } else {
        cudaEventRecord(ctx->tev[id][ctx->tn[id] + 1], st);
        ctx->tn[id] += 2;
    }
}

extern "C" int le_timing_enable(le_ctx *ctx, int on)
{
    if (!ctx) return LE_ERR_INVALID;
    ctx->timing_on = on ? 1 : 0;
    return LE_OK;
}

extern "C" int le_timing_read(le_ctx *ctx, int id, double *total_ms, int *launches)
{
    LE_ON_DEVICE(ctx);
    if (!ctx || id < 0 || id >= LE_K_COUNT || !total_ms || !launches) return LE_ERR_INVALID;
    double tot = 0;
    for (int i = 0; i + 1 < ctx->tn[id]; i += 2) {
        LE_CUDA(ctx, cudaEventSynchronize(ctx->tev[id][i + 1]));
        float ms = 0;
        LE_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->tev[id][i], ctx->tev[id][i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = ctx->tn[id] / 2;
    ctx->tn[id] = 0;
    return LE_OK;
}

__global__ void k_collect_overflow(const int *counters, int n, int *out)
{
    int any = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) any |= counters[i * LE_C_STRIDE + LE_C_OVERFLOW];
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) *out = any;
}

extern "C" int le_check(le_ctx *ctx, void *stream)
{
    LE_ON_DEVICE(ctx);
    if (!ctx) return LE_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    LE_CUDA(ctx, cudaStreamSynchronize(st));
    if (!ctx->counters.p) return LE_OK;
    int rc = le_ensure(ctx, &ctx->misc, 256);
    if (rc) return rc;
    int n = (int)(ctx->counters.bytes / (sizeof(int) * LE_C_STRIDE));
    if (ctx->edge_list_n > 0 && ctx->edge_list_n < n) n = ctx->edge_list_n;
    k_collect_overflow<<<1, 256, 0, st>>>((const int *)ctx->counters.p, n, (int *)ctx->misc.p);
    LE_LAUNCH_CHECK(ctx);
    int h = 0;
    LE_CUDA(ctx, cudaMemcpyAsync(&h, ctx->misc.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LE_CUDA(ctx, cudaStreamSynchronize(st));
    if (h) return le_fail(ctx, LE_ERR_OVERFLOW, "device work queue overflowed (flags 0x%x)", h);
    return LE_OK;
}
