// le_common.cuh -- shared definitions of the line-engine CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "../../include/line_engine.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef unsigned long long u64;

// Per-frame device counters (int32 each).
enum {
    LE_C_QTAIL = 0,   // edge queue: number of entries (strong seeds + promoted weak)
    LE_C_WEAK = 1,    // weak list length (grows from the back of the queue buffer)
    LE_C_PEAKS = 2,   // Hough peaks found
    LE_C_OVERFLOW = 3,
    LE_C_MIN = 4,
    LE_C_MAX = 5,
    LE_C_AUX0 = 6,
    LE_C_AUX1 = 7,
    LE_C_STRIDE = 8
};

enum { LE_K_FRONT = 0, LE_K_HYST, LE_K_VOTE, LE_K_SORT, LE_K_EDGELIST, LE_K_PPHT_PREP, LE_K_PPHT, LE_K_LSD_PREP,
       LE_K_LSD_GROW, LE_K_MISC, LE_K_COUNT };

struct le_buf {
    void *p;
    size_t bytes;
};

struct le_ctx {
    int device;
    int max_h, max_w, max_batch;
    char err[512];
    long long launches;
    size_t ws_bytes;
    // scratch buffers (grown on demand)
    le_buf counters;    // int32 [n][LE_C_STRIDE]
    le_buf queue;       // u32   [n][h*w]
    le_buf lut;         // u8    [n][256]
    le_buf peaks;       // u64   [n][peak_cap]
    le_buf trig;        // float [2][numangle] + int [numangle] (rlo)
    le_buf tmp0, tmp1;  // generic images
    le_buf ppht_accum, ppht_mask, ppht_nz, ppht_rng;
    size_t ppht_rng_n;  // entries of the cv::RNG(-1) output table in ppht_rng
    le_buf lsd0, lsd1, lsd2, lsd3, lsd4, lsd5, lsd6, lsd7, lsd_tab;
    double lsd_key[4];
    int lsd_epoch;  // call counter carried by the speculative grower's tags (le_lsd.cu)
    le_buf misc;
    // state shared between le_front_canny_u8 and le_hough_lines
    const void *edge_list_for;  // edges pointer the queue currently describes (or null)
    int edge_list_n, edge_list_h, edge_list_w;
    // cached trig tables
    double trig_rho, trig_theta;
    int trig_kind, trig_h, trig_w;
    int sm_count;
    int smem_optin;
    // per-kernel event timing
    int timing_on;
    cudaEvent_t *tev[LE_K_COUNT];  // pairs (begin, end)
    int tn[LE_K_COUNT], tcap[LE_K_COUNT];
};

void le_time_mark(le_ctx *ctx, int id, int end, cudaStream_t st);
#define LE_T0(ctx, id, st) do { if ((ctx)->timing_on) le_time_mark(ctx, id, 0, st); } while (0)
#define LE_T1(ctx, id, st) do { if ((ctx)->timing_on) le_time_mark(ctx, id, 1, st); } while (0)

int le_fail(le_ctx *ctx, int code, const char *fmt, ...);

// cudaFuncSetAttribute is a per-DEVICE setting: every call site remembers what it has set per device (a process
// may hold contexts on several GPUs).  `v` is the size (or 1) already set on that device.
#define LE_MAX_DEVICES 64
struct le_dev_once {
    size_t v[LE_MAX_DEVICES];
    bool needs(int dev, size_t want) const { return dev < 0 || dev >= LE_MAX_DEVICES || v[dev] < want; }
    void done(int dev, size_t want) { if (dev >= 0 && dev < LE_MAX_DEVICES) v[dev] = want; }
};

// Development knobs (A/B measurements, the comparison runs of the tests) are environment variables read ONCE per
// process and call site, never on the per-call path.  `present`: the variable exists (value ignored).
#include <stdlib.h>
#define LE_KNOB_INT(name, dflt) ([] { static const int v = [] { const char *e = getenv(name); return e ? atoi(e) : (dflt); }(); return v; }())
#define LE_KNOB_SET(name) ([] { static const bool v = getenv(name) != nullptr; return v; }())

// Every ABI entry runs on its context's device and leaves the caller's current device as it found it.
struct le_dev_guard {
    int prev;
    explicit le_dev_guard(int dev) : prev(-1)
    {
        if (dev < 0) return;  // null context: the entry point reports it
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~le_dev_guard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define LE_ON_DEVICE(ctx) le_dev_guard le_dev_guard__((ctx) ? (ctx)->device : -1)
int le_ensure(le_ctx *ctx, le_buf *b, size_t bytes);

#define LE_CUDA(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return le_fail(ctx, LE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                           __FILE__, __LINE__);                                                    \
    } while (0)

#define LE_LAUNCH_CHECK(ctx)                                                                   \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess)                                                                \
            return le_fail(ctx, LE_ERR_CUDA, "kernel launch failed: %s (%s:%d)",               \
                           cudaGetErrorString(e__), __FILE__, __LINE__);                       \
    } while (0)

#define LE_REQUIRE(ctx, cond, msg)                                   \
    do {                                                             \
        if (!(cond)) return le_fail(ctx, LE_ERR_INVALID, "%s", msg); \
    } while (0)

static __device__ __forceinline__ int d_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}
static __device__ __forceinline__ int d_clamp(int p, int lo, int hi) { return min(max(p, lo), hi); }

// internal entry points shared between translation units
int le_front_launch(le_ctx *ctx, const u8 *src, int ch, int do_blur, int do_norm, u8 *edges, int n, int h, int w,
                    int lo, int hi, cudaStream_t st);
int le_edge_list_from_map(le_ctx *ctx, const u8 *edges, int n, int h, int w, cudaStream_t st);
