/*
 * line_engine.h -- C ABI of the B200-native line-detection engine (liblineengine.so).
 *
 * Drop-in boundary for the hot path of pklito/Graphics-Project.  The reference has no FFI of
 * its own: its hot path is Python calling the `cv2` module (import cv2 as cv: opencv.py:1,
 * opencv_fit_color.py:1, opencv_fit.py:1).  Each entry point below stands where one cv2 call
 * (or one pure-Python loop of the reference) stands today; the citation after "replaces:" is
 * the reference call site.  The Python shim (graphics-project_b200/cv_shim.py) binds these
 * with ctypes; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", every function returns an int status (LE_OK == 0, negative == error class)
 *     and never throws.  le_last_error(ctx) returns a human-readable message.
 *   - All image / result pointers are DEVICE pointers owned by the caller.  `stream` is a
 *     cudaStream_t passed as void*.  Calls are asynchronous on that stream unless stated.
 *   - Images are batched, dense, row-major: n frames of h x w (x3 for BGR), no padding.
 *   - Scratch memory is owned by the opaque context and grows on demand (first call of a
 *     given size allocates; steady state does not).  A context is bound to one device and is
 *     not thread-safe; use one context per (process, device, stream).
 *     Several contexts may live in one process, on the same or on different devices; every entry
 *     point runs on its context's device and restores the caller's current device before it
 *     returns.  A call that has to grow scratch waits for the device once (allocation + clear).
 *   - Parameter combinations the kernels cannot handle exactly (16-bit vote counters, an order-free
 *     HoughLinesP) return LE_ERR_UNSUPPORTED; an internal list that overflowed (edge queue, Hough
 *     peak list) is reported by le_check as LE_ERR_OVERFLOW.  Nothing is silently approximated.
 *   - The reference-level geometry entries (le_pair_intersections, le_segments_to_polar,
 *     le_combine_parallel_lines, le_face_quads, le_vp_*) compute in float64 where the reference's
 *     numpy stays in float32 for float32 inputs, and le_vp_fit_batch is a deterministic search where
 *     the reference draws random candidates: INTEGRATION.md lists these differences.
 *   - Variable-length results use caller-provided padded buffers [n][max_k][...] plus
 *     counts[n] (int32).  counts[i] holds the TRUE count even if it exceeds max_k (only
 *     max_k entries are written), so the caller can detect truncation and retry.
 */
#ifndef LINE_ENGINE_H_
#define LINE_ENGINE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct le_ctx le_ctx;

enum {
    LE_OK = 0,
    LE_ERR_INVALID = -1,  /* bad argument (null pointer, non-positive size, unsupported parameter) */
    LE_ERR_CUDA = -2,     /* a CUDA runtime call failed; see le_last_error */
    LE_ERR_NOMEM = -3,    /* scratch allocation failed */
    LE_ERR_OVERFLOW = -4, /* an internal work queue overflowed (reported by le_check) */
    LE_ERR_UNSUPPORTED = -5
};

/* Library / build identification: returns e.g. "lineengine 0.1 sm_100a". */
const char *le_version(void);

/* Context life cycle.  max_h/max_w/max_batch are hints used to pre-size scratch. */
int le_create(le_ctx **out, int device, int max_h, int max_w, int max_batch);
int le_destroy(le_ctx *ctx);
const char *le_last_error(const le_ctx *ctx);
/* Bytes of device scratch currently held by the context. */
size_t le_workspace_bytes(const le_ctx *ctx);
/* Synchronises `stream` and reports deferred device-side conditions (queue overflow). */
int le_check(le_ctx *ctx, void *stream);
/* Number of kernel launches issued through this context since creation (for bench accounting). */
long long le_launch_count(const le_ctx *ctx);

/* Per-kernel device timing (CUDA events recorded around each launch on the caller's stream).
 * Kernel ids: 0 front_nms, 1 hysteresis, 2 hough_vote, 3 hough_sort, 4 edge_list, 5 ppht_prep,
 * 6 ppht, 7 lsd_prep, 8 lsd_grow, 9 misc.  le_timing_read synchronises the recorded events, returns
 * the summed duration and launch count since the last read, and resets that kernel's record.     */
int le_timing_enable(le_ctx *ctx, int on);
int le_timing_read(le_ctx *ctx, int kernel_id, double *total_ms, int *launches);
const char *le_kernel_name(int kernel_id);

/* ---- a1  replaces: cv.cvtColor(image, cv.COLOR_BGR2GRAY)   opencv.py:22,205,213; opencv_fit_color.py:74
 * Y = (3735 B + 19235 G + 9798 R + 16384) >> 15                                              */
int le_bgr2gray_u8(le_ctx *ctx, const uint8_t *bgr, uint8_t *gray, int n, int h, int w, void *stream);

/* ---- a2  replaces: cv.GaussianBlur(gray, (5,5), 0)          opencv.py:25; opencv_fit_color.py:77
 * separable [1 4 6 4 1]/16, BORDER_REFLECT_101, out = (sum + 128) >> 8                        */
int le_gauss5_u8(le_ctx *ctx, const uint8_t *src, uint8_t *dst, int n, int h, int w, void *stream);

/* ---- a3  replaces: cv.normalize(x, None, 0, 255, cv.NORM_MINMAX).astype(np.uint8)   opencv.py:26
 * per-frame min/max, then rint(fmaf(v, 255/(max-min), -min*255/(max-min))) saturated          */
int le_minmax_normalize_u8(le_ctx *ctx, const uint8_t *src, uint8_t *dst, int n, int h, int w, void *stream);

/* ---- a4  replaces: cv.Canny(img, lo, hi, apertureSize=3)    opencv.py:26,214; opencv_fit_color.py:78
 * 3x3 Sobel (replicate), L1 magnitude, NMS, double threshold, 8-connected hysteresis; {0,255}. */
int le_canny_u8(le_ctx *ctx, const uint8_t *src, uint8_t *edges, int n, int h, int w, double lo, double hi,
                void *stream);

/* cv.Canny on a 3-CHANNEL image (a cv2 convention; no call site of the reference passes one):
 * Sobel per channel, every pixel keeps the channel with the largest |dx| + |dy| (the first on ties), then NMS
 * and hysteresis as above.  src: uint8 [n][h][w][3].  Leaves the per-frame edge lists like le_front_canny_u8. */
int le_canny_u8c3(le_ctx *ctx, const uint8_t *src, uint8_t *edges, int n, int h, int w, double lo, double hi,
                  void *stream);

/* Fused front end: [BGR->gray] -> [blur5] -> [minmax normalize] -> Canny, one pass over the input.
 *   src_channels: 3 (BGR) or 1 (gray);  do_blur, do_normalize: 0/1.
 * Covers prob() (3,0,0,5,150; opencv.py:213-214), get_camera_angles 'hough' (3,1,0,5,10;
 * opencv_fit_color.py:74-78) and doCanny (3,1,1,30,50; opencv.py:21-28).
 * After this call the context also holds the per-frame edge-pixel lists; passing edges == NULL
 * to le_hough_lines (same n, h, w) votes from those lists instead of re-scanning the map.      */
int le_front_canny_u8(le_ctx *ctx, const uint8_t *src, int src_channels, int do_blur, int do_normalize,
                      uint8_t *edges, int n, int h, int w, double lo, double hi, void *stream);

/* ---- a5  replaces: cv.HoughLines(edges, rho, theta, threshold, None, 0, 0)
 *                    opencv.py:88,118; opencv_fit_color.py:83,194; opencv_fit.py:52,115
 * edges: u8 [n][h][w], non-zero == edge; or NULL to reuse the lists of the preceding
 * le_front_canny_u8 call on this context.
 * accum (optional, may be NULL): int32 [n][numangle+2][numrho+2], fully written (no pre-zeroing needed).
 * lines: float32 [n][max_lines][2] (rho, theta), sorted by (votes desc, accumulator index asc);
 * votes (optional): int32 [n][max_lines];  counts: int32 [n].
 * le_hough_dims gives numangle / numrho for (h, w, rho, theta).
 * counts[f] is the TRUE number of peaks; when it exceeds the peak list (max(4 * max_lines, 65536) entries) the
 * frame's lines are an arbitrary subset and le_check reports LE_ERR_OVERFLOW.  (ceil(rho) + 1) * diagonal must
 * stay below 65535 (16-bit stripe counters; cv2 counts in int32), else LE_ERR_UNSUPPORTED.                  */
int le_hough_dims(int h, int w, double rho, double theta, int *numangle, int *numrho);
int le_hough_lines(le_ctx *ctx, const uint8_t *edges, int n, int h, int w, double rho, double theta, int threshold,
                   int32_t *accum, float *lines, int32_t *votes, int32_t *counts, int max_lines, void *stream);

/* ---- a6  replaces: cv.HoughLinesP(edges, rho, theta, threshold, minLineLength, maxLineGap)
 *                    opencv.py:35-37,217
 * Progressive probabilistic Hough, exact replay of cv::RNG(-1) visiting order.
 * segs: int32 [n][max_segs][4] (x1,y1,x2,y2) in cv2's emission order; counts: int32 [n].
 * exact_order: must be non-zero.  SURVEY 8(b) reserved 0 for an order-free mode graded by "every cv2 segment
 * matched within 1 px, recall >= 99 %"; cv2's own algorithm under another RNG seed recalls 22-37 % of its
 * segments at 1 px (profiles/r2_ppht_order_sensitivity.txt, tests/test_ppht_order.py), so no order-free variant
 * can meet it: 0 returns LE_ERR_UNSUPPORTED instead of results that would fail the stated criterion.
 * rho: (ceil(rho) + 1) * diagonal must stay below 32767 (16-bit accumulator cells), else LE_ERR_UNSUPPORTED.  */
int le_hough_lines_p(le_ctx *ctx, const uint8_t *edges, int n, int h, int w, double rho, double theta,
                     int threshold, int min_line_length, int max_line_gap, int32_t *segs, int32_t *counts,
                     int max_segs, int exact_order, void *stream);

/* ---- a7  replaces: cv.createLineSegmentDetector(refine, scale, sigma_scale, quant, ang_th, log_eps,
 *                    density_th, n_bins).detect(gray)          opencv.py:206-207 (refine 0), :142-145 (refine 2)
 * refine 0 (LSD_REFINE_NONE), 1 (LSD_REFINE_STD), 2 (LSD_REFINE_ADV: rectangle improvement + NFA validation
 * against log_eps).  segs: float32 [n][max_segs][4]; width, prec (optional): float64 [n][max_segs];
 * nfa (optional, written for refine 2 only): float64 [n][max_segs] = cv2's fourth output; counts: int32 [n].
 * Parity: all three modes reproduce cv2 4.13 including order (segments, widths, precisions, NFA values; the
 * rectangle's cos / sin are evaluated in double-double so that its vertices equal glibc's). */
int le_lsd_scaled_dims(int h, int w, double scale, int *sh, int *sw);
int le_lsd(le_ctx *ctx, const uint8_t *gray, int n, int h, int w, int refine, double scale, double sigma_scale,
           double quant, double ang_th, double log_eps, double density_th, int n_bins, float *segs, double *width,
           double *prec, double *nfa, int32_t *counts, int max_segs, void *stream);

/* ---- a10 replaces: edges_to_polar_lines(edges)               opencv_renewed.py:16-17
 * segs: float64 [m][4] (ax, ay, bx, by) -> polar float64 [m][2] (rho, phi).                       */
int le_segments_to_polar(le_ctx *ctx, const double *segs, int m, double *polar, void *stream);

/* ---- a11/a12 replaces: sum_loss(phi, theta, lines) over many candidates
 *                    opencv_fit_color.py:14-26, 95-101 (the loop body of regress_lines :28-47)
 * lines: float64 [m][2] (rho, phi); cand: float64 [k][2] (phi, theta); loss: float64 [k].
 * width/height are the reference's WIDTH/HEIGHT (600/400).                                       */
int le_vp_loss(le_ctx *ctx, const double *lines, int m, const double *cand, int k, double width, double height,
               double *loss, void *stream);

/* Gaussian-sphere vote of pairwise segment intersections (north_star VP voting; new design, seeds
 * candidates for le_vp_loss).  segs: float64 [m][4]; hist: int32 [n_el][n_az], zeroed by the call. */
int le_vp_sphere_hist(le_ctx *ctx, const double *segs, int m, double width, double height, int32_t *hist, int n_az,
                      int n_el, void *stream);

/* ---- a9  replaces: _getIntersections(lines)                  opencv.py:171-189
 * lines: float32 [m][2] (rho, theta).  out: float64 [max_out][2] in the reference's (i<j) order;
 * count: int32 [1] true number of intersections.                                                  */
int le_pair_intersections(le_ctx *ctx, const float *lines, int m, double *out, int32_t *count, int max_out,
                          void *stream);

/* ---- a14 ('lsd' branch) for a whole batch; BASELINE config 5 (--pipeline lines: LSD segments + vanishing-point
 *      Gaussian-sphere voting).  Replaces, for n frames at once and without a host round trip:
 *      the list comprehension of get_camera_angles                opencv_fit_color.py:88-89
 *      (keep ||b - a|| > min_len, then edges_to_polar_lines       opencv_renewed.py:16-17)
 * segs: float32 [n][max_segs][4] and counts: int32 [n] exactly as le_lsd writes them (16-byte aligned).
 * polar: float64 [n][max_segs][2] (rho, phi), kept lines in segment order; pcounts: int32 [n].          */
int le_polar_batch(le_ctx *ctx, const float *segs, const int32_t *counts, int n, int max_segs, double min_len,
                   double *polar, int32_t *pcounts, void *stream);

/*      and regress_lines(lines, iterations, refinement_iterations, refinement_area)   opencv_fit_color.py:28-47
 *      per frame, seeded by the frame's Gaussian-sphere histogram (le_vp_sphere_hist of the lines drawn through
 *      the image).  One CTA per frame evaluates sum_loss (:14-26) for every candidate:
 *      phase 1 = n_phi x n_th grid over [0,pi) x [0,pi/2) + the seeds of the `top` (<= 16) highest histogram
 *      bins; phase 2 = `rounds` grids of n_ref x n_ref around the best so far, the first of side ref_area,
 *      each next one 4/n_ref of the last; a later candidate replaces the best only if strictly lower.
 * polar: float64 [n][max_lines][2]; pcounts: int32 [n]; hist: int32 [n][n_el][n_az] (written) or NULL (no
 * seeds); fit: float64 [n][3] = (phi, theta, loss); a frame with no lines gives (0, 0, 100000), the reference's
 * initial best.                                                                                         */
int le_vp_fit_batch(le_ctx *ctx, const double *polar, const int32_t *pcounts, int n, int max_lines, double width,
                    double height, int n_phi, int n_th, int n_ref, int rounds, double ref_area, int32_t *hist,
                    int n_az, int n_el, int top, double *fit, void *stream);

/* ---- n2  replaces: combineParallelLines(lines, max_distance=5, max_angle=3)   util.py:141-159
 *      (with combineEdges util.py:115-139, edgeDistance :103-108, getEdgeProjection :27-52; callers
 *      opencv.py:225 linesToPlanarGraph, graph.py:116 makeGraphFromLines, opencv_renewed.py:56-60 smoothEdges)
 * Batched: segs float64 [n][max_k][4] (x1,y1,x2,y2), counts int32 [n] or NULL (= max_k each), max_k <= 4096.
 * Repeats "merge the lexicographically first pair (i<j) whose |cos angle| > cos_max_angle and whose
 * edgeDistance < max_distance into slot i, delete j" until no pair is left.  As in the reference, the
 * thresholds passed apply until the first merge and (default_distance, default_cos) = (5, cos 3 deg)
 * afterwards.  out_segs [n][max_k][4] holds the surviving segments in list order, out_counts [n] their
 * number, out_src [n][max_k] (optional) the original index of each slot, bit 30 set if it was merged.  */
int le_combine_parallel_lines(le_ctx *ctx, const double *segs, const int32_t *counts, int n, int max_k,
                              double max_distance, double cos_max_angle, double default_distance,
                              double default_cos, double *out_segs, int32_t *out_src, int32_t *out_counts,
                              void *stream);

/* ---- n4  float32 frame-buffer ingest.  replaces: the CV_32F calls doCanny (opencv.py:21-28) makes
 *      when postProcessFbo (opencv.py:110-113) passes it _fboToImage's float32 view (opencv.py:101-108):
 *      cv.cvtColor(f32 BGR, BGR2GRAY) -> cv.GaussianBlur(f32,(5,5),0) -> cv.normalize(f32,None,0,255,NORM_MINMAX)
 *      -> .astype(np.uint8) -> cv.Canny.  Arithmetic = cv2's FMA3 kernels operation for operation (le_f32.cu).
 * src is addressed with ELEMENT strides (frame, row, pixel, channel; any sign), so the reference's flipped view
 * `raw.reshape(h, w, 3)[::-1, :, ::-1]` of the raw frame-buffer bytes is consumed in place.
 * le_f32_minmax_normalize writes the float32 result (dst_f32) and/or its uint8 truncation (dst_u8).        */
int le_f32_bgr2gray(le_ctx *ctx, const float *src, long long stride_n, long long stride_y, long long stride_x,
                    long long stride_c, float *gray, int n, int h, int w, void *stream);
int le_f32_gauss5(le_ctx *ctx, const float *src, float *dst, int n, int h, int w, void *stream);
int le_f32_minmax_normalize(le_ctx *ctx, const float *src, float *dst_f32, uint8_t *dst_u8, int n, int h, int w,
                            void *stream);
/* Fused: float32 BGR (strided) -> gray -> blur5 -> min-max normalise -> uint8 -> Canny(lo, hi); also leaves the
 * per-frame edge lists for le_hough_lines(edges == NULL) like le_front_canny_u8.                            */
int le_f32_front_canny(le_ctx *ctx, const float *src, long long stride_n, long long stride_y, long long stride_x,
                       long long stride_c, uint8_t *edges, int n, int h, int w, double lo, double hi, void *stream);

/* ---- n3  replaces: get_faces_from_pairs(edges1, edges2, threshold=15)   opencv_renewed.py:62-112
 *      (with util.py segments_distance :198-209, lineIntersection :54-68, pointInConvexPolygon :162-168,
 *      faceCircumference :170-171; callers opencv_renewed.py:176-178, 268-270, 278-280, 310-312)
 * Batched over n problems: segs1 float64 [n][max1][4], segs2 float64 [n][max2][4] (x1,y1,x2,y2), counts1/counts2
 * int32 [n] or NULL (= max each); max1, max2 <= 1024.  Every quad i<k (edges1), j<l (edges2) whose four
 * neighbouring segments lie within `threshold` of each other, in the reference's (i, j, k, l) order:
 * quads int32 [n][max_out][4]; faces float64 [n][max_out][4][2] = the four corner intersections, already
 * oriented as the reference returns them (a corner of two parallel segments is NaN where the reference raises);
 * keep uint8 [n][max_out] = 1 if the face survives the reference's duplicate rule; counts int32 [n] = true
 * number of candidate quads.                                                                              */
int le_face_quads(le_ctx *ctx, const double *segs1, const int32_t *counts1, int max1, const double *segs2,
                  const int32_t *counts2, int max2, int n, double threshold, int32_t *quads, double *faces,
                  uint8_t *keep, int32_t *counts, int max_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LINE_ENGINE_H_ */
