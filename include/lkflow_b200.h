/*
 * lkflow_b200.h -- C ABI of the B200-native optical-flow hot path (liblkflow_b200.so).
 *
 * The reference (chiahuilin0531/LK-Optical-Flow-Method) has no FFI of its own: its hot path is the set
 * of cv2 Python calls made by VP_detection_using_optical_flow.py ("VP.py") and LK_Final.py ("LKF.py").
 * Each entry point below names the call site(s) it replaces.  All image / point pointers are DEVICE
 * pointers (HBM) unless an argument is explicitly called host_*; the caller owns every buffer; the
 * library owns only the opaque workspace `lkf_ctx`.  Every call enqueues work on `stream`
 * (a cudaStream_t passed as void*) and returns without synchronising, so a sequence of calls can be
 * captured in a CUDA graph.  Return value: 0 on success, a negative lkf_status otherwise.  No C++
 * exceptions cross this boundary.  One ctx per (device, stream); calls on one ctx are not re-entrant.
 *
 * Data layout: images are pitched, batched planes.  Pixel (b, y, x, c) of an image lives at
 *   data + b*image_pitch + y*row_pitch + (x*channels + c)*sizeof(element)      (pitches in BYTES).
 * Pyramid level l of a WxH image has size ((W+1)>>1 applied l times, same for H) and NO border
 * (REFLECT_101 is applied by index mirroring inside the kernels).
 */
#ifndef LKFLOW_B200_H
#define LKFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LKF_ABI_VERSION 2
#define LKF_MAX_LEVELS 8

typedef enum lkf_status {
    LKF_OK = 0,
    LKF_ERR_ARG = -1,          /* bad argument (what cv2 rejects with an assertion)              */
    LKF_ERR_CUDA = -2,         /* a CUDA runtime call failed; see lkf_last_error()               */
    LKF_ERR_UNSUPPORTED = -3,  /* valid for cv2 but outside this path (SURVEY 8b "never exercised") */
    LKF_ERR_NOMEM = -4,
    LKF_ERR_OVERFLOW = -5      /* a caller-provided capacity was too small                        */
} lkf_status;

typedef struct lkf_ctx lkf_ctx;    /* opaque workspace (scratch buffers, device properties) */
typedef void *lkf_stream;          /* cudaStream_t */

typedef struct lkf_image {
    void *data;                    /* device pointer */
    int32_t batch, height, width, channels;
    int64_t row_pitch;             /* bytes between rows   */
    int64_t image_pitch;           /* bytes between images */
} lkf_image;

/* cv2 constants used by the scripts (VP.py:100, VP.py:406) */
#define LKF_TERM_CRITERIA_COUNT 1
#define LKF_TERM_CRITERIA_EPS 2
#define LKF_OPTFLOW_USE_INITIAL_FLOW 4
#define LKF_OPTFLOW_LK_GET_MIN_EIGENVALS 8

/* lk_params of VP.py:98-100 / LKF.py:96-98 plus the two cv2 defaults the scripts do not pass */
typedef struct lkf_lk_params {
    int32_t win_w, win_h;          /* winSize (15,15)                                   */
    int32_t max_level;             /* maxLevel 2 (0-based; reduced like cv2 when too small) */
    int32_t crit_type;             /* TERM_CRITERIA_EPS | TERM_CRITERIA_COUNT           */
    int32_t max_count;             /* 10                                                */
    int32_t flags;                 /* 0                                                 */
    double epsilon;                /* 0.03 (squared inside, like cv2)                   */
    double min_eig_threshold;      /* 1e-4                                              */
} lkf_lk_params;

/* feature_params of VP.py:92-95 / LKF.py:90-93 */
typedef struct lkf_gftt_params {
    int32_t max_corners;           /* <=0: unlimited (bounded by the output capacity)   */
    int32_t block_size;            /* 7 in the scripts, cv2 default 3                   */
    int32_t dy_tail_cols;          /* -1: W mod 32 (cv2 4.13 AVX-512 host rule) */
    int32_t reserved;
    double quality_level;          /* 0.3                                               */
    double min_distance;           /* 7                                                 */
} lkf_gftt_params;

/* ---- library / workspace ------------------------------------------------------------------ */
int lkf_abi_version(void);
/* sizeof of an ABI struct as compiled into the library (host arithmetic; bindings check their mirror against it):
 * 0 lkf_image, 1 lkf_lk_params, 2 lkf_gftt_params, 3 lkf_vp_params, 4 lkf_vp_state; -1 otherwise */
int lkf_struct_size(int which);
const char *lkf_strerror(int status);
const char *lkf_last_error(void);                 /* text of the last CUDA failure on this thread */
int lkf_ctx_create(int device, lkf_ctx **out);
void lkf_ctx_destroy(lkf_ctx *ctx);
/* number of kernels launched through this ctx since creation (bench.py "gpu_launches") */
int64_t lkf_ctx_launch_count(const lkf_ctx *ctx);

/* ---- a1: cv.cvtColor(img, cv.COLOR_BGR2GRAY)            VP.py:406, LKF.py:402 -------------- */
int lkf_bgr2gray_u8(const lkf_image *bgr, const lkf_image *gray, lkf_stream stream);
/* ---- a2: cv.GaussianBlur(img, (3,3), 0)                 VP.py:420, LKF.py:416 -------------- */
int lkf_gaussian_blur3_u8(const lkf_image *src, const lkf_image *dst, lkf_stream stream);
/* ---- a1+a2 fused: process_img()                         VP.py:404-425, LKF.py:400-421 ------ */
int lkf_process_img_u8(const lkf_image *bgr, const lkf_image *dst, lkf_stream stream);
/* ---- a3: one pyrDown step of the pyramid cv.calcOpticalFlowPyrLK builds   VP.py:545 -------- */
int lkf_pyr_down_u8(const lkf_image *src, const lkf_image *dst, lkf_stream stream);
/* ---- a1+a2+a3 for a whole pyramid (process_img VP.py:404-425 + the pyramid of VP.py:545): levels[0] = gray+blur of
 *      `bgr`, levels[l] = pyrDown(levels[l-1]); `levels` is an array of n_levels descriptors (one call, the
 *      per-level kernels back to back). */
int lkf_process_pyramid_u8(const lkf_image *bgr, const lkf_image *levels, int n_levels, lkf_stream stream);
/* ---- a4: Scharr Ix/Iy (int16 x2 interleaved) of one level, as PyrLK computes it  VP.py:545 -- */
int lkf_scharr_s16(const lkf_image *src, const lkf_image *dst_ixiy, lkf_stream stream);
/* levels cv2 really uses for a WxH image (maxLevel is reduced when a level would be <= winSize) */
int lkf_effective_max_level(int height, int width, int win_w, int win_h, int max_level);

/* ---- a7: the tracker of cv.calcOpticalFlowPyrLK(prev, next, p0, None, **lk_params)  VP.py:545-546,
 *      LKF.py:531-532.  prev_pyr/next_pyr: arrays of (n_levels) descriptors, level 0 first, all with
 *      the same batch B.  Points are padded per image: prev_pts[b][i] for i < pt_counts[b]
 *      (pt_counts == NULL: every image has max_pts points).  next_pts is also an input when
 *      LKF_OPTFLOW_USE_INITIAL_FLOW is set.  status/err follow cv2 ((N,1) u8 / f32); err of a
 *      status-0 point is 0.  iters (nullable) receives executed LK iterations per point. */
int lkf_pyrlk_sparse(lkf_ctx *ctx, const lkf_lk_params *params, int n_levels,
                     const lkf_image *prev_pyr, const lkf_image *next_pyr,
                     const float *prev_pts, const int32_t *pt_counts, int max_pts,
                     float *next_pts, uint8_t *status, float *err, int32_t *iters,
                     lkf_stream stream);

/* ---- a7 on the raster grid (BASELINE config C4, dense per-pixel flow): the same tracker for
 *      prevPts = every pixel (x, y) of level 0, in raster order; no point array is read.  Outputs hold
 *      height*width entries per image (next_pts [B][H*W][2], status, err, iters), bit-identical to
 *      lkf_pyrlk_sparse on those points.  Supported: odd square windows 3/5/7, at most 4 levels,
 *      flags without LKF_OPTFLOW_USE_INITIAL_FLOW (LKF_ERR_UNSUPPORTED otherwise).
 *      lkf_pyrlk_sparse itself switches to this kernel when it is handed exactly that grid (checked
 *      on the device, no host round trip). */
int lkf_pyrlk_dense(lkf_ctx *ctx, const lkf_lk_params *params, int n_levels,
                    const lkf_image *prev_pyr, const lkf_image *next_pyr,
                    float *next_pts, uint8_t *status, float *err, int32_t *iters, lkf_stream stream);

/* ---- a5: cornerMinEigenVal inside cv.goodFeaturesToTrack          VP.py:503,707 ------------- */
int lkf_corner_min_eigen_f32(const lkf_image *src_u8, const lkf_image *eig_f32, int block_size,
                             int dy_tail_cols, lkf_stream stream);
/* ---- a6: selection part of cv.goodFeaturesToTrack(img, mask=, **feature_params)  VP.py:503,707.
 *      eig: B planes; mask (nullable): n_masks planes per image or shared (mask->batch == n_masks
 *      or B*n_masks).  For image b and mask m the result is corners[(b*n_masks+m)][0..count) as
 *      (x,y) float32, strongest first, and counts[b*n_masks+m].  max_out = row capacity. */
int lkf_gftt_select(lkf_ctx *ctx, const lkf_gftt_params *params, const lkf_image *eig_f32,
                    const lkf_image *mask_u8, int n_masks, float *corners, int32_t *counts,
                    int max_out, lkf_stream stream);

/* ---- a5+a6 fused for a batch: cv.goodFeaturesToTrack(img, mask=m, **feature_params) for every image and each of
 *      its n_masks masks (the scripts call it 4x per refresh on the same image, VP.py:500-507, 702-711); the eig map
 *      lives in the workspace.  image_enable (nullable, device int32[B]): images with 0 are skipped on the device. */
int lkf_gftt_detect(lkf_ctx *ctx, const lkf_gftt_params *params, const lkf_image *src_u8, const lkf_image *mask_u8,
                    int n_masks, const int32_t *image_enable, float *corners, int32_t *counts, int max_out,
                    lkf_stream stream);

/* ---- next row (8f.1): imutils.resize(frame, width=860) = cv2.resize(.., INTER_AREA), non-integer down-scale
 *      (VP.py:433,531).  Tables (host-built, device-resident): destination index d uses entries
 *      [ofs[d], ofs[d+1]) of (src index, alpha) -- cv2's computeResizeAreaTab. */
int lkf_resize_area_u8(const lkf_image *src, const lkf_image *dst, const int32_t *xofs, const int32_t *xsi,
                       const float *xalpha, const int32_t *yofs, const int32_t *ysi, const float *yalpha,
                       lkf_stream stream);

/* ---- a8-a12: tracks -> flow lines -> cross points -> vanishing point (one frame of every clip per call).
 *      Replaces the Python glue of VP.py:326-349 (checkInside), 276-283 (FlowLine), 565-574 (acceptance),
 *      296-324 + 581-595 (cross points), 157-194 + 596-653 (VP state machine, CSV rows), 700 (refresh test).
 *      Host-computed thresholds are passed as the doubles Python would compute (e.g. TP_NUM*TP_UPDATE_RATE). */
#define LKF_VP_RING 32
typedef struct lkf_vp_params {
    double cp_gate_x, cp_gate_y;       /* WID*CP_THOLD, Hei*CP_THOLD            (VP.py:192-194)      */
    double min_ang_dif;                /* MIN_ANG_DIF = 25                      (VP.py:586)          */
    double max_cp_std;                 /* MAX_CP_STD = 1.0                      (VP.py:617-618)      */
    double vp_update_rate;             /* VP_UPDATE_RATE = 0.5                  (VP.py:181-182)      */
    double fl_update_rate;             /* FL_UPDATE_RATE = 0.05                 (VP.py:571)          */
    double min_fl_len;                 /* MIN_FL_LEN = 1.0 (LK_Final.py: 1.5)   (VP.py:570)          */
    double tp_refresh_below;           /* TP_NUM*TP_UPDATE_RATE                 (VP.py:700)          */
    float start_x_sep;                 /* float32(WID*0.05); < 0 disables (LK_Final.py has no such filter, VP.py:588) */
    int32_t vp_ref_num;                /* VP_REF_NUM = 10 (LK_Final.py: 15), <= LKF_VP_RING          */
    int32_t hide_vp_thold;             /* HIDE_VP_THOLD = 50                    (VP.py:644)          */
    int32_t reset_avg_on_hide;         /* 1: VP.py:647, 0: LK_Final.py                               */
    int32_t tp_update_time;            /* TP_UPDATE_TIME = 10                   (VP.py:700)          */
    int32_t vl_ref;                    /* VP_REF = 300: history rows of the vanishing-line fits (VP.py:52, 654); 0 = the whole history (Python's all_vp[-0:]); < 0 = no VL */
} lkf_vp_params;

typedef struct lkf_vp_state {          /* one per clip, device resident, zero-initialised except avg_len */
    double vp_x, vp_y;
    double ring_x[LKF_VP_RING], ring_y[LKF_VP_RING];   /* recent cross points (last LKF_VP_RING) */
    double avg_len[2];                 /* per-group flow-length EMA, starts at MIN_FL_LEN            */
    int32_t ring_alias[LKF_VP_RING];   /* 1: this entry IS the VP object (init aliasing) */
    int32_t vp_init, ring_count, ring_head, vp_ult, tp_ult, n_rows, row_overflow, refresh;
    int64_t n_lines, n_cps;
    /* vanishing-line pair, class VL (VP.py:196-250): least-squares lines y = m x + b and x = mv y + bv through the last
     * vl_ref rows, re-anchored at the VP; updated once per frame while the VP is shown and has moved (VP.py:654, 226) */
    int32_t vp_moved;                  /* Point._moved of the current VP object (VP.py:180); cleared by init / hide */
    int32_t vl_init;                   /* VL._isinit: the fields below are valid                              */
    double vl_m, vl_b, vl_mv, vl_bv;   /* slope, intercept of y(x); slope, intercept of x(y)                  */
    double vl_bp[2];                   /* VP at the last update                                               */
    double vl_lp_y, vl_rp_y, vl_up_x, vl_dp_x;   /* end points (0, lp_y), (WID-1, rp_y), (up_x, 0), (dp_x, Hei-1) */
} lkf_vp_state;

/* p0/p1: [n_clips*2][max_pts_per_group][2] float32 (group g of clip b at list 2b+g), status likewise u8,
 * counts [n_clips*2] int32 (in: tracked points per group; out: survivors, which p0 now holds = `p0s[n] = good_new`).
 * rows: [n_clips][row_capacity][2] float64, appended at states[b].n_rows.  clip_enable (nullable): skip clip b if 0. */
int lkf_vp_frame(const lkf_vp_params *params, lkf_vp_state *states, int n_clips, float *p0, const float *p1,
                 const uint8_t *status, int32_t *counts, int max_pts_per_group, const lkf_image *roi_mask /* gives WID, Hei */,
                 double *rows, int row_capacity, int32_t *refresh_out /* nullable: [n_clips] copy of state.refresh */,
                 const int32_t *clip_enable, lkf_stream stream);
/* REPlace refresh (VP.py:700-718) from a 4-mask corner detection: lists 4b+0,1 -> group 0, 4b+2,3 -> group 1,
 * applied where states[b].refresh is set (or everywhere if force) and both new groups are non-empty. */
int lkf_vp_apply_refresh(lkf_vp_state *states, int n_clips, const float *corners, const int32_t *corner_counts,
                         int max_out, float *p0, int32_t *counts, int max_pts_per_group, int force, lkf_stream stream);

/* ---- measurement helpers (bench.py): instruction-rate probes of the non-tensor pipes.  lkf_probe_pipe: kind 0 FFMA,
 *      1 IMAD, 2 IDP.2A (dp2a), 3 SHF+LOP3 pairs, 4 FFMA+SHF pairs, 5 IMAD/SHF, 6 IMAD/FFMA, 7 IDP.2A/SHF/FFMA alternating; each of blocks*256 threads executes iters*128 probe
 *      instructions (kinds 3, 4: iters*256).  lkf_probe_fp32_fma = kind 0: blocks*256*iters*128*2 flops. ------------- */
int lkf_probe_pipe(int kind, int blocks, int iters, float *out, lkf_stream stream);
int lkf_probe_fp32_fma(int blocks, int iters, float *out, lkf_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* LKFLOW_B200_H */
