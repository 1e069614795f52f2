/*
 * wtracker_b200 — C ABI of the B200-native detect+predict hot path.
 *
 * The reference (giladfrid009/WTracker) has no FFI of its own: its plugin boundary is the Python
 * ABC `SimController` (wtracker/sim/simulator.py:197-293) and the arithmetic below runs inside
 * ultralytics / torch / cv2 / numpy on its behalf.  Each entry point cites the reference code whose
 * arithmetic it replaces.  Conventions:
 *
 *   - every pointer is a raw DEVICE pointer unless named `h_*`; PyTorch (or any caller) owns all
 *     memory including the engine workspace; no entry point allocates, frees or synchronises,
 *     except wt_engine_create/destroy (host-side bookkeeping only) and wt_selftest_* (test only);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value: 0 = ok, non-zero = error; wt_last_error() gives a thread-local message;
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute call fails.
 */
#ifndef WTRACKER_B200_H
#define WTRACKER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WT_ABI_VERSION 12

/* ------------------------------------------------------------------------------------------ */
/* errors / info                                                                              */
/* ------------------------------------------------------------------------------------------ */
const char* wt_last_error(void);
int wt_abi_version(void);
/* Number of kernel launches issued by this library since load (all threads). bench.py's
 * `gpu_launches` is the difference of two readings. */
uint64_t wt_launch_count(void);
/* Fills sm count / compute capability of the current device. */
int wt_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------ */
/* K1-K4  camera-view crop + letterbox resize                                                 */
/*   replaces ViewController.read/_custom_view (wtracker/sim/view_controller.py:45-61,158-172) */
/*   cv.cvtColor GRAY2BGR (wtracker/sim/sim_controllers/yolo_controller.py:68-69) and the      */
/*   ultralytics LetterBox + preprocess that predict() runs (yolo_controller.py:72-78).        */
/* ------------------------------------------------------------------------------------------ */
typedef struct wt_letterbox {
    int32_t src_w, src_h;     /* camera view size (w,h) in px (crop taken from the frame)      */
    int32_t dst_w, dst_h;     /* letterboxed network input size                                */
    int32_t new_w, new_h;     /* resized (unpadded) size; new == src -> no resampling          */
    int32_t pad_left, pad_top;/* letterbox padding (value 114)                                 */
    /* cv2 INTER_LINEAR u8 fixed-point tables, device pointers (ignored when new == src):      */
    const int32_t* xofs;      /* [new_w]   left source column                                  */
    const int16_t* xcoef;     /* [new_w*2] 11-bit weights                                      */
    const int32_t* yofs;      /* [new_h]   top source row                                      */
    const int16_t* ycoef;     /* [new_h*2]                                                     */
} wt_letterbox;

/* frames : u8 [n_frames][frame_h][frame_w] grayscale, resident in HBM
 * frame_idx[i], crop_x[i], crop_y[i] : per output image i, source frame and the crop origin in
 *          FRAME coordinates (= position - size//2, may be negative or run past the frame:
 *          border pixels replicate, as cv.copyMakeBorder(BORDER_REPLICATE) + slicing does)
 * out_u8 : u8 [n][dst_h][dst_w] letterboxed grey image (the product path; conv0 folds /255 and
 *          the 3 identical channels), or NULL
 * out_f32: f32 [n][3][dst_h][dst_w] in [0,1] — the tensor ultralytics would feed the net, or NULL */
int wt_preprocess(const uint8_t* frames, int n_frames, int frame_h, int frame_w,
                  const int32_t* frame_idx, const int32_t* crop_x, const int32_t* crop_y, int n,
                  const wt_letterbox* lb, uint8_t* out_u8, float* out_f32, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* K5  YOLOv8s backbone/neck/head as a program of fused ops over NHWC bf16 buffers            */
/*   replaces ultralytics AutoBackend.forward (called from yolo_controller.py:72-78)           */
/* ------------------------------------------------------------------------------------------ */
enum { WT_OP_CONV0 = 0, WT_OP_CONV = 1, WT_OP_SPPF_POOL = 2 };
enum { WT_ACT_NONE = 0, WT_ACT_SILU = 1 };
enum { WT_DT_BF16 = 0, WT_DT_F32 = 1, WT_DT_U8 = 2 };

typedef struct wt_buf {          /* one NHWC activation buffer (batch = engine batch)          */
    int32_t h, w, c;             /* c = total channels (concat destinations hold all parts)    */
    int32_t dtype;               /* WT_DT_*                                                    */
} wt_buf;

typedef struct wt_op {
    int32_t kind;                /* WT_OP_*                                                    */
    int32_t src, src_coff;       /* source buffer id, first channel                            */
    int32_t dst, dst_coff;       /* destination buffer id, first channel                       */
    int32_t res, res_coff;       /* residual buffer id (-1 = none), first channel              */
    int32_t cin, cout;           /* conv channels (POOL: cin = channels pooled)                 */
    int32_t k, stride;           /* kernel size (1|3), stride (1|2); pad = k/2                 */
    int32_t act;                 /* WT_ACT_*                                                   */
    int64_t w_off, b_off;        /* byte offsets into the weight blob:                         */
                                 /*   CONV : bf16 [cout][k][k][cin], f32 bias[cout]            */
                                 /*   CONV0: f32 [cout][3][3] (grey-folded, /255 folded), f32 bias */
    int64_t dot_off;             /* CONV only, -1 = none.  Otherwise the byte offset of        */
                                 /* f32 [cout + 1] = weights w[cout] then a bias b of a FUSED   */
                                 /* following 1x1 convolution with ONE output channel (the     */
                                 /* class-logit conv of the head, nc = 1): the activated output */
                                 /* is not stored; dst (f32, c = 1) receives                    */
                                 /* sum_c out[c] * w[c] + b per pixel.  Needs cout <= 256.      */
    int32_t add_buf, add_coff;   /* CONV only, add_buf = -1: none.  Otherwise an f32 buffer of HALF the  */
                                 /* destination's spatial size whose channels [add_coff, add_coff+cout) */
                                 /* are added, nearest-2x-upsampled, to the accumulator BEFORE the      */
                                 /* activation.  A 1x1 conv over concat(upsample2x(a), b) is split by    */
                                 /* linearity into conv_a(a) at half resolution (this addend, bias       */
                                 /* included) + conv_b(b): the upsampled tensor is never materialised.   */
    int32_t lane;                /* 0 = the caller's stream; 1 = the engine's side stream.  Ops are launched in    */
                                 /* program order; an op that reads or overwrites what an op of the OTHER lane     */
                                 /* wrote or read waits for it through an event (the engine derives the hazards    */
                                 /* from the buffer ids and channel ranges), and the caller's stream joins the    */
                                 /* side stream before wt_engine_forward returns.  Independent branches (the head  */
                                 /* of one pyramid level vs the rest of the neck) then overlap: the drain of one   */
                                 /* persistent kernel is filled by the ramp-up of the other lane's next kernel.   */
    int32_t chain_act;           /* CONV only: activation of the chained conv below (WT_ACT_*)                      */
    int64_t chain_w_off;         /* CONV only, -1 = none.  Otherwise a 1x1 convolution cout -> cout is CHAINED onto  */
    int64_t chain_b_off;         /* this one: dst receives act2(W2 * bf16(act(conv(src))) + b2).  Byte offsets of   */
                                 /* bf16 [cout][cout] and f32 bias[cout].  The intermediate map is rounded to bf16  */
                                 /* exactly as if it had been stored, but only exists as tiles in shared memory     */
                                 /* (they are the A operand of a second UMMA chain).  Needs cout = 64 | 128, no      */
                                 /* residual / addend / dot head.  (YOLOv8: a stride-2 conv followed by C2f.cv1,     */
                                 /* which is the only consumer of the conv's output.)                               */
    int32_t cat_buf, cat_coff;   /* CONV with a chain only, cat_buf = -1: none.  Otherwise a CONCAT chain: the chained  */
    int32_t cat_c, chain_cout;   /* 1x1 conv runs over concat(channels [cat_coff, cat_coff + cat_c) of buffer cat_buf,  */
                                 /* this conv's output) and has chain_cout outputs; chain weights are bf16              */
                                 /* [chain_cout][cat_c + cout].  This conv's residual must be the upper half of that    */
                                 /* slice.  Implemented for the exit of a C2f block with one bottleneck and 32 hidden   */
                                 /* channels (3x3 32 -> 32 + y1, then cv2 over [y0 | y1 | b], 96 -> 64): the bottleneck */
                                 /* output never leaves shared memory and [y0 | y1] is read once.                       */
} wt_op;

typedef struct wt_engine wt_engine;

/* bytes of workspace needed for `batch` images with these buffers */
int64_t wt_engine_workspace_bytes(const wt_buf* bufs, int n_bufs, int batch);
/* conv_impl: 0 = tcgen05 implicit GEMM (product), 1 = scalar validation kernel (tests only).
 * weights/workspace are device pointers that must outlive the engine. */
int wt_engine_create(const wt_buf* bufs, int n_bufs, const wt_op* ops, int n_ops, int batch,
                     const void* weights, int64_t weight_bytes, void* workspace, int64_t workspace_bytes,
                     int conv_impl, wt_engine** out);
void wt_engine_destroy(wt_engine* e);
/* device address of buffer `id` ([batch][h][w][c]) */
void* wt_engine_buffer(wt_engine* e, int id);
/* run ops [first, last) for the first n (<= batch) images; buffer 0 must hold the input */
int wt_engine_forward(wt_engine* e, int n, int first_op, int last_op, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* K6-K8  DFL decode + confidence filter + NMS + scale_boxes                                  */
/*   replaces ultralytics Detect._inference / non_max_suppression / scale_boxes and the        */
/*   first-box / NaN-row logic of YoloController.predict (yolo_controller.py:80-90)            */
/* ------------------------------------------------------------------------------------------ */
typedef struct wt_head_level {
    const void* box;        /* [n][h*w][64] box-branch logits (4 sides x 16 DFL bins), box_dtype; */
                            /* or NULL when box_feat is given                                    */
    const void* cls_feat;   /* [n][h*w][cls_c] bf16 features feeding the final 1x1 cls conv, or  */
    const float* cls_logit; /* [n][h*w] f32 ready-made class logits (oracle-fed tests); one of   */
                            /* cls_feat / cls_logit is non-NULL                                  */
    int32_t h, w, stride;   /* grid and stride (8/16/32)                                         */
    int32_t box_dtype;      /* WT_DT_BF16 | WT_DT_F32                                            */
    int32_t cls_c;          /* channels of cls_feat (128)                                        */
    const void* cls_w;      /* bf16 [cls_c] weight of the 1x1 cls conv (nc = 1)                  */
    float cls_b;            /* its bias                                                          */
    /* Alternative to `box` (box == NULL): the final 1x1 conv of the box branch is evaluated ONLY  */
    /* for the anchors that survive the confidence filter (a few per image instead of all 8400):   */
    const void* box_feat;   /* [n][h*w][box_c] bf16 features feeding that conv                    */
    const void* box_w;      /* bf16 [64][box_c] its weight                                        */
    const float* box_b;     /* f32 [64] its bias                                                  */
    int32_t box_c;          /* channels of box_feat (64), a multiple of 8                         */
} wt_head_level;

typedef struct wt_post_params {
    float conf_thres;       /* keep anchors with conf > conf_thres                    (0.1)      */
    float iou_thres;        /* suppress when IoU > iou_thres                          (0.7)      */
    int32_t max_det;        /* boxes kept per image                                   (1)        */
    int32_t net_w, net_h;   /* letterboxed input size                                            */
    int32_t img_w, img_h;   /* original image size boxes are scaled back to                      */
    float gain;             /* scale_boxes gain = min(net_h/img_h, net_w/img_w)                  */
    float pad_x, pad_y;     /* scale_boxes padding = round((net - img*gain)/2 - 0.1)             */
} wt_post_params;

/* out_boxes : f32 [n][max_det][6] = x1,y1,x2,y2 (original image px, clipped), conf, anchor index
 * out_count : i32 [n] kept boxes per image (0 => the reference returns a NaN row)
 * scratch   : >= wt_post_scratch_bytes(n, total_anchors) bytes, 16-byte aligned, ZEROED once before its first use
 *             (the max_det == 1 form keeps two words per image in it and leaves them zeroed after every call)    */
int64_t wt_post_scratch_bytes(int n, int total_anchors);
int wt_decode_nms(const wt_head_level* levels, int n_levels, int n, const wt_post_params* p,
                  float* out_boxes, int32_t* out_count, void* scratch, void* stream);

/* Rows of the tracking log from one batch of detections (the absolute-coordinate shift of
 * LoggingController._log_cycle, wtracker/sim/sim_controllers/logging_controller.py:152-155, and the
 * camera / microscope boxes of ViewController.camera_position / micro_position,
 * wtracker/sim/view_controller.py:93-117):
 *   worm_xywh[i] = best box of image i as (x, y, w, h) in FRAME pixels; when count == 0: a NaN row, or with
 *                  none_as_zero the row 0, 0, 0, 0 that bboxes.csv holds for such a frame (discretize zeroes the
 *                  NaN prediction in place before it is logged, logging_controller.py:153-158)
 *   mic_xywh[i]  = microscope box centred like the camera view
 * boxes: f32 [n][max_det][6], count: i32 [n] (outputs of wt_decode_nms); crop_x/y: camera-view origin. */
int wt_track_rows(const float* boxes, const int32_t* count, int max_det, const int32_t* crop_x, const int32_t* crop_y,
                  int cam_w, int cam_h, int mic_w, int mic_h, double* worm_xywh, double* mic_xywh, int64_t n,
                  int none_as_zero, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* K9  ResMLP position predictor                                                              */
/*   replaces WormPredictor.forward / RMLP.forward (wtracker/neural/mlp.py:47-48,176-188)      */
/*   as called by MLPController.provide_movement_vector (mlp_controllers.py:59)                */
/* ------------------------------------------------------------------------------------------ */
typedef struct wt_resmlp_desc {
    int32_t in_dim, hidden, out_dim;  /* 28, 40, 2                                               */
    int32_t n_blocks, block_len;      /* 4 blocks of block_len=4 layers                          */
    int32_t block_dims[8];            /* output width of each layer in a block (10,4,10,40)      */
    /* weights: f32 device blob, BN folded, layers in execution order, each [out][in] then [out] */
    const float* weights;
    int32_t n_weights;                /* floats in the blob                                      */
} wt_resmlp_desc;
/* x: f32 [n][in_dim] -> y: f32 [n][out_dim] */
int wt_resmlp_forward(const wt_resmlp_desc* d, const float* x, float* y, int64_t n, void* stream);
/* Network input rows from a bbox table — the gather + "relative to the first box" step of
 * MLPController.provide_movement_vector (mlp_controllers.py:38-56) and CsvController.predict
 * (csv_controller.py:25-37): for sample i and input offset j the row table[frame[i] + offsets[j]]
 * (NaN when outside [0, table_rows)); x, y of every box minus x, y of box 0; float64 arithmetic,
 * cast to f32.  valid[i] = 0 when any gathered value is non-finite (the controller then answers (0,0)).
 * table: f64 [table_rows][4]; frame: i32 [n]; offsets: i32 [k] (device); x: f32 [n][4k]; valid: u8 [n]. */
int wt_mlp_gather(const double* table, int64_t table_rows, const int32_t* frame, const int32_t* offsets, int k,
                  float* x, uint8_t* valid, int64_t n, void* stream);

/* The tail of the hot path as ONE launch (the product path of HotPath / bench.py): for batch items i = 0..n-1, whose
 * results go to rows first_row + i of the tracking table,
 *   table[first_row + i], mic_table[first_row + i]  <- wt_track_rows
 *   x[i], valid[i]                                  <- wt_mlp_gather with frame[i] = first_row + i (rows of this batch
 *                                                      are taken straight from boxes / count, earlier rows from table)
 *   y[i]                                            <- wt_resmlp_forward(x[i])      (bit-identical to that entry point)
 *   err[i]                                          <- wt_bbox_error(table row, mic row)
 * replaces, per frame: logging_controller.py:152-155 + view_controller.py:93-117, mlp_controllers.py:38-56,
 * neural/mlp.py:176-188, eval/error_calculator.py:163-195.
 * weights_t: the ResMLP blob with every layer TRANSPOSED, f32 [in][out] then bias[out], layers in execution order,
 * padded with zeros to a multiple of 4 floats, 16-byte aligned (mlp.weights is not read).  x may be NULL. */
#define WT_TAIL_MAX_K 16
typedef struct wt_tail_args {
    const float* boxes;            /* f32 [n][max_det][6] (wt_decode_nms)                          */
    const int32_t* count;          /* i32 [n]                                                       */
    int32_t max_det;
    const int32_t* crop_x;         /* i32 [n] camera-view origin in frame px                        */
    const int32_t* crop_y;
    int32_t cam_w, cam_h, mic_w, mic_h;
    double* table;                 /* f64 [table_rows][4] worm xywh by row (read + written)         */
    double* mic_table;             /* f64 [table_rows][4]                                           */
    int64_t table_rows, first_row, n;
    int32_t k;                     /* input boxes per sample (= in_dim / 4)                         */
    int32_t offsets[WT_TAIL_MAX_K];/* row offsets of the k input boxes (IOConfig.input_frames)      */
    wt_resmlp_desc mlp;
    const float* weights_t;
    float* x;                      /* f32 [n][4k] or NULL                                           */
    uint8_t* valid;                /* u8 [n]                                                        */
    float* y;                      /* f32 [n][out_dim]                                              */
    double* err;                   /* f64 [n]                                                       */
} wt_tail_args;
int wt_hot_tail(const wt_tail_args* args, void* stream);

/* Rows of the per-frame result table that the ranks gather at the end of an offline run (the values
 * YoloController.predict returns, yolo_controller.py:80-90): for detection i of frame first_frame + i
 *   rows[i] = 8 x 32 bit: f32 x, y, w, h (view px; NaN when nothing passed conf), f32 conf, i32 kept anchor index
 *             (-1 = none), i32 frame index, i32 valid (0 | 1).   rows: 16-byte aligned. */
int wt_result_rows(const float* boxes, const int32_t* count, int max_det, int64_t first_frame, int32_t* rows, int64_t n,
                   void* stream);

/* ------------------------------------------------------------------------------------------ */
/* K10  per-step bbox metrics                                                                 */
/*   replaces ErrorCalculator.calculate_bbox_error / calculate_mse_error                       */
/*   (wtracker/eval/error_calculator.py:163-195, 197-212); float64 like the reference          */
/* ------------------------------------------------------------------------------------------ */
int wt_bbox_error(const double* worm_xywh, const double* mic_xywh, double* err, int64_t n, void* stream);
int wt_mse_error(const double* worm_xywh, const double* mic_xywh, double* err, int64_t n, void* stream);

/* Rows of the tracking log `bboxes.csv` for n consecutive frames starting at first_frame (replaces the per-frame
 * part of LoggingController._log_cycle, wtracker/sim/sim_controllers/logging_controller.py:145-185, and
 * BoxUtils.discretize, wtracker/utils/bbox_utils.py:119-167):
 *   worm_rel  : [n][4] camera-relative worm boxes (x, y, w, h), f64 or f32 (worm_is_f32), NaN row = no prediction
 *   cam_xywh, mic_xywh : i32 [n][4] camera / microscope boxes; plt_xy : i32 [n][2] platform positions
 *   table     : f64 [n][17] = frame, cycle, phase (0 imaging | 1 moving), plt_x, plt_y, cam_x, cam_y, cam_w, cam_h,
 *               mic_x, mic_y, mic_w, mic_h, wrm_x, wrm_y, wrm_w, wrm_h   (the csv column order; wrm absolute,
 *               rows without a prediction are 0, 0, 0, 0 exactly as the reference logs them)
 *   crop_xywh : i32 [n][4] integer crop of the worm view clipped to the frame (0 if empty), crop_legal: u8 [n]
 * cycle = frame / cycle_frame_num, phase = (frame % cycle_frame_num) < imaging_frame_num.
 * worm_rel, cam_xywh, mic_xywh, crop_xywh: 16-byte aligned (rows move as 128-bit words); plt_xy, table: 8-byte aligned. */
int wt_log_rows(const void* worm_rel, int worm_is_f32, const int32_t* cam_xywh, const int32_t* mic_xywh,
                const int32_t* plt_xy, int64_t n, int64_t first_frame, int cycle_frame_num, int imaging_frame_num,
                int frame_h, int frame_w, double* table, int32_t* crop_xywh, uint8_t* crop_legal, void* stream);

/* Derived columns of the analysed log (replaces DataAnalyzer.initialize, wtracker/eval/data_analyzer.py:54-107):
 *   table : f64 [n][17] rows of bboxes.csv in wt_log_rows' layout (one simulation, consecutive frames)
 *   out   : f64 [n][30] = frame, cycle, plt_x, plt_y, cam_x, cam_y, cam_w, cam_h, mic_x, mic_y, mic_w, mic_h, wrm_x, wrm_y,
 *           wrm_w, wrm_h, time, cycle_step, wrm_center_x, wrm_center_y, mic_center_x, mic_center_y, wrm_speed_x,
 *           wrm_speed_y, wrm_speed, worm_deviation_x, worm_deviation_y, worm_deviation, bbox_error, precise_error (NaN)
 * speed = centre difference over `period` rows / frame difference (NaN for the first `period` rows); every float column
 * rounded to 5 decimals exactly as DataFrame.round(5) does (rint(x * 1e5) / 1e5).                                       */
int wt_analysis_columns(const double* table, int64_t n, int period, int cycle_frame_num, double* out, void* stream);

/* Row masks of the analysed log (replaces DataAnalyzer.clean, wtracker/eval/data_analyzer.py:121-159, and the masks of
 * DataAnalyzer.calc_anomalies, :326-374) over wt_analysis_columns' f64 [n][30] table:
 *   moving   : u8 [n] phase of each row (0 imaging | 1 moving), may be NULL unless imaging_only
 *   h_bounds : HOST pointer to (x_min, y_min, x_max, y_max) or NULL: a row survives when its finite worm box lies inside,
 *              or when it has no prediction OR its worm box fails the x range and its microscope box lies inside (the
 *              reference's in-place mask aliasing, :140-148; golden tests/golden/reference_masks.npz)
 *   keep     : u8 [n] = the row survives imaging_only and bounds (trim_cycles needs the largest cycle of the SURVIVING
 *              rows, a reduction the host mirror does on the kept cycle column)
 *   anomaly  : u8 [n] bit 0 wrm_speed >= min_speed, 1 bbox_error >= min_bbox_error, 2 worm_deviation >= min_dist_error,
 *              3 wrm_w >= min_size, 4 wrm_h >= min_size, 5 no_preds and the worm box is not finite
 * Comparisons with NaN are false, as pandas evaluates them; thresholds may be +inf. */
int wt_analysis_masks(const double* table30, const uint8_t* moving, int64_t n, int imaging_only, const double* h_bounds,
                      int no_preds, double min_bbox_error, double min_dist_error, double min_speed, double min_size,
                      uint8_t* keep, uint8_t* anomaly, void* stream);

/* Segmentation-based tracking error (replaces ErrorCalculator.calculate_precise / calculate_segmentation,
 * wtracker/eval/error_calculator.py:19-161): for row i the fraction of the segmented worm — pixels of the
 * discretized worm box whose |frame - background| exceeds diff_thresh — that lies outside the microscope box.
 *   frames : u8 [n_frames][frame_h][frame_w] grey frames in HBM; frame_idx : i32 [n] frame of each row (the worm
 *            view the reference reads through `worm_reader` is that frame cropped at the discretized worm box)
 *   view_off : NULL, or i64 [n]: `frames` is then a packed buffer of the rows' own crops (row-major, w x h of the
 *            discretized worm box) and row i's crop starts at byte view_off[i]; frame_idx is not read
 *   background : u8 [frame_h][frame_w];  worm_xywh, mic_xywh : f64 [n][4]
 *   err : f64 [n]: 0 when nothing is segmented, NaN when the worm box is non-finite or empty after clipping.
 * (The reference writes the results of the legal rows to err[0..n_legal): that quirk is applied by the host mirror.) */
int wt_precise_error(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* frame_idx,
                     const int64_t* view_off, const uint8_t* background, const double* worm_xywh, const double* mic_xywh, double diff_thresh,
                     double* err, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* test-only helpers (allocate + synchronise; never called by the product path)               */
/* ------------------------------------------------------------------------------------------ */
/* Runs one conv through the tcgen05 kernel and the scalar validation kernel on seeded data and
 * returns the max abs difference (bf16 outputs) in *max_abs_diff; prints one line when verbose & 1; verbose & 2: the
 * source slice is the whole source buffer (what the stride-2 pixel-pair kernel needs). */
int wt_selftest_conv(int batch, int h, int w, int cin, int cout, int k, int stride, int act,
                     int with_residual, int out_f32, int verbose, double* max_abs_diff);
/* Chained form (wt_op.chain_w_off): conv (k, stride, SiLU) followed by a 1x1 conv cout -> cout (SiLU), once as ONE
 * chained launch and once as two tcgen05 launches through a bf16 buffer; the results must be identical.  */
int wt_selftest_conv_chain(int batch, int h, int w, int cin, int cout, int k, int stride, int verbose,
                           double* max_abs_diff);
/* Concat chain (wt_op.cat_buf): the exit of a C2f block (3x3 32 -> 32 + residual, then 1x1 96 -> 64 over the concat
 * buffer) as one launch vs two tcgen05 launches; the results must be identical. */
int wt_selftest_conv_cat(int batch, int h, int w, int verbose, double* max_abs_diff);

#ifdef __cplusplus
}
#endif
#endif /* WTRACKER_B200_H */
