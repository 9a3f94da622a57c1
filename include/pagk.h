/*
 * pagk.h -- C-ABI of the B200-native pixel-aware gyro-aided KLT hot path.
 *
 * This is the drop-in boundary for the path
 *     GyroAidedTracker::TrackFeatures()            (reference src/gyro_aided_tracker.cpp:344-426)
 *       -> IntegrateGyroMeasurements / SetRcl       (:511-587)
 *       -> GyroPredictFeatures                      (:118-256)
 *       -> PatchMatch::OpticalFlowMultiLevel        (reference src/patch_match.cpp:79-142)
 *            CreatePyramids (:61-76), the per-feature Gauss-Newton loop (:167-367),
 *            DistortPoints (:409-416), SetMatcher (:370-388)
 *       -> threshold filter                         (src/gyro_aided_tracker.cpp:289-336)
 *
 * The reference has no FFI layer: callers construct a C++ GyroAidedTracker over cv::Mat /
 * std::vector references (include/gyro_aided_tracker.h:109-126), call TrackFeatures() (:134) and
 * read public result vectors (:173-259).  This header restates exactly those inputs and outputs as
 * plain pointers and sizes.  include/pagk_tracker.hpp re-creates the two C++ classes on top of it.
 *
 * Everything here is POD; no C++ types, no exceptions cross the boundary.  All entry points return
 * PAGK_OK (0) or a negative pagk_status.  There is NO CPU fallback behind this ABI: every compute
 * entry point fails with PAGK_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef PAGK_H_
#define PAGK_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAGK_VERSION 100

typedef enum pagk_status {
  PAGK_OK = 0,
  PAGK_ERR_INVALID = -1,     /* bad argument (null pointer, size out of the handle's limits, ...) */
  PAGK_ERR_UNSUPPORTED = -2, /* eType 0 (cv::calcOpticalFlowPyrLK baseline), unknown eType, inverse mode */
  PAGK_ERR_NO_DEVICE = -3,   /* no CUDA device / driver: the product path never falls back to the CPU */
  PAGK_ERR_CUDA = -4,        /* a CUDA runtime call failed; see pagk_last_error() */
  PAGK_ERR_NOMEM = -5
} pagk_status;

/* GyroAidedTracker::eType, reference include/gyro_aided_tracker.h:55-63 */
enum {
  PAGK_OPENCV_OPTICAL_FLOW_PYR_LK = 0, /* not provided: returns PAGK_ERR_UNSUPPORTED */
  PAGK_GYRO_PREDICT = 1,
  PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED = 2,
  PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION = 3,
  PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION = 4, /* default */
  PAGK_IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION = 5,
  PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR = 6
};

/* GyroAidedTracker::ePredictMethod, reference include/gyro_aided_tracker.h:65-68 */
enum { PAGK_PIXEL_AWARE_PREDICTION = 1, PAGK_SINGLE_HOMOGRAPHY = 2 };

/* Capacity of one handle.  One handle lives on one device; handles on different devices are
 * independent (multi-GPU = one handle per device, streams sharded by the caller). */
typedef struct pagk_config {
  int device;         /* CUDA ordinal */
  int max_width;      /* largest image, pixels */
  int max_height;
  int max_keys;       /* keypoints per frame pair */
  int max_pairs;      /* frame pairs per pagk_track_batch call */
  int max_imu;        /* gyro samples per pair */
  int max_levels;     /* pyramid levels (reference hard-codes 3, src/gyro_aided_tracker.cpp:277) */
  int max_half_patch; /* reference default 5 (include/gyro_aided_tracker.h:115) */
} pagk_config;

/* Algorithm parameters.  Defaults (pagk_default_params) are the reference's hard-coded values:
 * iterations 10 / pyramids 3 / inverse false (src/gyro_aided_tracker.cpp:276-278),
 * lambda 1, alpha 0.5, max_distance 25 (src/patch_match.cpp:48-50), half_patch 5. */
typedef struct pagk_params {
  int e_type;         /* PAGK_* eType */
  int predict_method; /* PAGK_PIXEL_AWARE_PREDICTION | PAGK_SINGLE_HOMOGRAPHY */
  int half_patch;     /* 0 means 5, as in GyroAidedTracker::Initialize (:61) */
  int iterations;
  int pyramids;
  int inverse;        /* must be 0: the reference marks the inverse path "not support yet" (src/patch_match.cpp:220) */
  int calc_ncc;       /* PatchMatch bCalculateNCC_ (include/patch_match.h:49); the tracker never enables it */
  float lambda;
  float alpha;
  int max_distance;
} pagk_params;

/* Inputs of one frame pair = what the GyroAidedTracker ctor binds (src/gyro_aided_tracker.cpp:30-49). */
typedef struct pagk_pair_in {
  const uint8_t *img_ref;   /* mImgGrayRef: CV_8UC1, height rows of `pitch` bytes.  NULL on EVERY pair of a batch = stream
                             * continuation: the reference image of pair p is the current image of pair p of the previous
                             * batch on this handle (same size, levels; its pyramid is still on the device), as in the
                             * drivers' frame loop where curFrame becomes lastFrame.  Only img_cur crosses PCIe. */
  const uint8_t *img_cur;   /* mImgGrayCur */
  int width, height, pitch; /* all pairs of one batch must share width and height */
  int n_keys;               /* mN = mvKeysRef.size() */
  const float *keys_ref_un; /* mvKeysRefUn[i].pt, [n_keys][2] */
  const float *keys_ref;    /* mvKeysRef[i].pt, [n_keys][2]; read only by eType 5 (:266); may be NULL otherwise */
  int n_imu;                /* mvImuFromLastFrame.size() */
  const double *imu_t;      /* IMU::Point::t, [n_imu] */
  const float *imu_w;       /* IMU::Point::w, [n_imu][3] */
  double t_ref, t_cur;      /* mTimeStampRef, mTimeStamp */
  float bias_g[3];          /* mBias */
  float K[9];               /* mK row-major 3x3 */
  float dist[5];            /* mDistCoef k1 k2 p1 p2 k3 */
  int n_dist;               /* mDistCoef.total(): k3 is read only when 5 (:70) */
  float Rbc[9];             /* imuCalib.Tbc(0:3,0:3) row-major */
  const float *normalize_table; /* mNormalizeTable, CV_32FC2 height x width, or NULL (drivers pass cv::Mat()) */
  const float *Rcl_override;    /* NULL, or a 3x3 row-major rotation: SetRcl() instead of gyro integration */
} pagk_pair_in;

/* Outputs of one frame pair = the public result members of GyroAidedTracker
 * (include/gyro_aided_tracker.h:173-259).  Every pointer is caller-allocated for n_keys entries and
 * may be NULL when the caller does not want that vector. */
typedef struct pagk_pair_out {
  float *pt_predict_un;      /* mvPtPredictUn              [N][2] */
  float *pt_predict;         /* mvPtPredict                [N][2] */
  uint8_t *status;           /* mvStatus                   [N]    */
  float *pt_gyro_predict_un; /* mvPtGyroPredictUn          [N][2] */
  float *pt_gyro_predict;    /* mvPtGyroPredict            [N][2] */
  float *flows_predict_un;   /* mvFlowsPredictUn           [N][2] */
  float *affine;             /* mvAffineDeformationMatrix  [N][4] row-major 2x2; all-zero where the cv::Mat is empty */
  float *corner_flows;       /* mvvFlowsPredictCorners     [N][4][2] */
  float *pt_corners_un;      /* mvvPtPredictCornersUn      [N][4][2] */
  float *pt_corners;         /* mvvPtPredictCorners        [N][4][2] */
  float *pm_pt_un;           /* mvPtPredictAfterPatchMatchedUn [N][2] */
  float *pm_pt;              /* mvPtPredictAfterPatchMatched   [N][2] */
  uint8_t *pm_status;        /* mvStatusAfterPatchMatched  [N] */
  double *pixel_error;       /* mvPixelErrorsOfPatchMatched [N] */
  double *distance;          /* mvDistanceBetweenPredictedAndPatchMatched [N] */
  float *ncc;                /* mvNccAfterPatchMatched     [N] */
  int32_t *iters;            /* not in the reference: Gauss-Newton passes executed per feature, all levels */
  float Rcl[9];              /* mRcl */
  float KRKinv[9];           /* mKRKinv */
  int n_predict;             /* return value of TrackFeatures() */
  int64_t n_iterations;      /* sum of iters[] (the throughput unit: feature x iteration) */
  float t_gyro_predict;      /* mTimeCostGyroPredict,            seconds (device time of the batch / n_pairs) */
  float t_opt_flow;          /* mTimeCostOptFlow */
  float t_filter;            /* mTimeCostOptFlowResultFilterOut */
} pagk_pair_out;

/* Explicit inputs of PatchMatch (include/patch_match.h:44-49): what it reads from the tracker. */
typedef struct pagk_patch_match_in {
  const uint8_t *img_ref, *img_cur;
  int width, height, pitch;
  int n_keys;
  const float *keys_ref_un;   /* mpMatcher->mvKeysRefUn[i].pt */
  const float *pt_predict_un; /* mpMatcher->mvPtPredictUn */
  const uint8_t *status;      /* mpMatcher->mvStatus at construction (mvGyroPredictStatus) */
  const float *affine;        /* mpMatcher->mvAffineDeformationMatrix [N][4] */
  float K[9];
  float dist[5];
  int n_dist;
  int half_patch, iterations, pyramids;
  int has_gyro_predict_initial, inverse, consider_illumination, consider_affine_deformation,
      regularization_penalty, calc_ncc;
  float lambda, alpha;
  int max_distance;
} pagk_patch_match_in;

typedef struct pagk_handle pagk_handle;

void pagk_default_params(pagk_params *p);
int pagk_version(void);
const char *pagk_last_error(void);
int pagk_device_count(void);

int pagk_create(const pagk_config *cfg, pagk_handle **out);
void pagk_destroy(pagk_handle *h);

/* == GyroAidedTracker(...).TrackFeatures() for n_pairs independent frame pairs.
 * Host buffers in, host buffers out; H2D/D2H copies are inside the call (pinned memory is used
 * as-is, pageable memory goes through the driver's staging).  out[p].n_predict is -1 and the call
 * returns PAGK_ERR_UNSUPPORTED for an unsupported eType, as TrackFeatures() returns -1 (:415-418). */
int pagk_track_batch(pagk_handle *h, const pagk_params *prm, int n_pairs, const pagk_pair_in *in,
                     pagk_pair_out *out);

/* Asynchronous form of pagk_track_batch: pagk_submit_batch enqueues the copies and kernels of one batch on
 * the handle's stream and returns; pagk_wait_batch blocks until the results are in `out`.  A caller that
 * rotates over two or three handles overlaps the upload of the next batch with the kernels of the
 * current one.  `in`, `out` and every buffer they point to must stay valid until pagk_wait_batch. */
int pagk_submit_batch(pagk_handle *h, const pagk_params *prm, int n_pairs, const pagk_pair_in *in,
                      pagk_pair_out *out);
int pagk_wait_batch(pagk_handle *h);

/* The same work split so that inputs can stay resident in HBM between runs:
 *   pagk_upload_batch   H2D of images / keypoints / per-pair constants (gyro integrated on the host)
 *   pagk_run_resident   the kernels only (pyramid, predict, patch match, filter), asynchronous on
 *                       the handle's stream; may be called repeatedly on the same resident inputs
 *   pagk_download_batch D2H of the result vectors (synchronises) */
int pagk_upload_batch(pagk_handle *h, const pagk_params *prm, int n_pairs, const pagk_pair_in *in);
int pagk_run_resident(pagk_handle *h);
int pagk_download_batch(pagk_handle *h, int n_pairs, pagk_pair_out *out);
int pagk_synchronize(pagk_handle *h);
/* Stage clocks.  By default a run records a CUDA event between its kernels so that pagk_last_run_ms and the
 * t_gyro_predict / t_opt_flow / t_filter fields (mTimeCostGyroPredict, mTimeCostOptFlow,
 * mTimeCostOptFlowResultFilterOut, include/gyro_aided_tracker.h:226-231) are per stage.  on = 0 drops those events --
 * a throughput pipeline over several handles runs about 3 % faster without them -- and books the whole device time of
 * the batch on the patch alignment (t_opt_flow); the other two read 0. */
int pagk_set_stage_timing(pagk_handle *h, int on);
/* Throughput pipelines over several handles on one device (one handle per stream, all kept busy): n_handles > 1 makes a
 * launch of the patch-alignment kernel take 1/n_handles of every SM's CTA slots, so that the launches of the other handles
 * run beside it instead of behind it -- two batches interleaved on an SM halve the share of a launch's tail (config B: 7 % more
 * features/s over three rotating handles with n_handles = 2, 12 % over four with 3).  It is the caller's statement about its own pipeline: a launch that turns
 * out to be alone on the device runs on 1/n_handles of it.  Default 1 (a launch fills the device); results do not depend on it. */
int pagk_set_device_share(pagk_handle *h, int n_handles);
/* device time of the last pagk_run_resident, milliseconds (CUDA events on the handle's stream) */
int pagk_last_run_ms(pagk_handle *h, float *total_ms, float *pyramid_ms, float *predict_ms,
                     float *lk_ms, float *filter_ms);
/* the cudaStream_t of the handle, as void* (for callers that time with their own events) */
void *pagk_stream(pagk_handle *h);
/* make `h` issue its work on the stream of `other` (several resident batches, one in-order stream) */
int pagk_share_stream(pagk_handle *h, pagk_handle *other);
/* CUDA-event timing of the patch-alignment kernel over many runs: reset, run any number of times
 * (up to 1024 are recorded), then read the count and the summed device milliseconds (synchronises). */
int pagk_timing_reset(pagk_handle *h);
int pagk_timing_read(pagk_handle *h, int *n_runs, float *lk_ms_sum);
/* kernels launched by this handle since creation */
int64_t pagk_launch_count(pagk_handle *h);

/* == PatchMatch::CreatePyramids (src/patch_match.cpp:61-76) on the resident images.
 * pagk_get_pyramid_level copies level `level` of image `which` (0 = ref, 1 = cur) of pair `pair`
 * back to the host as a continuous cols x rows buffer. */
int pagk_build_pyramids(pagk_handle *h, int n_images, const uint8_t *const *imgs, int width, int height,
                        int pitch, int levels);
int pagk_pyramid_level_size(int width, int height, int level, int *cols, int *rows);
int pagk_get_pyramid_level(pagk_handle *h, int image, int level, uint8_t *dst, size_t dst_bytes);

/* == IntegrateGyroMeasurements + SetRcl (src/gyro_aided_tracker.cpp:511-587). Host arithmetic
 * (about ten 3x3 products per pair; uses libm sinf/cosf exactly like the reference). */
int pagk_integrate_gyro(const pagk_pair_in *in, float Rcl[9], float KRKinv[9]);

/* == GyroPredictFeatures (src/gyro_aided_tracker.cpp:118-185) for one pair, given Rcl. */
int pagk_gyro_predict(pagk_handle *h, const pagk_params *prm, const pagk_pair_in *in, pagk_pair_out *out);

/* == PatchMatch(...).OpticalFlowMultiLevel() (src/patch_match.cpp:79-142) for one pair with explicit
 * inputs; fills pm_pt_un, pm_pt, pm_status, pixel_error, distance, ncc, iters of `out`. */
int pagk_patch_match(pagk_handle *h, const pagk_patch_match_in *in, pagk_pair_out *out);

/* == GyroAidedTracker::GeometryValidation() (src/gyro_aided_tracker.cpp:429-508, CheckHomography :589-678,
 * CheckFundamental :680-768): the step both reference drivers run right after TrackFeatures(): two robust models of the
 * status-1 correspondences, the symmetric-transfer / epipolar chi-square scoring, the model choice
 * RH = SH / (SH + SF) > 0.45 and the outlier marking.
 *   estimate == 0  the caller supplies H21 / F21 -- what cv::findHomography(vPts1, vPts2, RANSAC, 3) and
 *                  cv::findFundamentalMat(vPts1, vPts2, FM_RANSAC, 3., 0.99) returned -- and everything after them is
 *                  bit-exact against the reference's own functions.
 *   estimate != 0  H21 / F21 are ignored: the device estimates both (csrc/pagk_ransac.h: RANSAC over 1024 hypotheses each
 *                  from a counter-based generator seeded with `seed`, refit on the inliers, same 3 px tests as OpenCV) and
 *                  the whole step stays on the device.  OpenCV's hypotheses come from OpenCV's own RNG, so these models
 *                  are not OpenCV's bit for bit: they are accepted statistically (tests/test_ransac.py).
 * keys_ref_un / pt_predict_un / status NULL: use the vectors of the handle's last run, resident on the device. */
typedef struct pagk_geometry_in {
  int n_keys;
  const float *keys_ref_un;   /* mvKeysRefUn[i].pt  [n_keys][2] */
  const float *pt_predict_un; /* mvPtPredictUn      [n_keys][2] */
  const uint8_t *status;      /* mvStatus           [n_keys]    */
  double H21[9];              /* row-major 3x3, CV_64F as cv::findHomography returns it */
  double F21[9];
  float sigma;                /* 1.0 in the reference (:447) */
  int estimate;               /* 0: H21 / F21 given; 1: estimated on the device */
  unsigned int seed;          /* of the hypothesis generator when estimate != 0 */
  int reserved;
} pagk_geometry_in;

typedef struct pagk_geometry_out {
  uint8_t *status;   /* mvStatus after the call [n_keys]; may be NULL */
  float score_H;     /* CheckHomography's score */
  float score_F;     /* CheckFundamental's score */
  int used_H;        /* 1: RH > 0.45, the homography's inliers were kept; 0: the fundamental matrix's */
  int n_candidates;  /* vPts1.size(): features with status 1 before the call */
  int n_inlier;      /* return value of GeometryValidation() (0 when n_candidates <= 8: nothing is validated) */
  double H21[9];     /* the models that were scored (the caller's, or the device's estimates) */
  double F21[9];
} pagk_geometry_out;

int pagk_geometry_validation(pagk_handle *h, int n_pairs, const pagk_geometry_in *in, pagk_geometry_out *out);

/* == Frame::SetPredictKeyPointsAndMask() (src/frame.cpp:115-153): the step that turns the surviving predictions of one
 * frame pair into the reference keypoints of the next one -- compaction in index order of mvKeys / mvKeysUn /
 * mvKeysNormal / mvPtIndexInLastFrame, the flow velocity of the last frame's features in the normalised plane, and the
 * occupancy mask (ones, with a 14 x 14 square zeroed around every survivor) that the detector top-up reads.
 * pt_predict / pt_predict_un / status NULL: the vectors of the handle's last run (or last pagk_geometry_validation),
 * resident on the device. */
typedef struct pagk_carry_in {
  int n_keys;                    /* mvStatus.size() */
  const float *pt_predict;       /* mvPtPredict    [n_keys][2] */
  const float *pt_predict_un;    /* mvPtPredictUn  [n_keys][2] */
  const uint8_t *status;         /* mvStatus       [n_keys]    */
  const float *keys_normal_last; /* mpLastFrame->mvKeysNormal[i].pt [n_keys][2] */
  float fx, fy, cx, cy;          /* mfx, mfy, mcx, mcy (mfx_inv = 1.0 / mfx as the Frame constructor computes it) */
  double t_cur, t_last;          /* mTimeStamp, mpLastFrame->mTimeStamp */
  int width, height;             /* mpCameraParams->width / height = size of mMask */
} pagk_carry_in;

typedef struct pagk_carry_out {
  int n_out;                 /* survivors = size of the four vectors below */
  float *keys;               /* mvKeys[k].pt        [n_keys][2] capacity */
  float *keys_un;            /* mvKeysUn[k].pt */
  float *keys_normal;        /* mvKeysNormal[k].pt */
  int32_t *index_in_last;    /* mvPtIndexInLastFrame[k] */
  float *flow_velocity_last; /* mpLastFrame->mvFlowVelocityInNormalPlane[k] */
  uint8_t *mask;             /* mMask, height x width, or NULL */
} pagk_carry_out;

int pagk_set_predict_keypoints_and_mask(pagk_handle *h, int n_pairs, const pagk_carry_in *in, pagk_carry_out *out);

/* == cv::FAST(image, keypoints, threshold, nonmaxSuppression, TYPE_9_16): the detector primitive under the keypoint top-up
 * of the drivers (ORBextractor::ComputeKeyPointsOctTree calls it per 30-pixel cell, src/ORBextractor.cc:833-840).
 * Keypoints come back in OpenCV's order (row-major), xy[k] = (x, y), response[k] = the corner score (0 without
 * non-maximum suppression, as in OpenCV).  mask (height x width, e.g. the one pagk_set_predict_keypoints_and_mask
 * returned) may be NULL; a keypoint on a zero mask byte is dropped, as ORBextractor::DetectFeatures does
 * (src/ORBextractor.cc:1200-1203).  At most max_out keypoints are written; *n_out is the number found. */
int pagk_fast_detect(pagk_handle *h, const uint8_t *img, int width, int height, int pitch, int threshold, int nonmax,
                     const uint8_t *mask, int max_out, float *xy, float *response, int *n_out);

/* == The detection half of ORBextractor::DetectFeatures for one pyramid level (src/ORBextractor.cc:1148-1205 with
 * ComputeKeyPointsOctTree :789-876), as Frame::DetectKeyPoints uses it to top keypoints up (src/frame.cpp:155-219): the
 * image inside a 16-pixel border is cut into cells of about 30 pixels, every cell runs cv::FAST with non-maximum suppression
 * at ini_th and, if that yields nothing, at min_th; keypoints on a zero mask byte are dropped (:1200-1203).  Keypoints come
 * back cell by cell (cells row-major, row-major inside a cell).  The thinning to nfeatures (DistributeOctTree) is
 * pagk_distribute_octtree / pagk_orb_detect_features below; with nfeatures above the candidate count it keeps every keypoint,
 * which is how the tests compare this call with the reference's own DetectFeatures. */
int pagk_orb_cell_detect(pagk_handle *h, const uint8_t *img, int width, int height, int pitch, int ini_th, int min_th,
                         const uint8_t *mask, int max_out, float *xy, float *response, int *n_out);

/* == ORBextractor::DistributeOctTree (src/ORBextractor.cc:563-787): thins candidates to about n_features keypoints, one per
 * leaf of a quadtree over [min_x, max_x) x [min_y, max_y), the strongest corner of each leaf.  xy are RELATIVE to (min_x,
 * min_y) as ComputeKeyPointsOctTree hands them over (:846-851).  Host code (a serial walk over a linked list of nodes; needs
 * no device and no handle).  Same splitting rule, list order and early exit as the reference; nodes of EQUAL size are split
 * in the order of their upper-left corner where the reference compares node addresses (:706-707).  out_index receives the
 * indices of the kept candidates in the reference's output order; *n_out their number (at most n). */
int pagk_distribute_octtree(int n, const float *xy, const float *response, int min_x, int max_x, int min_y, int max_y,
                            int n_features, int *out_index, int *n_out);

/* == ORBextractor(n_features, 1.2, 1, ini_th, min_th).DetectFeatures(img, mask) (src/ORBextractor.cc:1148-1205): the keypoint
 * top-up of Frame::DetectKeyPoints (src/frame.cpp:155-219) for one pyramid level: per-cell FAST on the device
 * (pagk_orb_cell_detect without a mask), DistributeOctTree down to n_features (pagk_distribute_octtree), then the mask filter
 * of :1200-1203, in that order as in the reference. */
int pagk_orb_detect_features(pagk_handle *h, const uint8_t *img, int width, int height, int pitch, int n_features, int ini_th,
                             int min_th, const uint8_t *mask, int max_out, float *xy, float *response, int *n_out);

/* == cv::remap(src, dst, map_x, map_y, cv::INTER_LINEAR) on CV_8UC1 with CV_32FC1 maps and the default constant (0) border:
 * the rectification both drivers run on every frame before tracking (Examples/Demo/RealSenseD435i.cpp:202,
 * Examples/ROS/.../feature_tracker.cpp:137; the maps come from cv::initUndistortRectifyMap, include/imu_types.h:63-65).
 * OpenCV's fixed-point arithmetic: coordinates rounded to 1/32 pixel, 15-bit weights, (sum + 2^14) >> 15.  Maps and dst are
 * dst_height x dst_width, contiguous. */
int pagk_remap_linear(pagk_handle *h, const uint8_t *src, int width, int height, int pitch, const float *map_x,
                      const float *map_y, int dst_width, int dst_height, uint8_t *dst);

/* The same rectification inside the pipeline: once maps are set, every img_ref / img_cur a batch brings is taken as the
 * DISTORTED image; it is uploaded as it is and remapped on the device straight into level 0 of its pyramid slot (nothing more
 * crosses PCIe, the maps stay on the device).  The maps must have the size of the images.  map_x = map_y = NULL switches the
 * rectification off again. */
int pagk_set_rectify_maps(pagk_handle *h, const float *map_x, const float *map_y, int width, int height);

#ifdef __cplusplus
}
#endif
#endif /* PAGK_H_ */
