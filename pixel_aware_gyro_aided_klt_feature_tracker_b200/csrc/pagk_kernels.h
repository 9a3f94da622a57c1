// pagk_kernels.h -- host-side launch wrappers of the kernels in pagk_kernels.cu
#pragma once
#include <cuda.h>  // CUtensorMap (the type only: the driver entry point is looked up at run time)
#include "pagk_device.cuh"

// one TMA tensor map per pyramid level: the level with its padding as (x, y), the image slot as z
struct PagkTmaLevels {
  CUtensorMap lv[PAGK_MAX_LEVELS];
};

int pagk_pyramid_fused_max_level();
// image z of the launch lives in slot z * z_stride (1: every slot, 2: every other one, i.e. only the current images)
int pagk_launch_pyramids(unsigned char *images, const PagkGeom &g, int n_images, int z_stride, cudaStream_t st,
                         long long *launches);
int pagk_launch_predict(const PagkPairConst *pcs, const float2 *keys_un, const float2 *keys, const PagkOutPtrs &out,
                        const PagkMode &mode, int max_keys, int n_max, int n_pairs, int width, int height,
                        const float *ntab, unsigned long long ntab_stride, cudaStream_t st, long long *launches);
size_t pagk_lk_smem_per_warp(int half);
int pagk_launch_lk(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                   const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs, cudaStream_t st,
                   long long *launches);
int pagk_launch_ncc(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                    const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs, cudaStream_t st,
                    long long *launches);
int pagk_launch_epilogue(const PagkPairConst *pcs, const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max,
                         int n_pairs, PagkPairResult *res, int do_filter, cudaStream_t st, long long *launches);
// GeometryValidation minus its RANSAC estimators (reference src/gyro_aided_tracker.cpp:429-508, 589-768)
int pagk_launch_geometry(const PagkGeoModel *models, const float2 *keys_un, const float2 *pred_un, unsigned char *status,
                         int max_keys, int n_pairs, PagkGeoResult *res, cudaStream_t st, long long *launches);
// the two RANSAC estimators of GeometryValidation on the device (pagk_ransac.h): fills H21 / H12 / F21 of models[p] where estimate[p]
int pagk_launch_ransac(const float2 *keys_un, const float2 *pred_un, const unsigned char *status, int max_keys, int n_pairs,
                       unsigned int seed, int *scratch_idx, unsigned char *scratch_in, PagkGeoModel *models,
                       const unsigned char *estimate, cudaStream_t st, long long *launches);
// Frame::SetPredictKeyPointsAndMask (reference src/frame.cpp:115-153); mask may be null
int pagk_launch_carry(const PagkCarryConst *cc, const float2 *pt_predict, const float2 *pt_predict_un, const unsigned char *status,
                      const float2 *normal_last, int max_keys, int n_pairs, float2 *keys, float2 *keys_un, float2 *keys_normal,
                      int *index_in_last, float2 *flow_last, int *n_out, unsigned char *mask, unsigned long long mask_stride,
                      cudaStream_t st, long long *launches);
// cv::FAST TYPE_9_16 (OpenCV features2d fast.cpp): scores -> keep flags + row counts -> row offsets -> ordered emit
int pagk_launch_fast(const unsigned char *img, int cols, int rows, int threshold, int nonmax, const unsigned char *mask,
                     unsigned short *score, unsigned char *keep, int *row_count, int *row_offset, int max_out, float2 *xy,
                     float *response, cudaStream_t st, long long *launches);
// per-cell FAST of ORBextractor::ComputeKeyPointsOctTree (one level): scores at min_th, per-cell threshold choice + keep flags,
// scan of the cell counts, ordered emit (cells row-major, row-major inside a cell)
struct PagkCellGrid { int n_cols, n_rows, w_cell, h_cell, min_x, min_y, max_x, max_y; };
int pagk_launch_orb_cells(const unsigned char *img, int cols, int rows, int ini_th, int min_th, const unsigned char *mask,
                          unsigned short *score, unsigned char *keep, int *cell_count, int *cell_offset, PagkCellGrid grid,
                          int max_out, float2 *xy, float *response, cudaStream_t st, long long *launches);
// cv::remap INTER_LINEAR, CV_8UC1, float maps, constant 0 border (OpenCV imgwarp.cpp remapBilinear)
int pagk_launch_remap(const unsigned char *src, int cols, int rows, const float *map_x, const float *map_y, int dcols, int drows,
                      unsigned char *dst, cudaStream_t st, long long *launches);
// the same, from a contiguous stack of raw images into level 0 of pyramid slots: image z -> slot z * z_stride + z_offset
int pagk_launch_remap_slots(const unsigned char *raw, unsigned char *images, const PagkGeom &g, const float *map_x, const float *map_y,
                            int n_images, int z_stride, int z_offset, cudaStream_t st, long long *launches);
int pagk_launch_count_status(const PagkPairConst *pcs, const PagkOutPtrs &out, int max_keys, int n_pairs,
                             PagkPairResult *res, cudaStream_t st, long long *launches);

// per device, from pagk_create after cudaSetDevice: dynamic shared memory opt-in of every kernel that needs it
int pagk_configure_kernels();
int pagk_lk_lanes_configure();

// production patch-alignment kernels (pagk_lk_lanes.cu): template staging pass + persistent CTAs, one lane per feature.
// tmpl: [max_pairs * max_keys][levels] template records of pagk_lk_lanes_record_bytes(half) bytes each.
bool pagk_lk_lanes_supported(const PagkMode &mode);
size_t pagk_lk_lanes_record_bytes(int half);
void pagk_lk_lanes_tma_box(int half, int *box_w, int *box_h);
int pagk_launch_lk_lanes(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                         const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs,
                         int *work_counters, int parity, int *progress, int epoch, int n_sms, unsigned char *tmpl,
                         const PagkTmaLevels *tmaps, cudaStream_t st, long long *launches, long long *prof, int share);
