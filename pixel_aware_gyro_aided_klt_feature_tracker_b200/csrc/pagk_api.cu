// pagk_api.cu -- the C-ABI of include/pagk.h: handle, HBM layout, copies, kernel sequencing.
//
// HBM layout of one handle (all allocated once in pagk_create, sized by pagk_config):
//   images   [2 * max_pairs] slots; a slot holds every pyramid level of one image, each level with a
//            4-byte aligned row pitch, its wrap column and guard row (PagkLevelGeom), level bases
//            256-byte aligned.  Slot 2p is the reference image of pair p, slot 2p+1 the current image.
//   tmpl     [levels][pairs * max_keys] template records of the alignment kernel, level-major (pagk_lk_lanes.cu)
//   handover [max_pairs * max_keys] 32-byte records: a feature's position and pass count from one level to the next
//   in       one block like the pinned staging block: PagkPairConst [max_pairs] (KRK^-1, r31..r33, intrinsics,
//            distortion, n_keys), then float2 [max_pairs][max_keys] x 2 (undistorted keypoints, raw keypoints)
//   results  structure of arrays [max_pairs][max_keys] (PagkOutPtrs), PagkPairResult [max_pairs]
// Host side: one pinned staging block mirrors keys+consts (upload) and one mirrors the results
// (download) so that a batch moves with a handful of large copies.
#include "../../include/pagk.h"
#include "pagk_host_math.h"
#include "pagk_kernels.h"
#include "pagk_octree.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *a = "", const char *b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) return fail(PAGK_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool make_geom(int width, int height, int levels, PagkGeom *g) {
  if (levels < 1 || levels > PAGK_MAX_LEVELS || width < 1 || height < 1) return false;
  g->levels = levels; g->width = width; g->height = height;
  size_t off = 0;
  int c = width, r = height;
  for (int l = 0; l < levels; ++l) {
    if (c < 1 || r < 1) return false;
    // rows are 16-byte aligned (TMA tensor maps need that of every stride): continuous when the width allows it, else
    // padded with room for the explicit wrap column
    const int pitch = (c % 16 == 0) ? c : (int)align_up((size_t)c + 1, 16);
    g->lv[l].cols = c; g->lv[l].rows = r; g->lv[l].pitch = pitch; g->lv[l].offset = (unsigned int)off;
    // rows + guard row, at least 32 rows, and 64 elements: a staged window may overhang the image (its content is not used)
    off += align_up((size_t)std::max(r + 1, 32) * pitch + 64, 256);
    c = (int)(c * 0.5); r = (int)(r * 0.5);  // cv::Size(cols * 0.5, rows * 0.5), reference src/patch_match.cpp:69
  }
  for (int l = levels; l < PAGK_MAX_LEVELS; ++l) g->lv[l] = PagkLevelGeom{0, 0, 0, 0};
  g->slot_bytes = off;
  return true;
}

enum OutKind {
  O_PT_PREDICT_UN, O_PT_PREDICT, O_PT_GYRO_UN, O_PT_GYRO, O_FLOWS, O_AFFINE, O_CFLOWS, O_CORNERS_UN, O_CORNERS,
  O_PM_UN, O_PM, O_STATUS, O_PM_STATUS, O_GYRO_STATUS, O_PIX_ERR, O_DIST, O_NCC, O_ITERS, O_COUNT
};
const size_t kOutElt[O_COUNT] = {8, 8, 8, 8, 8, 16, 32, 32, 32, 8, 8, 1, 1, 1, 8, 8, 4, 4};

}  // namespace

struct pagk_handle {
  pagk_config cfg;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  bool ran_stages = true;      // whether the last run recorded them
  int cur_pairs = 0;           // pairs whose CURRENT-image slot holds a finished pyramid (of geometry `geom`) from the last run
  bool cont = false;           // the uploaded batch continues streams: its reference images are the previous current ones
  int stage_timing = 1;        // CUDA events between the kernels of a run (pagk_set_stage_timing)
  int device_share = 1;        // handles that keep launches in flight on this device (pagk_set_device_share)
  cudaStream_t aux = nullptr;  // the gyro prediction runs here, beside the pyramid build (independent kernels)
  cudaEvent_t ev_aux = nullptr;
  std::vector<cudaEvent_t> tev;  // start/end event pairs around the LK kernel, pagk_timing_*
  int tev_used = -1;             // -1: timing off
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t slot_capacity = 0;
  unsigned char *d_images = nullptr;
  unsigned char *d_in = nullptr;  // [constants][undistorted keypoints][raw keypoints], laid out like the pinned h_in
  float2 *d_keys_un = nullptr, *d_keys = nullptr;
  PagkPairConst *d_pc = nullptr;
  PagkPairResult *d_res = nullptr;
  unsigned char *d_out = nullptr;
  size_t out_off[O_COUNT];
  size_t out_bytes = 0;
  float *d_ntab = nullptr;
  size_t ntab_stride = 0;  // floats per pair
  int *d_work = nullptr;   // work counters of the persistent LK kernel ([16..17])
  PagkTmaLevels tmaps;              // tensor maps of the u8 levels for the template kernel's tile loads ...
  PagkGeom tmaps_geom;              // ... of this geometry and patch size (rebuilt when either changes)
  int tmaps_half = 0;
  unsigned char *d_tmpl = nullptr;  // template records of the alignment kernel, [max_pairs * max_keys][max_levels]
  size_t tmpl_rec = 0;              // bytes per record the buffer was sized for (0: patch size without a lanes kernel)
  int lk_parity = 0;       // which of the two lanes-kernel counters the next launch uses
  // pagk_set_predict_keypoints_and_mask: one block, allocated on first use (consts, last normals, 5 result vectors, counts)
  unsigned char *h_aux = nullptr;   // pinned staging of the entry points around the path (grown on demand)
  size_t h_aux_bytes = 0;
  unsigned char *d_carry = nullptr;
  unsigned char *d_mask = nullptr;
  size_t mask_stride = 0;
  float *d_maps = nullptr;          // pagk_set_rectify_maps: map_x | map_y of map_w x map_h; batches then bring distorted images
  int map_w = 0, map_h = 0;
  unsigned char *d_raw = nullptr;   // the distorted images of a batch, as uploaded (2 * max_pairs images)
  unsigned char *d_remap = nullptr;  // pagk_remap_linear: source | map_x | map_y | destination (first use)
  unsigned char *d_fast = nullptr;  // pagk_fast_detect: image | mask | score | keep | row counts | row offsets (first use)
  unsigned char *d_fast_out = nullptr;
  int fast_out_cap = 0;
  PagkGeoModel *d_geo = nullptr;    // pagk_geometry_validation: models in, results out (allocated on first use)
  PagkGeoResult *d_geo_res = nullptr;
  unsigned char *d_ransac = nullptr;  // its estimators: per pair and model an index list and inlier flags, then the estimate flags
  int *d_progress = nullptr;  // lanes kernel, level-granular work items: per feature a 32-byte hand-over record (tag = epoch * 8 + levels finished)
  int lk_epoch = 0;           // launch number of the lanes kernel on this handle (values of earlier launches never match)
  int n_sms = 0;
  long long *d_dbg = nullptr;  // PAGK_LK_TIMELINE=<file>: clock64 timeline of CTA 0 of the LK kernel (developer aid)
  const char *dbg_path = nullptr;
  int lk_kernel = 0;       // PAGK_LK_KERNEL=generic: use the any-patch-size kernel instead of the lane-per-feature
                           // one (tests compare them)
  // pinned staging
  unsigned char *h_in = nullptr;   // keys_un | keys | consts
  size_t h_in_keys_un = 0, h_in_keys = 0, h_in_pc = 0, h_in_bytes = 0;
  unsigned char *h_out = nullptr;  // mirror of d_out
  PagkPairResult *h_res = nullptr;
  // state of the resident batch
  PagkGeom geom;
  PagkMode mode;
  int n_pairs = 0, n_max = 0, e_type = 0;
  bool resident = false, ran = false;
  std::vector<PagkPairConst> pcs;
  std::vector<float> rcl, krk;  // [n_pairs][9]
  long long launches = 0;
  float ms[5] = {0, 0, 0, 0, 0};
  // a download that has been enqueued but not finished (pagk_submit_batch / pagk_wait_batch)
  pagk_pair_out *pending_out = nullptr;
  int pending_n = 0;
  std::vector<char> pending_staged;

  PagkOutPtrs outs() const {
    PagkOutPtrs o;
    o.pt_predict_un = (float2 *)(d_out + out_off[O_PT_PREDICT_UN]); o.pt_predict = (float2 *)(d_out + out_off[O_PT_PREDICT]);
    o.pt_gyro_un = (float2 *)(d_out + out_off[O_PT_GYRO_UN]); o.pt_gyro = (float2 *)(d_out + out_off[O_PT_GYRO]);
    o.flows = (float2 *)(d_out + out_off[O_FLOWS]); o.affine = (float4 *)(d_out + out_off[O_AFFINE]);
    o.cflows = (float2 *)(d_out + out_off[O_CFLOWS]); o.corners_un = (float2 *)(d_out + out_off[O_CORNERS_UN]);
    o.corners = (float2 *)(d_out + out_off[O_CORNERS]); o.pm_un = (float2 *)(d_out + out_off[O_PM_UN]);
    o.pm = (float2 *)(d_out + out_off[O_PM]); o.status = d_out + out_off[O_STATUS];
    o.pm_status = d_out + out_off[O_PM_STATUS]; o.gyro_status = d_out + out_off[O_GYRO_STATUS];
    o.pix_err = (double *)(d_out + out_off[O_PIX_ERR]); o.dist = (double *)(d_out + out_off[O_DIST]);
    o.ncc = (float *)(d_out + out_off[O_NCC]); o.iters = (int *)(d_out + out_off[O_ITERS]);
    return o;
  }
};

namespace {

int mode_from_etype(int e_type, PagkMode *m) {
  switch (e_type) {  // reference src/gyro_aided_tracker.cpp:384-414
    case PAGK_IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION: m->gyro_init = 0; m->illum = 1; m->affine = 1; m->regular = 0; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED: m->gyro_init = 1; m->illum = 0; m->affine = 0; m->regular = 0; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION: m->gyro_init = 1; m->illum = 1; m->affine = 0; m->regular = 0; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION: m->gyro_init = 1; m->illum = 1; m->affine = 1; m->regular = 0; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR: m->gyro_init = 1; m->illum = 1; m->affine = 1; m->regular = 1; return 0;
    case PAGK_GYRO_PREDICT: m->gyro_init = 1; m->illum = 0; m->affine = 0; m->regular = 0; return 0;
    default: return -1;
  }
}

void fill_mode_common(PagkMode *m, int half, int iterations, int levels, int calc_ncc, int predict_method, float lambda,
                      float alpha, int max_distance) {
  m->half = half; m->iterations = iterations; m->levels = levels; m->calc_ncc = calc_ncc;
  m->predict_method = predict_method; m->lambda = lambda; m->alpha = alpha;
  // PatchMatch ctor, reference src/patch_match.cpp:51,57
  m->inv_log_max_dist = (float)(1.0 / (double)logf(alpha * (float)max_distance + 1.0f));
  m->win_size_inv = (double)(1.0f / (2.0f * half + 1.0f) / (2.0f * half + 1.0f));
  m->bb_inv = pagk_host::bbt_inverse_diag(half);
}

// GyroAidedTracker::Initialize caches (:64-70) + IntegrateGyroMeasurements/SetRcl
void pair_const(const pagk_pair_in &in, PagkPairConst *pc, float Rcl[9], float KRK[9]) {
  using namespace pagk_host;
  const Mat3 K = from_array(in.K);
  Mat3 R;
  if (in.Rcl_override) R = from_array(in.Rcl_override);
  else R = integrate(in.n_imu, in.imu_t, in.imu_w, in.t_ref, in.t_cur, in.bias_g, from_array(in.Rbc));
  const Mat3 M = krkinv(K, R);
  std::memcpy(Rcl, R.v, 9 * sizeof(float));
  std::memcpy(KRK, M.v, 9 * sizeof(float));
  std::memcpy(pc->M, M.v, 9 * sizeof(float));
  pc->r31 = R.v[2][0]; pc->r32 = R.v[2][1]; pc->r33 = R.v[2][2];
  pc->fx = in.K[0]; pc->fy = in.K[4]; pc->cx = in.K[2]; pc->cy = in.K[5];
  pc->fx_inv = (float)(1.0 / pc->fx); pc->fy_inv = (float)(1.0 / pc->fy);
  pc->k1 = in.dist[0]; pc->k2 = in.dist[1]; pc->p1 = in.dist[2]; pc->p2 = in.dist[3];
  pc->k3 = (in.n_dist == 5) ? in.dist[4] : 0.0f;
  pc->n_keys = in.n_keys;
  pc->has_table = in.normalize_table ? 1 : 0;
}

// with rectification maps set: the images go to the raw stack as they are and are remapped into level 0 of their slots
int upload_images_rectify(pagk_handle *h, int n_pairs, const uint8_t *const *refs, const uint8_t *const *curs, int width, int height,
                          const int *pitches) {
  const PagkGeom &g = h->geom;
  const size_t img_bytes = (size_t)width * height;
  if (h->map_w != width || h->map_h != height) return fail(PAGK_ERR_INVALID, "rectification maps do not have the size of the images");
  if (!h->d_raw) CU(cudaMalloc(&h->d_raw, (size_t)2 * h->cfg.max_pairs * h->cfg.max_width * h->cfg.max_height));
  const bool cont = refs == nullptr;
  const int per = cont ? 1 : 2;
  for (int p = 0; p < n_pairs; ++p) {
    if (!cont) CU(cudaMemcpy2DAsync(h->d_raw + (size_t)(2 * p) * img_bytes, width, refs[p], pitches[p], width, height, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpy2DAsync(h->d_raw + (size_t)(per * p + per - 1) * img_bytes, width, curs[p], pitches[p], width, height, cudaMemcpyHostToDevice, h->stream));
  }
  CU((cudaError_t)pagk_launch_remap_slots(h->d_raw, h->d_images, g, h->d_maps, h->d_maps + img_bytes, per * n_pairs, cont ? 2 : 1,
                                          cont ? 1 : 0, h->stream, &h->launches));
  return PAGK_OK;
}

int upload_images(pagk_handle *h, int n_pairs, const uint8_t *const *refs, const uint8_t *const *curs, int width,
                  int height, const int *pitches) {
  if (h->d_maps) return upload_images_rectify(h, n_pairs, refs, curs, width, height, pitches);
  const PagkGeom &g = h->geom;
  const size_t img_bytes = (size_t)width * height;
  const size_t dp = (size_t)g.lv[0].pitch;
  bool contiguous = dp == (size_t)width;  // level 0 is stored continuous on the device as well
  for (int p = 0; p < n_pairs && contiguous; ++p) {
    if (pitches[p] != width) contiguous = false;
    if (curs[p] != refs[p] + img_bytes) contiguous = false;
    if (p + 1 < n_pairs && refs[p + 1] != curs[p] + img_bytes) contiguous = false;
  }
  if (contiguous && n_pairs > 0) {  // [n_pairs][2][H][W] in one strided copy
    CU(cudaMemcpy2DAsync(h->d_images + g.lv[0].offset, g.slot_bytes, refs[0], img_bytes, img_bytes, (size_t)2 * n_pairs,
                         cudaMemcpyHostToDevice, h->stream));
    return PAGK_OK;
  }
  for (int p = 0; p < n_pairs; ++p) {
    CU(cudaMemcpy2DAsync(h->d_images + (size_t)(2 * p) * g.slot_bytes + g.lv[0].offset, dp, refs[p], pitches[p], width,
                         height, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpy2DAsync(h->d_images + (size_t)(2 * p + 1) * g.slot_bytes + g.lv[0].offset, dp, curs[p], pitches[p],
                         width, height, cudaMemcpyHostToDevice, h->stream));
  }
  return PAGK_OK;
}

// Stream continuation: the reference image of pair p is the current image of pair p of the previous batch, whose pyramid
// is still on the device.  One device-to-device copy moves every current slot (all levels) into its reference slot; only
// the new current images cross PCIe.
int upload_images_continue(pagk_handle *h, int n_pairs, const uint8_t *const *curs, int width, int height, const int *pitches) {
  const PagkGeom &g = h->geom;
  const size_t img_bytes = (size_t)width * height;
  CU(cudaMemcpy2DAsync(h->d_images, 2 * g.slot_bytes, h->d_images + g.slot_bytes, 2 * g.slot_bytes, g.slot_bytes, (size_t)n_pairs,
                       cudaMemcpyDeviceToDevice, h->stream));
  if (h->d_maps) return upload_images_rectify(h, n_pairs, nullptr, curs, width, height, pitches);
  const size_t dp = (size_t)g.lv[0].pitch;
  bool contiguous = dp == (size_t)width;
  for (int p = 0; p < n_pairs && contiguous; ++p) {
    if (pitches[p] != width) contiguous = false;
    if (p + 1 < n_pairs && curs[p + 1] != curs[p] + img_bytes) contiguous = false;
  }
  if (contiguous && n_pairs > 0) {
    CU(cudaMemcpy2DAsync(h->d_images + g.slot_bytes + g.lv[0].offset, 2 * g.slot_bytes, curs[0], img_bytes, img_bytes, (size_t)n_pairs,
                         cudaMemcpyHostToDevice, h->stream));
    return PAGK_OK;
  }
  for (int p = 0; p < n_pairs; ++p)
    CU(cudaMemcpy2DAsync(h->d_images + (size_t)(2 * p + 1) * g.slot_bytes + g.lv[0].offset, dp, curs[p], pitches[p], width,
                         height, cudaMemcpyHostToDevice, h->stream));
  return PAGK_OK;
}

// Validates a batch against the handle's capacity and computes its geometry into *g.  The handle's own geometry
// (h->geom: the layout of the pyramids that ARE on the device) only changes when the caller commits *g, i.e. when
// images of that layout are on their way into the slots.
int check_batch(pagk_handle *h, int n_pairs, int width, int height, int levels, int half, PagkGeom *g) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  if (n_pairs < 0 || n_pairs > h->cfg.max_pairs) return fail(PAGK_ERR_INVALID, "n_pairs exceeds pagk_config.max_pairs");
  if (levels < 1 || levels > h->cfg.max_levels) return fail(PAGK_ERR_INVALID, "pyramids exceeds pagk_config.max_levels");
  if (half < 1 || half > h->cfg.max_half_patch) return fail(PAGK_ERR_INVALID, "half_patch exceeds pagk_config.max_half_patch");
  // buffers beside the image slots (normalize tables, raw images of the rectification) are sized by max_width * max_height
  if (width > h->cfg.max_width || height > h->cfg.max_height) return fail(PAGK_ERR_INVALID, "image exceeds pagk_config.max_width/max_height");
  if (!make_geom(width, height, levels, g)) return fail(PAGK_ERR_INVALID, "image too small for the requested pyramid");
  if (g->slot_bytes > h->slot_capacity) return fail(PAGK_ERR_INVALID, "image exceeds pagk_config.max_width/max_height");
  if (pagk_lk_smem_per_warp(half) > 227 * 1024) return fail(PAGK_ERR_INVALID, "half_patch too large for shared memory");
  return PAGK_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// One tensor map per level: x = the padded row (pitch bytes), y = the allocated rows, z = the image slot.  Boxes that
// overhang the tensor are filled with zeros, so the kernels never clamp a tile's origin.
int build_tmaps(pagk_handle *h, int half) {
  if (h->tmaps_half == half && std::memcmp(&h->tmaps_geom, &h->geom, sizeof(PagkGeom)) == 0) return PAGK_OK;
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(PAGK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  int bw = 0, bh = 0;
  pagk_lk_lanes_tma_box(half, &bw, &bh);
  std::memset(&h->tmaps, 0, sizeof(h->tmaps));
  for (int l = 0; l < h->geom.levels; ++l) {
    const PagkLevelGeom &L = h->geom.lv[l];
    const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)std::max(L.rows + 1, 32), (cuuint64_t)(2 * h->cfg.max_pairs)};
    const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)h->geom.slot_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(&h->tmaps.lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, h->d_images + L.offset, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { h->tmaps_half = 0; return fail(PAGK_ERR_CUDA, "cuTensorMapEncodeTiled failed"); }
  }
  h->tmaps_geom = h->geom; h->tmaps_half = half;
  return PAGK_OK;
}

int launch_lk_kernel(pagk_handle *h, const PagkOutPtrs &o, const PagkMode &m, int n_max, int n_pairs) {
  if (h->lk_kernel == 0 && pagk_lk_lanes_supported(m) && h->d_tmpl && pagk_lk_lanes_record_bytes(m.half) <= h->tmpl_rec) {
    if (build_tmaps(h, m.half) != PAGK_OK) return (int)cudaErrorUnknown;
    // PAGK_LK_TIMELINE=<file> with a -DPAGK_LANES_PROF build: per-warp phase cycles (developer aid)
    if (h->d_dbg) CU(cudaMemsetAsync(h->d_dbg, 0, 2048 * 16 * sizeof(long long), h->stream));
    if (++h->lk_epoch >= 0x0fffffff) {  // the epoch is about to repeat: forget every progress word written so far
      CU(cudaMemsetAsync(h->d_progress, 0, (size_t)h->cfg.max_pairs * h->cfg.max_keys * 8 * sizeof(int), h->stream));
      h->lk_epoch = 1;
    }
    const int rc = pagk_launch_lk_lanes(h->d_images, h->geom, h->d_pc, h->d_keys_un, o, m, h->cfg.max_keys, n_max, n_pairs,
                                        h->d_work + 16, h->lk_parity, h->d_progress, h->lk_epoch, h->n_sms, h->d_tmpl, &h->tmaps,
                                        h->stream, &h->launches, h->d_dbg, h->device_share);
    if (rc == 0 && n_max > 0 && n_pairs > 0) h->lk_parity ^= 1;
    if (h->d_dbg && rc == 0) {
      std::vector<long long> tl(2048 * 16);
      CU(cudaMemcpyAsync(tl.data(), h->d_dbg, tl.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      if (FILE *f = fopen(h->dbg_path, "w")) {
        for (size_t w = 0; w < tl.size() / 16; ++w) {
          if (!tl[w * 16 + 6]) continue;
          for (int k = 0; k < 16; ++k) fprintf(f, "%lld%c", tl[w * 16 + k], k == 15 ? '\n' : ' ');
        }
        fclose(f);
      }
    }
    return rc;
  }
  return pagk_launch_lk(h->d_images, h->geom, h->d_pc, h->d_keys_un, o, m, h->cfg.max_keys, n_max, n_pairs, h->stream,
                        &h->launches);
}

// patch alignment, then PatchMatch::NCC when the caller asked for it (bCalculateNCC_)
int launch_lk(pagk_handle *h, const PagkOutPtrs &o, const PagkMode &m, int n_max, int n_pairs) {
  const int rc = launch_lk_kernel(h, o, m, n_max, n_pairs);
  if (rc != 0 || !m.calc_ncc) return rc;
  return pagk_launch_ncc(h->d_images, h->geom, h->d_pc, h->d_keys_un, o, m, h->cfg.max_keys, n_max, n_pairs, h->stream,
                         &h->launches);
}

template <class T>
void scatter(const pagk_handle *h, OutKind k, int n_pairs, const std::vector<int> &nk, T *const *dst) {
  for (int p = 0; p < n_pairs; ++p)
    if (dst[p]) std::memcpy(dst[p], h->h_out + h->out_off[k] + (size_t)p * h->cfg.max_keys * kOutElt[k], (size_t)nk[p] * kOutElt[k]);
}

}  // namespace

extern "C" {

void pagk_default_params(pagk_params *p) {
  if (!p) return;
  p->e_type = PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION;
  p->predict_method = PAGK_PIXEL_AWARE_PREDICTION;
  p->half_patch = 5; p->iterations = 10; p->pyramids = 3; p->inverse = 0; p->calc_ncc = 0;
  p->lambda = 1.0f; p->alpha = 0.5f; p->max_distance = 25;
}

int pagk_version(void) { return PAGK_VERSION; }
const char *pagk_last_error(void) { return g_err; }

int pagk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int pagk_create(const pagk_config *cfg, pagk_handle **out) {
  if (!cfg || !out) return fail(PAGK_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->max_width < 1 || cfg->max_height < 1 || cfg->max_keys < 1 || cfg->max_pairs < 1 || cfg->max_levels < 1 ||
      cfg->max_levels > PAGK_MAX_LEVELS || cfg->max_half_patch < 1)
    return fail(PAGK_ERR_INVALID, "bad pagk_config");
  // the pair (and the image, 2 * pair) is a grid dimension of the per-pair kernels: 65535 at most
  if (cfg->max_pairs > 32767) return fail(PAGK_ERR_INVALID, "pagk_config.max_pairs exceeds 32767");
  if ((long long)cfg->max_pairs * cfg->max_keys > 0x3fffffffLL) return fail(PAGK_ERR_INVALID, "pagk_config: max_pairs * max_keys exceeds 2^30");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(PAGK_ERR_NO_DEVICE, "no CUDA device: the pagk hot path has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(PAGK_ERR_INVALID, "pagk_config.device out of range");
  CU(cudaSetDevice(cfg->device));
  pagk_handle *h = new (std::nothrow) pagk_handle();
  if (!h) return fail(PAGK_ERR_NOMEM, "out of host memory");
  h->cfg = *cfg;
  PagkGeom g;
  // capacity: the requested number of levels on the largest image, every level present
  int lv = cfg->max_levels;
  while (lv > 1 && !make_geom(cfg->max_width, cfg->max_height, lv, &g)) --lv;
  make_geom(cfg->max_width, cfg->max_height, lv, &g);
  h->slot_capacity = g.slot_bytes + 512 * (size_t)cfg->max_levels;
  const size_t NK = (size_t)cfg->max_pairs * cfg->max_keys;
  size_t off = 0;
  for (int k = 0; k < O_COUNT; ++k) { h->out_off[k] = off; off += align_up(NK * kOutElt[k], 256); }
  h->out_bytes = off;
  // per-pair constants, undistorted keypoints, raw keypoints: one block on the host and the same block on the device, so that
  // a batch's constants and keypoints go up in one copy (the raw keypoints, last, only in the mode that reads them)
  h->h_in_pc = 0;
  h->h_in_keys_un = align_up((size_t)cfg->max_pairs * sizeof(PagkPairConst), 256);
  h->h_in_keys = h->h_in_keys_un + align_up(NK * sizeof(float2), 256);
  h->h_in_bytes = h->h_in_keys + align_up(NK * sizeof(float2), 256);
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; return r == cudaSuccess; };
  ok(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  ok(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
  ok(cudaEventCreateWithFlags(&h->ev_aux, cudaEventDisableTiming));
  for (int i = 0; i < 6; ++i) ok(cudaEventCreate(&h->ev[i]));
  ok(cudaMalloc(&h->d_images, h->slot_capacity * 2 * (size_t)cfg->max_pairs));
  ok(cudaMalloc(&h->d_in, h->h_in_bytes));
  if (e == cudaSuccess) {
    h->d_pc = reinterpret_cast<PagkPairConst *>(h->d_in + h->h_in_pc);
    h->d_keys_un = reinterpret_cast<float2 *>(h->d_in + h->h_in_keys_un);
    h->d_keys = reinterpret_cast<float2 *>(h->d_in + h->h_in_keys);
  }
  ok(cudaMalloc(&h->d_res, (size_t)cfg->max_pairs * sizeof(PagkPairResult)));
  ok(cudaMalloc(&h->d_out, h->out_bytes));
  ok(cudaMalloc(&h->d_work, 256));
  if (e == cudaSuccess) ok(cudaMemset(h->d_work, 0, 256));
  ok(cudaMalloc(&h->d_progress, NK * 8 * sizeof(int)));
  if (e == cudaSuccess) ok(cudaMemset(h->d_progress, 0, NK * 8 * sizeof(int)));
  // test hook: start the launch counter of the lanes kernel near its wrap (the hand-over tags are epoch * 8 + levels finished)
  if (const char *ep = getenv("PAGK_DEBUG_LK_EPOCH")) h->lk_epoch = atoi(ep);
  ok(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, cfg->device));
  // function attributes belong to the device: set them for this handle's device (no process-wide "done" flag)
  ok((cudaError_t)pagk_configure_kernels());
  ok((cudaError_t)pagk_lk_lanes_configure());
  // template records of the alignment kernel for the largest patch this handle may be asked for (5 and 10 have a lanes kernel)
  h->tmpl_rec = cfg->max_half_patch >= 10 ? pagk_lk_lanes_record_bytes(10) : cfg->max_half_patch >= 5 ? pagk_lk_lanes_record_bytes(5) : 0;
  if (h->tmpl_rec) ok(cudaMalloc(&h->d_tmpl, NK * (size_t)cfg->max_levels * h->tmpl_rec));
  {
    const char *k = getenv("PAGK_LK_KERNEL");
    h->lk_kernel = (k && std::strcmp(k, "generic") == 0) ? 2 : 0;
    h->dbg_path = getenv("PAGK_LK_TIMELINE");
    if (h->dbg_path) { ok(cudaMalloc(&h->d_dbg, 2048 * 16 * sizeof(long long))); }
  }
  ok(cudaMallocHost(&h->h_in, h->h_in_bytes));
  ok(cudaMallocHost(&h->h_out, h->out_bytes));
  ok(cudaMallocHost(&h->h_res, (size_t)cfg->max_pairs * sizeof(PagkPairResult)));
  if (e == cudaSuccess) ok(cudaMemsetAsync(h->d_out, 0, h->out_bytes, h->stream));
  if (e == cudaSuccess) ok(cudaStreamSynchronize(h->stream));
  if (e != cudaSuccess) {
    const int code = (e == cudaErrorMemoryAllocation) ? PAGK_ERR_NOMEM : PAGK_ERR_CUDA;
    fail(code, "pagk_create: %s", cudaGetErrorString(e));
    pagk_destroy(h);
    return code;
  }
  *out = h;
  return PAGK_OK;
}

void pagk_destroy(pagk_handle *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  // a borrowed stream (pagk_share_stream) may already be gone with its owner: synchronise the device instead
  if (h->stream && h->own_stream) cudaStreamSynchronize(h->stream);
  else cudaDeviceSynchronize();
  cudaFree(h->d_images); cudaFree(h->d_in); cudaFree(h->d_res);
  cudaFree(h->d_out); cudaFree(h->d_ntab); cudaFree(h->d_work); cudaFree(h->d_tmpl); cudaFree(h->d_progress); cudaFree(h->d_geo); cudaFree(h->d_geo_res); cudaFree(h->d_ransac); cudaFree(h->d_carry); cudaFree(h->d_mask); cudaFreeHost(h->h_aux); cudaFree(h->d_fast); cudaFree(h->d_fast_out); cudaFree(h->d_remap); cudaFree(h->d_maps); cudaFree(h->d_raw); cudaFree(h->d_dbg);
  cudaFreeHost(h->h_in); cudaFreeHost(h->h_out); cudaFreeHost(h->h_res);
  for (int i = 0; i < 6; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
  if (h->aux) { cudaStreamSynchronize(h->aux); cudaStreamDestroy(h->aux); }
  if (h->ev_aux) cudaEventDestroy(h->ev_aux);
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
}

int pagk_upload_batch(pagk_handle *h, const pagk_params *prm, int n_pairs, const pagk_pair_in *in) {
  if (!h || !prm || (n_pairs > 0 && !in)) return fail(PAGK_ERR_INVALID, "null argument");
  h->resident = false; h->ran = false;
  PagkMode m;
  std::memset(&m, 0, sizeof(m));
  if (prm->e_type == PAGK_OPENCV_OPTICAL_FLOW_PYR_LK)
    return fail(PAGK_ERR_UNSUPPORTED, "eType 0 (cv::calcOpticalFlowPyrLK comparison baseline) is not part of the hot path");
  if (mode_from_etype(prm->e_type, &m) != 0) return fail(PAGK_ERR_UNSUPPORTED, "Unsupport type!!! return -1;");
  if (prm->inverse) return fail(PAGK_ERR_UNSUPPORTED, "inverse mode is \"not support yet\" in the reference (src/patch_match.cpp:220)");
  if (prm->predict_method != PAGK_PIXEL_AWARE_PREDICTION && prm->predict_method != PAGK_SINGLE_HOMOGRAPHY)
    return fail(PAGK_ERR_INVALID, "bad predict_method");
  const int half = prm->half_patch == 0 ? 5 : prm->half_patch;  // Initialize(), :61
  if (n_pairs == 0) { h->n_pairs = 0; h->resident = true; return PAGK_OK; }
  const int W = in[0].width, H = in[0].height;
  // img_ref == NULL on every pair: the batch continues the streams of the previous batch of this handle
  int n_null = 0;
  for (int p = 0; p < n_pairs; ++p) n_null += in[p].img_ref ? 0 : 1;
  const bool cont = n_null == n_pairs;
  if (n_null != 0 && !cont) return fail(PAGK_ERR_INVALID, "img_ref: null on every pair of the batch (stream continuation) or on none");
  PagkGeom geom;
  int rc = check_batch(h, n_pairs, W, H, prm->pyramids, half, &geom);
  if (rc != PAGK_OK) return rc;
  if (cont && (h->cur_pairs < n_pairs || geom.width != h->geom.width || geom.height != h->geom.height || geom.levels != h->geom.levels)) {
    h->cur_pairs = 0;
    return fail(PAGK_ERR_INVALID, "stream continuation without a previous batch of the same geometry and at least as many pairs");
  }
  if (prm->iterations < 0) return fail(PAGK_ERR_INVALID, "iterations < 0");
  fill_mode_common(&m, half, prm->iterations, prm->pyramids, prm->calc_ncc, prm->predict_method, prm->lambda, prm->alpha,
                   prm->max_distance);
  CU(cudaSetDevice(h->cfg.device));
  h->pcs.assign(n_pairs, PagkPairConst());
  h->rcl.assign((size_t)n_pairs * 9, 0.f); h->krk.assign((size_t)n_pairs * 9, 0.f);
  int n_max = 0;
  bool any_table = false;
  std::vector<const uint8_t *> refs(n_pairs), curs(n_pairs);
  std::vector<int> pitches(n_pairs);
  CU(cudaStreamSynchronize(h->stream));  // the pinned staging block may still feed an earlier copy
  float2 *s_un = (float2 *)(h->h_in + h->h_in_keys_un), *s_k = (float2 *)(h->h_in + h->h_in_keys);
  for (int p = 0; p < n_pairs; ++p) {
    const pagk_pair_in &q = in[p];
    if (q.width != W || q.height != H) return fail(PAGK_ERR_INVALID, "all pairs of a batch must share width and height");
    if ((!cont && !q.img_ref) || !q.img_cur || q.pitch < W) return fail(PAGK_ERR_INVALID, "bad image pointer or pitch");
    if (q.n_keys < 0 || q.n_keys > h->cfg.max_keys) return fail(PAGK_ERR_INVALID, "n_keys exceeds pagk_config.max_keys");
    if (q.n_keys > 0 && !q.keys_ref_un) return fail(PAGK_ERR_INVALID, "keys_ref_un is null");
    if (!m.gyro_init && q.n_keys > 0 && !q.keys_ref) return fail(PAGK_ERR_INVALID, "eType 5 reads keys_ref");
    if (!q.Rcl_override && (q.n_imu < 0 || (q.n_imu > 0 && (!q.imu_t || !q.imu_w)))) return fail(PAGK_ERR_INVALID, "bad imu arrays");
    pair_const(q, &h->pcs[p], &h->rcl[(size_t)p * 9], &h->krk[(size_t)p * 9]);
    n_max = q.n_keys > n_max ? q.n_keys : n_max;
    any_table |= (q.normalize_table != nullptr);
    refs[p] = q.img_ref; curs[p] = q.img_cur; pitches[p] = q.pitch;
    std::memcpy(s_un + (size_t)p * h->cfg.max_keys, q.keys_ref_un, (size_t)q.n_keys * sizeof(float2));
    if (q.keys_ref) std::memcpy(s_k + (size_t)p * h->cfg.max_keys, q.keys_ref, (size_t)q.n_keys * sizeof(float2));
  }
  std::memcpy(h->h_in + h->h_in_pc, h->pcs.data(), (size_t)n_pairs * sizeof(PagkPairConst));
  const size_t key_bytes = (size_t)n_pairs * h->cfg.max_keys * sizeof(float2);
  // constants and undistorted keypoints in one copy (the gap between them is the constants of the pairs not in this batch:
  // at most 256 bytes per unused pair)
  const size_t pc_bytes = (size_t)n_pairs * sizeof(PagkPairConst);
  if (h->h_in_keys_un - pc_bytes <= 65536) {
    CU(cudaMemcpyAsync(h->d_in, h->h_in, h->h_in_keys_un + key_bytes, cudaMemcpyHostToDevice, h->stream));
  } else {  // a handle sized for far more pairs than this batch has
    CU(cudaMemcpyAsync(h->d_pc, h->h_in + h->h_in_pc, pc_bytes, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_keys_un, s_un, key_bytes, cudaMemcpyHostToDevice, h->stream));
  }
  if (!m.gyro_init) CU(cudaMemcpyAsync(h->d_keys, s_k, key_bytes, cudaMemcpyHostToDevice, h->stream));
  if (any_table) {
    const size_t per = (size_t)W * H * 2;
    if (!h->d_ntab || h->ntab_stride < per) {
      cudaFree(h->d_ntab); h->d_ntab = nullptr;
      h->ntab_stride = (size_t)h->cfg.max_width * h->cfg.max_height * 2;
      CU(cudaMalloc(&h->d_ntab, h->ntab_stride * sizeof(float) * h->cfg.max_pairs));
    }
    for (int p = 0; p < n_pairs; ++p)
      if (in[p].normalize_table)
        CU(cudaMemcpyAsync(h->d_ntab + (size_t)p * h->ntab_stride, in[p].normalize_table, per * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  // from here on the slots change: the geometry is committed together with the images, and whatever pyramids a failed
  // copy leaves behind are not continued from
  h->geom = geom;
  h->cur_pairs = 0;  // until the run has built the new current pyramids
  rc = cont ? upload_images_continue(h, n_pairs, curs.data(), W, H, pitches.data())
            : upload_images(h, n_pairs, refs.data(), curs.data(), W, H, pitches.data());
  if (rc != PAGK_OK) return rc;
  h->cont = cont;
  h->mode = m; h->n_pairs = n_pairs; h->n_max = n_max; h->e_type = prm->e_type;
  h->resident = true;
  return PAGK_OK;
}

int pagk_run_resident(pagk_handle *h) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  if (!h->resident) return fail(PAGK_ERR_INVALID, "pagk_run_resident: no resident batch (call pagk_upload_batch)");
  if (h->n_pairs == 0) { h->ran = true; return PAGK_OK; }
  CU(cudaSetDevice(h->cfg.device));
  const PagkOutPtrs o = h->outs();
  cudaStream_t st = h->stream;
  const bool lk = (h->e_type != PAGK_GYRO_PREDICT);
  if (!lk) h->cur_pairs = 0;  // no pyramids are built: nothing a later batch could continue from
  CU(cudaEventRecord(h->ev[0], st));
  // stage clocks off: no event sits between two kernels (the whole device time is booked on the patch alignment)
  const bool stages = h->stage_timing != 0;
  h->ran_stages = stages;
  if (lk) {
    // K2 (prediction) and K1 (pyramids) are independent: K2 runs on the handle's second stream beside K1 and joins
    // before K3.  Stage clocks: ev[0]..ev[1] = the pyramid build with the prediction beside it, ev[1]..ev[2] = what
    // is left of the prediction after the pyramids are done.
    // (a small batch is bound by the host's launch rate, not by the device: there the fork and the join -- four runtime calls
    // for three microseconds of kernel -- cost more than they hide, and K2 simply follows K1 on the handle's stream)
    const bool beside = (long long)h->n_pairs * h->n_max >= 16384;
    if (beside) {
      CU(cudaStreamWaitEvent(h->aux, h->ev[0], 0));
      CU((cudaError_t)pagk_launch_predict(h->d_pc, h->d_keys_un, h->d_keys, o, h->mode, h->cfg.max_keys, h->n_max, h->n_pairs,
                                          h->geom.width, h->geom.height, h->d_ntab, h->ntab_stride, h->aux, &h->launches));
      CU(cudaEventRecord(h->ev_aux, h->aux));
    }
    // a stream continuation already has the reference pyramids (copied from the previous current ones at upload)
    if (h->cont) CU((cudaError_t)pagk_launch_pyramids(h->d_images + h->geom.slot_bytes, h->geom, h->n_pairs, 2, st, &h->launches));
    else CU((cudaError_t)pagk_launch_pyramids(h->d_images, h->geom, 2 * h->n_pairs, 1, st, &h->launches));
    h->cur_pairs = h->n_pairs;
    if (stages) CU(cudaEventRecord(h->ev[1], st));
    if (beside) CU(cudaStreamWaitEvent(st, h->ev_aux, 0));
    else CU((cudaError_t)pagk_launch_predict(h->d_pc, h->d_keys_un, h->d_keys, o, h->mode, h->cfg.max_keys, h->n_max, h->n_pairs,
                                             h->geom.width, h->geom.height, h->d_ntab, h->ntab_stride, st, &h->launches));
  } else {
    if (stages) CU(cudaEventRecord(h->ev[1], st));
    CU((cudaError_t)pagk_launch_predict(h->d_pc, h->d_keys_un, h->d_keys, o, h->mode, h->cfg.max_keys, h->n_max, h->n_pairs,
                                        h->geom.width, h->geom.height, h->d_ntab, h->ntab_stride, st, &h->launches));
  }
  if (stages) CU(cudaEventRecord(h->ev[2], st));
  const bool timed = lk && h->tev_used >= 0 && h->tev_used < 1024;
  if (timed) {
    while ((int)h->tev.size() < 2 * (h->tev_used + 1)) { cudaEvent_t e; CU(cudaEventCreate(&e)); h->tev.push_back(e); }
    CU(cudaEventRecord(h->tev[2 * h->tev_used], st));
  }
  if (lk) CU((cudaError_t)launch_lk(h, o, h->mode, h->n_max, h->n_pairs));
  if (timed) { CU(cudaEventRecord(h->tev[2 * h->tev_used + 1], st)); ++h->tev_used; }
  if (stages) CU(cudaEventRecord(h->ev[3], st));
  if (lk) CU((cudaError_t)pagk_launch_epilogue(h->d_pc, o, h->mode, h->cfg.max_keys, h->n_max, h->n_pairs, h->d_res, 1, st, &h->launches));
  else {
    // eType 1 never constructs a PatchMatch: its six result vectors stay empty in the reference
    const size_t NK = (size_t)h->n_pairs * h->cfg.max_keys;
    for (int k : {O_PM_UN, O_PM, O_PM_STATUS, O_PIX_ERR, O_DIST, O_NCC, O_ITERS})
      CU(cudaMemsetAsync(h->d_out + h->out_off[k], 0, NK * kOutElt[k], st));
    CU((cudaError_t)pagk_launch_count_status(h->d_pc, o, h->cfg.max_keys, h->n_pairs, h->d_res, st, &h->launches));
  }
  CU(cudaEventRecord(h->ev[4], st));
  h->ran = true;
  return PAGK_OK;
}

int pagk_set_stage_timing(pagk_handle *h, int on) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  h->stage_timing = on ? 1 : 0;
  return PAGK_OK;
}

int pagk_set_device_share(pagk_handle *h, int n_handles) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  if (n_handles < 1 || n_handles > 8) return fail(PAGK_ERR_INVALID, "pagk_set_device_share: 1 <= n_handles <= 8");
  h->device_share = n_handles;
  return PAGK_OK;
}

int pagk_synchronize(pagk_handle *h) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  return PAGK_OK;
}

int pagk_last_run_ms(pagk_handle *h, float *total_ms, float *pyramid_ms, float *predict_ms, float *lk_ms, float *filter_ms) {
  if (!h || !h->ran) return fail(PAGK_ERR_INVALID, "no run to time");
  if (h->n_pairs == 0) { if (total_ms) *total_ms = 0; return PAGK_OK; }
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaEventSynchronize(h->ev[4]));
  float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  CU(cudaEventElapsedTime(&t[0], h->ev[0], h->ev[4]));
  if (h->ran_stages) { for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&t[i + 1], h->ev[i], h->ev[i + 1])); }
  else t[3] = t[0];  // no stage clocks: everything is booked on the patch alignment
  if (total_ms) *total_ms = t[0];
  if (pyramid_ms) *pyramid_ms = t[1];
  if (predict_ms) *predict_ms = t[2];
  if (lk_ms) *lk_ms = t[3];
  if (filter_ms) *filter_ms = t[4];
  return PAGK_OK;
}

void *pagk_stream(pagk_handle *h) { return h ? (void *)h->stream : nullptr; }

int pagk_share_stream(pagk_handle *h, pagk_handle *other) {
  if (!h || !other || h->cfg.device != other->cfg.device) return fail(PAGK_ERR_INVALID, "pagk_share_stream: handles must live on one device");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream != other->stream) cudaStreamDestroy(h->stream);
  h->stream = other->stream;
  h->own_stream = false;
  return PAGK_OK;
}

int pagk_timing_reset(pagk_handle *h) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  h->tev_used = 0;
  return PAGK_OK;
}

int pagk_timing_read(pagk_handle *h, int *n_runs, float *lk_ms_sum) {
  if (!h || !n_runs || !lk_ms_sum) return fail(PAGK_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  float sum = 0.f;
  const int n = h->tev_used < 0 ? 0 : h->tev_used;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->tev[2 * i], h->tev[2 * i + 1]));
    sum += ms;
  }
  *n_runs = n; *lk_ms_sum = sum;
  return PAGK_OK;
}
int64_t pagk_launch_count(pagk_handle *h) { return h ? h->launches : 0; }

}  // extern "C"

namespace {
struct Want { OutKind k; size_t field_off; };
const Want kWants[] = {
    {O_PT_PREDICT_UN, offsetof(pagk_pair_out, pt_predict_un)}, {O_PT_PREDICT, offsetof(pagk_pair_out, pt_predict)},
    {O_STATUS, offsetof(pagk_pair_out, status)}, {O_PT_GYRO_UN, offsetof(pagk_pair_out, pt_gyro_predict_un)},
    {O_PT_GYRO, offsetof(pagk_pair_out, pt_gyro_predict)}, {O_FLOWS, offsetof(pagk_pair_out, flows_predict_un)},
    {O_AFFINE, offsetof(pagk_pair_out, affine)}, {O_CFLOWS, offsetof(pagk_pair_out, corner_flows)},
    {O_CORNERS_UN, offsetof(pagk_pair_out, pt_corners_un)}, {O_CORNERS, offsetof(pagk_pair_out, pt_corners)},
    {O_PM_UN, offsetof(pagk_pair_out, pm_pt_un)}, {O_PM, offsetof(pagk_pair_out, pm_pt)},
    {O_PM_STATUS, offsetof(pagk_pair_out, pm_status)}, {O_PIX_ERR, offsetof(pagk_pair_out, pixel_error)},
    {O_DIST, offsetof(pagk_pair_out, distance)}, {O_NCC, offsetof(pagk_pair_out, ncc)}, {O_ITERS, offsetof(pagk_pair_out, iters)}};
const int kNumWants = (int)(sizeof(kWants) / sizeof(kWants[0]));
inline unsigned char *out_field(pagk_pair_out *out, int p, size_t off) {
  return *reinterpret_cast<unsigned char *const *>(reinterpret_cast<const char *>(&out[p]) + off);
}

// enqueue the device-to-host copies of the requested result vectors (no synchronisation)
int download_enqueue(pagk_handle *h, int n_pairs, pagk_pair_out *out) {
  if (!h || (n_pairs > 0 && !out)) return fail(PAGK_ERR_INVALID, "null argument");
  if (!h->ran || n_pairs != h->n_pairs) return fail(PAGK_ERR_INVALID, "pagk_download_batch: no finished run of that size");
  h->pending_out = out; h->pending_n = n_pairs;
  h->pending_staged.assign(kNumWants, 0);
  if (n_pairs == 0) return PAGK_OK;
  CU(cudaSetDevice(h->cfg.device));
  // A small batch is bound by the number of runtime calls, not by bytes: when the requested vectors span at most 1 MB of the
  // result block they come down in ONE copy into the pinned mirror and are scattered from there.
  {
    size_t lo = (size_t)-1, hi = 0;
    for (int w = 0; w < kNumWants; ++w) {
      bool any = false;
      for (int p = 0; p < n_pairs && !any; ++p) any = out_field(out, p, kWants[w].field_off) != nullptr;
      if (!any) continue;
      const size_t a = h->out_off[kWants[w].k], b = a + (size_t)n_pairs * h->cfg.max_keys * kOutElt[kWants[w].k];
      lo = a < lo ? a : lo; hi = b > hi ? b : hi;
      h->pending_staged[w] = 1;
    }
    if (hi > lo && hi - lo <= (size_t)1 << 20) {
      CU(cudaMemcpyAsync(h->h_out + lo, h->d_out + lo, hi - lo, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaMemcpyAsync(h->h_res, h->d_res, (size_t)n_pairs * sizeof(PagkPairResult), cudaMemcpyDeviceToHost, h->stream));
      return PAGK_OK;
    }
    h->pending_staged.assign(kNumWants, 0);
  }
  for (int w = 0; w < kNumWants; ++w) {
    const size_t elt = kOutElt[kWants[w].k];
    bool any = false, direct = true;
    unsigned char *first = out_field(out, 0, kWants[w].field_off);
    for (int p = 0; p < n_pairs; ++p) {
      unsigned char *d = out_field(out, p, kWants[w].field_off);
      if (d) any = true;
      // contiguous [n_pairs][max_keys] caller block: copy straight into it
      if (!d || h->pcs[p].n_keys != h->cfg.max_keys || d != first + (size_t)p * h->cfg.max_keys * elt) direct = false;
    }
    if (!any) continue;
    const size_t bytes = (size_t)n_pairs * h->cfg.max_keys * elt;
    if (direct) {
      CU(cudaMemcpyAsync(first, h->d_out + h->out_off[kWants[w].k], bytes, cudaMemcpyDeviceToHost, h->stream));
    } else {
      CU(cudaMemcpyAsync(h->h_out + h->out_off[kWants[w].k], h->d_out + h->out_off[kWants[w].k], bytes, cudaMemcpyDeviceToHost, h->stream));
      h->pending_staged[w] = 1;
    }
  }
  CU(cudaMemcpyAsync(h->h_res, h->d_res, (size_t)n_pairs * sizeof(PagkPairResult), cudaMemcpyDeviceToHost, h->stream));
  return PAGK_OK;
}

// wait for the copies, scatter the staged vectors, fill the scalar results
int download_finish(pagk_handle *h) {
  pagk_pair_out *out = h->pending_out;
  const int n_pairs = h->pending_n;
  h->pending_out = nullptr; h->pending_n = 0;
  if (!out || n_pairs == 0) return PAGK_OK;
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  for (int w = 0; w < kNumWants; ++w) {
    if (!h->pending_staged[w]) continue;
    const size_t elt = kOutElt[kWants[w].k];
    for (int p = 0; p < n_pairs; ++p) {
      unsigned char *d = out_field(out, p, kWants[w].field_off);
      if (d) std::memcpy(d, h->h_out + h->out_off[kWants[w].k] + (size_t)p * h->cfg.max_keys * elt, (size_t)h->pcs[p].n_keys * elt);
    }
  }
  float t[5] = {0, 0, 0, 0, 0};
  pagk_last_run_ms(h, &t[0], &t[1], &t[2], &t[3], &t[4]);
  for (int p = 0; p < n_pairs; ++p) {
    std::memcpy(out[p].Rcl, &h->rcl[(size_t)p * 9], 9 * sizeof(float));
    std::memcpy(out[p].KRKinv, &h->krk[(size_t)p * 9], 9 * sizeof(float));
    out[p].n_predict = h->h_res[p].n_predict;
    out[p].n_iterations = h->h_res[p].n_iterations;
    // per-pair share of the batch's device time, seconds (the reference records wall time per tracker)
    out[p].t_gyro_predict = t[2] * 1e-3f / n_pairs;
    out[p].t_opt_flow = (t[1] + t[3]) * 1e-3f / n_pairs;
    out[p].t_filter = t[4] * 1e-3f / n_pairs;
  }
  return PAGK_OK;
}
}  // namespace

extern "C" {

int pagk_download_batch(pagk_handle *h, int n_pairs, pagk_pair_out *out) {
  const int rc = download_enqueue(h, n_pairs, out);
  if (rc != PAGK_OK) return rc;
  return download_finish(h);
}

int pagk_submit_batch(pagk_handle *h, const pagk_params *prm, int n_pairs, const pagk_pair_in *in, pagk_pair_out *out) {
  int rc = pagk_upload_batch(h, prm, n_pairs, in);
  if (rc != PAGK_OK) {
    if (rc == PAGK_ERR_UNSUPPORTED && out)
      for (int p = 0; p < n_pairs; ++p) out[p].n_predict = -1;  // TrackFeatures() returns -1 (:415-418)
    return rc;
  }
  rc = pagk_run_resident(h);
  if (rc != PAGK_OK) return rc;
  return download_enqueue(h, n_pairs, out);
}

int pagk_wait_batch(pagk_handle *h) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  return download_finish(h);
}

int pagk_track_batch(pagk_handle *h, const pagk_params *prm, int n_pairs, const pagk_pair_in *in, pagk_pair_out *out) {
  const int rc = pagk_submit_batch(h, prm, n_pairs, in, out);
  if (rc != PAGK_OK) return rc;
  return pagk_wait_batch(h);
}

int pagk_pyramid_level_size(int width, int height, int level, int *cols, int *rows) {
  PagkGeom g;
  if (level < 0 || level >= PAGK_MAX_LEVELS || !make_geom(width, height, level + 1, &g)) return fail(PAGK_ERR_INVALID, "bad level");
  if (cols) *cols = g.lv[level].cols;
  if (rows) *rows = g.lv[level].rows;
  return PAGK_OK;
}

int pagk_build_pyramids(pagk_handle *h, int n_images, const uint8_t *const *imgs, int width, int height, int pitch, int levels) {
  if (!h || !imgs) return fail(PAGK_ERR_INVALID, "null argument");
  if (n_images < 1 || n_images > 2 * h->cfg.max_pairs) return fail(PAGK_ERR_INVALID, "n_images exceeds 2 * max_pairs");
  PagkGeom geom;
  int rc = check_batch(h, 1, width, height, levels, 1, &geom);
  if (rc != PAGK_OK) return rc;
  for (int i = 0; i < n_images; ++i)
    if (!imgs[i] || pitch < width) return fail(PAGK_ERR_INVALID, "bad image pointer or pitch");
  CU(cudaSetDevice(h->cfg.device));
  h->resident = false; h->ran = false;
  h->geom = geom;
  h->cur_pairs = 0;
  for (int i = 0; i < n_images; ++i)
    CU(cudaMemcpy2DAsync(h->d_images + (size_t)i * h->geom.slot_bytes + h->geom.lv[0].offset, (size_t)h->geom.lv[0].pitch, imgs[i], pitch,
                         width, height, cudaMemcpyHostToDevice, h->stream));
  CU((cudaError_t)pagk_launch_pyramids(h->d_images, h->geom, n_images, 1, h->stream, &h->launches));
  CU(cudaStreamSynchronize(h->stream));
  return PAGK_OK;
}

int pagk_get_pyramid_level(pagk_handle *h, int image, int level, uint8_t *dst, size_t dst_bytes) {
  if (!h || !dst) return fail(PAGK_ERR_INVALID, "null argument");
  if (image < 0 || image >= 2 * h->cfg.max_pairs || level < 0 || level >= h->geom.levels) return fail(PAGK_ERR_INVALID, "bad image or level");
  const PagkLevelGeom &L = h->geom.lv[level];
  const size_t bytes = (size_t)L.cols * L.rows;
  if (dst_bytes < bytes) return fail(PAGK_ERR_INVALID, "destination too small");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpy2DAsync(dst, (size_t)L.cols, h->d_images + (size_t)image * h->geom.slot_bytes + L.offset, (size_t)L.pitch, (size_t)L.cols,
                       (size_t)L.rows, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return PAGK_OK;
}

int pagk_integrate_gyro(const pagk_pair_in *in, float Rcl[9], float KRKinv[9]) {
  if (!in || !Rcl || !KRKinv) return fail(PAGK_ERR_INVALID, "null argument");
  PagkPairConst pc;
  pair_const(*in, &pc, Rcl, KRKinv);
  return PAGK_OK;
}

int pagk_gyro_predict(pagk_handle *h, const pagk_params *prm, const pagk_pair_in *in, pagk_pair_out *out) {
  if (!h || !prm || !in || !out) return fail(PAGK_ERR_INVALID, "null argument");
  pagk_params p = *prm;
  p.e_type = PAGK_GYRO_PREDICT;
  if (p.pyramids < 1) p.pyramids = 1;
  return pagk_track_batch(h, &p, 1, in, out);
}

static int aux_staging(pagk_handle *h, size_t bytes) {
  if (bytes <= h->h_aux_bytes) return PAGK_OK;
  cudaFreeHost(h->h_aux); h->h_aux = nullptr; h->h_aux_bytes = 0;
  CU(cudaMallocHost(&h->h_aux, bytes));
  h->h_aux_bytes = bytes;
  return PAGK_OK;
}

int pagk_geometry_validation(pagk_handle *h, int n_pairs, const pagk_geometry_in *in, pagk_geometry_out *out) {
  if (!h || !in || !out) return fail(PAGK_ERR_INVALID, "null argument");
  if (n_pairs < 0 || n_pairs > h->cfg.max_pairs) return fail(PAGK_ERR_INVALID, "n_pairs exceeds pagk_config.max_pairs");
  if (n_pairs == 0) return PAGK_OK;
  CU(cudaSetDevice(h->cfg.device));
  if (!h->d_geo) {
    CU(cudaMalloc(&h->d_geo, (size_t)h->cfg.max_pairs * sizeof(PagkGeoModel)));
    CU(cudaMalloc(&h->d_geo_res, (size_t)h->cfg.max_pairs * sizeof(PagkGeoResult)));
  }
  const PagkOutPtrs o = h->outs();
  cudaStream_t st = h->stream;
  std::vector<PagkGeoModel> models((size_t)n_pairs);
  for (int p = 0; p < n_pairs; ++p) {
    const pagk_geometry_in &g = in[p];
    if (g.n_keys < 0 || g.n_keys > h->cfg.max_keys) return fail(PAGK_ERR_INVALID, "n_keys exceeds pagk_config.max_keys");
    const bool given = g.keys_ref_un && g.pt_predict_un && g.status;
    if (!given && (g.keys_ref_un || g.pt_predict_un || g.status))
      return fail(PAGK_ERR_INVALID, "keys_ref_un, pt_predict_un and status: all three or none");
    if (!given && !(h->ran && p < h->n_pairs)) return fail(PAGK_ERR_INVALID, "no resident run to validate");
    PagkGeoModel &M = models[(size_t)p];
    std::memcpy(M.H21, g.H21, sizeof(M.H21)); std::memcpy(M.F21, g.F21, sizeof(M.F21));
    pagk_inv3_f64(g.H21, M.H12);  // cv::Mat H12 = H21.inv(), src/gyro_aided_tracker.cpp:597
    M.sigma = g.sigma; M.n_keys = g.n_keys;
    const size_t off = (size_t)p * h->cfg.max_keys, n = (size_t)g.n_keys;
    if (given && n) {
      CU(cudaMemcpyAsync(h->d_keys_un + off, g.keys_ref_un, n * sizeof(float2), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(o.pt_predict_un + off, g.pt_predict_un, n * sizeof(float2), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(o.status + off, g.status, n, cudaMemcpyHostToDevice, st));
    }
  }
  CU(cudaMemcpyAsync(h->d_geo, models.data(), models.size() * sizeof(PagkGeoModel), cudaMemcpyHostToDevice, st));
  bool any_est = false;
  for (int p = 0; p < n_pairs; ++p) any_est |= in[p].estimate != 0;
  if (any_est) {  // the two robust estimators on the device: H21 / H12 / F21 of those pairs are overwritten in d_geo
    const size_t NK2 = (size_t)2 * h->cfg.max_pairs * h->cfg.max_keys;
    const size_t off_in = align_up(NK2 * sizeof(int), 256), off_est = off_in + align_up(NK2, 256);
    if (!h->d_ransac) CU(cudaMalloc(&h->d_ransac, off_est + align_up((size_t)h->cfg.max_pairs, 256)));
    std::vector<unsigned char> est((size_t)n_pairs);
    for (int p = 0; p < n_pairs; ++p) est[(size_t)p] = in[p].estimate != 0 ? 1 : 0;
    CU(cudaMemcpyAsync(h->d_ransac + off_est, est.data(), est.size(), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // `est` and `models` are host temporaries
    CU((cudaError_t)pagk_launch_ransac(h->d_keys_un, o.pt_predict_un, o.status, h->cfg.max_keys, n_pairs, in[0].seed, (int *)h->d_ransac,
                                       h->d_ransac + off_in, h->d_geo, h->d_ransac + off_est, st, &h->launches));
  }
  CU((cudaError_t)pagk_launch_geometry(h->d_geo, h->d_keys_un, o.pt_predict_un, o.status, h->cfg.max_keys, n_pairs, h->d_geo_res,
                                       st, &h->launches));
  std::vector<PagkGeoResult> res((size_t)n_pairs);
  CU(cudaMemcpyAsync(res.data(), h->d_geo_res, res.size() * sizeof(PagkGeoResult), cudaMemcpyDeviceToHost, st));
  if (any_est) CU(cudaMemcpyAsync(models.data(), h->d_geo, models.size() * sizeof(PagkGeoModel), cudaMemcpyDeviceToHost, st));
  // one copy of the status rows into pinned memory, then per pair on the host
  const size_t st_bytes = (size_t)n_pairs * h->cfg.max_keys;
  if (aux_staging(h, st_bytes) != PAGK_OK) return PAGK_ERR_CUDA;
  CU(cudaMemcpyAsync(h->h_aux, o.status, st_bytes, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int p = 0; p < n_pairs; ++p)
    if (out[p].status && in[p].n_keys) std::memcpy(out[p].status, h->h_aux + (size_t)p * h->cfg.max_keys, (size_t)in[p].n_keys);
  for (int p = 0; p < n_pairs; ++p) {
    out[p].score_H = res[(size_t)p].score_H; out[p].score_F = res[(size_t)p].score_F; out[p].used_H = res[(size_t)p].used_H;
    out[p].n_candidates = res[(size_t)p].n_candidates; out[p].n_inlier = res[(size_t)p].n_inlier;
    std::memcpy(out[p].H21, models[(size_t)p].H21, sizeof(out[p].H21)); std::memcpy(out[p].F21, models[(size_t)p].F21, sizeof(out[p].F21));
  }
  return PAGK_OK;
}

int pagk_set_predict_keypoints_and_mask(pagk_handle *h, int n_pairs, const pagk_carry_in *in, pagk_carry_out *out) {
  if (!h || !in || !out) return fail(PAGK_ERR_INVALID, "null argument");
  if (n_pairs < 0 || n_pairs > h->cfg.max_pairs) return fail(PAGK_ERR_INVALID, "n_pairs exceeds pagk_config.max_pairs");
  if (n_pairs == 0) return PAGK_OK;
  CU(cudaSetDevice(h->cfg.device));
  const size_t MP = (size_t)h->cfg.max_pairs, NK = MP * h->cfg.max_keys;
  // layout of d_carry: consts | normal_last | keys | keys_un | keys_normal | flow_last | index_in_last | n_out
  const size_t off_nl = align_up(MP * sizeof(PagkCarryConst), 256), off_k = off_nl + align_up(NK * 8, 256), off_ku = off_k + align_up(NK * 8, 256),
               off_kn = off_ku + align_up(NK * 8, 256), off_fl = off_kn + align_up(NK * 8, 256), off_ix = off_fl + align_up(NK * 8, 256),
               off_n = off_ix + align_up(NK * 4, 256), total = off_n + align_up(MP * 4, 256);
  if (!h->d_carry) CU(cudaMalloc(&h->d_carry, total));
  const PagkOutPtrs o = h->outs();
  cudaStream_t st = h->stream;
  std::vector<PagkCarryConst> cc((size_t)n_pairs);
  const size_t rows8 = (size_t)n_pairs * h->cfg.max_keys * 8;  // bytes of n_pairs rows of a float2 vector
  if (aux_staging(h, 5 * rows8) != PAGK_OK) return PAGK_ERR_CUDA;
  CU(cudaStreamSynchronize(st));  // the staging block may still feed an earlier copy
  bool any_mask = false;
  int W = 0, H = 0;
  for (int p = 0; p < n_pairs; ++p) {
    const pagk_carry_in &g = in[p];
    if (g.n_keys < 0 || g.n_keys > h->cfg.max_keys) return fail(PAGK_ERR_INVALID, "n_keys exceeds pagk_config.max_keys");
    if (g.n_keys > 0 && (!g.keys_normal_last || !out[p].keys || !out[p].keys_un || !out[p].keys_normal || !out[p].index_in_last || !out[p].flow_velocity_last))
      return fail(PAGK_ERR_INVALID, "null vector");
    if (g.width < 14 || g.height < 14 || g.width > h->cfg.max_width || g.height > h->cfg.max_height) return fail(PAGK_ERR_INVALID, "bad mask size");
    const bool given = g.pt_predict && g.pt_predict_un && g.status;
    if (!given && (g.pt_predict || g.pt_predict_un || g.status)) return fail(PAGK_ERR_INVALID, "pt_predict, pt_predict_un and status: all three or none");
    if (!given && !(h->ran && p < h->n_pairs)) return fail(PAGK_ERR_INVALID, "no resident run to carry over");
    if (p == 0) { W = g.width; H = g.height; }
    if (g.width != W || g.height != H) return fail(PAGK_ERR_INVALID, "all pairs of a call must share the mask size");
    PagkCarryConst &c = cc[(size_t)p];
    c.cx = g.cx; c.cy = g.cy;
    c.fx_inv = (float)(1.0 / g.fx); c.fy_inv = (float)(1.0 / g.fy);  // Frame ctor, src/frame.cpp:71
    c.dt = g.t_cur - g.t_last;
    c.n_keys = g.n_keys; c.width = g.width; c.height = g.height; c.pad = 0;
    any_mask |= out[p].mask != nullptr;
    const size_t off = (size_t)p * h->cfg.max_keys, n = (size_t)g.n_keys;
    if (n) std::memcpy(h->h_aux + off * 8, g.keys_normal_last, n * 8);  // one H2D copy below
    if (given && n) {
      CU(cudaMemcpyAsync(o.pt_predict + off, g.pt_predict, n * 8, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(o.pt_predict_un + off, g.pt_predict_un, n * 8, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(o.status + off, g.status, n, cudaMemcpyHostToDevice, st));
    }
  }
  if (any_mask) {
    if (!h->d_mask) {
      h->mask_stride = align_up((size_t)h->cfg.max_width * h->cfg.max_height, 256);
      CU(cudaMalloc(&h->d_mask, h->mask_stride * MP));
    }
    CU(cudaMemsetAsync(h->d_mask, 1, h->mask_stride * (size_t)n_pairs, st));  // cv::Mat::ones
  }
  CU(cudaMemcpyAsync(h->d_carry + off_nl, h->h_aux, rows8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->d_carry, cc.data(), cc.size() * sizeof(PagkCarryConst), cudaMemcpyHostToDevice, st));
  CU((cudaError_t)pagk_launch_carry((const PagkCarryConst *)h->d_carry, o.pt_predict, o.pt_predict_un, o.status,
                                    (const float2 *)(h->d_carry + off_nl), h->cfg.max_keys, n_pairs, (float2 *)(h->d_carry + off_k),
                                    (float2 *)(h->d_carry + off_ku), (float2 *)(h->d_carry + off_kn), (int *)(h->d_carry + off_ix),
                                    (float2 *)(h->d_carry + off_fl), (int *)(h->d_carry + off_n), any_mask ? h->d_mask : nullptr,
                                    h->mask_stride, st, &h->launches));
  std::vector<int> n_out((size_t)n_pairs);
  CU(cudaMemcpyAsync(n_out.data(), h->d_carry + off_n, n_out.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  // the five result vectors: one copy each into pinned memory (the normals staged above are consumed by then), per pair on the host
  CU(cudaMemcpyAsync(h->h_aux, h->d_carry + off_k, rows8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->h_aux + rows8, h->d_carry + off_ku, rows8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->h_aux + 2 * rows8, h->d_carry + off_kn, rows8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->h_aux + 3 * rows8, h->d_carry + off_fl, rows8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->h_aux + 4 * rows8, h->d_carry + off_ix, rows8 / 2, cudaMemcpyDeviceToHost, st));
  for (int p = 0; p < n_pairs; ++p)
    if (out[p].mask) CU(cudaMemcpyAsync(out[p].mask, h->d_mask + (size_t)p * h->mask_stride, (size_t)W * H, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int p = 0; p < n_pairs; ++p) {
    const size_t off = (size_t)p * h->cfg.max_keys, n = (size_t)in[p].n_keys;  // survivors <= n_keys: copy the capacity
    if (!n) continue;
    std::memcpy(out[p].keys, h->h_aux + off * 8, n * 8);
    std::memcpy(out[p].keys_un, h->h_aux + rows8 + off * 8, n * 8);
    std::memcpy(out[p].keys_normal, h->h_aux + 2 * rows8 + off * 8, n * 8);
    std::memcpy(out[p].flow_velocity_last, h->h_aux + 3 * rows8 + off * 8, n * 8);
    std::memcpy(out[p].index_in_last, h->h_aux + 4 * rows8 + off * 4, n * 4);
  }
  for (int p = 0; p < n_pairs; ++p) out[p].n_out = n_out[(size_t)p];
  return PAGK_OK;
}

// device block of the detector entry points: image | mask | score (u16) | keep | row or cell counts | their offsets
static int fast_buffers(pagk_handle *h, size_t *off_mask, size_t *off_score, size_t *off_keep, size_t *off_rc, size_t *off_ro) {
  const size_t WH = align_up((size_t)h->cfg.max_width * h->cfg.max_height, 256);
  const size_t n_cells = (size_t)(h->cfg.max_width / 30 + 1) * (size_t)(h->cfg.max_height / 30 + 1);
  const size_t n_int = std::max((size_t)h->cfg.max_height, n_cells) + 1;
  const size_t RW = align_up(n_int * sizeof(int), 256);
  *off_mask = WH; *off_score = 2 * WH; *off_keep = 4 * WH; *off_rc = 5 * WH; *off_ro = *off_rc + RW;
  if (!h->d_fast) CU(cudaMalloc(&h->d_fast, *off_ro + RW));
  return PAGK_OK;
}

static int fast_out_buffers(pagk_handle *h, int max_out) {
  if (max_out > h->fast_out_cap) {
    cudaFree(h->d_fast_out); h->d_fast_out = nullptr; h->fast_out_cap = 0;
    CU(cudaMalloc(&h->d_fast_out, align_up((size_t)max_out * 8, 256) + align_up((size_t)max_out * 4, 256)));
    h->fast_out_cap = max_out;
  }
  return PAGK_OK;
}

int pagk_orb_cell_detect(pagk_handle *h, const uint8_t *img, int width, int height, int pitch, int ini_th, int min_th,
                         const uint8_t *mask, int max_out, float *xy, float *response, int *n_out) {
  if (!h || !img || !n_out || max_out < 0 || (max_out > 0 && (!xy || !response))) return fail(PAGK_ERR_INVALID, "null argument");
  if (width < 1 || height < 1 || pitch < width || width > h->cfg.max_width || height > h->cfg.max_height)
    return fail(PAGK_ERR_INVALID, "image exceeds pagk_config.max_width/max_height");
  *n_out = 0;
  // the cell grid, with the reference's own float arithmetic (src/ORBextractor.cc:793-811)
  const int EDGE_THRESHOLD = 19;
  const float W = 30;
  PagkCellGrid g;
  g.min_x = EDGE_THRESHOLD - 3; g.min_y = g.min_x;
  g.max_x = width - EDGE_THRESHOLD + 3; g.max_y = height - EDGE_THRESHOLD + 3;
  const float fw = (float)(g.max_x - g.min_x), fh = (float)(g.max_y - g.min_y);
  g.n_cols = (int)(fw / W); g.n_rows = (int)(fh / W);
  if (g.n_cols < 1 || g.n_rows < 1) return PAGK_OK;
  g.w_cell = (int)std::ceil(fw / g.n_cols); g.h_cell = (int)std::ceil(fh / g.n_rows);
  CU(cudaSetDevice(h->cfg.device));
  size_t off_mask, off_score, off_keep, off_rc, off_ro;
  if (fast_buffers(h, &off_mask, &off_score, &off_keep, &off_rc, &off_ro) != PAGK_OK) return PAGK_ERR_CUDA;
  if (fast_out_buffers(h, max_out) != PAGK_OK) return PAGK_ERR_CUDA;
  cudaStream_t st = h->stream;
  CU(cudaMemcpy2DAsync(h->d_fast, (size_t)width, img, (size_t)pitch, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, st));
  if (mask) CU(cudaMemcpyAsync(h->d_fast + off_mask, mask, (size_t)width * height, cudaMemcpyHostToDevice, st));
  float2 *d_xy = (float2 *)h->d_fast_out;
  float *d_rs = (float *)(h->d_fast_out + align_up((size_t)h->fast_out_cap * 8, 256));
  const int n_cells = g.n_cols * g.n_rows;
  CU((cudaError_t)pagk_launch_orb_cells(h->d_fast, width, height, ini_th, min_th, mask ? h->d_fast + off_mask : nullptr,
                                        (unsigned short *)(h->d_fast + off_score), h->d_fast + off_keep, (int *)(h->d_fast + off_rc),
                                        (int *)(h->d_fast + off_ro), g, max_out, d_xy, d_rs, st, &h->launches));
  int total_found = 0;
  CU(cudaMemcpyAsync(&total_found, h->d_fast + off_ro + (size_t)n_cells * sizeof(int), sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const int n = total_found < max_out ? total_found : max_out;
  if (n > 0) {
    CU(cudaMemcpyAsync(xy, d_xy, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(response, d_rs, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  *n_out = total_found;
  return PAGK_OK;
}

int pagk_distribute_octtree(int n, const float *xy, const float *response, int min_x, int max_x, int min_y, int max_y,
                            int n_features, int *out_index, int *n_out) {
  if (n < 0 || !n_out || (n > 0 && (!xy || !response || !out_index))) return fail(PAGK_ERR_INVALID, "null argument");
  const std::vector<int> keep = pagk_octree::distribute(xy, response, n, min_x, max_x, min_y, max_y, n_features);
  for (size_t k = 0; k < keep.size(); ++k) out_index[k] = keep[k];
  *n_out = (int)keep.size();
  return PAGK_OK;
}

int pagk_orb_detect_features(pagk_handle *h, const uint8_t *img, int width, int height, int pitch, int n_features, int ini_th,
                             int min_th, const uint8_t *mask, int max_out, float *xy, float *response, int *n_out) {
  if (!h || !img || !n_out || max_out < 0 || (max_out > 0 && (!xy || !response))) return fail(PAGK_ERR_INVALID, "null argument");
  *n_out = 0;
  // every candidate of the per-cell FAST (no mask: the reference filters AFTER the thinning, :1200-1203)
  int cap = 1 << 16, found = 0;
  std::vector<float> cxy, crs;
  for (int attempt = 0; attempt < 2; ++attempt) {
    cxy.assign((size_t)cap * 2, 0.f); crs.assign((size_t)cap, 0.f);
    const int rc = pagk_orb_cell_detect(h, img, width, height, pitch, ini_th, min_th, nullptr, cap, cxy.data(), crs.data(), &found);
    if (rc != PAGK_OK) return rc;
    if (found <= cap) break;
    cap = found;
  }
  const int EDGE_THRESHOLD = 19;  // src/ORBextractor.cc:65; ComputeKeyPointsOctTree :797-800
  const int min_x = EDGE_THRESHOLD - 3, min_y = min_x, max_x = width - EDGE_THRESHOLD + 3, max_y = height - EDGE_THRESHOLD + 3;
  for (int k = 0; k < found; ++k) { cxy[2 * (size_t)k] -= (float)min_x; cxy[2 * (size_t)k + 1] -= (float)min_y; }
  const std::vector<int> keep = pagk_octree::distribute(cxy.data(), crs.data(), found, min_x, max_x, min_y, max_y, n_features);
  int n = 0;
  for (int k : keep) {
    const float x = cxy[2 * (size_t)k] + (float)min_x, y = cxy[2 * (size_t)k + 1] + (float)min_y;  // :866-867
    if (mask && !mask[(size_t)(int)y * width + (int)x]) continue;
    if (n < max_out) { xy[2 * n] = x; xy[2 * n + 1] = y; response[n] = crs[(size_t)k]; }
    ++n;
  }
  *n_out = n;
  return PAGK_OK;
}

int pagk_fast_detect(pagk_handle *h, const uint8_t *img, int width, int height, int pitch, int threshold, int nonmax,
                     const uint8_t *mask, int max_out, float *xy, float *response, int *n_out) {
  if (!h || !img || !n_out || max_out < 0 || (max_out > 0 && (!xy || !response))) return fail(PAGK_ERR_INVALID, "null argument");
  if (width < 1 || height < 1 || pitch < width || width > h->cfg.max_width || height > h->cfg.max_height)
    return fail(PAGK_ERR_INVALID, "image exceeds pagk_config.max_width/max_height");
  CU(cudaSetDevice(h->cfg.device));
  size_t off_mask, off_score, off_keep, off_rc, off_ro;
  if (fast_buffers(h, &off_mask, &off_score, &off_keep, &off_rc, &off_ro) != PAGK_OK) return PAGK_ERR_CUDA;
  if (fast_out_buffers(h, max_out) != PAGK_OK) return PAGK_ERR_CUDA;
  cudaStream_t st = h->stream;
  CU(cudaMemcpy2DAsync(h->d_fast, (size_t)width, img, (size_t)pitch, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, st));
  if (mask) CU(cudaMemcpyAsync(h->d_fast + off_mask, mask, (size_t)width * height, cudaMemcpyHostToDevice, st));
  float2 *d_xy = (float2 *)h->d_fast_out;
  float *d_rs = (float *)(h->d_fast_out + align_up((size_t)h->fast_out_cap * 8, 256));
  CU((cudaError_t)pagk_launch_fast(h->d_fast, width, height, threshold, nonmax ? 1 : 0, mask ? h->d_fast + off_mask : nullptr,
                                   (unsigned short *)(h->d_fast + off_score), h->d_fast + off_keep, (int *)(h->d_fast + off_rc),
                                   (int *)(h->d_fast + off_ro), max_out, d_xy, d_rs, st, &h->launches));
  int total_found = 0;
  CU(cudaMemcpyAsync(&total_found, h->d_fast + off_ro + (size_t)height * sizeof(int), sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const int n = total_found < max_out ? total_found : max_out;
  if (n > 0) {
    CU(cudaMemcpyAsync(xy, d_xy, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(response, d_rs, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  *n_out = total_found;
  return PAGK_OK;
}

int pagk_set_rectify_maps(pagk_handle *h, const float *map_x, const float *map_y, int width, int height) {
  if (!h) return fail(PAGK_ERR_INVALID, "null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  h->cur_pairs = 0;  // pyramids built so far belong to the other kind of image
  if (!map_x && !map_y) { cudaFree(h->d_maps); h->d_maps = nullptr; h->map_w = h->map_h = 0; return PAGK_OK; }
  if (!map_x || !map_y || width < 1 || height < 1 || width > h->cfg.max_width || height > h->cfg.max_height)
    return fail(PAGK_ERR_INVALID, "bad rectification maps");
  const size_t n = (size_t)width * height;
  cudaFree(h->d_maps); h->d_maps = nullptr;
  CU(cudaMalloc(&h->d_maps, 2 * n * sizeof(float)));
  CU(cudaMemcpy(h->d_maps, map_x, n * sizeof(float), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_maps + n, map_y, n * sizeof(float), cudaMemcpyHostToDevice));
  h->map_w = width; h->map_h = height;
  return PAGK_OK;
}

int pagk_remap_linear(pagk_handle *h, const uint8_t *src, int width, int height, int pitch, const float *map_x,
                      const float *map_y, int dst_width, int dst_height, uint8_t *dst) {
  if (!h || !src || !map_x || !map_y || !dst) return fail(PAGK_ERR_INVALID, "null argument");
  if (width < 1 || height < 1 || pitch < width || width > h->cfg.max_width || height > h->cfg.max_height || dst_width < 1 ||
      dst_height < 1 || dst_width > h->cfg.max_width || dst_height > h->cfg.max_height)
    return fail(PAGK_ERR_INVALID, "image exceeds pagk_config.max_width/max_height");
  CU(cudaSetDevice(h->cfg.device));
  const size_t WH = align_up((size_t)h->cfg.max_width * h->cfg.max_height, 256);
  if (!h->d_remap) CU(cudaMalloc(&h->d_remap, WH * 10));  // u8 source, two float maps, u8 destination
  cudaStream_t st = h->stream;
  const size_t n = (size_t)dst_width * dst_height;
  float *d_mx = (float *)(h->d_remap + WH), *d_my = (float *)(h->d_remap + 5 * WH);
  unsigned char *d_dst = h->d_remap + 9 * WH;
  CU(cudaMemcpy2DAsync(h->d_remap, (size_t)width, src, (size_t)pitch, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_mx, map_x, n * sizeof(float), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_my, map_y, n * sizeof(float), cudaMemcpyHostToDevice, st));
  CU((cudaError_t)pagk_launch_remap(h->d_remap, width, height, d_mx, d_my, dst_width, dst_height, d_dst, st, &h->launches));
  CU(cudaMemcpyAsync(dst, d_dst, n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return PAGK_OK;
}

int pagk_patch_match(pagk_handle *h, const pagk_patch_match_in *in, pagk_pair_out *out) {
  if (!h || !in || !out) return fail(PAGK_ERR_INVALID, "null argument");
  if (in->inverse) return fail(PAGK_ERR_UNSUPPORTED, "inverse mode is \"not support yet\" in the reference (src/patch_match.cpp:220)");
  if (!in->img_ref || !in->img_cur || in->pitch < in->width) return fail(PAGK_ERR_INVALID, "bad image pointer or pitch");
  if (in->n_keys < 0 || in->n_keys > h->cfg.max_keys) return fail(PAGK_ERR_INVALID, "n_keys exceeds pagk_config.max_keys");
  if (in->n_keys > 0 && (!in->keys_ref_un || !in->pt_predict_un || !in->status || (in->consider_affine_deformation && !in->affine)))
    return fail(PAGK_ERR_INVALID, "null input vector");
  PagkGeom geom;
  int rc = check_batch(h, 1, in->width, in->height, in->pyramids, in->half_patch, &geom);
  if (rc != PAGK_OK) return rc;
  CU(cudaSetDevice(h->cfg.device));
  h->resident = false; h->ran = false;
  h->geom = geom;
  h->cur_pairs = 0;
  PagkMode m;
  std::memset(&m, 0, sizeof(m));
  m.gyro_init = in->has_gyro_predict_initial ? 1 : 0; m.illum = in->consider_illumination ? 1 : 0;
  m.affine = in->consider_affine_deformation ? 1 : 0; m.regular = in->regularization_penalty ? 1 : 0;
  fill_mode_common(&m, in->half_patch, in->iterations, in->pyramids, in->calc_ncc, PAGK_PIXEL_AWARE_PREDICTION, in->lambda,
                   in->alpha, in->max_distance);
  pagk_pair_in pin;
  std::memset(&pin, 0, sizeof(pin));
  std::memcpy(pin.K, in->K, sizeof(pin.K)); std::memcpy(pin.dist, in->dist, sizeof(pin.dist));
  pin.n_dist = in->n_dist; pin.n_keys = in->n_keys;
  const float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  pin.Rcl_override = eye;
  h->pcs.assign(1, PagkPairConst());
  h->rcl.assign(9, 0.f); h->krk.assign(9, 0.f);
  pair_const(pin, &h->pcs[0], h->rcl.data(), h->krk.data());
  const size_t n = (size_t)in->n_keys;
  const PagkOutPtrs o = h->outs();
  cudaStream_t st = h->stream;
  CU(cudaMemcpyAsync(h->d_pc, h->pcs.data(), sizeof(PagkPairConst), cudaMemcpyHostToDevice, st));
  if (n) {
    CU(cudaMemcpyAsync(h->d_keys_un, in->keys_ref_un, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(o.pt_predict_un, in->pt_predict_un, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(o.gyro_status, in->status, n, cudaMemcpyHostToDevice, st));
    if (in->affine) CU(cudaMemcpyAsync(o.affine, in->affine, n * 16, cudaMemcpyHostToDevice, st));
  }
  const uint8_t *r = in->img_ref, *c = in->img_cur;
  rc = upload_images(h, 1, &r, &c, in->width, in->height, &in->pitch);
  if (rc != PAGK_OK) return rc;
  CU(cudaStreamSynchronize(st));  // the caller's (pageable) inputs may go away after this call
  h->mode = m; h->n_pairs = 1; h->n_max = in->n_keys; h->e_type = PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED;
  h->ran_stages = true;
  CU(cudaEventRecord(h->ev[0], st));
  h->cur_pairs = 0;
  CU((cudaError_t)pagk_launch_pyramids(h->d_images, h->geom, 2, 1, st, &h->launches));
  CU(cudaEventRecord(h->ev[1], st));
  CU(cudaEventRecord(h->ev[2], st));
  CU((cudaError_t)launch_lk(h, o, m, in->n_keys, 1));
  CU(cudaEventRecord(h->ev[3], st));
  CU((cudaError_t)pagk_launch_epilogue(h->d_pc, o, m, h->cfg.max_keys, in->n_keys, 1, h->d_res, 0, st, &h->launches));
  CU(cudaEventRecord(h->ev[4], st));
  h->ran = true;
  pagk_pair_out tmp = *out;  // only the PatchMatch outputs (SetMatcher, src/patch_match.cpp:370-388)
  tmp.pt_predict_un = nullptr; tmp.pt_predict = nullptr; tmp.status = nullptr; tmp.pt_gyro_predict_un = nullptr;
  tmp.pt_gyro_predict = nullptr; tmp.flows_predict_un = nullptr; tmp.affine = nullptr; tmp.corner_flows = nullptr;
  tmp.pt_corners_un = nullptr; tmp.pt_corners = nullptr;
  rc = pagk_download_batch(h, 1, &tmp);
  out->n_predict = tmp.n_predict; out->n_iterations = tmp.n_iterations;
  out->t_gyro_predict = tmp.t_gyro_predict; out->t_opt_flow = tmp.t_opt_flow; out->t_filter = tmp.t_filter;
  return rc;
}

}  // extern "C"
