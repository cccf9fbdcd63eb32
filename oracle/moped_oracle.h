/*
 * moped_oracle.h — CPU restatement of MOPED's recognition hot path. TEST INFRASTRUCTURE ONLY:
 * linked/loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the
 * product (moped_b200/, libmoped_cuda.so).
 *
 * Parity status: PINNED. Every function is checked against the reference's own stage classes
 * compiled unmodified (oracle/_ref/libmoped_ref.so, see ref_harness.cpp) by tests/test_oracle_*.py
 * and against committed golden vectors generated from that build (tests/golden/).
 */
#ifndef MOPED_ORACLE_H
#define MOPED_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* MATCH ------------------------------------------------------------------------------------ */
void mo_norm_rows(float *desc, int n, int D);
void mo_match_2nn(const float *db, int N, int D, const float *q, int Q, int *idx2, float *dist2);
int  mo_match_emit(const int *idx2, const float *dist2, int Q, float ratio, const int *model_of_row,
                   int n_models, int *match_query, int *match_row, int *model_offsets);

/* CLUSTER ---------------------------------------------------------------------------------- */
int mo_meanshift(const float *xy, int n, float radius, float merge, int minpts, int maxiter,
                 int *cluster_offsets, int *members);
int mo_cluster(const int *match_offsets, const int *match_image, const float *match_xy, int n_models, int n_images,
               float radius, float merge, int minpts, int maxiter,
               int *cluster_model, int *cluster_offsets, int *members);

/* POSE ------------------------------------------------------------------------------------- */
typedef struct { float K[4]; float TM[12]; } mo_camera;   /* K=(fx,fy,cx,cy); TM = 3x4 of cameraPose */
void mo_camera_init(mo_camera *cam, const float *K4, const float *cam_pose7);
int  mo_rand(uint64_t *state);
int  mo_rand_sample(uint64_t *state, const float *xy, const int *image, const int *tie_ids, int n, int n_samples, int *sample_pos);
void mo_init_pose(uint64_t *state, float *pose7);
void mo_lm_func(const float *p7, float *res, int n_pts, const float *xy, const float *xyz, const int *image, const mo_camera *cams);
int  mo_levmar_dif(float *p7, int n_pts, int itmax, const float *xy, const float *xyz, const int *image,
                   const mo_camera *cams, float *info10);
float mo_optimize_camera(float *pose7, int n_pts, int itmax, const float *xy, const float *xyz, const int *image, const mo_camera *cams);
void mo_project(const float *pose7, const float *xyz3, const mo_camera *cam, float *uv2);
int  mo_test_all_points(const float *pose7, int n, const float *xy, const float *xyz, const int *image,
                        const mo_camera *cams, float err_thr, unsigned char *mask);
int  mo_hypothesis(int n, const float *xy, const float *xyz, const int *image, const mo_camera *cams,
                   const int *sample_pos, int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts,
                   float *pose_lm, float *pose_refit, float *lm_err2, unsigned char *mask);
int  mo_ransac(uint64_t *state, int n, const float *xy, const float *xyz, const int *image, const int *tie_ids, const mo_camera *cams,
               int max_ransac, int max_lm, int n_pts_align, int min_npts, float err_thr, float *pose7, int *iters);

/* FILTER ----------------------------------------------------------------------------------- */
int mo_filter(int n_models, const int *match_offsets, const int *match_image, const float *match_xy, const float *match_xyz,
              const mo_camera *cams, int n_obj, const int *obj_model, const float *obj_pose, int min_points, float feat_dist,
              float min_score, unsigned char *keep, float *score, int *cluster_offsets, int *members);

/* FILTER, moped3d's depth-map variant (FILTER_PROJECTION_DEPTH_CPU) */
int mo_filter_depth_select(uint64_t *state, int n_keypoints, int sample_size, int *out_idx);
float mo_filter_depth_penalty(const float *pose7, int n_test, const float *test_xyz, const mo_camera *depth_cam, int width, int height,
                              const float *depth, const float *fill_distance, float depth_fraction, int *used);
int mo_filter_depth(int n_models, const int *match_offsets, const int *match_image, const float *match_xy, const float *match_xyz,
                    const mo_camera *cams, int n_obj, const int *obj_model, const float *obj_pose, int min_points, float feat_dist,
                    float plausible_dist, float min_score, float depth_fraction, float min_keypoint_fraction,
                    const int *test_offsets, const float *test_xyz, const mo_camera *depth_cam, int width, int height,
                    const float *depth, const float *fill_distance,
                    unsigned char *keep, float *score, int *cluster_offsets, int *members);

void mo_set_lm_finite_check(int on);   /* 0 (default): -ffinite-math-only semantics of the reference build; 1: levmar's stop=7 as in a strict build */

/* POSE, moped3d depth-aware variant (oracle only so far) ---------------------------------- */
float mo_cauchy_weight(float fill_distance);
void mo_lm_func_depth(const float *p7, float *res, int n_pts, const float *xyz, const float *world, const float *cauchy,
                      const int *image, const mo_camera *cams, float alpha);
float mo_optimize_camera_depth(float *pose7, int n_pts, int itmax, const float *xyz, const float *world, const float *cauchy,
                               const int *image, const mo_camera *cams, float alpha);
void mo_init_translation_depth(const float *world, const int *sample_pos, int n_samples, float *t3);
int mo_hypothesis_depth(int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image, const mo_camera *cams,
                        float alpha, const int *sample_pos, int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts,
                        float *pose_lm, float *pose_refit, float *lm_err2, unsigned char *mask);
int mo_ransac_depth(uint64_t *state, int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image,
                    const int *tie_ids, const mo_camera *cams, float alpha, int max_ransac, int max_lm, int n_pts_align, int min_npts,
                    float err_thr, float *pose7, int *iters);

/* second depth variant, POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU: three residuals per correspondence, Cauchy scale 25 */
float mo_cauchy_weight_v1(float fill_distance);
void mo_lm_func_depth_v1(const float *p7, float *res, int n_pts, const float *xy, const float *xyz, const float *world, const float *cauchy,
                         const int *image, const mo_camera *cams, float alpha);
int mo_hypothesis_depth_v1(int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image, const mo_camera *cams,
                           float alpha, const int *sample_pos, int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts,
                           float *pose_lm, float *pose_refit, float *lm_err2, unsigned char *mask);
int mo_ransac_depth_v1(uint64_t *state, int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image,
                       const int *tie_ids, const mo_camera *cams, float alpha, int max_ransac, int max_lm, int n_pts_align, int min_npts,
                       float err_thr, float *pose7, int *iters);

/* CLUSTER, moped3d linkage variant (oracle only so far) — moped_linkage_oracle.c ----------- */
void mo_linkage_similarity(int n, const float *xy, const float *xyz, const float *world, int W, int H, const float *depth, const float *distance,
                           int use3DFilter, float sigma2D, float sigma3D, float *K /* n x n */);
int mo_linkage_agglomerate(const float *K, int n, float cutoff, int min_pts, int linkage_type, int *cluster_offsets, int *members);
int mo_cluster_linkage(int n, const float *xy, const float *xyz, const float *world, int W, int H, const float *depth, const float *distance,
                       float cutoff, int min_pts, int use3DFilter, int linkage_type, float sigma2D, float sigma3D,
                       int *cluster_offsets, int *members);

/* FEAT (SIFT) — moped_sift_oracle.c -------------------------------------------- */
typedef struct { int octave, index, scan_row, scan_col, row, col; float X[3]; float fsize; int first_kp; } mo_sift_trace;
int mo_sift_gauss_kernel(float fblur, float *kernel /* >= 64 floats */);
void mo_sift_set_conv_fma(int on);   /* 1 (default): taps as FMA, like the CUDA kernels; 0: multiply then add, like a strict-IEEE build of the reference */
int mo_sift(const uint8_t *gray, int height, int width, int double_size, int max_kp,
            float *xy /* (col,row) */, float *scale_ori, float *desc /* x128 */);
int mo_sift_debug(const uint8_t *gray, int height, int width, int double_size, int dbg_octave, float *dbg_gauss /* 6 images */,
                  float *dbg_dog /* 5 images */, int max_trace, mo_sift_trace *trace, int *n_trace);

#ifdef __cplusplus
}
#endif
#endif
