/* rbd_oracle.h — CPU oracle (TEST INFRASTRUCTURE ONLY, parity unpinned: see rbd_oracle.c header). */
#ifndef RBD_ORACLE_H
#define RBD_ORACLE_H
#include "../include/rbd_b200.h" /* only for the rbd_model_desc POD and the RBD_JOINT_* codes */
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_workspace orc_workspace;
orc_workspace* orc_create(const rbd_model_desc* d);
void orc_destroy(orc_workspace* w);
void orc_compute_joint_jacobians(orc_workspace* w, const double* q);
void orc_update_frame_placements(orc_workspace* w);
void orc_get_frame_jacobian_lwa(orc_workspace* w, int frame, double* J6xnv_colmajor_zeroed);
void orc_rot_to_quat(const double* R_rowmajor, double* xyzw);
void orc_apply(orc_workspace* w, const double* q_joints, const double* root_pose, double* out);
void orc_applyJ(orc_workspace* w, const double* dq_joints, const double* root_vel, double* out);
void orc_applyJT(orc_workspace* w, const double* in, double* out_tau, double* out_root);
void orc_apply_loop(orc_workspace* w, long n, const double* q_joints, const double* root_pose, double* out);
int orc_applyJT_row(orc_workspace* w, int nnz, const int* col, const double* val, double* row_out);
void orc_crba(orc_workspace* w, const double* q, double* M_rowmajor);
void orc_rnea(orc_workspace* w, const double* q, const double* v, const double* a, double* tau);
void orc_aba(orc_workspace* w, const double* q, const double* v, const double* tau, double* ddq);
void orc_rnea_derivatives(orc_workspace* w, const double* q, const double* v, const double* a, double* dtau_dq_rowmajor,
                          double* dtau_dv_rowmajor, double* M_rowmajor_or_null);
void orc_eval(orc_workspace* w, const double* q, const double* v, const double* a, double* oMi, double* J,
              double* Mpacked, double* tau);
int orc_eval_batch(const rbd_model_desc* d, long n, const double* q, const double* v, const double* a,
                   double* oMi, double* J, double* Mpacked, double* tau, int nthreads);
const double* orc_data_J(const orc_workspace* w);
void orc_get_oMi(const orc_workspace* w, int i, double* out12);
void orc_get_oMf(const orc_workspace* w, int f, double* out12);
int orc_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
