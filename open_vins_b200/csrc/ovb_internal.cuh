// ovb_internal.cuh — context, device-side views and launch declarations shared by the .cu files of libovb200.so.
#pragma once
#include "../../include/ovb200.h"
#include <cuda_runtime.h>
#include <stdint.h>

#define OVB_MAX_MEAS_PER_FEAT 128 // K*C of config 4 is 124 (SURVEY.md §8 sizes table)
#define OVB_MAX_COLS 512          // >= 6*OVB_MAX_CLONES + 14*OVB_MAX_CAMS; also the widest H of ovb_ekf_update (config 5: n = 500)
#define OVB_NB 16                 // TSQR panel width
#define OVB_CR 256                // TSQR rows per chunk

// ---- device-resident copy of the slice of State the path reads (uploaded once per update call)
struct DevFrame {
  int n_clones, n_cams;
  int n_slots;        // variables a feature can touch (clones + calibrated extrinsics/intrinsics)
  int n_all;          // canonical column count = sum of slot sizes
  double clone_R[OVB_MAX_CLONES][9];
  double clone_p[OVB_MAX_CLONES][3];
  double clone_R_fej[OVB_MAX_CLONES][9];
  double clone_p_fej[OVB_MAX_CLONES][3];
  double cam_R[OVB_MAX_CAMS][9];
  double cam_p[OVB_MAX_CAMS][3];
  double cam_intr[OVB_MAX_CAMS][8];
  int cam_model[OVB_MAX_CAMS];
  // slots in canonical (ascending covariance offset) order
  int slot_off[OVB_MAX_VARS];  // covariance offset
  int slot_size[OVB_MAX_VARS]; // 6 or 8
  int slot_col[OVB_MAX_VARS];  // first canonical column
  int clone_slot[OVB_MAX_CLONES];
  int cam_ext_slot[OVB_MAX_CAMS];  // -1 when not calibrated
  int cam_intr_slot[OVB_MAX_CAMS]; // -1 when not calibrated
};

// camera-at-clone poses (update/UpdaterMSCKF.cpp:98-115), filled on the device by k_cam_poses
struct DevCamPoses {
  double cc_R[OVB_MAX_CAMS][OVB_MAX_CLONES][9]; // R_GtoCi
  double cc_p[OVB_MAX_CAMS][OVB_MAX_CLONES][3]; // p_CiinG
};

struct DevOpts {
  ovb_opts o;
  double sigma_pix_sq;
  int rep; // effective MSCKF representation (SINGLE remapped, UpdaterMSCKF.cpp:180-183)
};

// per-feature device record (inputs derived on the host while packing + outputs)
struct DevFeat {
  int m0, m1;        // measurement range
  int row0;          // first row in the stacked staging matrix (prefix sum of max(2M-3,0))
  int key0, key1;    // camera-key range
  int status;        // ovb_feat_status
  int anchor_cam, anchor_clone;
  int sched, pad0;   // CTA i of the per-feature kernels works on feature feats[i].sched (longest tracks first)
  double p_FinA[3], p_FinG[3];
  double chi2;
  // SLAM update only (landmark already in the state, update/UpdaterSLAM.cpp:333-341, :389-408)
  double p_FinG_fej[3]; // Landmark::get_xyz(true) for the global representations
  double sigma_sq;      // per-class pixel noise variance
  double chi2_mult;     // per-class gate multiplier
  int lm_slot, pad1;    // slot of the landmark's own 3-wide variable
};

// written by the column-map kernel; read by TSQR re-order, EKF and the host (D2H with the outputs)
struct DevUpdateInfo {
  int n_used;                   // columns of the stacked H in the requested order (ct_jacob)
  int n_feats_used;             // accepted features
  int rows_stacked;             // Σ (2M-3) over accepted features
  int n_order;                  // variables in Hx_order_big
  int order_slot[OVB_MAX_VARS]; // slot id of each variable in stacked order
  int col_state[OVB_MAX_COLS];  // covariance index of each stacked column (order applied)
  int col_canon[OVB_MAX_COLS];  // canonical column each stacked column comes from
  int neg_diag_index;           // EKF: -1 or first negative diagonal
  int not_spd;                  // EKF: Cholesky pivot failure flag
  int nonfinite;
};

// packed measurement blob layout (device): [meas_off int32 (F+1)][cam u8 (M)][pad][clone u16 (M)][pad][uv f32 2M][uvn f32 2M][keys u8]
struct BlobView {
  const uint8_t *cam;
  const uint16_t *clone;
  const float *uv;
  const float *uvn;
  const uint8_t *keys;
};

struct ovb_ctx {
  ovb_config cfg;
  int device;
  cudaStream_t stream;
  int own_stream;
  cudaStream_t side_stream;      // column bookkeeping runs here, concurrently with the compression
  cudaStream_t side_stream2;     // with side_stream: the per-feature kernel's size classes run side by side
  cudaEvent_t ev_fork, ev_join, ev_join2;
  int feat_classes;              // split the per-feature kernel into size classes (OVB_FEAT_CLASSES=0 disables: A/B timing only)
  cudaEvent_t ev[8];
  char err[256];
  // covariance (double buffered for clone/marginalize), row-major with leading dimension ldP
  int N, ldP;
  double *P[2];
  int cur;
  // per-call device inputs
  // one input arena (single H2D copy per call): [DevFrame][DevOpts][DevFeat x max_feats][blob]
  unsigned char *d_arena, *h_arena;
  size_t arena_bytes, off_opts, off_feat, off_blob;
  DevFrame *d_frame;
  DevOpts *d_opts;
  DevFeat *d_feat;
  unsigned char *d_blob; // packed SoA measurements
  DevFrame *h_frame;
  DevOpts *h_opts;
  DevFeat *h_feat;
  unsigned char *h_blob;
  size_t blob_cap;
  DevCamPoses *d_cc;
  unsigned char *d_feat_order; // [max_feats][OVB_MAX_VARS+1] per-feature Hx_order (slot ids, first-seen)
  DevUpdateInfo *d_info;
  DevUpdateInfo *h_info;
  double *h_dx;
  double *h_stage; // pinned staging for dense H / Phi uploads (grown on demand)
  size_t stage_cap;
  // device work buffers
  double *d_chi2_table;
  double *d_Hs; // stacked staging matrix [max_rows][ldH]
  size_t Hs_cap; // doubles
  double *d_W[2]; // TSQR panel ping-pong workspaces
  size_t W_cap;
  double *d_R;   // TSQR output [OVB_MAX_COLS][ldR]
  double *d_R2;  // re-ordered / re-triangularised R
  double *d_M;   // EKF: P[:,cols] H'   [max_state][ldM]
  double *d_S;   // EKF: innovation covariance / Cholesky factor
  double *d_Y;   // EKF: M L^-T
  double *d_w;   // EKF: L^-1 z
  double *d_dx;
  double *d_scratch; // per-CTA scratch for large features in the gate kernel
  size_t scratch_per_cta;
  int scratch_ctas;
  double *d_dump; // debug dumps for ovb_feature_jacobians
  size_t dump_cap;
  int dump_rows;
  int max_rows;
  int sm_count;
  size_t info_bytes; // DevUpdateInfo rounded up: d_dx / h_dx start right behind d_info / h_info
  int attr_done[8]; // per-context (= per-device) one-time cudaFuncSetAttribute flags: 0 tsqr, 1 feature, 2 ekf, 3 gram, 4 cholqr
  int tsqr_pdl;     // programmatic dependent launch between the TSQR level kernels (OVB_TSQR_PDL=0 disables: A/B timing only)
  int tsqr_cluster; // upper TSQR levels as one thread-block cluster (OVB_TSQR_CLUSTER=0 disables: A/B timing only)
  int gram_cluster;  // k_cq_gram: clusters of 4 slabs pre-reduce in distributed shared memory (OVB_GRAM_CLUSTER=0 disables: A/B timing only)
  int ekf_chol_dmma; // EKF Cholesky on the DMMA kernel of k_cholqr.cu (OVB_EKF_CHOL_DMMA=0 disables: A/B timing only)
  mutable float stage_ms[6];
  mutable int stage_pending; // stage_ms[0..4] of the last update not read back from the events yet (done on demand: each read costs ~1.5 us of host time)
  double host_us[4]; // host wall clock of the last ovb_msckf_update: marshalling + H2D enqueue, kernel enqueue, wait, result unpack
  // replay of the last update on device-resident inputs (bench: `value` leg; see ovb_msckf_replay)
  int replay_enabled, last_pk_valid;
  int last_n_feats, last_max_M, last_m_total, last_ldH, last_n_all, last_col_order;
  BlobView last_bv;
  double *P_snap;
  void *d_flush;
  // bookkeeping for bench.py: kernels launched by the last update pipeline, bytes moved by the last ovb_msckf_update
  int n_launch, n_launch_tsqr_level;
  // normal-equations compression (k_gram.cu)
  double *d_Gpart, *d_G;
  size_t Gpart_cap, G_cap;
  double *d_cqw; // wide systems (k_cholqr.cu): [G1 | G2 | packed diagonal-block factors | scalars]
  size_t cqw_cap;
  size_t last_h2d_bytes, last_d2h_bytes;
  // per-kernel profile (ovb_set_profile): CUDA events around every ovb_launch of the main stream; PDL is off while it is on
  int prof_on, prof_n;
  cudaEvent_t prof_ev[2 * 96];
  const void *prof_fn[96];
};

#define OVB_CUDA_CHECK(ctx, call)                                                                                     \
  do {                                                                                                                \
    cudaError_t e_ = (call);                                                                                          \
    if (e_ != cudaSuccess) {                                                                                          \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return OVB_ERR_CUDA;                                                                                            \
    }                                                                                                                 \
  } while (0)

// ---- launchers (each enqueues on ctx->stream; no host sync)
void launch_cam_poses(ovb_ctx *ctx);
void launch_triangulate(ovb_ctx *ctx, int n_feats, BlobView bv);
// mode 0: normal (write post-nullspace rows to Hs, gate on chi²); mode 1: dump pre-nullspace dense rows to d_dump
// mode 2: like 0 but features keep the status/p_FinG given (no triangulation ran) — used by ovb_feature_jacobians
void launch_feature_system(ovb_ctx *ctx, int n_feats, BlobView bv, int ldH, int mode, int max_M);
void launch_column_map(ovb_ctx *ctx, int n_feats, BlobView bv, int rows_drop = 3);
// TSQR of A [m x (n+1)] (last column = residual) in place; R (n x (n+1), diag>=0) to Rout with leading dimension ldR
void launch_tsqr(ovb_ctx *ctx, double *A, int m, int n, int ldA, double *Rout, int ldR);
// gather columns of Rin in the order info->col_canon (n_used of them) into Hs scratch and re-triangularise into Rout
// [R | z] <- chol([H r]'[H r]) (k_gram.cu); returns the number of kernels launched or -1
int launch_compress_gram(ovb_ctx *ctx, const double *A, int m, int n, int ldA, double *Rout, int ldR);
// [R | z] <- shifted CholeskyQR2 of A (k_cholqr.cu; A is overwritten); returns kernels launched, or -1 when n+1 exceeds
// what the single-CTA Cholesky takes (callers then use launch_tsqr)
int launch_compress_cholqr2(ovb_ctx *ctx, double *A, int m, int n, int ldA, double *Rout, int ldR);
// EKF Cholesky on the DMMA kernel of k_cholqr.cu; false when r does not fit (caller uses k_ekf_chol)
bool launch_chol_ekf_dmma(ovb_ctx *ctx, double *S, int ldS, int r, const double *res, double *w, double *invdiag, double **Lpk_out);
// A <- A (L')^-1 for the rows of A [m x nt] with the packed factor of the DMMA Cholesky; false when it does not fit
bool launch_trsm_rows(ovb_ctx *ctx, double *A, int ldA, int m, int nt, const double *Lpk);
// wide EKF (r > 160): blocked DMMA Cholesky of S (r x r lower, residual staged in w) and Y = M L^-T in place; false when r
// exceeds OVB_MAX_COLS or the leading dimensions are odd
bool launch_chol_solve_wide(ovb_ctx *ctx, double *S, int ldS, int r, double *w, double *invdiag, double *M, int ldM, int N, bool gate_only);
void launch_reorder_R(ovb_ctx *ctx, const double *Rin, int n_all, int ldRin, double *Rout, int ldRout);
// EKF update from an upper-trapezoidal / dense H [r x n] with column->state map in d_info (device-side sizes)
void launch_ekf_update(ovb_ctx *ctx, const double *H, int ldHm, int r_max, int n_max, bool sizes_from_info, double sigma2,
                       const double *Rdiag_dev);
// false: the launch was refused (shared-memory footprint) or failed; the caller must not grow N
bool launch_cov_init_augment(ovb_ctx *ctx, int k, int n, const double *Hx_dev, const double *Hinv_dev, double sigma2);
void launch_cov_clone(ovb_ctx *ctx, int old_off, int size, const double *dnc_dt_dev, int dt_off);
void launch_cov_marginalize(ovb_ctx *ctx, int off, int size);
void launch_cov_propagate(ovb_ctx *ctx, int new_off, int p, int q, const int *old_idx_dev, const double *Phi_dev, const double *Q_dev);

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------------
// Every kernel of the update pipeline starts with OVB_PDL_ENTER(): it lets the NEXT kernel of the stream be scheduled
// while this one runs (its CTAs become resident on idle SMs and block), then waits until the PREVIOUS kernel has
// completed and flushed. Semantics are those of ordinary stream order; what is saved is the launch latency between the
// ~30 dependent, latency-bound kernels of one update. Both instructions are no-ops in a kernel launched without the
// attribute. ovb_launch() adds the attribute when ctx->tsqr_pdl is set (OVB_TSQR_PDL=0 disables it).
#define OVB_PDL_ENTER()                                                     \
  do {                                                                      \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");         \
    asm volatile("griddepcontrol.wait;" ::: "memory");                      \
  } while (0)

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline void ovb_launch(ovb_ctx *ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (ctx->tsqr_pdl && !ctx->prof_on) ? 1 : 0;
  const bool prof = ctx->prof_on && ctx->prof_n < 96 && ctx->prof_ev[0] != nullptr;
  if (prof)
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream);
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (prof) {
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n + 1], ctx->stream);
    ctx->prof_fn[ctx->prof_n++] = (const void *)kern;
  }
}
#endif
