// ovb_api.cu — the C ABI of include/ovb200.h: context, host-side marshalling into one pinned arena, stream orchestration.
// No arithmetic of the path happens on the host: it packs the inputs, launches the kernels of k_*.cu and copies results back.
#include "ovb_internal.cuh"
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

static const double g_chi2_table[OVB_CHI2_TABLE_LEN] = {
#include "chi2_table.inc"
};

unsigned char *ovb_feat_order_ptr(ovb_ctx *ctx) { return ctx->d_feat_order; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" {

static ovb_status ensure_stage(ovb_ctx *ctx, size_t doubles);

int ovb_abi_version(void) { return OVB_ABI_VERSION; }

double ovb_chi2_quantile95(int dof) {
  if (dof < 1)
    return 0.0;
  if (dof >= OVB_CHI2_TABLE_LEN)
    dof = OVB_CHI2_TABLE_LEN - 1;
  return g_chi2_table[dof];
}

void ovb_opts_default(ovb_opts *o) {
  memset(o, 0, sizeof(*o));
  o->triangulate_1d = 0;
  o->refine_features = 1;
  o->max_runs = 5;
  o->init_lamda = 1e-3;
  o->max_lamda = 1e10;
  o->min_dx = 1e-6;
  o->min_dcost = 1e-6;
  o->lam_mult = 10;
  o->min_dist = 0.10;
  o->max_dist = 60;
  o->max_baseline = 40;
  o->max_cond_number = 10000;
  o->sigma_pix = 1;
  o->chi2_multipler = 5;
  o->do_fej = 1;
  o->feat_rep = OVB_REP_GLOBAL_3D;
  o->do_calib_camera_pose = 0;
  o->do_calib_camera_intrinsics = 0;
  o->col_order = OVB_COLS_CANONICAL;
  o->compress = OVB_COMPRESS_CHOLQR2;
}

const char *ovb_last_error(const ovb_ctx *ctx) { return ctx ? ctx->err : "null context"; }

ovb_status ovb_create(const ovb_config *cfg, ovb_ctx **out) {
  if (!cfg || !out)
    return OVB_ERR_ARG;
  *out = nullptr;
  ovb_ctx *ctx = new (std::nothrow) ovb_ctx();
  if (!ctx)
    return OVB_ERR_CAPACITY;
  memset(ctx, 0, sizeof(*ctx));
  ctx->cfg = *cfg;
  if (ctx->cfg.max_state < 32)
    ctx->cfg.max_state = 32;
  if (ctx->cfg.max_feats < 1)
    ctx->cfg.max_feats = 1;
  if (ctx->cfg.max_meas < 2)
    ctx->cfg.max_meas = 2;
  ctx->device = cfg->device;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) {
    delete ctx;
    return OVB_ERR_CUDA; // no CUDA device: there is no CPU fallback
  }
#define CK(call)                                                                                                      \
  do {                                                                                                                \
    cudaError_t e_ = (call);                                                                                          \
    if (e_ != cudaSuccess) {                                                                                          \
      fprintf(stderr, "ovb_create: %s failed: %s\n", #call, cudaGetErrorString(e_));                                  \
      ovb_destroy(ctx);                                                                                               \
      return OVB_ERR_CUDA;                                                                                            \
    }                                                                                                                 \
  } while (0)
  CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  ctx->sm_count = prop.multiProcessorCount;
  {
    const char *e = getenv("OVB_TSQR_CLUSTER");
    ctx->tsqr_cluster = e ? atoi(e) : 1;
    const char *e2 = getenv("OVB_TSQR_PDL");
    ctx->tsqr_pdl = e2 ? atoi(e2) : 1;
    const char *e3 = getenv("OVB_FEAT_CLASSES");
    ctx->feat_classes = e3 ? atoi(e3) : 1;
    const char *e4 = getenv("OVB_EKF_CHOL_DMMA");
    ctx->ekf_chol_dmma = e4 ? atoi(e4) : 1;
    const char *e5 = getenv("OVB_GRAM_CLUSTER");
    ctx->gram_cluster = e5 ? atoi(e5) : 0; // measured slower on B200 (37 clusters of 4 do not co-schedule on 148 SMs: second wave), profiles/README.md
  }
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = 1;
  CK(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&ctx->side_stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < 8; i++)
    CK(cudaEventCreate(&ctx->ev[i]));
  const int ms = ctx->cfg.max_state;
  ctx->ldP = ms;
  ctx->N = 0;
  ctx->cur = 0;
  CK(cudaMalloc(&ctx->P[0], sizeof(double) * (size_t)ms * ms));
  CK(cudaMalloc(&ctx->P[1], sizeof(double) * (size_t)ms * ms));
  CK(cudaMemset(ctx->P[0], 0, sizeof(double) * (size_t)ms * ms));
  CK(cudaMemset(ctx->P[1], 0, sizeof(double) * (size_t)ms * ms));
  // input arena
  ctx->off_opts = align_up(sizeof(DevFrame), 256);
  ctx->off_feat = ctx->off_opts + align_up(sizeof(DevOpts), 256);
  ctx->off_blob = ctx->off_feat + align_up(sizeof(DevFeat) * (size_t)ctx->cfg.max_feats, 256);
  ctx->blob_cap = align_up((size_t)ctx->cfg.max_meas * (1 + 2 + 8 + 8) + (size_t)ctx->cfg.max_feats * OVB_MAX_CAMS + 64, 256);
  ctx->arena_bytes = ctx->off_blob + ctx->blob_cap;
  CK(cudaMalloc(&ctx->d_arena, ctx->arena_bytes));
  CK(cudaMallocHost(&ctx->h_arena, ctx->arena_bytes));
  ctx->d_frame = (DevFrame *)ctx->d_arena;
  ctx->d_opts = (DevOpts *)(ctx->d_arena + ctx->off_opts);
  ctx->d_feat = (DevFeat *)(ctx->d_arena + ctx->off_feat);
  ctx->d_blob = ctx->d_arena + ctx->off_blob;
  ctx->h_frame = (DevFrame *)ctx->h_arena;
  ctx->h_opts = (DevOpts *)(ctx->h_arena + ctx->off_opts);
  ctx->h_feat = (DevFeat *)(ctx->h_arena + ctx->off_feat);
  ctx->h_blob = ctx->h_arena + ctx->off_blob;
  CK(cudaMalloc(&ctx->d_cc, sizeof(DevCamPoses)));
  CK(cudaMalloc(&ctx->d_feat_order, (size_t)ctx->cfg.max_feats * (OVB_MAX_VARS + 1)));
  // update info + dx live in ONE device block and ONE pinned host block: a single D2H copy returns both
  ctx->info_bytes = align_up(sizeof(DevUpdateInfo), 256);
  CK(cudaMalloc(&ctx->d_info, ctx->info_bytes + sizeof(double) * (size_t)ms));
  ctx->d_dx = (double *)((char *)ctx->d_info + ctx->info_bytes);
  CK(cudaMallocHost(&ctx->h_info, ctx->info_bytes + sizeof(double) * (size_t)ms));
  ctx->h_dx = (double *)((char *)ctx->h_info + ctx->info_bytes);
  CK(cudaMalloc(&ctx->d_chi2_table, sizeof(g_chi2_table)));
  CK(cudaMemcpy(ctx->d_chi2_table, g_chi2_table, sizeof(g_chi2_table), cudaMemcpyHostToDevice));
  // stacked staging matrix
  ctx->max_rows = ctx->cfg.max_rows > 0 ? ctx->cfg.max_rows : 2 * ctx->cfg.max_meas;
  if (ctx->max_rows < 2 * ctx->cfg.max_meas)
    ctx->max_rows = 2 * ctx->cfg.max_meas;
  int ldcap = std::min(OVB_MAX_COLS, ms) + 8;
  ctx->Hs_cap = (size_t)ctx->max_rows * ldcap;
  CK(cudaMalloc(&ctx->d_Hs, sizeof(double) * ctx->Hs_cap));
  ctx->h_stage = nullptr; // pinned dense staging is grown on demand (ensure_stage)
  ctx->stage_cap = 0;
  ctx->W_cap = ((size_t)ctx->max_rows / OVB_CR + 2 + (size_t)ctx->sm_count) * OVB_NB * OVB_NB + 4096; // level-0 chunks can be as short as max_rows / sm_count
  CK(cudaMalloc(&ctx->d_W[0], sizeof(double) * ctx->W_cap));
  CK(cudaMalloc(&ctx->d_W[1], sizeof(double) * ctx->W_cap));
  size_t rsz = (size_t)(ms + 8) * (ms + 8);
  CK(cudaMalloc(&ctx->d_R, sizeof(double) * rsz));
  CK(cudaMalloc(&ctx->d_R2, sizeof(double) * rsz));
  CK(cudaMalloc(&ctx->d_M, sizeof(double) * (size_t)ms * ms));
  CK(cudaMalloc(&ctx->d_S, sizeof(double) * (size_t)(ms + 1) * ms));
  CK(cudaMalloc(&ctx->d_Y, sizeof(double) * (size_t)ms * ms));
  CK(cudaMalloc(&ctx->d_w, sizeof(double) * (size_t)ms * 4));
  ctx->scratch_per_cta = (size_t)(2 * OVB_MAX_MEAS_PER_FEAT + 1) * (2 * OVB_MAX_MEAS_PER_FEAT + 1);
  ctx->scratch_ctas = 2 * ctx->sm_count;
  CK(cudaMalloc(&ctx->d_scratch, sizeof(double) * ctx->scratch_per_cta * ctx->scratch_ctas));
  ctx->dump_cap = 0;
  ctx->d_dump = nullptr; // allocated on first use by ovb_feature_jacobians(stage 0)
#undef CK
  *out = ctx;
  return OVB_OK;
}

void ovb_destroy(ovb_ctx *ctx) {
  if (ctx)
    for (int i = 0; i < 2 * 96; i++)
      if (ctx->prof_ev[i])
        cudaEventDestroy(ctx->prof_ev[i]);
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  if (ctx->stream)
    cudaStreamSynchronize(ctx->stream);
  void *dev[] = {ctx->P[0],   ctx->P[1], ctx->d_arena, ctx->d_cc, ctx->d_feat_order, ctx->d_info, ctx->d_chi2_table, ctx->d_Hs, ctx->d_W[0],
                 ctx->d_W[1], ctx->d_R,  ctx->d_R2,    ctx->d_M,  ctx->d_S,          ctx->d_Y,    ctx->d_w,          ctx->d_scratch,
                 ctx->d_dump, ctx->P_snap, ctx->d_flush, ctx->d_Gpart, ctx->d_G, ctx->d_cqw};
  for (void *p : dev)
    if (p)
      cudaFree(p);
  void *host[] = {ctx->h_arena, ctx->h_info, ctx->h_stage};
  for (void *p : host)
    if (p)
      cudaFreeHost(p);
  for (int i = 0; i < 8; i++)
    if (ctx->ev[i])
      cudaEventDestroy(ctx->ev[i]);
  if (ctx->ev_fork)
    cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join)
    cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_join2)
    cudaEventDestroy(ctx->ev_join2);
  if (ctx->side_stream2)
    cudaStreamDestroy(ctx->side_stream2);
  if (ctx->side_stream)
    cudaStreamDestroy(ctx->side_stream);
  if (ctx->stream && ctx->own_stream)
    cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// ------------------------------------------------------------------------------------------------ covariance residency
int ovb_cov_dim(const ovb_ctx *ctx) { return ctx ? ctx->N : 0; }

ovb_status ovb_cov_set(ovb_ctx *ctx, const double *P, int N) {
  if (!ctx || !P || N < 1)
    return OVB_ERR_ARG;
  if (N > ctx->cfg.max_state) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_cov_set: N=%d exceeds max_state=%d", N, ctx->cfg.max_state);
    return OVB_ERR_CAPACITY;
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  OVB_CUDA_CHECK(ctx, cudaMemcpy2DAsync(ctx->P[ctx->cur], sizeof(double) * ctx->ldP, P, sizeof(double) * N, sizeof(double) * N, N,
                                        cudaMemcpyHostToDevice, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->N = N;
  return OVB_OK;
}

ovb_status ovb_cov_get(ovb_ctx *ctx, double *P, int N) {
  if (!ctx || !P || N != ctx->N)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  OVB_CUDA_CHECK(ctx, cudaMemcpy2DAsync(P, sizeof(double) * N, ctx->P[ctx->cur], sizeof(double) * ctx->ldP, sizeof(double) * N, N,
                                        cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return OVB_OK;
}

ovb_status ovb_cov_get_marginal(ovb_ctx *ctx, const int *off, const int *sz, int nvar, double *out) {
  if (!ctx || !off || !sz || !out || nvar < 1)
    return OVB_ERR_ARG;
  int n = 0;
  for (int i = 0; i < nvar; i++) {
    if (off[i] < 0 || sz[i] < 1 || off[i] + sz[i] > ctx->N)
      return OVB_ERR_ARG;
    n += sz[i];
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  // block copies (small): one strided copy per (i,k) block
  int ii = 0;
  for (int i = 0; i < nvar; i++) {
    int kk = 0;
    for (int k = 0; k < nvar; k++) {
      OVB_CUDA_CHECK(ctx, cudaMemcpy2DAsync(out + (size_t)ii * n + kk, sizeof(double) * n,
                                            ctx->P[ctx->cur] + (size_t)off[i] * ctx->ldP + off[k], sizeof(double) * ctx->ldP,
                                            sizeof(double) * sz[k], sz[i], cudaMemcpyDeviceToHost, ctx->stream));
      kk += sz[k];
    }
    ii += sz[i];
  }
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return OVB_OK;
}

ovb_status ovb_cov_clone(ovb_ctx *ctx, int old_off, int size, const double *dnc_dt, int dt_off) {
  if (!ctx || size < 1 || old_off < 0 || old_off + size > ctx->N)
    return OVB_ERR_ARG;
  if (ctx->N + size > ctx->cfg.max_state) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_cov_clone: N+size=%d exceeds max_state=%d", ctx->N + size, ctx->cfg.max_state);
    return OVB_ERR_CAPACITY;
  }
  if (dnc_dt && (dt_off < 0 || dt_off >= ctx->N || size > 64))
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const double *dnc_dev = nullptr;
  if (dnc_dt) {
    memcpy(ctx->h_dx, dnc_dt, sizeof(double) * size);
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_w, ctx->h_dx, sizeof(double) * size, cudaMemcpyHostToDevice, ctx->stream));
    dnc_dev = ctx->d_w;
  }
  launch_cov_clone(ctx, old_off, size, dnc_dev, dt_off);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->N += size;
  return OVB_OK;
}

ovb_status ovb_cov_marginalize(ovb_ctx *ctx, int off, int size) {
  if (!ctx || size < 1 || off < 0 || off + size > ctx->N)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  launch_cov_marginalize(ctx, off, size);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->cur ^= 1;
  ctx->N -= size;
  return OVB_OK;
}

ovb_status ovb_cov_propagate(ovb_ctx *ctx, int new_off, int p, const int *old_off, const int *old_sz, int nold, const double *Phi,
                             const double *Q) {
  if (!ctx || !old_off || !old_sz || !Phi || !Q || p < 1 || nold < 1 || new_off < 0 || new_off + p > ctx->N)
    return OVB_ERR_ARG; // the reference exits on empty variable lists (StateHelper.cpp:41-44)
  int q = 0;
  for (int i = 0; i < nold; i++) {
    if (old_off[i] < 0 || old_sz[i] < 1 || old_off[i] + old_sz[i] > ctx->N)
      return OVB_ERR_ARG;
    q += old_sz[i];
  }
  if (q > ctx->cfg.max_state || p > ctx->cfg.max_state || (size_t)p * q + (size_t)p * p > ctx->Hs_cap)
    return OVB_ERR_CAPACITY;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  // stage: [Phi p*q][Q p*p] doubles, then q ints
  {
    ovb_status es = ensure_stage(ctx, (size_t)p * q + (size_t)p * p + (size_t)q + 16);
    if (es != OVB_OK)
      return es;
  }
  double *hs = ctx->h_stage;
  memcpy(hs, Phi, sizeof(double) * (size_t)p * q);
  memcpy(hs + (size_t)p * q, Q, sizeof(double) * (size_t)p * p);
  int *hidx = (int *)(hs + (size_t)p * q + (size_t)p * p);
  int c = 0;
  for (int i = 0; i < nold; i++)
    for (int k = 0; k < old_sz[i]; k++)
      hidx[c++] = old_off[i] + k;
  size_t bytes = sizeof(double) * ((size_t)p * q + (size_t)p * p) + sizeof(int) * q;
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_Hs, hs, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const double *Phi_dev = ctx->d_Hs;
  const double *Q_dev = ctx->d_Hs + (size_t)p * q;
  const int *idx_dev = (const int *)(ctx->d_Hs + (size_t)p * q + (size_t)p * p);
  launch_cov_propagate(ctx, new_off, p, q, idx_dev, Phi_dev, Q_dev);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(DevUpdateInfo), cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->h_info->neg_diag_index != 0x7fffffff) {
    snprintf(ctx->err, sizeof(ctx->err), "EKFPropagation: diagonal at %d is negative", ctx->h_info->neg_diag_index);
    return OVB_ERR_NEG_DIAG;
  }
  return OVB_OK;
}

// ------------------------------------------------------------------------------------------------ marshalling
struct Packed {
  int n_feats, n_meas, max_M, m_total, ldH, n_all;
  BlobView bv;
};

// lm != nullptr: SLAM batch — every feature brings its landmark (a 3-wide state variable that becomes one more slot),
// rows are NOT nullspace-projected (2M per feature instead of 2M-3), values/anchors come from the landmark.
static ovb_status pack_inputs(ovb_ctx *ctx, const ovb_frame *fr, const ovb_feat_batch *fb, const ovb_opts *op, const ovb_feat_out *given,
                              Packed *pk, const ovb_landmarks *lm = nullptr) {
  if (!fr || !fb || !op)
    return OVB_ERR_ARG;
  if (lm && (!lm->lm_off || !lm->value || !lm->value_fej))
    return OVB_ERR_ARG;
  if (fr->n_clones < 1 || fr->n_clones > OVB_MAX_CLONES || fr->n_cams < 1 || fr->n_cams > OVB_MAX_CAMS) {
    snprintf(ctx->err, sizeof(ctx->err), "frame: n_clones=%d (max %d) n_cams=%d (max %d)", fr->n_clones, OVB_MAX_CLONES, fr->n_cams, OVB_MAX_CAMS);
    return OVB_ERR_CAPACITY;
  }
  if (fb->n_feats < 0 || fb->n_feats > ctx->cfg.max_feats || fb->n_meas < 0 || fb->n_meas > ctx->cfg.max_meas) {
    snprintf(ctx->err, sizeof(ctx->err), "batch: n_feats=%d (max %d) n_meas=%d (max %d)", fb->n_feats, ctx->cfg.max_feats, fb->n_meas,
             ctx->cfg.max_meas);
    return OVB_ERR_CAPACITY;
  }
  DevFrame *hf = ctx->h_frame;
  memset(hf, 0, sizeof(*hf));
  hf->n_clones = fr->n_clones;
  hf->n_cams = fr->n_cams;
  memcpy(hf->clone_R, fr->clone_R, sizeof(double) * 9 * fr->n_clones);
  memcpy(hf->clone_p, fr->clone_p, sizeof(double) * 3 * fr->n_clones);
  memcpy(hf->clone_R_fej, fr->clone_R_fej ? fr->clone_R_fej : fr->clone_R, sizeof(double) * 9 * fr->n_clones);
  memcpy(hf->clone_p_fej, fr->clone_p_fej ? fr->clone_p_fej : fr->clone_p, sizeof(double) * 3 * fr->n_clones);
  memcpy(hf->cam_R, fr->cam_R, sizeof(double) * 9 * fr->n_cams);
  memcpy(hf->cam_p, fr->cam_p, sizeof(double) * 3 * fr->n_cams);
  memcpy(hf->cam_intr, fr->cam_intr, sizeof(double) * 8 * fr->n_cams);
  for (int k = 0; k < fr->n_cams; k++)
    hf->cam_model[k] = fr->cam_model ? fr->cam_model[k] : OVB_CAM_RADTAN;
  // ---- slots in ascending covariance offset
  struct SlotRec {
    int off, size, kind, idx;
  }; // kind 0 clone, 1 ext, 2 intr, 3 landmark
  std::vector<SlotRec> slots;
  std::vector<int> lm_slot_of((size_t)(lm ? fb->n_feats : 0), -1);
  const bool lm_single = lm && op->feat_rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE; // 1-wide landmark (inverse depth only)
  if (lm)
    for (int f = 0; f < fb->n_feats; f++)
      slots.push_back({lm->lm_off[f], lm_single ? 1 : 3, 3, f});
  for (int k = 0; k < fr->n_cams; k++) {
    hf->cam_ext_slot[k] = hf->cam_intr_slot[k] = -1;
    if (op->do_calib_camera_pose) {
      if (!fr->cam_ext_off || fr->cam_ext_off[k] < 0) {
        snprintf(ctx->err, sizeof(ctx->err), "do_calib_camera_pose set but cam_ext_off[%d] < 0", k);
        return OVB_ERR_ARG;
      }
      slots.push_back({fr->cam_ext_off[k], 6, 1, k});
    }
    if (op->do_calib_camera_intrinsics) {
      if (!fr->cam_intr_off || fr->cam_intr_off[k] < 0) {
        snprintf(ctx->err, sizeof(ctx->err), "do_calib_camera_intrinsics set but cam_intr_off[%d] < 0", k);
        return OVB_ERR_ARG;
      }
      slots.push_back({fr->cam_intr_off[k], 8, 2, k});
    }
  }
  for (int c = 0; c < fr->n_clones; c++)
    slots.push_back({fr->clone_off[c], 6, 0, c});
  std::sort(slots.begin(), slots.end(), [](const SlotRec &a, const SlotRec &b) { return a.off < b.off; });
  if ((int)slots.size() > OVB_MAX_VARS) {
    snprintf(ctx->err, sizeof(ctx->err), "%d state variables in one update (clones + calibration + landmarks), max %d: split the batch",
             (int)slots.size(), OVB_MAX_VARS);
    return OVB_ERR_CAPACITY;
  }
  int col = 0;
  for (size_t s = 0; s < slots.size(); s++) {
    if (slots[s].off < 0 || slots[s].off + slots[s].size > ctx->N) {
      snprintf(ctx->err, sizeof(ctx->err), "variable offset %d(+%d) outside the covariance (N=%d)", slots[s].off, slots[s].size, ctx->N);
      return OVB_ERR_ARG;
    }
    if (s > 0 && slots[s].off < slots[s - 1].off + slots[s - 1].size) {
      snprintf(ctx->err, sizeof(ctx->err), "overlapping state variables at offset %d", slots[s].off);
      return OVB_ERR_ARG;
    }
    hf->slot_off[s] = slots[s].off;
    hf->slot_size[s] = slots[s].size;
    hf->slot_col[s] = col;
    col += slots[s].size;
    if (slots[s].kind == 0)
      hf->clone_slot[slots[s].idx] = (int)s;
    else if (slots[s].kind == 1)
      hf->cam_ext_slot[slots[s].idx] = (int)s;
    else if (slots[s].kind == 2)
      hf->cam_intr_slot[slots[s].idx] = (int)s;
    else
      lm_slot_of[(size_t)slots[s].idx] = (int)s;
  }
  hf->n_slots = (int)slots.size();
  hf->n_all = col;
  if (col > OVB_MAX_COLS || col + 1 > std::min(OVB_MAX_COLS, ctx->cfg.max_state) + 8)
    return OVB_ERR_CAPACITY;
  // ---- options
  DevOpts *ho = ctx->h_opts;
  ho->o = *op;
  ho->sigma_pix_sq = std::pow(op->sigma_pix, 2);
  ho->rep = op->feat_rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE ? OVB_REP_ANCHORED_MSCKF_INVERSE_DEPTH : op->feat_rep;
  if (ho->rep < 0 || ho->rep > OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE)
    return OVB_ERR_ARG;
  // ---- blob: [cam u8 M][pad2][clone u16 M][pad4][uv f32 2M][uvn f32 2M][keys u8]
  const int F = fb->n_feats, M = fb->n_meas;
  size_t o_cam = 0;
  size_t o_clone = align_up(o_cam + (size_t)M, 16);
  size_t o_uv = align_up(o_clone + 2 * (size_t)M, 16);
  size_t o_uvn = align_up(o_uv + 8 * (size_t)M, 16);
  size_t o_keys = align_up(o_uvn + 8 * (size_t)M, 16);
  unsigned char *hb = ctx->h_blob;
  if (M > 0) {
    memcpy(hb + o_cam, fb->cam, (size_t)M);
    memcpy(hb + o_clone, fb->clone, 2 * (size_t)M);
    memcpy(hb + o_uv, fb->uv, 8 * (size_t)M);
    memcpy(hb + o_uvn, fb->uvn, 8 * (size_t)M);
  }
  unsigned char *hkeys = hb + o_keys;
  size_t nkeys = 0;
  int row = 0, maxM = 0;
  for (int f = 0; f < F; f++) {
    DevFeat &d = ctx->h_feat[f];
    d.m0 = fb->meas_off[f];
    d.m1 = fb->meas_off[f + 1];
    if (d.m0 < 0 || d.m1 < d.m0 || d.m1 > M)
      return OVB_ERR_ARG;
    int Mf = d.m1 - d.m0;
    if (Mf > OVB_MAX_MEAS_PER_FEAT) {
      snprintf(ctx->err, sizeof(ctx->err), "feature %d has %d measurements (max %d)", f, Mf, OVB_MAX_MEAS_PER_FEAT);
      return OVB_ERR_CAPACITY;
    }
    maxM = std::max(maxM, Mf);
    d.row0 = row;
    if (lm)
      row += lm_single ? (Mf >= 2 ? 2 * Mf - 2 : 0) : 2 * Mf; // UpdaterSLAM.cpp:344-387: all 2M rows kept (SINGLE: 2 projected out)
    else
      row += Mf >= 2 ? 2 * Mf - 3 : 0;
    d.key0 = (int)nkeys;
    // validate BEFORE anything is stored in the pinned blob: camera ids in range, at most one key per camera (the
    // reference's feat->timestamps is a map keyed by camera id), room for the keys
    for (int i = d.m0; i < d.m1; i++)
      if (fb->cam[i] >= fr->n_cams || fb->clone[i] >= fr->n_clones) {
        snprintf(ctx->err, sizeof(ctx->err), "feature %d: measurement %d refers to camera %d / clone %d outside the frame", f, i, (int)fb->cam[i],
                 (int)fb->clone[i]);
        return OVB_ERR_ARG;
      }
    if (o_keys + nkeys + (size_t)OVB_MAX_CAMS > ctx->blob_cap)
      return OVB_ERR_CAPACITY;
    {
      unsigned seen = 0;
      if (fb->cam_keys_off && fb->cam_keys) {
        const int k0 = fb->cam_keys_off[f], k1 = fb->cam_keys_off[f + 1];
        if (k0 < 0 || k1 < k0 || k1 - k0 > fr->n_cams) {
          snprintf(ctx->err, sizeof(ctx->err), "feature %d: %d camera keys for %d cameras", f, k1 - k0, fr->n_cams);
          return OVB_ERR_ARG;
        }
        for (int k = k0; k < k1; k++) {
          const unsigned c = fb->cam_keys[k];
          if ((int)c >= fr->n_cams || (seen >> c) & 1u) {
            snprintf(ctx->err, sizeof(ctx->err), "feature %d: camera key %u out of range or repeated", f, c);
            return OVB_ERR_ARG;
          }
          seen |= 1u << c;
          hkeys[nkeys++] = (unsigned char)c;
        }
      } else {
        int last = -1;
        for (int i = d.m0; i < d.m1; i++)
          if ((int)fb->cam[i] != last) {
            last = fb->cam[i];
            if ((seen >> last) & 1u) { // measurements must be grouped by camera (include/ovb200.h, ovb_feat_batch)
              snprintf(ctx->err, sizeof(ctx->err), "feature %d: measurements are not grouped by camera", f);
              return OVB_ERR_ARG;
            }
            seen |= 1u << last;
            hkeys[nkeys++] = (unsigned char)last;
          }
      }
    }
    d.key1 = (int)nkeys;
    for (int i = d.m0; i < d.m1; i++)
      if (fb->cam[i] >= fr->n_cams || fb->clone[i] >= fr->n_clones)
        return OVB_ERR_ARG;
    d.status = OVB_FEAT_OK;
    d.anchor_cam = d.anchor_clone = -1;
    d.chi2 = NAN;
    for (int k = 0; k < 3; k++)
      d.p_FinA[k] = d.p_FinG[k] = NAN;
    d.sigma_sq = std::pow(op->sigma_pix, 2);
    d.chi2_mult = op->chi2_multipler;
    d.lm_slot = -1;
    for (int k = 0; k < 3; k++)
      d.p_FinG_fej[k] = NAN;
    if (lm) {
      d.lm_slot = lm_slot_of[(size_t)f];
      d.status = Mf >= (lm_single ? 2 : 1) ? OVB_FEAT_OK : OVB_FEAT_FEW_MEAS; // UpdaterSLAM.cpp:278-290
      const bool rel = op->feat_rep >= OVB_REP_ANCHORED_3D;
      d.anchor_cam = rel && lm->anchor_cam ? lm->anchor_cam[f] : -1;
      d.anchor_clone = rel && lm->anchor_clone ? lm->anchor_clone[f] : -1;
      if (rel && (d.anchor_cam < 0 || d.anchor_cam >= fr->n_cams || d.anchor_clone < 0 || d.anchor_clone >= fr->n_clones)) {
        snprintf(ctx->err, sizeof(ctx->err), "landmark %d: anchored representation without a valid anchor", f);
        return OVB_ERR_ARG;
      }
      for (int k = 0; k < 3; k++) {
        d.p_FinA[k] = lm->value[3 * f + k]; // meaning depends on the representation: the kernel reads the right one
        d.p_FinG[k] = lm->value[3 * f + k];
        d.p_FinG_fej[k] = lm->value_fej[3 * f + k];
      }
      if (lm->sigma_pix)
        d.sigma_sq = std::pow(lm->sigma_pix[f], 2);
      if (lm->chi2_multipler)
        d.chi2_mult = lm->chi2_multipler[f];
    }
    if (given) {
      d.status = given->status ? given->status[f] : OVB_FEAT_OK;
      d.anchor_cam = given->anchor_cam ? given->anchor_cam[f] : -1;
      d.anchor_clone = given->anchor_clone ? given->anchor_clone[f] : -1;
      for (int k = 0; k < 3; k++) {
        d.p_FinA[k] = given->p_FinA ? given->p_FinA[3 * f + k] : NAN;
        d.p_FinG[k] = given->p_FinG ? given->p_FinG[3 * f + k] : NAN;
      }
    }
  }
  {
    // CTA schedule of the per-feature kernels: longest tracks first (counting sort on the track length, stable)
    int cnt[OVB_MAX_MEAS_PER_FEAT + 2] = {0};
    for (int f = 0; f < F; f++)
      cnt[OVB_MAX_MEAS_PER_FEAT - (ctx->h_feat[f].m1 - ctx->h_feat[f].m0) + 1]++;
    for (int i = 1; i <= OVB_MAX_MEAS_PER_FEAT + 1; i++)
      cnt[i] += cnt[i - 1];
    for (int f = 0; f < F; f++)
      ctx->h_feat[cnt[OVB_MAX_MEAS_PER_FEAT - (ctx->h_feat[f].m1 - ctx->h_feat[f].m0)]++].sched = f;
  }
  pk->n_feats = F;
  pk->n_meas = M;
  pk->max_M = maxM;
  pk->m_total = row;
  pk->n_all = col;
  pk->ldH = (int)align_up((size_t)col + 1, 4);
  if ((size_t)std::max(row, col) * pk->ldH > ctx->Hs_cap || row > ctx->max_rows) {
    snprintf(ctx->err, sizeof(ctx->err), "stacked system %d x %d exceeds the reserved staging matrix", row, pk->ldH);
    return OVB_ERR_CAPACITY;
  }
  pk->bv.cam = ctx->d_blob + o_cam;
  pk->bv.clone = (const uint16_t *)(ctx->d_blob + o_clone);
  pk->bv.uv = (const float *)(ctx->d_blob + o_uv);
  pk->bv.uvn = (const float *)(ctx->d_blob + o_uvn);
  pk->bv.keys = ctx->d_blob + o_keys;
  size_t used = ctx->off_blob + o_keys + nkeys;
  ctx->last_h2d_bytes = used;
  cudaError_t e = cudaMemcpyAsync(ctx->d_arena, ctx->h_arena, used, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    snprintf(ctx->err, sizeof(ctx->err), "H2D arena copy: %s", cudaGetErrorString(e));
    return OVB_ERR_CUDA;
  }
  return OVB_OK;
}

static void unpack_feats(ovb_ctx *ctx, int F, ovb_feat_out *out) {
  if (!out)
    return;
  for (int f = 0; f < F; f++) {
    const DevFeat &d = ctx->h_feat[f];
    if (out->status)
      out->status[f] = d.status;
    if (out->anchor_cam)
      out->anchor_cam[f] = d.anchor_cam;
    if (out->anchor_clone)
      out->anchor_clone[f] = d.anchor_clone;
    if (out->chi2)
      out->chi2[f] = d.chi2;
    for (int k = 0; k < 3; k++) {
      if (out->p_FinA)
        out->p_FinA[3 * f + k] = d.p_FinA[k];
      if (out->p_FinG)
        out->p_FinG[3 * f + k] = d.p_FinG[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ hot path
ovb_status ovb_triangulate(ovb_ctx *ctx, const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, ovb_feat_out *out) {
  if (!ctx || !out)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  Packed pk;
  int saveN = ctx->N;
  if (ctx->N == 0)
    ctx->N = ctx->cfg.max_state; // offsets are not dereferenced by this stage
  ovb_status st = pack_inputs(ctx, frame, feats, opts, nullptr, &pk);
  ctx->N = saveN;
  if (st != OVB_OK)
    return st;
  launch_cam_poses(ctx);
  launch_triangulate(ctx, pk.n_feats, pk.bv);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_feat, ctx->d_feat, sizeof(DevFeat) * (size_t)pk.n_feats, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unpack_feats(ctx, pk.n_feats, out);
  return OVB_OK;
}

ovb_status ovb_feature_jacobians(ovb_ctx *ctx, const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, ovb_feat_out *out,
                                 int stage, double *Hf_out, double *Hx_out, double *res_out, int32_t *row_off_out, int32_t *ncols_out,
                                 int32_t *col_index_out, int ld_out) {
  if (!ctx || !out || !row_off_out || !ncols_out || !col_index_out)
    return OVB_ERR_ARG;
  if (ctx->N < 1) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_feature_jacobians: no covariance loaded (ovb_cov_set)");
    return OVB_ERR_ARG;
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  Packed pk;
  ovb_status st = pack_inputs(ctx, frame, feats, opts, out, &pk);
  if (st != OVB_OK)
    return st;
  const int F = pk.n_feats, n_all = pk.n_all;
  if (ld_out < n_all)
    return OVB_ERR_ARG;
  *ncols_out = n_all;
  for (int s = 0, c = 0; s < ctx->h_frame->n_slots; s++)
    for (int k = 0; k < ctx->h_frame->slot_size[s]; k++)
      col_index_out[c++] = ctx->h_frame->slot_off[s] + k;
  launch_cam_poses(ctx);
  if (stage == 0) {
    int rows = 2 * pk.n_meas;
    size_t need = (size_t)rows * (OVB_MAX_COLS + 4);
    if (!ctx->d_dump || need > ctx->dump_cap) {
      if (ctx->d_dump)
        cudaFree(ctx->d_dump);
      ctx->dump_cap = need;
      OVB_CUDA_CHECK(ctx, cudaMalloc(&ctx->d_dump, sizeof(double) * ctx->dump_cap));
    }
    ctx->dump_rows = rows;
    OVB_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_dump, 0, sizeof(double) * need, ctx->stream));
    launch_feature_system(ctx, F, pk.bv, pk.ldH, 1, pk.max_M);
    OVB_CUDA_CHECK(ctx, cudaGetLastError());
    std::vector<double> host(need);
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(host.data(), ctx->d_dump, sizeof(double) * need, cudaMemcpyDeviceToHost, ctx->stream));
    OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    int dump_rows = rows;
    const double *dHf = host.data(), *dres = host.data() + (size_t)dump_rows * 3, *dHx = host.data() + (size_t)dump_rows * 4;
    for (int f = 0; f <= F; f++)
      row_off_out[f] = 2 * feats->meas_off[f];
    for (int i = 0; i < rows; i++) {
      if (Hf_out)
        for (int k = 0; k < 3; k++)
          Hf_out[(size_t)i * 3 + k] = dHf[(size_t)i * 3 + k];
      if (res_out)
        res_out[i] = dres[i];
      if (Hx_out)
        for (int j = 0; j < n_all; j++)
          Hx_out[(size_t)i * ld_out + j] = dHx[(size_t)i * OVB_MAX_COLS + j];
    }
    return OVB_OK;
  }
  launch_feature_system(ctx, F, pk.bv, pk.ldH, 0, pk.max_M);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  std::vector<double> host((size_t)pk.m_total * pk.ldH);
  if (pk.m_total > 0)
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(host.data(), ctx->d_Hs, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_feat, ctx->d_feat, sizeof(DevFeat) * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unpack_feats(ctx, F, out);
  for (int f = 0; f < F; f++)
    row_off_out[f] = ctx->h_feat[f].row0;
  row_off_out[F] = pk.m_total;
  for (int i = 0; i < pk.m_total; i++) {
    if (res_out)
      res_out[i] = host[(size_t)i * pk.ldH + n_all];
    if (Hx_out)
      for (int j = 0; j < n_all; j++)
        Hx_out[(size_t)i * ld_out + j] = host[(size_t)i * pk.ldH + j];
  }
  return OVB_OK;
}

// UpdaterSLAM::delayed_init in one call (see include/ovb200.h). Composition of the staged entry points with the state mean
// moved by the caller's callback between the features, exactly the reference's sequential structure.
ovb_status ovb_slam_delayed_init(ovb_ctx *ctx, const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, const double *sigma_pix,
                                 const double *chi2_multipler, ovb_init_callback on_init, void *user, ovb_feat_out *out, int32_t *lm_off_out) {
  if (!ctx || !frame || !feats || !opts || !out || !out->status || !out->p_FinA || !out->p_FinG || !out->anchor_cam || !out->anchor_clone || !lm_off_out)
    return OVB_ERR_ARG;
  if (opts->feat_rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_slam_delayed_init: the 1-wide SINGLE representation uses the staged route (INTEGRATION.md)");
    return OVB_ERR_ARG;
  }
  const int F = feats->n_feats;
  for (int f = 0; f < F; f++)
    lm_off_out[f] = -1;
  if (F <= 0)
    return OVB_OK;
  // 1. triangulate + GN all tracks at the current state (UpdaterSLAM.cpp:118-142)
  ovb_status st = ovb_triangulate(ctx, frame, feats, opts, out);
  if (st != OVB_OK)
    return st;
  // 2. one feature after the other
  std::vector<double> Hf, Hx, res, HR, dx;
  std::vector<int32_t> colidx(OVB_MAX_COLS);
  for (int f = 0; f < F; f++) {
    if (out->status[f] != OVB_FEAT_OK)
      continue;
    const int m0 = feats->meas_off[f], m1 = feats->meas_off[f + 1], rows = 2 * (m1 - m0);
    // shallow one-feature view of the batch
    int32_t moff[2] = {0, m1 - m0}, koff[2] = {0, 0};
    ovb_feat_batch v = *feats;
    v.n_feats = 1;
    v.n_meas = m1 - m0;
    v.meas_off = moff;
    v.cam = feats->cam + m0;
    v.clone = feats->clone + m0;
    v.uv = feats->uv + 2 * (size_t)m0;
    v.uvn = feats->uvn + 2 * (size_t)m0;
    if (feats->cam_keys_off && feats->cam_keys) {
      koff[1] = feats->cam_keys_off[f + 1] - feats->cam_keys_off[f];
      v.cam_keys_off = koff;
      v.cam_keys = feats->cam_keys + feats->cam_keys_off[f];
    }
    int32_t st1 = OVB_FEAT_OK, ac = out->anchor_cam[f], acl = out->anchor_clone[f];
    double pA[3] = {out->p_FinA[3 * f], out->p_FinA[3 * f + 1], out->p_FinA[3 * f + 2]};
    double pG[3] = {out->p_FinG[3 * f], out->p_FinG[3 * f + 1], out->p_FinG[3 * f + 2]}, c2 = 0;
    ovb_feat_out o1{&st1, pA, pG, &ac, &acl, &c2};
    Hf.assign((size_t)rows * 3, 0.0);
    Hx.assign((size_t)rows * OVB_MAX_COLS, 0.0);
    res.assign((size_t)rows, 0.0);
    int32_t row_off[2], ncols = 0;
    st = ovb_feature_jacobians(ctx, frame, &v, opts, &o1, 0, Hf.data(), Hx.data(), res.data(), row_off, &ncols, colidx.data(), OVB_MAX_COLS);
    if (st != OVB_OK)
      return st;
    // Hx_order: the variables this feature touches = runs of consecutive covariance columns with a non-zero entry
    std::vector<int> used;
    for (int j = 0; j < ncols; j++) {
      bool nz = false;
      for (int i = 0; i < rows && !nz; i++)
        nz = Hx[(size_t)i * OVB_MAX_COLS + j] != 0.0;
      if (nz)
        used.push_back(j);
    }
    std::vector<int> off, sz;
    for (size_t a = 0; a < used.size();) {
      size_t b = a + 1;
      while (b < used.size() && colidx[(size_t)used[b]] == colidx[(size_t)used[b - 1]] + 1)
        b++;
      off.push_back(colidx[(size_t)used[a]]);
      sz.push_back((int)(b - a));
      a = b;
    }
    const int n = (int)used.size();
    HR.assign((size_t)rows * n, 0.0);
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < n; j++)
        HR[(size_t)i * n + j] = Hx[(size_t)i * OVB_MAX_COLS + used[(size_t)j]];
    const double sp = sigma_pix ? sigma_pix[f] : opts->sigma_pix, cm = chi2_multipler ? chi2_multipler[f] : opts->chi2_multipler;
    const int N0 = ctx->N;
    int accepted = 0;
    double dx_new[3] = {0, 0, 0};
    dx.assign((size_t)N0 + 3, 0.0);
    st = ovb_cov_initialize(ctx, off.data(), sz.data(), (int)off.size(), HR.data(), Hf.data(), res.data(), rows, 3, sp * sp, cm, &accepted, dx_new, dx.data());
    if (st != OVB_OK)
      return st;
    if (!accepted) {
      out->status[f] = OVB_FEAT_CHI2;
      continue;
    }
    lm_off_out[f] = N0;
    if (on_init)
      on_init(user, f, N0, 3, dx_new, dx.data(), N0 + 3); // the host moves its mean and refreshes the frame arrays
  }
  return OVB_OK;
}

// copy column n (the residual z) of the n x (n+1) R into d_w so the Cholesky kernel can append it
__global__ void k_take_z(const double *R, int ldR, int rows, int col, double *w) {
  OVB_PDL_ENTER();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows)
    w[i] = R[(size_t)i * ldR + col];
}
// col_state for the canonical layout (used by ovb_compress-less paths): info->col_state[j] = slot_off + k
__global__ void k_fill_zero_dx(double *dx, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N)
    dx[i] = 0.0;
}

// measurement_compress_inplace in the requested mode; returns the number of rows of [R | z] handed to the EKF update.
// The Cholesky-based modes always produce n rows; the Householder path leaves min(m, n).
static int compress_system(ovb_ctx *ctx, int mode, double *A, int m, int n, int ldA, double *Rout, int ldR) {
  if (mode == OVB_COMPRESS_CHOLQR2 && launch_compress_cholqr2(ctx, A, m, n, ldA, Rout, ldR) >= 0)
    return n;
  if (mode == OVB_COMPRESS_NORMAL_EQUATIONS && launch_compress_gram(ctx, A, m, n, ldA, Rout, ldR) >= 0)
    return n;
  launch_tsqr(ctx, A, m, n, ldA, Rout, ldR);
  return std::min(m, n);
}

// The device pipeline of one update on inputs already in the arena: steps 2-6 of UpdaterMSCKF::update.
// ev (optional): ev[1] after triangulation, ev[2] after the per-feature systems, ev[3] after the column map, ev[4] after
// compression, ev[5] after the EKF update. Returns the row count handed to the EKF update.
// slam: UpdaterSLAM::update — landmarks come from the state (no triangulation), rows are kept unprojected and whitened.
static int enqueue_update(ovb_ctx *ctx, int F, BlobView bv, int ldH, int max_M, int m_total, int n_all, int col_order, cudaEvent_t *ev,
                          bool slam = false) {
  const int N = ctx->N;
  ctx->n_launch = 0;
  ctx->n_launch_tsqr_level = 0;
  ctx->prof_n = 0;
  if (!slam) {
    launch_cam_poses(ctx);
    launch_triangulate(ctx, F, bv);
    ctx->n_launch += 2;
  }
  ctx->n_launch += 1; // column map (the per-feature kernel counts its own launches: one, or one per size class)
  if (ev)
    cudaEventRecord(ev[1], ctx->stream);
  launch_feature_system(ctx, F, bv, ldH, slam ? 2 : 0, max_M);
  if (ev)
    cudaEventRecord(ev[2], ctx->stream);
  // the column bookkeeping (a single serial CTA) only feeds the re-ordering and the EKF update: it runs on the side
  // stream while the compression owns the GPU, and is joined before its first consumer
  {
    cudaEventRecord(ctx->ev_fork, ctx->stream);
    cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0);
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->side_stream;
    launch_column_map(ctx, F, bv, slam ? (ctx->h_opts->o.feat_rep == OVB_REP_ANCHORED_INVERSE_DEPTH_SINGLE ? 2 : 0) : 3);
    ctx->stream = main_stream;
    cudaEventRecord(ctx->ev_join, ctx->side_stream);
  }
  if (ev)
    cudaEventRecord(ev[3], ctx->stream);
  const int ldR = ldH;
  const double *Rfinal = ctx->d_R;
  int r = 0;
  if (m_total > 0) {
    r = compress_system(ctx, ctx->h_opts->o.compress, ctx->d_Hs, m_total, n_all, ldH, ctx->d_R, ldR);
    cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    if (col_order == OVB_COLS_REFERENCE_FIRST_SEEN) {
      launch_reorder_R(ctx, ctx->d_R, n_all, ldR, ctx->d_R2, ldR);
      Rfinal = ctx->d_R2;
    }
  }
  else
    cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
  if (ev)
    cudaEventRecord(ev[4], ctx->stream);
  if (r > 0) {
    ovb_launch(ctx, k_take_z, dim3((r + 127) / 128), dim3(128), (size_t)(0), Rfinal, ldR, r, n_all, ctx->d_w);
    launch_ekf_update(ctx, Rfinal, ldR, r, n_all, false, slam ? 1.0 : ctx->h_opts->sigma_pix_sq, nullptr);
    ctx->n_launch += 7; // take_z + prep, 2 gemm, chol, trsm, downdate
  } else {
    k_fill_zero_dx<<<(N + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_dx, N);
    ctx->n_launch += 1;
  }
  if (ev)
    cudaEventRecord(ev[5], ctx->stream);
  return r;
}

// Per-kernel timing of the update pipeline (bench.py's roofline block): while on, every kernel launched through
// ovb_launch on the context stream is bracketed by CUDA events and programmatic dependent launch is disabled, so each
// duration is that kernel alone, in stream order, with whatever the previous kernels left in L2.
ovb_status ovb_set_profile(ovb_ctx *ctx, int enabled) {
  if (!ctx)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (enabled && ctx->prof_ev[0] == nullptr)
    for (int i = 0; i < 2 * 96; i++)
      OVB_CUDA_CHECK(ctx, cudaEventCreate(&ctx->prof_ev[i]));
  ctx->prof_on = enabled ? 1 : 0;
  ctx->prof_n = 0;
  return OVB_OK;
}

// kernels of the LAST update call, in launch order: names (NUL-separated, truncated to name_cap bytes in total) and their
// durations in microseconds. *n = number of entries (<= cap). Call after the update returned (the stream is idle).
ovb_status ovb_profile_read(ovb_ctx *ctx, char *names, int name_cap, float *us, int cap, int *n) {
  if (!ctx || !names || !us || !n || name_cap < 1)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int w = 0, k = 0;
  for (; k < ctx->prof_n && k < cap; k++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->prof_ev[2 * k], ctx->prof_ev[2 * k + 1]) != cudaSuccess)
      cudaGetLastError();
    us[k] = 1e3f * ms;
    const char *nm = nullptr;
    if (cudaFuncGetName(&nm, ctx->prof_fn[k]) != cudaSuccess || !nm) {
      cudaGetLastError();
      nm = "?";
    }
    for (const char *c = nm; *c && w < name_cap - 2; c++)
      names[w++] = *c;
    names[w++] = 0;
  }
  if (w < name_cap)
    names[w] = 0;
  *n = k;
  return OVB_OK;
}

ovb_status ovb_last_host_us(const ovb_ctx *ctx, double out[4]) {
  if (!ctx || !out)
    return OVB_ERR_ARG;
  for (int i = 0; i < 4; i++)
    out[i] = ctx->host_us[i];
  return OVB_OK;
}

ovb_status ovb_last_counters(const ovb_ctx *ctx, int64_t out[4]) {
  if (!ctx || !out)
    return OVB_ERR_ARG;
  out[0] = ctx->n_launch;            // kernels launched by the last update pipeline
  out[1] = ctx->n_launch_tsqr_level; // of which k_tsqr_level
  out[2] = (int64_t)ctx->last_h2d_bytes;
  out[3] = (int64_t)ctx->last_d2h_bytes;
  return OVB_OK;
}

ovb_status ovb_set_replay(ovb_ctx *ctx, int enabled) {
  if (!ctx)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (enabled && !ctx->P_snap)
    OVB_CUDA_CHECK(ctx, cudaMalloc(&ctx->P_snap, sizeof(double) * (size_t)ctx->ldP * ctx->ldP));
  ctx->replay_enabled = enabled ? 1 : 0;
  ctx->last_pk_valid = 0;
  return OVB_OK;
}

ovb_status ovb_msckf_replay(ovb_ctx *ctx, int steps, int flush_l2, float *ms_per_step, float stage_ms_sum[5]) {
  if (!ctx || steps < 1 || !ms_per_step)
    return OVB_ERR_ARG;
  if (!ctx->replay_enabled || !ctx->last_pk_valid) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_msckf_replay: call ovb_set_replay(ctx,1) and ovb_msckf_update first");
    return OVB_ERR_ARG;
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const size_t flush_bytes = (size_t)256 << 20; // > 126 MB of L2
  if (flush_l2 && !ctx->d_flush)
    OVB_CUDA_CHECK(ctx, cudaMalloc(&ctx->d_flush, flush_bytes));
  std::vector<cudaEvent_t> evs((size_t)steps * 6);
  for (auto &e : evs)
    OVB_CUDA_CHECK(ctx, cudaEventCreate(&e));
  const size_t Pbytes = sizeof(double) * (size_t)ctx->ldP * ctx->N;
  for (int s = 0; s < steps; s++) {
    if (flush_l2)
      cudaMemsetAsync(ctx->d_flush, s & 0xff, flush_bytes, ctx->stream);
    // restore the prior saved by the last ovb_msckf_update (outside the timed bracket: it is not part of an update)
    cudaMemcpyAsync(ctx->P[ctx->cur], ctx->P_snap, Pbytes, cudaMemcpyDeviceToDevice, ctx->stream);
    cudaEvent_t *ev = &evs[(size_t)s * 6];
    cudaEventRecord(ev[0], ctx->stream);
    enqueue_update(ctx, ctx->last_n_feats, ctx->last_bv, ctx->last_ldH, ctx->last_max_M, ctx->last_m_total, ctx->last_n_all, ctx->last_col_order,
                   ev);
  }
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (stage_ms_sum)
    for (int k = 0; k < 5; k++)
      stage_ms_sum[k] = 0.f;
  for (int s = 0; s < steps; s++) {
    cudaEvent_t *ev = &evs[(size_t)s * 6];
    cudaEventElapsedTime(&ms_per_step[s], ev[0], ev[5]);
    if (stage_ms_sum)
      for (int k = 0; k < 5; k++) {
        float t;
        cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
        stage_ms_sum[k] += t;
      }
  }
  for (auto &e : evs)
    cudaEventDestroy(e);
  return OVB_OK;
}

ovb_status ovb_msckf_update(ovb_ctx *ctx, const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, ovb_feat_out *out,
                            double *dx, ovb_stats *stats) {
  if (!ctx || !dx)
    return OVB_ERR_ARG;
  if (ctx->N < 1) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_msckf_update: no covariance loaded (ovb_cov_set)");
    return OVB_ERR_ARG;
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const int N = ctx->N;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->neg_diag_index = -1;
  }
  for (int i = 0; i < N; i++)
    dx[i] = 0.0;
  if (!feats || feats->n_feats <= 0) // UpdaterMSCKF.cpp:61-62
    return OVB_OK;
  cudaEventRecord(ctx->ev[0], ctx->stream);
  const auto h0 = std::chrono::steady_clock::now();
  Packed pk;
  ovb_status st = pack_inputs(ctx, frame, feats, opts, nullptr, &pk);
  if (st != OVB_OK)
    return st;
  const auto h1 = std::chrono::steady_clock::now();
  const int F = pk.n_feats;
  if (ctx->replay_enabled) {
    // keep the prior so that ovb_msckf_replay can re-run this exact update on device-resident inputs
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->P_snap, ctx->P[ctx->cur], sizeof(double) * (size_t)ctx->ldP * N, cudaMemcpyDeviceToDevice,
                                        ctx->stream));
    ctx->last_pk_valid = 1;
    ctx->last_n_feats = pk.n_feats;
    ctx->last_max_M = pk.max_M;
    ctx->last_m_total = pk.m_total;
    ctx->last_ldH = pk.ldH;
    ctx->last_n_all = pk.n_all;
    ctx->last_bv = pk.bv;
    ctx->last_col_order = opts->col_order;
  }
  const int r = enqueue_update(ctx, pk.n_feats, pk.bv, pk.ldH, pk.max_M, pk.m_total, pk.n_all, opts->col_order, ctx->ev);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_feat, ctx->d_feat, sizeof(DevFeat) * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, ctx->info_bytes + sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost,
                                      ctx->stream)); // info + dx in one copy
  ctx->last_d2h_bytes = sizeof(DevFeat) * (size_t)F + ctx->info_bytes + sizeof(double) * (size_t)N;
  cudaEventRecord(ctx->ev[6], ctx->stream);
  const auto h2 = std::chrono::steady_clock::now();
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const auto h3 = std::chrono::steady_clock::now();
  unpack_feats(ctx, F, out);
  for (int i = 0; i < N; i++)
    dx[i] = ctx->h_dx[i];
  {
    const auto h4 = std::chrono::steady_clock::now();
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::micro>(b - a).count();
    };
    ctx->host_us[0] = us(h0, h1), ctx->host_us[1] = us(h1, h2), ctx->host_us[2] = us(h2, h3), ctx->host_us[3] = us(h3, h4);
  }
  ctx->stage_pending = 1; // the five stage times are read from the events when ovb_last_stage_ms asks for them
  cudaEventElapsedTime(&ctx->stage_ms[5], ctx->ev[0], ctx->ev[6]);
  const DevUpdateInfo *inf = ctx->h_info;
  if (stats) {
    stats->n_feats_in = F;
    stats->n_feats_used = inf->n_feats_used;
    stats->rows_stacked = inf->rows_stacked;
    stats->cols_stacked = inf->n_used;
    stats->rows_update = std::min(inf->rows_stacked, inf->n_used);
    stats->neg_diag_index = (r > 0 && inf->neg_diag_index != 0x7fffffff) ? inf->neg_diag_index : -1;
    stats->ms_total = ctx->stage_ms[5];
  }
  if (r > 0) {
    if (inf->not_spd) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: innovation covariance not positive definite");
      return OVB_ERR_NOT_SPD;
    }
    if (inf->nonfinite) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: non-finite covariance entry");
      return OVB_ERR_NONFINITE;
    }
    if (inf->neg_diag_index != 0x7fffffff) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: diagonal at %d is negative", inf->neg_diag_index);
      return OVB_ERR_NEG_DIAG;
    }
  }
  return OVB_OK;
}

// UpdaterSLAM::update steps 4-5 (update/UpdaterSLAM.cpp:310-470) on the device: same per-feature kernel in its SLAM mode
// (landmark block appended, no nullspace projection, per-class noise and gate), then the shared compress + EKF stages.
ovb_status ovb_slam_update(ovb_ctx *ctx, const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_landmarks *landmarks,
                           const ovb_opts *opts, ovb_feat_out *out, double *dx, ovb_stats *stats) {
  if (!ctx || !dx || !landmarks || !opts)
    return OVB_ERR_ARG;
  if (ctx->N < 1) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_slam_update: no covariance loaded (ovb_cov_set)");
    return OVB_ERR_ARG;
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const int N = ctx->N;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->neg_diag_index = -1;
  }
  for (int i = 0; i < N; i++)
    dx[i] = 0.0;
  if (!feats || feats->n_feats <= 0) // UpdaterSLAM.cpp:256-257
    return OVB_OK;
  cudaEventRecord(ctx->ev[0], ctx->stream);
  Packed pk;
  ovb_status st = pack_inputs(ctx, frame, feats, opts, nullptr, &pk, landmarks);
  if (st != OVB_OK)
    return st;
  const int F = pk.n_feats;
  ctx->last_pk_valid = 0; // the replay path re-runs MSCKF updates only
  const int r = enqueue_update(ctx, pk.n_feats, pk.bv, pk.ldH, pk.max_M, pk.m_total, pk.n_all, opts->col_order, ctx->ev, true);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_feat, ctx->d_feat, sizeof(DevFeat) * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, ctx->info_bytes + sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost,
                                      ctx->stream)); // info + dx in one copy
  ctx->last_d2h_bytes = sizeof(DevFeat) * (size_t)F + ctx->info_bytes + sizeof(double) * (size_t)N;
  cudaEventRecord(ctx->ev[6], ctx->stream);
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (out) {
    for (int f = 0; f < F; f++) {
      if (out->status)
        out->status[f] = ctx->h_feat[f].status;
      if (out->chi2)
        out->chi2[f] = ctx->h_feat[f].chi2;
    }
  }
  for (int i = 0; i < N; i++)
    dx[i] = ctx->h_dx[i];
  ctx->stage_pending = 1; // the five stage times are read from the events when ovb_last_stage_ms asks for them
  cudaEventElapsedTime(&ctx->stage_ms[5], ctx->ev[0], ctx->ev[6]);
  const DevUpdateInfo *inf = ctx->h_info;
  if (stats) {
    stats->n_feats_in = F;
    stats->n_feats_used = inf->n_feats_used;
    stats->rows_stacked = inf->rows_stacked;
    stats->cols_stacked = inf->n_used;
    stats->rows_update = inf->rows_stacked; // what the reference hands to EKFUpdate (it never compresses here)
    stats->neg_diag_index = (r > 0 && inf->neg_diag_index != 0x7fffffff) ? inf->neg_diag_index : -1;
    stats->ms_total = ctx->stage_ms[5];
  }
  if (r > 0) {
    if (inf->not_spd) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: innovation covariance not positive definite");
      return OVB_ERR_NOT_SPD;
    }
    if (inf->nonfinite) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: non-finite covariance entry");
      return OVB_ERR_NONFINITE;
    }
    if (inf->neg_diag_index != 0x7fffffff) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: diagonal at %d is negative", inf->neg_diag_index);
      return OVB_ERR_NEG_DIAG;
    }
  }
  return OVB_OK;
}

ovb_status ovb_last_stage_ms(const ovb_ctx *ctx, float ms[6]) {
  if (!ctx || !ms)
    return OVB_ERR_ARG;
  if (ctx->stage_pending) {
    for (int s = 0; s < 5; s++)
      cudaEventElapsedTime(&ctx->stage_ms[s], ctx->ev[s], ctx->ev[s + 1]);
    ctx->stage_pending = 0;
  }
  for (int i = 0; i < 6; i++)
    ms[i] = ctx->stage_ms[i];
  return OVB_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU (features sharded)
// stages A (triangulate, Jacobian, nullspace, gate) and the local compression run on this rank's feature
// shard; the ranks exchange their compressed [R_g | z_g] blocks with ONE all-gather (done by the caller, e.g. NCCL through
// torch.distributed on the stream adopted with ovb_set_stream); every rank then compresses the stack and performs the
// identical EKF update on its replica of P (bitwise identical kernels on identical inputs keep the replicas in sync).
ovb_status ovb_set_stream(ovb_ctx *ctx, void *cuda_stream) {
  if (!ctx)
    return OVB_ERR_ARG;
  if (!cuda_stream) {
    // the legacy default stream (handle 0) has no ordering with the engine's non-blocking streams: adopting it would
    // silently leave the engine on its own stream and race with the caller's collectives
    snprintf(ctx->err, sizeof(ctx->err), "ovb_set_stream: pass an explicit (non-default) CUDA stream handle");
    return OVB_ERR_ARG;
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream)
    cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = 0;
  return OVB_OK;
}

// contiguous feature ranges with (nearly) equal stacked-row counts sum(max(2M-3, 0)); bounds[world + 1]
ovb_status ovb_shard_partition(const int32_t *meas_off, int n_feats, int world, int32_t *bounds) {
  if (!meas_off || !bounds || n_feats < 0 || world < 1)
    return OVB_ERR_ARG;
  long long total = 0;
  for (int f = 0; f < n_feats; f++)
    total += std::max(2 * (meas_off[f + 1] - meas_off[f]) - 3, 0);
  bounds[0] = 0;
  long long cum = 0;
  int f = 0;
  for (int r = 1; r < world; r++) {
    const double target = (double)total * r / world;
    while (f < n_feats && (double)cum < target) { // first f with cum(f) >= target
      cum += std::max(2 * (meas_off[f + 1] - meas_off[f]) - 3, 0);
      f++;
    }
    bounds[r] = f;
  }
  bounds[world] = n_feats;
  return OVB_OK;
}

ovb_status ovb_msckf_shard_compress_range(ovb_ctx *ctx, const ovb_frame *frame, const ovb_feat_batch *feats, int f0, int f1, const ovb_opts *opts,
                                          double *R_dev, int R_cap_doubles, int *n_cols, int *ld) {
  if (!ctx || !feats || f0 < 0 || f1 < f0 || f1 > feats->n_feats)
    return OVB_ERR_ARG;
  // shallow view of features [f0, f1): pointers advanced, offsets rebased
  const int F = f1 - f0, m0 = feats->meas_off[f0], m1 = feats->meas_off[f1];
  std::vector<int32_t> moff((size_t)F + 1), koff;
  for (int f = 0; f <= F; f++)
    moff[(size_t)f] = feats->meas_off[f0 + f] - m0;
  ovb_feat_batch v = *feats;
  v.n_feats = F;
  v.n_meas = m1 - m0;
  v.meas_off = moff.data();
  v.cam = feats->cam + m0;
  v.clone = feats->clone + m0;
  v.uv = feats->uv + 2 * (size_t)m0;
  v.uvn = feats->uvn + 2 * (size_t)m0;
  if (feats->cam_keys_off && feats->cam_keys) {
    const int k0 = feats->cam_keys_off[f0];
    koff.resize((size_t)F + 1);
    for (int f = 0; f <= F; f++)
      koff[(size_t)f] = feats->cam_keys_off[f0 + f] - k0;
    v.cam_keys_off = koff.data();
    v.cam_keys = feats->cam_keys + k0;
  }
  return ovb_msckf_shard_compress(ctx, frame, &v, opts, R_dev, R_cap_doubles, n_cols, ld); // pack_inputs copies: the view may die here
}

ovb_status ovb_msckf_shard_compress(ovb_ctx *ctx, const ovb_frame *frame, const ovb_feat_batch *feats, const ovb_opts *opts, double *R_dev,
                                    int R_cap_doubles, int *n_cols, int *ld) {
  if (!ctx || !frame || !feats || !opts || !R_dev || !n_cols || !ld || ctx->N < 1)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaEventRecord(ctx->ev[0], ctx->stream);
  Packed pk;
  ovb_opts o2 = *opts;
  o2.col_order = OVB_COLS_CANONICAL; // the column order must be fixed before sharding
  ovb_status st = pack_inputs(ctx, frame, feats, &o2, nullptr, &pk);
  if (st != OVB_OK)
    return st;
  cudaEventRecord(ctx->ev[1], ctx->stream); // inputs resident from here on
  *n_cols = pk.n_all;
  *ld = pk.ldH;
  if ((size_t)pk.n_all * pk.ldH > (size_t)R_cap_doubles)
    return OVB_ERR_CAPACITY;
  ctx->last_n_feats = pk.n_feats;
  ctx->last_n_all = pk.n_all;
  ctx->last_ldH = pk.ldH;
  ctx->n_launch = 0;
  ctx->n_launch_tsqr_level = 0;
  launch_cam_poses(ctx);
  launch_triangulate(ctx, pk.n_feats, pk.bv);
  launch_feature_system(ctx, pk.n_feats, pk.bv, pk.ldH, 0, pk.max_M);
  launch_column_map(ctx, pk.n_feats, pk.bv);
  ctx->n_launch += 3; // cam poses, triangulate, column map (+ the per-feature kernel's own count)
  if (pk.m_total > 0)
    compress_system(ctx, o2.compress, ctx->d_Hs, pk.m_total, pk.n_all, pk.ldH, R_dev, pk.ldH); // unused rows of the block read as zero
  else
    OVB_CUDA_CHECK(ctx, cudaMemsetAsync(R_dev, 0, sizeof(double) * (size_t)pk.n_all * pk.ldH, ctx->stream));
  cudaEventRecord(ctx->ev[4], ctx->stream);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  return OVB_OK;
}

ovb_status ovb_msckf_shard_finish(ovb_ctx *ctx, double *stacked_dev, int n_blocks, ovb_feat_out *out, double *dx, ovb_stats *stats) {
  if (!ctx || !stacked_dev || n_blocks < 1 || !dx || ctx->N < 1)
    return OVB_ERR_ARG;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const int N = ctx->N, n_all = ctx->last_n_all, ld = ctx->last_ldH, F = ctx->last_n_feats;
  const double *Rfinal = stacked_dev;
  if (n_blocks > 1) {
    compress_system(ctx, ctx->h_opts->o.compress, stacked_dev, n_blocks * n_all, n_all, ld, ctx->d_R, ld);
    Rfinal = ctx->d_R;
  }
  ovb_launch(ctx, k_take_z, dim3((n_all + 127) / 128), dim3(128), (size_t)(0), Rfinal, ld, n_all, n_all, ctx->d_w);
  launch_ekf_update(ctx, Rfinal, ld, n_all, n_all, false, ctx->h_opts->sigma_pix_sq, nullptr);
  ctx->n_launch += 7;
  cudaEventRecord(ctx->ev[5], ctx->stream);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_feat, ctx->d_feat, sizeof(DevFeat) * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, ctx->info_bytes + sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost,
                                      ctx->stream)); // info + dx in one copy
  ctx->last_d2h_bytes = sizeof(DevFeat) * (size_t)F + ctx->info_bytes + sizeof(double) * (size_t)N;
  cudaEventRecord(ctx->ev[6], ctx->stream);
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unpack_feats(ctx, F, out);
  for (int i = 0; i < N; i++)
    dx[i] = ctx->h_dx[i];
  ctx->stage_pending = 0;
  cudaEventElapsedTime(&ctx->stage_ms[5], ctx->ev[0], ctx->ev[6]);
  cudaEventElapsedTime(&ctx->stage_ms[3], ctx->ev[0], ctx->ev[4]);
  cudaEventElapsedTime(&ctx->stage_ms[4], ctx->ev[4], ctx->ev[5]);
  cudaEventElapsedTime(&ctx->stage_ms[0], ctx->ev[1], ctx->ev[5]); // inputs resident -> EKF update done (collective included)
  const DevUpdateInfo *inf = ctx->h_info;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->n_feats_in = F;
    int used = 0, rows = 0;
    for (int f = 0; f < F; f++)
      if (ctx->h_feat[f].status == OVB_FEAT_OK) {
        used++;
        rows += 2 * (ctx->h_feat[f].m1 - ctx->h_feat[f].m0) - 3;
      }
    stats->n_feats_used = used; // this rank's shard
    stats->rows_stacked = rows;
    stats->cols_stacked = n_all;
    stats->rows_update = n_all;
    stats->neg_diag_index = inf->neg_diag_index != 0x7fffffff ? inf->neg_diag_index : -1;
    stats->ms_total = ctx->stage_ms[5];
  }
  if (inf->not_spd)
    return OVB_ERR_NOT_SPD;
  if (inf->nonfinite)
    return OVB_ERR_NONFINITE;
  if (inf->neg_diag_index != 0x7fffffff)
    return OVB_ERR_NEG_DIAG;
  return OVB_OK;
}

// ------------------------------------------------------------------------------------------------ staged dense entry points
// rows scaled by 1/sqrt(Rdiag): whitening turns R = diag(Rdiag) into the identity so that compression applies
__global__ void k_whiten_rows(double *A, int ld, int m, int ncols) {
  int i = blockIdx.x;
  double s = 1.0 / sqrt(A[(size_t)i * ld + ncols]); // the row's noise variance rides in column ncols
  __syncthreads();
  for (int j = threadIdx.x; j < ncols; j += blockDim.x)
    A[(size_t)i * ld + j] *= s;
}

// pinned staging buffer, grown on demand (dense H uploads are a test/microbench path, not the per-frame path)
static ovb_status ensure_stage(ovb_ctx *ctx, size_t doubles) {
  if (doubles <= ctx->stage_cap && ctx->h_stage)
    return OVB_OK;
  if (ctx->h_stage)
    cudaFreeHost(ctx->h_stage);
  ctx->h_stage = nullptr;
  ctx->stage_cap = 0;
  OVB_CUDA_CHECK(ctx, cudaMallocHost(&ctx->h_stage, sizeof(double) * doubles));
  ctx->stage_cap = doubles;
  return OVB_OK;
}

// [H | res | extra] rows into the device staging matrix; column n = res, column n+1 = extra (or 0)
static ovb_status stage_dense(ovb_ctx *ctx, const double *H, int m, int n, const double *res, const double *extra, int *ld_out) {
  int ld = (int)align_up((size_t)n + 2, 4);
  if ((size_t)std::max(m, n) * ld > ctx->Hs_cap) {
    snprintf(ctx->err, sizeof(ctx->err), "dense system %d x %d exceeds the reserved staging matrix (max_rows/max_state)", m, n);
    return OVB_ERR_CAPACITY;
  }
  ovb_status st = ensure_stage(ctx, (size_t)std::max(m, n) * ld);
  if (st != OVB_OK)
    return st;
  double *hs = ctx->h_stage;
  for (int i = 0; i < m; i++) {
    memcpy(hs + (size_t)i * ld, H + (size_t)i * n, sizeof(double) * n);
    hs[(size_t)i * ld + n] = res[i];
    for (int j = n + 1; j < ld; j++)
      hs[(size_t)i * ld + j] = 0.0;
    if (extra)
      hs[(size_t)i * ld + n + 1] = extra[i];
  }
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_Hs, hs, sizeof(double) * (size_t)m * ld, cudaMemcpyHostToDevice, ctx->stream));
  *ld_out = ld;
  return OVB_OK;
}

ovb_status ovb_compress(ovb_ctx *ctx, const double *H, int m, int n, const double *res, double *R_out, double *z_out) {
  if (!ctx || !H || !res || !R_out || !z_out || m < 1 || n < 1)
    return OVB_ERR_ARG;
  if (n + 8 > ctx->cfg.max_state + 8)
    return OVB_ERR_CAPACITY;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int ld;
  ovb_status st = stage_dense(ctx, H, m, n, res, nullptr, &ld);
  if (st != OVB_OK)
    return st;
  launch_tsqr(ctx, ctx->d_Hs, m, n, ld, ctx->d_R, ld);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  double *hs = ctx->h_stage;
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(hs, ctx->d_R, sizeof(double) * (size_t)n * ld, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++)
      R_out[(size_t)i * n + j] = hs[(size_t)i * ld + j];
    z_out[i] = hs[(size_t)i * ld + n];
  }
  return OVB_OK;
}

ovb_status ovb_compress_gram(ovb_ctx *ctx, const double *H, int m, int n, const double *res, double *R_out, double *z_out) {
  if (!ctx || !H || !res || !R_out || !z_out || m < 1 || n < 1)
    return OVB_ERR_ARG;
  if (n > ctx->cfg.max_state)
    return OVB_ERR_CAPACITY;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int ld;
  ovb_status st = stage_dense(ctx, H, m, n, res, nullptr, &ld);
  if (st != OVB_OK)
    return st;
  if (launch_compress_gram(ctx, ctx->d_Hs, m, n, ld, ctx->d_R, ld) < 0)
    return OVB_ERR_CUDA;
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  double *hs = ctx->h_stage;
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(hs, ctx->d_R, sizeof(double) * (size_t)n * ld, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++)
      R_out[(size_t)i * n + j] = hs[(size_t)i * ld + j];
    z_out[i] = hs[(size_t)i * ld + n];
  }
  return OVB_OK;
}

ovb_status ovb_compress_cholqr2(ovb_ctx *ctx, const double *H, int m, int n, const double *res, double *R_out, double *z_out) {
  if (!ctx || !H || !res || !R_out || !z_out || m < 1 || n < 1)
    return OVB_ERR_ARG;
  if (n > ctx->cfg.max_state)
    return OVB_ERR_CAPACITY;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int ld;
  ovb_status st = stage_dense(ctx, H, m, n, res, nullptr, &ld);
  if (st != OVB_OK)
    return st;
  if (launch_compress_cholqr2(ctx, ctx->d_Hs, m, n, ld, ctx->d_R, ld) < 0) {
    snprintf(ctx->err, sizeof(ctx->err), "ovb_compress_cholqr2: %d columns exceed what this path takes", n);
    return OVB_ERR_CAPACITY;
  }
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  double *hs = ctx->h_stage;
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(hs, ctx->d_R, sizeof(double) * (size_t)n * ld, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++)
      R_out[(size_t)i * n + j] = hs[(size_t)i * ld + j];
    z_out[i] = hs[(size_t)i * ld + n];
  }
  return OVB_OK;
}

ovb_status ovb_ekf_update(ovb_ctx *ctx, const int *off, const int *sz, int nvar, const double *H, int r, const double *res, double sigma2,
                          const double *Rdiag, double *dx) {
  if (!ctx || !off || !sz || !H || !res || !dx || nvar < 1 || r < 1)
    return OVB_ERR_ARG;
  if (ctx->N < 1)
    return OVB_ERR_ARG;
  const int N = ctx->N;
  int n = 0;
  for (int i = 0; i < nvar; i++) {
    if (off[i] < 0 || sz[i] < 1 || off[i] + sz[i] > N)
      return OVB_ERR_ARG;
    n += sz[i];
  }
  if (n > OVB_MAX_COLS || n > ctx->cfg.max_state)
    return OVB_ERR_CAPACITY;
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (Rdiag)
    for (int i = 0; i < r; i++)
      if (!(Rdiag[i] > 0.0))
        return OVB_ERR_ARG;
  int ld;
  ovb_status st = stage_dense(ctx, H, r, n, res, Rdiag, &ld);
  if (st != OVB_OK)
    return st;
  DevUpdateInfo *hi = ctx->h_info;
  memset(hi, 0, sizeof(*hi));
  int c = 0;
  for (int i = 0; i < nvar; i++)
    for (int k = 0; k < sz[i]; k++)
      hi->col_state[c++] = off[i] + k;
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_info, hi, sizeof(DevUpdateInfo), cudaMemcpyHostToDevice, ctx->stream));
  const double *Hdev = ctx->d_Hs;
  int rr = r;
  double s2 = sigma2;
  if (Rdiag) {
    // whiten each row by 1/sqrt(R_ii) so that R = I; then compression is admissible (UpdaterSLAM.cpp:444 uses diagonal R)
    k_whiten_rows<<<r, 128, 0, ctx->stream>>>(ctx->d_Hs, ld, r, n + 1);
    s2 = 1.0;
  }
  if (r > n) {
    // more rows than columns: compress first (identical update, UpdaterMSCKF.cpp:275 does the same before EKFUpdate)
    rr = compress_system(ctx, OVB_COMPRESS_CHOLQR2, ctx->d_Hs, r, n, ld, ctx->d_R, ld);
    Hdev = ctx->d_R;
  }
  ovb_launch(ctx, k_take_z, dim3((rr + 127) / 128), dim3(128), (size_t)(0), Hdev, ld, rr, n, ctx->d_w);
  // k_take_z reads column n: for the uncompressed case that is the staged residual column
  launch_ekf_update(ctx, Hdev, ld, rr, n, false, s2, nullptr);
  OVB_CUDA_CHECK(ctx, cudaGetLastError());
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(DevUpdateInfo), cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_dx, ctx->d_dx, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
  OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < N; i++)
    dx[i] = ctx->h_dx[i];
  if (ctx->h_info->not_spd) {
    snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: innovation covariance not positive definite");
    return OVB_ERR_NOT_SPD;
  }
  if (ctx->h_info->nonfinite)
    return OVB_ERR_NONFINITE;
  if (ctx->h_info->neg_diag_index != 0x7fffffff) {
    snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: diagonal at %d is negative", ctx->h_info->neg_diag_index);
    return OVB_ERR_NEG_DIAG;
  }
  return OVB_OK;
}

// StateHelper::initialize on the device-resident covariance. The Givens split of the (tiny) system and the 3x3 inverse are
// host work in the reference's operation order; the gate, the augmentation and the update run on the GPU.
ovb_status ovb_cov_initialize(ovb_ctx *ctx, const int *off, const int *sz, int nvar, const double *H_R_in, const double *H_L_in,
                              const double *res_in, int r, int k, double sigma2, double chi2_mult, int *accepted, double *dx_new, double *dx) {
  if (!ctx || !off || !sz || !H_R_in || !H_L_in || !res_in || !accepted || !dx_new || !dx || nvar < 1 || k < 1 || k > 3 || r < k)
    return OVB_ERR_ARG;
  if (ctx->N < 1 || !(sigma2 > 0.0))
    return OVB_ERR_ARG;
  const int N = ctx->N;
  int n = 0;
  for (int i = 0; i < nvar; i++) {
    if (off[i] < 0 || sz[i] < 1 || off[i] + sz[i] > N)
      return OVB_ERR_ARG;
    n += sz[i];
  }
  if (n > OVB_MAX_COLS || n > ctx->cfg.max_state)
    return OVB_ERR_CAPACITY;
  if (N + k > ctx->ldP) {
    snprintf(ctx->err, sizeof(ctx->err), "initialize: covariance would grow to %d (max_state %d)", N + k, ctx->ldP);
    return OVB_ERR_CAPACITY;
  }
  OVB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  *accepted = 0;
  // ---- Givens split (StateHelper.cpp:429-440), row-major copies
  std::vector<double> HR(H_R_in, H_R_in + (size_t)r * n), HL(H_L_in, H_L_in + (size_t)r * k), res(res_in, res_in + r);
  auto rot = [](double c, double s, double &x, double &y) { // applyOnTheLeft(0, 1, G.adjoint()) on the pair (x, y)
    const double x0 = x, y0 = y;
    x = c * x0 - s * y0;
    y = s * x0 + c * y0;
  };
  for (int c0 = 0; c0 < k; ++c0) {
    for (int m = r - 1; m > c0; m--) {
      // JacobiRotation::makeGivens(p, q) (Eigen/src/Jacobi/Jacobi.h, real branch): G' [p; q] = [r; 0], r >= 0
      const double p = HL[(size_t)(m - 1) * k + c0], q = HL[(size_t)m * k + c0];
      double gc, gs;
      if (q == 0.0) {
        gc = p < 0.0 ? -1.0 : 1.0;
        gs = 0.0;
      } else if (p == 0.0) {
        gc = 0.0;
        gs = q < 0.0 ? 1.0 : -1.0;
      } else if (std::fabs(p) > std::fabs(q)) {
        const double t = q / p;
        double u = std::sqrt(1.0 + t * t);
        if (p < 0.0)
          u = -u;
        gc = 1.0 / u;
        gs = -t * gc;
      } else {
        const double t = p / q;
        double u = std::sqrt(1.0 + t * t);
        if (q < 0.0)
          u = -u;
        gs = -1.0 / u;
        gc = -t * gs;
      }
      for (int j = c0; j < k; j++)
        rot(gc, gs, HL[(size_t)(m - 1) * k + j], HL[(size_t)m * k + j]);
      rot(gc, gs, res[m - 1], res[m]);
      for (int j = 0; j < n; j++)
        rot(gc, gs, HR[(size_t)(m - 1) * n + j], HR[(size_t)m * n + j]);
    }
  }
  // ---- H_L^-1 of the invertible k x k block (Gauss-Jordan, partial pivoting)
  double A[9], Inv[9];
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) {
      A[i * k + j] = HL[(size_t)i * k + j];
      Inv[i * k + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c0 = 0; c0 < k; c0++) {
    int piv = c0;
    for (int i = c0 + 1; i < k; i++)
      if (std::fabs(A[i * k + c0]) > std::fabs(A[piv * k + c0]))
        piv = i;
    if (!(std::fabs(A[piv * k + c0]) > 0.0)) {
      snprintf(ctx->err, sizeof(ctx->err), "initialize: H_L is rank deficient");
      return OVB_ERR_ARG;
    }
    if (piv != c0)
      for (int j = 0; j < k; j++) {
        std::swap(A[c0 * k + j], A[piv * k + j]);
        std::swap(Inv[c0 * k + j], Inv[piv * k + j]);
      }
    const double d = A[c0 * k + c0];
    for (int j = 0; j < k; j++) {
      A[c0 * k + j] /= d;
      Inv[c0 * k + j] /= d;
    }
    for (int i = 0; i < k; i++) {
      if (i == c0)
        continue;
      const double f = A[i * k + c0];
      for (int j = 0; j < k; j++) {
        A[i * k + j] -= f * A[c0 * k + j];
        Inv[i * k + j] -= f * Inv[c0 * k + j];
      }
    }
  }
  // ---- columns of the measuring variables
  DevUpdateInfo *hi = ctx->h_info;
  memset(hi, 0, sizeof(*hi));
  {
    int c = 0;
    for (int i = 0; i < nvar; i++)
      for (int q = 0; q < sz[i]; q++)
        hi->col_state[c++] = off[i] + q;
  }
  const int rup = r - k;
  const double *Hdev = nullptr;
  int ld = 0, rr = 0;
  if (rup > 0) {
    // ---- gate on the projected part: chi2 = resup' (Hup P Hup' + s2 I)^-1 resup (StateHelper.cpp:458-470)
    ovb_status st = stage_dense(ctx, HR.data() + (size_t)k * n, rup, n, res.data() + k, nullptr, &ld);
    if (st != OVB_OK)
      return st;
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_info, hi, sizeof(DevUpdateInfo), cudaMemcpyHostToDevice, ctx->stream));
    Hdev = ctx->d_Hs;
    rr = rup;
    double res2 = 0.0;
    for (int i = k; i < r; i++)
      res2 += res[i] * res[i];
    const bool compressed = rup > n;
    if (compressed) { // orthogonal compression keeps S's relevant block; the dropped rows add |z2|^2 / s2 to chi2
      launch_tsqr(ctx, ctx->d_Hs, rup, n, ld, ctx->d_R, ld);
      Hdev = ctx->d_R;
      rr = n;
    }
    ovb_launch(ctx, k_take_z, dim3((rr + 127) / 128), dim3(128), (size_t)(0), Hdev, ld, rr, n, ctx->d_w);
    std::vector<double> zh((size_t)rr), wh((size_t)rr);
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(zh.data(), ctx->d_w, sizeof(double) * (size_t)rr, cudaMemcpyDeviceToHost, ctx->stream));
    launch_ekf_update(ctx, Hdev, ld, rr, n, true, sigma2, nullptr);
    OVB_CUDA_CHECK(ctx, cudaGetLastError());
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(wh.data(), ctx->d_w, sizeof(double) * (size_t)rr, cudaMemcpyDeviceToHost, ctx->stream));
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(DevUpdateInfo), cudaMemcpyDeviceToHost, ctx->stream));
    OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    double chi2 = 0.0, z2 = 0.0;
    for (int i = 0; i < rr; i++) {
      chi2 += wh[(size_t)i] * wh[(size_t)i];
      z2 += zh[(size_t)i] * zh[(size_t)i];
    }
    if (compressed)
      chi2 += std::max(0.0, res2 - z2) / sigma2;
    if (ctx->h_info->not_spd)
      chi2 = NAN;
    const double chi2_check = g_chi2_table[std::min(r, OVB_CHI2_TABLE_LEN - 1)];
    if (!(chi2 <= chi2_mult * chi2_check))
      return OVB_OK; // rejected: nothing was modified
  } else {
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_info, hi, sizeof(DevUpdateInfo), cudaMemcpyHostToDevice, ctx->stream));
  }
  // ---- initialize_invertible: augment P (StateHelper.cpp:484-577); Hxinit | Inv staged in the free d_Y buffer
  {
    std::vector<double> stage((size_t)k * n + (size_t)k * k);
    for (int i = 0; i < k; i++)
      for (int j = 0; j < n; j++)
        stage[(size_t)i * n + j] = HR[(size_t)i * n + j];
    for (int i = 0; i < k * k; i++)
      stage[(size_t)k * n + i] = Inv[i];
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_Y, stage.data(), sizeof(double) * stage.size(), cudaMemcpyHostToDevice, ctx->stream));
    const bool launched = launch_cov_init_augment(ctx, k, n, ctx->d_Y, ctx->d_Y + (size_t)k * n, sigma2);
    OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); // `stage` is pageable host memory owned by this scope
    if (!launched) { // nothing was written: N and the covariance are unchanged
      snprintf(ctx->err, sizeof(ctx->err), "ovb_cov_initialize: covariance augmentation kernel could not be launched (N=%d, k=%d)", N, k);
      return OVB_ERR_CAPACITY;
    }
  }
  ctx->N = N + k;
  for (int q = 0; q < k; q++) {
    double acc = 0.0;
    for (int i = 0; i < k; i++)
      acc += Inv[q * k + i] * res[i];
    dx_new[q] = acc;
  }
  *accepted = 1;
  for (int i = 0; i < N + k; i++)
    dx[i] = 0.0;
  if (rup > 0) {
    // ---- EKFUpdate with the projected part on the augmented covariance (StateHelper.cpp:476-479)
    ovb_launch(ctx, k_take_z, dim3((rr + 127) / 128), dim3(128), (size_t)(0), Hdev, ld, rr, n, ctx->d_w);
    launch_ekf_update(ctx, Hdev, ld, rr, n, false, sigma2, nullptr);
    OVB_CUDA_CHECK(ctx, cudaGetLastError());
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(DevUpdateInfo), cudaMemcpyDeviceToHost, ctx->stream));
    OVB_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_dx, ctx->d_dx, sizeof(double) * (size_t)(N + k), cudaMemcpyDeviceToHost, ctx->stream));
    OVB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < N + k; i++)
      dx[i] = ctx->h_dx[i];
    if (ctx->h_info->not_spd) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: innovation covariance not positive definite");
      return OVB_ERR_NOT_SPD;
    }
    if (ctx->h_info->nonfinite)
      return OVB_ERR_NONFINITE;
    if (ctx->h_info->neg_diag_index != 0x7fffffff) {
      snprintf(ctx->err, sizeof(ctx->err), "EKFUpdate: diagonal at %d is negative", ctx->h_info->neg_diag_index);
      return OVB_ERR_NEG_DIAG;
    }
  }
  return OVB_OK;
}

} // extern "C"
