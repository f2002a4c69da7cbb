// rbd_capi.cu — implementation of the C ABI declared in include/rbd_b200.h.
//
// The boundary a patched KinematicChainMapping binds: opaque model / workspace handles,
// int status codes, raw pointers and sizes, no exceptions, no CPU fallback.  Device-pointer (*_soa/_dev)
// entry points only enqueue kernels on the workspace stream; host-pointer (*_host) entry points wrap them in
// pipelined host<->device copies plus the instance-major <-> tiled-SoA adapters.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <cxxabi.h>
#include <stdlib.h>

#include <new>
#include <string>
#include <vector>

#include "../../../include/rbd_b200.h"
#include "rbd_kernels.h"
#include "rbd_model_build.h"

using namespace rbd;

#ifndef RBD_JT_MINB
#define RBD_JT_MINB 1
#endif
#define RBD_JT_REGS (RBD_JT_MINB >= 2 ? 128 : 255)

// --------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int set_err(int code, const std::string& msg) { g_err = msg; return code; }
int rbd_internal_set_error(int code, const char* msg) { return set_err(code, msg ? msg : ""); }  // other translation units
static int cuda_fail(cudaError_t e, const char* what) {
  // cudaGetLastError clears non-sticky launch errors so the next call starts clean
  cudaGetLastError();
  return set_err(RBD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define CK(call)                                        \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

struct rbd_model {
  DevModel dm;
  HostTables ht;
  rbd_model_info info;
};

static const long long kRowsCacheChunk = 65536;  // instances per fill pass of the row path's Jacobian cache
static const int kHostChunk = 32768;  // instances per pipelined chunk of the *_host entry points (A/B r1i, e2e M evals/s: 8192 4.54, 16384 4.66, 32768 4.72, 65536 4.74)
static const int kSlots = 2;

struct HostSlot {
  cudaStream_t st = nullptr;
  double* aos = nullptr;  // device, instance-major staging
  double* soa = nullptr;  // device, tiled-SoA staging (tile = 32)
};

struct rbd_workspace {
  const rbd_model* model = nullptr;
  int device = 0;
  long long capacity = 0;
  cudaStream_t stream = nullptr;
  // device copies of the per-output tables
  int *d_sorted_k = nullptr, *d_outk_parent = nullptr, *d_sorted_ident = nullptr;
  double *d_sorted_pl = nullptr, *d_outk_pl = nullptr;
  DevTables tb{};
  // fused eval: one planned copy of the joint table per variant (bit 1 = CRBA, bit 0 = RNEA), host + device;
  // the kernel stages the device copy into shared memory
  static const int kPlans = 19;
  DevModel* h_plan[kPlans] = {};  // index: bit 2 = oMi requested, bit 1 = CRBA, bit 0 = RNEA; [8] = the applyJT sweep,
                                  // [9] = its ring variant (full tiles of 32); [11 + i] = eval variant i with the
                                  // support-only M format (rbd_eval_soa_sparse)
  DevModel* d_plan[kPlans] = {};  // [10] = applyDJT on the ring kernel
  int djt_ring_slots = -1;    // ... of plan [10]
  int jt_ring_slots = -1;     // ring stages per warp of plan [9]; 0 = the ring kernel cannot run this model; -1 = not planned yet
  char plan_failed[kPlans] = {};  // the planner found no configuration for this variant (reset by drop_plans): do not retry per call
  int use_jt_ring = 1;
  DevModelT<float>* h_plan32[8] = {};  // FP32 mode of the fused eval
  DevModelT<float>* d_plan32[8] = {};
  LaunchCfg cfg_eval32[8];
  int use_tmem = 1;
  int num_sms = 0;
  double* qcache = nullptr;  // [nq][capacity] assembled configuration of the last apply
  // Jacobian cache of the constraint-row applyJT (the reference's data.J / data.oMi left behind by apply,
  // KinematicChainMapping.inl:365-373): world joint Jacobian (6 nv fields) and oMi (12 (njoints-1) fields) of the applied
  // instances, instance-major; allocated on the first row call, refilled lazily after every apply
  double *jcache = nullptr, *omicache = nullptr, *jc_scratch = nullptr;  // instance-major caches + tile-32 scratch
  long long apply_epoch = 0, jc_epoch = -1;
  long long jcache_cap = 0;  // instances the Jacobian cache holds
  int use_rows_cache = 1;
  long long applied_n = 0;   // instances covered by the cache (0 = apply not called yet)
  LaunchCfg cfg_eval[8], cfg_eval_sp[8], cfg_apply[2], cfg_applyJ[2], cfg_applyJT[2], cfg_rows, cfg_jt_ring, cfg_djt[2], cfg_djt_ring;
  DevModel* d_model = nullptr;  // device copy of the joint table for the mapping kernels
  int forced_eval_block = 0;
  long long launches = 0;
  // host-path staging
  HostSlot slot[kSlots];
  size_t slot_doubles = 0;  // capacity of each of aos / soa per slot
  long long chunk = 0;
  cudaEvent_t ev = nullptr;
  // grow-only buffers for the rows entry point
  void* rows_buf = nullptr;
  size_t rows_bytes = 0;
  double* aba_scratch[2] = {nullptr, nullptr};  // forward dynamics: one record per resident warp, per chunk stream of the
                                                // host path ([1]: launches on slot[1].st, which may overlap the others)
  // the fused forward-dynamics sweep (aba2): its planned joint table (host + device), state: 0 = not planned yet, 1 = in
  // use, -1 = the planner found no configuration (or the knob is off): the global-memory sweep runs
  DevModel* h_aba2 = nullptr;
  DevModel* d_aba2 = nullptr;
  int aba2_state = 0;
  int use_aba2 = 1;
  int use_arm = 1;  // fused eval of 6- / 7-joint serial chains on the unrolled arm kernel
  double* aba2_scratch[2] = {nullptr, nullptr};
  // M^-1 sweep (minv): planned joint table, state as aba2_state (-1: the model's level records do not fit an SM)
  DevModel* h_minv = nullptr;
  DevModel* d_minv = nullptr;
  int minv_state = 0;
  double* minv_scratch[2] = {nullptr, nullptr};
  // derivatives of inverse dynamics (drnea): planned joint table, state as aba2_state; cold level records per resident warp
  DevModel* h_drnea = nullptr;
  DevModel* d_drnea = nullptr;
  int drnea_state = 0;
  double* drnea_scratch[2] = {nullptr, nullptr};
  void* sel_buf = nullptr;  // getJ: selection + block pointers
  size_t sel_bytes = 0;
};

static void drop_plans(rbd_workspace* ws);

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err != cudaSuccess) { ok = false; return; }
    if (prev != dev) {
      err = cudaSetDevice(dev);
      if (err != cudaSuccess) ok = false;
    }
  }
  ~DeviceGuard() {
    if (ok && prev >= 0) {
      int cur = -1;
      if (cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
  }
};
#define GUARD(ws)                                                        \
  if (!(ws)) return set_err(RBD_ERR_INVALID_ARG, "null workspace");      \
  DeviceGuard guard_((ws)->device);                                      \
  if (!guard_.ok) return cuda_fail(guard_.err, "cudaSetDevice")

// --------------------------------------------------------------------------------------------------------
// All RBD_API functions below were declared extern "C" in include/rbd_b200.h, so they get C linkage.

RBD_API int rbd_version(void) { return RBD_B200_VERSION; }
RBD_API const char* rbd_last_error(void) { return g_err.c_str(); }
RBD_API void rbd_clear_error(void) { g_err.clear(); }

RBD_API const char* rbd_last_kernel_name(void) {
  static thread_local std::string name;
  const char* mangled = rbd::last_kernel_mangled();
  if (!mangled) return "";
  int status = 0;
  char* dem = abi::__cxa_demangle(mangled, nullptr, nullptr, &status);
  name = (status == 0 && dem) ? dem : mangled;
  free(dem);
  return name.c_str();
}

RBD_API int rbd_device_count(int* count) {
  if (!count) return set_err(RBD_ERR_INVALID_ARG, "null count");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
  return RBD_OK;
}

RBD_API int rbd_model_create(const rbd_model_desc* desc, rbd_model** out) {
  if (!desc || !out) return set_err(RBD_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  rbd_model* m = new (std::nothrow) rbd_model();
  if (!m) return set_err(RBD_ERR_ALLOC, "out of host memory");
  std::string err;
  int rc = build_dev_model(desc, &m->dm, &m->ht, &err);
  if (rc != RBD_OK) { delete m; return set_err(rc, err); }
  rbd_model_info& I = m->info;
  const DevModel& d = m->dm;
  I.njoints = d.njoints; I.nq = d.nq; I.nv = d.nv; I.nframes = desc->nframes; I.n_out = d.n_out;
  I.has_freeflyer_root = d.has_ff_root; I.nq_joints = d.nq_joints; I.nv_joints = d.nv_joints;
  I.omi_fields = 12 * (d.njoints - 1); I.jac_fields = 6 * d.nv; I.mass_fields = d.nv * (d.nv + 1) / 2;
  I.max_depth = d.max_depth;
  {
    DevModel tmp = d;
    if (plan_eval(&tmp, true, true, true, 0, true)) {
      I.eval_smem_doubles = tmp.plan.smem_doubles; I.eval_tmem_doubles = tmp.plan.tmem_doubles;
      I.eval_warps_per_sm = tmp.plan.warps * tmp.plan.blocks_per_sm;
    } else {
      I.eval_smem_doubles = I.eval_tmem_doubles = I.eval_warps_per_sm = -1;
    }
  }
  I.eval_bytes = 8LL * ((d.nq + 2 * d.nv) + I.omi_fields + I.jac_fields + I.mass_fields + d.nv);
  I.mass_support_fields = mass_support_pattern(d, nullptr, nullptr);
  *out = m;
  return RBD_OK;
}
RBD_API void rbd_model_destroy(rbd_model* model) { delete model; }
RBD_API int rbd_model_get_info(const rbd_model* model, rbd_model_info* info) {
  if (!model || !info) return set_err(RBD_ERR_INVALID_ARG, "null argument");
  *info = model->info;
  return RBD_OK;
}

RBD_API int rbd_model_drnea_pattern(const rbd_model* model, int32_t* row_ptr, int32_t* col_idx, int32_t* nnz) {
  if (!model) return set_err(RBD_ERR_INVALID_ARG, "null model");
  const int n = drnea_support_pattern(model->dm, row_ptr, col_idx);
  if (nnz) *nnz = n;
  return RBD_OK;
}
RBD_API int rbd_model_mass_pattern(const rbd_model* model, int32_t* col_ptr, int32_t* row_idx) {
  if (!model) return set_err(RBD_ERR_INVALID_ARG, "null model");
  (void)mass_support_pattern(model->dm, col_ptr, row_idx);
  return RBD_OK;
}

// ---- workspace ------------------------------------------------------------------------------------------
static void ws_free(rbd_workspace* ws) {
  if (!ws) return;
  cudaFree(ws->d_sorted_k); cudaFree(ws->d_outk_parent); cudaFree(ws->d_sorted_ident); cudaFree(ws->d_sorted_pl); cudaFree(ws->d_outk_pl);
  cudaFree(ws->sel_buf); cudaFree(ws->aba_scratch[0]); cudaFree(ws->aba_scratch[1]);
  cudaFree(ws->aba2_scratch[0]); cudaFree(ws->aba2_scratch[1]); cudaFree(ws->d_aba2); delete ws->h_aba2;
  cudaFree(ws->minv_scratch[0]); cudaFree(ws->minv_scratch[1]); cudaFree(ws->d_minv); delete ws->h_minv;
  cudaFree(ws->drnea_scratch[0]); cudaFree(ws->drnea_scratch[1]); cudaFree(ws->d_drnea); delete ws->h_drnea;
  cudaFree(ws->qcache); cudaFree(ws->rows_buf); cudaFree(ws->d_model); cudaFree(ws->jcache); cudaFree(ws->omicache); cudaFree(ws->jc_scratch);
  for (int k = 0; k < rbd_workspace::kPlans; ++k) { cudaFree(ws->d_plan[k]); delete ws->h_plan[k]; }
  for (int k = 0; k < 8; ++k) { cudaFree(ws->d_plan32[k]); delete ws->h_plan32[k]; }
  for (int s = 0; s < kSlots; ++s) {
    cudaFree(ws->slot[s].aos); cudaFree(ws->slot[s].soa);
    if (ws->slot[s].st) cudaStreamDestroy(ws->slot[s].st);
  }
  if (ws->ev) cudaEventDestroy(ws->ev);
  delete ws;
}

RBD_API int rbd_workspace_create(const rbd_model* model, int device, int64_t capacity, void* stream,
                                 rbd_workspace** out) {
  if (!model || !out) return set_err(RBD_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (capacity < 1) return set_err(RBD_ERR_INVALID_ARG, "capacity must be >= 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return e != cudaSuccess ? cuda_fail(e, "cudaGetDeviceCount")
                            : set_err(RBD_ERR_CUDA, "no CUDA device present (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(RBD_ERR_INVALID_ARG, "device index out of range");
  DeviceGuard g(device);
  if (!g.ok) return cuda_fail(g.err, "cudaSetDevice");
  rbd_workspace* ws = new (std::nothrow) rbd_workspace();
  if (!ws) return set_err(RBD_ERR_ALLOC, "out of host memory");
  ws->model = model; ws->device = device; ws->capacity = capacity; ws->stream = (cudaStream_t)stream;
  const HostTables& ht = model->ht;
  const size_t no = ht.outk_parent.size();
#define WS_CK(call)                                                  \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) { ws_free(ws); return cuda_fail(e_, #call); } \
  } while (0)
  if (no > 0) {
    WS_CK(cudaMalloc(&ws->d_sorted_k, no * sizeof(int)));
    WS_CK(cudaMalloc(&ws->d_outk_parent, no * sizeof(int)));
    WS_CK(cudaMalloc(&ws->d_sorted_ident, no * sizeof(int)));
    WS_CK(cudaMemcpy(ws->d_sorted_ident, ht.sorted_ident.data(), no * sizeof(int), cudaMemcpyHostToDevice));
    WS_CK(cudaMalloc(&ws->d_sorted_pl, no * 12 * sizeof(double)));
    WS_CK(cudaMalloc(&ws->d_outk_pl, no * 12 * sizeof(double)));
    WS_CK(cudaMemcpy(ws->d_sorted_k, ht.sorted_k.data(), no * sizeof(int), cudaMemcpyHostToDevice));
    WS_CK(cudaMemcpy(ws->d_outk_parent, ht.outk_parent.data(), no * sizeof(int), cudaMemcpyHostToDevice));
    WS_CK(cudaMemcpy(ws->d_sorted_pl, ht.sorted_pl.data(), no * 12 * sizeof(double), cudaMemcpyHostToDevice));
    WS_CK(cudaMemcpy(ws->d_outk_pl, ht.outk_pl.data(), no * 12 * sizeof(double), cudaMemcpyHostToDevice));
  }
  ws->tb.sorted_k = ws->d_sorted_k; ws->tb.sorted_pl = ws->d_sorted_pl; ws->tb.sorted_ident = ws->d_sorted_ident;
  ws->tb.outk_parent = ws->d_outk_parent; ws->tb.outk_pl = ws->d_outk_pl;
  WS_CK(cudaMalloc(&ws->d_model, sizeof(DevModel)));
  WS_CK(cudaMemcpy(ws->d_model, &model->dm, sizeof(DevModel), cudaMemcpyHostToDevice));
  // cached configuration of the last apply: tile-32 layout, whole tiles
  WS_CK(cudaMalloc(&ws->qcache, (size_t)(model->dm.nq > 0 ? model->dm.nq : 1) * (size_t)((capacity + 31) / 32 * 32) * sizeof(double)));
  WS_CK(cudaEventCreateWithFlags(&ws->ev, cudaEventDisableTiming));
  WS_CK(cudaDeviceGetAttribute(&ws->num_sms, cudaDevAttrMultiProcessorCount, device));
#undef WS_CK
  *out = ws;
  return RBD_OK;
}

RBD_API void rbd_workspace_destroy(rbd_workspace* ws) {
  if (!ws) return;
  DeviceGuard g(ws->device);
  ws_free(ws);
}
RBD_API int rbd_workspace_set_stream(rbd_workspace* ws, void* stream) {
  if (!ws) return set_err(RBD_ERR_INVALID_ARG, "null workspace");
  ws->stream = (cudaStream_t)stream;
  return RBD_OK;
}
RBD_API int rbd_workspace_synchronize(rbd_workspace* ws) {
  GUARD(ws);
  CK(cudaStreamSynchronize(ws->stream));
  return RBD_OK;
}
RBD_API int64_t rbd_workspace_launch_count(const rbd_workspace* ws) { return ws ? ws->launches : 0; }
RBD_API int rbd_workspace_set_eval_block(rbd_workspace* ws, int threads) {
  if (!ws) return set_err(RBD_ERR_INVALID_ARG, "null workspace");
  if (threads < 0 || threads > 256 || threads % 32 != 0) return set_err(RBD_ERR_INVALID_ARG, "block must be 0 or a multiple of 32 <= 256");
  GUARD(ws);
  if (ws->forced_eval_block != threads) { ws->forced_eval_block = threads; drop_plans(ws); }
  return RBD_OK;
}
RBD_API int rbd_workspace_set_eval_tmem(rbd_workspace* ws, int enable) {
  GUARD(ws);
  if ((ws->use_tmem != 0) != (enable != 0)) { ws->use_tmem = enable ? 1 : 0; drop_plans(ws); }
  return RBD_OK;
}

RBD_API int rbd_workspace_set_eval_arm(rbd_workspace* ws, int enable) {
  GUARD(ws);
  ws->use_arm = enable ? 1 : 0;
  return RBD_OK;
}
RBD_API int rbd_workspace_set_aba_fused(rbd_workspace* ws, int enable) {
  GUARD(ws);
  if ((ws->use_aba2 != 0) != (enable != 0)) { ws->use_aba2 = enable ? 1 : 0; drop_plans(ws); }
  return RBD_OK;
}

static int ensure_plan(rbd_workspace* ws, int pk);
RBD_API int rbd_workspace_set_applyJT_ring(rbd_workspace* ws, int enable, int32_t* ring_slots) {
  GUARD(ws);
  ws->use_jt_ring = enable ? 1 : 0;
  if (ring_slots) {
    int rc = ensure_plan(ws, 9);
    if (rc) return rc;
    *ring_slots = ws->jt_ring_slots;
  }
  return RBD_OK;
}

static int check_batch(const rbd_workspace* ws, int64_t n, int64_t tile) {
  if (n < 0) return set_err(RBD_ERR_INVALID_ARG, "negative batch size");
  if (n > ws->capacity) return set_err(RBD_ERR_INVALID_ARG, "batch size exceeds the workspace capacity");
  if (tile < 1) return set_err(RBD_ERR_INVALID_ARG, "tile must be >= 1");
  return RBD_OK;
}

// ---- device SoA entry points (enqueue only) -----------------------------------------------------------------
// plan (once per variant and tuning-knob setting) + upload of the planned joint table
static int ensure_plan(rbd_workspace* ws, int pk) {
  if (ws->h_plan[pk]) return RBD_OK;
  if (ws->plan_failed[pk])
    return set_err(RBD_ERR_UNSUPPORTED, "the sweep state of this model does not fit one SM's shared + tensor memory"
                                        " at the requested block size");
  if (pk == 9 && ws->jt_ring_slots == 0) return RBD_OK;
  if (pk == 10 && ws->djt_ring_slots == 0) return RBD_OK;
  DevModel* h = new (std::nothrow) DevModel(ws->model->dm);
  if (!h) return set_err(RBD_ERR_ALLOC, "out of host memory");
  bool planned;
  std::vector<int> jt_prog;
  if (pk == 9) {
    ws->jt_ring_slots = plan_jt_ring(h, ws->model->ht, ws->use_tmem != 0, &jt_prog);
    if (ws->jt_ring_slots == 0) { delete h; return RBD_OK; }  // not an error: the caller keeps the register path
    planned = true;
  } else if (pk == 10) {
    ws->djt_ring_slots = plan_djt_ring(h, ws->model->ht, ws->use_tmem != 0, &jt_prog);
    if (ws->djt_ring_slots == 0) { delete h; return RBD_OK; }
    planned = true;
  } else {
    const int ek = pk >= 11 ? pk - 11 : pk;  // eval variant bits (the storage plan does not depend on the M format)
    planned = pk == 8 ? plan_eval(h, false, true, ws->use_tmem != 0, 0, false, /*jt=*/true, RBD_JT_REGS)
                      : plan_eval(h, (ek & 2) != 0, (ek & 1) != 0, ws->use_tmem != 0,
                                  ws->forced_eval_block / 32, (ek & 4) != 0);
  }
  if (!planned) {
    delete h;
    ws->plan_failed[pk] = 1;
    return set_err(RBD_ERR_UNSUPPORTED, "the sweep state of this model does not fit one SM's shared + tensor memory"
                                        " at the requested block size");
  }
  const std::vector<unsigned long long> blob = (pk == 9 || pk == 10) ? build_jt_blob(h, jt_prog) : build_eval_blob(h, pk >= 11);
  if (!ws->d_plan[pk]) {
    // sized for the largest blob any plan of this model can produce (the ColProg length is plan-independent)
    cudaError_t e = cudaMalloc(&ws->d_plan[pk], blob.size() * 8);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaMalloc(plan)"); }
  }
  // synchronous copy, ordered before any later launch on any stream of this device
  cudaError_t e = cudaMemcpy(ws->d_plan[pk], blob.data(), blob.size() * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaMemcpy(plan)"); }
  ws->h_plan[pk] = h;
  return RBD_OK;
}
static void drop_plans(rbd_workspace* ws) {
  // kernels still in flight keep reading the device copies: wait for them before the next upload overwrites
  cudaDeviceSynchronize();
  for (int k = 0; k < rbd_workspace::kPlans; ++k) { delete ws->h_plan[k]; ws->h_plan[k] = nullptr; }
  ws->jt_ring_slots = -1; ws->djt_ring_slots = -1;
  memset(ws->plan_failed, 0, sizeof(ws->plan_failed));
  for (int k = 0; k < 8; ++k) { delete ws->h_plan32[k]; ws->h_plan32[k] = nullptr; }
  delete ws->h_aba2; ws->h_aba2 = nullptr; ws->aba2_state = 0;
  delete ws->h_minv; ws->h_minv = nullptr; ws->minv_state = 0;
  delete ws->h_drnea; ws->h_drnea = nullptr; ws->drnea_state = 0;
  cudaFree(ws->drnea_scratch[0]); cudaFree(ws->drnea_scratch[1]); ws->drnea_scratch[0] = ws->drnea_scratch[1] = nullptr;
  cudaFree(ws->minv_scratch[0]); cudaFree(ws->minv_scratch[1]); ws->minv_scratch[0] = ws->minv_scratch[1] = nullptr;
  cudaFree(ws->aba2_scratch[0]); cudaFree(ws->aba2_scratch[1]); ws->aba2_scratch[0] = ws->aba2_scratch[1] = nullptr;
}

RBD_API int rbd_workspace_get_eval_plan(rbd_workspace* ws, int with_M, int with_tau, int32_t out[8]) {
  GUARD(ws);
  if (!out) return set_err(RBD_ERR_INVALID_ARG, "null out");
  const int pk = 4 | (with_M ? 2 : 0) | (with_tau ? 1 : 0);  // all outputs requested
  int rc = ensure_plan(ws, pk);
  if (rc) return rc;
  const DevModel& h = *ws->h_plan[pk];
  out[0] = h.plan.warps; out[1] = h.plan.blocks_per_sm; out[2] = h.plan.smem_doubles; out[3] = h.plan.tmem_doubles;
  out[4] = h.plan.tmem_cols; out[5] = h.plan.y_in_tmem;
  out[6] = (int32_t)(eval_model_bytes(h) + (size_t)h.plan.warps * 32 * (h.plan.smem_doubles + 3 * h.plan.in_slots) * 8);
  out[7] = h.plan.in_slots;
  return RBD_OK;
}

// sparse_m: M receives the support-only fields (structural zeros not stored) instead of the packed upper triangle
static int eval_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, const double* q, const double* v,
                   const double* a, double* oMi, double* J, double* M, double* tau, bool sparse_m = false) {
  if (!q) return set_err(RBD_ERR_INVALID_ARG, "q is null");
  if (tau && (!v || !a)) return set_err(RBD_ERR_INVALID_ARG, "rnea needs v and a");
  if ((tau != nullptr) != (M != nullptr) && (oMi || J)) {
    // tau or M alone with oMi / J: the single-algorithm instantiations have no oMi / J stores compiled in
    // (rbd_sweeps_body.h, RBD_MF_NO_OMIJ) — the FK + Jacobian sweep writes them, then the RNEA / CRBA sweep runs alone
    int rc2 = eval_on(ws, st, n, tile, q, nullptr, nullptr, oMi, J, nullptr, nullptr);
    if (rc2 == RBD_OK) rc2 = eval_on(ws, st, n, tile, q, v, a, nullptr, nullptr, M, tau, sparse_m);
    return rc2;
  }
  const int ek = (oMi ? 4 : 0) | (M ? 2 : 0) | (tau ? 1 : 0);
  // fixed-base serial chains of 6 / 7 joints, M and tau requested: the unrolled sweep with the joint table in the constant
  // bank (rbd_kernels_arm.cu).  A chain's M has no structural zeros: the support-only format is the same array.
  if (ws->use_arm && ws->forced_eval_block == 0 && M && tau && v && a) {
    cudaError_t ce = cudaSuccess;
    EvalArgs AA{n, tile, q, v, a, oMi, J, M, tau};
    if (launch_eval_arm(ws->model->dm, AA, ws->num_sms, st, &ce)) {
      if (ce != cudaSuccess) return cuda_fail(ce, "launch_eval_arm");
      if (n > 0) ws->launches++;
      return RBD_OK;
    }
  }
  const int pk = (sparse_m && M) ? 11 + ek : ek;
  int rc = ensure_plan(ws, pk);
  if (rc == RBD_ERR_UNSUPPORTED && M && tau && ws->forced_eval_block == 0) {
    // very deep trees: the state of the fused CRBA + RNEA sweep does not fit one SM; the two halves may
    rc = eval_on(ws, st, n, tile, q, v, a, oMi, J, M, nullptr, sparse_m);
    if (rc == RBD_OK) rc = eval_on(ws, st, n, tile, q, v, a, nullptr, nullptr, nullptr, tau);
    if (rc == RBD_OK) g_err.clear();  // the "does not fit" message of the fused plan is not an error of this call
    return rc;
  }
  if (rc) return rc;
  EvalArgs A{n, tile, q, v, a, oMi, J, M, tau};
  LaunchCfg* cfg = pk >= 11 ? ws->cfg_eval_sp : ws->cfg_eval;
  CK(launch_eval_any(*ws->h_plan[pk], ws->d_plan[pk], A, cfg, ws->num_sms, st));
  if (n > 0) ws->launches++;
  return RBD_OK;
}

// FP32 mode: the plan is made on a double copy of the model (tree structure only) with 4-byte state elements, then
// the constants are converted to float for the staged blob
static int ensure_plan32(rbd_workspace* ws, int pk) {
  if (ws->h_plan32[pk]) return RBD_OK;
  DevModel* hd = new (std::nothrow) DevModel(ws->model->dm);
  DevModelT<float>* hf = new (std::nothrow) DevModelT<float>();
  if (!hd || !hf) { delete hd; delete hf; return set_err(RBD_ERR_ALLOC, "out of host memory"); }
  if (!plan_eval(hd, (pk & 2) != 0, (pk & 1) != 0, ws->use_tmem != 0, ws->forced_eval_block / 32, (pk & 4) != 0, false,
                 /*regs=*/128, /*elem_bytes=*/4)) {
    delete hd; delete hf;
    return set_err(RBD_ERR_UNSUPPORTED, "the FP32 sweep state of this model does not fit one SM");
  }
  const std::vector<unsigned long long> blob = build_eval_blob_f32(hd, hf);
  delete hd;
  if (!ws->d_plan32[pk]) {
    cudaError_t e = cudaMalloc(&ws->d_plan32[pk], blob.size() * 8);
    if (e != cudaSuccess) { delete hf; return cuda_fail(e, "cudaMalloc(plan32)"); }
  }
  cudaError_t e = cudaMemcpy(ws->d_plan32[pk], blob.data(), blob.size() * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete hf; return cuda_fail(e, "cudaMemcpy(plan32)"); }
  ws->h_plan32[pk] = hf;
  return RBD_OK;
}

RBD_API int rbd_workspace_get_eval_plan_f32(rbd_workspace* ws, int with_M, int with_tau, int32_t out[8]) {
  GUARD(ws);
  if (!out) return set_err(RBD_ERR_INVALID_ARG, "null out");
  const int pk = 4 | (with_M ? 2 : 0) | (with_tau ? 1 : 0);
  int rc = ensure_plan32(ws, pk);
  if (rc) return rc;
  const DevModelT<float>& h = *ws->h_plan32[pk];
  out[0] = h.plan.warps; out[1] = h.plan.blocks_per_sm; out[2] = h.plan.smem_doubles; out[3] = h.plan.tmem_doubles;
  out[4] = h.plan.tmem_cols; out[5] = h.plan.y_in_tmem;
  out[6] = (int32_t)(eval_model_bytes(h) + (size_t)h.plan.warps * 32 * h.plan.smem_doubles * 4); out[7] = 1;
  return RBD_OK;
}

RBD_API int rbd_eval_soa_f32(rbd_workspace* ws, int64_t n, int64_t tile, const float* q, const float* v, const float* a,
                             float* oMi, float* J, float* M, float* tau) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  if (!q) return set_err(RBD_ERR_INVALID_ARG, "q is null");
  if (tau && (!v || !a)) return set_err(RBD_ERR_INVALID_ARG, "rnea needs v and a");
  if ((tau != nullptr) != (M != nullptr) && (oMi || J)) {  // as in the FP64 path: FK + Jacobian sweep, then RNEA / CRBA alone
    rc = rbd_eval_soa_f32(ws, n, tile, q, nullptr, nullptr, oMi, J, nullptr, nullptr);
    if (rc == RBD_OK) rc = rbd_eval_soa_f32(ws, n, tile, q, v, a, nullptr, nullptr, M, tau);
    return rc;
  }
  const int pk = (oMi ? 4 : 0) | (M ? 2 : 0) | (tau ? 1 : 0);
  rc = ensure_plan32(ws, pk);
  if (rc) return rc;
  EvalArgsT<float> A{n, tile, q, v, a, oMi, J, M, tau};
  CK(launch_eval_f32(*ws->h_plan32[pk], ws->d_plan32[pk], A, ws->cfg_eval32, ws->num_sms, ws->stream));
  if (n > 0) ws->launches++;
  return RBD_OK;
}

RBD_API int rbd_eval_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                         const double* a, double* oMi, double* J, double* M, double* tau) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  return eval_on(ws, ws->stream, n, tile, q, v, a, oMi, J, M, tau);
}
RBD_API int rbd_eval_soa_sparse(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                                const double* a, double* oMi, double* J, double* Mval, double* tau) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  return eval_on(ws, ws->stream, n, tile, q, v, a, oMi, J, Mval, tau, true);
}
RBD_API int rbd_crba_soa_sparse(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, double* Mval) {
  if (!Mval) return set_err(RBD_ERR_INVALID_ARG, "Mval is null");
  return rbd_eval_soa_sparse(ws, n, tile, q, nullptr, nullptr, nullptr, nullptr, Mval, nullptr);
}
RBD_API int rbd_crba_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, double* M) {
  if (!M) return set_err(RBD_ERR_INVALID_ARG, "M is null");
  return rbd_eval_soa(ws, n, tile, q, nullptr, nullptr, nullptr, nullptr, M, nullptr);
}
RBD_API int rbd_rnea_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                         const double* a, double* tau) {
  if (!tau) return set_err(RBD_ERR_INVALID_ARG, "tau is null");
  return rbd_eval_soa(ws, n, tile, q, v, a, nullptr, nullptr, nullptr, tau);
}

// Mass / ForceField products on the RNEA sweep: out <- out + scale * rnea(q, v | 0, a | 0) with gravity on or off
static int mass_force_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, const double* q, const double* v,
                         const double* a, double* out, double scale, int gravity_on) {
  if (!q || !out) return set_err(RBD_ERR_INVALID_ARG, "null q / output");
  int rc = ensure_plan(ws, 1);
  if (rc) return rc;
  EvalArgs A{n, tile, q, v, a, nullptr, nullptr, nullptr, out};
  A.tau_scale = scale; A.tau_acc = 1; A.gravity_on = gravity_on;
  CK(launch_eval_any(*ws->h_plan[1], ws->d_plan[1], A, ws->cfg_eval, ws->num_sms, st));
  if (n > 0) ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_addMDx_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* dx, double factor,
                           double* res) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  if (!dx) return set_err(RBD_ERR_INVALID_ARG, "dx is null");
  return mass_force_on(ws, ws->stream, n, tile, q, nullptr, dx, res, factor, 0);
}
RBD_API int rbd_addForce_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v, double* f) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  return mass_force_on(ws, ws->stream, n, tile, q, v, nullptr, f, -1.0, 1);
}

// forward dynamics (articulated-body algorithm)
static int aba_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, const double* q, const double* v,
                  const double* tau, double* ddq) {
  if (!q || !v || !tau || !ddq) return set_err(RBD_ERR_INVALID_ARG, "null q / v / tau / ddq");
  const DevModel& dm = ws->model->dm;
  for (int i = 1; i < dm.njoints; ++i)
    if (dm.joint[i].type == JT_FF && !(i == 1 && dm.joint[i].parent == 0))
      return set_err(RBD_ERR_UNSUPPORTED, "aba supports a free-flyer only as the root joint (joint 1 on the universe)");
  const int sc = (ws->slot[1].st && st == ws->slot[1].st) ? 1 : 0;
  // the fused sweep (passes 1 + 2 on on-chip level records, aba2_instance) whenever its storage plan fits the SM
  if (ws->aba2_state == 0) {
    ws->aba2_state = -1;
    if (ws->use_aba2) {
      DevModel* h = new (std::nothrow) DevModel(dm);
      if (!h) return set_err(RBD_ERR_ALLOC, "out of host memory");
      if (plan_aba2(h, ws->use_tmem != 0)) {
        if (!ws->d_aba2) CK(cudaMalloc(&ws->d_aba2, sizeof(DevModel)));
        CK(cudaMemcpyAsync(ws->d_aba2, h, dev_model_bytes(*h), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));  // `h` may be replaced by drop_plans; the upload must not read it later
        ws->h_aba2 = h; ws->aba2_state = 1;
      } else {
        delete h;
      }
    }
  }
  if (ws->aba2_state == 1) {
    if (!ws->aba2_scratch[sc]) {
      int warps = 0;
      AbaArgs S{0, 32, nullptr, nullptr, nullptr, nullptr, nullptr};
      CK(launch_aba2(*ws->h_aba2, ws->d_aba2, S, ws->num_sms, &warps, st));
      const size_t bytes = (size_t)warps * (size_t)aba2_scratch_doubles(dm) * 32 * sizeof(double);
      cudaError_t e = cudaMalloc(&ws->aba2_scratch[sc], bytes > 0 ? bytes : 256);
      if (e != cudaSuccess) { cudaGetLastError(); return set_err(RBD_ERR_ALLOC, std::string("cudaMalloc(aba scratch): ") + cudaGetErrorString(e)); }
    }
    AbaArgs A{n, tile, q, v, tau, ddq, ws->aba2_scratch[sc]};
    CK(launch_aba2(*ws->h_aba2, ws->d_aba2, A, ws->num_sms, nullptr, st));
    if (n > 0) ws->launches++;
    return RBD_OK;
  }
  if (!ws->aba_scratch[sc]) {
    int warps = 0;
    AbaArgs S{0, 32, nullptr, nullptr, nullptr, nullptr, nullptr};
    CK(launch_aba(dm, ws->d_model, S, ws->num_sms, &warps, st));
    const size_t bytes = (size_t)warps * (size_t)aba_scratch_doubles(dm) * 32 * sizeof(double);
    cudaError_t e = cudaMalloc(&ws->aba_scratch[sc], bytes > 0 ? bytes : 256);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(RBD_ERR_ALLOC, std::string("cudaMalloc(aba scratch): ") + cudaGetErrorString(e)); }
  }
  AbaArgs A{n, tile, q, v, tau, ddq, ws->aba_scratch[sc]};
  CK(launch_aba(dm, ws->d_model, A, ws->num_sms, nullptr, st));
  if (n > 0) ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_aba_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v, const double* tau,
                        double* ddq) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  return aba_on(ws, ws->stream, n, tile, q, v, tau, ddq);
}

// inverse of the joint-space mass matrix (minv_instance), packed upper triangle
static int minv_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, const double* q, double* Minv) {
  if (!q || !Minv) return set_err(RBD_ERR_INVALID_ARG, "null q / Minv");
  const DevModel& dm = ws->model->dm;
  for (int i = 1; i < dm.njoints; ++i)
    if (dm.joint[i].type == JT_FF && !(i == 1 && dm.joint[i].parent == 0))
      return set_err(RBD_ERR_UNSUPPORTED, "minv supports a free-flyer only as the root joint (joint 1 on the universe)");
  const int sc = (ws->slot[1].st && st == ws->slot[1].st) ? 1 : 0;
  if (ws->minv_state == 0) {
    ws->minv_state = -1;
    DevModel* h = new (std::nothrow) DevModel(dm);
    if (!h) return set_err(RBD_ERR_ALLOC, "out of host memory");
    if (plan_minv(h, ws->use_tmem != 0)) {
      if (!ws->d_minv) CK(cudaMalloc(&ws->d_minv, sizeof(DevModel)));
      CK(cudaMemcpyAsync(ws->d_minv, h, dev_model_bytes(*h), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      ws->h_minv = h; ws->minv_state = 1;
    } else {
      delete h;
    }
  }
  if (ws->minv_state != 1)
    return set_err(RBD_ERR_UNSUPPORTED, "minv: the level records of this tree do not fit the shared + tensor memory of an SM");
  if (!ws->minv_scratch[sc]) {
    int warps = 0;
    MinvArgs S{0, 32, nullptr, nullptr, nullptr};
    CK(launch_minv(*ws->h_minv, ws->d_minv, S, ws->num_sms, &warps, st));
    const size_t bytes = (size_t)warps * (size_t)(RBD_MINV_REC * (dm.njoints - 1) + RBD_MINV_ROOT_REC) * 32 * sizeof(double);
    cudaError_t e = cudaMalloc(&ws->minv_scratch[sc], bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(RBD_ERR_ALLOC, std::string("cudaMalloc(minv scratch): ") + cudaGetErrorString(e)); }
  }
  MinvArgs A{n, tile, q, Minv, ws->minv_scratch[sc]};
  CK(launch_minv(*ws->h_minv, ws->d_minv, A, ws->num_sms, nullptr, st));
  if (n > 0) ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_minv_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, double* Minv) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  return minv_on(ws, ws->stream, n, tile, q, Minv);
}

// derivatives of inverse dynamics (drnea_instance): dtau_dq, dtau_dv nv x nv row-major fields, M optional (packed upper)
static int drnea_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, const double* q, const double* v,
                    const double* a, double* dq, double* dv, double* M, bool sparse = false) {
  if (!q || !v || !a || !dq || !dv) return set_err(RBD_ERR_INVALID_ARG, "null q / v / a / dtau_dq / dtau_dv");
  const DevModel& dm = ws->model->dm;
  for (int i = 1; i < dm.njoints; ++i)
    if (dm.joint[i].type == JT_FF && !(i == 1 && dm.joint[i].parent == 0))
      return set_err(RBD_ERR_UNSUPPORTED, "rnea_derivatives supports a free-flyer only as the root joint (joint 1 on the universe)");
  if (ws->drnea_state == 0) {
    ws->drnea_state = -1;
    DevModel* h = new (std::nothrow) DevModel(dm);
    if (!h) return set_err(RBD_ERR_ALLOC, "out of host memory");
    if (plan_drnea(h, ws->use_tmem != 0)) {
      if (!ws->d_drnea) CK(cudaMalloc(&ws->d_drnea, sizeof(DevModel)));
      CK(cudaMemcpyAsync(ws->d_drnea, h, dev_model_bytes(*h), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      ws->h_drnea = h; ws->drnea_state = 1;
    } else {
      delete h;
    }
  }
  if (ws->drnea_state != 1)
    return set_err(RBD_ERR_UNSUPPORTED, "rnea_derivatives: the level records of this tree do not fit the shared + tensor memory of an SM");
  const int sc = (ws->slot[1].st && st == ws->slot[1].st) ? 1 : 0;
  if (!ws->drnea_scratch[sc]) {
    int warps = 0;
    DrneaArgs S{0, 32, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    CK(launch_drnea(*ws->h_drnea, ws->d_drnea, S, ws->num_sms, &warps, st));
    const size_t bytes = (size_t)warps * (size_t)(ws->h_drnea->slot_total[0] > 0 ? ws->h_drnea->slot_total[0] : 1) * 32 * sizeof(double);
    cudaError_t e = cudaMalloc(&ws->drnea_scratch[sc], bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(RBD_ERR_ALLOC, std::string("cudaMalloc(rnea_derivatives scratch): ") + cudaGetErrorString(e)); }
  }
  DrneaArgs A{n, tile, q, v, a, dq, dv, M, ws->drnea_scratch[sc], sparse ? 1 : 0};
  CK(launch_drnea(*ws->h_drnea, ws->d_drnea, A, ws->num_sms, nullptr, st));
  if (n > 0) ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_rnea_derivatives_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                                     const double* a, double* dtau_dq, double* dtau_dv, double* M) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  return drnea_on(ws, ws->stream, n, tile, q, v, a, dtau_dq, dtau_dv, M);
}

RBD_API int rbd_rnea_derivatives_soa_sparse(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                                            const double* a, double* dq_val, double* dv_val) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  return drnea_on(ws, ws->stream, n, tile, q, v, a, dq_val, dv_val, nullptr, true);
}

static int apply_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, long long inst0, const double* qj,
                    const double* root, double* out) {
  const DevModel& dm = ws->model->dm;
  if (!qj || !out) return set_err(RBD_ERR_INVALID_ARG, "null q_joints / out");
  if (root && !dm.has_ff_root) return set_err(RBD_ERR_INVALID_ARG, "root pose given but the model has no free-flyer root");
  MapArgs A{n, tile, qj, root, nullptr, out, ws->qcache, inst0};
  CK(launch_apply(dm, ws->d_model, ws->tb, A, smem_apply(dm), ws->cfg_apply, st));
  if (n > 0) ws->launches++;
  ws->apply_epoch++;  // the cached configuration changed: the Jacobian cache of the row path is stale
  return RBD_OK;
}
RBD_API int rbd_apply_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q_joints, const double* root_pose,
                          double* out) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  rc = apply_on(ws, ws->stream, n, tile, 0, q_joints, root_pose, out);
  if (rc == RBD_OK) ws->applied_n = n;
  return rc;
}

static int applyJ_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, long long inst0, const double* dq,
                     const double* rootv, double* out) {
  const DevModel& dm = ws->model->dm;
  if (!dq || !out) return set_err(RBD_ERR_INVALID_ARG, "null dq_joints / out");
  if (rootv && !dm.has_ff_root) return set_err(RBD_ERR_INVALID_ARG, "root velocity given but the model has no free-flyer root");
  MapArgs A{n, tile, dq, rootv, nullptr, out, ws->qcache, inst0};
  CK(launch_applyJ(dm, ws->d_model, ws->tb, A, smem_applyJ(dm), ws->cfg_applyJ, st));
  if (n > 0) ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_applyJ_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* dq_joints, const double* root_vel,
                           double* out) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  if (n > ws->applied_n) return set_err(RBD_ERR_NO_APPLY, "applyJ before apply: no cached configuration for these instances");
  return applyJ_on(ws, ws->stream, n, tile, 0, dq_joints, root_vel, out);
}

static int applyJT_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, long long inst0, const double* in,
                      double* out_tau, double* out_root) {
  const DevModel& dm = ws->model->dm;
  if (!in || !out_tau) return set_err(RBD_ERR_INVALID_ARG, "null in / out_tau");
  if (out_root && !dm.has_ff_root) return set_err(RBD_ERR_INVALID_ARG, "root output given but the model has no free-flyer root");
  MapArgs A{n, tile, out_tau, out_root, in, nullptr, ws->qcache, inst0};
  int rc;
  // Full tiles of a tile-32 batch go through the ring kernel (bulk asynchronous copies of the wrench stream into
  // shared memory); a partial last tile, other tilings and models whose state leaves no room for the ring take the
  // register path.
  if (tile == 32 && n >= 32 && ws->use_jt_ring && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    if ((rc = ensure_plan(ws, 9))) return rc;
    if (ws->jt_ring_slots > 0) {
      const long long nfull = n / 32 * 32;
      MapArgs F = A;
      F.n = nfull;
      CK(launch_applyJT_ring(*ws->h_plan[9], ws->d_plan[9], F, ws->jt_ring_slots, &ws->cfg_jt_ring, ws->num_sms, st, ws->use_arm != 0));
      ws->launches++;
      if (nfull == n) return RBD_OK;
      const DevModel& m = ws->model->dm;
      A.n = n - nfull;
      A.in = in + nfull * 6 * (long long)m.n_out;
      A.joints = out_tau + nfull * (long long)(m.nv - (out_root ? 6 : 0));  // field count as the kernels see it
      A.root = out_root ? out_root + nfull * 6 : nullptr;
      A.cache_inst0 = inst0 + nfull;
    }
  }
  if ((rc = ensure_plan(ws, 8))) return rc;
  CK(launch_applyJT_plan(*ws->h_plan[8], ws->d_plan[8], ws->tb, A, ws->cfg_applyJT, ws->num_sms, st));
  if (A.n > 0) ws->launches++;
  return RBD_OK;
}
// applyDJT: geometric stiffness of the mapping (optional extension: the reference's applyDJT is a no-op)
static int applyDJT_on(rbd_workspace* ws, cudaStream_t st, int64_t n, int64_t tile, long long inst0, const double* dq,
                       const double* rootv, const double* in, double kfac, double* out_tau, double* out_root) {
  const DevModel& dm = ws->model->dm;
  if (!dq || !in || !out_tau) return set_err(RBD_ERR_INVALID_ARG, "null dq_joints / in / out_tau");
  if ((rootv || out_root) && !dm.has_ff_root) return set_err(RBD_ERR_INVALID_ARG, "root arrays given but the model has no free-flyer root");
  if (dm.has_ff_root && (!rootv) != (!out_root)) return set_err(RBD_ERR_INVALID_ARG, "root velocity and root output go together");
  for (int i = 2; i < dm.njoints; ++i)
    if (dm.joint[i].type == JT_FF) return set_err(RBD_ERR_UNSUPPORTED, "applyDJT supports a free-flyer only as joint 1");
  if (dm.njoints > 1 && dm.joint[1].type == JT_FF && dm.joint[1].parent != 0)
    return set_err(RBD_ERR_UNSUPPORTED, "applyDJT supports a free-flyer only as the root joint");
  DjtArgs A{n, tile, dq, rootv, in, out_tau, out_root, kfac, ws->qcache, inst0};
  // full tiles of a tile-32 batch: the ring kernel (wrench stream staged by the copy engine, forward projection); a
  // partial last tile, other tilings and trees too deep for the tensor-memory plan: the shared-memory kernel
  if (tile == 32 && n >= 32 && ws->use_jt_ring && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    int rc = ensure_plan(ws, 10);
    if (rc) return rc;
    if (ws->djt_ring_slots > 0) {
      const long long nfull = n / 32 * 32;
      MapArgs F{nfull, tile, out_tau, out_root, in, nullptr, ws->qcache, inst0};
      F.dq_joints = dq; F.dq_root = rootv; F.kfac = kfac;
      CK(launch_applyJT_ring(*ws->h_plan[10], ws->d_plan[10], F, ws->djt_ring_slots, &ws->cfg_djt_ring, ws->num_sms, st));
      ws->launches++;
      if (nfull == n) return RBD_OK;
      const long long kj = dm.nv - (out_root ? 6 : 0);
      A.n = n - nfull;
      A.in = in + nfull * 6 * (long long)dm.n_out;
      A.dq_joints = dq + nfull * kj; A.dq_root = rootv ? rootv + nfull * 6 : nullptr;
      A.out_joints = out_tau + nfull * kj; A.out_root = out_root ? out_root + nfull * 6 : nullptr;
      A.cache_inst0 = inst0 + nfull;
    }
  }
  CK(launch_applyDJT(dm, ws->d_model, ws->tb, A, ws->cfg_djt, st));
  if (A.n > 0) ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_applyDJT_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* dq_joints, const double* root_vel,
                             const double* in, double kfactor, double* out_tau, double* out_root) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  if (n > ws->applied_n) return set_err(RBD_ERR_NO_APPLY, "applyDJT before apply: no cached configuration for these instances");
  return applyDJT_on(ws, ws->stream, n, tile, 0, dq_joints, root_vel, in, kfactor, out_tau, out_root);
}

RBD_API int rbd_applyJT_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* in, double* out_tau,
                            double* out_root) {
  GUARD(ws);
  int rc = check_batch(ws, n, tile);
  if (rc) return rc;
  if (n > ws->applied_n) return set_err(RBD_ERR_NO_APPLY, "applyJT before apply: no cached configuration for these instances");
  return applyJT_on(ws, ws->stream, n, tile, 0, in, out_tau, out_root);
}

static int ensure_jcache(rbd_workspace* ws, cudaStream_t st);
// constraint rows: refresh the Jacobian cache if an apply happened since it was filled, then one light kernel
static int rows_on(rbd_workspace* ws, cudaStream_t st, const RowArgs& A) {
  const DevModel& dm = ws->model->dm;
  if (!ws->use_rows_cache || dm.njoints < 2) {
    CK(launch_applyJT_rows(dm, ws->d_model, ws->tb, A, smem_applyJT(dm), &ws->cfg_rows, st));
    ws->launches++;
    return RBD_OK;
  }
  int rc0 = ensure_jcache(ws, st);
  if (rc0 == RBD_ERR_ALLOC) {
    // no room for the cache (a few constraint rows on a very large batch): the per-row sweep needs none
    g_err.clear();
    CK(launch_applyJT_rows(dm, ws->d_model, ws->tb, A, smem_applyJT(dm), &ws->cfg_rows, st));
    ws->launches++;
    return RBD_OK;
  }
  if (rc0) return rc0;
  CK(launch_applyJT_rows_cached(dm, ws->d_model, ws->tb, A, ws->jcache, ws->omicache, ws->num_sms, st));
  ws->launches++;
  return RBD_OK;
}
// Jacobian cache of the applied instances (see rbd_workspace::jcache): sized for the applied instances (grow-only, whole
// tiles), allocated on first use, refilled after an apply.  RBD_ERR_ALLOC when the device has no room for it.
static int ensure_jcache(rbd_workspace* ws, cudaStream_t st) {
  const DevModel& dm = ws->model->dm;
  const int kj = 6 * dm.nv, ko = 12 * (dm.njoints - 1);
  const long long need = (ws->applied_n + 31) / 32 * 32;
  const long long chunk = std::min<long long>(need, kRowsCacheChunk);
  if (!ws->jcache || ws->jcache_cap < need) {
    // all three or none: a failed allocation must not leave a half-built cache behind
    if (ws->jcache) {
      CK(cudaStreamSynchronize(st));
      cudaFree(ws->jcache); cudaFree(ws->omicache); cudaFree(ws->jc_scratch);
      ws->jcache = ws->omicache = ws->jc_scratch = nullptr; ws->jcache_cap = 0;
    }
    double *j = nullptr, *o = nullptr, *sc = nullptr;
    cudaError_t e = cudaMalloc(&j, (size_t)need * (size_t)kj * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&o, (size_t)need * (size_t)ko * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sc, (size_t)chunk * (size_t)(kj + ko) * sizeof(double));
    if (e != cudaSuccess) {
      cudaFree(j); cudaFree(o); cudaFree(sc);
      cudaGetLastError();
      return set_err(RBD_ERR_ALLOC, std::string("cudaMalloc(Jacobian cache): ") + cudaGetErrorString(e));
    }
    ws->jcache = j; ws->omicache = o; ws->jc_scratch = sc; ws->jcache_cap = need;
    ws->jc_epoch = -1;
  }
  if (ws->jc_epoch != ws->apply_epoch) {
    // FK + world joint Jacobian of the applied instances from the cached configuration (eval_kernel<false,false,32>,
    // tile-32 fast path), transposed chunk by chunk into the instance-major cache the row kernel reads
    double* sj = ws->jc_scratch;
    double* so = ws->jc_scratch + (size_t)chunk * kj;
    for (long long i0 = 0; i0 < ws->applied_n; i0 += chunk) {
      const long long cn = std::min<long long>(chunk, ws->applied_n - i0);
      int rc = eval_on(ws, st, cn, 32, ws->qcache + i0 * dm.nq, nullptr, nullptr, so, sj, nullptr, nullptr);
      if (rc) return rc;
      CK(launch_soa_to_aos(cn, kj, 32, sj, ws->jcache + i0 * kj, st));
      CK(launch_soa_to_aos(cn, ko, 32, so, ws->omicache + i0 * ko, st));
      ws->launches += 2;
    }
    ws->jc_epoch = ws->apply_epoch;
  }
  return RBD_OK;
}
RBD_API int rbd_getFrameJacobian_dev(rbd_workspace* ws, int64_t npairs, const int32_t* instance, const int32_t* out_index,
                                     double* J) {
  GUARD(ws);
  if (npairs < 0) return set_err(RBD_ERR_INVALID_ARG, "negative pair count");
  if (npairs == 0) return RBD_OK;
  if (!instance || !out_index || !J) return set_err(RBD_ERR_INVALID_ARG, "null pair arrays");
  if ((reinterpret_cast<uintptr_t>(J) & 15) != 0) return set_err(RBD_ERR_INVALID_ARG, "J must be 16-byte aligned");
  if (ws->applied_n <= 0) return set_err(RBD_ERR_NO_APPLY, "getFrameJacobian before apply: no cached configuration");
  const DevModel& dm = ws->model->dm;
  if (dm.njoints < 2) { CK(cudaMemsetAsync(J, 0, (size_t)npairs * 6 * dm.nv * sizeof(double), ws->stream)); return RBD_OK; }
  int rc = ensure_jcache(ws, ws->stream);
  if (rc) return rc;
  CK(launch_frame_jacobian(dm, ws->d_model, ws->tb, npairs, instance, out_index, J, ws->jcache, ws->omicache, ws->stream));
  ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_getFrameJacobian_host(rbd_workspace* ws, int64_t npairs, const int32_t* instance, const int32_t* out_index,
                                      double* J) {
  GUARD(ws);
  if (npairs < 0) return set_err(RBD_ERR_INVALID_ARG, "negative pair count");
  if (npairs == 0) return RBD_OK;
  if (!instance || !out_index || !J) return set_err(RBD_ERR_INVALID_ARG, "null pair arrays");
  if (ws->applied_n <= 0) return set_err(RBD_ERR_NO_APPLY, "getFrameJacobian before apply: no cached configuration");
  const int nv = ws->model->info.nv, n_out = ws->model->info.n_out;
  for (int64_t p = 0; p < npairs; ++p) {
    if (instance[p] < 0 || instance[p] >= ws->applied_n) return set_err(RBD_ERR_INVALID_ARG, "instance out of the applied range");
    if (out_index[p] < 0 || out_index[p] >= n_out) return set_err(RBD_ERR_INVALID_ARG, "output index out of range");
  }
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_i = al((size_t)npairs * 4), b_o = al((size_t)npairs * 4), b_j = al((size_t)npairs * 6 * nv * 8);
  if (ws->rows_bytes < b_i + b_o + b_j) {
    CK(cudaStreamSynchronize(ws->stream));
    cudaFree(ws->rows_buf); ws->rows_buf = nullptr; ws->rows_bytes = 0;
    CK(cudaMalloc(&ws->rows_buf, b_i + b_o + b_j));
    ws->rows_bytes = b_i + b_o + b_j;
  }
  char* base = (char*)ws->rows_buf;
  int* d_i = (int*)base; int* d_o = (int*)(base + b_i); double* d_j = (double*)(base + b_i + b_o);
  cudaStream_t st = ws->stream;
  CK(cudaMemcpyAsync(d_i, instance, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_o, out_index, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
  ws->stream = st;
  int rc = rbd_getFrameJacobian_dev(ws, npairs, d_i, d_o, d_j);
  if (rc) return rc;
  CK(cudaMemcpyAsync(J, d_j, (size_t)npairs * 6 * nv * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RBD_OK;
}
RBD_API int rbd_model_getJ_pattern(const rbd_model* model, int32_t n_sel, const int32_t* sel_out, int32_t* blk_ptr,
                                   int32_t* cols) {
  if (!model || !blk_ptr) return set_err(RBD_ERR_INVALID_ARG, "null model / blk_ptr");
  if (n_sel < 0 || (!sel_out && n_sel > model->info.n_out)) return set_err(RBD_ERR_INVALID_ARG, "bad selection size");
  if (!getJ_pattern(model->dm, model->ht, n_sel, sel_out, blk_ptr, cols)) return set_err(RBD_ERR_INVALID_ARG, "selected output index out of range");
  return RBD_OK;
}
// selection + block pointers to the device (grow-only buffer of the workspace), then one kernel over the cache
static int getJ_on(rbd_workspace* ws, int64_t n, int32_t n_sel, const int32_t* sel_out, double* d_val, int64_t* fields) {
  if (n < 0 || n_sel < 0) return set_err(RBD_ERR_INVALID_ARG, "negative size");
  if (n > ws->applied_n) return set_err(RBD_ERR_NO_APPLY, "getJ before apply: no cached configuration for these instances");
  const DevModel& dm = ws->model->dm;
  std::vector<int> blk((size_t)n_sel + 1), sel((size_t)n_sel);
  if (!sel_out && n_sel > dm.n_out) return set_err(RBD_ERR_INVALID_ARG, "bad selection size");
  if (!getJ_pattern(dm, ws->model->ht, n_sel, sel_out, blk.data(), nullptr)) return set_err(RBD_ERR_INVALID_ARG, "selected output index out of range");
  for (int s = 0; s < n_sel; ++s) sel[s] = sel_out ? sel_out[s] : s;
  if (fields) *fields = 6LL * blk[n_sel];
  if (n == 0 || n_sel == 0 || blk[n_sel] == 0 || !d_val) return RBD_OK;
  const size_t need = ((size_t)2 * n_sel + 1) * sizeof(int);
  if (ws->sel_bytes < need) {
    CK(cudaStreamSynchronize(ws->stream));
    cudaFree(ws->sel_buf); ws->sel_buf = nullptr; ws->sel_bytes = 0;
    CK(cudaMalloc(&ws->sel_buf, need));
    ws->sel_bytes = need;
  }
  int* d_sel = (int*)ws->sel_buf; int* d_blk = d_sel + n_sel;
  CK(cudaMemcpyAsync(d_sel, sel.data(), (size_t)n_sel * sizeof(int), cudaMemcpyHostToDevice, ws->stream));
  CK(cudaMemcpyAsync(d_blk, blk.data(), ((size_t)n_sel + 1) * sizeof(int), cudaMemcpyHostToDevice, ws->stream));
  CK(cudaStreamSynchronize(ws->stream));  // sel / blk are stack vectors: the copies must have read them
  int rc = ensure_jcache(ws, ws->stream);
  if (rc) return rc;
  CK(launch_getJ(dm, ws->d_model, ws->tb, n, n_sel, d_sel, d_blk, d_val, ws->jcache, ws->omicache, ws->stream));
  ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_getJ_dev(rbd_workspace* ws, int64_t n, int32_t n_sel, const int32_t* sel_out, double* val) {
  GUARD(ws);
  if (!val) return set_err(RBD_ERR_INVALID_ARG, "val is null");
  if ((reinterpret_cast<uintptr_t>(val) & 15) != 0) return set_err(RBD_ERR_INVALID_ARG, "val must be 16-byte aligned");
  return getJ_on(ws, n, n_sel, sel_out, val, nullptr);
}
RBD_API int rbd_getJ_host(rbd_workspace* ws, int64_t n, int32_t n_sel, const int32_t* sel_out, double* val) {
  GUARD(ws);
  if (!val) return set_err(RBD_ERR_INVALID_ARG, "val is null");
  int64_t fields = 0;
  int rc = getJ_on(ws, n, n_sel, sel_out, nullptr, &fields);   // validates and sizes
  if (rc || n == 0 || fields == 0) return rc;
  const size_t bytes = (size_t)n * (size_t)fields * sizeof(double);
  if (ws->rows_bytes < bytes) {
    CK(cudaStreamSynchronize(ws->stream));
    cudaFree(ws->rows_buf); ws->rows_buf = nullptr; ws->rows_bytes = 0;
    CK(cudaMalloc(&ws->rows_buf, bytes));
    ws->rows_bytes = bytes;
  }
  rc = getJ_on(ws, n, n_sel, sel_out, (double*)ws->rows_buf, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(val, ws->rows_buf, bytes, cudaMemcpyDeviceToHost, ws->stream));
  CK(cudaStreamSynchronize(ws->stream));
  return RBD_OK;
}

RBD_API int rbd_workspace_set_rows_cache(rbd_workspace* ws, int enable) {
  if (!ws) return set_err(RBD_ERR_INVALID_ARG, "null workspace");
  ws->use_rows_cache = enable ? 1 : 0;
  return RBD_OK;
}

RBD_API int rbd_applyJT_rows_dev(rbd_workspace* ws, int64_t nrows, const int32_t* row_ptr, const int32_t* row_instance,
                                 const int32_t* col, const double* val, double* out_rows, int32_t* keep) {
  GUARD(ws);
  if (nrows < 0) return set_err(RBD_ERR_INVALID_ARG, "negative row count");
  if (nrows == 0) return RBD_OK;
  if (!row_ptr || !row_instance || !out_rows || !keep) return set_err(RBD_ERR_INVALID_ARG, "null row arrays");
  if (ws->applied_n <= 0) return set_err(RBD_ERR_NO_APPLY, "applyJT(rows) before apply: no cached configuration");
  const DevModel& dm = ws->model->dm;
  RowArgs A{nrows, row_ptr, row_instance, col, val, out_rows, keep, ws->qcache};
  (void)dm;
  return rows_on(ws, ws->stream, A);
}

RBD_API int rbd_aos_to_soa_dev(rbd_workspace* ws, int64_t n, int32_t K, int64_t tile, const double* aos, double* soa) {
  GUARD(ws);
  if (n < 0 || K < 0 || tile < 1 || (n > 0 && K > 0 && (!aos || !soa))) return set_err(RBD_ERR_INVALID_ARG, "bad transpose arguments");
  CK(launch_aos_to_soa(n, K, tile, aos, soa, ws->stream));
  if (n > 0 && K > 0) ws->launches++;
  return RBD_OK;
}
RBD_API int rbd_soa_to_aos_dev(rbd_workspace* ws, int64_t n, int32_t K, int64_t tile, const double* soa, double* aos) {
  GUARD(ws);
  if (n < 0 || K < 0 || tile < 1 || (n > 0 && K > 0 && (!aos || !soa))) return set_err(RBD_ERR_INVALID_ARG, "bad transpose arguments");
  CK(launch_soa_to_aos(n, K, tile, soa, aos, ws->stream));
  if (n > 0 && K > 0) ws->launches++;
  return RBD_OK;
}

// ---- host (SOFA AoS) entry points ----------------------------------------------------------------------------
// Each chunk of <= ws->chunk instances goes through one of two slots, each with its own stream:
//   H2D (instance-major) -> aos_to_soa -> kernel -> soa_to_aos -> D2H.
// Consecutive chunks alternate slots so that the copies of one overlap the kernels of the other.
static int ensure_slots(rbd_workspace* ws, size_t fields_total) {
  const long long chunk = ws->capacity < kHostChunk ? ws->capacity : kHostChunk;
  const size_t need = fields_total * (size_t)((chunk + 31) / 32 * 32);  // tile-32 staging holds whole tiles
  if (ws->slot[0].st && ws->slot_doubles >= need && ws->chunk == chunk) return RBD_OK;
  for (int s = 0; s < kSlots; ++s) {
    if (!ws->slot[s].st) CK(cudaStreamCreateWithFlags(&ws->slot[s].st, cudaStreamNonBlocking));
    if (ws->slot_doubles < need) {
      CK(cudaStreamSynchronize(ws->slot[s].st));
      cudaFree(ws->slot[s].aos); cudaFree(ws->slot[s].soa);
      ws->slot[s].aos = ws->slot[s].soa = nullptr;
      CK(cudaMalloc(&ws->slot[s].aos, need * sizeof(double)));
      CK(cudaMalloc(&ws->slot[s].soa, need * sizeof(double)));
    }
  }
  if (ws->slot_doubles < need) ws->slot_doubles = need;
  ws->chunk = chunk;
  return RBD_OK;
}

struct Piece {  // one array of a host call: K fields, host pointer (may be null), direction
  int K;
  const double* h_in;
  double* h_out;
  bool load;   // copy host -> device before the kernel
  bool store;  // copy device -> host after the kernel
  double* d_soa = nullptr;
  double* d_aos = nullptr;
};

// The arrays of one host call: at most kMaxPieces, on the stack (no allocation on the hot path).
static const int kMaxPieces = 8;
struct Pieces {
  Piece p[kMaxPieces];
  int n = 0;
  Piece& operator[](int i) { return p[i]; }
  void add(int K, const double* h_in, double* h_out, bool load, bool store) {
    Piece& x = p[n++];
    x.K = K; x.h_in = h_in; x.h_out = h_out; x.load = load; x.store = store; x.d_soa = x.d_aos = nullptr;
  }
};

template <class KernelFn>
static int run_host_chunks(rbd_workspace* ws, int64_t n, Pieces& pieces, KernelFn kernel) {
  size_t fields = 0;
  for (int k = 0; k < pieces.n; ++k) fields += (size_t)pieces[k].K;
  int rc = ensure_slots(ws, fields);
  if (rc) return rc;
  // order after whatever the caller queued on the workspace stream
  CK(cudaEventRecord(ws->ev, ws->stream));
  for (int s = 0; s < kSlots; ++s) CK(cudaStreamWaitEvent(ws->slot[s].st, ws->ev, 0));
  const long long chunk = ws->chunk;
  const long long cpad = (chunk + 31) / 32 * 32, tile = 32;  // device staging is tile-32 (the kernels' fast path)
  int which = 0;
  for (long long i0 = 0; i0 < n; i0 += chunk, which ^= 1) {
    const long long cn = (n - i0) < chunk ? (n - i0) : chunk;
    HostSlot& sl = ws->slot[which];
    size_t off = 0;
    for (int k = 0; k < pieces.n; ++k) {
      Piece& p = pieces[k];
      p.d_aos = sl.aos + off;
      p.d_soa = sl.soa + off;
      off += (size_t)p.K * (size_t)cpad;
      const bool present = p.h_in != nullptr || p.h_out != nullptr;
      if (!present) { p.d_aos = p.d_soa = nullptr; continue; }
      if (p.load && p.K > 0) {
        CK(cudaMemcpyAsync(p.d_aos, p.h_in + (size_t)i0 * p.K, (size_t)cn * p.K * sizeof(double), cudaMemcpyHostToDevice, sl.st));
        CK(launch_aos_to_soa(cn, p.K, tile, p.d_aos, p.d_soa, sl.st));
        ws->launches++;
      }
    }
    rc = kernel(sl.st, i0, cn, tile);
    if (rc) return rc;
    for (int k = 0; k < pieces.n; ++k) {
      Piece& p = pieces[k];
      if (p.store && p.h_out && p.K > 0) {
        CK(launch_soa_to_aos(cn, p.K, tile, p.d_soa, p.d_aos, sl.st));
        ws->launches++;
        CK(cudaMemcpyAsync(p.h_out + (size_t)i0 * p.K, p.d_aos, (size_t)cn * p.K * sizeof(double), cudaMemcpyDeviceToHost, sl.st));
      }
    }
  }
  for (int s = 0; s < kSlots; ++s) CK(cudaStreamSynchronize(ws->slot[s].st));
  return RBD_OK;
}

static int eval_host_impl(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a, double* oMi,
                          double* J, double* M, double* tau, bool sparse_m) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q) return set_err(RBD_ERR_INVALID_ARG, "q is null");
  if (tau && (!v || !a)) return set_err(RBD_ERR_INVALID_ARG, "rnea needs v and a");
  const rbd_model_info& I = ws->model->info;
  // only what the requested outputs need travels: v and a are not copied when tau is not requested
  Pieces P;
  P.add(I.nq, q, nullptr, true, false);
  P.add(I.nv, tau ? v : nullptr, nullptr, true, false);
  P.add(I.nv, tau ? a : nullptr, nullptr, true, false);
  P.add(I.omi_fields, nullptr, oMi, false, true);
  P.add(I.jac_fields, nullptr, J, false, true);
  P.add(sparse_m ? I.mass_support_fields : I.mass_fields, nullptr, M, false, true);
  P.add(I.nv, nullptr, tau, false, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long, long long cn, long long tile) {
    return eval_on(ws, st, cn, tile, P[0].d_soa, P[1].d_soa, P[2].d_soa, P[3].d_soa, P[4].d_soa, P[5].d_soa, P[6].d_soa,
                   sparse_m);
  });
}
RBD_API int rbd_eval_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a, double* oMi,
                          double* J, double* M, double* tau) {
  return eval_host_impl(ws, n, q, v, a, oMi, J, M, tau, false);
}
RBD_API int rbd_eval_host_sparse(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a,
                                 double* oMi, double* J, double* Mval, double* tau) {
  return eval_host_impl(ws, n, q, v, a, oMi, J, Mval, tau, true);
}

// Mass / ForceField products on host vectors: the accumulated vector travels both ways (load, +=, store)
RBD_API int rbd_addMDx_host(rbd_workspace* ws, int64_t n, const double* q, const double* dx, double factor, double* res) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q || !dx || !res) return set_err(RBD_ERR_INVALID_ARG, "null q / dx / res");
  const rbd_model_info& I = ws->model->info;
  Pieces P;
  P.add(I.nq, q, nullptr, true, false); P.add(I.nv, dx, nullptr, true, false); P.add(I.nv, res, res, true, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long, long long cn, long long tile) {
    return mass_force_on(ws, st, cn, tile, P[0].d_soa, nullptr, P[1].d_soa, P[2].d_soa, factor, 0);
  });
}
RBD_API int rbd_addForce_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, double* f) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q || !f) return set_err(RBD_ERR_INVALID_ARG, "null q / f");
  const rbd_model_info& I = ws->model->info;
  Pieces P;
  P.add(I.nq, q, nullptr, true, false); P.add(I.nv, v, nullptr, true, false); P.add(I.nv, f, f, true, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long, long long cn, long long tile) {
    return mass_force_on(ws, st, cn, tile, P[0].d_soa, v ? P[1].d_soa : nullptr, nullptr, P[2].d_soa, -1.0, 1);
  });
}

RBD_API int rbd_aba_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* tau, double* ddq) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q || !v || !tau || !ddq) return set_err(RBD_ERR_INVALID_ARG, "null q / v / tau / ddq");
  const rbd_model_info& I = ws->model->info;
  Pieces P;
  P.add(I.nq, q, nullptr, true, false); P.add(I.nv, v, nullptr, true, false); P.add(I.nv, tau, nullptr, true, false);
  P.add(I.nv, nullptr, ddq, false, true);
  // (each chunk stream has its own scratch records: aba_on)
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long, long long cn, long long tile) {
    return aba_on(ws, st, cn, tile, P[0].d_soa, P[1].d_soa, P[2].d_soa, P[3].d_soa);
  });
}

RBD_API int rbd_minv_host(rbd_workspace* ws, int64_t n, const double* q, double* Minv) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q || !Minv) return set_err(RBD_ERR_INVALID_ARG, "null q / Minv");
  const rbd_model_info& I = ws->model->info;
  Pieces P;
  P.add(I.nq, q, nullptr, true, false);
  P.add(I.mass_fields, nullptr, Minv, false, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long, long long cn, long long tile) {
    return minv_on(ws, st, cn, tile, P[0].d_soa, P[1].d_soa);
  });
}

RBD_API int rbd_rnea_derivatives_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a,
                                      double* dtau_dq, double* dtau_dv, double* M) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q || !v || !a || !dtau_dq || !dtau_dv) return set_err(RBD_ERR_INVALID_ARG, "null q / v / a / dtau_dq / dtau_dv");
  const rbd_model_info& I = ws->model->info;
  Pieces P;
  P.add(I.nq, q, nullptr, true, false);
  P.add(I.nv, v, nullptr, true, false);
  P.add(I.nv, a, nullptr, true, false);
  P.add(I.nv * I.nv, nullptr, dtau_dq, false, true);
  P.add(I.nv * I.nv, nullptr, dtau_dv, false, true);
  P.add(I.mass_fields, nullptr, M, false, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long, long long cn, long long tile) {
    return drnea_on(ws, st, cn, tile, P[0].d_soa, P[1].d_soa, P[2].d_soa, P[3].d_soa, P[4].d_soa, P[5].d_soa);
  });
}

RBD_API int rbd_rnea_derivatives_host_sparse(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a,
                                             double* dq_val, double* dv_val) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q || !v || !a || !dq_val || !dv_val) return set_err(RBD_ERR_INVALID_ARG, "null q / v / a / dq_val / dv_val");
  const rbd_model_info& I = ws->model->info;
  const int nnz = drnea_support_pattern(ws->model->dm, nullptr, nullptr);
  Pieces P;
  P.add(I.nq, q, nullptr, true, false);
  P.add(I.nv, v, nullptr, true, false);
  P.add(I.nv, a, nullptr, true, false);
  P.add(nnz, nullptr, dq_val, false, true);
  P.add(nnz, nullptr, dv_val, false, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long, long long cn, long long tile) {
    return drnea_on(ws, st, cn, tile, P[0].d_soa, P[1].d_soa, P[2].d_soa, P[3].d_soa, P[4].d_soa, nullptr, true);
  });
}

RBD_API int rbd_apply_host(rbd_workspace* ws, int64_t n, const double* q_joints, const double* root_pose, double* out) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!q_joints || !out) return set_err(RBD_ERR_INVALID_ARG, "null q_joints / out");
  const rbd_model_info& I = ws->model->info;
  if (root_pose && !I.has_freeflyer_root) return set_err(RBD_ERR_INVALID_ARG, "root pose given but the model has no free-flyer root");
  const int kq = root_pose ? I.nq - 7 : I.nq;
  Pieces P;
  P.add(kq, q_joints, nullptr, true, false); P.add(7, root_pose, nullptr, true, false);
  P.add(7 * I.n_out, nullptr, out, false, true);
  rc = run_host_chunks(ws, n, P, [&](cudaStream_t st, long long i0, long long cn, long long tile) {
    return apply_on(ws, st, cn, tile, i0, P[0].d_soa, P[1].d_soa, P[2].d_soa);
  });
  if (rc == RBD_OK) ws->applied_n = n;
  return rc;
}

RBD_API int rbd_applyJ_host(rbd_workspace* ws, int64_t n, const double* dq_joints, const double* root_vel, double* out) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!dq_joints || !out) return set_err(RBD_ERR_INVALID_ARG, "null dq_joints / out");
  if (n > ws->applied_n) return set_err(RBD_ERR_NO_APPLY, "applyJ before apply: no cached configuration for these instances");
  const rbd_model_info& I = ws->model->info;
  if (root_vel && !I.has_freeflyer_root) return set_err(RBD_ERR_INVALID_ARG, "root velocity given but the model has no free-flyer root");
  const int kv = root_vel ? I.nv - 6 : I.nv;
  Pieces P;
  P.add(kv, dq_joints, nullptr, true, false); P.add(6, root_vel, nullptr, true, false);
  P.add(6 * I.n_out, nullptr, out, false, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long i0, long long cn, long long tile) {
    return applyJ_on(ws, st, cn, tile, i0, P[0].d_soa, P[1].d_soa, P[2].d_soa);
  });
}

RBD_API int rbd_applyJT_host(rbd_workspace* ws, int64_t n, const double* in, double* out_tau, double* out_root) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!in || !out_tau) return set_err(RBD_ERR_INVALID_ARG, "null in / out_tau");
  if (n > ws->applied_n) return set_err(RBD_ERR_NO_APPLY, "applyJT before apply: no cached configuration for these instances");
  const rbd_model_info& I = ws->model->info;
  if (out_root && !I.has_freeflyer_root) return set_err(RBD_ERR_INVALID_ARG, "root output given but the model has no free-flyer root");
  const int kv = out_root ? I.nv - 6 : I.nv;
  // out_tau / out_root are accumulated (+=, reference KinematicChainMapping.inl:498-511): load, add, store
  Pieces P;
  P.add(6 * I.n_out, in, nullptr, true, false); P.add(kv, out_tau, out_tau, true, true);
  P.add(6, out_root, out_root, true, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long i0, long long cn, long long tile) {
    return applyJT_on(ws, st, cn, tile, i0, P[0].d_soa, P[1].d_soa, P[2].d_soa);
  });
}

RBD_API int rbd_applyDJT_host(rbd_workspace* ws, int64_t n, const double* dq_joints, const double* root_vel,
                              const double* in, double kfactor, double* out_tau, double* out_root) {
  GUARD(ws);
  int rc = check_batch(ws, n, 1);
  if (rc) return rc;
  if (!dq_joints || !in || !out_tau) return set_err(RBD_ERR_INVALID_ARG, "null dq_joints / in / out_tau");
  if (n > ws->applied_n) return set_err(RBD_ERR_NO_APPLY, "applyDJT before apply: no cached configuration for these instances");
  const rbd_model_info& I = ws->model->info;
  if ((root_vel || out_root) && !I.has_freeflyer_root) return set_err(RBD_ERR_INVALID_ARG, "root arrays given but the model has no free-flyer root");
  const int kv = root_vel ? I.nv - 6 : I.nv;
  Pieces P;
  P.add(kv, dq_joints, nullptr, true, false); P.add(6, root_vel, nullptr, true, false);
  P.add(6 * I.n_out, in, nullptr, true, false); P.add(kv, out_tau, out_tau, true, true);
  P.add(6, out_root, out_root, true, true);
  return run_host_chunks(ws, n, P, [&](cudaStream_t st, long long i0, long long cn, long long tile) {
    return applyDJT_on(ws, st, cn, tile, i0, P[0].d_soa, P[1].d_soa, P[2].d_soa, kfactor, P[3].d_soa, P[4].d_soa);
  });
}

RBD_API int rbd_applyJT_rows_host(rbd_workspace* ws, int64_t nrows, const int32_t* row_ptr, const int32_t* row_instance,
                                  const int32_t* col, const double* val, double* out_rows, int32_t* keep) {
  GUARD(ws);
  if (nrows < 0) return set_err(RBD_ERR_INVALID_ARG, "negative row count");
  if (nrows == 0) return RBD_OK;
  if (!row_ptr || !row_instance || !out_rows || !keep) return set_err(RBD_ERR_INVALID_ARG, "null row arrays");
  if (ws->applied_n <= 0) return set_err(RBD_ERR_NO_APPLY, "applyJT(rows) before apply: no cached configuration");
  const int nv = ws->model->info.nv;
  const long long nnz = row_ptr[nrows];
  if (nnz < 0 || (nnz > 0 && (!col || !val))) return set_err(RBD_ERR_INVALID_ARG, "bad CSR arrays");
  // a malformed constraint matrix would index the per-output tables and the Jacobian cache out of bounds on the device
  if (row_ptr[0] != 0) return set_err(RBD_ERR_INVALID_ARG, "row_ptr[0] must be 0");
  const int n_out = ws->model->info.n_out;
  for (int64_t r = 0; r < nrows; ++r) {
    if (row_instance[r] < 0 || row_instance[r] >= ws->applied_n) return set_err(RBD_ERR_INVALID_ARG, "row_instance out of the applied range");
    if (row_ptr[r + 1] < row_ptr[r]) return set_err(RBD_ERR_INVALID_ARG, "row_ptr is not non-decreasing");
  }
  for (long long j = 0; j < nnz; ++j)
    if (col[j] < 0 || col[j] >= n_out) return set_err(RBD_ERR_INVALID_ARG, "col entry is not a mapped-output index (0 <= col < n_out)");
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_rp = al((size_t)(nrows + 1) * 4), b_ri = al((size_t)nrows * 4), b_col = al((size_t)nnz * 4),
               b_val = al((size_t)nnz * 48), b_out = al((size_t)nrows * nv * 8), b_keep = al((size_t)nrows * 4);
  const size_t total = b_rp + b_ri + b_col + b_val + b_out + b_keep;
  if (ws->rows_bytes < total) {
    CK(cudaStreamSynchronize(ws->stream));
    cudaFree(ws->rows_buf); ws->rows_buf = nullptr; ws->rows_bytes = 0;
    CK(cudaMalloc(&ws->rows_buf, total));
    ws->rows_bytes = total;
  }
  char* base = (char*)ws->rows_buf;
  int* d_rp = (int*)base; int* d_ri = (int*)(base + b_rp); int* d_col = (int*)(base + b_rp + b_ri);
  double* d_val = (double*)(base + b_rp + b_ri + b_col);
  double* d_out = (double*)(base + b_rp + b_ri + b_col + b_val);
  int* d_keep = (int*)(base + b_rp + b_ri + b_col + b_val + b_out);
  cudaStream_t st = ws->stream;
  CK(cudaMemcpyAsync(d_rp, row_ptr, (size_t)(nrows + 1) * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_ri, row_instance, (size_t)nrows * 4, cudaMemcpyHostToDevice, st));
  if (nnz > 0) {
    CK(cudaMemcpyAsync(d_col, col, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_val, val, (size_t)nnz * 48, cudaMemcpyHostToDevice, st));
  }
  const DevModel& dm = ws->model->dm;
  RowArgs A{nrows, d_rp, d_ri, d_col, d_val, d_out, d_keep, ws->qcache};
  (void)dm;
  { int rc2 = rows_on(ws, st, A); if (rc2) return rc2; }
  CK(cudaMemcpyAsync(out_rows, d_out, (size_t)nrows * nv * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(keep, d_keep, (size_t)nrows * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RBD_OK;
}



// ---- multi-robot batching context ---------------------------------------------------------------------------------
// Host-side gather / scatter around the *_host entry points: the robots' SOFA vectors are scattered host buffers, the
// library wants one instance-major array per quantity.  Staging is pinned so that the chunked copies run
// asynchronously.
struct rbd_batch {
  rbd_workspace* ws = nullptr;
  long long n = 0;
  std::vector<const double*> a_q, a_root, j_dq, j_root, t_in;
  std::vector<double*> a_out, j_out, t_tau, t_rootout;
  std::vector<char> a_bound, j_bound, t_bound;
  double *h_in1 = nullptr, *h_in2 = nullptr, *h_out = nullptr, *h_out2 = nullptr;  // pinned staging
  long long epoch_apply = -1, epoch_J = -1, epoch_JT = -1;
  bool any_epoch_apply = false, any_epoch_J = false, any_epoch_JT = false;
};

RBD_API int rbd_batch_create(rbd_workspace* ws, int64_t n_slots, rbd_batch** out) {
  if (!out) return set_err(RBD_ERR_INVALID_ARG, "null out");
  *out = nullptr;
  GUARD(ws);
  if (n_slots < 1 || n_slots > ws->capacity) return set_err(RBD_ERR_INVALID_ARG, "n_slots must be in [1, workspace capacity]");
  rbd_batch* b = new (std::nothrow) rbd_batch();
  if (!b) return set_err(RBD_ERR_ALLOC, "out of host memory");
  b->ws = ws; b->n = n_slots;
  const size_t n = (size_t)n_slots;
  b->a_q.assign(n, nullptr); b->a_root.assign(n, nullptr); b->j_dq.assign(n, nullptr); b->j_root.assign(n, nullptr);
  b->t_in.assign(n, nullptr); b->a_out.assign(n, nullptr); b->j_out.assign(n, nullptr); b->t_tau.assign(n, nullptr);
  b->t_rootout.assign(n, nullptr); b->a_bound.assign(n, 0); b->j_bound.assign(n, 0); b->t_bound.assign(n, 0);
  const rbd_model_info& I = ws->model->info;
  const size_t big = (size_t)(7 > 6 ? 7 : 6) * (size_t)I.n_out;  // poses (7) / twists and wrenches (6) per robot
  const size_t in1 = (size_t)(I.nq > I.nv ? I.nq : I.nv);
  cudaError_t e = cudaHostAlloc(&b->h_in1, n * (in1 > big ? in1 : big) * sizeof(double), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&b->h_in2, n * 7 * sizeof(double), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&b->h_out, n * big * sizeof(double), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&b->h_out2, n * 7 * sizeof(double), cudaHostAllocDefault);
  if (e != cudaSuccess) { rbd_batch_destroy(b); return cuda_fail(e, "cudaHostAlloc(batch staging)"); }
  *out = b;
  return RBD_OK;
}

RBD_API void rbd_batch_destroy(rbd_batch* b) {
  if (!b) return;
  cudaFreeHost(b->h_in1); cudaFreeHost(b->h_in2); cudaFreeHost(b->h_out); cudaFreeHost(b->h_out2);
  delete b;
}

static int batch_slot_ok(const rbd_batch* b, int64_t slot) {
  if (!b) return set_err(RBD_ERR_INVALID_ARG, "null batch");
  if (slot < 0 || slot >= b->n) return set_err(RBD_ERR_INVALID_ARG, "slot out of range");
  return RBD_OK;
}

RBD_API int rbd_batch_bind_apply(rbd_batch* b, int64_t slot, const double* q_joints, const double* root_pose, double* out) {
  int rc = batch_slot_ok(b, slot);
  if (rc) return rc;
  const rbd_model_info& I = b->ws->model->info;
  if (!q_joints || !out) return set_err(RBD_ERR_INVALID_ARG, "null q_joints / out");
  if ((root_pose != nullptr) != (I.has_freeflyer_root != 0))
    return set_err(RBD_ERR_INVALID_ARG, "root pose must be bound iff the model has a free-flyer root");
  b->a_q[slot] = q_joints; b->a_root[slot] = root_pose; b->a_out[slot] = out; b->a_bound[slot] = 1;
  return RBD_OK;
}
RBD_API int rbd_batch_bind_applyJ(rbd_batch* b, int64_t slot, const double* dq_joints, const double* root_vel, double* out) {
  int rc = batch_slot_ok(b, slot);
  if (rc) return rc;
  const rbd_model_info& I = b->ws->model->info;
  if (!dq_joints || !out) return set_err(RBD_ERR_INVALID_ARG, "null dq_joints / out");
  if ((root_vel != nullptr) != (I.has_freeflyer_root != 0))
    return set_err(RBD_ERR_INVALID_ARG, "root velocity must be bound iff the model has a free-flyer root");
  b->j_dq[slot] = dq_joints; b->j_root[slot] = root_vel; b->j_out[slot] = out; b->j_bound[slot] = 1;
  return RBD_OK;
}
RBD_API int rbd_batch_bind_applyJT(rbd_batch* b, int64_t slot, const double* in, double* out_tau, double* out_root) {
  int rc = batch_slot_ok(b, slot);
  if (rc) return rc;
  const rbd_model_info& I = b->ws->model->info;
  if (!in || !out_tau) return set_err(RBD_ERR_INVALID_ARG, "null in / out_tau");
  if ((out_root != nullptr) != (I.has_freeflyer_root != 0))
    return set_err(RBD_ERR_INVALID_ARG, "root output must be bound iff the model has a free-flyer root");
  b->t_in[slot] = in; b->t_tau[slot] = out_tau; b->t_rootout[slot] = out_root; b->t_bound[slot] = 1;
  return RBD_OK;
}

static int batch_all_bound(const std::vector<char>& v, const char* what) {
  for (char c : v)
    if (!c) return set_err(RBD_ERR_INVALID_ARG, std::string("not every slot is bound for ") + what);
  return RBD_OK;
}

RBD_API int rbd_batch_apply(rbd_batch* b, int64_t slot, int64_t epoch) {
  int rc = batch_slot_ok(b, slot);
  if (rc) return rc;
  if (b->any_epoch_apply && b->epoch_apply == epoch) return 0;
  if ((rc = batch_all_bound(b->a_bound, "apply"))) return rc;
  const rbd_model_info& I = b->ws->model->info;
  const bool ff = I.has_freeflyer_root != 0;
  const size_t kq = (size_t)I.nq_joints, ko = 7 * (size_t)I.n_out;
  for (long long i = 0; i < b->n; ++i) {
    memcpy(b->h_in1 + (size_t)i * kq, b->a_q[i], kq * sizeof(double));
    if (ff) memcpy(b->h_in2 + (size_t)i * 7, b->a_root[i], 7 * sizeof(double));
  }
  rc = rbd_apply_host(b->ws, b->n, b->h_in1, ff ? b->h_in2 : nullptr, b->h_out);
  if (rc) return rc;
  for (long long i = 0; i < b->n; ++i) memcpy(b->a_out[i], b->h_out + (size_t)i * ko, ko * sizeof(double));
  b->epoch_apply = epoch; b->any_epoch_apply = true;
  return 1;
}

RBD_API int rbd_batch_applyJ(rbd_batch* b, int64_t slot, int64_t epoch) {
  int rc = batch_slot_ok(b, slot);
  if (rc) return rc;
  if (b->any_epoch_J && b->epoch_J == epoch) return 0;
  if ((rc = batch_all_bound(b->j_bound, "applyJ"))) return rc;
  const rbd_model_info& I = b->ws->model->info;
  const bool ff = I.has_freeflyer_root != 0;
  const size_t kv = (size_t)I.nv_joints, ko = 6 * (size_t)I.n_out;
  for (long long i = 0; i < b->n; ++i) {
    memcpy(b->h_in1 + (size_t)i * kv, b->j_dq[i], kv * sizeof(double));
    if (ff) memcpy(b->h_in2 + (size_t)i * 6, b->j_root[i], 6 * sizeof(double));
  }
  rc = rbd_applyJ_host(b->ws, b->n, b->h_in1, ff ? b->h_in2 : nullptr, b->h_out);
  if (rc) return rc;
  for (long long i = 0; i < b->n; ++i) memcpy(b->j_out[i], b->h_out + (size_t)i * ko, ko * sizeof(double));
  b->epoch_J = epoch; b->any_epoch_J = true;
  return 1;
}

RBD_API int rbd_batch_applyJT(rbd_batch* b, int64_t slot, int64_t epoch) {
  int rc = batch_slot_ok(b, slot);
  if (rc) return rc;
  if (b->any_epoch_JT && b->epoch_JT == epoch) return 0;
  if ((rc = batch_all_bound(b->t_bound, "applyJT"))) return rc;
  const rbd_model_info& I = b->ws->model->info;
  const bool ff = I.has_freeflyer_root != 0;
  const size_t kv = (size_t)I.nv_joints, ki = 6 * (size_t)I.n_out;
  // the library accumulates (+=) into out_tau / out_root: gather their current contents, scatter the sums back
  for (long long i = 0; i < b->n; ++i) {
    memcpy(b->h_in1 + (size_t)i * ki, b->t_in[i], ki * sizeof(double));
    memcpy(b->h_out + (size_t)i * kv, b->t_tau[i], kv * sizeof(double));
    if (ff) memcpy(b->h_out2 + (size_t)i * 6, b->t_rootout[i], 6 * sizeof(double));
  }
  rc = rbd_applyJT_host(b->ws, b->n, b->h_in1, b->h_out, ff ? b->h_out2 : nullptr);
  if (rc) return rc;
  for (long long i = 0; i < b->n; ++i) {
    memcpy(b->t_tau[i], b->h_out + (size_t)i * kv, kv * sizeof(double));
    if (ff) memcpy(b->t_rootout[i], b->h_out2 + (size_t)i * 6, 6 * sizeof(double));
  }
  b->epoch_JT = epoch; b->any_epoch_JT = true;
  return 1;
}
