// cvb200.cu -- C ABI of libcvb200.so (see include/cvb200.h for the contract and the
// reference call sites each entry point replaces).
#include "../../include/cvb200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <string>
#include <vector>

#include "fwd_simt.cuh"
#include "fc4_tc.cuh"
#include "conv_tc.cuh"
#include "conv_tc_slab.cuh"
#include "tail_tc.cuh"
#include <math.h>
#include <stdlib.h>
#include <sys/mman.h>
#include "train_simt.cuh"
#include "gemm_tc.cuh"

using namespace cvb;

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

extern "C" const char* cvb_last_error(void) { return g_err; }
// for the host-only translation units of the library (text_feed.cpp)
void cvb_internal_set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
extern "C" int cvb_version(void) { return 100; }

// ------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------
struct VarInfo {
  std::string name;
  int ndim;
  int64_t dims[4];
  int64_t numel, offset;
};

// sites per device pass: bounds the intermediates and makes k_fc4's grid exactly one wave
// (M-tile 96 sites x #SMs for v3; 224 x #SMs for slim)

static const int64_t kFc4SplitSites = 48 * 128;  // largest batch whose FC4 is split over K (96 CTAs x 3 slices = two waves)

struct cvb_model {
  int variant = 0, device = 0, compute_mode = CVB_COMPUTE_FP32, num_sms = 148;
  int64_t CHUNK = 96 * 148;
  std::vector<VarInfo> vars;
  int64_t nparams = 0, step = 0;
  float *d_params = nullptr, *d_m = nullptr, *d_v = nullptr, *d_grad = nullptr;
  // forward work buffers
  float *d_p2 = nullptr, *d_p3 = nullptr, *d_h4 = nullptr, *d_h5 = nullptr;
  int64_t p2_site = 0, p3_site = 0, h4_site = 0;
  // host path: 2 slots
  // host-call pipeline slots.  Two are not enough: with H2D(c) ~ kernels(c) ~ 0.7 ms the copy engine only gets its next
  // transfer after the host has seen D2H(c-1), so every host wake-up and de-interleave shows up as a PCIe bubble.
  static constexpr int NSLOT = 4;
  float *d_x[NSLOT] = {}, *d_out[NSLOT] = {}, *d_lg[NSLOT] = {};
  __half* d_x16[NSLOT] = {};  // narrow input slots (fp16 values, int16 / uint8 raw counts: cvb_predict_host_f16 / _counts_*)
  float* d_fc4ws = nullptr;   // split-K FC4 of small batches: per-chunk partial sums [9][kFc4SplitSites][352]
  float* d_xw = nullptr;      // fp32 scratch of one chunk for the front kernels that cannot read a narrow feed themselves
  float *h_x[NSLOT] = {}, *h_out[NSLOT] = {}, *h_lg[NSLOT] = {};
  cudaStream_t s_comp = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  cudaStream_t s_aux = nullptr;  // second branch of the backward pass: weight gradients run beside the data-gradient chain
  cudaEvent_t e_fork = nullptr, e_join = nullptr;
  cudaEvent_t e_h2d[NSLOT], e_comp[NSLOT], e_d2h[NSLOT];
  // cvb_predict_submit / cvb_predict_collect: one small batch per slot in flight
  struct Ticket { bool busy = false; bool logits = false; int64_t n = 0; uint32_t gen = 0; };
  Ticket tickets[NSLOT];
  int next_slot = 0;
  int tickets_out() const { int k = 0; for (auto& t : tickets) k += t.busy; return k; }
  bool events = false;
  int64_t launches = 0;
  // tensor-core path (CVB_COMPUTE_FP16X3): pre-split FC4 weights + TMA descriptors
  __half *d_w4t_hi = nullptr, *d_w4t_lo = nullptr;
  unsigned int* d_absmax = nullptr;
  float* d_inv_scale = nullptr;
  bool tc_ready = false, tc_weights_dirty = true;
  CUtensorMap map_a_hi, map_a_lo, map_b_hi, map_b_lo;
  CUtensorMap map_bh_hi, map_bh_lo;  // FC4 weights with NH/2-row boxes (cluster multicast)
  // conv3 on tensor cores: B = rearranged conv3 weights [3*192][128], A = p2 hi/lo [sites*28][128]
  __half *d_w3b_hi = nullptr, *d_w3b_lo = nullptr;
  // conv2 on tensor cores: A = p1 hi/lo [sites*30][64] (written by k_v3_c1_reg), B = rearranged conv2 weights [2*128][64]
  __half *d_p1 = nullptr, *d_w2b_hi = nullptr, *d_w2b_lo = nullptr;
  // rows per fp16 plane of p1 / p2 (sites*RPS + slack); lo plane = hi plane + rows*KROW
  int64_t p1_rows = 0, p2_rows = 0;
  size_t p2_bytes = 0;
  CUtensorMap map_c2b2, map_c2b3, map_c2b4, map_c3b2, map_c3b3, map_c3b4;  // weight boxes of 2, 3, 4 output columns, hi + lo planes
  // v3_slim tensor path: every layer but conv1 and the tail on tcgen05
  __half *d_p1s = nullptr, *d_w2s = nullptr, *d_w4h = nullptr;  // p1 [rows][32] hi|lo; dense conv2 taps hi|lo; fc4/kernel^T [36][4224] hi|lo
  int64_t p1s_rows = 0;
  CUtensorMap map_s2slab, map_s2b;
  tc::BiasParam hb2 = {}, hb3 = {};  // host copies of conv2/bias, conv3/bias: passed by value to the inference conv kernels
  CUtensorMap map_c2slab, map_c3slab;
  // fused tail (FC5 + heads) on tensor cores: A = h4 hi/lo [sites][336], B = [W5 | Wb]^T [176][336]
  __half *d_h4s = nullptr, *d_wtail = nullptr;
  CUtensorMap map_ta_hi, map_ta_lo, map_tb_hi, map_tb_lo;
  int64_t alloc_sites = 0;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;  // 6 per chunk: start, after SIMT front, conv2(tc), conv3, fc4, tail
  size_t prof_used = 0;
  TrainWork* train = nullptr;
  int train_mode = CVB_TRAIN_BF16X3;  // arithmetic of the FC4 contractions in a training step (v3)
  void* nccl_comm = nullptr;  // data-parallel training: gradient all-reduce on s_comp inside cvb_train_step_host / cvb_apply_adam
  bool nccl_owned = false;
  int nccl_ranks = 1;
  float drop5 = 0.f, drop5_now = 0.f;  // dropoutRateFC5 (cvb_set_dropout_fc5) / the rate of the pass being run (0 in getLoss)
  const float* var(const char* n) const {
    for (auto& v : vars)
      if (v.name == n) return d_params + v.offset;
    return nullptr;
  }
  const VarInfo* info(const char* n) const {
    for (auto& v : vars)
      if (v.name == n) return &v;
    return nullptr;
  }
};

static void add_var(cvb_model* m, const std::string& name, std::initializer_list<int64_t> dims) {
  VarInfo v;
  v.name = name;
  v.ndim = (int)dims.size();
  v.numel = 1;
  int i = 0;
  for (auto d : dims) { v.dims[i++] = d; v.numel *= d; }
  for (; i < 4; ++i) v.dims[i] = 1;
  v.offset = m->nparams;
  m->nparams += (v.numel + 3) / 4 * 4;  // keep every variable 16-byte aligned
  m->vars.push_back(v);
}

static void build_vars(cvb_model* m) {
  // names/shapes: jupyter_nb/visualization.ipynb:100-121; clairvoyante_v3_slim.py:9-11
  if (m->variant == CVB_V3) {
    add_var(m, "conv1/kernel", {1, 4, 4, 16});   add_var(m, "conv1/bias", {16});
    add_var(m, "conv2/kernel", {2, 4, 16, 32});  add_var(m, "conv2/bias", {32});
    add_var(m, "conv3/kernel", {3, 4, 32, 48});  add_var(m, "conv3/bias", {48});
    add_var(m, "fc4/kernel", {4608, 336});       add_var(m, "fc4/bias", {336});
    add_var(m, "fc5/kernel", {336, 168});        add_var(m, "fc5/bias", {168});
    add_var(m, "YBaseChangeSigmoid/kernel", {336, 4}); add_var(m, "YBaseChangeSigmoid/bias", {4});
    add_var(m, "YZygosityFC/kernel", {168, 2});        add_var(m, "YZygosityFC/bias", {2});
    add_var(m, "YVarTypeFC/kernel", {168, 4});         add_var(m, "YVarTypeFC/bias", {4});
    add_var(m, "YIndelLengthFC/kernel", {168, 6});     add_var(m, "YIndelLengthFC/bias", {6});
    m->p2_site = 28 * 128; m->p3_site = 4608; m->h4_site = 336;
  } else {
    add_var(m, "conv1/kernel", {1, 4, 4, 8});    add_var(m, "conv1/bias", {8});
    add_var(m, "conv2/kernel", {3, 4, 8, 16});   add_var(m, "conv2/bias", {16});
    add_var(m, "conv3/kernel", {5, 4, 16, 32});  add_var(m, "conv3/bias", {32});
    add_var(m, "fc4/kernel", {4224, 36});        add_var(m, "fc4/bias", {36});
    add_var(m, "fc5/kernel", {36, 18});          add_var(m, "fc5/bias", {18});
    add_var(m, "YBaseChangeSigmoid/kernel", {36, 4}); add_var(m, "YBaseChangeSigmoid/bias", {4});
    add_var(m, "YZygosityFC/kernel", {18, 2});        add_var(m, "YZygosityFC/bias", {2});
    add_var(m, "YVarTypeFC/kernel", {18, 4});         add_var(m, "YVarTypeFC/bias", {4});
    add_var(m, "YIndelLengthFC/kernel", {18, 6});     add_var(m, "YIndelLengthFC/bias", {6});
    m->p2_site = 37 * 64; m->p3_site = 4224; m->h4_site = 36;
  }
}

// ------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen of the libnccl.so.2 the process already holds, e.g. torch's): the library has no
// link-time dependency on it.  Prototypes restated from nccl.h (ncclFloat32 = 7, ncclSum = 0).
// ------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*CommCount)(void*, int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {getenv("CVB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names)
    if (nm && *nm && (h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) { fail("NCCL not found (dlopen libnccl.so.2: %s); set CVB_NCCL_LIB", dlerror()); return nullptr; }
  api.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  api.CommCount = (int (*)(void*, int*))dlsym(h, "ncclCommCount");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.CommCount) {
    fail("libnccl lacks an expected symbol");
    return nullptr;
  }
  api.lib = h;
  return &api;
}
#define NK(api, call)                                                                                              \
  do {                                                                                                             \
    int r_ = (call);                                                                                               \
    if (r_ != 0) return fail("%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, r_,                      \
                             (api)->GetErrorString ? (api)->GetErrorString(r_) : "?");                             \
  } while (0)

static void nccl_release(cvb_model* m);
static int allreduce_gradients(cvb_model* m, cudaStream_t st);

extern "C" int cvb_create(int variant, int device, cvb_model** out) {
  if (!out) return fail("cvb_create: out is NULL");
  if (variant != CVB_V3 && variant != CVB_V3_SLIM) return fail("cvb_create: unknown variant %d", variant);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("cvb_create: device %d not in [0,%d)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("cvb_create: libcvb200 needs an sm_100 GPU, device %d is sm_%d%d", device, prop.major, prop.minor);
  cvb_model* m = new cvb_model();
  m->variant = variant;
  m->device = device;
  m->num_sms = prop.multiProcessorCount;
  build_vars(m);
  m->CHUNK = (int64_t)m->num_sms * (variant == CVB_V3 ? 96 : 224);
  m->alloc_sites = (int64_t)m->num_sms * 224;  // every mode's chunk fits (fp32: 96/SM, tensor: 128/SM, slim: 224/SM)
  const int64_t CHUNK = m->alloc_sites;
  const size_t pb = (size_t)m->nparams * 4;
  CK(cudaMalloc(&m->d_params, pb)); CK(cudaMemset(m->d_params, 0, pb));
  CK(cudaMalloc(&m->d_m, pb));      CK(cudaMemset(m->d_m, 0, pb));
  CK(cudaMalloc(&m->d_v, pb));      CK(cudaMemset(m->d_v, 0, pb));
  CK(cudaMalloc(&m->d_grad, pb + 256)); CK(cudaMemset(m->d_grad, 0, pb + 256));
  if (variant == CVB_V3) m->p2_rows = ((CHUNK * 28 + 160 + 29) / 30) * 30;   // multiple of conv3's quadrant step
  else m->p2_rows = ((CHUNK * 37 + 160 + 31) / 32) * 32;                      // slim: 37 rows/site, quadrant step 32
  m->p1_rows = ((CHUNK * 30 + 160 + 28) / 29) * 29;
  m->p2_bytes = std::max((size_t)CHUNK * m->p2_site * 4, (size_t)m->p2_rows * (variant == CVB_V3 ? 128 : 64) * 2 * 2) + 4096;
  const size_t p2_bytes = m->p2_bytes;
  CK(cudaMalloc(&m->d_p2, p2_bytes)); CK(cudaMemset(m->d_p2, 0, p2_bytes));
  CK(cudaMalloc(&m->d_p3, (size_t)CHUNK * m->p3_site * 4));
  CK(cudaMalloc(&m->d_h4, (size_t)CHUNK * m->h4_site * 4));
  CK(cudaMalloc(&m->d_h5, (size_t)CHUNK * 168 * 4));
  CK(cudaStreamCreateWithFlags(&m->s_comp, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&m->s_h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&m->s_d2h, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&m->s_aux, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&m->e_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&m->e_join, cudaEventDisableTiming));
  for (int i = 0; i < cvb_model::NSLOT; ++i) {
    CK(cudaEventCreateWithFlags(&m->e_h2d[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->e_comp[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->e_d2h[i], cudaEventDisableTiming));
  }
  m->events = true;
  *out = m;
  return 0;
}

extern "C" int cvb_destroy(cvb_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  train_work_free(m->train);
  nccl_release(m);
  for (auto e : m->prof_events) cudaEventDestroy(e);
  cudaFree(m->d_params); cudaFree(m->d_m); cudaFree(m->d_v); cudaFree(m->d_grad);
  cudaFree(m->d_p2); cudaFree(m->d_p3); cudaFree(m->d_h4); cudaFree(m->d_h5);
  cudaFree(m->d_w3b_hi); cudaFree(m->d_p1); cudaFree(m->d_w2b_hi); cudaFree(m->d_h4s); cudaFree(m->d_wtail);
  cudaFree(m->d_p1s); cudaFree(m->d_w2s); cudaFree(m->d_w4h);
  cudaFree(m->d_w4t_hi); cudaFree(m->d_w4t_lo); cudaFree(m->d_absmax); cudaFree(m->d_inv_scale);
  for (int i = 0; i < cvb_model::NSLOT; ++i) {
    if (i == 0) { cudaFree(m->d_xw); cudaFree(m->d_fc4ws); }
    cudaFree(m->d_x[i]); cudaFree(m->d_x16[i]); cudaFree(m->d_out[i]); cudaFree(m->d_lg[i]);
    cudaFreeHost(m->h_x[i]); cudaFreeHost(m->h_out[i]); cudaFreeHost(m->h_lg[i]);
    if (m->events) { cudaEventDestroy(m->e_h2d[i]); cudaEventDestroy(m->e_comp[i]); cudaEventDestroy(m->e_d2h[i]); }
  }
  if (m->s_comp) cudaStreamDestroy(m->s_comp);
  if (m->s_h2d) cudaStreamDestroy(m->s_h2d);
  if (m->s_d2h) cudaStreamDestroy(m->s_d2h);
  if (m->s_aux) cudaStreamDestroy(m->s_aux);
  if (m->e_fork) cudaEventDestroy(m->e_fork);
  if (m->e_join) cudaEventDestroy(m->e_join);
  delete m;
  return 0;
}

extern "C" int cvb_num_variables(const cvb_model* m) { return m ? (int)m->vars.size() : 0; }
extern "C" int64_t cvb_num_parameters(const cvb_model* m) {
  int64_t t = 0;
  if (m) for (auto& v : m->vars) t += v.numel;
  return t;
}
extern "C" int cvb_variable_info(const cvb_model* m, int idx, char* name, int name_cap, int64_t* numel, int* ndim,
                                 int64_t dims[4]) {
  if (!m || idx < 0 || idx >= (int)m->vars.size()) return fail("cvb_variable_info: bad index %d", idx);
  const VarInfo& v = m->vars[idx];
  if (name && name_cap > 0) { strncpy(name, v.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (numel) *numel = v.numel;
  if (ndim) *ndim = v.ndim;
  if (dims) for (int i = 0; i < 4; ++i) dims[i] = v.dims[i];
  return 0;
}

static float* slot_ptr(cvb_model* m, int slot) {
  return slot == 0 ? m->d_params : slot == 1 ? m->d_m : slot == 2 ? m->d_v : nullptr;
}
extern "C" int cvb_set_variable(cvb_model* m, const char* name, int slot, const float* host, int64_t n) {
  if (!m || !name || !host) return fail("cvb_set_variable: NULL argument");
  if (m->tickets_out()) return fail("cvb_set_variable: %d submitted batch(es) still in flight (cvb_predict_collect)", m->tickets_out());
  const VarInfo* v = m->info(name);
  if (!v) return fail("cvb_set_variable: no variable named '%s'", name);
  if (n != v->numel) return fail("cvb_set_variable: '%s' has %lld elements, got %lld", name, (long long)v->numel, (long long)n);
  float* base = slot_ptr(m, slot);
  if (!base) return fail("cvb_set_variable: bad slot %d", slot);
  CK(cudaSetDevice(m->device));
  CK(cudaMemcpy(base + v->offset, host, (size_t)n * 4, cudaMemcpyHostToDevice));
  if (slot == 0) m->tc_weights_dirty = true;
  return 0;
}
extern "C" int cvb_get_variable(cvb_model* m, const char* name, int slot, float* host, int64_t n) {
  if (!m || !name || !host) return fail("cvb_get_variable: NULL argument");
  const VarInfo* v = m->info(name);
  if (!v) return fail("cvb_get_variable: no variable named '%s'", name);
  if (n != v->numel) return fail("cvb_get_variable: '%s' has %lld elements, got %lld", name, (long long)v->numel, (long long)n);
  float* base = slot_ptr(m, slot);
  if (!base) return fail("cvb_get_variable: bad slot %d", slot);
  CK(cudaSetDevice(m->device));
  CK(cudaStreamSynchronize(m->s_comp));
  CK(cudaMemcpy(host, base + v->offset, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}
// init_op (clairvoyante_v3.py:177-178): the reference's initialisers, drawn on the host from a counter-based stream --
// conv* / fc4 / fc5 kernels: tf.contrib.layers.variance_scaling_initializer(factor=2.0, mode='FAN_IN', uniform=False)
// (clairvoyante_v3.py:57,72,87,106,116) = normal with stddev sqrt(1.3 * 2 / fan_in), redrawn outside +-2 sigma; head kernels:
// tf.layers.dense's default glorot_uniform (:125-135); biases zero.  Also zeroes the Adam slots and the step counter.
// (The reference is unseeded; the Python twin initializers.init_weights uses NumPy's generator, so the two streams differ.)
static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
extern "C" int cvb_init_weights(cvb_model* m, uint64_t seed) {
  if (!m) return fail("cvb_init_weights: NULL model");
  if (m->tickets_out()) return fail("cvb_init_weights: %d submitted batch(es) still in flight (cvb_predict_collect)", m->tickets_out());
  CK(cudaSetDevice(m->device));
  std::vector<float> h((size_t)m->nparams, 0.f);
  uint64_t ctr = 0;
  const uint64_t key = mix64(seed + 0x9E3779B97F4A7C15ull);
  auto uni = [&]() { return (double)(mix64(key + (++ctr) * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0); };  // [0,1)
  for (auto& v : m->vars) {
    if (v.name.find("bias") != std::string::npos) continue;
    float* w = h.data() + v.offset;
    const int64_t fan_out = v.dims[v.ndim - 1], fan_in = v.numel / fan_out;
    if (v.name[0] == 'Y') {
      const double limit = sqrt(6.0 / (double)(fan_in + fan_out));
      for (int64_t i = 0; i < v.numel; ++i) w[i] = (float)((2.0 * uni() - 1.0) * limit);
    } else {
      const double sd = sqrt(1.3 * 2.0 / (double)fan_in);
      for (int64_t i = 0; i < v.numel; ++i) {
        double z;
        do {  // Box-Muller, redrawn outside two standard deviations (tf.truncated_normal)
          const double u1 = 1.0 - uni(), u2 = uni();
          z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        } while (fabs(z) > 2.0);
        w[i] = (float)(z * sd);
      }
    }
  }
  CK(cudaStreamSynchronize(m->s_comp));
  const size_t pb = (size_t)m->nparams * 4;
  CK(cudaMemcpy(m->d_params, h.data(), pb, cudaMemcpyHostToDevice));
  CK(cudaMemset(m->d_m, 0, pb));
  CK(cudaMemset(m->d_v, 0, pb));
  m->step = 0;
  m->tc_weights_dirty = true;
  return 0;
}

// ------------------------------------------------------------------------------------
// tensor-core path setup: TMA descriptors (driver entry point fetched through the runtime,
// so the library has no link-time dependency on libcuda) and pre-split FC4 weights
// ------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
// fp16 row-major [rows][cols] tensor, box = box_cols x box_rows, given swizzle
static int make_map_f16(CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows,
                        CUtensorMapSwizzle sw) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// general fp16 tensor map: dims/box innermost first, strides (bytes) for dims 1..rank-1
static int make_map_nd(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, CUtensorMapSwizzle sw) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(rank %d) failed with CUresult %d", rank, (int)r);
  return 0;
}

// weight views for k_conv_slab: 3-D (k, row, plane) over the prepared taps at `wts` (hi plane then lo plane), one box
// {BK, nb*COUT, 2} per number nb of output columns an input column feeds
template <class C>
static int make_conv_weight_maps(void* wts, CUtensorMapSwizzle sw, CUtensorMap* b2, CUtensorMap* b3, CUtensorMap* b4) {
  const uint64_t bd[3] = {(uint64_t)C::KROW, (uint64_t)C::B_ROWS_TOTAL, 2};
  const uint64_t bs[2] = {(uint64_t)C::KROW * 2, (uint64_t)C::B_ROWS_TOTAL * C::KROW * 2};
  CUtensorMap* bm[3] = {b2, b3, b4};
  for (int nb = 2; nb <= 4; ++nb) {
    const uint32_t bb[3] = {(uint32_t)C::BK, (uint32_t)(nb * C::COUT), 2};
    if (make_map_nd(bm[nb - 2], wts, 3, bd, bs, bb, sw)) return 1;
  }
  return 0;
}

// slab view for k_conv_slab: plain 3-D (k, row, plane) tensor, box {BK, SLAB_ROWS, 2}
template <class C, class S>
static int make_slab_map(void* act, int64_t rows, CUtensorMapSwizzle sw, CUtensorMap* out) {
  const uint64_t d[3] = {(uint64_t)C::KROW, (uint64_t)rows, 2};
  const uint64_t st[2] = {(uint64_t)C::KROW * 2, (uint64_t)rows * C::KROW * 2};
  const uint32_t b[3] = {(uint32_t)C::BK, (uint32_t)S::SLAB_ROWS, 2};
  return make_map_nd(out, act, 3, d, st, b, sw);
}

static int tc_setup(cvb_model* m) {
  if (m->tc_ready) return 0;
  using F = tc::Fc4Tc;
  const int K = 4608;
  CK(cudaMalloc(&m->d_w4t_hi, (size_t)F::N * K * 2));
  CK(cudaMalloc(&m->d_w4t_lo, (size_t)F::N * K * 2));
  CK(cudaMalloc(&m->d_absmax, 16));
  CK(cudaMalloc(&m->d_inv_scale, 16));
  __half* a_hi = reinterpret_cast<__half*>(m->d_p3);
  __half* a_lo = a_hi + m->alloc_sites * K;
  if (make_map_f16(&m->map_a_hi, a_hi, (uint64_t)m->alloc_sites, K, F::BK, F::BM, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  if (make_map_f16(&m->map_a_lo, a_lo, (uint64_t)m->alloc_sites, K, F::BK, F::BM, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  if (make_map_f16(&m->map_b_hi, m->d_w4t_hi, F::N, K, F::BK, F::NH, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  if (make_map_f16(&m->map_b_lo, m->d_w4t_lo, F::N, K, F::BK, F::NH, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  if (make_map_f16(&m->map_bh_hi, m->d_w4t_hi, F::N, K, F::BK, F::NH / 2, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  if (make_map_f16(&m->map_bh_lo, m->d_w4t_lo, F::N, K, F::BK, F::NH / 2, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  CK(cudaFuncSetAttribute(tc::k_fc4_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES));
  CK(cudaFuncSetAttribute(tc::k_fc4_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES));
  {
    using C = tc::Conv3Tc;
    using S = tc::Conv3SlabRes;
    CK(cudaMalloc(&m->d_w3b_hi, (size_t)C::B_ROWS_TOTAL * C::KROW * 2 * 2));  // hi plane then lo plane
    m->d_w3b_lo = m->d_w3b_hi + (size_t)C::B_ROWS_TOTAL * C::KROW;
    if (make_slab_map<C, S>(m->d_p2, m->p2_rows, CU_TENSOR_MAP_SWIZZLE_64B, &m->map_c3slab)) return 1;
    if (make_conv_weight_maps<C>(m->d_w3b_hi, CU_TENSOR_MAP_SWIZZLE_64B, &m->map_c3b2, &m->map_c3b3, &m->map_c3b4)) return 1;
    CK(cudaFuncSetAttribute(tc::k_conv_slab<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
  }
  {
    using C = tc::Conv2Tc;
    using S = tc::Conv2Slab;
    const size_t p1_halves = (size_t)m->p1_rows * C::KROW;
    CK(cudaMalloc(&m->d_p1, p1_halves * 2 * 2));
    CK(cudaMemset(m->d_p1, 0, p1_halves * 2 * 2));  // row 29 of every site stays zero (conv2's bottom SAME pad)
    CK(cudaMalloc(&m->d_w2b_hi, (size_t)C::B_ROWS_TOTAL * C::KROW * 2 * 2));  // hi plane then lo plane
    m->d_w2b_lo = m->d_w2b_hi + (size_t)C::B_ROWS_TOTAL * C::KROW;
    if (make_slab_map<C, S>(m->d_p1, m->p1_rows, CU_TENSOR_MAP_SWIZZLE_32B, &m->map_c2slab)) return 1;
    if (make_conv_weight_maps<C>(m->d_w2b_hi, CU_TENSOR_MAP_SWIZZLE_32B, &m->map_c2b2, &m->map_c2b3, &m->map_c2b4)) return 1;
    CK(cudaFuncSetAttribute(tc::k_conv_slab<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
  }
  {
    using T = tc::TailTc;
    const size_t plane = (size_t)m->alloc_sites * T::N4;
    CK(cudaMalloc(&m->d_h4s, plane * 2 * 2));
    CK(cudaMalloc(&m->d_wtail, (size_t)T::NB * T::N4 * 2 * 2));
    if (make_map_f16(&m->map_ta_hi, m->d_h4s, (uint64_t)m->alloc_sites, T::N4, T::BK, T::BM, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (make_map_f16(&m->map_ta_lo, m->d_h4s + plane, (uint64_t)m->alloc_sites, T::N4, T::BK, T::BM, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (make_map_f16(&m->map_tb_hi, m->d_wtail, T::NB, T::N4, T::BK, T::NB, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (make_map_f16(&m->map_tb_lo, m->d_wtail + (size_t)T::NB * T::N4, T::NB, T::N4, T::BK, T::NB, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    CK(cudaFuncSetAttribute(tc::k_tail_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
  }
  m->tc_ready = true;
  m->tc_weights_dirty = true;
  return 0;
}

// v3_slim: conv1 SIMT (registers), conv2 / conv3 / FC4 on tcgen05, tail SIMT
static int tc_setup_slim(cvb_model* m) {
  if (m->tc_ready) return 0;
  using C = tc::SlimConv3TcH;
  CK(cudaMalloc(&m->d_absmax, 16));
  CK(cudaMalloc(&m->d_inv_scale, 16));
  CK(cudaMalloc(&m->d_w3b_hi, (size_t)C::B_ROWS_TOTAL * C::KROW * 2 * 2));
  m->d_w3b_lo = m->d_w3b_hi + (size_t)C::B_ROWS_TOTAL * C::KROW;
  if (make_conv_weight_maps<C>(m->d_w3b_hi, CU_TENSOR_MAP_SWIZZLE_32B, &m->map_c3b2, &m->map_c3b3, &m->map_c3b4)) return 1;
  if (make_slab_map<C, tc::SlimConv3HSlabRes>(m->d_p2, m->p2_rows, CU_TENSOR_MAP_SWIZZLE_32B, &m->map_c3slab)) return 1;
  {
    // conv2 as a dense row-shifted GEMM on p1 [site][35][32] (k_slim_c1_reg writes it), conv3 with fp16 hi / lo output planes,
    // FC4 as a split-fp16 GEMM over those planes (gemm_tc.cuh)
    using C2 = tc::SlimConv2Tc;
    using S2 = tc::SlimConv2SlabRes;
    m->p1s_rows = m->alloc_sites * C2::RPS + 256;
    const size_t p1_halves = (size_t)m->p1s_rows * C2::KROW;
    CK(cudaMalloc(&m->d_p1s, p1_halves * 2 * 2));
    CK(cudaMemset(m->d_p1s, 0, p1_halves * 2 * 2));  // rows 0 and 34 of every site stay zero (conv2's SAME padding)
    const size_t w2_halves = (size_t)C2::B_ROWS_TOTAL * C2::KROW;
    CK(cudaMalloc(&m->d_w2s, w2_halves * 2 * 2));
    CK(cudaMalloc(&m->d_w4h, (size_t)36 * 4224 * 2 * 2));
    if (make_slab_map<C2, S2>(m->d_p1s, m->p1s_rows, CU_TENSOR_MAP_SWIZZLE_64B, &m->map_s2slab)) return 1;
    const uint64_t bd[3] = {(uint64_t)C2::KROW, (uint64_t)C2::B_ROWS_TOTAL, 2};
    const uint64_t bs[2] = {(uint64_t)C2::KROW * 2, (uint64_t)C2::B_ROWS_TOTAL * C2::KROW * 2};
    const uint32_t bb[3] = {(uint32_t)C2::BK, (uint32_t)C2::NOUT, 2};
    if (make_map_nd(&m->map_s2b, m->d_w2s, 3, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    CK(cudaFuncSetAttribute(tc::k_conv_slab<C2, S2>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::k_conv_slab<tc::SlimConv3TcH, tc::SlimConv3HSlabRes>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SlimConv3HSlabRes::SMEM_BYTES));
  }
  m->tc_ready = true;
  m->tc_weights_dirty = true;
  return 0;
}

// (re)build the split fp16 copy of fc4/kernel on `st` if the fp32 master changed
static int tc_refresh_weights(cvb_model* m, cudaStream_t st) {
  if (!m->tc_weights_dirty) return 0;
  {  // host copies of the conv biases for the by-value kernel parameter (BiasParam); weights change rarely on this path
    const VarInfo* b2 = m->info("conv2/bias");
    const VarInfo* b3 = m->info("conv3/bias");
    CK(cudaMemcpyAsync(m->hb2.v, m->d_params + b2->offset, (size_t)b2->numel * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(m->hb3.v, m->d_params + b3->offset, (size_t)b3->numel * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  if (m->variant == CVB_V3_SLIM) {
    using C = tc::SlimConv3Tc;
    CK(cudaMemsetAsync(m->d_absmax + 1, 0, 4, st));
    tc::k_absmax<<<32, 256, 0, st>>>(m->var("conv3/kernel"), 5 * 4 * 16 * 32, m->d_absmax + 1);
    tc::k_prep_conv_weights<C><<<(C::B_ROWS_TOTAL * C::KROW + 255) / 256, 256, 0, st>>>(m->var("conv3/kernel"), m->d_absmax + 1,
                                                                                        m->d_w3b_hi, m->d_w3b_lo, m->d_inv_scale + 1);
    CK(cudaGetLastError());
    m->launches += 2;
    {
      using C2 = tc::SlimConv2Tc;
      CK(cudaMemsetAsync(m->d_absmax + 2, 0, 4, st));
      tc::k_absmax<<<16, 256, 0, st>>>(m->var("conv2/kernel"), 3 * 4 * 8 * 16, m->d_absmax + 2);
      tc::k_prep_conv_weights_dense<C2, 8, 16><<<(C2::B_ROWS_TOTAL * C2::KROW + 255) / 256, 256, 0, st>>>(
          m->var("conv2/kernel"), m->d_absmax + 2, m->d_w2s, m->d_w2s + (size_t)C2::B_ROWS_TOTAL * C2::KROW, m->d_inv_scale + 2);
      CK(cudaMemsetAsync(m->d_absmax, 0, 4, st));
      tc::k_absmax<<<64, 256, 0, st>>>(m->var("fc4/kernel"), (int64_t)4224 * 36, m->d_absmax);
      tc::k_prep_fc_weights<<<dim3(4224 / 32, 2), dim3(32, 8), 0, st>>>(m->var("fc4/kernel"), 4224, 36, m->d_absmax, m->d_w4h,
                                                                        m->d_w4h + (size_t)36 * 4224, m->d_inv_scale);
      CK(cudaGetLastError());
      m->launches += 4;
    }
    m->tc_weights_dirty = false;
    return 0;
  }
  using F = tc::Fc4Tc;
  const int K = 4608;
  CK(cudaMemsetAsync(m->d_absmax, 0, 4, st));
  tc::k_absmax<<<256, 256, 0, st>>>(m->var("fc4/kernel"), (int64_t)K * F::N, m->d_absmax);
  CK(cudaGetLastError());
  dim3 grid((K + 31) / 32, (F::N + 31) / 32), block(32, 8);
  tc::k_prep_fc_weights<<<grid, block, 0, st>>>(m->var("fc4/kernel"), K, F::N, m->d_absmax, m->d_w4t_hi, m->d_w4t_lo,
                                                 m->d_inv_scale);
  CK(cudaGetLastError());
  // conv3: separate |w|max and scale (d_absmax[1], d_inv_scale[1])
  CK(cudaMemsetAsync(m->d_absmax + 1, 0, 4, st));
  tc::k_absmax<<<64, 256, 0, st>>>(m->var("conv3/kernel"), 3 * 4 * 32 * 48, m->d_absmax + 1);
  CK(cudaGetLastError());
  tc::k_prep_conv_weights<tc::Conv3Tc><<<(3 * 192 * 128 + 255) / 256, 256, 0, st>>>(
      m->var("conv3/kernel"), m->d_absmax + 1, m->d_w3b_hi, m->d_w3b_lo, m->d_inv_scale + 1);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(m->d_absmax + 2, 0, 4, st));
  tc::k_absmax<<<16, 256, 0, st>>>(m->var("conv2/kernel"), 2 * 4 * 16 * 32, m->d_absmax + 2);
  CK(cudaGetLastError());
  tc::k_prep_conv_weights<tc::Conv2Tc><<<(2 * 128 * 64 + 255) / 256, 256, 0, st>>>(
      m->var("conv2/kernel"), m->d_absmax + 2, m->d_w2b_hi, m->d_w2b_lo, m->d_inv_scale + 2);
  CK(cudaGetLastError());
  {
    using T = tc::TailTc;
    CK(cudaMemsetAsync(m->d_absmax + 3, 0, 4, st));
    tc::k_absmax<<<64, 256, 0, st>>>(m->var("fc5/kernel"), (int64_t)T::N4 * T::N5, m->d_absmax + 3);
    tc::k_absmax<<<4, 256, 0, st>>>(m->var("YBaseChangeSigmoid/kernel"), (int64_t)T::N4 * 4, m->d_absmax + 3);
    tc::k_prep_tail_weights<<<(T::NB * T::N4 + 255) / 256, 256, 0, st>>>(m->var("fc5/kernel"), m->var("YBaseChangeSigmoid/kernel"),
                                                                         m->d_absmax + 3, m->d_wtail,
                                                                         m->d_wtail + (size_t)T::NB * T::N4, m->d_inv_scale + 3);
    CK(cudaGetLastError());
  }
  m->launches += 9;
  m->tc_weights_dirty = false;
  return 0;
}

extern "C" int cvb_set_step(cvb_model* m, int64_t t) { if (!m) return fail("NULL model"); m->step = t; return 0; }
extern "C" int cvb_get_step(const cvb_model* m, int64_t* t) { if (!m || !t) return fail("NULL argument"); *t = m->step; return 0; }
extern "C" int cvb_set_compute_mode(cvb_model* m, int mode) {
  if (!m) return fail("NULL model");
  if (mode != CVB_COMPUTE_FP32 && mode != CVB_COMPUTE_FP16X3 && mode != CVB_COMPUTE_FP16)
    return fail("cvb_set_compute_mode: unknown mode %d", mode);
  if (mode == CVB_COMPUTE_FP16 && m->variant != CVB_V3_SLIM)
    return fail("cvb_set_compute_mode: plain fp16 (BASELINE configs[2]) is implemented for v3_slim");
  CK(cudaSetDevice(m->device));
  if (mode != CVB_COMPUTE_FP32 && (m->variant == CVB_V3 ? tc_setup(m) : tc_setup_slim(m))) return 1;
  if ((mode == CVB_COMPUTE_FP32) != (m->compute_mode == CVB_COMPUTE_FP32)) {
    // p2's zero padding rows sit at different byte offsets in the fp32 and the fp16 hi/lo layouts
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(m->d_p2, 0, m->p2_bytes));
  }
  m->compute_mode = mode;
  m->CHUNK = (int64_t)m->num_sms * (m->variant == CVB_V3 ? (mode == CVB_COMPUTE_FP32 ? 96 : 128) : 224);
  return 0;
}
extern "C" int cvb_set_train_mode(cvb_model* m, int mode) {
  if (!m) return fail("NULL model");
  if (mode != CVB_TRAIN_FP32 && mode != CVB_TRAIN_BF16X3 && mode != CVB_TRAIN_BF16)
    return fail("cvb_set_train_mode: unknown mode %d", mode);
  m->train_mode = mode;
  return 0;
}
extern "C" int cvb_set_dropout_fc5(cvb_model* m, float rate) {
  if (!m) return fail("NULL model");
  if (!(rate >= 0.f && rate < 1.f)) return fail("cvb_set_dropout_fc5: rate %g outside [0,1)", rate);
  m->drop5 = rate;
  return 0;
}
extern "C" int64_t cvb_kernel_launches(const cvb_model* m) { return m ? m->launches : 0; }

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
// cudaLaunchKernelEx with the two launch attributes this library uses: a cluster shape and programmatic dependent launch
// (the kernel may be scheduled while its predecessor in the stream is still in its last wave; it calls pdl_wait() before it
// touches the predecessor's output, common.cuh)
template <class... P, class... A>
static cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, int cluster_y,
                            bool pdl, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (cluster_x * cluster_y > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cluster_x; at[na].val.clusterDim.y = cluster_y; at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
static bool use_pdl(const cvb_model* m, int which = 0) {
  static const bool on = !(getenv("CVB_PDL") && getenv("CVB_PDL")[0] == '0');
  static const int off_mask = getenv("CVB_PDL_OFF") ? atoi(getenv("CVB_PDL_OFF")) : 0;  // diagnosis: 1 conv2, 2 conv3, 4 fc4, 8 after fc4
  return on && !(off_mask & which) && !m->profiling;  // (event records between the kernels would serialise them anyway)
}

template <class K>
static cudaError_t set_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

static HeadPtrs head_ptrs(const cvb_model* m) {
  HeadPtrs h;
  h.w5 = m->var("fc5/kernel"); h.b5 = m->var("fc5/bias");
  h.wb = m->var("YBaseChangeSigmoid/kernel"); h.bb = m->var("YBaseChangeSigmoid/bias");
  h.wz = m->var("YZygosityFC/kernel"); h.bz = m->var("YZygosityFC/bias");
  h.wt = m->var("YVarTypeFC/kernel"); h.bt = m->var("YVarTypeFC/bias");
  h.wl = m->var("YIndelLengthFC/kernel"); h.bl = m->var("YIndelLengthFC/bias");
  return h;
}

// one persistent launch of k_conv_slab<T, S> over n sites (inference: bias by value where S::BC says so)
template <class T, class S>
static int launch_conv_slab(cvb_model* m, int64_t n, cudaStream_t st, bool pdl, const CUtensorMap& slab, const CUtensorMap& b2,
                            const CUtensorMap& b3, const CUtensorMap& b4, const float* bias, const tc::BiasParam& bc,
                            const float* inv_scale, __half* out_hi, __half* out_lo, int flags = 0) {
  static const int ablate = getenv("CVB_ABLATE") ? atoi(getenv("CVB_ABLATE")) : 0;  // timing experiments only (tools/ablate.py)
  const int64_t tiles = (n * T::RPS + S::TILE_STEP - 1) / S::TILE_STEP;
  CK(launch_k(tc::k_conv_slab<T, S>, dim3((unsigned)std::min<int64_t>(tiles, m->num_sms)), dim3(S::THREADS), S::SMEM_BYTES, st, 1, 1,
              pdl, slab, b2, b3, b4, n, bias, inv_scale, out_hi, out_lo, ablate | flags, bc));
  CK(cudaGetLastError());
  return 0;
}

static int prof_mark(cvb_model* m, cudaStream_t st) {
  if (!m->profiling) return 0;
  if (m->prof_used == m->prof_events.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    m->prof_events.push_back(e);
  }
  CK(cudaEventRecord(m->prof_events[m->prof_used++], st));
  return 0;
}

struct GemmExtra {  // batched / split-K launches, see tc::GemmArgs
  int batches = 1, a_batch_rows = 0, a_kshift0 = 0, a_kshift_per_batch = 0;
  int64_t c_batch_stride = 0;
  int kslices = 1;
  int64_t c_slice_stride = 0;  // > 0: K slice z stores to C + z * c_slice_stride (no atomics)
  int f16 = 0, terms = 0;  // fp16 operands; terms 0 = what the training mode says (3, or 1 for CVB_TRAIN_BF16)
  const float* inv_scale = nullptr;
  bool pdl = false;
};
template <int BN, bool CHUNKED, int EPI, bool MN = false>
static int launch_gemm_tc(cvb_model* m, const uint16_t* a, int64_t a_plane, int64_t lda, const uint16_t* b, int64_t b_plane,
                          int64_t ldb, int M, int N, int K, float* C, int64_t ldc, const float* bias, cudaStream_t st,
                          const GemmExtra& ex = GemmExtra());

// narrow feed -> fp32 scratch (only for the front kernels that do not widen on their own)
static int widen_chunk(cvb_model* m, const void* xin, int kind, int64_t n, cudaStream_t st) {
  if (!m->d_xw) CK(cudaMalloc(&m->d_xw, (size_t)m->alloc_sites * 528 * 4));
  const int64_t npos = n * 132;
  const int g = (int)std::min<int64_t>((npos + 255) / 256, (int64_t)m->num_sms * 16);
  float4* o = reinterpret_cast<float4*>(m->d_xw);
  if (kind == X_F16) k_widen<X_F16><<<g, 256, 0, st>>>(xin, o, npos);
  else if (kind == X_I16) k_widen<X_I16><<<g, 256, 0, st>>>(xin, o, npos);
  else k_widen<X_U8><<<g, 256, 0, st>>>(xin, o, npos);
  CK(cudaGetLastError());
  m->launches += 1;
  return 0;
}

// one chunk (n <= CHUNK) of the forward pass on `st`; xin holds elements of `kind` (X_F32 .. X_U8)
static int forward_chunk(cvb_model* m, const void* xin, int kind, int64_t n, OutDst out16, float* logits16, cudaStream_t st) {
  if (n <= 0) return 0;
  const int sms = m->num_sms;
  const bool tensor = m->compute_mode != CVB_COMPUTE_FP32;
  if (tensor && tc_refresh_weights(m, st)) return 1;
  // k_v3_c1_reg<KIND> / k_slim_c1_reg<KIND> (tensor path) widen a narrow feed on their own
  if (kind != X_F32 && !tensor) {
    if (widen_chunk(m, xin, kind, n, st)) return 1;
    xin = m->d_xw;
    kind = X_F32;
  }
  const float* x = static_cast<const float*>(xin);
  if (prof_mark(m, st)) return 1;
  if (m->variant == CVB_V3 && tensor) {
    // ---- v3 on the tensor path: conv1 (SIMT, registers) -> conv2 -> conv3 (row-shifted implicit GEMMs) -> FC4 -> fused tail
    {
      using T = tc::Conv2Tc;
      const size_t p1_halves = (size_t)m->p1_rows * T::KROW;
      const unsigned g1 = (unsigned)((n * 16 + 127) / 128);
      const float *w1 = m->var("conv1/kernel"), *b1 = m->var("conv1/bias");
      __half *phi = m->d_p1, *plo = m->d_p1 + p1_halves;
      if (kind == X_F32) k_v3_c1_reg<X_F32><<<g1, 128, 0, st>>>(xin, n, w1, b1, phi, plo);
      else if (kind == X_F16) k_v3_c1_reg<X_F16><<<g1, 128, 0, st>>>(xin, n, w1, b1, phi, plo);
      else if (kind == X_I16) k_v3_c1_reg<X_I16><<<g1, 128, 0, st>>>(xin, n, w1, b1, phi, plo);
      else k_v3_c1_reg<X_U8><<<g1, 128, 0, st>>>(xin, n, w1, b1, phi, plo);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
      __half* p2_hi = reinterpret_cast<__half*>(m->d_p2);
      if (launch_conv_slab<T, tc::Conv2Slab>(m, n, st, use_pdl(m, 1), m->map_c2slab, m->map_c2b2, m->map_c2b3, m->map_c2b4,
                                             m->var("conv2/bias"), m->hb2, m->d_inv_scale + 2, p2_hi, p2_hi + m->p2_rows * 128))
        return 1;
      if (prof_mark(m, st)) return 1;
    }
    {
      using T = tc::Conv3Tc;
      __half* p3_hi = reinterpret_cast<__half*>(m->d_p3);
      if (launch_conv_slab<T, tc::Conv3SlabRes>(m, n, st, use_pdl(m, 2), m->map_c3slab, m->map_c3b2, m->map_c3b3, m->map_c3b4,
                                                m->var("conv3/bias"), m->hb3, m->d_inv_scale + 1, p3_hi,
                                                p3_hi + m->alloc_sites * 4608))
        return 1;
      if (prof_mark(m, st)) return 1;
    }
    {
      using F = tc::Fc4Tc;
      const unsigned tiles4 = (unsigned)((n + F::BM - 1) / F::BM);
      __half* h4hi = m->d_h4s;  // fp16 hi / lo planes of h4 = the A operand of the fused tail
      __half* h4lo = m->d_h4s + (size_t)m->alloc_sites * 336;
      // small batches: split the 9 K-chunks over gridDim.z so that ~every SM streams a slice of W4 (k_fc4_tc, `ws`)
      unsigned ksplit = 1;
      const unsigned ctas = 2 * ((tiles4 + 1) & ~1u);
      for (unsigned k : {9u, 3u})
        if ((int64_t)n <= kFc4SplitSites && ctas * k <= 2 * (unsigned)sms) { ksplit = k; break; }
      float* ws = nullptr;
      const int64_t ws_plane = (int64_t)kFc4SplitSites * 2 * F::NH;
      if (ksplit > 1) {
        if (!m->d_fc4ws) CK(cudaMalloc(&m->d_fc4ws, (size_t)(4608 / F::KCH) * ws_plane * 4));
        ws = m->d_fc4ws;
      }
      if (tiles4 >= 2) {  // pairs of site tiles share every weight box (2-CTA multicast); a padding tile loads zeros and stores nothing
        const dim3 g4(2, (tiles4 + 1) & ~1u, ksplit);
        CK(launch_k(tc::k_fc4_tc<2>, g4, dim3(F::THREADS), F::SMEM_BYTES, st, 1, 2, use_pdl(m, 4), m->map_a_hi, m->map_a_lo,
                    m->map_bh_hi, m->map_bh_lo, n, 4608, m->var("fc4/bias"), (const float*)m->d_inv_scale, m->d_h4, h4hi, h4lo, ws,
                    ws_plane));
      } else {
        tc::k_fc4_tc<1><<<dim3(2, tiles4, ksplit), F::THREADS, F::SMEM_BYTES, st>>>(m->map_a_hi, m->map_a_lo, m->map_b_hi, m->map_b_lo, n,
                                                                                    4608, m->var("fc4/bias"), m->d_inv_scale, m->d_h4,
                                                                                    h4hi, h4lo, ws, ws_plane);
      }
      CK(cudaGetLastError());
      if (ws) {
        CK(launch_k(tc::k_fc4_reduce, dim3((unsigned)std::min<int64_t>((n * 84 + 127) / 128, (int64_t)sms * 8)), dim3(128), 0, st, 1, 1,
                    use_pdl(m), (const float*)ws, ws_plane, 4608 / F::KCH, n, m->var("fc4/bias"), (const float*)m->d_inv_scale, m->d_h4,
                    h4hi, h4lo));
        CK(cudaGetLastError());
        m->launches += 1;
      }
      if (prof_mark(m, st)) return 1;
    }
    {
      using T = tc::TailTc;
      tc::TailHeads th{m->var("fc5/bias"), m->var("YBaseChangeSigmoid/bias"), m->var("YZygosityFC/kernel"), m->var("YZygosityFC/bias"),
                       m->var("YVarTypeFC/kernel"), m->var("YVarTypeFC/bias"), m->var("YIndelLengthFC/kernel"),
                       m->var("YIndelLengthFC/bias")};
      CK(launch_k(tc::k_tail_tc, dim3((unsigned)((n + T::BM - 1) / T::BM)), dim3(T::THREADS), T::SMEM_BYTES, st, 1, 1, use_pdl(m, 8),
                  m->map_ta_hi, m->map_ta_lo, m->map_tb_hi, m->map_tb_lo, n, th, (const float*)(m->d_inv_scale + 3), out16, logits16));
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    m->launches += 5;  // c1, conv2, conv3, FC4, tail
  } else if (m->variant == CVB_V3) {
    // ---- v3, fp32 SIMT kernels (CVB_COMPUTE_FP32)
    {
      using F = FrontV3<4>;
      auto k = k_v3_front<4>;
      CK(set_smem(k, F::SMEM_BYTES));
      const int grid = (int)std::min<int64_t>((n + 3) / 4, 2 * sms);
      k<<<grid, 256, F::SMEM_BYTES, st>>>(x, n, m->var("conv1/kernel"), m->var("conv1/bias"), m->var("conv2/kernel"),
                                          m->var("conv2/bias"), m->d_p2);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
      if (prof_mark(m, st)) return 1;  // (no separate conv2 kernel)
    }
    {
      using C = ConvCfg<32, 48, 3, 26, 3, 8, 8>;
      using L = ConvLayerSmem<C, 3>;
      auto k = k_conv_layer<C, 3, 256>;
      CK(set_smem(k, L::SMEM_BYTES));
      const int grid = (int)std::min<int64_t>((n + C::S - 1) / C::S, sms);
      k<<<grid, 256, L::SMEM_BYTES, st>>>(m->d_p2, n, m->var("conv3/kernel"), m->var("conv3/bias"), m->d_p3);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    {
      using F = FcCfg<336, 21, 16, 12, 8>;
      auto k = k_fc4<F>;
      CK(set_smem(k, F::SMEM_BYTES));
      int grid = (int)((n + F::M - 1) / F::M);
      k<<<grid, 256, F::SMEM_BYTES, st>>>(m->d_p3, n, 4608, m->var("fc4/kernel"), m->var("fc4/bias"), m->d_h4, 336, 336);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    {
      using F5 = FcCfg<168, 21, 8, 12, 8>;  // FC5: h5 = SELU(h4 @ W5 + b5), same SGEMM as the fp32 FC4
      auto k = k_fc4<F5>;
      CK(set_smem(k, F5::SMEM_BYTES));
      int grid = (int)((n + F5::M - 1) / F5::M);
      k<<<grid, 256, F5::SMEM_BYTES, st>>>(m->d_h4, n, 336, m->var("fc5/kernel"), m->var("fc5/bias"), m->d_h5, 168, 168);
      CK(cudaGetLastError());
      k_heads<336, 168><<<(int)((n + 15) / 16), 256, 0, st>>>(m->d_h4, m->d_h5, n, head_ptrs(m), out16, logits16);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    m->launches += 5;  // front, conv3, FC4, FC5, heads
  } else if (tensor) {
    // ---- v3_slim on the tensor path: conv1 (SIMT, registers) -> conv2 (dense row-shifted GEMM) -> conv3 -> FC4 (GEMM) -> tail
    const bool hi_only = m->compute_mode == CVB_COMPUTE_FP16;  // plain fp16: one MMA term, hi planes only
    using C2 = tc::SlimConv2Tc;
    using C3 = tc::SlimConv3TcH;
    {
      const unsigned g1 = (unsigned)((n * 8 + 63) / 64);
      const float *w1 = m->var("conv1/kernel"), *b1 = m->var("conv1/bias");
      __half *phi = m->d_p1s, *plo = m->d_p1s + (size_t)m->p1s_rows * C2::KROW;
#define CVB_SLIM_C1(KIND)                                                                                  \
  do {                                                                                                     \
    if (hi_only) k_slim_c1_reg<KIND, false><<<g1, 64, 0, st>>>(xin, n, w1, b1, phi, plo);                  \
    else k_slim_c1_reg<KIND, true><<<g1, 64, 0, st>>>(xin, n, w1, b1, phi, plo);                           \
  } while (0)
      if (kind == X_F32) CVB_SLIM_C1(X_F32);
      else if (kind == X_F16) CVB_SLIM_C1(X_F16);
      else if (kind == X_I16) CVB_SLIM_C1(X_I16);
      else CVB_SLIM_C1(X_U8);
#undef CVB_SLIM_C1
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    const int hi_flag = hi_only ? 64 : 0;  // k_conv_slab: only the hi x hi term, no lo plane written
    {
      __half* p2_hi = reinterpret_cast<__half*>(m->d_p2);
      if (launch_conv_slab<C2, tc::SlimConv2SlabRes>(m, n, st, use_pdl(m, 1), m->map_s2slab, m->map_s2b, m->map_s2b, m->map_s2b,
                                                     m->var("conv2/bias"), m->hb2, m->d_inv_scale + 2, p2_hi,
                                                     p2_hi + m->p2_rows * 64, hi_flag))
        return 1;
      if (prof_mark(m, st)) return 1;
    }
    {
      __half* p3_hi = reinterpret_cast<__half*>(m->d_p3);
      if (launch_conv_slab<C3, tc::SlimConv3HSlabRes>(m, n, st, use_pdl(m, 2), m->map_c3slab, m->map_c3b2, m->map_c3b3, m->map_c3b4,
                                                      m->var("conv3/bias"), m->hb3, m->d_inv_scale + 1, p3_hi,
                                                      p3_hi + m->alloc_sites * 4224, hi_flag))
        return 1;
      if (prof_mark(m, st)) return 1;
    }
    {
      GemmExtra ex;
      ex.f16 = 1;
      ex.terms = hi_only ? 1 : 3;
      ex.inv_scale = m->d_inv_scale;
      // (launched the ordinary way: as a programmatic dependent of conv3 this 8-CTA GEMM took 37 us longer whenever
      //  conv3 had 2..7 CTAs with a single tile -- 999 <= n <= 1016, the reference's batch size among them -- and saved 4 us
      //  otherwise; measured with tools/small_n_probe.py)
      ex.pdl = false;
      if (launch_gemm_tc<48, true, tc::GEMM_EPI_BIAS_SELU>(m, reinterpret_cast<const uint16_t*>(m->d_p3), m->alloc_sites * 4224, 4224,
                                                           reinterpret_cast<const uint16_t*>(m->d_w4h), 36 * 4224, 4224, (int)n, 36, 4224,
                                                           m->d_h4, 36, m->var("fc4/bias"), st, ex))
        return 1;
      if (prof_mark(m, st)) return 1;
    }
    k_tail_site<36, 18><<<(int)((n + 127) / 128), 128, 0, st>>>(m->d_h4, n, head_ptrs(m), out16, logits16);
    CK(cudaGetLastError());
    if (prof_mark(m, st)) return 1;
    m->launches += 5;
  } else {
    // ---- v3_slim, fp32 SIMT kernels (CVB_COMPUTE_FP32)
    {
      using F = FrontSlim<6>;
      int64_t tiles = (n + 5) / 6;
      int grid = (int)std::min<int64_t>(tiles, 4 * sms);
      auto k = k_slim_front<6>;
      CK(set_smem(k, F::SMEM_BYTES));
      k<<<grid, 256, F::SMEM_BYTES, st>>>(x, n, m->var("conv1/kernel"), m->var("conv1/bias"), m->var("conv2/kernel"),
                                          m->var("conv2/bias"), m->d_p2);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
      if (prof_mark(m, st)) return 1;  // (no separate conv2 kernel)
    }
    {
      using C = ConvCfg<16, 32, 5, 33, 3, 8, 8>;
      using L = ConvLayerSmem<C, 1>;
      auto k = k_conv_layer<C, 1, 256>;
      CK(set_smem(k, L::SMEM_BYTES));
      int64_t tiles = (n + C::S - 1) / C::S;
      int grid = (int)std::min<int64_t>(tiles, sms);
      k<<<grid, 256, L::SMEM_BYTES, st>>>(m->d_p2, n, m->var("conv3/kernel"), m->var("conv3/bias"), m->d_p3);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    {
      using F = FcCfg<36, 9, 4, 28, 8>;
      auto k = k_fc4<F>;
      CK(set_smem(k, F::SMEM_BYTES));
      int grid = (int)((n + F::M - 1) / F::M);
      k<<<grid, 256, F::SMEM_BYTES, st>>>(m->d_p3, n, 4224, m->var("fc4/kernel"), m->var("fc4/bias"), m->d_h4, 36, 36);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    {
      k_tail_site<36, 18><<<(int)((n + 127) / 128), 128, 0, st>>>(m->d_h4, n, head_ptrs(m), out16, logits16);
      CK(cudaGetLastError());
      if (prof_mark(m, st)) return 1;
    }
    m->launches += 4;
  }
  return 0;
}

static const int kXBytes[4] = {4, 2, 2, 1};  // element size of CVB_X_F32 / F16 / I16 / U8

extern "C" int cvb_predict_device_x(cvb_model* m, const void* x, int x_kind, int64_t n, float* out16, float* logits16, void* stream) {
  if (!m) return fail("cvb_predict_device: NULL model");
  if (n < 0) return fail("cvb_predict_device: negative n");
  if (x_kind < 0 || x_kind > 3) return fail("cvb_predict_device: unknown element kind %d", x_kind);
  if (n == 0) return 0;
  if (!x || !out16) return fail("cvb_predict_device: NULL buffer");
  if ((uintptr_t)x & 15) return fail("cvb_predict_device: x must be 16-byte aligned");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t CHUNK = m->CHUNK;
  const char* xb = static_cast<const char*>(x);
  for (int64_t s = 0; s < n; s += CHUNK) {
    int64_t c = std::min<int64_t>(CHUNK, n - s);
    if (forward_chunk(m, xb + (size_t)s * 528 * kXBytes[x_kind], x_kind, c, out_interleaved(out16 + s * 16),
                      logits16 ? logits16 + s * 16 : nullptr, st))
      return 1;
  }
  return 0;
}
extern "C" int cvb_predict_device(cvb_model* m, const float* x, int64_t n, float* out16, float* logits16, void* stream) {
  return cvb_predict_device_x(m, x, X_F32, n, out16, logits16, stream);
}

static int ensure_host_slots(cvb_model* m) {
  if (m->d_x[0]) return 0;
  const int64_t CHUNK = m->alloc_sites;  // large enough for every compute mode's chunk
  for (int i = 0; i < cvb_model::NSLOT; ++i) {
    CK(cudaMalloc(&m->d_x[i], (size_t)CHUNK * 528 * 4));
    CK(cudaMalloc(&m->d_x16[i], (size_t)CHUNK * 528 * 2));
    CK(cudaMalloc(&m->d_out[i], (size_t)CHUNK * 16 * 4));
    CK(cudaMalloc(&m->d_lg[i], (size_t)CHUNK * 16 * 4));
    CK(cudaMallocHost(&m->h_x[i], (size_t)CHUNK * 528 * 4));
    CK(cudaMallocHost(&m->h_out[i], (size_t)CHUNK * 16 * 4));
    CK(cudaMallocHost(&m->h_lg[i], (size_t)CHUNK * 16 * 4));
  }
  return 0;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Large fresh output arrays cost one page fault per 4 KB on first touch; ask for 2 MB pages.
static void advise_hugepages(void* p, size_t bytes) {
  const uintptr_t a = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
  const uintptr_t e = ((uintptr_t)p + bytes) & ~(uintptr_t)((2u << 20) - 1);
  if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
}

static int predict_host_impl(cvb_model* m, const void* xv, int kind, int64_t n, float* base, float* zygosity, float* var_type,
                             float* indel_length, float* logits16);

extern "C" int cvb_predict_host(cvb_model* m, const float* x, int64_t n, float* base, float* zygosity, float* var_type,
                                float* indel_length, float* logits16) {
  return predict_host_impl(m, x, X_F32, n, base, zygosity, var_type, indel_length, logits16);
}
extern "C" int cvb_predict_host_f16(cvb_model* m, const uint16_t* x, int64_t n, float* base, float* zygosity, float* var_type,
                                    float* indel_length, float* logits16) {
  return predict_host_impl(m, x, X_F16, n, base, zygosity, var_type, indel_length, logits16);
}
extern "C" int cvb_predict_host_counts_i16(cvb_model* m, const int16_t* counts, int64_t n, float* base, float* zygosity,
                                           float* var_type, float* indel_length, float* logits16) {
  return predict_host_impl(m, counts, X_I16, n, base, zygosity, var_type, indel_length, logits16);
}
extern "C" int cvb_predict_host_counts_u8(cvb_model* m, const uint8_t* counts, int64_t n, float* base, float* zygosity,
                                          float* var_type, float* indel_length, float* logits16) {
  return predict_host_impl(m, counts, X_U8, n, base, zygosity, var_type, indel_length, logits16);
}

// The device writes a chunk's results as four per-head blocks [cap*4 | cap*2 | cap*4 | cap*6] floats (OutDst), so the way
// back is four plain copies per chunk: straight into the caller's arrays when those are pinned, through the pinned staging
// slot and four memcpy otherwise -- no per-site de-interleave on the host.
static const int kHeadW[4] = {4, 2, 4, 6};

static int predict_host_impl(cvb_model* m, const void* xv, int kind, int64_t n, float* base, float* zygosity, float* var_type,
                             float* indel_length, float* logits16) {
  const char* x = static_cast<const char*>(xv);
  const size_t esz = (size_t)kXBytes[kind];
  if (!m) return fail("cvb_predict_host: NULL model");
  if (n < 0) return fail("cvb_predict_host: negative n");
  if (n == 0) return 0;
  if (!x || !base || !zygosity || !var_type || !indel_length) return fail("cvb_predict_host: NULL buffer");
  if (m->tickets_out()) return fail("cvb_predict_host: %d submitted batch(es) not collected yet (cvb_predict_collect)", m->tickets_out());
  CK(cudaSetDevice(m->device));
  if (ensure_host_slots(m)) return 1;
  if (n >= (1 << 18)) {
    advise_hugepages(base, (size_t)n * 16); advise_hugepages(zygosity, (size_t)n * 8);
    advise_hugepages(var_type, (size_t)n * 16); advise_hugepages(indel_length, (size_t)n * 24);
    if (logits16) advise_hugepages(logits16, (size_t)n * 64);
  }
  float* const heads[4] = {base, zygosity, var_type, indel_length};
  const bool pinned_in = is_pinned(x);
  // (a caller that pins its outputs pins its input too: skipping five pointer queries is worth 5-8 us on a 1,000-site call)
  const bool pinned_out = pinned_in && is_pinned(base) && is_pinned(zygosity) && is_pinned(var_type) && is_pinned(indel_length) &&
                          (!logits16 || is_pinned(logits16));
  const int64_t CHUNK = m->CHUNK, cap = m->alloc_sites;
  const int64_t nchunks = (n + CHUNK - 1) / CHUNK;
  if (nchunks == 1) {
    // One chunk -- every call of the reference's drivers (predictBatchSize = 1000, callVar.py:184): nothing to pipeline, so
    // the copy in, the kernels and the copy out go down ONE stream with no events in between, the four head blocks are packed
    // back to back ([4 | 2 | 4 | 6] x np floats, np = n rounded up to 4 for the vector stores) and come back in one copy.
    cudaStream_t st = m->s_comp;
    void* dx = kind == X_F32 ? (void*)m->d_x[0] : (void*)m->d_x16[0];
    const size_t bytes = (size_t)n * 528 * esz;
    if (pinned_in) {
      CK(cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, st));
    } else {
      // pageable input: staged through the pinned slot in four pieces, each on its way to the device while the next is copied
      char* hs = reinterpret_cast<char*>(m->h_x[0]);
      const size_t piece = ((bytes / 4) + 4095) & ~(size_t)4095;
      for (size_t o = 0; o < bytes; o += piece) {
        const size_t len = std::min(piece, bytes - o);
        memcpy(hs + o, x + o, len);
        CK(cudaMemcpyAsync(static_cast<char*>(dx) + o, hs + o, len, cudaMemcpyHostToDevice, st));
      }
    }
    const int64_t np = (n + 3) & ~(int64_t)3;
    float* o = m->d_out[0];
    const OutDst dst{nullptr, o, o + np * 4, o + np * 6, o + np * 10};
    if (forward_chunk(m, dx, kind, n, dst, logits16 ? m->d_lg[0] : nullptr, st)) return 1;
    if (pinned_out) {
      int64_t off = 0;
      for (int h = 0; h < 4; off += np * kHeadW[h], ++h)
        CK(cudaMemcpyAsync(heads[h], o + off, (size_t)n * kHeadW[h] * 4, cudaMemcpyDeviceToHost, st));
      if (logits16) CK(cudaMemcpyAsync(logits16, m->d_lg[0], (size_t)n * 64, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      return 0;
    }
    CK(cudaMemcpyAsync(m->h_out[0], o, (size_t)np * 64, cudaMemcpyDeviceToHost, st));
    if (logits16) CK(cudaMemcpyAsync(m->h_lg[0], m->d_lg[0], (size_t)n * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int64_t off = 0;
    for (int h = 0; h < 4; off += np * kHeadW[h], ++h) memcpy(heads[h], m->h_out[0] + off, (size_t)n * kHeadW[h] * 4);
    if (logits16) memcpy(logits16, m->h_lg[0], (size_t)n * 64);
    return 0;
  }
  // software pipeline over chunks: the host enqueues H2D(c), kernels(c), D2H(c) and only then collects chunk c-LAG, so
  // the copy engines always have LAG chunks of work queued ahead of the host
  constexpr int NS = cvb_model::NSLOT, LAG = NS - 1;
  for (int64_t c = 0; c < nchunks + LAG; ++c) {
    if (c < nchunks) {
      const int sl = (int)(c % NS);
      const int64_t s0 = c * CHUNK, cn = std::min<int64_t>(CHUNK, n - s0);
      const char* src = x + (size_t)s0 * 528 * esz;
      if (!pinned_in) {
        if (c >= NS) CK(cudaEventSynchronize(m->e_h2d[sl]));  // staging buffer free again
        memcpy(m->h_x[sl], src, (size_t)cn * 528 * esz);
        src = reinterpret_cast<const char*>(m->h_x[sl]);
      }
      void* dx = kind == X_F32 ? (void*)m->d_x[sl] : (void*)m->d_x16[sl];
      if (c >= NS) CK(cudaStreamWaitEvent(m->s_h2d, m->e_comp[sl], 0));  // the slot's input buffer has been consumed
      CK(cudaMemcpyAsync(dx, src, (size_t)cn * 528 * esz, cudaMemcpyHostToDevice, m->s_h2d));
      CK(cudaEventRecord(m->e_h2d[sl], m->s_h2d));
      CK(cudaStreamWaitEvent(m->s_comp, m->e_h2d[sl], 0));
      if (c >= NS) CK(cudaStreamWaitEvent(m->s_comp, m->e_d2h[sl], 0));  // d_out[sl] drained
      float* o = m->d_out[sl];
      const OutDst dst{nullptr, o, o + cap * 4, o + cap * 6, o + cap * 10};
      if (forward_chunk(m, dx, kind, cn, dst, logits16 ? m->d_lg[sl] : nullptr, m->s_comp)) return 1;
      CK(cudaEventRecord(m->e_comp[sl], m->s_comp));
      CK(cudaStreamWaitEvent(m->s_d2h, m->e_comp[sl], 0));
      int64_t off = 0;
      for (int h = 0; h < 4; off += cap * kHeadW[h], ++h) {
        float* to = pinned_out ? heads[h] + s0 * kHeadW[h] : m->h_out[sl] + off;
        CK(cudaMemcpyAsync(to, o + off, (size_t)cn * kHeadW[h] * 4, cudaMemcpyDeviceToHost, m->s_d2h));
      }
      if (logits16)
        CK(cudaMemcpyAsync(pinned_out ? logits16 + s0 * 16 : m->h_lg[sl], m->d_lg[sl], (size_t)cn * 64, cudaMemcpyDeviceToHost, m->s_d2h));
      CK(cudaEventRecord(m->e_d2h[sl], m->s_d2h));
    }
    if (c >= LAG) {
      const int64_t p = c - LAG;
      const int sl = (int)(p % NS);
      const int64_t s0 = p * CHUNK, cn = std::min<int64_t>(CHUNK, n - s0);
      CK(cudaEventSynchronize(m->e_d2h[sl]));
      if (!pinned_out) {
        int64_t off = 0;
        for (int h = 0; h < 4; off += cap * kHeadW[h], ++h)
          memcpy(heads[h] + s0 * kHeadW[h], m->h_out[sl] + off, (size_t)cn * kHeadW[h] * 4);
        if (logits16) memcpy(logits16 + s0 * 16, m->h_lg[sl], (size_t)cn * 64);
      }
    }
  }
  return 0;
}

// ---- pipelined small batches: submit k+1 .. k+3 while k is on the device ---------------------------------------------------
// A synchronous 1,000-site call (the reference's predictBatchSize) spends ~50 us staging 528 KB of pageable counts, ~73 us in
// the five kernels and ~20 us coming back, one after the other.  With up to NSLOT batches in flight the staging copy of the
// next batch and the copy-out of the previous one run on the host while the device works: the call rate becomes
// max(host work, device chain) instead of their sum.  Each ticket owns one slot (device input / output buffers, pinned
// staging buffers, three events); the kernels of consecutive tickets are serialised on s_comp, so the shared intermediates
// (p1 .. h4) are safe.
extern "C" int cvb_predict_submit(cvb_model* m, const void* xv, int x_kind, int64_t n, int want_logits, int* ticket) {
  if (!m || !ticket) return fail("cvb_predict_submit: NULL argument");
  if (x_kind < 0 || x_kind > 3) return fail("cvb_predict_submit: unknown element kind %d", x_kind);
  if (n < 0) return fail("cvb_predict_submit: negative n");
  if (n > 0 && !xv) return fail("cvb_predict_submit: NULL buffer");
  if (n > m->CHUNK)
    return fail("cvb_predict_submit: %lld sites exceed one device pass (%lld); use cvb_predict_host, which pipelines its chunks",
                (long long)n, (long long)m->CHUNK);
  const int sl = m->next_slot;
  cvb_model::Ticket& t = m->tickets[sl];
  if (t.busy) return fail("cvb_predict_submit: all %d slots are in flight; collect the oldest ticket first", cvb_model::NSLOT);
  CK(cudaSetDevice(m->device));
  if (ensure_host_slots(m)) return 1;
  const char* x = static_cast<const char*>(xv);
  const size_t bytes = (size_t)n * 528 * kXBytes[x_kind];
  if (n > 0) {
    void* dx = x_kind == X_F32 ? (void*)m->d_x[sl] : (void*)m->d_x16[sl];
    // always staged (the caller's buffer is free again when this returns), in four pieces: each is on its way to the device
    // while the next is copied
    char* hs = reinterpret_cast<char*>(m->h_x[sl]);
    const size_t piece = ((bytes / 4) + 4095) & ~(size_t)4095;
    for (size_t o = 0; o < bytes; o += piece) {
      const size_t len = std::min(piece, bytes - o);
      memcpy(hs + o, x + o, len);
      CK(cudaMemcpyAsync(static_cast<char*>(dx) + o, hs + o, len, cudaMemcpyHostToDevice, m->s_h2d));
    }
    CK(cudaEventRecord(m->e_h2d[sl], m->s_h2d));
    CK(cudaStreamWaitEvent(m->s_comp, m->e_h2d[sl], 0));
    const int64_t np = (n + 3) & ~(int64_t)3;
    float* o = m->d_out[sl];
    const OutDst dst{nullptr, o, o + np * 4, o + np * 6, o + np * 10};
    if (forward_chunk(m, dx, x_kind, n, dst, want_logits ? m->d_lg[sl] : nullptr, m->s_comp)) return 1;
    CK(cudaEventRecord(m->e_comp[sl], m->s_comp));
    CK(cudaStreamWaitEvent(m->s_d2h, m->e_comp[sl], 0));
    CK(cudaMemcpyAsync(m->h_out[sl], o, (size_t)np * 64, cudaMemcpyDeviceToHost, m->s_d2h));
    if (want_logits) CK(cudaMemcpyAsync(m->h_lg[sl], m->d_lg[sl], (size_t)n * 64, cudaMemcpyDeviceToHost, m->s_d2h));
    CK(cudaEventRecord(m->e_d2h[sl], m->s_d2h));
  }
  t.busy = true;
  t.logits = want_logits != 0;
  t.n = n;
  t.gen = (t.gen + 1) & 0xffffff;
  *ticket = (int)(t.gen << 4) | sl;
  m->next_slot = (sl + 1) % cvb_model::NSLOT;
  return 0;
}

extern "C" int cvb_predict_collect(cvb_model* m, int ticket, float* base, float* zygosity, float* var_type, float* indel_length,
                                   float* logits16) {
  if (!m) return fail("cvb_predict_collect: NULL model");
  const int sl = ticket & 15;
  if (ticket < 0 || sl >= cvb_model::NSLOT) return fail("cvb_predict_collect: bad ticket %d", ticket);
  cvb_model::Ticket& t = m->tickets[sl];
  if (!t.busy || (int)(t.gen << 4 | sl) != ticket) return fail("cvb_predict_collect: ticket %d is not in flight", ticket);
  const int64_t n = t.n;
  if (n > 0 && (!base || !zygosity || !var_type || !indel_length)) return fail("cvb_predict_collect: NULL buffer");
  if (logits16 && !t.logits) return fail("cvb_predict_collect: ticket %d was submitted without logits", ticket);
  t.busy = false;  // (also on a CUDA error below: the slot is not left blocked)
  if (n == 0) return 0;
  CK(cudaSetDevice(m->device));
  CK(cudaEventSynchronize(m->e_d2h[sl]));
  float* const heads[4] = {base, zygosity, var_type, indel_length};
  const int64_t np = (n + 3) & ~(int64_t)3;
  int64_t off = 0;
  for (int h = 0; h < 4; off += np * kHeadW[h], ++h) memcpy(heads[h], m->h_out[sl] + off, (size_t)n * kHeadW[h] * 4);
  if (logits16) memcpy(logits16, m->h_lg[sl], (size_t)n * 64);
  return 0;
}

extern "C" int cvb_debug_read(cvb_model* m, int which, float* host, int64_t n) {
  if (!m || !host) return fail("cvb_debug_read: NULL argument");
  if (which == 3 || which == 4) {  // fp16 hi/lo pair of p2 (3) / p3 (4) in tensor mode, recombined to fp32
    const int64_t per3 = which == 3 ? m->p2_site : m->p3_site;
    const __half* hi = reinterpret_cast<const __half*>(which == 3 ? m->d_p2 : m->d_p3);
    const __half* lo = hi + (which == 3 ? m->p2_rows * (m->variant == CVB_V3 ? 128 : 64) : m->alloc_sites * per3);
    if (n < 0 || n > m->alloc_sites * per3) return fail("cvb_debug_read: n out of range");
    std::vector<__half> h((size_t)n), l((size_t)n);
    CK(cudaSetDevice(m->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), hi, (size_t)n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(l.data(), lo, (size_t)n * 2, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) host[i] = __half2float(h[i]) + __half2float(l[i]);
    return 0;
  }
  const float* src = which == 0 ? m->d_p2 : which == 1 ? m->d_p3 : which == 2 ? m->d_h4 : nullptr;
  const int64_t per = which == 0 ? m->p2_site : which == 1 ? m->p3_site : m->h4_site;
  if (!src) return fail("cvb_debug_read: bad selector %d", which);
  if (n < 0 || n > m->alloc_sites * per) return fail("cvb_debug_read: n out of range");
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(host, src, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int cvb_profile_begin(cvb_model* m) {
  if (!m) return fail("cvb_profile_begin: NULL model");
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  m->profiling = true;
  m->prof_used = 0;
  return 0;
}
extern "C" int cvb_profile_read(cvb_model* m, double ms[5], int64_t launches[5]) {
  if (!m || !ms || !launches) return fail("cvb_profile_read: NULL argument");
  CK(cudaSetDevice(m->device));
  CK(cudaDeviceSynchronize());
  for (int k = 0; k < 5; ++k) { ms[k] = 0; launches[k] = 0; }
  for (size_t i = 0; i + 6 <= m->prof_used; i += 6)
    for (int k = 0; k < 5; ++k) {
      float t = 0;
      CK(cudaEventElapsedTime(&t, m->prof_events[i + k], m->prof_events[i + k + 1]));
      ms[k] += t;
      launches[k] += 1;
    }
  m->profiling = false;
  return 0;
}

extern "C" int cvb_alloc_pinned(int64_t bytes, void** out) {
  if (!out || bytes <= 0) return fail("cvb_alloc_pinned: bad argument");
  CK(cudaMallocHost(out, (size_t)bytes));
  return 0;
}
extern "C" int cvb_free_pinned(void* p) {
  if (p) CK(cudaFreeHost(p));
  return 0;
}

// ------------------------------------------------------------------------------------
// loss / training (train_simt.cuh), fp32 SIMT kernels, both variants.
// ------------------------------------------------------------------------------------
static const int64_t TRAIN_CHUNK = 5120;  // sites per micro-chunk (activations + gradients ~180 KB/site)

static int ensure_train_work(cvb_model* m) {
  if (m->train) return 0;
  TrainWork* w = new TrainWork();
  w->cap = TRAIN_CHUNK;
  CK(cudaMalloc(&w->seedbuf, 16));
  CK(cudaMemset(w->seedbuf, 0, 16));
  w->graphs_on = !(getenv("CVB_TRAIN_GRAPH") && getenv("CVB_TRAIN_GRAPH")[0] == '0');
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&w->ev_up[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&w->ev_done[i], cudaEventDisableTiming));
  }
  const int64_t c = w->cap;
  struct Item { float** p; int64_t n; };
  // v3_slim (clairvoyante_v3_slim.py:54-118): no pools, so the "pooled + padded" buffers are just the conv outputs moved
  // into the next conv's zero-padded row layout (35 rows for KH=3, 37 for KH=5) and p3 is c3 itself
  Item slim_items[] = {
      {&w->xs[0], c * 528}, {&w->ys[0], c * 16}, {&w->xs[1], c * 528}, {&w->ys[1], c * 16}, {&w->c1, c * 33 * 32}, {&w->p1p, c * 35 * 32}, {&w->c2, c * 33 * 64},
      {&w->p2p, c * 37 * 64}, {&w->c3, c * 33 * 128}, {&w->h4, c * 36}, {&w->d4, c * 36},
      {&w->h5, c * 24}, {&w->logits, c * 16}, {&w->out16, c * 16}, {&w->dlog, c * 16}, {&w->g5, c * 24},
      {&w->g4, c * 36}, {&w->g4b, c * 36}, {&w->gp3, c * 33 * 128}, {&w->g3p, c * 37 * 128}, {&w->gp2, c * 33 * 64},
      {&w->g2p, c * 35 * 64}, {&w->gp1, c * 33 * 32}, {&w->g1, c * 33 * 32}, {&w->w3t, 5 * 4 * 32 * 16},
      {&w->w2t, 3 * 4 * 16 * 8}, {&w->tmpb, 36 * 16}, {&w->tmph, 24 * 16}, {&w->tmpw, 5 * 64 * 128}, {&w->fsc, 16},
      {&w->loss, 16}};
  if (m->variant != CVB_V3) {
    int64_t total = 0;
    for (auto& it : slim_items) total += (it.n + 63) / 64 * 64;
    CK(cudaMalloc(&w->all, (size_t)total * 4));
    CK(cudaMemset(w->all, 0, (size_t)total * 4));
    int64_t off = 0;
    for (auto& it : slim_items) { *it.p = w->all + off; off += (it.n + 63) / 64 * 64; }
    w->p3 = w->c3;
    {
      // tensor-core conv3 + FC4 of the v3_slim step (CVB_TRAIN_FP32 keeps the SIMT kernels): operand planes [hi | lo]
      struct Item16 { uint16_t** p; int64_t n; };
      Item16 it16[] = {{&w->p2h, 2 * c * 37 * 64}, {&w->p2b, 2 * c * 37 * 64}, {&w->g3h, 2 * c * 37 * 128}, {&w->p3s, 2 * c * 4224},
                       {&w->g4s, 2 * c * 40}, {&w->w4s, 2 * 4224 * 40}, {&w->w4ts, 2 * 36 * 4224}, {&w->wf3, 2 * 5 * 128 * 64},
                       {&w->wd3, 2 * 5 * 64 * 128}};
      int64_t t16 = 0;
      for (auto& it : it16) t16 += (it.n + 127) / 128 * 128;
      CK(cudaMalloc(&w->all16, (size_t)t16 * 2));
      CK(cudaMemset(w->all16, 0, (size_t)t16 * 2));  // pad rows / pad columns (g4s, w4s: 36 -> 40) stay zero
      int64_t o16 = 0;
      for (auto& it : it16) { *it.p = w->all16 + o16; o16 += (it.n + 127) / 128 * 128; }
      CK(cudaMalloc(&w->amax, 16));
      const float one = 1.f;
      CK(cudaMemcpy(w->fsc + 2, &one, 4, cudaMemcpyHostToDevice));
    }
    m->train = w;
    return 0;
  }
  Item items[] = {
      {&w->xs[0], c * 528}, {&w->ys[0], c * 16}, {&w->xs[1], c * 528}, {&w->ys[1], c * 16}, {&w->c1, c * 33 * 64}, {&w->p1p, c * 30 * 64}, {&w->c2, c * 29 * 128},
      {&w->p2p, c * 28 * 128}, {&w->c3, c * 26 * 192}, {&w->p3, c * 24 * 192}, {&w->h4, c * 336}, {&w->d4, c * 336},
      {&w->h5, c * 168}, {&w->logits, c * 16}, {&w->out16, c * 16}, {&w->dlog, c * 16}, {&w->g5, c * 176},
      {&w->g4, c * 336}, {&w->g4b, c * 336}, {&w->gp3, c * 24 * 192}, {&w->g3p, c * 28 * 192}, {&w->gp2, c * 26 * 128},
      {&w->g2p, c * 30 * 128}, {&w->gp1, c * 29 * 64}, {&w->g1, c * 33 * 64}, {&w->w3t, 3 * 4 * 48 * 32},
      {&w->w2t, 2 * 4 * 32 * 16}, {&w->w4t, 336 * 4608}, {&w->w5t, 176 * 336}, {&w->tmpb, 336 * 16}, {&w->tmph, 168 * 16},
      {&w->tmp5, 336 * 184}, {&w->tmpw, 3 * 128 * 256}, {&w->fsc, 16}, {&w->loss, 16}};
  int64_t total = 0;
  for (auto& it : items) total += (it.n + 63) / 64 * 64;
  CK(cudaMalloc(&w->all, (size_t)total * 4));
  CK(cudaMemset(w->all, 0, (size_t)total * 4));  // padding rows / columns of the padded layouts stay zero forever
  int64_t off = 0;
  for (auto& it : items) { *it.p = w->all + off; off += (it.n + 63) / 64 * 64; }
  {
    // split-bf16 operand copies for the tensor-core FC4 contractions: each buffer = hi plane followed by lo plane
    w->ldt = c;
    struct Item16 { uint16_t** p; int64_t n; };
    Item16 it16[] = {{&w->p3s, 2 * c * 4608}, {&w->g4s, 2 * c * 336}, {&w->w4s, 2 * 4608 * 336}, {&w->w4ts, 2 * 336 * 4608},
                     {&w->d4t, 2 * 336 * c},  {&w->h5t, 2 * 168 * c},    {&w->gct, 2 * 184 * c},
                     {&w->p1b, 2 * c * 30 * 64}, {&w->p2b, 2 * c * 28 * 128}, {&w->d4s, 2 * c * 336}, {&w->g5s, 2 * c * 168},
                     {&w->w5s, 2 * 336 * 168}, {&w->w5ts, 2 * 168 * 336},
                     {&w->p1h, 2 * c * 30 * 64}, {&w->p2h, 2 * c * 28 * 128}, {&w->g3h, 2 * c * 28 * 256}, {&w->g2h, 2 * c * 30 * 128},
                     {&w->wf2, 2 * 2 * 128 * 64}, {&w->wf3, 2 * 3 * 192 * 128}, {&w->wd2, 2 * 2 * 64 * 128},
                     {&w->wd3, 2 * 3 * 128 * 256}};
    CK(cudaMalloc(&w->amax, 16));
    const float one = 1.f;
    CK(cudaMemcpy(w->fsc + 2, &one, 4, cudaMemcpyHostToDevice));
    int64_t t16 = 0;
    for (auto& it : it16) t16 += (it.n + 127) / 128 * 128;
    CK(cudaMalloc(&w->all16, (size_t)t16 * 2));
    CK(cudaMemset(w->all16, 0, (size_t)t16 * 2));
    int64_t o16 = 0;
    for (auto& it : it16) { *it.p = w->all16 + o16; o16 += (it.n + 127) / 128 * 128; }
  }
  m->train = w;
  return 0;
}

static inline int gsz(int64_t total, int block = 256) { return (int)std::min<int64_t>((total + block - 1) / block, 148 * 16); }


// ---- tensor-core FC4 contractions of the training path (gemm_tc.cuh) ------------------------------------------------
// C[M][N] (op)= A[M][K] . B[N][K]^T ; a / b point at the hi plane, the lo plane follows `a_plane` / `b_plane` elements later;
// lda / ldb = row pitch in elements (multiple of 8: TMA strides are 16-byte granular)
// MN = false: a [M][K], b [N][K] (K-major; lda / ldb = row pitch).  MN = true: a [K][M], b [K][N] (transposed operands read in
// place as MN-major tiles; lda / ldb = pitch of a K row).
template <int BN, bool CHUNKED, int EPI, bool MN>
static int launch_gemm_tc(cvb_model* m, const uint16_t* a, int64_t a_plane, int64_t lda, const uint16_t* b, int64_t b_plane,
                          int64_t ldb, int M, int N, int K, float* C, int64_t ldc, const float* bias, cudaStream_t st,
                          const GemmExtra& ex) {
  using G = tc::GemmTc<BN, MN>;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  const uint64_t sa[1] = {(uint64_t)lda * 2}, sb[1] = {(uint64_t)ldb * 2};
  if (MN) {
    const uint64_t da[2] = {(uint64_t)M, (uint64_t)K}, db[2] = {(uint64_t)N, (uint64_t)K};
    const uint32_t bx[2] = {64, (uint32_t)G::BK};
    if (make_map_nd(&ma_hi, (void*)a, 2, da, sa, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (make_map_nd(&ma_lo, (void*)(a + a_plane), 2, da, sa, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (make_map_nd(&mb_hi, (void*)b, 2, db, sb, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (make_map_nd(&mb_lo, (void*)(b + b_plane), 2, db, sb, bx, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {
    const uint64_t da[2] = {(uint64_t)K, (uint64_t)(M + (ex.batches - 1) * ex.a_batch_rows)}, db[2] = {(uint64_t)K, (uint64_t)N};
    const uint32_t ba[2] = {(uint32_t)G::BK, (uint32_t)G::BM}, bb[2] = {(uint32_t)G::BK, (uint32_t)BN};
    if (make_map_nd(&ma_hi, (void*)a, 2, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (make_map_nd(&ma_lo, (void*)(a + a_plane), 2, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (make_map_nd(&mb_hi, (void*)b, 2, db, sb, bb, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (make_map_nd(&mb_lo, (void*)(b + b_plane), 2, db, sb, bb, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  }
  auto k = tc::k_gemm_tc<BN, CHUNKED, EPI, MN>;
  CK(set_smem(k, G::SMEM_BYTES));
  tc::GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.terms = ex.terms ? ex.terms : (m->train_mode == CVB_TRAIN_BF16 ? 1 : 3);
  g.f16 = ex.f16;
  g.inv_scale = ex.inv_scale;
  g.c_slice_stride = ex.c_slice_stride;
  g.C = C; g.ldc = ldc; g.bias = bias;
  g.m_tiles = (M + G::BM - 1) / G::BM;
  g.a_batch_rows = ex.a_batch_rows; g.c_batch_stride = ex.c_batch_stride;
  g.a_kshift0 = ex.a_kshift0; g.a_kshift_per_batch = ex.a_kshift_per_batch;
  const int nkb = (K + G::BK - 1) / G::BK;
  g.kb_per_slice = (nkb + ex.kslices - 1) / ex.kslices;
  const int kslices = (nkb + g.kb_per_slice - 1) / g.kb_per_slice;  // no empty slice
  const dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)(g.m_tiles * ex.batches), (unsigned)kslices);
  CK(launch_k(k, grid, dim3(G::THREADS), G::SMEM_BYTES, st, 1, 1, ex.pdl, ma_hi, ma_lo, mb_hi, mb_lo, g));
  CK(cudaGetLastError());
  return 0;
}
static inline __nv_bfloat16* bf(uint16_t* p) { return reinterpret_cast<__nv_bfloat16*>(p); }
static inline __half* hp(uint16_t* p) { return reinterpret_cast<__half*>(p); }
static int split_rows_bf16(const float* src, int64_t rows, int cols, uint16_t* dst, int64_t plane, cudaStream_t st, int64_t ld_src = 0,
                           int64_t ld_dst = 0) {
  tc::k_split_bf16<<<gsz(rows * (cols / 4)), 256, 0, st>>>(src, rows, cols, ld_src ? ld_src : cols, bf(dst), bf(dst + plane),
                                                           ld_dst ? ld_dst : cols);
  CK(cudaGetLastError());
  return 0;
}
static int split_transpose_bf16(const float* src, int64_t R, int C, int64_t ld_src, uint16_t* dst, int64_t plane, int64_t ld_dst,
                                cudaStream_t st, int row_shift = 0) {
  if (R <= 0) return 0;
  tc::k_split_transpose_bf16<<<dim3((unsigned)((C + 31) / 32), (unsigned)((R + 63) / 64)), dim3(32, 8), 0, st>>>(
      src, R, C, ld_src, bf(dst), bf(dst + plane), ld_dst, row_shift);
  CK(cudaGetLastError());
  return 0;
}

// Conv weight gradient on tcgen05 (training_op, clairvoyante_v3.py:174, for conv2 / conv3):
//   dW[kh][kw][c][co] = sum_R in[R - 1 + kh][(w', c)] * g[R][(w, co)],   w' = w + kw - 1,
// over the flattened (site, row) index R of the padded layouts (g stored one row down, its pad rows are zero, so terms that
// cross a site boundary vanish).  K = R is the slow index of both factors, i.e. they are the TRANSPOSED operands of the GEMM
// tmpW[kh] = in(shifted by kh - 1 rows)^T . g: one batched (batch = kh), split-K launch of k_gemm_tc in MN-major mode reads
// the split-bf16 planes written by the forward pool kernel (in) and by pool-backward (g, COUT padded to CP) in place and
// accumulates every (w', c) x (w, co) product with fp32 atomics; k_scatter_conv_wgrad adds the 13 of 16 (w', w) pairs that
// are taps into dW.
// SHIFT0 = -(row of the g layout that holds output row 0): -1 for the v3 layers, -2 for v3_slim's 5-row conv3
template <int CIN, int COUT, int KH, int CP, int SHIFT0 = -1>
static int launch_conv_wgrad_tc(cvb_model* m, const uint16_t* in, int64_t in_plane, const uint16_t* g, int64_t g_plane, int64_t R,
                                float* dW, cudaStream_t st) {
  TrainWork* w = m->train;
  constexpr int MA = 4 * CIN, NB = 4 * CP;
  CK(cudaMemsetAsync(w->tmpw, 0, (size_t)KH * MA * NB * 4, st));
  GemmExtra ex;
  ex.batches = KH; ex.a_kshift0 = SHIFT0; ex.a_kshift_per_batch = 1; ex.c_batch_stride = (int64_t)MA * NB;
  ex.kslices = std::max(1, m->num_sms / KH);
  if (launch_gemm_tc<NB, false, tc::GEMM_EPI_ATOMIC, true>(m, in, in_plane, MA, g, g_plane, NB, MA, NB, (int)R, w->tmpw, NB, nullptr,
                                                           st, ex))
    return 1;
  k_scatter_conv_wgrad<CIN, COUT, KH, CP><<<(KH * 4 * CIN * COUT + 255) / 256, 256, 0, st>>>(w->tmpw, dW);
  CK(cudaGetLastError());
  m->launches += 2;
  return 0;
}

// ---- tcgen05 convs of the training path (conv_tc_slab.cuh): POOL = 1, fp32 output, every SELU output kept ------------
// forward (split fp16, scaled weights, bias + SELU) and data gradient (split bf16, flipped kernel, PADL = 2, no activation)
namespace trc {
using Conv2F = tc::ConvTcCfg<30, 2, 16, 32, 29, 1, 29, 0, 6, true>;                      // p1h [.][30][64]  -> c2 [.][29][128]
using Conv3F = tc::ConvTcCfg<28, 3, 32, 48, 26, 1, 26, 0, 4, true>;                      // p2h [.][28][128] -> c3 [.][26][192]
using Conv3D = tc::ConvTcCfg<28, 3, 64, 32, 26, 1, 26, 0, 4, true, 2, false, true>;      // g3h [.][28][256] -> gp2 [.][26][128]
using Conv2D = tc::ConvTcCfg<30, 2, 32, 16, 29, 1, 29, 0, 4, true, 2, false, true>;      // g2h [.][30][128] -> gp1 [.][29][64]
// v3_slim conv3 (5x4, 16 -> 32, rows padded 2 + 33 + 2): forward = the inference configuration; data gradient 32 -> 16
using SlimConv3D = tc::ConvTcCfg<37, 5, 32, 16, 33, 1, 33, 0, 4, true, 2, false, true>;  // g3h [.][37][128] -> gp2 [.][33][64]
using SlimConv3DS = tc::ConvSlabCfg<SlimConv3D, 3, 6>;
using SlimConv3FS = tc::ConvSlabCfg<tc::SlimConv3Tc, 4, 8>;             // the inference geometry, fp32 output, bias read from memory
using Conv2FS = tc::ConvSlabCfg<Conv2F, 4, 8>;
using Conv3FS = tc::ConvSlabCfg<Conv3F, 3, 6>;
using Conv3DS = tc::ConvSlabCfg<Conv3D, 3, 3>;
using Conv2DS = tc::ConvSlabCfg<Conv2D, 3, 6>;
}  // namespace trc

// act: hi plane [rows = nc * RPS][KROW], lo plane act_plane elements later; wts: B [KH * NOUT][KROW] hi then lo plane
template <class F, class S>
static int launch_train_conv(cvb_model* m, const uint16_t* act, int64_t act_plane, const uint16_t* wts, int64_t nc, const float* bias,
                             const float* inv_scale, float* out, cudaStream_t st) {
  const CUtensorMapSwizzle sw = F::ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                    : (F::ROW_BYTES == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap ma, mb[3];
  const uint64_t ad[3] = {(uint64_t)F::KROW, (uint64_t)(nc * F::RPS), 2};
  const uint64_t as[2] = {(uint64_t)F::KROW * 2, (uint64_t)act_plane * 2};
  const uint32_t ab[3] = {(uint32_t)F::BK, (uint32_t)S::SLAB_ROWS, 2};
  if (make_map_nd(&ma, (void*)act, 3, ad, as, ab, sw)) return 1;
  const uint64_t bd[3] = {(uint64_t)F::KROW, (uint64_t)F::B_ROWS_TOTAL, 2};
  const uint64_t bs[2] = {(uint64_t)F::KROW * 2, (uint64_t)F::B_ROWS_TOTAL * F::KROW * 2};
  for (int nb = 2; nb <= 4; ++nb) {
    const uint32_t bb[3] = {(uint32_t)F::BK, (uint32_t)(nb * F::COUT), 2};
    if (make_map_nd(&mb[nb - 2], (void*)wts, 3, bd, bs, bb, sw)) return 1;
  }
  const int64_t tiles = (nc * F::RPS + S::TILE_STEP - 1) / S::TILE_STEP;
  const int grid = (int)std::min<int64_t>(tiles, m->num_sms);
  // (weights through the ring: keeping the taps resident was measured on these launches too -- 3.27 -> 3.25 ms per
  //  10,000-tensor step, inside the noise -- and not kept)
  auto k = tc::k_conv_slab<F, S>;
  CK(set_smem(k, S::SMEM_BYTES));
  k<<<grid, S::THREADS, S::SMEM_BYTES, st>>>(ma, mb[0], mb[1], mb[2], nc, bias, inv_scale, reinterpret_cast<__half*>(out), nullptr, 0, tc::BiasParam{});
  CK(cudaGetLastError());
  m->launches += 1;
  return 0;
}

// training-mode forward of one micro-chunk already resident in tw->x: keeps every SELU output
template <class C>
static int launch_conv_keep(cvb_model* m, const float* in, int64_t nc, const float* wg, const float* bg, float* out, bool act,
                            cudaStream_t st) {
  using L = ConvLayerSmem<C, 1>;
  const int grid = (int)std::min<int64_t>((nc + C::S - 1) / C::S, 2 * m->num_sms);
  if (act) {
    auto k = k_conv_layer<C, 1, 256, true>;
    CK(set_smem(k, L::SMEM_BYTES));
    k<<<grid, 256, L::SMEM_BYTES, st>>>(in, nc, wg, bg, out);
  } else {
    auto k = k_conv_layer<C, 1, 256, false>;
    CK(set_smem(k, L::SMEM_BYTES));
    k<<<grid, 256, L::SMEM_BYTES, st>>>(in, nc, wg, bg, out);
  }
  CK(cudaGetLastError());
  return 0;
}

// dropout5 (clairvoyante_v3.py:121): h5 -> d5 under the FC5 stream of the step's seed; returns what the heads read
static const uint64_t kSeed5 = 0x5D5D5D5D5D5D5D5Dull;
static int train_dropout5(cvb_model* m, int64_t nc, int n5, uint64_t seed, int64_t site0, cudaStream_t st, const float** h5in) {
  TrainWork* w = m->train;
  *h5in = w->h5;
  if (m->drop5_now <= 0.f) return 0;
  if (!w->d5) CK(cudaMalloc(&w->d5, (size_t)w->cap * 168 * 4));
  k_dropout_fwd<<<gsz(nc * n5), 256, 0, st>>>(w->h5, w->d5, nc * n5, site0 * n5, SeedRef{w->seedbuf, kSeed5}, drop_const(m->drop5_now));
  CK(cudaGetLastError());
  m->launches += 1;
  *h5in = w->d5;
  return 0;
}

static int train_forward_slim(cvb_model* m, int64_t nc, float drop4, uint64_t seed, int64_t index0, cudaStream_t st) {
  TrainWork* w = m->train;
  if (launch_conv_keep<ConvCfg<4, 8, 1, 33, 12, 8, 8>>(m, w->x, nc, m->var("conv1/kernel"), m->var("conv1/bias"), w->c1, true, st)) return 1;
  k_pool_fwd<1><<<gsz(nc * 33 * 8), 256, 0, st>>>(w->c1, nc, 33, 32, w->p1p, 35, 1, nullptr, nullptr);
  if (launch_conv_keep<ConvCfg<8, 16, 3, 33, 6, 8, 8>>(m, w->p1p, nc, m->var("conv2/kernel"), m->var("conv2/bias"), w->c2, true, st)) return 1;
  const bool stc = m->train_mode != CVB_TRAIN_FP32;
  k_pool_fwd<1><<<gsz(nc * 33 * 16), 256, 0, st>>>(w->c2, nc, 33, 64, w->p2p, 37, 2, stc ? hp(w->p2h) : nullptr,
                                                   stc ? hp(w->p2h) + w->cap * 37 * 64 : nullptr, stc ? bf(w->p2b) : nullptr,
                                                   stc ? bf(w->p2b) + w->cap * 37 * 64 : nullptr);
  if (stc) {
    // conv3 on the tcgen05 slab kernel (the inference configuration already keeps every SELU output: no pooling in slim),
    // FC4 = c3 [sites][4224] . W4 as a split-bf16 GEMM (N = 36 in one 48-column tile)
    if (launch_train_conv<tc::SlimConv3Tc, trc::SlimConv3FS>(m, w->p2h, w->cap * 37 * 64, w->wf3, nc, m->var("conv3/bias"), w->fsc + 1,
                                                              w->c3, st))
      return 1;
    if (split_rows_bf16(w->c3, nc, 4224, w->p3s, w->cap * 4224, st)) return 1;
    if (launch_gemm_tc<48, true, tc::GEMM_EPI_BIAS_SELU>(m, w->p3s, w->cap * 4224, 4224, w->w4ts, 36 * 4224, 4224, (int)nc, 36, 4224,
                                                         w->h4, 36, m->var("fc4/bias"), st))
      return 1;
    m->launches += 2;
  } else {
    if (launch_conv_keep<ConvCfg<16, 32, 5, 33, 3, 8, 8>>(m, w->p2p, nc, m->var("conv3/kernel"), m->var("conv3/bias"), w->c3, true, st)) return 1;
    using F = FcCfg<36, 9, 4, 28, 8>;
    auto k = k_fc4<F, true>;
    CK(set_smem(k, F::SMEM_BYTES));
    k<<<(int)((nc + F::M - 1) / F::M), 256, F::SMEM_BYTES, st>>>(w->c3, nc, 4224, m->var("fc4/kernel"), m->var("fc4/bias"), w->h4, 36, 36);
    CK(cudaGetLastError());
  }
  const float* d4 = w->h4;
  if (drop4 > 0.f) {
    k_dropout_fwd<<<gsz(nc * 36), 256, 0, st>>>(w->h4, w->d4, nc * 36, index0, SeedRef{w->seedbuf, 0}, drop_const(drop4));
    d4 = w->d4;
  }
  k_dense_small<true><<<gsz(nc * 18), 256, 0, st>>>(d4, 36, 36, m->var("fc5/kernel"), 18, m->var("fc5/bias"), w->h5, 18, nc);
  const float* h5in;
  if (train_dropout5(m, nc, 18, seed, index0 / 36, st, &h5in)) return 1;
  k_heads<36, 18><<<(int)((nc + 15) / 16), 256, 0, st>>>(d4, h5in, nc, head_ptrs(m), out_interleaved(w->out16), w->logits);
  CK(cudaGetLastError());
  m->launches += 8 + (drop4 > 0.f ? 1 : 0);
  return 0;
}

static float* gvar(cvb_model* m, const char* name) { return m->d_grad + m->info(name)->offset; }

// The weight gradients (split-K GEMMs + atomics, ~190 us of a 680 us step at 1,250 tensors) do not feed anything later in
// the backward pass: they run on a second stream, forked from `st` wherever their operands are complete and joined at the
// end of the micro-chunk, so that in the captured graph they are a parallel branch beside the data-gradient chain
// (pool-backward -> dgrad conv -> ...).  CVB_TRAIN_AUX=0 keeps everything on one stream.
static cudaStream_t bwd_fork(cvb_model* m, cudaStream_t st, bool tensor_path, bool* forked) {
  static const bool aux_on = !(getenv("CVB_TRAIN_AUX") && getenv("CVB_TRAIN_AUX")[0] == '0');
  if (!aux_on || !tensor_path) return st;
  cudaEventRecord(m->e_fork, st);
  cudaStreamWaitEvent(m->s_aux, m->e_fork, 0);
  *forked = true;
  return m->s_aux;
}
static int bwd_join(cvb_model* m, cudaStream_t st, bool forked) {
  if (!forked) return 0;  // the branch rejoins before the chunk ends (and before the next chunk reuses its operands)
  CK(cudaEventRecord(m->e_join, m->s_aux));
  CK(cudaStreamWaitEvent(st, m->e_join, 0));
  return 0;
}

static int train_backward_slim(cvb_model* m, int64_t nc, float drop4, uint64_t seed, int64_t index0, cudaStream_t st) {
  TrainWork* w = m->train;
  bool forked = false;
  const int sms = m->num_sms;
  const float* d4 = drop4 > 0.f ? w->d4 : w->h4;
  // heads
  CK(cudaMemsetAsync(w->tmpb, 0, 36 * 16 * 4, st));
  CK(cudaMemsetAsync(w->tmph, 0, 24 * 16 * 4, st));
  k_gemm_tn<<<dim3(1, 1), 256, 0, st>>>(d4, 36, w->dlog, 16, w->tmpb, 16, 36, 16, nc);
  const float* h5in = m->drop5_now > 0.f ? w->d5 : w->h5;
  k_gemm_tn<<<dim3(1, 1), 256, 0, st>>>(h5in, 18, w->dlog, 16, w->tmph, 16, 18, 16, nc);
  HeadG hg{gvar(m, "YBaseChangeSigmoid/kernel"), gvar(m, "YZygosityFC/kernel"), gvar(m, "YVarTypeFC/kernel"),
           gvar(m, "YIndelLengthFC/kernel")};
  k_scatter_heads<<<1, 256, 0, st>>>(w->tmpb, w->tmph, 36, 18, hg);
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 0, nc, 16, 4, gvar(m, "YBaseChangeSigmoid/bias"));
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 4, nc, 16, 2, gvar(m, "YZygosityFC/bias"));
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 6, nc, 16, 4, gvar(m, "YVarTypeFC/bias"));
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 10, nc, 16, 6, gvar(m, "YIndelLengthFC/bias"));
  HeadW hw{m->var("YBaseChangeSigmoid/kernel"), m->var("YZygosityFC/kernel"), m->var("YVarTypeFC/kernel"),
           m->var("YIndelLengthFC/kernel")};
  k_heads_bwd<<<gsz(nc * (36 + 18)), 256, 0, st>>>(w->dlog, w->h5, nc, 36, 18, hw, w->g4, w->g5, 24, m->drop5_now > 0.f ? 1 : 0,
                                                   SeedRef{w->seedbuf, kSeed5}, index0 / 36 * 18, drop_const(m->drop5_now));
  // FC5
  k_gemm_tn<<<dim3(1, 1), 256, 0, st>>>(d4, 36, w->g5, 24, gvar(m, "fc5/kernel"), 18, 36, 18, nc);
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->g5, nc, 24, 18, gvar(m, "fc5/bias"));
  k_dense_bwd_small<<<gsz(nc * 36), 256, 0, st>>>(w->g5, 24, 18, m->var("fc5/kernel"), 36, w->g4b, 36, nc);
  k_fc4_bwd_elem<<<gsz(nc * 36), 256, 0, st>>>(w->g4, w->g4b, w->h4, nc * 36, index0, SeedRef{w->seedbuf, 0}, drop_const(drop4), drop4 > 0.f ? 1 : 0);
  // FC4
  const bool stc = m->train_mode != CVB_TRAIN_FP32;
  k_colsum<<<dim3(2, 32), 256, 0, st>>>(w->g4, nc, 36, 36, gvar(m, "fc4/bias"));
  if (stc) {
    // dpre4 [sites][36] as split bf16 with rows padded to 40 (TMA strides are 16-byte granular); weight gradient reads
    // c3 / dpre4 in place as MN-major operands (K = sites, 33 output tiles x 4 K slices), data gradient is K-major (K = 36)
    if (split_rows_bf16(w->g4, nc, 36, w->g4s, w->cap * 40, st, 36, 40)) return 1;
    GemmExtra k4;
    k4.kslices = 4;
    if (launch_gemm_tc<64, true, tc::GEMM_EPI_ATOMIC, true>(m, w->p3s, w->cap * 4224, 4224, w->g4s, w->cap * 40, 40, 4224, 36, (int)nc,
                                                            gvar(m, "fc4/kernel"), 36, nullptr, bwd_fork(m, st, true, &forked), k4))
      return 1;
    if (launch_gemm_tc<192, false, tc::GEMM_EPI_STORE>(m, w->g4s, w->cap * 40, 40, w->w4s, 4224 * 40, 40, (int)nc, 4224, 36, w->gp3, 4224,
                                                       nullptr, st))
      return 1;
    k_pool_bwd_selu<1, 128, 256, 32><<<gsz(nc * 33 * 32), 256, 0, st>>>(w->gp3, w->c3, nc, 33, w->g3p, 37, 2, gvar(m, "conv3/bias"),
                                                                        bf(w->g3h), bf(w->g3h) + w->cap * 37 * 128);
    if (launch_conv_wgrad_tc<16, 32, 5, 32, -2>(m, w->p2b, w->cap * 37 * 64, w->g3h, w->cap * 37 * 128, nc * 37, gvar(m, "conv3/kernel"),
                                                bwd_fork(m, st, true, &forked)))
      return 1;
    if (launch_train_conv<trc::SlimConv3D, trc::SlimConv3DS>(m, w->g3h, w->cap * 37 * 128, w->wd3, nc, nullptr, w->fsc + 2, w->gp2, st))
      return 1;
  } else {
  k_gemm_tn<<<dim3(4224 / 64, 1), 256, 0, st>>>(w->c3, 4224, w->g4, 36, gvar(m, "fc4/kernel"), 36, 4224, 36, nc);
  k_dense_bwd_small<<<gsz(nc * 4224), 256, 0, st>>>(w->g4, 36, 36, m->var("fc4/kernel"), 4224, w->gp3, 4224, nc);
  // conv3 (5x4, 16 -> 32)
  k_pool_bwd_selu<1, 128, 256><<<gsz(nc * 33 * 32), 256, 0, st>>>(w->gp3, w->c3, nc, 33, w->g3p, 37, 2, gvar(m, "conv3/bias"));
  {
    using W = WgradCfg<16, 32, 5, 33, 4, 8, 4>;
    auto k = k_conv_wgrad<16, 32, 5, 33, 4, 8, 4>;
    CK(set_smem(k, W::SMEM_BYTES));
    k<<<(int)std::min<int64_t>((nc + 3) / 4, sms), W::THREADS, W::SMEM_BYTES, st>>>(w->p2p, w->g3p, 37, 2, nc, gvar(m, "conv3/kernel"));
    CK(cudaGetLastError());
    if (launch_conv_keep<ConvCfg<32, 16, 5, 33, 3, 8, 8, 2>>(m, w->g3p, nc, w->w3t, nullptr, w->gp2, false, st)) return 1;
  }
  }
  // conv2 (3x4, 8 -> 16)
  k_pool_bwd_selu<1, 64, 256><<<gsz(nc * 33 * 16), 256, 0, st>>>(w->gp2, w->c2, nc, 33, w->g2p, 35, 1, gvar(m, "conv2/bias"));
  {
    using W = WgradCfg<8, 16, 3, 33, 4, 4, 4>;
    auto k = k_conv_wgrad<8, 16, 3, 33, 4, 4, 4>;
    CK(set_smem(k, W::SMEM_BYTES));
    k<<<(int)std::min<int64_t>((nc + 3) / 4, sms), W::THREADS, W::SMEM_BYTES, st>>>(w->p1p, w->g2p, 35, 1, nc, gvar(m, "conv2/kernel"));
    CK(cudaGetLastError());
    if (launch_conv_keep<ConvCfg<16, 8, 3, 33, 6, 8, 8, 2>>(m, w->g2p, nc, w->w2t, nullptr, w->gp1, false, st)) return 1;
  }
  // conv1 (1x4, 4 -> 8)
  k_pool_bwd_selu<1, 32, 256><<<gsz(nc * 33 * 8), 256, 0, st>>>(w->gp1, w->c1, nc, 33, w->g1, 33, 0, gvar(m, "conv1/bias"));
  k_conv1_wgrad<8, 4><<<(int)std::min<int64_t>((nc + 3) / 4, 4 * sms), 128, 0, st>>>(w->x, w->g1, nc, gvar(m, "conv1/kernel"));
  CK(cudaGetLastError());
  if (bwd_join(m, st, forked)) return 1;
  m->launches += 24;
  return 0;
}

static int train_forward(cvb_model* m, int64_t nc, float drop4, uint64_t seed, int64_t index0, cudaStream_t st) {
  if (m->variant != CVB_V3) return train_forward_slim(m, nc, drop4, seed, index0, st);
  TrainWork* w = m->train;
  const int sms = m->num_sms;
  const bool tcm = m->train_mode != CVB_TRAIN_FP32;
  {
    using C = ConvCfg<4, 16, 1, 33, 6, 8, 8>;
    using L = ConvLayerSmem<C, 1>;
    auto k = k_conv_layer<C, 1, 256, true>;
    CK(set_smem(k, L::SMEM_BYTES));
    k<<<(int)std::min<int64_t>((nc + C::S - 1) / C::S, 2 * sms), 256, L::SMEM_BYTES, st>>>(w->x, nc, m->var("conv1/kernel"),
                                                                                          m->var("conv1/bias"), w->c1);
    CK(cudaGetLastError());
    k_pool_fwd<5><<<gsz(nc * 29 * 16), 256, 0, st>>>(w->c1, nc, 33, 64, w->p1p, 30, 0, tcm ? hp(w->p1h) : nullptr,
                                                     tcm ? hp(w->p1h) + w->cap * 30 * 64 : nullptr, tcm ? bf(w->p1b) : nullptr,
                                                     tcm ? bf(w->p1b) + w->cap * 30 * 64 : nullptr);
  }
  if (tcm) {
    if (launch_train_conv<trc::Conv2F, trc::Conv2FS>(m, w->p1h, w->cap * 30 * 64, w->wf2, nc, m->var("conv2/bias"), w->fsc + 0, w->c2, st))
      return 1;
    k_pool_fwd<4><<<gsz(nc * 26 * 32), 256, 0, st>>>(w->c2, nc, 29, 128, w->p2p, 28, 1, hp(w->p2h), hp(w->p2h) + w->cap * 28 * 128,
                                                     bf(w->p2b), bf(w->p2b) + w->cap * 28 * 128);
    if (launch_train_conv<trc::Conv3F, trc::Conv3FS>(m, w->p2h, w->cap * 28 * 128, w->wf3, nc, m->var("conv3/bias"), w->fsc + 1, w->c3, st))
      return 1;
    // pool3 also writes p3 as the split-bf16 A operand of the FC4 forward GEMM
    k_pool_fwd<3, true><<<gsz(nc * 24 * 48), 256, 0, st>>>(w->c3, nc, 26, 192, w->p3, 24, 0, hp(w->p3s), hp(w->p3s) + w->cap * 4608);
    CK(cudaGetLastError());
  } else {
    {
      using C = ConvCfg<16, 32, 2, 29, 4, 8, 8>;
      using L = ConvLayerSmem<C, 1>;
      auto k = k_conv_layer<C, 1, 256, true>;
      CK(set_smem(k, L::SMEM_BYTES));
      k<<<(int)std::min<int64_t>((nc + C::S - 1) / C::S, sms), 256, L::SMEM_BYTES, st>>>(w->p1p, nc, m->var("conv2/kernel"),
                                                                                        m->var("conv2/bias"), w->c2);
      CK(cudaGetLastError());
      k_pool_fwd<4><<<gsz(nc * 26 * 32), 256, 0, st>>>(w->c2, nc, 29, 128, w->p2p, 28, 1, nullptr, nullptr);
    }
    {
      using C = ConvCfg<32, 48, 3, 26, 3, 8, 8>;
      using L = ConvLayerSmem<C, 1>;
      auto k = k_conv_layer<C, 1, 256, true>;
      CK(set_smem(k, L::SMEM_BYTES));
      k<<<(int)std::min<int64_t>((nc + C::S - 1) / C::S, sms), 256, L::SMEM_BYTES, st>>>(w->p2p, nc, m->var("conv3/kernel"),
                                                                                        m->var("conv3/bias"), w->c3);
      CK(cudaGetLastError());
      k_pool_fwd<3><<<gsz(nc * 24 * 48), 256, 0, st>>>(w->c3, nc, 26, 192, w->p3, 24, 0, nullptr, nullptr);
    }
  }
  if (m->train_mode != CVB_TRAIN_FP32) {
    // FC4 on tcgen05: p3 -> split bf16 (K-major), B = W4^T prepared once per step (train_prepare_weights)
    if (nc <= 2560) {
      // few site tiles (a data-parallel shard of the reference's 10,000-tensor batch on 8 GPUs is 1,250 tensors = 20 CTAs,
      // each streaming all of W4: 64 us): K split six ways, every slice stores its partial sums (into gp3, a backward-pass
      // buffer that is free here), and a second pass adds them in slice order + bias + SELU -- deterministic
      GemmExtra k6;
      k6.kslices = 6;
      k6.c_slice_stride = nc * 336;
      if (launch_gemm_tc<176, true, tc::GEMM_EPI_STORE>(m, w->p3s, w->cap * 4608, 4608, w->w4ts, 336 * 4608, 4608, (int)nc, 336, 4608,
                                                        w->gp3, 336, nullptr, st, k6))
        return 1;
      k_bias_selu<<<gsz(nc * 84), 256, 0, st>>>(w->gp3, nc * 336, 6, w->h4, m->var("fc4/bias"), nc * 84, 84);
      CK(cudaGetLastError());
      m->launches += 2;
    } else if (launch_gemm_tc<176, true, tc::GEMM_EPI_BIAS_SELU>(m, w->p3s, w->cap * 4608, 4608, w->w4ts, 336 * 4608, 4608, (int)nc,
                                                                 336, 4608, w->h4, 336, m->var("fc4/bias"), st)) {
      return 1;
    }
    m->launches += 1;
  } else {
    using F = FcCfg<336, 21, 16, 12, 8>;
    auto k = k_fc4<F, true>;
    CK(set_smem(k, F::SMEM_BYTES));
    k<<<(int)((nc + F::M - 1) / F::M), 256, F::SMEM_BYTES, st>>>(w->p3, nc, 4608, m->var("fc4/kernel"), m->var("fc4/bias"), w->h4,
                                                                 336, 336);
    CK(cudaGetLastError());
  }
  const float* d4 = w->h4;
  if (drop4 > 0.f) {
    k_dropout_fwd<<<gsz(nc * 336), 256, 0, st>>>(w->h4, w->d4, nc * 336, index0, SeedRef{w->seedbuf, 0}, drop_const(drop4));
    d4 = w->d4;
  }
  if (tcm) {
    // FC5 on tcgen05: h5 = SELU(d4 . W5 + b5), B = W5^T [168][336] prepared once per step
    if (split_rows_bf16(d4, nc, 336, w->d4s, w->cap * 336, st)) return 1;
    if (launch_gemm_tc<176, false, tc::GEMM_EPI_BIAS_SELU>(m, w->d4s, w->cap * 336, 336, w->w5ts, 168 * 336, 336, (int)nc, 168, 336,
                                                           w->h5, 168, m->var("fc5/bias"), st))
      return 1;
    m->launches += 1;
  } else {
    using F5 = FcCfg<168, 21, 8, 12, 8>;
    auto k = k_fc4<F5, true>;
    CK(set_smem(k, F5::SMEM_BYTES));
    k<<<(int)((nc + F5::M - 1) / F5::M), 256, F5::SMEM_BYTES, st>>>(d4, nc, 336, m->var("fc5/kernel"), m->var("fc5/bias"), w->h5, 168,
                                                                    168);
    CK(cudaGetLastError());
  }
  const float* h5in;
  if (train_dropout5(m, nc, 168, seed, index0 / 336, st, &h5in)) return 1;
  k_heads<336, 168><<<(int)((nc + 15) / 16), 256, 0, st>>>(d4, h5in, nc, head_ptrs(m), out_interleaved(w->out16), w->logits);
  CK(cudaGetLastError());
  m->launches += 9 + (drop4 > 0.f ? 1 : 0);
  return 0;
}

static int train_backward(cvb_model* m, int64_t nc, float drop4, uint64_t seed, int64_t index0, cudaStream_t st) {
  if (m->variant != CVB_V3) return train_backward_slim(m, nc, drop4, seed, index0, st);
  TrainWork* w = m->train;
  const int sms = m->num_sms;
  const float* d4 = drop4 > 0.f ? w->d4 : w->h4;
  // heads: bias gradients; heads -> g4 (base branch), g5 (x selu')
  const bool tcm = m->train_mode != CVB_TRAIN_FP32;
  bool forked = false;
  auto fork = [&]() { return bwd_fork(m, st, tcm, &forked); };
  HeadG hg{gvar(m, "YBaseChangeSigmoid/kernel"), gvar(m, "YZygosityFC/kernel"), gvar(m, "YVarTypeFC/kernel"),
           gvar(m, "YIndelLengthFC/kernel")};
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 0, nc, 16, 4, gvar(m, "YBaseChangeSigmoid/bias"));
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 4, nc, 16, 2, gvar(m, "YZygosityFC/bias"));
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 6, nc, 16, 4, gvar(m, "YVarTypeFC/bias"));
  k_colsum<<<dim3(1, 64), 256, 0, st>>>(w->dlog + 10, nc, 16, 6, gvar(m, "YIndelLengthFC/bias"));
  HeadW hw{m->var("YBaseChangeSigmoid/kernel"), m->var("YZygosityFC/kernel"), m->var("YVarTypeFC/kernel"),
           m->var("YIndelLengthFC/kernel")};
  const float* h5in = m->drop5_now > 0.f ? w->d5 : w->h5;  // what the three softmax heads read
  k_heads_bwd<<<gsz(nc * (336 + 168)), 256, 0, st>>>(w->dlog, w->h5, nc, 336, 168, hw, w->g4, w->g5, 176, m->drop5_now > 0.f ? 1 : 0,
                                                     SeedRef{w->seedbuf, kSeed5}, index0 / 336 * 168, drop_const(m->drop5_now));
  if (tcm) {
    // weight gradients of FC5 and the four heads as two tcgen05 contractions over K = sites:
    //   tmp5 [336][184] = d4^T . [g5 | dlog]   (cols 0..167 -> fc5/kernel, 168..171 -> base head: its input is dropout4)
    //   tmph [168][16]  = h5^T . dlog          (cols 4..15 -> zygosity / varType / indelLength heads)
    const int64_t ldt = w->ldt;
    cudaStream_t sw = fork();
    if (split_transpose_bf16(d4, nc, 336, 336, w->d4t, 336 * ldt, ldt, sw)) return 1;
    if (split_transpose_bf16(h5in, nc, 168, 168, w->h5t, 168 * ldt, ldt, sw)) return 1;
    if (split_transpose_bf16(w->g5, nc, 168, 176, w->gct, 184 * ldt, ldt, sw)) return 1;
    if (split_transpose_bf16(w->dlog, nc, 16, 16, w->gct + 168 * ldt, 184 * ldt, ldt, sw)) return 1;
    // 3 and 2 output tiles only: K (= sites) is split over the SMs and the slices meet in fp32 atomics
    CK(cudaMemsetAsync(w->tmp5, 0, 336 * 184 * 4, sw));
    CK(cudaMemsetAsync(w->tmph, 0, 168 * 16 * 4, sw));
    GemmExtra sk;
    sk.kslices = std::max(1, m->num_sms / 3);
    if (launch_gemm_tc<192, true, tc::GEMM_EPI_ATOMIC>(m, w->d4t, 336 * ldt, ldt, w->gct, 184 * ldt, ldt, 336, 184, (int)nc, w->tmp5,
                                                       184, nullptr, sw, sk))
      return 1;
    sk.kslices = std::max(1, m->num_sms / 4);
    if (launch_gemm_tc<16, true, tc::GEMM_EPI_ATOMIC>(m, w->h5t, 168 * ldt, ldt, w->gct + 168 * ldt, 184 * ldt, ldt, 168, 16,
                                                      (int)nc, w->tmph, 16, nullptr, sw, sk))
      return 1;
    k_scatter_fc5_heads<<<(336 * 168 + 255) / 256, 256, 0, sw>>>(w->tmp5, w->tmph, gvar(m, "fc5/kernel"), hg);
    CK(cudaGetLastError());
  } else {
    CK(cudaMemsetAsync(w->tmpb, 0, 336 * 16 * 4, st));
    CK(cudaMemsetAsync(w->tmph, 0, 168 * 16 * 4, st));
    k_gemm_tn<<<dim3((336 + 63) / 64, 1), 256, 0, st>>>(d4, 336, w->dlog, 16, w->tmpb, 16, 336, 16, nc);
    k_gemm_tn<<<dim3((168 + 63) / 64, 1), 256, 0, st>>>(h5in, 168, w->dlog, 16, w->tmph, 16, 168, 16, nc);
    k_scatter_heads<<<(336 * 4 + 255) / 256, 256, 0, st>>>(w->tmpb, w->tmph, 336, 168, hg);
    k_gemm_tn<<<dim3((336 + 63) / 64, (168 + 63) / 64), 256, 0, st>>>(d4, 336, w->g5, 176, gvar(m, "fc5/kernel"), 168, 336, 168, nc);
  }
  // FC5
  k_colsum<<<dim3((168 + 31) / 32, 32), 256, 0, st>>>(w->g5, nc, 176, 168, gvar(m, "fc5/bias"));
  if (tcm) {
    // data gradient of FC5: g4b [sites][336] = g5 . W5^T   (B = W5 as stored: [336][168] is K-major for K = 168)
    if (split_rows_bf16(w->g5, nc, 168, w->g5s, w->cap * 168, st, 176)) return 1;
    if (launch_gemm_tc<176, false, tc::GEMM_EPI_STORE>(m, w->g5s, w->cap * 168, 168, w->w5s, 336 * 168, 168, (int)nc, 336, 168, w->g4b,
                                                       336, nullptr, st))
      return 1;
  } else {
    using F = FcCfg<336, 21, 16, 12, 8>;
    auto k = k_fc4<F, false>;
    CK(set_smem(k, F::SMEM_BYTES));
    k<<<(int)((nc + F::M - 1) / F::M), 256, F::SMEM_BYTES, st>>>(w->g5, nc, 176, w->w5t, nullptr, w->g4b, 336, 336);
    CK(cudaGetLastError());
  }
  k_fc4_bwd_elem<<<gsz(nc * 336), 256, 0, st>>>(w->g4, w->g4b, w->h4, nc * 336, index0, SeedRef{w->seedbuf, 0}, drop_const(drop4), drop4 > 0.f ? 1 : 0);
  // FC4
  k_colsum<<<dim3((336 + 31) / 32, 32), 256, 0, st>>>(w->g4, nc, 336, 336, gvar(m, "fc4/bias"));
  if (m->train_mode != CVB_TRAIN_FP32) {
    // weight gradient  dW4 [4608][336] += p3^T . dpre4   (K = sites)
    // p3s [sites][4608] (written by pool3) and g4s [sites][336] are its transposed operands: read in place, MN-major
    if (split_rows_bf16(w->g4, nc, 336, w->g4s, w->cap * 336, st)) return 1;
    GemmExtra k2;
    k2.kslices = 2;  // 72 output tiles x 2 K slices = 144 CTAs; the gradient buffer was zeroed at the start of the step
    if (launch_gemm_tc<192, true, tc::GEMM_EPI_ATOMIC, true>(m, w->p3s, w->cap * 4608, 4608, w->g4s, w->cap * 336, 336, 4608, 336,
                                                             (int)nc, gvar(m, "fc4/kernel"), 336, nullptr, fork(), k2))
      return 1;
    // data gradient  gp3 [sites][4608] = dpre4 . W4^T      (B = W4 as stored: [4608][336] is K-major for K = 336)
    if (launch_gemm_tc<192, false, tc::GEMM_EPI_STORE>(m, w->g4s, w->cap * 336, 336, w->w4s, 4608 * 336, 336, (int)nc, 4608, 336,
                                                       w->gp3, 4608, nullptr, st))
      return 1;
    m->launches += 3;
  } else {
    k_gemm_tn<<<dim3(4608 / 64, (336 + 63) / 64), 256, 0, st>>>(w->p3, 4608, w->g4, 336, gvar(m, "fc4/kernel"), 336, 4608, 336, nc);
    using F = FcCfg<192, 24, 8, 10, 8>;
    auto k = k_fc4<F, false>;
    CK(set_smem(k, F::SMEM_BYTES));
    k<<<dim3((unsigned)((nc + F::M - 1) / F::M), 4608 / 192), 256, F::SMEM_BYTES, st>>>(w->g4, nc, 336, w->w4t, nullptr, w->gp3, 4608,
                                                                                       4608);
    CK(cudaGetLastError());
  }
  // conv3
  k_pool_bwd_selu<3, 192, 192, 64><<<gsz(nc * 26 * 48, 192), 192, 0, st>>>(w->gp3, w->c3, nc, 26, w->g3p, 28, 1, gvar(m, "conv3/bias"),
                                                                            tcm ? bf(w->g3h) : nullptr,
                                                                            tcm ? bf(w->g3h) + w->cap * 28 * 256 : nullptr);
  {
    if (tcm) {
      if (launch_conv_wgrad_tc<32, 48, 3, 64>(m, w->p2b, w->cap * 28 * 128, w->g3h, w->cap * 28 * 256, nc * 28, gvar(m, "conv3/kernel"),
                                              fork()))
        return 1;
    } else {
      using W = WgradCfg<32, 48, 3, 26, 4, 12, 4>;
      auto k = k_conv_wgrad<32, 48, 3, 26, 4, 12, 4>;
      CK(set_smem(k, W::SMEM_BYTES));
      k<<<(int)std::min<int64_t>((nc + 3) / 4, sms), W::THREADS, W::SMEM_BYTES, st>>>(w->p2p, w->g3p, 28, 1, nc, gvar(m, "conv3/kernel"));
      CK(cudaGetLastError());
    }
    if (tcm) {
      if (launch_train_conv<trc::Conv3D, trc::Conv3DS>(m, w->g3h, w->cap * 28 * 256, w->wd3, nc, nullptr, w->fsc + 2, w->gp2, st)) return 1;
    } else {
      using C = ConvCfg<48, 32, 3, 26, 3, 8, 8, 2>;
      using L = ConvLayerSmem<C, 1>;
      auto kd = k_conv_layer<C, 1, 256, false>;
      CK(set_smem(kd, L::SMEM_BYTES));
      kd<<<(int)std::min<int64_t>((nc + C::S - 1) / C::S, sms), 256, L::SMEM_BYTES, st>>>(w->g3p, nc, w->w3t, nullptr, w->gp2);
      CK(cudaGetLastError());
    }
  }
  // conv2
  k_pool_bwd_selu<4, 128, 256, 32><<<gsz(nc * 29 * 32), 256, 0, st>>>(w->gp2, w->c2, nc, 29, w->g2p, 30, 1, gvar(m, "conv2/bias"),
                                                                       tcm ? bf(w->g2h) : nullptr,
                                                                       tcm ? bf(w->g2h) + w->cap * 30 * 128 : nullptr);
  {
    if (tcm) {
      if (launch_conv_wgrad_tc<16, 32, 2, 32>(m, w->p1b, w->cap * 30 * 64, w->g2h, w->cap * 30 * 128, nc * 30, gvar(m, "conv2/kernel"),
                                              fork()))
        return 1;
    } else {
      using W = WgradCfg<16, 32, 2, 29, 4, 4, 4>;
      auto k = k_conv_wgrad<16, 32, 2, 29, 4, 4, 4>;
      CK(set_smem(k, W::SMEM_BYTES));
      k<<<(int)std::min<int64_t>((nc + 3) / 4, sms), W::THREADS, W::SMEM_BYTES, st>>>(w->p1p, w->g2p, 30, 1, nc, gvar(m, "conv2/kernel"));
      CK(cudaGetLastError());
    }
    if (tcm) {
      if (launch_train_conv<trc::Conv2D, trc::Conv2DS>(m, w->g2h, w->cap * 30 * 128, w->wd2, nc, nullptr, w->fsc + 2, w->gp1, st)) return 1;
    } else {
      using C = ConvCfg<32, 16, 2, 29, 4, 8, 8, 2>;
      using L = ConvLayerSmem<C, 1>;
      auto kd = k_conv_layer<C, 1, 256, false>;
      CK(set_smem(kd, L::SMEM_BYTES));
      kd<<<(int)std::min<int64_t>((nc + C::S - 1) / C::S, sms), 256, L::SMEM_BYTES, st>>>(w->g2p, nc, w->w2t, nullptr, w->gp1);
      CK(cudaGetLastError());
    }
  }
  // conv1 (no data gradient needed)
  k_pool_bwd_selu<5, 64, 256><<<gsz(nc * 33 * 16), 256, 0, st>>>(w->gp1, w->c1, nc, 33, w->g1, 33, 0, gvar(m, "conv1/bias"));
  k_conv1_wgrad<16, 4><<<(int)std::min<int64_t>((nc + 3) / 4, 4 * sms), 256, 0, st>>>(w->x, w->g1, nc, gvar(m, "conv1/kernel"));
  CK(cudaGetLastError());
  if (bwd_join(m, st, forked)) return 1;
  m->launches += 24;
  return 0;
}

static int train_prepare_weights(cvb_model* m, cudaStream_t st, bool backward) {
  TrainWork* w = m->train;
  if (m->variant != CVB_V3) {
    const bool stc = m->train_mode != CVB_TRAIN_FP32;
    if (stc) {  // forward operands (getLoss included): conv3 weights rearranged + scaled fp16, W4^T [36][4224] split bf16
      using F3 = tc::SlimConv3Tc;
      CK(cudaMemsetAsync(w->amax, 0, 8, st));
      tc::k_absmax<<<32, 256, 0, st>>>(m->var("conv3/kernel"), 5 * 4 * 16 * 32, w->amax + 1);
      tc::k_prep_conv_weights<F3><<<(F3::B_ROWS_TOTAL * F3::KROW + 255) / 256, 256, 0, st>>>(
          m->var("conv3/kernel"), w->amax + 1, hp(w->wf3), hp(w->wf3) + F3::B_ROWS_TOTAL * F3::KROW, w->fsc + 1);
      if (split_transpose_bf16(m->var("fc4/kernel"), 4224, 36, 36, w->w4ts, 36 * 4224, 4224, st)) return 1;
      m->launches += 3;
    }
    if (!backward) return 0;
    k_flip_conv_weights<<<(5 * 4 * 16 * 32 + 255) / 256, 256, 0, st>>>(m->var("conv3/kernel"), 5, 16, 32, w->w3t);
    k_flip_conv_weights<<<(3 * 4 * 8 * 16 + 255) / 256, 256, 0, st>>>(m->var("conv2/kernel"), 3, 8, 16, w->w2t);
    if (stc) {
      using D3 = trc::SlimConv3D;
      if (split_rows_bf16(m->var("fc4/kernel"), 4224, 36, w->w4s, 4224 * 40, st, 36, 40)) return 1;
      tc::k_prep_conv_weights_bf16<D3, 32><<<(D3::B_ROWS_TOTAL * D3::KROW + 255) / 256, 256, 0, st>>>(
          w->w3t, bf(w->wd3), bf(w->wd3) + D3::B_ROWS_TOTAL * D3::KROW);
      m->launches += 2;
    }
    CK(cudaGetLastError());
    m->launches += 2;
    return 0;
  }
  if (m->train_mode != CVB_TRAIN_FP32) {  // the forward pass (getLoss included) reads W4^T, the data gradient W4
    if (split_transpose_bf16(m->var("fc4/kernel"), 4608, 336, 336, w->w4ts, 336 * 4608, 4608, st)) return 1;
    if (backward && split_rows_bf16(m->var("fc4/kernel"), 4608, 336, w->w4s, 4608 * 336, st)) return 1;
    if (split_transpose_bf16(m->var("fc5/kernel"), 336, 168, 168, w->w5ts, 168 * 336, 336, st)) return 1;
    if (backward && split_rows_bf16(m->var("fc5/kernel"), 336, 168, w->w5s, 336 * 168, st)) return 1;
    // forward conv2 / conv3 operands: rearranged, power-of-two scaled split-fp16 weights (as on the inference path)
    using F2 = trc::Conv2F;
    using F3 = trc::Conv3F;
    CK(cudaMemsetAsync(w->amax, 0, 8, st));
    tc::k_absmax<<<16, 256, 0, st>>>(m->var("conv2/kernel"), 2 * 4 * 16 * 32, w->amax + 0);
    tc::k_absmax<<<64, 256, 0, st>>>(m->var("conv3/kernel"), 3 * 4 * 32 * 48, w->amax + 1);
    tc::k_prep_conv_weights<F2><<<(F2::B_ROWS_TOTAL * F2::KROW + 255) / 256, 256, 0, st>>>(
        m->var("conv2/kernel"), w->amax + 0, hp(w->wf2), hp(w->wf2) + F2::B_ROWS_TOTAL * F2::KROW, w->fsc + 0);
    tc::k_prep_conv_weights<F3><<<(F3::B_ROWS_TOTAL * F3::KROW + 255) / 256, 256, 0, st>>>(
        m->var("conv3/kernel"), w->amax + 1, hp(w->wf3), hp(w->wf3) + F3::B_ROWS_TOTAL * F3::KROW, w->fsc + 1);
    CK(cudaGetLastError());
    m->launches += backward ? 6 : 5;
  } else if (backward) {
    k_transpose<<<dim3((336 + 31) / 32, 4608 / 32), dim3(32, 8), 0, st>>>(m->var("fc4/kernel"), 4608, 336, w->w4t);
    m->launches += 1;
  }
  if (!backward) return 0;
  k_flip_conv_weights<<<(3 * 4 * 32 * 48 + 255) / 256, 256, 0, st>>>(m->var("conv3/kernel"), 3, 32, 48, w->w3t);
  k_flip_conv_weights<<<(2 * 4 * 16 * 32 + 255) / 256, 256, 0, st>>>(m->var("conv2/kernel"), 2, 16, 32, w->w2t);
  k_transpose<<<dim3((168 + 31) / 32, (336 + 31) / 32), dim3(32, 8), 0, st>>>(m->var("fc5/kernel"), 336, 168, w->w5t);
  if (m->train_mode != CVB_TRAIN_FP32) {  // flipped kernels of the data-gradient convs as split-bf16 tcgen05 operands
    using D2 = trc::Conv2D;
    using D3 = trc::Conv3D;
    tc::k_prep_conv_weights_bf16<D3, 48><<<(D3::B_ROWS_TOTAL * D3::KROW + 255) / 256, 256, 0, st>>>(
        w->w3t, bf(w->wd3), bf(w->wd3) + D3::B_ROWS_TOTAL * D3::KROW);
    tc::k_prep_conv_weights_bf16<D2, 32><<<(D2::B_ROWS_TOTAL * D2::KROW + 255) / 256, 256, 0, st>>>(
        w->w2t, bf(w->wd2), bf(w->wd2) + D2::B_ROWS_TOTAL * D2::KROW);
    m->launches += 2;
  }
  CK(cudaGetLastError());
  m->launches += 3;
  return 0;
}

// Runs `body` (a fixed sequence of launches on `st`) through a CUDA graph: the first use of a key runs eagerly (lazy
// allocations, function attributes), the second is captured and instantiated, every later one is a single cudaGraphLaunch.
// A training step is ~60 launches per micro-chunk of a few microseconds each; at the data-parallel shard sizes (1,250 tensors
// per GPU at the reference's batch of 10,000 on 8 GPUs) the step was bound by launch overhead, not by the kernels.
// What varies between steps is kept out of the launch parameters: the dropout seed is read from device memory (SeedRef),
// the upload slot and the batch offset of the chunk are part of the key.
template <class Body>
static int run_captured(cvb_model* m, uint64_t key, cudaStream_t st, Body body) {
  TrainWork* w = m->train;
  if (!w->graphs_on) return body();
  TrainWork::GraphEntry& e = w->graphs[key];
  if (e.exec) {
    CK(cudaGraphLaunch(e.exec, st));
    m->launches += e.launches;
    return 0;
  }
  if (e.uses++ == 0) return body();
  const int64_t l0 = m->launches;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    w->graphs_on = false;
    return body();
  }
  const int rc = body();
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &g);
  if (rc || ce != cudaSuccess || !g || cudaGraphInstantiate(&e.exec, g, 0) != cudaSuccess) {
    cudaGetLastError();  // something in the sequence cannot be captured on this driver: eager launches from now on
    if (g) cudaGraphDestroy(g);
    e.exec = nullptr;
    w->graphs_on = false;
    m->launches = l0;
    return body();
  }
  cudaGraphDestroy(g);
  e.launches = m->launches - l0;
  CK(cudaGraphLaunch(e.exec, st));
  return 0;
}
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// forward + loss (+ backward) over a host batch in micro-chunks; loss terms accumulate in train->loss[0..3]
static int train_pass(cvb_model* m, const void* xv, int kind, const float* y, int64_t n, float drop4, uint64_t seed, bool backward) {
  if (ensure_train_work(m)) return 1;
  TrainWork* w = m->train;
  cudaStream_t st = m->s_comp;
  const char* x = static_cast<const char*>(xv);
  const size_t esz = (size_t)kXBytes[kind];
  if (kind != X_F32 && !w->xn[0]) {  // narrow upload slots (raw counts / fp16 values), widened on the device per micro-chunk
    for (int i = 0; i < 2; ++i) CK(cudaMalloc(&w->xn[i], (size_t)w->cap * 528 * 2));
  }
  m->drop5_now = backward ? m->drop5 : 0.f;  // phase = False in getLoss (clairvoyante_v3.py:207-216)
  CK(cudaMemcpyAsync(w->seedbuf, &seed, 8, cudaMemcpyHostToDevice, st));  // (pageable source: staged before the call returns)
  // what a captured sequence depends on besides the chunk: pass kind, arithmetic, dropout rates (baked into DropConst)
  const uint64_t cfg = (uint64_t)(backward ? 1 : 0) | ((uint64_t)m->train_mode << 1) | ((uint64_t)(fbits(drop4) >> 8) << 4) |
                       ((uint64_t)(fbits(m->drop5_now) >> 8) << 28) | ((uint64_t)kind << 52);
  auto key_of = [&](int64_t ci, int64_t nc) { return cfg * 0x9E3779B97F4A7C15ull + (uint64_t)(ci + 1) * 1000003ull + (uint64_t)nc * 7919ull; };
  if (run_captured(m, key_of(-1, 0), st, [&]() -> int {
        CK(cudaMemsetAsync(w->loss, 0, 16 * 4, st));
        if (backward) CK(cudaMemsetAsync(m->d_grad, 0, (size_t)(m->nparams + 16) * 4, st));
        return train_prepare_weights(m, st, backward);
      }))
    return 1;
  // micro-chunks alternate between two upload slots: chunk c+1 is copied on the H2D stream (a pageable source blocks the
  // host inside cudaMemcpyAsync) while chunk c's kernels, already enqueued, run on the compute stream
  int64_t ci = 0;
  for (int64_t s0 = 0; s0 < n; s0 += w->cap, ++ci) {
    const int64_t nc = std::min<int64_t>(w->cap, n - s0);
    const int slot = (int)(ci & 1);
    if (ci >= 2) CK(cudaStreamWaitEvent(m->s_h2d, w->ev_done[slot], 0));
    void* xdst = kind == X_F32 ? (void*)w->xs[slot] : w->xn[slot];
    CK(cudaMemcpyAsync(xdst, x + (size_t)s0 * 528 * esz, (size_t)nc * 528 * esz, cudaMemcpyHostToDevice, m->s_h2d));
    CK(cudaMemcpyAsync(w->ys[slot], y + s0 * 16, (size_t)nc * 16 * 4, cudaMemcpyHostToDevice, m->s_h2d));
    CK(cudaEventRecord(w->ev_up[slot], m->s_h2d));
    CK(cudaStreamWaitEvent(st, w->ev_up[slot], 0));
    w->x = w->xs[slot];
    w->y = w->ys[slot];
    const int64_t n4 = m->variant == CVB_V3 ? 336 : 36;  // dropout counter = flat index into the whole batch's FC4 output
    if (run_captured(m, key_of(ci, nc), st, [&]() -> int {
          if (kind != X_F32) {  // counts -> the fp32 tensors the layers read, channel 0 subtracted (utils_v2.py:46)
            const int64_t npos = nc * 132;
            const int g = (int)std::min<int64_t>((npos + 255) / 256, (int64_t)m->num_sms * 16);
            float4* o = reinterpret_cast<float4*>(w->xs[slot]);
            if (kind == X_F16) k_widen<X_F16><<<g, 256, 0, st>>>(w->xn[slot], o, npos);
            else if (kind == X_I16) k_widen<X_I16><<<g, 256, 0, st>>>(w->xn[slot], o, npos);
            else k_widen<X_U8><<<g, 256, 0, st>>>(w->xn[slot], o, npos);
            CK(cudaGetLastError());
            m->launches += 1;
          }
          if (train_forward(m, nc, drop4, seed, s0 * n4, st)) return 1;
          k_loss_grad<<<gsz(nc, 128), 128, 0, st>>>(w->logits, w->out16, w->y, nc, backward ? w->dlog : nullptr, w->loss);
          CK(cudaGetLastError());
          m->launches += 1;
          if (backward && train_backward(m, nc, drop4, seed, s0 * n4, st)) return 1;
          return 0;
        }))
      return 1;
    CK(cudaEventRecord(w->ev_done[slot], st));
  }
  return 0;
}

extern "C" int cvb_loss_host(cvb_model* m, const float* x, const float* y, int64_t n, float* loss) {
  return cvb_loss_host_x(m, x, X_F32, y, n, loss);
}
extern "C" int cvb_loss_host_x(cvb_model* m, const void* x, int x_kind, const float* y, int64_t n, float* loss) {
  if (!m || !loss) return fail("cvb_loss_host: NULL argument");
  if (n < 0) return fail("cvb_loss_host: negative n");
  if (x_kind < 0 || x_kind > 3) return fail("cvb_loss_host: unknown element kind %d", x_kind);
  *loss = 0.f;
  if (n == 0) return 0;
  if (!x || !y) return fail("cvb_loss_host: NULL buffer");
  CK(cudaSetDevice(m->device));
  if (train_pass(m, x, x_kind, y, n, 0.f, 0, false)) return 1;
  float l[4];
  CK(cudaMemcpyAsync(l, m->train->loss, 16, cudaMemcpyDeviceToHost, m->s_comp));
  CK(cudaStreamSynchronize(m->s_comp));
  *loss = l[0] + l[1] + l[2] + l[3];  // getLoss feeds lambda = 0 (clairvoyante_v3.py:207-216)
  return 0;
}

// finishes a step on the compute stream: one k_adam_flat launch (Adam on all 18 variables + sum of squares of the
// pre-update kernels), then loss6 = [total, loss1, loss2, loss3, loss4, lossL2] from the (possibly all-reduced) sums at
// d_grad[nparams..+4) -- one launch and one host synchronisation where the first version had 27 launches and two.
extern "C" int cvb_apply_adam(cvb_model* m, float lr, float l2, float* loss6) {
  if (!m) return fail("cvb_apply_adam: NULL model");
  if (!m->train) return fail("cvb_apply_adam: no gradients (call cvb_train_step_host first)");
  if (m->tickets_out()) return fail("cvb_apply_adam: %d submitted batch(es) still in flight (cvb_predict_collect)", m->tickets_out());
  CK(cudaSetDevice(m->device));
  cudaStream_t st = m->s_comp;
  TrainWork* w = m->train;
  m->step += 1;
  const double b1 = 0.9, b2 = 0.999;
  const float lr_t = (float)(lr * sqrt(1.0 - pow(b2, (double)m->step)) / (1.0 - pow(b1, (double)m->step)));
  BiasRanges br;
  br.n = 0;
  for (auto& v : m->vars)
    if (v.name.find("bias") != std::string::npos) {  // clairvoyante_v3.py:150
      br.lo[br.n] = v.offset / 4;
      br.hi[br.n] = (v.offset + (v.numel + 3) / 4 * 4) / 4;
      ++br.n;
    }
  CK(cudaMemsetAsync(w->loss + 8, 0, 4, st));
  const int64_t n4 = m->nparams / 4;
  k_adam_flat<<<gsz(n4), 256, 0, st>>>(reinterpret_cast<float4*>(m->d_params), reinterpret_cast<float4*>(m->d_m),
                                       reinterpret_cast<float4*>(m->d_v), reinterpret_cast<const float4*>(m->d_grad), n4, lr_t,
                                       (float)b1, (float)b2, 1e-8f, l2, br, w->loss + 8);
  CK(cudaGetLastError());
  float l[5];
  CK(cudaMemcpyAsync(l, m->d_grad + m->nparams, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(l + 4, w->loss + 8, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (loss6) {
    loss6[5] = l2 * 0.5f * l[4];
    for (int i = 0; i < 4; ++i) loss6[1 + i] = l[i];
    loss6[0] = l[0] + l[1] + l[2] + l[3] + loss6[5];
  }
  m->launches += 1;
  m->tc_weights_dirty = true;
  return 0;
}

extern "C" int cvb_train_step_host(cvb_model* m, const float* x, const float* y, int64_t n, float lr, float l2, float drop4,
                                   uint64_t dropout_seed, int apply_update, float* loss6) {
  return cvb_train_step_host_x(m, x, X_F32, y, n, lr, l2, drop4, dropout_seed, apply_update, loss6);
}
extern "C" int cvb_train_step_host_x(cvb_model* m, const void* x, int x_kind, const float* y, int64_t n, float lr, float l2,
                                     float drop4, uint64_t dropout_seed, int apply_update, float* loss6) {
  if (!m) return fail("cvb_train_step_host: NULL model");
  if (x_kind < 0 || x_kind > 3) return fail("cvb_train_step_host: unknown element kind %d", x_kind);
  if (n <= 0) return fail("cvb_train_step_host: empty batch");
  if (m->tickets_out()) return fail("cvb_train_step_host: %d submitted batch(es) still in flight (cvb_predict_collect)", m->tickets_out());
  if (!x || !y) return fail("cvb_train_step_host: NULL buffer");
  if (drop4 < 0.f || drop4 >= 1.f) return fail("cvb_train_step_host: dropout rate %g outside [0,1)", drop4);
  CK(cudaSetDevice(m->device));
  if (train_pass(m, x, x_kind, y, n, drop4, dropout_seed, true)) return 1;
  // loss sums ride at the tail of the gradient buffer so that one all-reduce covers both
  CK(cudaMemcpyAsync(m->d_grad + m->nparams, m->train->loss, 16, cudaMemcpyDeviceToDevice, m->s_comp));
  if (apply_update) {
    if (allreduce_gradients(m, m->s_comp)) return 1;  // data-parallel: this rank's shard -> the global batch's sums
    return cvb_apply_adam(m, lr, l2, loss6);
  }
  CK(cudaStreamSynchronize(m->s_comp));
  return 0;
}

extern "C" int cvb_grad_buffer(cvb_model* m, void** dev_ptr, int64_t* numel) {
  if (!m || !dev_ptr || !numel) return fail("cvb_grad_buffer: NULL argument");
  *dev_ptr = m->d_grad;
  *numel = m->nparams + 4;  // gradients (16-byte aligned slots per variable) followed by the four loss sums
  return 0;
}
// ---- data-parallel training: the SUM all-reduce of [gradients | 4 loss sums] runs inside the library, on the compute
// stream, between the backward pass and the optimiser kernel -- no host synchronisation and nothing on another stream
// (the first version reduced through torch.distributed on torch's stream and had to synchronise the host before Adam).
static void nccl_release(cvb_model* m) {
  if (m->nccl_comm && m->nccl_owned) {
    NcclApi* a = nccl_api();
    if (a) a->CommDestroy(m->nccl_comm);
  }
  m->nccl_comm = nullptr;
  m->nccl_owned = false;
  m->nccl_ranks = 1;
}
extern "C" int cvb_nccl_unique_id(void* id128) {
  if (!id128) return fail("cvb_nccl_unique_id: NULL argument");
  NcclApi* a = nccl_api();
  if (!a) return 1;
  NK(a, a->GetUniqueId(static_cast<NcclId*>(id128)));
  return 0;
}
extern "C" int cvb_allreduce_init(cvb_model* m, const void* id128, int nranks, int rank) {
  if (!m || !id128) return fail("cvb_allreduce_init: NULL argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("cvb_allreduce_init: rank %d of %d", rank, nranks);
  NcclApi* a = nccl_api();
  if (!a) return 1;
  CK(cudaSetDevice(m->device));
  nccl_release(m);
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  NK(a, a->CommInitRank(&comm, nranks, id, rank));
  m->nccl_comm = comm;
  m->nccl_owned = true;
  m->nccl_ranks = nranks;
  return 0;
}
extern "C" int cvb_allreduce_attach(cvb_model* m, void* nccl_comm) {
  if (!m) return fail("cvb_allreduce_attach: NULL model");
  nccl_release(m);
  if (!nccl_comm) return 0;
  NcclApi* a = nccl_api();
  if (!a) return 1;
  int n = 0;
  NK(a, a->CommCount(nccl_comm, &n));
  m->nccl_comm = nccl_comm;
  m->nccl_ranks = n;
  return 0;
}
static int allreduce_gradients(cvb_model* m, cudaStream_t st) {
  if (!m->nccl_comm || m->nccl_ranks < 2) return 0;
  NcclApi* a = nccl_api();
  if (!a) return 1;
  NK(a, a->AllReduce(m->d_grad, m->d_grad, (size_t)(m->nparams + 4), 7 /* ncclFloat32 */, 0 /* ncclSum */, m->nccl_comm, st));
  return 0;
}
extern "C" int cvb_allreduce_gradients(cvb_model* m) {
  if (!m) return fail("cvb_allreduce_gradients: NULL model");
  CK(cudaSetDevice(m->device));
  return allreduce_gradients(m, m->s_comp);
}

extern "C" int cvb_get_gradient(cvb_model* m, const char* name, float* host, int64_t n) {
  if (!m || !name || !host) return fail("cvb_get_gradient: NULL argument");
  const VarInfo* v = m->info(name);
  if (!v) return fail("cvb_get_gradient: no variable named '%s'", name);
  if (n != v->numel) return fail("cvb_get_gradient: size mismatch for '%s'", name);
  CK(cudaSetDevice(m->device));
  CK(cudaStreamSynchronize(m->s_comp));
  CK(cudaMemcpy(host, m->d_grad + v->offset, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}
