// C ABI of libtpx.so (see include/tpx.h): argument checking, parameter packing and the
// dispatch onto the kernel instantiations.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "fcn_kernels.cuh"
#include "tc_kernels.h"
#include "tpx_internal.h"

namespace tpx {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

static int pad_width(int w) {
  if (w <= 32) return 32;
  if (w <= 64) return 64;
  if (w <= 128) return 128;
  if (w <= 256) return 256;
  return -1;
}

static int make_layout(const tpx_fcn_desc* net, Layout* lay) {
  if (!net) return fail(TPX_E_ARG, "net is NULL");
  if (net->n_hidden < 1 || net->n_hidden > TPX_MAX_HIDDEN)
    return fail(TPX_E_ARG, "n_hidden=%d outside 1..%d", net->n_hidden, TPX_MAX_HIDDEN);
  if (net->d_in < 1 || net->d_in > TPX_MAX_DIN) return fail(TPX_E_UNSUPPORTED, "d_in=%d outside 1..%d", net->d_in, TPX_MAX_DIN);
  if (net->d_out < 1 || net->d_out > TPX_MAX_DOUT) return fail(TPX_E_UNSUPPORTED, "d_out=%d outside 1..%d", net->d_out, TPX_MAX_DOUT);
  if (net->activation != TPX_ACT_TANH && net->activation != TPX_ACT_SIN)
    return fail(TPX_E_UNSUPPORTED, "activation id %d", net->activation);
  int wmax = 0;
  for (int l = 0; l < net->n_hidden; ++l) {
    if (net->widths[l] < 1) return fail(TPX_E_ARG, "width[%d]=%d", l, net->widths[l]);
    if (net->widths[l] > wmax) wmax = net->widths[l];
  }
  int hp = pad_width(wmax);
  // narrow nets with more hidden matrices than the resident-weight kernels hold run on the layer-major tensor-core
  // kernels as one zero-padded 128-neuron block per layer
  if (hp > 0 && hp < 128 && tcw_takes_narrow(net)) hp = 128;
  if (hp < 0) return fail(TPX_E_UNSUPPORTED, "hidden width %d > %d", wmax, TPX_MAX_WIDTH);
  lay->L = net->n_hidden;
  lay->HP = hp;
  lay->d_in = net->d_in;
  lay->d_out = net->d_out;
  return 0;
}

static int check_jet(const tpx_fcn_desc* net, const tpx_jet_desc* jet, JetArgs* ja) {
  if (!jet) return fail(TPX_E_ARG, "jet is NULL");
  if (jet->nd < 0 || jet->nd > TPX_MAX_ND) return fail(TPX_E_UNSUPPORTED, "nd=%d outside 0..%d", jet->nd, TPX_MAX_ND);
  if (jet->lap && jet->nd == 0) return fail(TPX_E_ARG, "lap requires nd >= 1");
  ja->nd = jet->nd;
  ja->out_major = jet->out_major ? 1 : 0;
  ja->out_ld = 0;
  if (ja->out_major) {
    Layout lay;
    int rc = make_layout(net, &lay);
    if (rc) return rc;
    if (!lay.wide_out() || !tcw_supported(net, jet->nd, jet->lap ? 1 : 0))
      return fail(TPX_E_UNSUPPORTED, "output-major streams need 9 or more outputs on the layer-major tensor-core kernels");
  }
  for (int k = 0; k < TPX_MAX_ND; ++k) {
    ja->dcol[k] = 0;
    ja->lap_w[k] = 0.f;
  }
  for (int k = 0; k < jet->nd; ++k) {
    if (jet->dcol[k] < 0 || jet->dcol[k] >= net->d_in) return fail(TPX_E_ARG, "dcol[%d]=%d", k, jet->dcol[k]);
    ja->dcol[k] = jet->dcol[k];
    ja->lap_w[k] = jet->lap ? jet->lap_w[k] : 0.f;
  }
  return 0;
}

struct PackArgs {
  const float* W[TPX_MAX_HIDDEN + 1];
  const float* b[TPX_MAX_HIDDEN + 1];
  int in_w[TPX_MAX_HIDDEN + 1];
  int out_w[TPX_MAX_HIDDEN + 1];
};
struct UnpackArgs {
  float* W[TPX_MAX_HIDDEN + 1];
  float* b[TPX_MAX_HIDDEN + 1];
  int in_w[TPX_MAX_HIDDEN + 1];
  int out_w[TPX_MAX_HIDDEN + 1];
};

// one thread per packed float; zero fill for the padding
__global__ void pack_kernel(Layout lay, PackArgs pa, float* __restrict__ packed) {
  const size_t total = lay.total();
  const int HP = lay.HP;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (e < lay.b0()) {  // W0t[c][j] = W0[j][c]
      const int c = (int)(e / HP), j = (int)(e % HP);
      if (j < pa.out_w[0]) v = pa.W[0][(size_t)j * pa.in_w[0] + c];
    } else if (e < lay.hidden_base(1)) {
      const int j = (int)(e - lay.b0());
      if (j < pa.out_w[0]) v = pa.b[0][j];
    } else if (e < lay.hidden_base(lay.L)) {
      const size_t per = 2 * (size_t)HP * HP + HP;
      const size_t r = e - lay.hidden_base(1);
      const int l = 1 + (int)(r / per);
      const size_t q = r % per;
      if (q < (size_t)HP * HP) {  // Wt[i][j] = W[j][i]
        const int i = (int)(q / HP), j = (int)(q % HP);
        if (j < pa.out_w[l] && i < pa.in_w[l]) v = pa.W[l][(size_t)j * pa.in_w[l] + i];
      } else if (q < 2 * (size_t)HP * HP) {  // W[j][i]
        const size_t q2 = q - (size_t)HP * HP;
        const int j = (int)(q2 / HP), i = (int)(q2 % HP);
        if (j < pa.out_w[l] && i < pa.in_w[l]) v = pa.W[l][(size_t)j * pa.in_w[l] + i];
      } else {
        const int j = (int)(q - 2 * (size_t)HP * HP);
        if (j < pa.out_w[l]) v = pa.b[l][j];
      }
    } else if (e < lay.wl()) {               // wide_out only: WLt[i][j] = WL[j][i], rows of dop() columns
      const size_t q = e - lay.hidden_base(lay.L);
      const int i = (int)(q / lay.dop()), j = (int)(q % lay.dop());
      if (j < lay.d_out && i < pa.in_w[lay.L]) v = pa.W[lay.L][(size_t)j * pa.in_w[lay.L] + i];
    } else if (e < lay.bl()) {
      const size_t q = e - lay.wl();
      const int o = (int)(q / HP), i = (int)(q % HP);
      if (o < lay.d_out && i < pa.in_w[lay.L]) v = pa.W[lay.L][(size_t)o * pa.in_w[lay.L] + i];
    } else {
      const int o = (int)(e - lay.bl());
      if (o < lay.d_out) v = pa.b[lay.L][o];
    }
    packed[e] = v;
  }
}

// one thread per real parameter element
__global__ void unpack_kernel(Layout lay, UnpackArgs ua, const double* __restrict__ gacc, int accumulate) {
  const int HP = lay.HP;
  for (int l = 0; l <= lay.L; ++l) {
    const int ow = ua.out_w[l], iw = ua.in_w[l];
    const size_t nW = (size_t)ow * iw;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nW + ow; e += (size_t)gridDim.x * blockDim.x) {
      if (e < nW) {
        if (!ua.W[l]) continue;
        const int j = (int)(e / iw), i = (int)(e % iw);
        size_t src;
        if (l == 0) src = lay.w0() + (size_t)i * HP + j;
        else if (l == lay.L) src = lay.wl() + (size_t)j * HP + i;
        else src = lay.w(l) + (size_t)j * HP + i;
        const float g = (float)gacc[src];
        ua.W[l][e] = accumulate ? ua.W[l][e] + g : g;
      } else {
        if (!ua.b[l]) continue;
        const int j = (int)(e - nW);
        size_t src;
        if (l == 0) src = lay.b0() + j;
        else if (l == lay.L) src = lay.bl() + j;
        else src = lay.b(l) + j;
        const float g = (float)gacc[src];
        ua.b[l][j] = accumulate ? ua.b[l][j] + g : g;
      }
    }
  }
}

// the packed buffer is [FP32-kernel image | tensor-core image]; the second starts 1 KB aligned
static size_t tc_image_offset(const Layout& lay) { return (lay.total() + 255) / 256 * 256; }
// ... followed by the image of the points-as-lanes kernels (1 KB aligned as well)
static size_t tcp_image_offset(const tpx_fcn_desc* net, const Layout& lay) {
  return tc_image_offset(lay) + (tc_net_supported(net) ? (tc_packed_floats(net) + 255) / 256 * 256 : 0);
}

static void fill_dims(const tpx_fcn_desc* net, int* in_w, int* out_w) {
  for (int l = 0; l <= net->n_hidden; ++l) {
    in_w[l] = (l == 0) ? net->d_in : net->widths[l - 1];
    out_w[l] = (l == net->n_hidden) ? net->d_out : net->widths[l];
  }
}

static int sm_count() {           // of the CURRENT device (cached per device: a process may drive several)
  static int cached[64] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

int cuda_fail(const char* what) {
  cudaError_t e = cudaGetLastError();
  return fail(TPX_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

}  // namespace tpx

using namespace tpx;

extern "C" {

int tpx_version(void) { return TPX_VERSION; }
const char* tpx_last_error(void) { return g_err; }
// development hook (not in tpx.h): device buffer that receives phase timestamps of one tile
void tpx_debug_set_timestamps(long long* device_buffer) { tpx::g_tc_debug = device_buffer; }

int tpx_fcn_packed_floats(const tpx_fcn_desc* net, size_t* n_floats) {
  Layout lay;
  int rc = make_layout(net, &lay);
  if (rc) return rc;
  if (!n_floats) return fail(TPX_E_ARG, "n_floats is NULL");
  *n_floats = tcp_image_offset(net, lay) + (tcp_net_supported(net) ? tcp_packed_floats(net) : 0);
  return 0;
}

int tpx_fcn_pack(const tpx_fcn_desc* net, const void* const* params_host, float* packed, void* stream) {
  Layout lay;
  int rc = make_layout(net, &lay);
  if (rc) return rc;
  if (!params_host || !packed) return fail(TPX_E_ARG, "NULL pointer");
  PackArgs pa;
  memset(&pa, 0, sizeof(pa));
  fill_dims(net, pa.in_w, pa.out_w);
  for (int l = 0; l <= net->n_hidden; ++l) {
    pa.W[l] = (const float*)params_host[2 * l];
    pa.b[l] = (const float*)params_host[2 * l + 1];
    if (!pa.W[l] || !pa.b[l]) return fail(TPX_E_ARG, "parameter %d is NULL", l);
  }
  const size_t total = lay.total();
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lay, pa, packed);
  if (cudaGetLastError() != cudaSuccess) return cuda_fail("pack_kernel");
  if (tc_net_supported(net)) {
    if (tc_pack(net, params_host, packed + tc_image_offset(lay), (cudaStream_t)stream)) return cuda_fail("tc_pack");
  }
  if (tcp_net_supported(net)) {
    if (tcp_pack(net, params_host, packed + tcp_image_offset(net, lay), (cudaStream_t)stream)) return cuda_fail("tcp_pack");
  }
  return 0;
}

static int fill_launch(const tpx_fcn_desc* net, const tpx_jet_desc* jet, LaunchArgs* a) {
  int rc = make_layout(net, &a->lay);
  if (rc) return rc;
  rc = check_jet(net, jet, &a->jet);
  if (rc) return rc;
  a->lap = jet->lap ? 1 : 0;
  a->activation = net->activation;
  a->sm_count = sm_count();
  if (a->sm_count <= 0) return fail(TPX_E_CUDA, "no CUDA device");
  return 0;
}

static int jet_forward(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                       const float* x, int64_t n, float* u, float* du, float* lap, float* saved, void* stream) {
  LaunchArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_launch(net, jet, &a);
  if (rc) return rc;
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!packed || !x || !u) return fail(TPX_E_ARG, "NULL pointer");
  if (jet->nd > 0 && !du) return fail(TPX_E_ARG, "du is NULL but nd > 0");
  if (jet->lap && !lap) return fail(TPX_E_ARG, "lap is NULL but lap requested");
  a.packed = packed; a.x = x; a.n = n; a.u = u; a.du = du; a.lapo = lap;
  a.stream = (cudaStream_t)stream;
  a.jet.out_ld = n;
  if (tcp_supported(net, jet->nd, a.lap) && (!saved || tce_bwd_supported(net, jet->nd, a.lap))) {
    float *ck = nullptr, *st = nullptr, *dw = nullptr;
    if (saved) tce_saved_split(net, jet->nd, a.lap, (n + 127) / 128, saved, &ck, &st, &dw);
    return tcp_forward(net, a.jet, a.lap, packed + tcp_image_offset(net, a.lay), x, n, u, du, lap, ck, st, a.sm_count, a.stream);
  }
  if (tc_supported(net, jet->nd, a.lap)) {
    rc = tc_forward(net, a.jet, a.lap, packed + tc_image_offset(a.lay), x, n, u, du, lap, saved, a.sm_count, a.stream);
    if (rc) return cuda_fail("tc_fwd_kernel");
    return 0;
  }
  if (saved && tcw_supported(net, jet->nd, a.lap))     // the launcher recorded the message
    return (net->activation == TPX_ACT_SIN ? tcws_forward : tcw_forward)(a.lay, a.jet, a.lap, packed, x, n, u, du, lap, saved, a.sm_count, a.stream);
  if (a.lay.d_out > 256 || a.jet.out_major)
    return fail(TPX_E_UNSUPPORTED, "d_out=%d > 256 / output-major streams run on the wide tensor-core path only: call tpx_fcn_jet_forward_ws or _save", a.lay.d_out);
  switch (a.lay.HP) {
    case 32: rc = launch_fwd<32>(a); break;
    case 64: rc = launch_fwd<64>(a); break;
    case 128: rc = launch_fwd<128>(a); break;
    case 256: rc = launch_fwd<256>(a); break;
    default: rc = TPX_E_UNSUPPORTED;
  }
  if (rc == TPX_E_UNSUPPORTED) return fail(rc, "no forward kernel for HP=%d nd=%d lap=%d", a.lay.HP, jet->nd, jet->lap);
  if (rc) return cuda_fail("fcn_fwd_kernel");
  return 0;
}

int tpx_fcn_jet_forward(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                        const float* x, int64_t n, float* u, float* du, float* lap, void* stream) {
  return jet_forward(net, jet, packed, x, n, u, du, lap, nullptr, stream);
}

// points per chunk of the wide tensor-core path when its activation arrays live in a caller workspace
static const int64_t kWideChunkPoints = 262144;
static int64_t wide_chunk_points(const tpx_fcn_desc* net, int nd, int lap, size_t workspace_bytes) {
  const size_t fixed = tcw_saved_bytes(net, nd, lap, 0);          // per-CTA partial sums of the reverse kernel
  const size_t per16 = tcw_saved_bytes(net, nd, lap, 16) - fixed; // bytes of one 16-point tile
  if (workspace_bytes <= fixed) return 0;
  return (int64_t)((workspace_bytes - fixed) / per16) * 16;
}

int tpx_fcn_jet_forward_ws(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                           const float* x, int64_t n, float* u, float* du, float* lap, void* workspace,
                           size_t workspace_bytes, void* stream) {
  Layout lay;
  int rc = make_layout(net, &lay);
  if (rc) return rc;
  JetArgs ja;
  rc = check_jet(net, jet, &ja);
  if (rc) return rc;
  const int lp = jet->lap ? 1 : 0;
  if (tcp_supported(net, jet->nd, lp) || tc_supported(net, jet->nd, lp) || !tcw_supported(net, jet->nd, lp) || n <= 0)
    return jet_forward(net, jet, packed, x, n, u, du, lap, nullptr, stream);
  if (!packed || !x || !u || !workspace) return fail(TPX_E_ARG, "NULL pointer");
  if (jet->nd > 0 && !du) return fail(TPX_E_ARG, "du is NULL but nd > 0");
  if (jet->lap && !lap) return fail(TPX_E_ARG, "lap is NULL but lap requested");
  const int64_t cpts = wide_chunk_points(net, jet->nd, lp, workspace_bytes);
  if (cpts < 16) return fail(TPX_E_WORKSPACE, "workspace of %zu bytes holds no 16-point tile of the wide path", workspace_bytes);
  const int sms = sm_count();
  if (sms <= 0) return fail(TPX_E_CUDA, "no CUDA device");
  const int d_in = net->d_in, d_out = net->d_out, nd = jet->nd;
  ja.out_ld = n;
  for (int64_t p0 = 0; p0 < n; p0 += cpts) {
    const int64_t m = (n - p0 < cpts) ? n - p0 : cpts;
    const int64_t ou = ja.out_major ? p0 : p0 * d_out, odu = ja.out_major ? p0 : p0 * d_out * nd;   // output-major: a column range
    rc = (net->activation == TPX_ACT_SIN ? tcws_forward : tcw_forward)(lay, ja, lp, packed, x + p0 * d_in, m, u + ou, du ? du + odu : nullptr,
                     lap ? lap + ou : nullptr, (float*)workspace, sms, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int tpx_fcn_saved_bytes(const tpx_fcn_desc* net, const tpx_jet_desc* jet, int64_t n, size_t* bytes) {
  Layout lay;
  int rc = make_layout(net, &lay);
  if (rc) return rc;
  JetArgs ja;
  rc = check_jet(net, jet, &ja);
  if (rc) return rc;
  if (!bytes || n < 0) return fail(TPX_E_ARG, "bad argument");
  const int lap = jet->lap ? 1 : 0;
  if (tce_bwd_supported(net, jet->nd, lap)) *bytes = tce_saved_bytes(net, jet->nd, lap, n);
  else if (tc_bwd_supported(net, jet->nd, lap)) *bytes = tc_saved_bytes(net, n);
  else if (tcw_supported(net, jet->nd, lap)) *bytes = tcw_saved_bytes(net, jet->nd, lap, n);
  else *bytes = 0;
  return 0;
}

int tpx_fcn_jet_forward_save(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                             const float* x, int64_t n, float* u, float* du, float* lap, void* saved,
                             size_t saved_bytes, void* stream) {
  size_t need = 0;
  int rc = tpx_fcn_saved_bytes(net, jet, n, &need);
  if (rc) return rc;
  if (need == 0) return fail(TPX_E_UNSUPPORTED, "this network / jet has no checkpoint-saving forward kernel");
  if (!saved || saved_bytes < need) return fail(TPX_E_WORKSPACE, "saved buffer of %zu bytes < %zu", saved_bytes, need);
  return jet_forward(net, jet, packed, x, n, u, du, lap, (float*)saved, stream);
}

int tpx_fcn_backward_workspace_bytes(const tpx_fcn_desc* net, const tpx_jet_desc* jet, size_t* bytes) {
  LaunchArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_launch(net, jet, &a);
  if (rc) return rc;
  if (!bytes) return fail(TPX_E_ARG, "bytes is NULL");
  if (a.lay.d_out > 256) {                   // no FP32 kernel: the wide tensor-core path or nothing
    if (!tcw_supported(net, jet->nd, a.lap)) return fail(TPX_E_UNSUPPORTED, "no reverse kernel for d_out=%d", a.lay.d_out);
    *bytes = tcw_saved_bytes(net, jet->nd, a.lap, kWideChunkPoints);
    return 0;
  }
  switch (a.lay.HP) {
    case 32: rc = bwd_workspace<32>(jet->nd, a.lap, a.lay.L, a.lay.d_out, a.sm_count, bytes); break;
    case 64: rc = bwd_workspace<64>(jet->nd, a.lap, a.lay.L, a.lay.d_out, a.sm_count, bytes); break;
    case 128: rc = bwd_workspace<128>(jet->nd, a.lap, a.lay.L, a.lay.d_out, a.sm_count, bytes); break;
    case 256: rc = bwd_workspace<256>(jet->nd, a.lap, a.lay.L, a.lay.d_out, a.sm_count, bytes); break;
    default: rc = TPX_E_UNSUPPORTED;
  }
  if (rc && tcw_supported(net, jet->nd, a.lap)) {     // (e.g. 256-wide with many outputs: too much shared memory for the FP32 reverse kernel)
    *bytes = 0;
    rc = 0;
  }
  if (rc) return fail(rc, "no reverse kernel for HP=%d nd=%d lap=%d d_out=%d", a.lay.HP, jet->nd, jet->lap, a.lay.d_out);
  if (tce_bwd_supported(net, jet->nd, a.lap)) {
    const size_t tcb = tce_saved_bytes(net, jet->nd, a.lap, (int64_t)a.sm_count * 2 * 128);   // two tiles per SM: the chunk stays in L2
    if (tcb > *bytes) *bytes = tcb;
  } else if (tc_bwd_supported(net, jet->nd, a.lap)) {
    const size_t tcb = tc_bwd_workspace_bytes(net, a.sm_count);
    if (tcb > *bytes) *bytes = tcb;
  } else if (tcw_supported(net, jet->nd, a.lap)) {
    const size_t wb = tcw_saved_bytes(net, jet->nd, a.lap, kWideChunkPoints);
    if (wb > *bytes) *bytes = wb;
  }
  return 0;
}

static int jet_backward(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                        const float* x, int64_t n, const float* gu, const float* gdu,
                        const float* glap, double* grad_acc, void* workspace,
                        size_t workspace_bytes, int saved, void* stream) {
  LaunchArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_launch(net, jet, &a);
  if (rc) return rc;
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!packed || !x || !grad_acc || !workspace) return fail(TPX_E_ARG, "NULL pointer");
  a.packed = packed; a.x = x; a.n = n;
  a.gu = gu; a.gdu = jet->nd > 0 ? gdu : nullptr; a.glap = jet->lap ? glap : nullptr;
  a.gacc = grad_acc; a.ckpt = (float*)workspace; a.ckpt_bytes = workspace_bytes;
  a.stream = (cudaStream_t)stream;
  a.jet.out_ld = n;
  if (tce_bwd_supported(net, jet->nd, a.lap))
    return tce_backward(net, a.lay, a.jet, a.lap, packed + tcp_image_offset(net, a.lay), x, n, a.gu, a.gdu, a.glap, grad_acc,
                        (float*)workspace, workspace_bytes, saved, a.sm_count, a.stream);   // the launcher recorded the message
  if (saved && !tc_bwd_supported(net, jet->nd, a.lap)) {
    if (!tcw_supported(net, jet->nd, a.lap)) return fail(TPX_E_UNSUPPORTED, "no reverse kernel that reads saved checkpoints");
    const size_t need = tcw_saved_bytes(net, jet->nd, a.lap, n);
    if (workspace_bytes < need) return fail(TPX_E_WORKSPACE, "saved buffer of %zu bytes < %zu", workspace_bytes, need);
    return (net->activation == TPX_ACT_SIN ? tcws_backward : tcw_backward)(a.lay, a.jet, a.lap, packed, x, n, a.gu, a.gdu, a.glap, grad_acc, (float*)workspace, a.sm_count, a.stream);
  }
  if (tc_bwd_supported(net, jet->nd, a.lap)) {
    rc = tc_backward(net, a.lay, a.jet, a.lap, packed + tc_image_offset(a.lay), x, n, a.gu, a.gdu, a.glap, grad_acc,
                     (float*)workspace, workspace_bytes, saved, a.sm_count, a.stream);
    return rc;                       // the launcher recorded the message
  }
  if (tcw_supported(net, jet->nd, a.lap)) {
    // no saved activations: chunks of (forward sweep writing checkpoints only) + reverse sweep through the workspace;
    // a workspace too small for one tile (a caller that sized it for the FP32 kernels) falls through to those
    const int64_t cpts = wide_chunk_points(net, jet->nd, a.lap, workspace_bytes);
    if (cpts >= 16) {
      const int d_in = net->d_in, d_out = net->d_out, nd = jet->nd;
      for (int64_t p0 = 0; p0 < n; p0 += cpts) {
        const int64_t m = (n - p0 < cpts) ? n - p0 : cpts;
        rc = (net->activation == TPX_ACT_SIN ? tcws_forward : tcw_forward)(a.lay, a.jet, a.lap, packed, x + p0 * d_in, m, nullptr, nullptr, nullptr, (float*)workspace, a.sm_count, a.stream);
        if (rc) return rc;
        const int64_t ou = a.jet.out_major ? p0 : p0 * d_out, odu = a.jet.out_major ? p0 : p0 * d_out * nd;
        rc = (net->activation == TPX_ACT_SIN ? tcws_backward : tcw_backward)(a.lay, a.jet, a.lap, packed, x + p0 * d_in, m, a.gu ? a.gu + ou : nullptr,
                          a.gdu ? a.gdu + odu : nullptr, a.glap ? a.glap + ou : nullptr, grad_acc,
                          (float*)workspace, a.sm_count, a.stream);
        if (rc) return rc;
      }
      return 0;
    }
  }
  if (a.lay.d_out > 256 || a.jet.out_major)
    return fail(TPX_E_UNSUPPORTED, "d_out=%d > 256 / output-major streams run on the wide tensor-core path only (workspace of %zu bytes holds no tile)", a.lay.d_out, workspace_bytes);
  switch (a.lay.HP) {
    case 32: rc = launch_bwd<32>(a); break;
    case 64: rc = launch_bwd<64>(a); break;
    case 128: rc = launch_bwd<128>(a); break;
    case 256: rc = launch_bwd<256>(a); break;
    default: rc = TPX_E_UNSUPPORTED;
  }
  if (rc == TPX_E_UNSUPPORTED) return fail(rc, "no reverse kernel for HP=%d nd=%d lap=%d", a.lay.HP, jet->nd, jet->lap);
  if (rc == TPX_E_WORKSPACE) return fail(rc, "workspace of %zu bytes is too small", workspace_bytes);
  if (rc) return cuda_fail("fcn_bwd_kernel");
  return 0;
}

int tpx_fcn_jet_backward(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                         const float* x, int64_t n, const float* gu, const float* gdu,
                         const float* glap, double* grad_acc, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return jet_backward(net, jet, packed, x, n, gu, gdu, glap, grad_acc, workspace, workspace_bytes, 0, stream);
}

int tpx_fcn_jet_backward_saved(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                               const float* x, int64_t n, const float* gu, const float* gdu,
                               const float* glap, double* grad_acc, const void* saved, size_t saved_bytes,
                               void* stream) {
  return jet_backward(net, jet, packed, x, n, gu, gdu, glap, grad_acc, (void*)saved, saved_bytes, 1, stream);
}

int tpx_fcn_unpack_grads(const tpx_fcn_desc* net, const double* grad_acc, void* const* grads_host,
                         int accumulate, void* stream) {
  Layout lay;
  int rc = make_layout(net, &lay);
  if (rc) return rc;
  if (!grad_acc || !grads_host) return fail(TPX_E_ARG, "NULL pointer");
  UnpackArgs ua;
  memset(&ua, 0, sizeof(ua));
  fill_dims(net, ua.in_w, ua.out_w);
  for (int l = 0; l <= net->n_hidden; ++l) {
    ua.W[l] = (float*)grads_host[2 * l];
    ua.b[l] = (float*)grads_host[2 * l + 1];
  }
  unpack_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(lay, ua, grad_acc, accumulate);
  if (cudaGetLastError() != cudaSuccess) return cuda_fail("unpack_kernel");
  return 0;
}

}  // extern "C"
