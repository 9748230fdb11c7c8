/*
 * tpx.h -- C ABI of libtpx.so, the B200 (sm_100a) engine behind the TorchPhysics API.
 *
 * The reference (boschresearch/torchphysics v1.1.0) is pure Python and has no FFI of
 * its own: its hot path is the chain of Python calls listed next to each entry point
 * below (paths relative to src/torchphysics in the reference).  This library is what a
 * maintainer binds with ctypes *inside* those Python functions.
 *
 * Conventions
 *   - every function returns 0 on success and a negative TPX_E_* code on failure; the
 *     message of the last failure on the calling thread is tpx_last_error();
 *   - all pointers are DEVICE pointers unless the name ends in _host; buffers are owned
 *     by the caller (PyTorch's allocator in the Python binding), nothing is retained;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work;
 *   - float32 row-major everywhere, gradient accumulators are float64;
 *   - re-entrant: no global mutable state except the per-thread error string.
 */
#ifndef TPX_H_
#define TPX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPX_VERSION 100
#define TPX_MAX_HIDDEN 16   /* hidden layers (Linear+activation pairs)            */
#define TPX_MAX_DIN 8       /* columns of the model input                          */
#define TPX_MAX_ND 3        /* derivative directions propagated through the net    */
#define TPX_MAX_WIDTH 256   /* widest hidden layer                                 */
#define TPX_MAX_DOUT 4096   /* outputs; more than 256 only on the layer-major tensor-core kernels (two or more hidden layers) */

#define TPX_E_ARG (-1)       /* invalid argument                      */
#define TPX_E_UNSUPPORTED (-2) /* valid network, but no kernel for it  */
#define TPX_E_CUDA (-3)      /* CUDA runtime error                    */
#define TPX_E_WORKSPACE (-4) /* caller workspace too small            */

#define TPX_ACT_TANH 0 /* nn.Tanh          (models/fcn.py:66)            */
#define TPX_ACT_SIN 1  /* tp.models.Sinus  (models/activation_fn.py:76)  */

/* Fully connected network  x -> W_L s(... s(W_0 x + b_0) ...) + b_L
 * (models/fcn.py:9-29, _construct_FC_layers; models/deeponet/trunknets.py:70-87). */
typedef struct tpx_fcn_desc {
  int32_t d_in;                     /* input columns                         */
  int32_t d_out;                    /* output columns                        */
  int32_t n_hidden;                 /* number of hidden layers, >= 1         */
  int32_t widths[TPX_MAX_HIDDEN];   /* hidden widths                         */
  int32_t activation;               /* TPX_ACT_*  (same for every layer)     */
} tpx_fcn_desc;

/* Which derivative streams ("jet") are propagated next to the value.
 *   nd      directions, 0..3;  dcol[k] = input column of direction k
 *   lap     0/1: also propagate  L u = sum_k lap_w[k] d^2 u / d x_dcol[k]^2
 * This is what tp.utils.grad / laplacian / div / jac / partial / normal_derivative /
 * convective (utils/differentialoperators/differentialoperators.py:11-342) read from. */
typedef struct tpx_jet_desc {
  int32_t nd;
  int32_t dcol[TPX_MAX_ND];
  int32_t lap;
  float lap_w[TPX_MAX_ND];
  int32_t out_major;   /* 0: u (n, d_out), du (n, d_out, nd), lap (n, d_out) -- and their adjoints -- as documented below;
                        * 1: output-major  u (d_out, n), du (nd, d_out, n), lap (d_out, n): the layout of a DeepONet's
                        *    (function, point) outputs when the inner product with the branch coefficients is folded into
                        *    the trunk's output layer (models/deeponet/deeponet.py:91-102).  Only for networks with 9 or
                        *    more outputs on the layer-major tensor-core kernels; TPX_E_UNSUPPORTED otherwise. */
} tpx_jet_desc;

int tpx_version(void);
const char* tpx_last_error(void);

/* Number of floats of the packed (padded, both orientations) parameter image, and of
 * the float64 gradient accumulator that uses the same indexing. */
int tpx_fcn_packed_floats(const tpx_fcn_desc* net, size_t* n_floats);

/* Re-layout the nn.Linear parameters (weight (out,in) row-major, bias) into the packed
 * image the kernels read.  params_host = {W0,b0,W1,b1,...,WL,bL}: a HOST array of
 * 2*(n_hidden+1) device pointers.  One launch.  Replaces nothing in the reference
 * (its nn.Linear modules are used as they are, models/fcn.py:18-28). */
int tpx_fcn_pack(const tpx_fcn_desc* net, const void* const* params_host, float* packed, void* stream);

/* Forward jet.  Replaces FCN.forward (models/fcn.py:85-87) and, for the requested
 * streams, the autograd sweeps of grad/laplacian/div/jac
 * (utils/differentialoperators/differentialoperators.py:34,40,64,160,254).
 *   x   (n, d_in)            u   (n, d_out)
 *   du  (n, d_out, nd)       lap (n, d_out)      (du/lap may be NULL if not requested) */
int tpx_fcn_jet_forward(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                        const float* x, int64_t n, float* u, float* du, float* lap, void* stream);

/* The same forward jet with a caller workspace of at least tpx_fcn_backward_workspace_bytes() bytes: networks of
 * hidden width 65..256 (tanh or sin, and narrower nets that are deeper than
 * 5 hidden layers or have more than 8 outputs or more than 4 inputs) then run on the layer-major tensor-core kernels
 * (csrc/tcw_kernels.cu), whose per-layer activation arrays live in the workspace (chunks of points); every other
 * network ignores the workspace.  tpx_fcn_jet_backward uses its workspace for those networks in the same way. */
int tpx_fcn_jet_forward_ws(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                           const float* x, int64_t n, float* u, float* du, float* lap, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Bytes of scratch tpx_fcn_jet_backward needs (per-CTA activation checkpoints that are
 * written and re-read within one kernel, sized to stay in the 126 MB L2). */
int tpx_fcn_backward_workspace_bytes(const tpx_fcn_desc* net, const tpx_jet_desc* jet, size_t* bytes);

/* Reverse sweep.  Replaces loss.backward() through the graphs built by the calls above
 * (solver.py:100-109 + every create_graph=True sweep).  Recomputes the forward jet per
 * tile of points on chip, then accumulates dLoss/dtheta into the float64 image
 * `grad_acc` (packed indexing; caller zeroes it).
 *   gu (n,d_out), gdu (n,d_out,nd), glap (n,d_out): dLoss/d(outputs); any may be NULL. */
int tpx_fcn_jet_backward(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                         const float* x, int64_t n, const float* gu, const float* gdu,
                         const float* glap, double* grad_acc, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Activation-jet checkpoints.  When a forward evaluation is followed by loss.backward() on the same points
 * (the normal training step: condition.py:293-304 then solver.py:100-109), the forward kernel can save the
 * per-layer jets (8 KB per point for a 4x64 net; HBM is 180 GB) and the reverse kernel reads them instead of
 * re-running the forward sweep.  *bytes = 0 when this network / jet has no saving kernel.  The size includes a
 * fixed tail (<= 13 MB) of per-CTA fp32 partial sums the reverse kernels flush their TMEM dW accumulators into.
 * Hidden widths 65..256 (tanh or sin) run on the layer-major tensor-core kernels
 * (csrc/tcw_kernels.cu); through this pair of calls their `saved` buffer holds the checkpoints of the
 * hidden layers 1..L-1 ([tile of 16 points][stream * 16 + point][128 neurons] fp32 per layer and 128-neuron
 * block: one block up to width 128, two above) followed by two scratch arrays of the same size for the reverse
 * sweep (+ one block of partial products for two blocks): (L + 1) * S * 512 bytes per point up to width 128,
 * (2 L + 3) * S * 512 above; the checkpoints are not consumed (the reverse may be repeated), the scratch part
 * is overwritten.
 * Without a saved buffer those networks run the same kernels in chunks through a caller workspace
 * (tpx_fcn_jet_forward_ws, tpx_fcn_jet_backward); plain tpx_fcn_jet_forward uses the FP32 kernels. */
int tpx_fcn_saved_bytes(const tpx_fcn_desc* net, const tpx_jet_desc* jet, int64_t n, size_t* bytes);
int tpx_fcn_jet_forward_save(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                             const float* x, int64_t n, float* u, float* du, float* lap, void* saved,
                             size_t saved_bytes, void* stream);
int tpx_fcn_jet_backward_saved(const tpx_fcn_desc* net, const tpx_jet_desc* jet, const float* packed,
                               const float* x, int64_t n, const float* gu, const float* gdu,
                               const float* glap, double* grad_acc, const void* saved, size_t saved_bytes,
                               void* stream);

/* Residual-loss reduction of PINNCondition (problem/conditions/condition.py:11-27 SquaredError, :488-489 mean):
 *   *loss_acc += (1 / rows) * sum of r^2 over all `total` elements  (float64 device scalar, caller zeroes it;
 *   warp-shuffle reduction, one atomic per CTA), and its adjoint  grad_r = grad_loss * 2 r / rows
 *   (grad_loss: device scalar, NULL = 1).  r is the (rows, ...) residual batch, contiguous fp32. */
int tpx_squared_error_mean(const float* r, int64_t total, int64_t rows, double* loss_acc, void* stream);
int tpx_squared_error_mean_backward(const float* r, int64_t total, int64_t rows, const float* grad_loss, float* grad_r,
                                    void* stream);

/* float64 packed accumulator -> per-parameter float32 gradients (layout of nn.Linear).
 * grads_host: HOST array of 2*(n_hidden+1) device pointers (entries may be NULL).
 * accumulate != 0 adds to the existing gradient (PyTorch .grad semantics). */
int tpx_fcn_unpack_grads(const tpx_fcn_desc* net, const double* grad_acc, void* const* grads_host,
                         int accumulate, void* stream);

/* ------------------------------------------------------------------------------------
 * Collocation-point samplers (problem/domains + problem/samplers of the reference).
 * A shape is one of the reference's basic domains with CONSTANT parameters; `col` is the
 * first column of the point row it occupies (product domains put shapes side by side,
 * domainoperations/product.py:233-267).  A region is a boolean combination of shapes as a
 * postfix program: Cut = A B NOT AND (cut.py:39-42), Union = A B OR (union.py:45-48),
 * Intersection = A B AND, Product = A B AND on different columns (product.py:108-111).
 * ------------------------------------------------------------------------------------ */
#define TPX_SHAPE_INTERVAL 1      /* p = {lower, upper}                  domain1D/interval.py       */
#define TPX_SHAPE_PARALLELOGRAM 2 /* p = {ox, oy, c1x, c1y, c2x, c2y}    domain2D/parallelogram.py  */
#define TPX_SHAPE_CIRCLE 3        /* p = {cx, cy, r}                     domain2D/circle.py         */
/* boundaries of the three primitives (membership = the reference's torch.isclose tests) */
#define TPX_SHAPE_INTERVAL_BOUNDARY 4      /* p = {lower, upper}          interval.py:96-149        */
#define TPX_SHAPE_INTERVAL_POINT 5         /* p = {side, normal}          interval.py:152-191       */
#define TPX_SHAPE_PARALLELOGRAM_BOUNDARY 6 /* p as the parallelogram      parallelogram.py:159-291  */
#define TPX_SHAPE_CIRCLE_BOUNDARY 7        /* p as the circle             circle.py:111-175         */
#define TPX_MAX_SHAPES 8
#define TPX_MAX_OPS 24
#define TPX_OP_SHAPE 0 /* push membership of shapes[arg]; op word = opcode | arg << 8 */
#define TPX_OP_AND 1
#define TPX_OP_OR 2
#define TPX_OP_NOT 3

typedef struct tpx_shape {
  int32_t kind;
  int32_t col;
  float p[6];
} tpx_shape;

typedef struct tpx_region {
  int32_t n_shapes;
  int32_t n_ops;
  tpx_shape shapes[TPX_MAX_SHAPES];
  int32_t ops[TPX_MAX_OPS];
} tpx_region;

/* mask[i] = 1 if row i of `points` (n rows of row_stride floats) lies in the region.
 * Replaces Domain._contains (interval.py:37-43, parallelogram.py:83-103, circle.py:34-40,
 * cut.py:39-42, union.py:45-48) with the same float32 operation order. */
int tpx_region_contains(const tpx_region* region, const float* points, int64_t n, int32_t row_stride,
                        uint8_t* mask, void* stream);

/* n uniform points of one shape written to columns [col, col+dim) of `out`; point i is a
 * pure function of (seed, offset + i) (Philox4x32-10).  Replaces sample_random_uniform of
 * interval.py:45-55, parallelogram.py:105-117, circle.py:52-68. */
int tpx_sample_uniform(const tpx_shape* shape, uint64_t seed, uint64_t offset, int64_t n, float* out,
                       int32_t row_stride, void* stream);

/* The same with the Philox key in DEVICE memory, for CUDA-graph capture of a whole training step (a replayed launch cannot
 * receive a fresh key as an argument): tpx_key_advance replaces *key_dev by splitmix64(*key_dev) (one thread), the sampling
 * launch that follows reads it.  Replaces the per-call host draw of sampler_base.py:130 / the domains' torch.rand calls. */
int tpx_key_advance(uint64_t* key_dev, void* stream);
int tpx_sample_uniform_dkey(const tpx_shape* shape, const uint64_t* key_dev, uint64_t offset, int64_t n, float* out,
                            int32_t row_stride, void* stream);

/* Rejection sampling without a host round trip (replaces the host-synchronising loop of
 * domainoperations/sampler_helper.py:50-76 and the filter loop of
 * samplers/random_samplers.py:60-90): n_candidates points of `generator` are tested
 * against `accept`; the accepted ones are appended IN CANDIDATE ORDER (stable warp-ballot
 * compaction) to `out` starting at row counters[0], never beyond row n_wanted.
 * counters (device, int64[4], caller zeroes before the first round):
 *   [0] rows filled so far (<= n_wanted)   [1] candidates offered so far
 *   [2] scratch                            [3] candidates accepted so far (unclamped)   */
int tpx_sample_reject_workspace_bytes(int64_t n_candidates, size_t* bytes);
int tpx_sample_uniform_reject(const tpx_shape* generator, const tpx_region* accept, uint64_t seed,
                              uint64_t offset, int64_t n_candidates, int64_t n_wanted, float* out,
                              int32_t row_stride, int64_t* counters, void* workspace,
                              size_t workspace_bytes, void* stream);

/* Deterministic grids: Interval interior linspace (interval.py:57-65), Parallelogram
 * barycentric n1 x n2 grid (parallelogram.py:119-152; n = n1*n2), Circle sunflower
 * arrangement (circle.py:70-99). */
int tpx_sample_grid(const tpx_shape* shape, int64_t n, int32_t n1, int32_t n2, float* out, int32_t row_stride,
                    void* stream);

/* Outward unit normals of a boundary shape at n points lying on it (rows of row_stride floats, the
 * shape's columns); normals = n x dim floats.  Replaces BoundaryDomain.normal of interval.py:139-144,
 * 181-186, parallelogram.py:248-286, circle.py:162-170 with the same float32 operation order (a
 * parallelogram point that fails the reference's isclose tests gets the reference's NaN). */
int tpx_boundary_normal(const tpx_shape* shape, const float* points, int64_t n, int32_t row_stride,
                        float* normals, void* stream);

/* Latin hypercube (samplers/random_samplers.py:203-221): keys = iid 63-bit sort keys (their
 * argsort is the random permutation of one axis); tpx_sample_lhs writes column c of row i
 * as box[2c] + len_c * perm[c*n+i]/n + len_c/n * U(0,1). */
int tpx_sample_keys(uint64_t seed, uint64_t offset, int64_t n, int64_t* keys, void* stream);
int tpx_sample_lhs(int32_t d, const float* box, const int64_t* perm, uint64_t seed, uint64_t offset, int64_t n,
                   float* out, int32_t row_stride, int32_t col, void* stream);

/* Union pick (domainoperations/union.py:72-93): out[i] = a[i] if b[i] lies in region in_a or
 * U(0,1) <= ratio (= vol_A / (vol_A + vol_B)), else b[i]. */
int tpx_union_pick(const tpx_region* in_a, const float* a, const float* b, int64_t n, int32_t d, int32_t row_stride,
                   float ratio, uint64_t seed, uint64_t offset, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Parameter update (reference: torch.optim.Adam configured by OptimizerSetting, solver.py:111-136, stepped by
 * Lightning after Solver.training_step): one launch for all parameter tensors, same arithmetic as
 * torch.optim.Adam (amsgrad = False).  Tensors are fp32, contiguous; a NULL grad skips the tensor (torch skips
 * parameters whose .grad is None).  `step` is the 1-based step count of the bias correction.
 * ------------------------------------------------------------------------------------ */
#define TPX_MAX_TENSORS 64
typedef struct tpx_tensor_list {
  int32_t n;
  int32_t reserved;
  void* param[TPX_MAX_TENSORS];
  const void* grad[TPX_MAX_TENSORS];
  void* exp_avg[TPX_MAX_TENSORS];
  void* exp_avg_sq[TPX_MAX_TENSORS];
  int64_t numel[TPX_MAX_TENSORS];
} tpx_tensor_list;

int tpx_adam_step(const tpx_tensor_list* list, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int32_t step, void* stream);

/* The same update with the step count in device memory (CUDA-graph capture): *step_dev is incremented, then used for the
 * bias corrections on the device.  Replaces torch.optim.Adam(capturable=True) behind solver.py:111-136. */
int tpx_adam_step_dstep(const tpx_tensor_list* list, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int32_t* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TPX_H_ */
