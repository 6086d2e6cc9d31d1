/* onetrans_b200.h — C-ABI of libonetrans_sm100.so: the B200 (sm_100a) kernels behind the OneTrans
 * ranking block of ScottHCL/recommend (rank/scaling_up/oneTrans/practice, "OT/" below).
 *
 * The reference has no FFI or plugin layer: its hot path is a Keras Model whose arithmetic is
 * TensorFlow library calls.  Each entry point below therefore replaces a group of
 * TensorFlow call sites of OT/model.py (cited per function); the host side that keeps the reference's
 * module API (recommend_b200/model.py) binds them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never frees caller memory, never synchronises the
 *     host and allocates no device memory on the hot path (its one allocation: a 4 KB ring of work counters per device
 *     for the persistent kernels' dynamic tile schedule, made on first use and kept for the life of the process);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - activations are bf16, row-major `[rows, cols]` with an element leading dimension (`ld*`);
 *     rows are TOKEN-MAJOR: row = token_position * B + sample  (DESIGN.md §3);
 *   - return value 0 = ok; negative = error (ot_last_error_string() describes it).  An unsupported
 *     shape is an error, never a fallback.
 */
#ifndef ONETRANS_B200_H_
#define ONETRANS_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OT_ABI_VERSION 14

int ot_version(void);
const char* ot_last_error_string(void);
/* SM count of the current device (0 if no device). */
int ot_num_sms(void);

/* ------------------------------------------------------------------------------------------------
 * Mixed-parameter GEMM  (tcgen05 / TMEM / TMA, persistent, grouped)
 *   out[row, :] = epilogue( A[row, :K] . W[group(row)]^T )           W[g] is [N, K] (K contiguous)
 * Replaces the per-position Dense loops of MixedMHA (OT/model.py:84-92), Wo (:117), MixedFFN
 * (:154-163), the sequence projections of the Tokenizer (:262-265) and, with transposed weights, their
 * input gradients.  "group(row)" is data: rows are described by up to 3 segments; a segment is
 * `n_units` runs of `rows_per_unit` rows; unit u uses weight group `group_start + u*group_stride`.
 *   shared S-token run : n_units=1, rows_per_unit=n_S*B, group_stride=0
 *   NS-token run       : n_units=n_NS, rows_per_unit=B,  group_stride=1     (OT/model.py:67-74, D4)
 * ---------------------------------------------------------------------------------------------- */
typedef struct ot_gemm_seg {
  int32_t row_start;     /* first output row of unit 0 */
  int32_t n_units;
  int32_t rows_per_unit;
  int32_t group_start;
  int32_t group_stride;
  int32_t a_row_start;   /* A row of unit 0 row 0 (2-D mode) or first unit coordinate (transposed mode) */
} ot_gemm_seg;

enum {
  OT_EPI_BIAS = 1,        /* + bias[group*bias_group_stride + n]                                   */
  OT_EPI_GELU = 2,        /* out = gelu_erf(v); if out2 != NULL, out2 = v (pre-activation)          */
  OT_EPI_RESIDUAL = 4,    /* + res[row, n]                                                          */
  OT_EPI_GELU_GRAD = 8,   /* v *= gelu_erf'(aux[row, n])                                            */
  OT_EPI_ROW_SCALE = 16,  /* v *= row_scale[row]   (applied first)                                  */
  OT_EPI_NORM = 64,       /* second output norm_out = rmsnorm(out row) * norm_gain (RMSNorm.call, OT/model.py:19-23, fused into
                             the producing GEMM); needs N == block_n <= 256 so that a tile holds whole rows               */
  OT_EPI_DROPOUT = 32     /* v = keep(row,n) ? v/(1-rate) : 0, after bias and before the residual add
                             (Keras inverted dropout on the branch output, OT/model.py:193,198)        */
};

typedef struct ot_gemm_params {
  /* A operand.  2-D mode: element (row,k) at A[row*a_stride1 + k], a_dim1 rows.
   * transposed mode (a_transposed=1): element (unit,row_in_unit,k) at
   * A[row_in_unit*a_stride2 + unit*a_stride1 + k], a_dim1 units, a_dim2 rows per unit
   * (the tokenizer reads `[B, L_i, 64]` events token-major this way). Strides in elements. */
  const void* A;
  int64_t a_dim1, a_dim2, a_stride1, a_stride2;
  int32_t a_transposed;
  int32_t n_groups;
  const void* W;         /* [n_groups, N, K] bf16, K contiguous, leading dimension ldw */
  int64_t ldw;
  int32_t N, K;
  int32_t n_segs;
  int32_t flags;
  ot_gemm_seg segs[3];
  void* out;   int64_t ldo;
  void* out2;  int64_t ldo2;
  const void* res; int64_t ldr;
  const void* aux; int64_t ldaux;
  const float* bias; int64_t bias_group_stride;
  const float* row_scale;
  int32_t block_n;       /* 0 = choose */
  int32_t swizzle;       /* 0 = 128-byte swizzle (default); 64 selects the 64-byte variant */
  /* High-precision residual stream of the NS-token rows: output rows >= hp_row0 (must be the
   * row_start of an NS segment) take their residual from res_hp (fp32, row 0 <-> output row hp_row0) instead
   * of `res`, and also write the fp32 result to out_hp.  Only with OT_EPI_RESIDUAL.  NULL = off. */
  const float* res_hp;
  float* out_hp;
  int64_t ld_hp;
  int64_t hp_row0;
  /* OT_EPI_DROPOUT: counter-based mask hash(seed, row, n); rate in [0, 1). */
  uint32_t drop_seed;
  float drop_rate;
  /* OT_EPI_NORM: norm_out[row, n] = x[row, n] * rstd[row] * norm_gain[n] with rstd = rsqrt(mean_n(x^2) + norm_eps), where x
   * is the row just written to `out` (the fp32 value on res_hp rows); rstd is stored to norm_rstd (fp32 [rows], may be
   * NULL).  Not together with out2. */
  void* norm_out; int64_t ld_norm;
  const float* norm_gain;
  float* norm_rstd;
  float norm_eps;
} ot_gemm_params;

int ot_mixed_gemm(const ot_gemm_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused MixedFFN forward  (one kernel for OT/model.py:149-163 + the residual / dropout of :198 + the next RMSNorm)
 *   y[row, :] = res[row, :] + drop( gelu_erf(zn[row, :] . W1[g] + b1[g]) . W2[g] + b2[g] ),   g = group(row)
 * The hidden activation [rows, F] stays on chip; `pre` (optional, bf16 [rows, F]) receives zn.W1 + b1 for the backward pass.
 * W1: [n_groups, F, d] (N = F, K = d), W2: [n_groups, d, F] (N = d, K = F), both bf16 with K contiguous (the same transposed
 * compute copies ot_mixed_gemm takes).  Segments as in ot_gemm_params (a_row_start must equal row_start).
 * flags: OT_EPI_RESIDUAL | OT_EPI_DROPOUT | OT_EPI_NORM (bias is always applied); res_hp / out_hp / hp_row0, drop_*, norm_* as
 * in ot_gemm_params.  Built for d == 256 and F % 128 == 0; any other shape is OT_ERR_UNSUPPORTED_SHAPE (callers use two
 * ot_mixed_gemm launches there). */
typedef struct ot_ffn_params {
  const void* zn; int64_t ldzn;
  const void* W1; int64_t ldw1;
  const void* W2; int64_t ldw2;
  const float* b1; int64_t b1_group_stride;
  const float* b2; int64_t b2_group_stride;
  int32_t n_groups, d, F;
  int32_t n_segs;
  int32_t flags;
  ot_gemm_seg segs[3];
  void* out; int64_t ldo;
  void* pre; int64_t ldpre;          /* may be NULL (evaluation) */
  void* h; int64_t ldh;              /* optional (NULL = not stored): bf16 [rows, F] copy of gelu_erf(pre) for the dW2 weight gradient */
  const void* res; int64_t ldr;
  const float* res_hp; float* out_hp; int64_t ld_hp; int64_t hp_row0;
  uint32_t drop_seed; float drop_rate;
  void* norm_out; int64_t ld_norm;
  const float* norm_gain;
  float* norm_rstd;
  float norm_eps;
} ot_ffn_params;
int ot_ffn_fwd(const ot_ffn_params* p, void* stream);
/* Input-gradient chain of the same layer in one kernel (tape.gradient of OT/model.py:149-163 w.r.t. the FFN input, OT/train.py:131):
 *   dpre = (dy . W2[g]^T) o gelu_erf'(pre)        dzn = dpre . W1[g]^T
 * with the SAME struct read as:  zn = dy [rows, d] (input),  W1 = the second Dense's kernel as [n_groups, F, d] (N = F, K = d: the
 * master layout of W2),  W2 = the first Dense's kernel as [n_groups, d, F] (the master layout of W1),  pre = saved pre-activation
 * (INPUT),  h = dpre [rows, F] (OUTPUT, required: the dW1 weight gradient reads it),  out = dzn [rows, d].  b1 / b2 are ignored,
 * flags must be 0.  dpre lives on chip between the two products and is stored from the kernel's own operand tile. */
int ot_ffn_bwd(const ot_ffn_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight gradient  (tcgen05, both operands MN-major, split over rows, fp32 atomic accumulate)
 *   C[group][m, n] += sum_rows P[row, m] * Q[row, n]
 * Replaces tape.gradient (OT/train.py:131) for every Dense kernel on the path.
 * A segment is `n_units` runs of `rows_per_unit` rows.  group_stride=1: unit u accumulates into
 * group group_start+u; group_stride=0: all units accumulate into group_start.
 * P/Q element (unit, row, col) at base[unit*stride_unit + row*stride_row + col]  (elements).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ot_wgrad_seg {
  const void* P; int64_t p_stride_row, p_stride_unit;
  const void* Q; int64_t q_stride_row, q_stride_unit;
  int32_t n_units;
  int32_t rows_per_unit;
  int32_t group_start;
  int32_t group_stride;
} ot_wgrad_seg;

typedef struct ot_wgrad_params {
  int32_t Mdim, Ndim;        /* Mdim % 128 == 0, Ndim % 64 == 0 */
  int32_t n_segs;
  int32_t swizzle;           /* 0 = 128 */
  ot_wgrad_seg segs[2];
  float* C;                  /* fp32, accumulated with atomics (caller zeroes) */
  int64_t c_group_stride, c_stride_m, c_stride_n;
  const float* p_row_scale;  /* optional: P rows scaled... reserved, must be NULL */
  int32_t block_n;           /* 0 = choose */
  int32_t target_ctas;       /* 0 = 2 waves of the device */
  /* optional bias gradient riding on the same pass (OT/train.py:131 for the Dense biases of OT/model.py:137-145):
   * q_colsum[g * q_colsum_group_stride + n] += sum_rows Q[row, n], fp32 atomics, caller zeroes; NULL = off */
  float* q_colsum;
  int64_t q_colsum_group_stride;
  /* p_gelu != 0: the kernel contracts gelu_erf(P) instead of P (P = the saved FFN pre-activation: dW2 = gelu(pre)^T dy after
   * ot_ffn_fwd, which keeps the hidden activation on chip).  block_n 256 / 128-byte swizzle only. */
  int32_t p_gelu;
} ot_wgrad_params;

int ot_wgrad(const ot_wgrad_params* p, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Causal attention with pyramid query pruning  (tcgen05 flash-style; forward and backward)
 * Replaces einsum/where/softmax/einsum of MixedMHA.call (OT/model.py:101-114) for the retained query
 * tail only (OT/model.py:356-371 computes every query and gathers; same values).
 *   queries: the last Lq positions; keys/values: all Lk positions (Lq <= Lk);
 *   query i attends keys 0 .. (Lk-Lq)+i; scores scaled by 1/sqrt(head_dim) (OT/model.py:106).
 * Buffers are token-major: element (l, b, h, e) at base[(l*B + b)*ld + h*head_dim + e].
 * lse / delta: fp32 [B, H, Lq].  ot_attn_bwd needs q,k,v,o,lse,d_o and writes dq,dk,dv (+ delta scratch).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ot_attn_params {
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  void* o;       int64_t ldo;
  float* lse;
  const void* d_o; int64_t lddo;
  void* dq; int64_t lddq;
  void* dk; int64_t lddk;
  void* dv; int64_t lddv;
  float* delta;
  int32_t B, H, Lq, Lk;
  int32_t head_dim;      /* 32, 64 or 96 */
  int32_t swizzle;       /* 0 = default; 64 forces the 64-byte-swizzle variant for head_dim 64 */
} ot_attn_params;

int ot_attn_fwd(const ot_attn_params* p, void* stream);
int ot_attn_bwd(const ot_attn_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Inference attention with a cross-candidate cache of the sequence-side K/V (PAPER:144-151; the reference's
 * own cache branch OT/model.py:95-98 is unusable).  One user, C candidates:
 *   shared keys/values : [Ls, ld_shared] rows of the S tokens, identical for every candidate (B = 1 layout)
 *   own keys/values    : [Tn*C, ld_own]  rows (token-major: row = t*C + c) of the Tn surviving NS tokens
 *   queries            : [Tq*C, ldq]     the last Tq NS tokens;  output o [Tq*C, ldo]
 * Query token i attends every shared key and own keys 0 .. (Tn-Tq)+i of ITS candidate.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ot_attn_cached_params {
  const void* q; int64_t ldq;
  const void* k_own; int64_t ld_own_k;
  const void* v_own; int64_t ld_own_v;
  const void* k_shared; int64_t ld_shared_k;
  const void* v_shared; int64_t ld_shared_v;
  void* o; int64_t ldo;
  int32_t C, H, Tq, Tn, Ls;
  int32_t head_dim;
} ot_attn_cached_params;

int ot_attn_ns_cached_fwd(const ot_attn_cached_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * RMSNorm (OT/model.py:19-23):  y = x * rsqrt(mean(x^2,-1) + eps) * gain.   HBM-bound, one warp per row.
 * forward : x, gain -> y, rstd (fp32 per row; may be NULL)
 * backward: dy, x, rstd, gain -> dx (+= dres if dres != NULL), dgain += sum_rows dy * x * rstd (fp32 atomics)
 * ---------------------------------------------------------------------------------------------- */
typedef struct ot_rmsnorm_params {
  const void* x;  int64_t ldx;
  void* y;        int64_t ldy;
  const float* gain;
  float* rstd;
  const void* dy;   int64_t lddy;
  const void* dres; int64_t lddres;
  void* dx;         int64_t lddx;
  float* dgain;
  int64_t rows;
  int32_t d;
  float eps;
  /* forward only: rows >= hp_row0 are read from x_hp (fp32 [rows-hp_row0, d]) instead of x.  NULL = off. */
  const float* x_hp;
  int64_t hp_row0;
  /* backward only: second output dx_drop[r, c] = keep(drop_row0 + r, c) ? dx[r, c] / (1 - rate) : 0 with the mask of
   * OT_EPI_DROPOUT / ot_dropout_mask (same seed, index space [*, d]): the gradient entering the dropped-out branch that
   * sits below this norm (OT/model.py:193,198), written in the same pass.  NULL = off. */
  void* dx_drop; int64_t lddx_drop;
  int64_t drop_row0;
  uint32_t drop_seed;
  float drop_rate;
} ot_rmsnorm_params;

int ot_rmsnorm_fwd(const ot_rmsnorm_params* p, void* stream);
int ot_rmsnorm_bwd(const ot_rmsnorm_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Non-sequence tokenizer (OT/model.py:211-214, 239-254): concat of scalar features -> Dense(d*L_NS)
 * with bias -> Reshape [L_NS, d], written straight into the token-major X0 rows row0 + j*B + b.
 * fp32 math (ids enter as magnitudes).  x: [B, n_feat] fp32, W: [n_feat, L_NS*d] fp32.
 * backward: dW += x^T dX0_ns, dbias += colsum(dX0_ns)  (fp32 atomics; caller zeroes)
 * ---------------------------------------------------------------------------------------------- */
typedef struct ot_ns_tokenizer_params {
  const float* x; const float* W; const float* bias;
  void* out; const void* dout; int64_t ldo;
  float* dW; float* dbias;
  int64_t row0;
  int32_t B, L_ns, d, n_feat;
  float* out_hp;         /* optional fp32 copy of the NS rows, [L_ns*B, d] (forward) */
} ot_ns_tokenizer_params;

int ot_ns_tokenizer_fwd(const ot_ns_tokenizer_params* p, void* stream);
int ot_ns_tokenizer_bwd(const ot_ns_tokenizer_params* p, void* stream);

/* rows [row0, row0+n_rows) of out <- bf16(vec[0..d)) : the [SEP] embedding rows (OT/model.py:269-272). */
int ot_fill_rows(const float* vec, void* out, int64_t ldo, int64_t row0, int64_t n_rows, int32_t d, void* stream);

/* Column sums for bias / [SEP] gradients: out[group_start + u*group_stride][n] += sum over the
 * rows of unit u of in[row, n]   (fp32 atomics; caller zeroes). */
typedef struct ot_colsum_params {
  const void* in; int64_t ld;
  int64_t row_start;
  int32_t n_units, rows_per_unit, group_start, group_stride;
  float* out; int64_t out_group_stride;
  int32_t N;
} ot_colsum_params;

int ot_colsum(const ot_colsum_params* p, void* stream);

/* out[r, c] = keep(r, c) ? in[r, c] / (1 - rate) : 0 with the mask of OT_EPI_DROPOUT (same seed, same [rows, cols]
 * index space): the backward of the dropout on a branch output.  bf16 in/out, cols % 8 == 0. */
int ot_dropout_mask(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t rows, int32_t cols,
                    uint32_t seed, float rate, void* stream);

/* ---- parameter update: per-tensor clip_by_norm + RMSprop (OT/train.py:131-138; optimizer OT/train.py:65-70,
 * hyper-parameters OT/config.py:39-52).  Keras 2.12 RMSprop.update_step semantics:
 *   g' = grad_scale * g * clip_norm / max(||grad_scale * g||_2, clip_norm)      (tf.clip_by_norm, per tensor; off if clip_norm <= 0)
 *   rms = rho * rms + (1 - rho) * g'^2 ;  inc = lr * g' * rsqrt(rms + eps)
 *   momentum > 0:  mom = momentum * mom + inc ; w -= mom        else  w -= inc
 * grad / rms / mom are flat fp32 buffers of n_flat elements in which tensor s occupies
 * [seg_off[s], seg_off[s] + seg_numel[s]); every seg_off and n_flat is a multiple of OT_OPT_CHUNK.  The fp32 masters
 * stay in the caller's own allocations: param_ptrs[s] (16-byte aligned).  All tables are DEVICE arrays.
 * sqnorm: device workspace of n_seg floats; on return it holds ||grad_scale * g_s||^2 (when clip_norm > 0). */
#define OT_OPT_CHUNK 1024
typedef struct ot_rmsprop_params {
  float* const* param_ptrs;
  const int64_t* seg_off;      /* [n_seg + 1] */
  const int64_t* seg_numel;    /* [n_seg] */
  int32_t n_seg;
  int64_t n_flat;
  float* grad;                 /* read; zeroed after use when zero_grad != 0 */
  float* rms;
  float* mom;                  /* may be NULL when momentum == 0 */
  float* sqnorm;
  float lr, rho, momentum, eps, clip_norm, grad_scale;
  int32_t zero_grad;
  /* Clip granularity (ABI 11).  tf.clip_by_norm runs per KERAS VARIABLE (OT/train.py:135), while a flat-buffer tensor
   * may pack several of them: Wqkv [G, d, 3d] holds 3*G Dense kernels (q|k|v column blocks of every weight group),
   * W1 / b1 / W2 / b2 [G, ...] hold G each.  Element `local` of tensor s belongs to clip slot
   *     slot_base[s] + (local / slot_outer[s]) * (slot_row[s] / slot_part[s]) + (local % slot_row[s]) / slot_part[s]
   * (slot_outer = elements per weight group, slot_row = row length, slot_part = columns per variable inside a row).
   * seg_slot == NULL: one slot per tensor (slot == s).  seg_slot: int64 DEVICE table [n_seg][4] =
   * {slot_base, slot_outer, slot_row, slot_part}; slot_part % 4 == 0.  sqnorm then has n_slots entries. */
  const int64_t* seg_slot;
  int32_t n_slots;
} ot_rmsprop_params;
int ot_clip_rmsprop_step(const ot_rmsprop_params* p, void* stream);

/* ---- ID front end of the sequence tokenizer (north_star item 1; extension of OT/model.py:217-219 along PAPER:89-109 and
 * the lookup-and-concat idiom of recall/bert_like/kuaiformer/practice/model.py:58-94).  One fp32 table holds the rows of
 * all fields back to back: field f owns rows [field_off[f], field_off[f] + field_rows[f]), every row has `ef` floats.
 *   ot_embed_gather_fwd  : events[e, f*ef + j] = bf16(table[(field_off[f] + ids[e, f]) * ef + j])      (bit-exact copy)
 *   ot_embed_scatter_bwd : grad[(field_off[f] + ids[e, f]) * ef + j] += events[e, f*ef + j]   (events = d loss / d events)
 *   ot_embed_adagrad_step: once per row present in ids:  acc += g*g ; table -= lr * g / sqrt(acc + eps) ; g = 0
 *                          (Keras Adagrad on the summed sparse gradient; OT/config.py:39-47 sparse_optimizer / sparse_lr)
 * ids: int32 [n_events, n_fields]; ids outside [0, field_rows[f]) give a zero row and are counted in *bad_ids (may be NULL).
 * stamp: int32 [total rows], zero-initialised by the caller, step_id != 0 and different from the previous step's.
 * field_off / field_rows are int64 DEVICE arrays. */
typedef struct ot_embed_params {
  const float* table;
  const int64_t* field_off;
  const int64_t* field_rows;
  const int32_t* ids;
  void* events; int64_t ld_events;   /* bf16 [n_events, n_fields*ef] */
  int64_t n_events;
  int32_t n_fields, ef;              /* ef % 8 == 0 */
  int32_t* bad_ids;
  float* grad;                       /* fp32, same shape as table */
  float* acc;                        /* Adagrad accumulator, same shape */
  int32_t* stamp;
  int32_t step_id;
  float lr, eps;
} ot_embed_params;
int ot_embed_gather_fwd(const ot_embed_params* p, void* stream);
int ot_embed_scatter_bwd(const ot_embed_params* p, void* stream);
int ot_embed_adagrad_step(const ot_embed_params* p, void* stream);

/* ---- output norm + task heads (+ BCE) on the last token, fp32 (OT/model.py:322-330 heads, :384-391 output norm and
 * last-token slice; loss OT/train.py:84-87, 124-128).  Per task t: pre = RMSNorm(x) W0[t] + b0[t], h = gelu_erf(pre),
 * logit = h . W1[t] + b1[t], prob = sigmoid(logit).  With labels: *loss += sum_t mean_b BCE(prob, y) (Keras semantics: clip
 * to [1e-7, 1-1e-7], log(p + 1e-7)) and g_bce[t, b] = d loss / d logit[t, b].
 * ot_heads_bwd takes dlogit [T, B] (any mix of g_bce * upstream and gradients w.r.t. probs / logits) and ACCUMULATES
 * dW0, db0, dW1, db1, dgain (caller zeroes), and writes dx.  All tensors fp32; W0[t] is [d, hidden] (Keras [in, out]). */
#define OT_MAX_TASKS 4
typedef struct ot_heads_params {
  const float* x; int64_t ldx;          /* [B, d] last-token rows of the residual stream */
  const float* gain; float eps;         /* output_norm.scale */
  int32_t B, d, hidden, n_tasks;        /* hidden = d/2 in the reference */
  const float* W0[OT_MAX_TASKS]; const float* b0[OT_MAX_TASKS]; const float* W1[OT_MAX_TASKS]; const float* b1[OT_MAX_TASKS];
  float* xn; float* rstd; float* pre;   /* saved for the backward: [B, d], [B], [n_tasks, B, hidden] */
  float* logits; float* probs;          /* [n_tasks, B] */
  const float* labels; float* loss; float* g_bce;   /* optional (all three or none): [n_tasks, B], scalar, [n_tasks, B] */
  const float* dlogit;                  /* backward input [n_tasks, B] */
  float* dpre;                          /* backward workspace [n_tasks, B, hidden] */
  float* dW0[OT_MAX_TASKS]; float* db0[OT_MAX_TASKS]; float* dW1[OT_MAX_TASKS]; float* db1[OT_MAX_TASKS];
  float* dgain;
  float* dx; int64_t lddx;
} ot_heads_params;
int ot_heads_fwd(const ot_heads_params* p, void* stream);
int ot_heads_bwd(const ot_heads_params* p, void* stream);

/* ---- streaming evaluation metrics (OT/train.py:95-109, 141-150, 178-187, 248-249; OT/evaluate.py:39-56, 91-99, 109-111):
 * per binary task the reference keeps Keras AUC() / BinaryAccuracy() / Precision() / Recall() (+ F1Score, absent from the
 * pinned TF 2.12 - SURVEY.md D11 - and the BinaryCrossentropy metric).  One pass over the predictions feeds all of them:
 *   state[task] (int64 words): pos_hist[NT] | neg_hist[NT] | tp fp tn fn count rejected bce_sum(double bits) reserved
 *   bucket  b = max(ceil(clip(p, 0, 1) * (NT - 1)) - 1, 0) in fp32   (Keras' update for evenly spaced thresholds)
 *   tp/fp/tn/fn at p > threshold;  bce as Keras (clip to [1e-7, 1 - 1e-7], log(. + 1e-7))
 *   samples with a NaN prediction or a label outside {0, 1} are skipped and counted in `rejected`
 * ot_metrics_update ACCUMULATES into state (caller zeroes to reset); ot_metrics_result writes, per task,
 *   result[task] (double words): auc accuracy precision recall f1 logloss count rejected
 * with fp32 element-wise arithmetic for auc / precision / recall / f1 as Keras' result() (ROC, trapezoid summation).
 * probs / labels: fp32, task t at base + t * ld, B samples each (the [n_tasks, B] layout ot_heads_fwd writes). */
#define OT_METRICS_MAX_THRESHOLDS 512
#define OT_METRICS_TAIL_WORDS 8
#define OT_METRICS_RESULT_WORDS 8
typedef struct ot_metrics_params {
  const float* probs; const float* labels; int64_t ld;
  int64_t B;
  int32_t n_tasks;
  int32_t num_thresholds;       /* NT; Keras default 200 */
  float threshold;              /* 0.5 */
  int64_t* state; int64_t state_stride;   /* >= 2 * NT + OT_METRICS_TAIL_WORDS words per task */
  double* result;               /* ot_metrics_result only: [n_tasks, OT_METRICS_RESULT_WORDS] */
} ot_metrics_params;
int ot_metrics_update(const ot_metrics_params* p, void* stream);
int ot_metrics_result(const ot_metrics_params* p, void* stream);

/* ---- exact (tie-aware) ROC-AUC, overall or per segment (per-user AUC), for the north_star's "AUC delta <= 1e-4" check
 * (Keras AUC() is a 200-threshold approximation).  Three steps:
 *   ot_auc_pack_keys : keys[i] = segment_ids[i] << 33 | ordered_bits(probs[i]) << 1 | label   (segment 0 if segment_ids NULL;
 *                      NaN / non-binary label / segment outside [0, n_segments) -> INT64_MAX key, counted in *rejected)
 *   caller sorts keys ascending (any int64 sort)
 *   ot_auc_ranksum   : per segment s:  seg_count[s] += #samples, seg_pos[s] += #positives,
 *                      seg_sum2[s] += sum over positives of (first + last + 1) of their tie group in the sorted order
 *                      = twice the 1-based mid-rank (ACCUMULATES; caller zeroes).
 * With start[s] the exclusive prefix sum of seg_count:  U_s = (seg_sum2[s] - 2*seg_pos[s]*start[s])/2 - P_s(P_s+1)/2,
 * AUC_s = U_s / (P_s * (count_s - P_s)).  n_segments < 2^30 - 1. */
typedef struct ot_auc_params {
  const float* probs; const float* labels; const int32_t* segment_ids;
  int64_t n;
  int32_t n_segments;
  int64_t* keys;
  int32_t* rejected;            /* may be NULL */
  int64_t* seg_count; int64_t* seg_pos; int64_t* seg_sum2;
} ot_auc_params;
int ot_auc_pack_keys(const ot_auc_params* p, void* stream);
int ot_auc_ranksum(const ot_auc_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ONETRANS_B200_H_ */
