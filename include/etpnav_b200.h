/* etpnav_b200 — C ABI of the B200-native ETPNav planner hot path.
 *
 * The reference (MarSaKi/ETPNav) is pure Python and has no FFI: its planner is reached as three
 * methods of one torch.nn.Module,
 *     GlocalTextPathNavCMT.forward_txt        vlnce_baselines/models/etp/vilmodel_cmt.py:684
 *     GlocalTextPathNavCMT.forward_panorama   vlnce_baselines/models/etp/vilmodel_cmt.py:690
 *     GlocalTextPathNavCMT.forward_navigation vlnce_baselines/models/etp/vilmodel_cmt.py:721
 * called from ETP.forward (vlnce_baselines/models/Policy_ViewSelection_ETP.py:167,346,352-357).
 * This header is the boundary a binding for that module sits on: plain pointers and sizes, no torch
 * types.  All pointers are DEVICE pointers unless stated; the caller (PyTorch in the shipped host
 * module etpnav_b200/planner.py) owns every buffer; every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*) and never synchronises.  Return value: 0 = ok, negative = error
 * (see etp_last_error()).  bf16 tensors are raw uint16 storage (`void*` here).
 *
 * Two layers are exported:
 *   (1) step level  — etp_forward_panorama / etp_forward_navigation / etp_forward_txt and their
 *       backward twins: one call = one reference method (INTEGRATION.md shows the binding).
 *   (2) operator level — the fused kernels the step level is made of (GEMM+epilogue, attention,
 *       LayerNorm, token packing, SAP head, AdamW); exported for the parity tests and for hosts that
 *       want to compose them differently.
 */
#ifndef ETPNAV_B200_H_
#define ETPNAV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ETP_OK 0
#define ETP_ERR_INVALID (-1)
#define ETP_ERR_CUDA (-2)
#define ETP_ERR_NO_DEVICE (-3)

/* ---------------------------------------------------------------------------------------------
 * library
 * ------------------------------------------------------------------------------------------- */
int etp_version(void);                /* 100 * major + minor */
const char* etp_last_error(void);     /* text of the last error on this thread */
/* 0 if the current CUDA device is sm_100 (B200); ETP_ERR_NO_DEVICE otherwise. */
int etp_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * operator level
 * ------------------------------------------------------------------------------------------- */

/* D[M,N] = epilogue(alpha * A.B^T) on tcgen05 tensor cores; replaces torch.nn.Linear / matmul on the
 * reference path (vilmodel_cmt.py:108-110,326-328,151,178,190,654; common/transformer.py:174-181).
 * a_mn/b_mn = 0: operand stored [rows, K] (K contiguous); 1: stored [K, rows] (rows contiguous).
 * epilogue: x = alpha*acc + bias[n]; out_pre = bf16(x); x = act(x); x *= f(aux); x += resid;
 *           out_f32 (=, or += atomically) x; out_bf16 = bf16(x). */
typedef struct {
  int32_t M, N, K;
  const void* A; int32_t lda; int32_t a_mn;
  const void* B; int32_t ldb; int32_t b_mn;
  float alpha;
  const float* bias;
  int32_t act;        /* 0 none, 1 gelu(erf), 2 relu */
  int32_t aux_mode;   /* 0 none, 1: *= gelu'(aux), 2: *= (aux > 0) */
  const void* aux; int32_t ld_aux;
  const float* resid; int32_t ld_resid;
  float* out_f32; int32_t ld_f32; int32_t atomic;
  void* out_bf16; int32_t ld_bf16;
  void* out_pre; int32_t ld_pre;
  int32_t k_splits;   /* >1 requires atomic */
  int32_t block_n;    /* 0 auto, 128, 256 */
} etp_gemm_args;
int etp_gemm(const etp_gemm_args* args, void* stream);

/* softmax(scale*q.k^T + bias).v, head dim 64; bias = (key_valid ? 0 : mask_value) + pair_w*pair + pair_b.
 * BertOutAttention (vilmodel_cmt.py:325-352), BertSelfAttention (:103-141, graph bias :391-393),
 * nn.MultiheadAttention in the pano encoder (common/transformer.py:176).  impl: 0 auto, 1 CUDA-core,
 * 2 tcgen05. */
typedef struct {
  int32_t B, heads, Sq, Sk;
  const void* q; int32_t ldq;
  const void* k; int32_t ldk;
  const void* v; int32_t ldv;
  float scale;
  const uint8_t* key_valid;  /* [B,Sk] or NULL */
  float mask_value;
  const float* pair;         /* [B,Sq,Sk] or NULL */
  float pair_w, pair_b;
  void* out; int32_t ldo;
  float* lse;                /* [B,heads,Sq] or NULL */
  int32_t impl;
} etp_attn_args;
int etp_attention_fwd(const etp_attn_args* args, void* stream);

/* LayerNorm over the last dim (768), biased variance (torch.nn.LayerNorm / BertLayerNorm,
 * vilmodel_cmt.py:24-28). */
int etp_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t H,
                      float* y_f32, void* y_bf16, float* mean, float* rstd, void* stream);
int etp_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      int32_t rows, int32_t H, float* dx_f32, int32_t accumulate_dx, void* dx_bf16, float* dgamma,
                      float* dbeta, void* stream);
int etp_colsum_bf16(const void* x, int32_t rows, int32_t cols, int32_t ld, float* out, void* stream);
int etp_colsum_f32(const float* x, int32_t rows, int32_t cols, int32_t ld, float* out, void* stream);
int etp_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);

/* view-token packing, vilmodel_cmt.py:695-711 */
typedef struct {
  int32_t rows;
  const float* rgb_lin; const float* dep_lin; const float* loc_fts; const int64_t* nav_types;
  const float* loc_w; const float* loc_b;
  const float* img_g; const float* img_b; const float* dep_g; const float* dep_b;
  const float* loc_g; const float* loc_bb; const float* out_g; const float* out_b;
  const float* nav_emb; const float* tok_emb1;
  float* x_f32; float* loc_lin; float* sum_pre; float* stats;
} etp_pano_pack_args;
int etp_pano_pack_fwd(const etp_pano_pack_args* args, void* stream);

/* node-token packing, vilmodel_cmt.py:728-730 */
typedef struct {
  int32_t rows;
  const float* img_fts; const int64_t* step_ids; const float* pos_fts;
  const float* pos_w; const float* pos_b; const float* pos_g; const float* pos_bb;
  const float* step_emb;
  float* x_f32; void* x_bf16; float* pos_lin; float* stats;
} etp_node_pack_args;
int etp_node_pack_fwd(const etp_node_pack_args* args, void* stream);

/* SAP head tail, vilmodel_cmt.py:654-658 (net.2, net.4) and :742-744 */
int etp_sap_tail_fwd(const float* relu_out, const float* gamma, const float* beta, const float* w4, const float* b4,
                     const uint8_t* visited, const uint8_t* valid, int32_t rows, int32_t H, float* logits,
                     float* mean, float* rstd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ETPNAV_B200_H_ */
