// extern "C" surface of libetpnav_b200.so (declared in include/etpnav_b200.h).
#include "../../include/etpnav_b200.h"

#include "host.h"
#include "ops.h"

namespace etp {
const char* last_error_cstr();
void prof_enable(bool on);
int prof_collect(double* ms, double* flops, long long* count, char* report, size_t cap);
int attention_tc_fwd(const AttnArgs& a, cudaStream_t stream);
int attention_tc_fwd_v1(const AttnArgs& a, cudaStream_t stream);
int attention_tc2_fwd(const AttnArgs& a, cudaStream_t stream);
int attention_dispatch(const AttnArgs& a, cudaStream_t stream);
int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t stream);
void attention_bwd_tc_set_debug(void* dev_buf);
}

using namespace etp;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define ETP_API __attribute__((visibility("default")))
extern "C" {

ETP_API int etp_version(void) { return 200; }
/* sizeof() of the public structs, in the order: gemm_args, attn_args, attn_bwd_args, pano_pack_args, node_pack_args,
 * dropout, layer_weights, nav_weights, nav_inputs, pano_layer_weights, pano_weights, pano_inputs, txt_weights — lets a
 * binding check its own struct mirrors (tests/test_host_cpu.py does it for the ctypes ones). */
ETP_API int etp_struct_sizes(int32_t* out, int32_t n) {
  const int32_t v[] = {(int32_t)sizeof(etp_gemm_args), (int32_t)sizeof(etp_attn_args), (int32_t)sizeof(etp_attn_bwd_args),
                       (int32_t)sizeof(etp_pano_pack_args), (int32_t)sizeof(etp_node_pack_args), (int32_t)sizeof(etp_dropout),
                       (int32_t)sizeof(etp_layer_weights), (int32_t)sizeof(etp_nav_weights), (int32_t)sizeof(etp_nav_inputs),
                       (int32_t)sizeof(etp_pano_layer_weights), (int32_t)sizeof(etp_pano_weights), (int32_t)sizeof(etp_pano_inputs),
                       (int32_t)sizeof(etp_txt_weights)};
  const int32_t m = static_cast<int32_t>(sizeof(v) / sizeof(v[0]));
  for (int32_t i = 0; i < n && i < m; ++i) out[i] = v[i];
  return m;
}
ETP_API const char* etp_last_error(void) { return last_error_cstr(); }

ETP_API long long etp_launch_count(void) { return g_launches.load(); }
ETP_API void* etp_event_create(void) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  return e;
}
ETP_API void etp_event_destroy(void* event) {
  if (event) cudaEventDestroy(static_cast<cudaEvent_t>(event));
}
ETP_API int etp_event_record(void* event, void* stream) {
  ETP_REQUIRE(event != nullptr, "etp_event_record: null event");
  ETP_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), S(stream)));
  return ETP_OK;
}
ETP_API int etp_stream_wait_event(void* stream, void* event) {
  ETP_REQUIRE(event != nullptr, "etp_stream_wait_event: null event");
  ETP_CHECK_CUDA(cudaStreamWaitEvent(S(stream), static_cast<cudaEvent_t>(event), 0));
  return ETP_OK;
}
ETP_API void etp_set_sm_reserve(int32_t n) { set_sm_reserve(n); }
ETP_API void etp_prof_gemm_enable(int on) { prof_enable(on != 0); }
ETP_API int etp_prof_gemm_collect(double* total_ms, double* total_flops, long long* launches) {
  ETP_REQUIRE(total_ms && total_flops && launches, "etp_prof_gemm_collect: null argument");
  return prof_collect(total_ms, total_flops, launches, nullptr, 0);
}
ETP_API int etp_prof_report(char* buf, size_t cap) {
  ETP_REQUIRE(buf && cap > 0, "etp_prof_report: null argument");
  return prof_collect(nullptr, nullptr, nullptr, buf, cap);
}

/* developer hook (not part of the ported interface): CTA-0 timeline of the next tcgen05 attention-backward launches */
ETP_API void etp_debug_attention_bwd_timeline(void* dev_buf_128_u64) { attention_bwd_tc_set_debug(dev_buf_128_u64); }

ETP_API int etp_dropout_mask(const etp_dropout* d, float p, uint32_t site, int64_t n, uint8_t* out, void* stream) {
  ETP_REQUIRE(d != nullptr, "etp_dropout_mask: null argument");
  return dropout_mask(make_drop_host(d->seed, p, site), n, out, S(stream));
}

ETP_API int etp_check_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(ETP_ERR_NO_DEVICE, "no CUDA device");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail(ETP_ERR_NO_DEVICE, "cannot query CUDA device");
  if (prop.major != 10) return fail(ETP_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) +
                                                             std::to_string(prop.minor) + ", kernels are sm_100a only");
  return ETP_OK;
}

ETP_API int etp_gemm(const etp_gemm_args* g, void* stream) {
  ETP_REQUIRE(g != nullptr, "etp_gemm: null args");
  GemmArgs a;
  a.M = g->M; a.N = g->N; a.K = g->K;
  a.A = static_cast<const bf16*>(g->A); a.lda = g->lda; a.a_mn = g->a_mn;
  a.B = static_cast<const bf16*>(g->B); a.ldb = g->ldb; a.b_mn = g->b_mn;
  a.alpha = g->alpha; a.bias = g->bias; a.act = g->act; a.aux_mode = g->aux_mode;
  a.aux = static_cast<const bf16*>(g->aux); a.ld_aux = g->ld_aux;
  a.resid = g->resid; a.ld_resid = g->ld_resid;
  a.out_f32 = g->out_f32; a.ld_f32 = g->ld_f32; a.atomic = g->atomic;
  a.out_bf16 = static_cast<bf16*>(g->out_bf16); a.ld_bf16 = g->ld_bf16;
  a.out_pre = static_cast<bf16*>(g->out_pre); a.ld_pre = g->ld_pre;
  a.k_splits = g->k_splits < 1 ? 1 : g->k_splits; a.block_n = g->block_n;
  a.colsum = g->colsum; a.pre_mode = g->pre_mode;
  a.drop_key = g->drop_key; a.drop_thr = g->drop_thr; a.drop_scale = g->drop_scale;
  return gemm(a, S(stream));
}

ETP_API int etp_attention_fwd(const etp_attn_args* g, void* stream) {
  ETP_REQUIRE(g != nullptr, "etp_attention_fwd: null args");
  AttnArgs a;
  a.B = g->B; a.heads = g->heads; a.Sq = g->Sq; a.Sk = g->Sk;
  a.q = static_cast<const bf16*>(g->q); a.ldq = g->ldq;
  a.k = static_cast<const bf16*>(g->k); a.ldk = g->ldk;
  a.v = static_cast<const bf16*>(g->v); a.ldv = g->ldv;
  a.scale = g->scale; a.key_valid = g->key_valid; a.mask_value = g->mask_value;
  a.pair = g->pair; a.pair_w = g->pair_w; a.pair_b = g->pair_b;
  a.pair_w_dev = g->pair_w_dev; a.pair_b_dev = g->pair_b_dev;
  a.out = static_cast<bf16*>(g->out); a.ldo = g->ldo; a.lse = g->lse;
  if (g->impl == 1) return attention_fwd(a, S(stream));
  if (g->impl == 2) return attention_tc_fwd(a, S(stream));      /* tcgen05, generation chosen by ETP_ATTN_V2 */
  if (g->impl == 3) return attention_tc2_fwd(a, S(stream));     /* tcgen05, two CTAs per SM */
  if (g->impl == 4) return attention_tc_fwd_v1(a, S(stream));   /* tcgen05, one CTA per SM */
  return attention_dispatch(a, S(stream));
}

ETP_API int etp_attention_bwd(const etp_attn_bwd_args* g, void* stream) {
  ETP_REQUIRE(g != nullptr, "etp_attention_bwd: null args");
  AttnBwdArgs a;
  a.B = g->B; a.heads = g->heads; a.Sq = g->Sq; a.Sk = g->Sk;
  a.q = static_cast<const bf16*>(g->q); a.k = static_cast<const bf16*>(g->k); a.v = static_cast<const bf16*>(g->v);
  a.ldq = g->ldq; a.ldk = g->ldk; a.ldv = g->ldv;
  a.out = static_cast<const bf16*>(g->out); a.ldo = g->ldo;
  a.dout = static_cast<const bf16*>(g->dout); a.lddo = g->lddo;
  a.lse = g->lse; a.dvec = g->dvec; a.scale = g->scale; a.key_valid = g->key_valid; a.mask_value = g->mask_value;
  a.pair = g->pair; a.pair_w = g->pair_w; a.pair_b = g->pair_b;
  a.dq = static_cast<bf16*>(g->dq); a.dk = static_cast<bf16*>(g->dk); a.dv = static_cast<bf16*>(g->dv);
  a.lddq = g->lddq; a.lddk = g->lddk; a.lddv = g->lddv; a.dpair_w = g->dpair_w; a.dpair_b = g->dpair_b;
  if (g->impl == 1) return attention_bwd(a, S(stream));
  if (g->impl == 2) return attention_bwd_tc(a, S(stream));
  return attention_bwd_dispatch(a, S(stream));
}

ETP_API int etp_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t H,
                      float* y_f32, void* y_bf16, float* mean, float* rstd, void* stream) {
  return layernorm_fwd(x, gamma, beta, eps, rows, H, y_f32, static_cast<bf16*>(y_bf16), mean, rstd, S(stream));
}
ETP_API int etp_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      int32_t rows, int32_t H, float* dx_f32, int32_t accumulate_dx, void* dx_bf16, float* dgamma,
                      float* dbeta, void* stream) {
  return layernorm_bwd(dy, x, gamma, mean, rstd, rows, H, dx_f32, accumulate_dx, static_cast<bf16*>(dx_bf16), dgamma,
                       dbeta, S(stream));
}
ETP_API int etp_colsum_bf16(const void* x, int32_t rows, int32_t cols, int32_t ld, float* out, void* stream) {
  return colsum_bf16(static_cast<const bf16*>(x), rows, cols, ld, out, S(stream));
}
ETP_API int etp_colsum_f32(const float* x, int32_t rows, int32_t cols, int32_t ld, float* out, void* stream) {
  return colsum_f32(x, rows, cols, ld, out, S(stream));
}
ETP_API int etp_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  return cast_f32_to_bf16(x, static_cast<bf16*>(y), n, S(stream));
}

ETP_API int etp_pano_pack_fwd(const etp_pano_pack_args* g, void* stream) {
  ETP_REQUIRE(g != nullptr, "etp_pano_pack_fwd: null args");
  PanoPackArgs a;
  a.rows = g->rows; a.rgb_lin = g->rgb_lin; a.dep_lin = g->dep_lin; a.loc_fts = g->loc_fts; a.nav_types = g->nav_types;
  a.loc_w = g->loc_w; a.loc_b = g->loc_b; a.img_g = g->img_g; a.img_b = g->img_b; a.dep_g = g->dep_g; a.dep_b = g->dep_b;
  a.loc_g = g->loc_g; a.loc_bb = g->loc_bb; a.out_g = g->out_g; a.out_b = g->out_b; a.nav_emb = g->nav_emb;
  a.tok_emb1 = g->tok_emb1; a.x_f32 = g->x_f32; a.loc_lin = g->loc_lin; a.sum_pre = g->sum_pre; a.stats = g->stats;
  return pano_pack_fwd(a, S(stream));
}

ETP_API int etp_node_pack_fwd(const etp_node_pack_args* g, void* stream) {
  ETP_REQUIRE(g != nullptr, "etp_node_pack_fwd: null args");
  NodePackArgs a;
  a.rows = g->rows; a.img_fts = g->img_fts; a.step_ids = g->step_ids; a.pos_fts = g->pos_fts;
  a.pos_w = g->pos_w; a.pos_b = g->pos_b; a.pos_g = g->pos_g; a.pos_bb = g->pos_bb; a.step_emb = g->step_emb;
  a.x_f32 = g->x_f32; a.x_bf16 = static_cast<bf16*>(g->x_bf16); a.pos_lin = g->pos_lin; a.stats = g->stats;
  return node_pack_fwd(a, S(stream));
}

ETP_API int etp_sap_tail_fwd(const float* relu_out, const float* gamma, const float* beta, const float* w4, const float* b4,
                     const uint8_t* visited, const uint8_t* valid, int32_t rows, int32_t H, float* logits,
                     float* mean, float* rstd, void* stream) {
  return sap_tail_fwd(relu_out, gamma, beta, w4, b4, visited, valid, rows, H, logits, mean, rstd, S(stream));
}

ETP_API int etp_step_loss(const float* logits, const int64_t* labels, int32_t B, int32_t N, int64_t ignore_index,
                          float grad_scale, float* loss_sum, float* dlogits, float* probs, int64_t* argmax, void* stream) {
  return step_loss(logits, labels, B, N, ignore_index, grad_scale, loss_sum, dlogits, probs, argmax, S(stream));
}

}  // extern "C"
