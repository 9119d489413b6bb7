// VAE train-step engine (included by engine.cu): Encoder x->h->(mu,log_var), reparam,
// Decoder z->h->x (src/vae.py:47-106), losses recon = sum (x-out)^2, kl (src/vae.py:203,212),
// backward of recon+kl in closed form (SURVEY.md A.1 "VAE" row), Adam with coupled weight
// decay (src/vae.py:139-142).  Same tensor-core GEMM kernel as the GAN engine.
//
// Flat fp32 layout: [W1 (h,x) | b1 (h) | Wm (z,h) | Wv (z,h) | bm (z) | bv (z) | W3 (h,z) | b3 (h) | W4 (x,h) | b4 (x)]
// (mu / log_var weights adjacent so that both heads are ONE 400 -> 2z GEMM).
struct VaePlans {
  GemmPlan e1, e2, d1, d2, d2_fwd, gw4, da3, gw3, dz, gwmv, da1, gw1;
};

struct gm_vae {
  gm_ctx* ctx;
  gm_vae_desc d;
  int X, H, Z, XP, HP, ZP, Bmax;
  int off_w1, off_b1, off_wmv, off_bmv, off_w3, off_b3, off_w4, off_b4, total;
  float *par = nullptr, *grd = nullptr, *am = nullptr, *av = nullptr;
  __nv_bfloat16 *Xin = nullptr, *H1 = nullptr, *Zb = nullptr, *H3 = nullptr, *DA4 = nullptr, *DA3 = nullptr, *DML = nullptr, *DA1 = nullptr;
  __nv_bfloat16 *W1_s = nullptr, *Wmv_s = nullptr, *Wmv_t = nullptr, *W3_s = nullptr, *W3_t = nullptr, *W4_s = nullptr, *W4_t = nullptr;
  float *MULV = nullptr, *DZ = nullptr, *EPS = nullptr, *slots_r = nullptr, *losses = nullptr;
  float *P1 = nullptr, *Pmv = nullptr, *P3 = nullptr, *P4 = nullptr;
  double *part_r = nullptr, *part_k = nullptr;
  bool lazy = false, pend = false;      // lazy gradients: gm_vae_apply gathers the split-K partials itself
  GradSegs pend_segs;
  long long pool_n = 0, pool_bpe = 0, pool_bs = 0;   // on-device epoch sampler (gm_vae_set_sampler)
  uint64_t pool_seed = 0;
  std::map<int, VaePlans> plans;
  std::vector<void*> allocs;
  BfArena arena;
  long long lo = 0;
};

static int vae_alloc(gm_vae* g, __nv_bfloat16** p, size_t count) {   // bf16 buffers come from the arena
  g->arena.request(p, count);
  return GM_OK;
}

template <typename T>
static int vae_alloc(gm_vae* g, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T));
  if (e != cudaSuccess) return fail(g->ctx, GM_ERR_CUDA, "cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e));
  cudaMemset(q, 0, count * sizeof(T));
  g->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return GM_OK;
}

extern "C" int gm_vae_destroy(gm_vae* g) {
  if (!g) return GM_OK;
  for (void* p : g->allocs) cudaFree(p);
  delete g;
  return GM_OK;
}

extern "C" int gm_vae_create(gm_ctx* c, const gm_vae_desc* d, gm_vae** out) {
  if (!c || !d || !out) return GM_ERR_ARG;
  *out = nullptr;
  if (!c->encode) return fail(c, GM_ERR_STATE, "context has no device");
  if (d->image_size % 16 || d->hidden_dim % 16 || d->image_size <= 0 || d->hidden_dim <= 0 || d->z_dim <= 0 || d->max_batch <= 0)
    return fail(c, GM_ERR_ARG, "image_size and hidden_dim must be positive multiples of 16");
  if (d->hidden_dim + 1 > 448 || 2 * d->z_dim > 64)
    return fail(c, GM_ERR_UNSUPPORTED, "hidden_dim <= 447 and z_dim <= 32 in this build (got %d, %d)", d->hidden_dim, d->z_dim);
  if (d->dtype_mode != GM_PREC_BF16 && d->dtype_mode != GM_PREC_SPLIT) return fail(c, GM_ERR_ARG, "unknown dtype_mode %d", d->dtype_mode);
  gm_vae* g = new gm_vae();
  g->ctx = c; g->d = *d;
  const int X = g->X = d->image_size, H = g->H = d->hidden_dim, Z = g->Z = d->z_dim;
  g->Bmax = d->max_batch;
  g->XP = rup(X + 1, 16); g->HP = rup(H + 1, 16); g->ZP = rup(Z + 1, 64);
  g->off_w1 = 0; g->off_b1 = H * X; g->off_wmv = g->off_b1 + H; g->off_bmv = g->off_wmv + 2 * Z * H;
  g->off_w3 = g->off_bmv + 2 * Z; g->off_b3 = g->off_w3 + H * Z; g->off_w4 = g->off_b3 + H; g->off_b4 = g->off_w4 + X * H;
  g->total = g->off_b4 + X;
  const size_t B = g->Bmax;
  int rc = GM_OK;
#define TRYV(x) do { rc = (x); if (rc) { gm_vae_destroy(g); return rc; } } while (0)
  TRYV(vae_alloc(g, &g->Xin, B * g->XP));
  TRYV(vae_alloc(g, &g->H1, B * g->HP));
  TRYV(vae_alloc(g, &g->Zb, B * g->ZP));
  TRYV(vae_alloc(g, &g->H3, B * g->HP));
  TRYV(vae_alloc(g, &g->DA4, B * g->XP));
  TRYV(vae_alloc(g, &g->DA3, B * g->HP));
  TRYV(vae_alloc(g, &g->DML, B * 64));
  TRYV(vae_alloc(g, &g->DA1, B * g->HP));
  TRYV(vae_alloc(g, &g->W1_s, size_t(H) * X));
  TRYV(vae_alloc(g, &g->Wmv_s, size_t(64) * H));
  TRYV(vae_alloc(g, &g->Wmv_t, size_t(H) * 64));
  TRYV(vae_alloc(g, &g->W3_s, size_t(H) * g->ZP));
  TRYV(vae_alloc(g, &g->W3_t, size_t(32) * H));
  TRYV(vae_alloc(g, &g->W4_s, size_t(X) * H));
  TRYV(vae_alloc(g, &g->W4_t, size_t(H) * X));
  TRYV(vae_alloc(g, &g->MULV, B * 64));
  TRYV(vae_alloc(g, &g->DZ, B * 32));
  TRYV(vae_alloc(g, &g->EPS, B * Z));
  TRYV(vae_alloc(g, &g->slots_r, size_t(2 * cdiv(X, 208)) * B));
  TRYV(vae_alloc(g, &g->losses, 4));
  TRYV(vae_alloc(g, &g->part_r, size_t(c->num_sms) * 2));
  TRYV(vae_alloc(g, &g->part_k, size_t(cdiv(int(B) * ((Z + 8) / 8), 256))));
  const int ns = c->num_sms;
  TRYV(vae_alloc(g, &g->P1, size_t(ns / cdiv(H, BM) > 0 ? ns / cdiv(H, BM) : 1) * H * 896 * cdiv(X + 1, 896)));
  TRYV(vae_alloc(g, &g->Pmv, size_t(ns) * 64 * 448));
  TRYV(vae_alloc(g, &g->P3, size_t(ns / cdiv(H, BM) > 0 ? ns / cdiv(H, BM) : 1) * H * 64));
  TRYV(vae_alloc(g, &g->P4, size_t(ns / cdiv(X, BM) > 0 ? ns / cdiv(X, BM) : 1) * X * 448));
  {
    cudaError_t e = g->arena.finalize(d->dtype_mode == GM_PREC_SPLIT);
    if (e != cudaSuccess) { rc = fail(c, GM_ERR_CUDA, "bf16 arena allocation failed: %s", cudaGetErrorString(e)); gm_vae_destroy(g); return rc; }
    g->allocs.push_back(g->arena.base);
    g->lo = g->arena.lo_off;
  }
#undef TRYV
  *out = g;
  return GM_OK;
}

// On-device epoch shuffling: gm_vae_grad(step) with gather_idx == NULL reads batch (step % batches_per_epoch)
// of the pseudo-random permutation of epoch (step / batches_per_epoch) over a resident pool (src/vae.py:150).
extern "C" int gm_vae_set_sampler(gm_vae* g, long long n_pool, long long batches_per_epoch, long long batch_size, uint64_t seed) {
  if (!g || n_pool < 0 || n_pool > 0x7FFFFFFFll || batches_per_epoch < 0 || batch_size < 0) return g ? fail(g->ctx, GM_ERR_ARG, "gm_vae_set_sampler: bad argument") : GM_ERR_ARG;
  g->pool_n = n_pool; g->pool_bpe = batches_per_epoch; g->pool_bs = batch_size; g->pool_seed = seed;
  return GM_OK;
}

extern "C" int gm_vae_param_count(const gm_vae* g) { return g ? g->total : GM_ERR_ARG; }

extern "C" int gm_vae_bind(gm_vae* g, float* p, float* gr, float* m, float* v) {
  if (!g || !p || !gr) return g ? fail(g->ctx, GM_ERR_ARG, "gm_vae_bind: bad argument") : GM_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(p) & 15) return fail(g->ctx, GM_ERR_ARG, "parameter buffer must be 16-byte aligned");
  g->par = p; g->grd = gr; g->am = m; g->av = v;
  g->plans.clear();
  return GM_OK;
}

static void vae_adam_segs(gm_vae* g, AdamParams& a) {
  a.lo_off = g->lo;
  a.total = g->total;
  a.nseg = 4;
  a.seg[0] = {g->off_w1, g->H * g->X, g->X, g->W1_s, g->X, nullptr, 0};
  a.seg[1] = {g->off_wmv, 2 * g->Z * g->H, g->H, g->Wmv_s, g->H, g->Wmv_t, 64};
  a.seg[2] = {g->off_w3, g->H * g->Z, g->Z, g->W3_s, g->ZP, g->W3_t, g->H};
  a.seg[3] = {g->off_w4, g->X * g->H, g->H, g->W4_s, g->H, g->W4_t, g->X};
}

extern "C" int gm_vae_sync_shadows(gm_vae* g, gm_stream stream) {
  if (!g) return GM_ERR_ARG;
  if (!g->par) return fail(g->ctx, GM_ERR_STATE, "not bound");
  AdamParams a;
  memset(&a, 0, sizeof a);
  a.p = g->par; a.update = 0;
  vae_adam_segs(g, a);
  launch_pdl("adam_kernel", adam_kernel, cdiv(a.total, 256), 256, 0, static_cast<cudaStream_t>(stream), a);
  g->ctx->launches++;
  CU_OK(g->ctx, cudaGetLastError());
  return GM_OK;
}

// the eps of the last gm_vae_grad / gm_vae_forward (caller tensor or in-kernel Philox, src/vae.py:104) -> out_dev [batch, z]
extern "C" int gm_vae_last_eps(gm_vae* g, float* out_dev, int batch, gm_stream stream) {
  if (!g || !out_dev || batch <= 0 || batch > g->Bmax) return g ? fail(g->ctx, GM_ERR_ARG, "gm_vae_last_eps: bad argument") : GM_ERR_ARG;
  CU_OK(g->ctx, cudaMemcpyAsync(out_dev, g->EPS, size_t(batch) * g->Z * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return GM_OK;
}

extern "C" int gm_vae_apply(gm_vae* g, const gm_adam_hp* hp, int step, gm_stream stream) {
  if (!g || !hp || step <= 0) return g ? fail(g->ctx, GM_ERR_ARG, "gm_vae_apply: bad argument") : GM_ERR_ARG;
  if (!g->par || !g->am || !g->av) return fail(g->ctx, GM_ERR_STATE, "not fully bound");
  AdamParams a;
  memset(&a, 0, sizeof a);
  a.p = g->par; a.g = g->grd; a.m = g->am; a.v = g->av;
  fill_adam(a, hp, step);
  vae_adam_segs(g, a);
  if (g->pend) { a.gather = 1; a.gout = g->grd; a.gsegs = g->pend_segs; g->pend = false; }
  launch_pdl("adam_kernel", adam_kernel, cdiv(a.total, 256), 256, 0, static_cast<cudaStream_t>(stream), a);
  g->ctx->launches++;
  CU_OK(g->ctx, cudaGetLastError());
  return GM_OK;
}

static void vae_flush_pending(gm_vae* g, cudaStream_t s) {
  if (!g->pend) return;
  launch_pdl("finalize_grads_kernel", finalize_grads_kernel, cdiv(g->pend_segs.total, 256), 256, 0, s, g->pend_segs, g->grd);
  g->ctx->launches++;
  g->pend = false;
}
// Lazy gradients (single-GPU fast path, like gm_gan_set_lazy_grads): gm_vae_grad leaves the split-K partials and the
// following gm_vae_apply gathers, stores the flat gradient and applies Adam in one kernel.
extern "C" int gm_vae_set_lazy_grads(gm_vae* g, int on, gm_stream stream) {
  if (!g) return GM_ERR_ARG;
  if (!on) vae_flush_pending(g, static_cast<cudaStream_t>(stream));
  g->lazy = on != 0;
  return GM_OK;
}
extern "C" int gm_vae_materialize_grads(gm_vae* g, gm_stream stream) {
  if (!g) return GM_ERR_ARG;
  vae_flush_pending(g, static_cast<cudaStream_t>(stream));
  CU_OK(g->ctx, cudaGetLastError());
  return GM_OK;
}

static void set_f32_epi(GemmParams& p, float* out, int ldp, const float* bias) {
  p.epi = EPI_F32; p.part = out; p.ldp = ldp; p.part_stride = 0; p.transpose = 0; p.bias = bias;
}

static int vae_plans(gm_vae* g, int B, VaePlans** out) {
  auto it = g->plans.find(B);
  if (it != g->plans.end()) { *out = &it->second; return GM_OK; }
  gm_ctx* c = g->ctx;
  PlanLoScope lo_scope(c, g->lo);
  VaePlans sp;
  const int X = g->X, H = g->H, Z = g->Z, XP = g->XP, HP = g->HP, ZP = g->ZP;
  const float* P = g->par;
  const int ms = c->num_sms;
  int rc;
  // encoder layer 1: H1 = relu(x W1^T + b1), ones column at H
  if ((rc = plan_gemm(c, &sp.e1, 0, B, H, X, g->Xin, XP, g->W1_s, X, HP, 1))) return rc;
  set_bf16_epi(sp.e1.p, g->H1, HP, HP, 1, P + g->off_b1, ACT_RELU);
  // both heads at once: [mu | log_var] = H1 [Wm; Wv]^T + [bm; bv]   (fp32 out)
  if ((rc = plan_gemm(c, &sp.e2, 0, B, 2 * Z, H, g->H1, HP, g->Wmv_s, H, 2 * Z, 1))) return rc;
  set_f32_epi(sp.e2.p, g->MULV, 64, P + g->off_bmv);
  // decoder layer 1 and 2; the training variant of layer 2 fuses the SSE loss and its gradient
  if ((rc = plan_gemm(c, &sp.d1, 0, B, H, rup(Z, 16), g->Zb, ZP, g->W3_s, ZP, HP, 1))) return rc;
  set_bf16_epi(sp.d1.p, g->H3, HP, HP, 1, P + g->off_b3, ACT_RELU);
  sp.d1.flops = 2.0 * B * H * Z;
  if ((rc = plan_gemm(c, &sp.d2, 0, B, X, H, g->H3, HP, g->W4_s, H, X, 1))) return rc;
  set_bf16_epi(sp.d2.p, g->DA4, XP, X, 0, P + g->off_b4, ACT_SIGMOID);
  sp.d2.p.aux = g->Xin; sp.d2.p.ld_aux = XP; sp.d2.p.aux_mode = AUX_VAE_OUT;
  sp.d2.p.dot_out = g->slots_r; sp.d2.p.dot_ld = g->Bmax;
  if ((rc = plan_gemm(c, &sp.d2_fwd, 0, B, X, H, g->H3, HP, g->W4_s, H, X, 1))) return rc;
  set_bf16_epi(sp.d2_fwd.p, g->DA4, XP, X, 0, P + g->off_b4, ACT_SIGMOID);
  // [dW4 | db4] = DA4^T [H3 | 1]
  if ((rc = plan_gemm(c, &sp.gw4, 1, X, H + 1, B, g->DA4, XP, g->H3, HP, H + 1, ms))) return rc;
  { GemmParams& p = sp.gw4.p; p.epi = EPI_F32; p.part = g->P4; p.ldp = 448; p.part_stride = (long long)X * 448; p.transpose = 0;
    sp.gw4.flops = 2.0 * X * H * B; }
  // DA3 = (DA4 W4) * 1[H3 > 0]
  if ((rc = plan_gemm(c, &sp.da3, 0, B, H, X, g->DA4, XP, g->W4_t, X, H, 1))) return rc;
  set_bf16_epi(sp.da3.p, g->DA3, HP, H, 0, nullptr, ACT_NONE);
  sp.da3.p.aux = g->H3; sp.da3.p.ld_aux = HP; sp.da3.p.aux_mode = AUX_RELU_MASK;
  // [dW3 | db3] = DA3^T [Zb | 1]
  if ((rc = plan_gemm(c, &sp.gw3, 1, H, Z + 1, B, g->DA3, HP, g->Zb, ZP, Z + 1, ms))) return rc;
  { GemmParams& p = sp.gw3.p; p.epi = EPI_F32; p.part = g->P3; p.ldp = 64; p.part_stride = (long long)H * 64; p.transpose = 0;
    sp.gw3.flops = 2.0 * H * Z * B; }
  // dz = DA3 W3   (fp32 out, N = z)
  if ((rc = plan_gemm(c, &sp.dz, 0, B, Z, H, g->DA3, HP, g->W3_t, H, Z, 1))) return rc;
  set_f32_epi(sp.dz.p, g->DZ, 32, nullptr);
  // [dWm; dWv | dbm; dbv] = DML^T [H1 | 1]
  if ((rc = plan_gemm(c, &sp.gwmv, 1, 2 * Z, H + 1, B, g->DML, 64, g->H1, HP, H + 1, ms))) return rc;
  { GemmParams& p = sp.gwmv.p; p.epi = EPI_F32; p.part = g->Pmv; p.ldp = 448; p.part_stride = (long long)64 * 448; p.transpose = 0;
    sp.gwmv.flops = 2.0 * 2 * Z * H * B; }
  // DA1 = (DML [Wm; Wv]) * 1[H1 > 0]
  if ((rc = plan_gemm(c, &sp.da1, 0, B, H, rup(2 * Z, 16), g->DML, 64, g->Wmv_t, 64, H, 1))) return rc;
  set_bf16_epi(sp.da1.p, g->DA1, HP, H, 0, nullptr, ACT_NONE);
  sp.da1.p.aux = g->H1; sp.da1.p.ld_aux = HP; sp.da1.p.aux_mode = AUX_RELU_MASK;
  sp.da1.flops = 2.0 * B * H * 2 * Z;
  // [dW1 | db1] = DA1^T [x | 1]
  if ((rc = plan_gemm(c, &sp.gw1, 1, H, X + 1, B, g->DA1, HP, g->Xin, XP, X + 1, ms))) return rc;
  { GemmParams& p = sp.gw1.p; p.epi = EPI_F32; p.part = g->P1; p.ldp = p.n_tiles * 448; p.part_stride = (long long)H * p.ldp; p.transpose = 0;
    sp.gw1.flops = 2.0 * H * X * B; }
  g->plans[B] = sp;
  *out = &g->plans[B];
  return GM_OK;
}

static int vae_forward_core(gm_vae* g, VaePlans* sp, const void* images, int fmt, const int* idx, int B, const float* eps,
                            uint64_t seed, uint64_t step, bool train, cudaStream_t s) {
  gm_ctx* c = g->ctx;
  int rc;
  Sampler smp = kNoSampler;
  if (!idx && g->pool_n > 0) {
    // batch (step mod batches_per_epoch) of this epoch's permutation: a true epoch like `for batch in train_iter` (src/vae.py:150)
    const uint64_t bpe = g->pool_bpe > 0 ? uint64_t(g->pool_bpe) : 1;
    smp = make_sampler(g->pool_n, g->pool_seed, step / bpe, (step % bpe) * uint64_t(g->pool_bs > 0 ? g->pool_bs : B));
  }
  launch_pdl("stage_images_kernel", stage_images_kernel, c->num_sms * 8, 256, 0, s, images, fmt, idx, g->Xin, B, g->X, g->XP, smp, g->lo);
  c->launches++;
  if ((rc = launch_plan(c, sp->e1, s))) return rc;
  if ((rc = launch_plan(c, sp->e2, s))) return rc;
  launch_pdl("vae_reparam_kernel", vae_reparam_kernel, cdiv(B * ((g->Z + 8) / 8), 256), 256, 0, s, g->MULV, 64, eps, g->EPS, g->Zb, g->ZP, B, g->Z, seed, step, g->part_k, g->lo);
  c->launches++;
  if ((rc = launch_plan(c, sp->d1, s))) return rc;
  if ((rc = launch_plan(c, train ? sp->d2 : sp->d2_fwd, s))) return rc;
  return GM_OK;
}

static void vae_losses(gm_vae* g, int B, cudaStream_t s) {
  const int nb = g->ctx->num_sms * 2;
  launch_pdl("vae_rowsum_kernel", vae_rowsum_kernel, nb, 256, 0, s, g->slots_r, 2 * cdiv(g->X, 208), g->Bmax, B, g->part_r);
  launch_pdl("vae_losses_final_kernel", vae_losses_final_kernel, 1, 256, 0, s, g->part_r, nb, g->part_k, cdiv(B * ((g->Z + 8) / 8), 256), g->losses);
  g->ctx->launches += 2;
}

// compute_batch + (recon + kl).backward()  (src/vae.py:157-161,193-212)
extern "C" int gm_vae_grad(gm_vae* g, const void* images, int img_fmt, const int* gather_idx, int batch,
                           const float* eps_dev, float grad_scale, uint64_t seed, uint64_t step, float* losses_dev,
                           gm_stream stream) {
  if (!g || !images) return GM_ERR_ARG;
  if (batch <= 0 || batch > g->Bmax) return fail(g->ctx, GM_ERR_ARG, "batch must be in (0, %d]", g->Bmax);
  if (!g->par || !g->grd) return fail(g->ctx, GM_ERR_STATE, "bind first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  gm_ctx* c = g->ctx;
  VaePlans* sp;
  int rc;
  if ((rc = vae_plans(g, batch, &sp))) return rc;
  const int B = batch;
  vae_flush_pending(g, s);      // a pending gradient lives in the partial buffers this call overwrites
  if ((rc = vae_forward_core(g, sp, images, img_fmt, gather_idx, B, eps_dev, seed, step, true, s))) return rc;
  vae_losses(g, B, s);
  (void)grad_scale;   // the reference's VAE losses are sums: data-parallel ranks SUM unscaled gradients
  if ((rc = launch_plan(c, sp->gw4, s))) return rc;
  if ((rc = launch_plan(c, sp->da3, s))) return rc;
  if ((rc = launch_plan(c, sp->gw3, s))) return rc;
  if ((rc = launch_plan(c, sp->dz, s))) return rc;
  launch_pdl("vae_dlatent_kernel", vae_dlatent_kernel, cdiv(B * 8, 256), 256, 0, s, g->MULV, 64, g->DZ, 32, g->EPS, g->DML, 64, B, g->Z, 1.f, g->lo);
  c->launches++;
  if ((rc = launch_plan(c, sp->gwmv, s))) return rc;
  if ((rc = launch_plan(c, sp->da1, s))) return rc;
  if ((rc = launch_plan(c, sp->gw1, s))) return rc;
  GradSegs gs;
  memset(&gs, 0, sizeof gs);
  const GemmParams &p1 = sp->gw1.p, &pm = sp->gwmv.p, &p3 = sp->gw3.p, &p4 = sp->gw4.p;
  gs.nseg = 8;
  gs.total = g->total;
  gs.s[0] = {g->off_w1, g->H * g->X, 0, g->X, p1.ldp, 0, p1.splits, p1.part_stride, g->P1};
  gs.s[1] = {g->off_b1, g->H, 2, 0, p1.ldp, g->X, p1.splits, p1.part_stride, g->P1};
  gs.s[2] = {g->off_wmv, 2 * g->Z * g->H, 0, g->H, pm.ldp, 0, pm.splits, pm.part_stride, g->Pmv};
  gs.s[3] = {g->off_bmv, 2 * g->Z, 2, 0, pm.ldp, g->H, pm.splits, pm.part_stride, g->Pmv};
  gs.s[4] = {g->off_w3, g->H * g->Z, 0, g->Z, p3.ldp, 0, p3.splits, p3.part_stride, g->P3};
  gs.s[5] = {g->off_b3, g->H, 2, 0, p3.ldp, g->Z, p3.splits, p3.part_stride, g->P3};
  gs.s[6] = {g->off_w4, g->X * g->H, 0, g->H, p4.ldp, 0, p4.splits, p4.part_stride, g->P4};
  gs.s[7] = {g->off_b4, g->X, 2, 0, p4.ldp, g->H, p4.splits, p4.part_stride, g->P4};
  if (g->lazy) { g->pend_segs = gs; g->pend = true; }
  else {
    launch_pdl("finalize_grads_kernel", finalize_grads_kernel, cdiv(gs.total, 256), 256, 0, s, gs, g->grd);
    c->launches++;
  }
  if (losses_dev) CU_OK(c, cudaMemcpyAsync(losses_dev, g->losses, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

// VAE.forward + losses without gradients (evaluate / reconstruct, src/vae.py:214-223,225-252)
extern "C" int gm_vae_forward(gm_vae* g, const void* images, int img_fmt, int n, const float* eps_dev, uint64_t seed,
                              uint64_t step, float* out_images_dev, float* mu_logvar_dev, float* losses_dev,
                              gm_stream stream) {
  if (!g || !images || n <= 0) return GM_ERR_ARG;
  if (n > g->Bmax) return fail(g->ctx, GM_ERR_ARG, "n (%d) exceeds max_batch", n);
  if (!g->par) return fail(g->ctx, GM_ERR_STATE, "bind first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VaePlans* sp;
  int rc;
  if ((rc = vae_plans(g, n, &sp))) return rc;
  if (losses_dev) {   // losses need the fused SSE epilogue (it overwrites DA4 with gradients)
    if ((rc = vae_forward_core(g, sp, images, img_fmt, nullptr, n, eps_dev, seed, step, true, s))) return rc;
    vae_losses(g, n, s);
    CU_OK(g->ctx, cudaMemcpyAsync(losses_dev, g->losses, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (out_images_dev) { if ((rc = launch_plan(g->ctx, sp->d2_fwd, s))) return rc; }
  } else {
    if ((rc = vae_forward_core(g, sp, images, img_fmt, nullptr, n, eps_dev, seed, step, false, s))) return rc;
  }
  if (out_images_dev) {
    const long long tot = (long long)n * g->X;
    launch_pdl("bf16_rows_to_f32_kernel", bf16_rows_to_f32_kernel, unsigned((tot + 255) / 256), 256, 0, s, g->DA4, g->XP, out_images_dev, n, g->X, g->lo);
    g->ctx->launches++;
  }
  if (mu_logvar_dev)
    CU_OK(g->ctx, cudaMemcpy2DAsync(mu_logvar_dev, size_t(2 * g->Z) * 4, g->MULV, 64 * 4, size_t(2 * g->Z) * 4, n,
                                    cudaMemcpyDeviceToDevice, s));
  CU_OK(g->ctx, cudaGetLastError());
  return GM_OK;
}

// Decoder.forward (src/vae.py:74-77): z [n, z_dim] fp32 -> images [n, image_size] fp32
extern "C" int gm_vae_decode(gm_vae* g, const float* z_dev, int n, float* out_images_dev, gm_stream stream) {
  if (!g || !z_dev || !out_images_dev || n <= 0) return GM_ERR_ARG;
  if (n > g->Bmax) return fail(g->ctx, GM_ERR_ARG, "n (%d) exceeds max_batch", n);
  if (!g->par) return fail(g->ctx, GM_ERR_STATE, "bind first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VaePlans* sp;
  int rc;
  if ((rc = vae_plans(g, n, &sp))) return rc;
  launch_pdl("stage_noise_kernel", stage_noise_kernel, cdiv(n * ((g->Z + 8) / 8), 256), 256, 0, s, z_dev, g->Zb, n, g->Z, g->ZP, 0, 0, g->lo, static_cast<const unsigned long long*>(nullptr));
  g->ctx->launches++;
  if ((rc = launch_plan(g->ctx, sp->d1, s))) return rc;
  if ((rc = launch_plan(g->ctx, sp->d2_fwd, s))) return rc;
  const long long tot = (long long)n * g->X;
  launch_pdl("bf16_rows_to_f32_kernel", bf16_rows_to_f32_kernel, unsigned((tot + 255) / 256), 256, 0, s, g->DA4, g->XP, out_images_dev, n, g->X, g->lo);
  g->ctx->launches++;
  CU_OK(g->ctx, cudaGetLastError());
  return GM_OK;
}
