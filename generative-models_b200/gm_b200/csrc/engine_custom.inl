// Custom-loss path (README.md:31: "edit train_D and train_G"): G and D as separately callable
// forward / backward pairs, so that a loss written in torch on DX_score / DG_score still trains
// through the same kernels.  The fused variants (gm_gan_d_grad / gm_gan_g_grad) never come here.
//
// D forward calls are numbered by `slot` (a row region of Xall / Aall / DHall; the regions the
// fused step uses for real / fake / xhat rows): the activations of up to gm_gan_num_slots()
// forwards stay live until their backward ran.  G keeps one set of activations.

namespace gm {
// ds[r] = dscore[r] * act'(s[r]) for one row region, sum(ds) -> db2[0].  One block.
__global__ void __launch_bounds__(1024) custom_ds_kernel(const float* __restrict__ slots, int nslots, int slot_ld,
                                                         const float* __restrict__ b2, int out_act,
                                                         const float* __restrict__ dscore, float* __restrict__ ds,
                                                         float* __restrict__ db2, int rows) {
  griddep_sync();
  __shared__ double sh[32];
  double acc = 0.0;
  for (int r = threadIdx.x; r < rows; r += 1024) {
    float s = b2[0];
    for (int k = 0; k < nslots; ++k) s += slots[(long long)k * slot_ld + r];
    const float d = act_out(s, out_act);
    const float v = dscore[r] * act_grad(s, d, out_act);
    ds[r] = v;
    acc += v;
  }
  acc = block_sum<1024>(acc, sh);
  if (threadIdx.x == 0) db2[0] = float(acc);
}

// Generator side: the saved fake rows (bf16) become da2 = dfake * fake * (1 - fake), in place.
__global__ void custom_da2_kernel(const float* __restrict__ dfake, __nv_bfloat16* __restrict__ fk, int rows, int x, int ld, long long lo_off) {
  griddep_sync();
  const int groups = ld / 8;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * groups) return;
  const int r = int(i / groups), c0 = int(i % groups) * 8;
  __nv_bfloat16* cell = fk + (long long)r * ld + c0;
  float f[8], v[8];
  load_bf16x8(cell, f, lo_off);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    v[j] = c < x ? dfake[(long long)r * x + c] * f[j] * (1.f - f[j]) : 0.f;
  }
  store_bf16x8(cell, v, lo_off);
}
}  // namespace gm

static int build_custom_plans(gm_gan* g, int B, CustomPlans** out) {
  auto it = g->cplans.find(B);
  if (it != g->cplans.end()) { *out = &it->second; return GM_OK; }
  gm_ctx* c = g->ctx;
  PlanLoScope lo_scope(c, g->lo);
  CustomPlans cp;
  const int X = g->X, H = g->H, XP = g->XP, HP = g->HP;
  const float* pD = g->par[GM_NET_D];
  const float* pG = g->par[GM_NET_G];
  int rc;
  for (int r = 0; r < g->nreg; ++r) {
    __nv_bfloat16* Xr = g->Xall + size_t(r) * g->Bmax * XP;
    __nv_bfloat16* Ar = g->Aall + size_t(r) * g->Bmax * HP;
    __nv_bfloat16* DHr = g->DHall + size_t(r) * g->Bmax * HP;
    // a = relu(x W1^T + b1), row-dot with w2 -> slots of this region (Discriminator.forward, src/ns_gan.py:57-60)
    if ((rc = plan_gemm(c, &cp.d1[r], 0, B, H, X, Xr, XP, g->W1d_s, X, H, 1))) return rc;
    set_bf16_epi(cp.d1[r].p, Ar, HP, H, 0, pD + g->D.off_b1, ACT_RELU);
    cp.d1[r].p.dot_w = pD + g->D.off_w2; cp.d1[r].p.dot_out = g->slots + size_t(r) * g->Bmax; cp.d1[r].p.dot_ld = g->nreg * g->Bmax;
    // [dW1 | db1]^T = [x | 1]^T dh over this region's rows
    if ((rc = plan_gemm(c, &cp.dw1[r], 1, X + 1, H, B, Xr, XP, DHr, HP, H, g->max_splits))) return rc;
    {
      GemmParams& p = cp.dw1[r].p;
      p.epi = EPI_F32; p.part = g->PD; p.ldp = p.m_tiles * BM; p.part_stride = (long long)H * p.ldp; p.transpose = 1;
      cp.dw1[r].flops = 2.0 * X * H * double(B);
    }
    // dx = dh W1 (fp32, to the caller: the gradient that flows on into G)
    if ((rc = plan_gemm(c, &cp.dx[r], 0, B, X, H, DHr, HP, g->W1d_t, H, X, 1))) return rc;
    {
      GemmParams& p = cp.dx[r].p;
      p.epi = EPI_F32; p.part = nullptr; p.ldp = X; p.part_stride = 0; p.transpose = 0; p.bias = nullptr;
    }
  }
  // fake = sigmoid(Hg W2g^T + b2g) kept in DA2 (turned into da2 in place by the backward)
  if ((rc = plan_gemm(c, &cp.g2, 0, B, X, H, g->Hg, HP, g->W2g_s, H, XP, 1))) return rc;
  set_bf16_epi(cp.g2.p, g->DA2, XP, X, 0, pG + g->G.off_b2, ACT_SIGMOID);
  auto ins = g->cplans.emplace(B, cp);
  *out = &ins.first->second;
  return GM_OK;
}

static int check_custom(gm_gan* g, int batch, gm_stream stream) {
  int rc = check_step_args(g, batch);
  if (rc) return rc;
  flush_pending(g, static_cast<cudaStream_t>(stream));   // lazy gradients of the fused step are formed before their partial buffers are reused
  if (g->d.variant == GM_BEGAN)
    return fail(g->ctx, GM_ERR_UNSUPPORTED, "the custom-loss path needs a scalar-output discriminator (BEGAN's D is an autoencoder)");
  return GM_OK;
}

extern "C" int gm_gan_num_slots(const gm_gan* g) { return g ? g->nreg : 0; }

extern "C" int gm_gan_d_forward(gm_gan* g, int slot, const float* x, int batch, float* scores, gm_stream stream) {
  int rc = check_custom(g, batch, stream);
  if (rc) return rc;
  if (!x || !scores) return fail(g->ctx, GM_ERR_ARG, "x / scores is null");
  if (slot < 0 || slot >= g->nreg) return fail(g->ctx, GM_ERR_ARG, "slot must be in [0, %d)", g->nreg);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CustomPlans* cp;
  if ((rc = build_custom_plans(g, batch, &cp))) return rc;
  gm_ctx* c = g->ctx;
  launch_pdl("stage_images_kernel", stage_images_kernel, c->num_sms * 8, 256, 0, s, static_cast<const void*>(x), int(GM_IMG_F32), static_cast<const int*>(nullptr),
             g->Xall + size_t(slot) * g->Bmax * g->XP, batch, g->X, g->XP, kNoSampler, g->lo);
  c->launches++;
  if ((rc = launch_plan(c, cp->d1[slot], s))) return rc;
  launch_pdl("scores_kernel", scores_kernel, cdiv(batch, 256), 256, 0, s, g->slots + size_t(slot) * g->Bmax, 2 * cdiv(g->H, 208), g->nreg * g->Bmax,
             g->par[GM_NET_D] + g->D.off_b2, g->d.d_out_act, scores, batch);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

extern "C" int gm_gan_d_backward(gm_gan* g, int slot, int batch, const float* dscore, float* dx, gm_stream stream) {
  int rc = check_custom(g, batch, stream);
  if (rc) return rc;
  if (!dscore) return fail(g->ctx, GM_ERR_ARG, "dscore is null");
  if (slot < 0 || slot >= g->nreg) return fail(g->ctx, GM_ERR_ARG, "slot must be in [0, %d)", g->nreg);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CustomPlans* cp;
  if ((rc = build_custom_plans(g, batch, &cp))) return rc;
  gm_ctx* c = g->ctx;
  const int B = batch, H = g->H, HP = g->HP;
  const float* w2 = g->par[GM_NET_D] + g->D.off_w2;
  float* ds = g->ds + size_t(slot) * g->Bmax;
  launch_pdl("custom_ds_kernel", custom_ds_kernel, 1, 1024, 0, s, g->slots + size_t(slot) * g->Bmax, 2 * cdiv(H, 208), g->nreg * g->Bmax,
             g->par[GM_NET_D] + g->D.off_b2, g->d.d_out_act, dscore, ds, g->lossbuf + 1, B);
  launch_pdl("dh_kernel", dh_kernel, g->dh_blocks, g->dh_threads, size_t(g->dh_rows_per_iter) * HP * sizeof(float), s,
             g->Aall + size_t(slot) * g->Bmax * HP, static_cast<const float*>(ds), w2, g->DHall + size_t(slot) * g->Bmax * HP, g->dw2p, B, H, HP,
             g->dh_rows_per_iter, g->lo, 0);
  launch_pdl("colsum_kernel", colsum_kernel, cdiv(HP * 32, 256), 256, 0, s, g->dw2p, g->dh_blocks, HP, HP, g->dw2sum);
  c->launches += 3;
  if ((rc = launch_plan(c, cp->dw1[slot], s))) return rc;
  GradSegs gs;
  memset(&gs, 0, sizeof gs);
  const GemmParams& pw = cp->dw1[slot].p;
  gs.nseg = 4;
  gs.total = g->D.total;
  gs.s[0] = {g->D.off_w1, g->H * g->X, 0, g->X, pw.ldp, 0, pw.splits, pw.part_stride, g->PD};
  gs.s[1] = {g->D.off_b1, g->H, 2, 0, pw.ldp, g->X, pw.splits, pw.part_stride, g->PD};
  gs.s[2] = {g->D.off_w2, g->H, 3, 0, 0, 0, 1, (long long)HP, g->dw2sum};
  gs.s[3] = {g->D.off_b2, 1, 3, 0, 0, 0, 1, 2, g->lossbuf + 1};
  launch_pdl("finalize_grads_kernel", finalize_grads_kernel, cdiv(gs.total, 256), 256, 0, s, gs, g->grd[GM_NET_D]);
  c->launches++;
  if (dx) {
    GemmPlan pl = cp->dx[slot];
    pl.p.part = dx;
    if ((rc = launch_plan(c, pl, s))) return rc;
  }
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

extern "C" int gm_gan_g_forward(gm_gan* g, const float* noise, int batch, float* images, gm_stream stream) {
  int rc = check_custom(g, batch, stream);
  if (rc) return rc;
  if (!noise || !images) return fail(g->ctx, GM_ERR_ARG, "noise / images is null");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StepPlans* sp;
  CustomPlans* cp;
  if ((rc = build_plans(g, batch, &sp))) return rc;
  if ((rc = build_custom_plans(g, batch, &cp))) return rc;
  gm_ctx* c = g->ctx;
  launch_pdl("stage_noise_kernel", stage_noise_kernel, cdiv(batch * ((g->Z + 8) / 8), 256), 256, 0, s, noise, g->Zb, batch, g->Z, g->ZP, uint64_t(0), uint64_t(0), g->lo, static_cast<const unsigned long long*>(nullptr));
  c->launches++;
  if ((rc = launch_plan(c, sp->g1, s))) return rc;
  if ((rc = launch_plan(c, cp->g2, s))) return rc;
  const long long tot = (long long)batch * g->X;
  launch_pdl("bf16_rows_to_f32_kernel", bf16_rows_to_f32_kernel, unsigned((tot + 255) / 256), 256, 0, s, static_cast<const __nv_bfloat16*>(g->DA2), g->XP, images, batch, g->X, g->lo);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

extern "C" int gm_gan_g_backward(gm_gan* g, int batch, const float* dimages, gm_stream stream) {
  int rc = check_custom(g, batch, stream);
  if (rc) return rc;
  if (!dimages) return fail(g->ctx, GM_ERR_ARG, "dimages is null");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StepPlans* sp;
  if ((rc = build_plans(g, batch, &sp))) return rc;
  gm_ctx* c = g->ctx;
  const long long cells = (long long)batch * (g->XP / 8);
  launch_pdl("custom_da2_kernel", custom_da2_kernel, unsigned((cells + 255) / 256), 256, 0, s, dimages, g->DA2, batch, g->X, g->XP, g->lo);
  c->launches++;
  if ((rc = launch_plan(c, sp->dw2g, s))) return rc;
  if ((rc = launch_plan(c, sp->dhg, s))) return rc;
  if ((rc = launch_plan(c, sp->dw1g, s))) return rc;
  began_finalize_g(g, sp, s, true);   // flat G gradient from the dW2g / dW1g partials (same layout for every variant)
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}
