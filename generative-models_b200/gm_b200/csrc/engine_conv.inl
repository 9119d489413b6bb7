// C-ABI entry points of the conv building blocks (conv_ops.cuh), included by engine.cu.  Stateless apart from a
// context-level scratch for the two-stage column reductions; the DCGAN engine that sequences them lives in
// gm_b200/dcgan.py (host language of the reference: Python).

static int conv_scratch(gm_ctx* c, size_t bytes, double** out) {
  if (bytes > c->red_bytes) {
    if (c->red) cudaFree(c->red);
    c->red = nullptr; c->red_bytes = 0;
    CU_OK(c, cudaMalloc(&c->red, bytes));
    c->red_bytes = bytes;
  }
  *out = static_cast<double*>(c->red);
  return GM_OK;
}

// grid of a grid-stride elementwise kernel: enough 256-thread blocks to fill the GPU (8 per SM), no more than the work needs
static unsigned conv_grid(const gm_ctx* c, unsigned long long items, int per_thread) {
  const unsigned long long want = (items + 256ull * per_thread - 1) / (256ull * per_thread);
  const unsigned long long cap = (unsigned long long)c->num_sms * 8;
  return unsigned(want < cap ? (want ? want : 1) : cap);
}

extern "C" int gm_im2col_k4s2(gm_ctx* c, const void* x, int B, int H, int W, int C, int ldx, void* col, int ldc, gm_stream stream) {
  if (!c || !x || !col || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (H & 1) || (W & 1)) return c ? fail(c, GM_ERR_ARG, "gm_im2col_k4s2: bad argument") : GM_ERR_ARG;
  if (ldc < 16 * C || ldx < C || (C % 8 == 0 && ((ldx % 8) || (ldc % 8)))) return fail(c, GM_ERR_ARG, "gm_im2col_k4s2: leading dimensions");
  const bool vec = C % 8 == 0;
  const unsigned long long rows = (unsigned long long)B * (H / 2) * (W / 2);
  const unsigned long long total = rows * 16 * (vec ? C / 8 : 1);
  if (total >= (1ull << 31)) return fail(c, GM_ERR_UNSUPPORTED, "gm_im2col_k4s2: more than 2^31 16-byte items (%llu): split the batch", total);
  const FastDiv dcg = make_fastdiv(vec ? C / 8 : 1), dwo = make_fastdiv(W / 2), dho = make_fastdiv(H / 2);
  const unsigned grid = conv_grid(c, total, kConvUnroll);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec)
    launch_pdl("im2col_k4s2_kernel", im2col_k4s2_kernel<true>, grid, 256, 0, s, static_cast<const __nv_bfloat16*>(x), H, W, C, ldx,
               static_cast<__nv_bfloat16*>(col), ldc, uint32_t(total), dcg, dwo, dho);
  else
    launch_pdl("im2col_k4s2_kernel", im2col_k4s2_kernel<false>, grid, 256, 0, s, static_cast<const __nv_bfloat16*>(x), H, W, C, ldx,
               static_cast<__nv_bfloat16*>(col), ldc, uint32_t(total), dcg, dwo, dho);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

extern "C" int gm_col2im_k4s2(gm_ctx* c, const void* col, int ldc, int B, int Hi, int Wi, int C, void* y, int ldy, int mode,
                              const void* aux, int ld_aux, float slope, gm_stream stream) {
  if (!c || !col || !y || B <= 0 || Hi <= 0 || Wi <= 0 || C <= 0 || mode < 0 || mode > 3) return c ? fail(c, GM_ERR_ARG, "gm_col2im_k4s2: bad argument") : GM_ERR_ARG;
  if (mode >= C2I_LRELU_GRAD && !aux) return fail(c, GM_ERR_ARG, "gm_col2im_k4s2: mode %d needs aux", mode);
  if (ldc < 16 * C || ldy < C) return fail(c, GM_ERR_ARG, "gm_col2im_k4s2: leading dimensions");
  const bool vec = C % 8 == 0;
  if (!vec && C > 8) return fail(c, GM_ERR_UNSUPPORTED, "gm_col2im_k4s2: C must be a multiple of 8 or below 8 (got %d)", C);
  if (vec && ((ldc % 8) || (ldy % 8) || (mode >= C2I_LRELU_GRAD && (ld_aux % 8)))) return fail(c, GM_ERR_ARG, "gm_col2im_k4s2: leading dimensions (multiples of 8)");
  const unsigned long long total = (unsigned long long)B * (2 * Hi) * (2 * Wi) * (vec ? C / 8 : 1);
  if (total >= (1ull << 31)) return fail(c, GM_ERR_UNSUPPORTED, "gm_col2im_k4s2: more than 2^31 items (%llu): split the batch", total);
  const FastDiv dcg = make_fastdiv(vec ? C / 8 : 1), dwo = make_fastdiv(2 * Wi), dho = make_fastdiv(2 * Hi);
  const unsigned grid = conv_grid(c, total, vec ? 2 : 1);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec)
    launch_pdl("col2im_k4s2_kernel", col2im_k4s2_kernel<true>, grid, 256, 0, s, static_cast<const __nv_bfloat16*>(col), ldc, Hi, Wi, C,
               static_cast<__nv_bfloat16*>(y), ldy, mode, static_cast<const __nv_bfloat16*>(aux), ld_aux, slope, uint32_t(total), dcg, dwo, dho);
  else
    launch_pdl("col2im_k4s2_kernel", col2im_k4s2_kernel<false>, grid, 256, 0, s, static_cast<const __nv_bfloat16*>(col), ldc, Hi, Wi, C,
               static_cast<__nv_bfloat16*>(y), ldy, mode, static_cast<const __nv_bfloat16*>(aux), ld_aux, slope, uint32_t(total), dcg, dwo, dho);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

// BatchNorm2d forward in training mode over x [rows, C] (NHWC rows): y = act(gamma xhat + beta); stats_dev [2][C] receives
// (mean, invstd) for the backward; running_dev [2][C] (nullable) is updated like torch (momentum, unbiased variance).
extern "C" int gm_bn_forward(gm_ctx* c, const void* x, long long rows, int C, int ld, const float* gamma, const float* beta, float eps,
                             int act, float slope, void* y, int ldy, float* stats_dev, float* running_dev, float momentum, gm_stream stream) {
  if (!c || !x || !y || !gamma || !beta || !stats_dev || rows <= 0 || C <= 0 || C % 8 || ld % 8 || ldy % 8)
    return c ? fail(c, GM_ERR_ARG, "gm_bn_forward: bad argument (C, ld multiples of 8)") : GM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nblk = c->num_sms * 4;
  double* part;
  int rc = conv_scratch(c, size_t(nblk) * 2 * C * sizeof(double), &part);
  if (rc) return rc;
  const int rpi = kBnThreads / (C / 8) > 0 ? kBnThreads / (C / 8) : 1;
  if (C / 8 > kBnThreads) return fail(c, GM_ERR_UNSUPPORTED, "gm_bn_forward: C <= %d", kBnThreads * 8);
  const size_t smem = size_t(rpi) * 2 * C * sizeof(double);
  static bool cfg = false;
  if (!cfg) { cudaFuncSetAttribute(bn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
              cudaFuncSetAttribute(bn_bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); cfg = true; }
  launch_pdl("bn_partial_kernel", bn_partial_kernel, nblk, kBnThreads, smem, s, static_cast<const __nv_bfloat16*>(x), rows, C, ld, part);
  launch_pdl("bn_finalize_kernel", bn_finalize_kernel, cdiv(C * 32, 256), 256, 0, s, static_cast<const double*>(part), nblk, C, double(rows), eps, stats_dev,
             running_dev, momentum);
  launch_pdl("bn_apply_kernel", bn_apply_kernel, conv_grid(c, (unsigned long long)rows * (C / 8), kBnUnroll), kBnThreads, 0, s, static_cast<const __nv_bfloat16*>(x), rows, C, ld,
             static_cast<const float*>(stats_dev), gamma, beta, act, slope, static_cast<__nv_bfloat16*>(y), ldy);
  c->launches += 3;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

// BatchNorm2d backward: dy = dL/d(act output), x = the layer's pre-normalisation input, stats from gm_bn_forward;
// writes dx [rows, lddx] and dgb_dev [2][C] = (dbeta, dgamma).
extern "C" int gm_bn_backward(gm_ctx* c, const void* dy, const void* x, long long rows, int C, int ld, const float* stats_dev,
                              const float* gamma, const float* beta, int act, float slope, void* dx, int lddx, float* dgb_dev,
                              gm_stream stream) {
  if (!c || !dy || !x || !dx || !gamma || !beta || !stats_dev || !dgb_dev || rows <= 0 || C <= 0 || C % 8 || ld % 8 || lddx % 8)
    return c ? fail(c, GM_ERR_ARG, "gm_bn_backward: bad argument") : GM_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nblk = c->num_sms * 2   /* 2 resident blocks per SM (register budget) */;
  double* part;
  int rc = conv_scratch(c, size_t(nblk) * 2 * C * sizeof(double), &part);
  if (rc) return rc;
  if (C / 8 > kBnThreads) return fail(c, GM_ERR_UNSUPPORTED, "gm_bn_backward: C <= %d", kBnThreads * 8);
  const int rpi = kBnThreads / (C / 8);
  const size_t smem = size_t(rpi) * 2 * C * sizeof(double);
  launch_pdl("bn_bwd_partial_kernel", bn_bwd_partial_kernel, nblk, kBnThreads, smem, s, static_cast<const __nv_bfloat16*>(dy),
             static_cast<const __nv_bfloat16*>(x), rows, C, ld, stats_dev, gamma, beta, act, slope, part);
  launch_pdl("bn_bwd_finalize_kernel", bn_bwd_finalize_kernel, cdiv(2 * C * 32, 256), 256, 0, s, static_cast<const double*>(part), nblk, C, dgb_dev);
  launch_pdl("bn_bwd_apply_kernel", bn_bwd_apply_kernel, conv_grid(c, (unsigned long long)rows * (C / 8), kBnBwdUnroll), kBnThreads, 0, s, static_cast<const __nv_bfloat16*>(dy),
             static_cast<const __nv_bfloat16*>(x), rows, C, ld, stats_dev, gamma, beta, act, slope, static_cast<const float*>(dgb_dev),
             float(1.0 / double(rows)), static_cast<__nv_bfloat16*>(dx), lddx);
  c->launches += 3;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

extern "C" int gm_cast_bf16(gm_ctx* c, const float* src, int R, int C, void* dst, int ld, void* dst_t, int ld_t, gm_stream stream) {
  if (!c || !src || R <= 0 || C <= 0 || (!dst && !dst_t)) return c ? fail(c, GM_ERR_ARG, "gm_cast_bf16: bad argument") : GM_ERR_ARG;
  const long long n = (long long)R * C;
  launch_pdl("cast_bf16_kernel", cast_bf16_kernel, unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), src, R, C,
             static_cast<__nv_bfloat16*>(dst), ld, static_cast<__nv_bfloat16*>(dst_t), ld_t);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

extern "C" int gm_pack_col0(gm_ctx* c, const float* v, int rows, void* out, int ld, gm_stream stream) {
  if (!c || !v || !out || rows <= 0 || ld <= 0) return c ? fail(c, GM_ERR_ARG, "gm_pack_col0: bad argument") : GM_ERR_ARG;
  const long long n = (long long)rows * ld;
  launch_pdl("pack_col0_kernel", pack_col0_kernel, unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), v, rows,
             static_cast<__nv_bfloat16*>(out), ld);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

// The per-variant adversarial loss + upstream gradient on a vector of logits (train_D: rows [0,B) real then [B,2B) fake;
// train_G: B fake rows) - loss_pass_kernel of the MLP path on logits that some other network produced (the conv D).
// ds_dev[rows] = dL/dlogit (scaled by inv_global_batch), d_out_dev (nullable) = D's outputs, loss_dev[0] = loss.
extern "C" int gm_loss_rows(gm_ctx* c, int variant, int out_act, const float* logits_dev, int batch, int g_step, float inv_global_batch,
                            float* ds_dev, float* d_out_dev, float* loss_dev, gm_stream stream) {
  if (!c || !logits_dev || !ds_dev || !loss_dev || batch <= 0) return c ? fail(c, GM_ERR_ARG, "gm_loss_rows: bad argument") : GM_ERR_ARG;
  if (variant == V_RA || variant == V_FISHER || variant == V_BEGAN || variant == V_INFO || variant == V_WGP || variant == V_DRA)
    return fail(c, GM_ERR_UNSUPPORTED, "gm_loss_rows: row-wise losses only (NS, MM, W, LS, f-GAN)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows = g_step ? batch : 2 * batch;
  const int nblk = cdiv(rows, kLossThreads) < c->num_sms * 2 ? cdiv(rows, kLossThreads) : c->num_sms * 2;
  double* ws;
  int rc = conv_scratch(c, size_t(3) * c->num_sms * 2 * 4 * sizeof(double) + 64, &ws);
  if (rc) return rc;
  if (!c->loss_zero) {
    CU_OK(c, cudaMalloc(&c->loss_zero, 64));
    CU_OK(c, cudaMemset(c->loss_zero, 0, 64));
  }
  LossParams lp;
  memset(&lp, 0, sizeof lp);
  lp.slots = logits_dev; lp.nslots = 1; lp.slot_ld = rows;
  lp.b2 = static_cast<const float*>(c->loss_zero);                       // the conv D's last layer has no bias
  lp.B = batch; lp.Bstat = batch; lp.g_step = g_step; lp.variant = variant; lp.out_act = out_act; lp.inv_b = inv_global_batch;
  lp.ds = ds_dev; lp.d_out = d_out_dev; lp.loss = loss_dev;
  lp.fisher = static_cast<float*>(c->loss_zero) + 4;
  lp.partA = ws; lp.partB = ws + size_t(c->num_sms) * 8; lp.partR = ws + size_t(c->num_sms) * 16;
  lp.nblk = nblk;
  lp.done = reinterpret_cast<unsigned int*>(static_cast<char*>(c->loss_zero) + 32);
  lp.ls_a = 0.f; lp.ls_b = 1.f; lp.ls_c = 1.f;
  launch_pdl("loss_pass_kernel<2>", loss_pass_kernel<2>, nblk, kLossThreads, 0, s, lp);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

// N(0,1) noise rows as the bf16 GEMM operand (compute_noise, src/ns_gan.py:218-220, as in-kernel Philox): out [rows, ld]
// with columns [0, z) noise, column z = 1 (unused by the conv generator), the rest 0; noise_dev != NULL converts a caller tensor.
extern "C" int gm_noise_rows(gm_ctx* c, const float* noise_dev, void* out, int rows, int z, int ld, uint64_t seed, uint64_t stream_id,
                             gm_stream stream) {
  if (!c || !out || rows <= 0 || z <= 0 || ld < z + 1 || ld % 8) return c ? fail(c, GM_ERR_ARG, "gm_noise_rows: bad argument") : GM_ERR_ARG;
  launch_pdl("stage_noise_kernel", stage_noise_kernel, cdiv(rows * ((z + 8) / 8), 256), 256, 0, static_cast<cudaStream_t>(stream), noise_dev,
             static_cast<__nv_bfloat16*>(out), rows, z, ld, (unsigned long long)seed, (unsigned long long)stream_id, 0ll,
             static_cast<const unsigned long long*>(nullptr));
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}

// images (gm_img_fmt: fp32 | u8 | 1-bit packed) -> bf16 rows [rows, ld] with a ones column at x (the bias-gradient trick of
// the MLP engines) - process_batch's flatten + to_cuda (src/ns_gan.py:222-226, src/ae.py:150-151) as a standalone step
extern "C" int gm_stage_images(gm_ctx* c, const void* images, int img_fmt, const int* gather_idx, void* out, int rows, int x, int ld,
                               gm_stream stream) {
  if (!c || !images || !out || rows <= 0 || x <= 0 || ld < x + 1 || ld % 8) return c ? fail(c, GM_ERR_ARG, "gm_stage_images: bad argument") : GM_ERR_ARG;
  launch_pdl("stage_images_kernel", stage_images_kernel, c->num_sms * 8, 256, 0, static_cast<cudaStream_t>(stream), images, img_fmt, gather_idx,
             static_cast<__nv_bfloat16*>(out), rows, x, ld, kNoSampler, 0ll);
  c->launches++;
  CU_OK(c, cudaGetLastError());
  return GM_OK;
}
