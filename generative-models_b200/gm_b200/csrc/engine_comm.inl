// Peer exchange for the fused all-reduce + Adam kernel (kernels.cuh: adam_allreduce_kernel).
// One cudaMalloc'ed region per rank = [2][16 source ranks][nfloats] exchange slots + flag array, exported with
// cudaIpcGetMemHandle; the host (gm_b200/parallel.py) gathers the 64-byte handles of all ranks
// through torch.distributed and gm_comm_open maps them (NVLink peer access on the B200 box).

extern "C" int gm_comm_create(gm_ctx* c, int nfloats, gm_comm** out) {
  if (!c || !out || nfloats <= 0) return GM_ERR_ARG;
  gm_comm* m = new gm_comm();
  m->ctx = c;
  m->nfloats = (long long)rup(nfloats, kCommChunk);
  m->nblocks = int(m->nfloats / kCommChunk);
  m->flag_off = size_t(2) * kCommMaxWorld * m->nfloats * sizeof(float);   // [parity][source rank][nfloats]: peers push their chunks here
  m->stat_off = m->flag_off + size_t(2) * kCommMaxWorld * m->nblocks * sizeof(unsigned long long);
  m->sflag_off = m->stat_off + size_t(2) * kCommMaxWorld * kCommStatVals * sizeof(double);
  const size_t bytes = m->sflag_off + size_t(2) * kCommMaxWorld * sizeof(unsigned long long);
  cudaError_t e = cudaMalloc(&m->base, bytes);
  if (e != cudaSuccess) { delete m; return fail(c, GM_ERR_CUDA, "gm_comm_create: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
  CU_OK(c, cudaMemset(m->base, 0, bytes));
  CU_OK(c, cudaDeviceSynchronize());
  *out = m;
  return GM_OK;
}

extern "C" int gm_comm_handle(gm_comm* m, void* out64) {
  if (!m || !out64) return GM_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CU_OK(m->ctx, cudaIpcGetMemHandle(&h, m->base));
  memcpy(out64, &h, 64);
  return GM_OK;
}

extern "C" int gm_comm_open(gm_comm* m, int rank, int world, const void* handles) {
  if (!m || !handles) return GM_ERR_ARG;
  if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world)
    return fail(m->ctx, GM_ERR_ARG, "gm_comm_open: rank %d / world %d (max %d ranks)", rank, world, kCommMaxWorld);
  m->rank = rank; m->world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { m->peer[r] = m->base; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + size_t(r) * 64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&m->peer[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail(m->ctx, GM_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (no peer access between the GPUs?)", r, cudaGetErrorString(e));
  }
  m->opened = true;
  return GM_OK;
}

extern "C" int gm_comm_destroy(gm_comm* m) {
  if (!m) return GM_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < m->world; ++r)
    if (r != m->rank && m->peer[r]) cudaIpcCloseMemHandle(m->peer[r]);
  if (m->base) cudaFree(m->base);
  delete m;
  return GM_OK;
}

// optimizer.step() of one net on every rank at once: SUM all-reduce of the flat gradient over the
// peer mappings fused with Adam (replaces dist.all_reduce + gm_gan_apply).  Every rank must call it
// with the same net / step, in the same order.
static int comm_dev(gm_gan* g, gm_comm* m, int total, CommDev& cm) {
  if (total > m->nfloats) return fail(g->ctx, GM_ERR_ARG, "exchange buffer too small (%lld < %d)", m->nfloats, total);
  memset(&cm, 0, sizeof cm);
  for (int r = 0; r < m->world; ++r) {
    cm.x[r] = static_cast<float*>(m->peer[r]);
    cm.f[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(m->peer[r]) + m->flag_off);
  }
  cm.rank = m->rank; cm.world = m->world; cm.nblocks = m->nblocks; cm.nfloats = m->nfloats;
  return GM_OK;
}

// First half of gm_gan_apply_allreduce: publish this rank's gradient of `net` to every peer and return - no waiting.  Work that
// does not depend on the update may be enqueued before the matching gm_gan_apply_allreduce (which then only waits, sums and
// applies Adam): the G step's generator forward under the D exchange, the next step's image staging under the G exchange.
extern "C" int gm_gan_exchange_begin(gm_gan* g, int net, gm_comm* m, gm_stream stream) {
  if (!g || net < 0 || net > 1 || !m) return g ? fail(g->ctx, GM_ERR_ARG, "gm_gan_exchange_begin: bad argument") : GM_ERR_ARG;
  if (!m->opened) return fail(g->ctx, GM_ERR_STATE, "gm_comm_open has not been called");
  if (m->begun[net]) return fail(g->ctx, GM_ERR_STATE, "exchange of net %d already begun", net);
  if (!g->grd[net]) return fail(g->ctx, GM_ERR_STATE, "net %d not bound", net);
  AdamParams a;
  memset(&a, 0, sizeof a);
  a.g = g->grd[net];
  a.total = net == GM_NET_G ? g->G.total : g->D.total;
  if (g->pend[net]) { a.gather = 1; a.gsegs = g->pend_segs[net]; g->pend[net] = false; }
  CommDev cm;
  int rc = comm_dev(g, m, a.total, cm);
  if (rc) return rc;
  cm.seq = ++m->seq;
  m->begun[net] = cm.seq;
  launch_pdl("adam_exchange_push_kernel", adam_exchange_push_kernel, cdiv(a.total, kCommChunk), 256, 0, static_cast<cudaStream_t>(stream), a, cm);
  g->ctx->launches++;
  CU_OK(g->ctx, cudaGetLastError());
  return GM_OK;
}

extern "C" int gm_gan_apply_allreduce(gm_gan* g, int net, const gm_adam_hp* hp, int step, gm_comm* m, gm_stream stream) {
  if (!g || net < 0 || net > 1 || !hp || step <= 0 || !m) return g ? fail(g->ctx, GM_ERR_ARG, "gm_gan_apply_allreduce: bad argument") : GM_ERR_ARG;
  if (!m->opened) return fail(g->ctx, GM_ERR_STATE, "gm_comm_open has not been called");
  if (!g->par[net] || !g->am[net] || !g->av[net] || !g->grd[net]) return fail(g->ctx, GM_ERR_STATE, "net %d not fully bound", net);
  AdamParams a;
  memset(&a, 0, sizeof a);
  a.p = g->par[net]; a.g = g->grd[net]; a.m = g->am[net]; a.v = g->av[net];
  fill_adam(a, hp, step);
  adam_segs(g, net, a);
  a.gout = g->grd[net];
  CommDev cm;
  int rc = comm_dev(g, m, a.total, cm);
  if (rc) return rc;
  if (m->begun[net]) {          // second half of a begun exchange: wait for the peers, sum, update
    cm.seq = m->begun[net];
    m->begun[net] = 0;
    launch_pdl("adam_exchange_finish_kernel", adam_exchange_finish_kernel, cdiv(a.total, kCommChunk), 256, 0, static_cast<cudaStream_t>(stream), a, cm);
  } else {
    if (g->pend[net]) { a.gather = 1; a.gsegs = g->pend_segs[net]; g->pend[net] = false; }
    cm.seq = ++m->seq;
    launch_pdl("adam_allreduce_kernel", adam_allreduce_kernel, cdiv(a.total, kCommChunk), 256, 0, static_cast<cudaStream_t>(stream), a, cm);
  }
  g->ctx->launches++;
  CU_OK(g->ctx, cudaGetLastError());
  return GM_OK;
}

// Batch statistics over the global batch: attach the communicator to the engine; gm_gan_d_grad then
// exchanges the RaNS / Fisher / DRAGAN / BEGAN partial sums between the ranks on the device
// (stats_exchange_kernel).  comm == NULL detaches (per-rank statistics).
extern "C" int gm_gan_attach_comm(gm_gan* g, gm_comm* m) {
  if (!g) return GM_ERR_ARG;
  if (m && !m->opened) return fail(g->ctx, GM_ERR_STATE, "gm_comm_open has not been called");
  g->comm = m;
  return GM_OK;
}
