// Persistent, warp-specialised tcgen05 GEMM for sm_100a (hand-written).
//
//   C[M,N] = sum_k A(m,k) * B(n,k)        bf16 operands, fp32 accumulation in TMEM
//
// Operand memory forms (row-major bf16 matrices in HBM, leading dim multiple of 8):
//   K-major  : matrix [MN rows, K cols]  (activations [batch, feat] as A; nn.Linear
//              weights [out, in] as B)            -> forward and dX GEMMs
//   MN-major : matrix [K rows, MN cols]  (contraction over the batch rows)
//                                                   -> dW GEMMs (A^T * B)
// Tiles: 128 x (BN1+BN2) x 64, 128-byte TMA/UMMA swizzle; BN2 > 0 issues two MMAs
// per k-step (N <= 256 each) into one wide accumulator.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> fused math -> HBM).  Pipelines:
// smem ring full/empty mbarriers (TMA <-> MMA), accumulator full/empty mbarriers
// (MMA <-> epilogue, double-buffered when 2*BN <= 512 TMEM columns), static
// round-robin tile scheduler (grid = #SMs).
#pragma once
#include "ptx.cuh"

namespace gm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kUmmaK = 16;
// epilogue warps EW (template parameter): 8 (2 per TMEM lane quarter) or 16.  Measured: the
// feed-bound long-K GEMMs (K = 784) want the extra smem pipeline stage 8 warps leave room for
// (D1: 83.0 -> 79.5 us), the epilogue-bound short-K ones (K <= 400, sigmoid / aux epilogues)
// want 16 warps (dX: 96.9 -> 84.9 us); the host picks per plan
constexpr int gemm_threads(int epi_warps) { return 64 + epi_warps * 32; }
constexpr int kSmemBudget = 225 * 1024;

enum : int { EPI_BF16 = 0, EPI_F32 = 1 };
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_LRELU = 3 };   // LeakyReLU(act_slope): universal epilogue only
// AUX_NONZERO_MASK: like AUX_RELU_MASK for an aux that holds the SIGNED mask form w2 * relu'(a) (nonzero <=> active unit)
enum : int { AUX_NONE = 0, AUX_SIGMOID_GRAD = 1, AUX_RELU_MASK = 2, AUX_VAE_OUT = 3, AUX_L1 = 4, AUX_NONZERO_MASK = 5 };

struct GemmParams {
  int M, N, K;          // logical extents; K counts contraction elements
  int m_tiles, n_tiles;
  int m_supers;         // ceil(m_tiles / cluster size): clusters walk (split, m_super, n_tile)
  int splits, kblocks, kb_per_split;
  int epi;
  // ---- EPI_BF16: out[m, n] = bf16( f(acc + bias[n]) * g(aux[m,n]) ), row-major, ld = ldo
  __nv_bfloat16* out;
  int ldo;
  int out_cols;         // columns [N, out_cols) are padding: zero, except col N = 1 if pad_one
  int pad_one;
  const float* bias;    // nullable
  int act;
  const __nv_bfloat16* aux;  // nullable, same [m, n] indexing, ld = ld_aux
  int ld_aux;
  int aux_mode;
  const float* dot_w;   // nullable: row-dot of the *stored* values with dot_w[n]
  int dot_sq;           // 1: dot_out = row sum of squares instead (dot_w unused)
  // AUX_L1 (BEGAN): out = sign(v - aux) * (row < row_split ? row_scale[0] : row_scale[1]), dot_out = sum |v - aux|
  const float* row_scale;
  int row_split;
  // dot_mask: with the row-dot, store dot_w[n] * 1[v > 0] instead of v (the G step needs only
  // M = w2 * relu'(a1) of D's hidden layer: dL/dx = ds * (M W1), src/ns_gan.py:57-60 backward)
  int dot_mask;
  // dot_mask == 2 (DOT_T 4; WGAN-GP's D forward over [real; fake; x_hat] rows): rows >= mask_row0 store the mask form
  // U = w2 * relu'(a1) (what the penalty's first gradient needs) at out_alt + (row - mask_row0) * ldo,
  // the rows below keep the activations at out + row * ldo
  // dot_mask == 3 (DOT_T 5; WGAN-GP's D forward when the x_hat rows are never materialised): the row-dot runs on relu(v)
  // but the PRE-activation v is stored (the penalty's mask is 1[eps pre_real + (1-eps) pre_fake > 0], gp_hat_kernel)
  int mask_row0;
  __nv_bfloat16* out_alt;
  // tma_store: full 32-column blocks of the bf16 output leave through the output tensor map
  // (cp.async.bulk.tensor store of the warp's swizzled staging tile) instead of LDS + STG
  int tma_store;
  // row_vec (AUX_SIGMOID_GRAD): additional per-row factor, out = v * row_vec[m] * aux (1 - aux)
  const float* row_vec;
  float* dot_out;       // partial slots [(n_tile*2 + half) * dot_ld + m]
  int dot_ld;
  // ---- EPI_F32: part[split*part_stride + (transpose ? n*ldp + m : m*ldp + n)] = acc
  float* part;
  long long part_stride;
  int ldp;
  int transpose;
  // optional phase timing (tools/time_phases.py): CTA 0 writes SM-clock stamps, see the kernel
  long long* dbg;
  // ---- split-bf16 operands (SPLIT kernels, gm_prec GM_PREC_SPLIT): A = A_hi + A_lo, B = B_hi + B_lo as bf16 planes
  // lo_off elements apart; the contraction runs over nparts = 3 operand pairs (hi,hi), (hi,lo), (lo,hi) of kb_part
  // k-blocks each into the same accumulator (kblocks = 3 kb_part; the dropped lo*lo term is below fp32 resolution).
  // bf16 outputs and aux inputs carry their residual plane at the same offset.
  int nparts, kb_part;
  long long lo_off;
  float act_slope;      // ACT_LRELU
};

// per-epilogue-warp staging tile for one 32-column block: 32 rows x (64 B data + 16 B pad):
// conflict-free row-wise writes (thread = row), row-contiguous reads (4 lanes = 64 B of a row)
constexpr int kEpiCols = 32;
constexpr int kEpiPitch = 80;
constexpr int kEpiStageBytes = 32 * kEpiPitch;
constexpr int kEpiVecBlocks = 8;                       // column blocks per warp the bias/dot staging holds
constexpr int kEpiVecBytes = kEpiVecBlocks * kEpiCols * 4 * 2;   // bias + row-dot weights

// Phase timing (tools/time_phases.py) is compiled in only with -DGM_PHASE_TIMING: the clock
// reads cost 8+ registers and CS2R stalls in the 96-register epilogue (profiles/r1e).
#ifdef GM_PHASE_TIMING
__device__ __forceinline__ long long phase_clock() { return clock64(); }
constexpr bool kPhaseTiming = true;
#else
__device__ __forceinline__ long long phase_clock() { return 0; }
constexpr bool kPhaseTiming = false;
#endif

// VEC_EXTRA: additional per-warp bytes behind the bias / row-dot staging (EpiVecExtra: room for the cp.async aux tile of
// the epilogues that have a bias AND an aux operand - VAE decoder output, BEGAN decoder)
template <int AUX_T, int BIAS_T, int DOT_T, int EW>
struct EpiVecExtra { static constexpr int value = (AUX_T > 0 && BIAS_T > 0 && DOT_T == 0 && EW == 16) ? 512 : 0; };
template <int BN1, int BN2, bool STAGED_EPI = true, bool PAIR = false, int EW = 8, int VEC_EXTRA = 0>
struct GemmCfg {
  static constexpr int BN = BN1 + BN2;
  static constexpr int NACC = (2 * BN <= 512) ? 2 : 1;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // a CTA pair splits the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // per-warp staging tile + per-warp copies of its blocks' bias / row-dot weight slices
  static constexpr int EPI_WARPS = EW;
  static_assert(EW == 8 || (EW == 16 && STAGED_EPI), "8 or 16 epilogue warps");
  static constexpr int VEC_BYTES = kEpiVecBytes + VEC_EXTRA;
  static constexpr int EPI_BYTES = STAGED_EPI ? EPI_WARPS * (kEpiStageBytes + VEC_BYTES) : 0;
  static_assert(!STAGED_EPI || (BN + kEpiCols - 1) / kEpiCols <= kEpiVecBlocks * (EPI_WARPS / 8), "bias/dot staging too small");
  static constexpr int STAGES_RAW = (kSmemBudget - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512 + EPI_BYTES;   // 512: barriers; staging tiles stay 512-B aligned (TMA 64B swizzle)
  // K-major B is loaded with boxes of BOXN rows (<= 256, divides BN1 and BN2)
  static constexpr int gcd(int a, int b) { return b == 0 ? a : gcd(b, a % b); }
  static constexpr int BOXN = (BN2 == 0) ? BN1 : gcd(BN1, BN2);
  static_assert(BN1 % 16 == 0 && BN2 % 16 == 0 && BN1 <= 256 && BN2 <= 256, "UMMA N constraint");
  static_assert(BN <= 512, "TMEM has 512 columns");
  static_assert(STAGES >= 2, "need a pipeline");
};

// Epilogue specialisation: ACT_T / AUX_T / BIAS_T / DOT_T >= 0 fix the fused epilogue at
// compile time (small code: the whole kernel must stay inside the instruction cache);
// -1 selects the universal variant that reads the choice from GemmParams at run time.
// CS: cluster size (1 or 2).  CS = 2 pairs two CTAs on consecutive m-tiles of the same
// (split, n-tile):
//  * K-major kernels: ONE tcgen05.mma.cta_group::2 (M = 256) per k-step spans both SMs; each
//    CTA stages its own 128 A rows and HALF of the B tile, so the bytes an SM has to pull into
//    shared memory per MMA cycle drop from A+B to A+B/2 (the measured limiter of the 1-CTA
//    kernel: the MMA thread waits on TMA, profiles/r1c).  The leader CTA issues the MMAs; TMA
//    completions of both CTAs are signalled on the leader's full barrier, slots / accumulators
//    are released with multicast commits, peer epilogue warps release the accumulator remotely.
//  * MN-major kernels: each CTA TMA-multicasts half of the shared B tile to both CTAs.
template <int BN1, int BN2, bool A_MN, bool B_MN, int ACT_T = -1, int AUX_T = -1, int BIAS_T = -1, int DOT_T = -1, int CS = 1, int EW = 8,
          bool SPLIT = false>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB2, const GemmParams p) {
  static_assert(!SPLIT || ACT_T < 0, "split operands: universal epilogue only");
  constexpr bool PAIR = (CS == 2) && !A_MN;
  using Cfg = GemmCfg<BN1, BN2, !A_MN, PAIR, EW, EpiVecExtra<AUX_T, BIAS_T, DOT_T, EW>::value>;
  constexpr int BN = Cfg::BN;
  constexpr int NACC = Cfg::NACC;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int kEpiWarps = Cfg::EPI_WARPS;
  constexpr int EPI_ARRIVALS = (NACC == 2) ? kEpiWarps / 2 : kEpiWarps;
  constexpr int kParts = kEpiWarps / 8;   // column partitions per accumulator stage (K-major kernels)
  static_assert(!B_MN || (BN1 % 64 == 0 && BN2 % 64 == 0), "MN-major B needs 64-wide atoms");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + NACC + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * NACC);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = CS > 1 ? int(cluster_ctarank()) : 0;
  constexpr uint16_t kMcMask = uint16_t((1u << CS) - 1u);
  static_assert(CS == 1 || CS == 2, "cluster size 1 or 2");
  static_assert(CS == 1 || B_MN || ((BN / CS) % 8 == 0 && Cfg::BOXN == BN), "B slice must keep the 8-row swizzle atoms whole");
  static_assert(!PAIR || BN2 == 0, "pair mode: one MMA per k-step");
  const bool pair_leader = !PAIR || crank == 0;

  griddep_launch();   // the next grid may start its prologue as soon as SMs free up
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (SPLIT) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    if (p.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), PAIR ? 2 : 1);          // pair: both producers arrive on the leader's barrier
      mbar_init(empty_bar(s), PAIR ? 1 : CS);        // multicast mode: every CTA releases the slot
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), PAIR ? 2 * EPI_ARRIVALS : EPI_ARRIVALS);   // pair: both CTAs' epilogues
    }
    fence_mbar_init();
  }
  if (warp == 1) { if constexpr (PAIR) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();   // peers' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_wait();     // barrier init / TMEM alloc above overlapped the previous grid's tail

  // work items are walked per CLUSTER: item -> (split, m_super, n_tile); this CTA's m-tile is
  // m_super*CS + rank (an m-tile past the end just loads zero rows and stores nothing)
  const int tiles = p.m_supers * p.n_tiles;
  const int total = tiles * p.splits;
  const int first_item = blockIdx.x / CS;
  const int item_step = gridDim.x / CS;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // The whole warp walks the loop (warp-uniform control flow and addresses keep the TMA
    // operands in uniform registers); one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    const bool leader = elect_one();
    for (int item = first_item; item < total; item += item_step) {
      const int split = item / tiles;
      const int rem = item - split * tiles;
      const int m0 = ((rem / p.n_tiles) * CS + crank) * BM;
      const int n0 = (rem % p.n_tiles) * BN;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (leader) {
          const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t b_dst = a_dst + Cfg::A_BYTES;
          int kk = kb, opart = 0;
          if constexpr (SPLIT) { opart = kb / p.kb_part; kk = kb - opart * p.kb_part; }
          const CUtensorMap* const ta = (SPLIT && opart == 2) ? &tmA2 : &tmA;   // (hi,hi), (hi,lo), (lo,hi)
          const CUtensorMap* const tb = (SPLIT && opart == 1) ? &tmB2 : &tmB;
          const int k0 = kk * BK;
          if constexpr (PAIR) {
            // own A rows + own half of B into own smem; bytes of BOTH CTAs are counted on the
            // leader's full barrier (leader expects 2 x stage bytes, the peer just arrives)
            const uint32_t fb = mapa_cluster(full_bar(stage), 0);
            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            else mbar_arrive_cluster(fb);
            tma_load_2d_pair(a_dst, ta, fb, k0, m0);
            tma_load_2d_pair(b_dst, tb, fb, k0, n0 + crank * (BN / 2));
          } else {
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, ta, full_bar(stage), k0, m0);
          } else {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a)
              tma_load_2d(a_dst + a * (BK * 128), ta, full_bar(stage), m0 + a * 64, k0);
          }
          if constexpr (CS == 1) {
            if constexpr (!B_MN) {
#pragma unroll
              for (int b = 0; b < BN / Cfg::BOXN; ++b)
                tma_load_2d(b_dst + b * (Cfg::BOXN * 128), tb, full_bar(stage), k0, n0 + b * Cfg::BOXN);
            } else {
#pragma unroll
              for (int b = 0; b < BN / 64; ++b)
                tma_load_2d(b_dst + b * (BK * 128), tb, full_bar(stage), n0 + b * 64, k0);
            }
          } else {
            // this CTA's share of the B tile, multicast to the whole cluster
            if constexpr (!B_MN) {
              constexpr int SL = BN / CS;   // rows per CTA (tensor-map box = SL rows)
              tma_load_2d_mc(b_dst + crank * (SL * 128), tb, full_bar(stage), k0, n0 + crank * SL, kMcMask);
            } else {
              for (int b = crank; b < BN / 64; b += CS)
                tma_load_2d_mc(b_dst + b * (BK * 128), tb, full_bar(stage), n0 + b * 64, k0, kMcMask);
            }
          }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // Whole warp in the loop, one elected lane issues tcgen05.mma / commit.  Descriptors are
    // a constant high word plus a 14-bit address field advanced by adds (no per-MMA rebuild).
    constexpr uint32_t idesc1 = make_idesc_bf16(PAIR ? 2 * BM : BM, BN1, A_MN, B_MN);
    constexpr uint32_t idesc2 = make_idesc_bf16(BM, BN2 > 0 ? BN2 : 16, A_MN, B_MN);
    // descriptor strides: K-major: SBO = 8 rows * 128 B; MN-major: LBO = atom stride
    // (BK*128 B), SBO = 8 k-rows * 128 B.  Per UMMA_K (16) advance: K-major 32 B,
    // MN-major 16 rows * 128 B.
    constexpr uint32_t A_LBO = A_MN ? BK * 128 : 0, A_KADV = A_MN ? kUmmaK * 128 : kUmmaK * 2;
    constexpr uint32_t B_LBO = B_MN ? BK * 128 : 0, B_KADV = B_MN ? kUmmaK * 128 : kUmmaK * 2;
    constexpr uint32_t B2_OFF = B_MN ? (BN1 / 64) * (BK * 128) : BN1 * 128;
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version 1, 128B swizzle
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int acc_iter = 0;
    for (int item = first_item; pair_leader && item < total; item += item_step, ++acc_iter) {
      const int split = item / tiles;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
      const int as = acc_iter % NACC;
      const uint32_t aphase = (acc_iter / NACC) & 1;
      const long long tm0 = phase_clock();
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const long long tm1 = phase_clock();
      long long full_wait = 0;
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        const long long tw0 = phase_clock();
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        full_wait += phase_clock() - tw0;
        const uint32_t a_src = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t b_src = a_src + Cfg::A_BYTES;
        int kk = kb;
        if constexpr (SPLIT) kk = kb % p.kb_part;
        const int krem = p.K - kk * BK;
        const int nk = krem >= BK ? BK / kUmmaK : (krem + kUmmaK - 1) / kUmmaK;
        if (leader) {
          uint32_t a_lo = ((a_src >> 4) & 0x3FFFu) | ((A_LBO >> 4) << 16);
          uint32_t b_lo = ((b_src >> 4) & 0x3FFFu) | ((B_LBO >> 4) << 16);
          uint32_t acc = kb > kb0 ? 1u : 0u;
          for (int k = 0; k < nk; ++k) {
            const uint64_t da = (uint64_t(DESC_HI) << 32) | a_lo;
            const uint64_t db = (uint64_t(DESC_HI) << 32) | b_lo;
            if constexpr (PAIR) umma_bf16_pair(d_tmem, da, db, idesc1, acc);
            else umma_bf16(d_tmem, da, db, idesc1, acc);
            if constexpr (BN2 > 0) {
              const uint64_t db2 = (uint64_t(DESC_HI) << 32) | (b_lo + (B2_OFF >> 4));
              umma_bf16(d_tmem + BN1, da, db2, idesc2, acc);
            }
            acc = 1u;
            a_lo += A_KADV >> 4;
            b_lo += B_KADV >> 4;
          }
          if constexpr (PAIR) umma_commit_pair(empty_bar(stage), kMcMask);   // frees the slot in both CTAs
          else if constexpr (CS == 1) umma_commit(empty_bar(stage));           // smem slot reusable once these MMAs retire
          else umma_commit_mc(empty_bar(stage), kMcMask);                      // ... in every CTA that multicasts into it
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) {                                 // accumulator complete -> epilogue (of both CTAs in pair mode)
        if constexpr (PAIR) umma_commit_pair(tfull_bar(as), kMcMask);
        else umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (kPhaseTiming && p.dbg != nullptr && blockIdx.x == 0 && acc_iter < 16 && leader) {
        long long* d = p.dbg + acc_iter * 4;   // [tile][wait accumulator free, issue loop, of which waiting for TMA, start stamp]
        d[0] = tm1 - tm0; d[1] = phase_clock() - tm1; d[2] = full_wait; d[3] = tm0;
      }
    }
  } else {
    // =========================== epilogue ===========================
    // 8 warps = 2 groups of 4 (one warp per TMEM lane quarter).  With a double-buffered
    // accumulator the groups alternate tiles (group g owns accumulator stage g), each
    // warp covering all columns of its 32 rows; with a single wide accumulator both
    // groups work on the same tile and split its 16-column chunks by parity.
    const int e = warp - 2;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int grp = e >> 2;         // 0..kEpiWarps/4-1
    const int my_stage = grp & 1;   // accumulator stage this warp serves when NACC == 2
    const int part = A_MN ? 0 : (grp >> 1);   // which interleaved set of 32-column blocks (K-major kernels)
    const uint32_t stage_s = bar_base + 512u + uint32_t(e) * kEpiStageBytes;  // staging tile (NT kernels)
    const uint32_t vec_s = bar_base + 512u + kEpiWarps * kEpiStageBytes + uint32_t(e) * Cfg::VEC_BYTES;
    const int act = ACT_T >= 0 ? ACT_T : p.act;
    const int aux_mode = AUX_T >= 0 ? AUX_T : p.aux_mode;
    const bool has_bias = BIAS_T >= 0 ? (BIAS_T != 0) : (p.bias != nullptr);
    const bool has_dot = DOT_T >= 0 ? (DOT_T == 1 || DOT_T == 3 || DOT_T == 4 || DOT_T == 5) : (p.dot_w != nullptr);
    const bool store_pre = DOT_T >= 0 ? (DOT_T == 5) : (p.dot_mask == 3);   // ReLU only inside the row-dot
    constexpr bool kRowMask = DOT_T == 4 || DOT_T < 0;   // per-row choice between activation and mask output
    const bool mask_all = DOT_T >= 0 ? (DOT_T == 3) : (p.dot_mask == 1);
    const bool mask_rows = DOT_T >= 0 ? (DOT_T == 4) : (p.dot_mask == 2);
    const bool has_sq = DOT_T >= 0 ? (DOT_T == 2) : (p.dot_sq != 0);
    const bool tma_st = !A_MN && p.tma_store != 0;
    bool st_pending = false;   // a bulk store may still be reading this warp's staging tile
    auto stage_acquire = [&]() {
      if (st_pending) {
        if (lane == 0) bulk_wait_read();
        __syncwarp();
        st_pending = false;
      }
    };
    int acc_iter = 0;
#pragma unroll 1
    for (int item = first_item; item < total; item += item_step, ++acc_iter) {
      const int as = acc_iter % NACC;
      if (NACC == 2 && as != my_stage) continue;
      const int split = item / tiles;
      const int rem = item - split * tiles;
      const int n_tile = rem % p.n_tiles;
      const int m0 = ((rem / p.n_tiles) * CS + crank) * BM;
      const int n0 = n_tile * BN;
      const uint32_t aphase = (acc_iter / NACC) & 1;
      const int wrow0 = m0 + quarter * 32;   // first row of this warp
      const int row = wrow0 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN;
      const bool mask_out = mask_all || (kRowMask && mask_rows && row >= p.mask_row0);
      constexpr bool kUniversal = ACT_T < 0;
      if (!A_MN && !(kUniversal && p.epi == EPI_F32)) {
       if constexpr (!A_MN) {
        // ================= bf16 epilogue (K-major kernels) =================
        // coalesced lane mapping for aux reads / output writes: 4 lanes x 16 B = one 64-byte
        // row segment, 8 rows per pass, 4 passes
        const int lr = lane >> 2, lc = lane & 3;
        constexpr int kBlocks = (BN + kEpiCols - 1) / kEpiCols;
        // aux tile (32 rows x 32 columns of the warp's next block):
        //  * kernels without bias / row-dot (dX, dHg, penalty T): cp.async straight into the
        //    warp's (otherwise unused) bias staging area, 64-byte rows with the 16-byte chunks
        //    XOR-swizzled by (row >> 1) & 3 so that the row-per-lane reads are conflict-free.
        //    No registers held across the block, no STS (the 96-register budget of the
        //    16-warp epilogue was spilling / re-reading S2R and constants, profiles/r1e);
        //  * otherwise: coalesced LDG into registers, transposed through the staging tile.
        //    With a bias (VAE decoder output, BEGAN decoder; 16 epilogue warps) the tile sits behind the warp's four bias
        //    blocks in an enlarged staging area: the register path (16 registers held across the block) spilled there.
        constexpr bool kAuxAsync = (AUX_T > 0) && (DOT_T == 0) && (BIAS_T == 0 || EpiVecExtra<AUX_T, BIAS_T, DOT_T, EW>::value > 0);
        const uint32_t auxt_s = vec_s + (BIAS_T > 0 ? 512u : 0u);   // 32 rows x 64 B, chunks XOR-swizzled
        uint4 pre[kAuxAsync ? 1 : 4];
        auto aux_fetch = [&](int c_first) {
          const int c = c_first + lc * 8;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rl = it * 8 + lr;
            const int r = wrow0 + rl;
            const bool ok = r < p.M && c < p.out_cols;
            if constexpr (kAuxAsync) {
              const __nv_bfloat16* src = ok ? p.aux + size_t(r) * p.ld_aux + c : p.aux;
              cp_async16_zfill(auxt_s + rl * 64 + ((lc ^ ((rl >> 1) & 3)) << 4), src, ok ? 16u : 0u);
            } else {
              pre[it] = make_uint4(0, 0, 0, 0);
              if (ok) pre[it] = __ldg(reinterpret_cast<const uint4*>(p.aux + size_t(r) * p.ld_aux + c));
            }
          }
          if constexpr (kAuxAsync) cp_async_commit();
        };
        // bias / row-dot weights of this warp's column blocks -> smem and the first aux tile ->
        // registers while the MMAs still run: no global-load latency after the TMEM read
        if (has_bias || has_dot) {
#pragma unroll
          for (int pass = 0; pass < kEpiVecBlocks / 4; ++pass) {
            const int blk = pass * 4 + (lane >> 3), c4 = (lane & 7) * 4;   // 4 blocks x 8 float4 per pass
            const int c = n0 + (part + blk * kParts) * kEpiCols + c4;
            uint4 bz = make_uint4(0, 0, 0, 0), wz = bz;
            if (part + blk * kParts < kBlocks && c < p.N) {
              if (has_bias) {
                bz = __ldg(reinterpret_cast<const uint4*>(p.bias + c));
                if (act == ACT_SIGMOID) {   // sigmoid(x + b) = 0.5 tanh(0.5 x + 0.5 b) + 0.5: stage 0.5 b, one FFMA later
                  bz.x = __float_as_uint(0.5f * __uint_as_float(bz.x)); bz.y = __float_as_uint(0.5f * __uint_as_float(bz.y));
                  bz.z = __float_as_uint(0.5f * __uint_as_float(bz.z)); bz.w = __float_as_uint(0.5f * __uint_as_float(bz.w));
                }
              }
              if (has_dot) wz = __ldg(reinterpret_cast<const uint4*>(p.dot_w + c));
            }
            if (blk < kEpiVecBlocks / kParts) {   // a warp of the 16-warp epilogue owns at most 4 blocks: the rest of the area may hold its aux tile
              if (has_bias) sts128(vec_s + (blk * kEpiCols + c4) * 4, bz);
              if (has_dot) sts128(vec_s + kEpiVecBytes / 2 + (blk * kEpiCols + c4) * 4, wz);
            }
          }
          __syncwarp();
        }
        if (aux_mode != AUX_NONE) {
          aux_fetch(n0 + part * kEpiCols);
          // pull the aux tile of the NEXT tile this warp will process (same accumulator stage)
          // into L2 now: its loads are otherwise HBM-latency-bound (dX epilogue measured 12-14 k
          // cycles per tile against 5 k for the same epilogue without aux; profiles/r1e)
          const int nitem = item + NACC * item_step;
          if (nitem < total) {
            const int nrem = nitem - (nitem / tiles) * tiles;
            const int nm0 = ((nrem / p.n_tiles) * CS + crank) * BM + quarter * 32;
            const int nn0 = (nrem % p.n_tiles) * BN;
            // 32 rows x 416 B: each of the kParts warps of this quarter takes every kParts-th row
            for (int t = lane; t < (32 / kParts) * 4; t += 32) {
              const int r = nm0 + (t >> 2) * kParts + part, c = nn0 + (t & 3) * 64;
              if (r < p.M && c < p.out_cols)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.aux + size_t(r) * p.ld_aux + c));
            }
          }
        }
        const long long te0 = phase_clock();
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const long long te1 = phase_clock();
        long long t_ld = 0;
        float dot = 0.f;
        bool released = false;
        float l1_scale = 0.f;
        if (aux_mode == AUX_L1) l1_scale = __ldg(p.row_scale + (row < p.row_split ? 0 : 1));
        if (aux_mode == AUX_SIGMOID_GRAD || aux_mode == AUX_RELU_MASK || aux_mode == AUX_NONZERO_MASK) l1_scale = (p.row_vec != nullptr && row_ok) ? __ldg(p.row_vec + row) : 1.f;
#pragma unroll 1
        for (int bi = 0; part + bi * kParts < kBlocks; ++bi) {
          const int cb = (part + bi * kParts) * kEpiCols;
          const int col0 = n0 + cb;
          if (col0 >= p.out_cols) break;
          const int nch = (BN - cb) >= kEpiCols ? 2 : (BN - cb) / 16;
          const int cb_next = cb + kParts * kEpiCols;
          const bool last_block = (cb_next >= BN) || (n0 + cb_next >= p.out_cols);
          uint4 ax[4];
          if (aux_mode != AUX_NONE) {   // prefetched aux (coalesced mapping) -> smem -> own row
            if constexpr (kAuxAsync) {
              cp_async_wait_all();
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 4; ++q) ax[q] = lds128(auxt_s + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
            } else {
              stage_acquire();
#pragma unroll
              for (int it = 0; it < 4; ++it) sts128(stage_s + (it * 8 + lr) * kEpiPitch + lc * 16, pre[it]);
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 4; ++q) ax[q] = lds128(stage_s + lane * kEpiPitch + q * 16);
            }
            __syncwarp();
            if (!last_block) aux_fetch(n0 + cb_next);   // overlaps with this block's math + stores
          }
          uint4 axl[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
          if constexpr (SPLIT) {   // residual plane of the aux values of this lane's row (plain loads: the split kernels are not tuned)
            if (p.lo_off != 0 && aux_mode != AUX_NONE && aux_mode != AUX_RELU_MASK && aux_mode != AUX_NONZERO_MASK && row_ok) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (col0 + q * 8 < p.out_cols)
                  axl[q] = __ldg(reinterpret_cast<const uint4*>(p.aux + p.lo_off + size_t(row) * p.ld_aux + col0 + q * 8));
            }
          }
          // accumulator columns -> registers: both chunks in flight, one wait
          uint32_t raw[2][16];
          const long long tl0 = phase_clock();
#pragma unroll
          for (int q = 0; q < 2; ++q)
            if (q < nch && col0 + q * 16 < p.N) tmem_ld16_issue(t_row + cb + q * 16, raw[q]);
          tmem_ld_wait();
          t_ld += phase_clock() - tl0;
          if (last_block) {   // accumulator stage fully read by this warp: hand it back to the MMA warp
            released = true;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(tempty_bar(as), 0));
              else mbar_arrive(tempty_bar(as));
            }
          }
          // full 32-column block -> one bulk tensor store of the (64-byte rows, XOR-swizzled) tile
          const bool blk_tma = tma_st && nch == 2 && wrow0 < p.M;
          stage_acquire();
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q < nch) {
              const int c0 = col0 + q * 16;
              float v[16];
              if (c0 < p.N) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[q][j]);
                if (has_bias) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    const uint4 b = lds128(vec_s + (bi * kEpiCols + q * 16 + k4 * 4) * 4);
                    if (act == ACT_SIGMOID) {   // staged bias is 0.5 b: v = 0.5 (acc + b), exact
                      v[4 * k4 + 0] = fmaf(v[4 * k4 + 0], 0.5f, __uint_as_float(b.x)); v[4 * k4 + 1] = fmaf(v[4 * k4 + 1], 0.5f, __uint_as_float(b.y));
                      v[4 * k4 + 2] = fmaf(v[4 * k4 + 2], 0.5f, __uint_as_float(b.z)); v[4 * k4 + 3] = fmaf(v[4 * k4 + 3], 0.5f, __uint_as_float(b.w));
                    } else {
                      v[4 * k4 + 0] += __uint_as_float(b.x); v[4 * k4 + 1] += __uint_as_float(b.y);
                      v[4 * k4 + 2] += __uint_as_float(b.z); v[4 * k4 + 3] += __uint_as_float(b.w);
                    }
                  }
                }
                if (act == ACT_RELU) {
                  if (!store_pre) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                  }
                } else if (ACT_T < 0 && act == ACT_LRELU) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : p.act_slope * v[j];
                } else if (act == ACT_SIGMOID) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    if constexpr (SPLIT) v[j] = 1.f / (1.f + __expf(has_bias ? -2.f * v[j] : -v[j]));   // fp32-grade (tanh.approx is 2^-11)
                    else v[j] = has_bias ? fast_sigmoid_half(v[j]) : fast_sigmoid(v[j]);
                  }
                }
                if (aux_mode != AUX_NONE) {
                  const uint32_t w[8] = {ax[2 * q].x, ax[2 * q].y, ax[2 * q].z, ax[2 * q].w,
                                         ax[2 * q + 1].x, ax[2 * q + 1].y, ax[2 * q + 1].z, ax[2 * q + 1].w};
                  const uint32_t wl[8] = {axl[2 * q].x, axl[2 * q].y, axl[2 * q].z, axl[2 * q].w,
                                          axl[2 * q + 1].x, axl[2 * q + 1].y, axl[2 * q + 1].z, axl[2 * q + 1].w};
#pragma unroll
                  for (int k2 = 0; k2 < 8; ++k2) {
                    float a_lo = bf16_lo(w[k2]), a_hi = bf16_hi(w[k2]);
                    if constexpr (SPLIT) { a_lo += bf16_lo(wl[k2]); a_hi += bf16_hi(wl[k2]); }
                    if (aux_mode == AUX_SIGMOID_GRAD) {
                      v[2 * k2] *= (l1_scale * a_lo) * (1.f - a_lo);
                      v[2 * k2 + 1] *= (l1_scale * a_hi) * (1.f - a_hi);
                    } else if (aux_mode == AUX_L1) {
                      // BEGAN: L1 reconstruction error of the autoencoder-discriminator and its
                      // (scaled) subgradient sign(r - x)   (src/be_gan.py:225-236)
                      const float d0 = v[2 * k2] - a_lo, d1 = v[2 * k2 + 1] - a_hi;
                      dot += fabsf(d0) + fabsf(d1);
                      v[2 * k2] = d0 > 0.f ? l1_scale : (d0 < 0.f ? -l1_scale : 0.f);
                      v[2 * k2 + 1] = d1 > 0.f ? l1_scale : (d1 < 0.f ? -l1_scale : 0.f);
                    } else if (aux_mode == AUX_VAE_OUT) {
                      // v = decoder output, aux = target x: accumulate (x-v)^2 and emit
                      // d/d(pre-sigmoid) of sum (x-v)^2 = -2 (x-v) v (1-v)   (src/vae.py:203)
                      const float d0 = a_lo - v[2 * k2], d1 = a_hi - v[2 * k2 + 1];
                      dot = fmaf(d0, d0, dot); dot = fmaf(d1, d1, dot);
                      v[2 * k2] = -2.f * d0 * v[2 * k2] * (1.f - v[2 * k2]);
                      v[2 * k2 + 1] = -2.f * d1 * v[2 * k2 + 1] * (1.f - v[2 * k2 + 1]);
                    } else {
                      // aux > 0 (post-ReLU activation), or aux != 0 for the signed mask form w2 * relu'(a)
                      const bool on0 = aux_mode == AUX_NONZERO_MASK ? a_lo != 0.f : a_lo > 0.f;
                      const bool on1 = aux_mode == AUX_NONZERO_MASK ? a_hi != 0.f : a_hi > 0.f;
                      v[2 * k2] = on0 ? v[2 * k2] * l1_scale : 0.f;
                      v[2 * k2 + 1] = on1 ? v[2 * k2 + 1] * l1_scale : 0.f;
                    }
                  }
                }
                if (has_sq) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) dot = fmaf(v[j], v[j], dot);
                }
                if (has_dot) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    const uint4 w = lds128(vec_s + kEpiVecBytes / 2 + (bi * kEpiCols + q * 16 + k4 * 4) * 4);
                    if (store_pre) {   // v holds pre-activations: the activation enters the dot product only
                      dot = fmaf(fmaxf(v[4 * k4 + 0], 0.f), __uint_as_float(w.x), dot); dot = fmaf(fmaxf(v[4 * k4 + 1], 0.f), __uint_as_float(w.y), dot);
                      dot = fmaf(fmaxf(v[4 * k4 + 2], 0.f), __uint_as_float(w.z), dot); dot = fmaf(fmaxf(v[4 * k4 + 3], 0.f), __uint_as_float(w.w), dot);
                    } else {
                      dot = fmaf(v[4 * k4 + 0], __uint_as_float(w.x), dot); dot = fmaf(v[4 * k4 + 1], __uint_as_float(w.y), dot);
                      dot = fmaf(v[4 * k4 + 2], __uint_as_float(w.z), dot); dot = fmaf(v[4 * k4 + 3], __uint_as_float(w.w), dot);
                    }
                    if (mask_out) {
                      v[4 * k4 + 0] = v[4 * k4 + 0] > 0.f ? __uint_as_float(w.x) : 0.f;
                      v[4 * k4 + 1] = v[4 * k4 + 1] > 0.f ? __uint_as_float(w.y) : 0.f;
                      v[4 * k4 + 2] = v[4 * k4 + 2] > 0.f ? __uint_as_float(w.z) : 0.f;
                      v[4 * k4 + 3] = v[4 * k4 + 3] > 0.f ? __uint_as_float(w.w) : 0.f;
                    }
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
                if (c0 == p.N && p.pad_one) v[0] = 1.f;
              }
              {
                const uint4 lo = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                const uint4 hi = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                if constexpr (SPLIT) {   // residual plane of the output: v - bf16(v), straight from registers
                  if (p.lo_off != 0 && p.out != nullptr && row_ok && c0 < p.out_cols) {
                    __nv_bfloat16* ob = (mask_rows && row >= p.mask_row0) ? p.out_alt + size_t(row - p.mask_row0) * p.ldo : p.out + size_t(row) * p.ldo;
                    uint4* ol = reinterpret_cast<uint4*>(ob + p.lo_off + c0);
                    ol[0] = make_uint4(pack_bf16x2_residual(v[0], v[1], lo.x), pack_bf16x2_residual(v[2], v[3], lo.y),
                                       pack_bf16x2_residual(v[4], v[5], lo.z), pack_bf16x2_residual(v[6], v[7], lo.w));
                    ol[1] = make_uint4(pack_bf16x2_residual(v[8], v[9], hi.x), pack_bf16x2_residual(v[10], v[11], hi.y),
                                       pack_bf16x2_residual(v[12], v[13], hi.z), pack_bf16x2_residual(v[14], v[15], hi.w));
                  }
                }
                if (blk_tma) {
                  const uint32_t sw = (lane >> 1) & 3, rowb = stage_s + lane * 64;
                  sts128(rowb + (((2 * q) ^ sw) << 4), lo);
                  sts128(rowb + (((2 * q + 1) ^ sw) << 4), hi);
                } else {
                  sts128(stage_s + lane * kEpiPitch + q * 32, lo);
                  sts128(stage_s + lane * kEpiPitch + q * 32 + 16, hi);
                }
              }
            }
          }
          if (blk_tma) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, stage_s, col0, wrow0);
              bulk_commit();
            }
            st_pending = true;
            continue;
          }
          __syncwarp();
          // coalesced store: 4 lanes cover 64 contiguous bytes of one row, 8 rows per pass
          if (p.out != nullptr && lc < nch * 2 && col0 + lc * 8 < p.out_cols) {
            __nv_bfloat16* o = p.out + size_t(wrow0 + lr) * p.ldo + col0 + lc * 8;
            const uint32_t sa = stage_s + lr * kEpiPitch + lc * 16;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rr = wrow0 + it * 8 + lr;
              if (rr < p.M) {
                __nv_bfloat16* dst = o + size_t(it * 8) * p.ldo;
                if constexpr (kRowMask) {
                  if (mask_rows && rr >= p.mask_row0) dst = p.out_alt + size_t(rr - p.mask_row0) * p.ldo + col0 + lc * 8;
                }
                *reinterpret_cast<uint4*>(dst) = lds128(sa + it * 8 * kEpiPitch);
              }
            }
          }
          __syncwarp();
        }
        if ((has_dot || has_sq || aux_mode == AUX_VAE_OUT || aux_mode == AUX_L1) && p.dot_out != nullptr && row_ok)
          p.dot_out[size_t(n_tile * 2 + (kParts == 2 ? part : 0)) * p.dot_ld + row] = dot;
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(tempty_bar(as), 0));
            else mbar_arrive(tempty_bar(as));
          }
        }
        if (kPhaseTiming && p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && (e & 3) == 0 && part == 0 && acc_iter < 16) {
          long long* d = p.dbg + 64 + acc_iter * 4;   // [tile][wait accumulator ready, epilogue work, of which TMEM ld+wait, start stamp]
          d[0] = te1 - te0; d[1] = phase_clock() - te1; d[2] = t_ld; d[3] = te0;
        }
       }
      } else {
        // ====== fp32 epilogue: split-K partials (MN-major kernels) or biased fp32 output ======
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const int c_first = (NACC == 2) ? (grp >> 1) : grp;
        const int c_step = (NACC == 2) ? kEpiWarps / 8 : kEpiWarps / 4;
        float* base = p.part + size_t(split) * p.part_stride;
#pragma unroll 1
        for (int c = c_first; c < BN / 16; c += c_step) {
          const int col0 = n0 + c * 16;
          if (col0 >= p.N) break;
          float v[16];
          tmem_ld16(t_row + c * 16, v);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
          if (row_ok) {
            if (p.transpose) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < p.N) base[size_t(col0 + j) * p.ldp + row] = v[j];
            } else if (col0 + 16 <= p.N) {
              float4* o = reinterpret_cast<float4*>(base + size_t(row) * p.ldp + col0);
#pragma unroll
              for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < p.N) base[size_t(row) * p.ldp + col0 + j] = v[j];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(tempty_bar(as), 0));
          else mbar_arrive(tempty_bar(as));
        }
      }
    }
    stage_acquire();   // the last bulk store has read its staging tile before shared memory goes away
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();   // no CTA leaves while a peer may still signal it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gm
