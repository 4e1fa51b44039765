// Integer tensor-core projection: V[t,c] = alpha * sum_p D[t,p] * U[c,p] - bias[c], the projection of every frame's
// motion energy (or binned frame) onto the masks (facemap/process.py:485, 495, 503, 510) with the frame operand as
// the INTEGERS it is: D = |S_t - S_{t-1}| (or S_t), a sum of sbin^2 bytes, held as one (D < 256) or two u8 planes
// D = 256 D_hi + D_lo, and every mask row as the s8 / u8 / u8 digits of fm_slice_rows_i8,
//   U[c,p] ~ 2^(exps[c] - 7) * (Q0 + Q1 / 2^8 + Q2 / 2^16).
// The ND x 3 plane products run on tcgen05.mma kind::i8 with exact int32 accumulation, one TMEM accumulator per weight
// class 256^(ND-1-j), j = d + e (products of equal weight share an accumulator): 3 or 4 accumulators of 128 columns.
// The subtraction of the average (avgmotion / avgframe) is the rank-one term bias[c] = sum_p avg[p] U[c,p], formed once
// in float64 (fm_row_dot).  Nothing rounds but the 2^-24 digit representation of the masks and the final float32 store.
//
// EXPERIMENTAL — written at the end of round 1 without GPU time left (like gemm_i8.cu, whose pipeline this shares):
// it compiles for sm_100a but has not run yet; nothing in the default path calls it; its GPU test is opt-in
// (tests/test_project_i8_gpu.py).  Why: the 3xTF32 projection issues 12 TF32 MMAs per 32 K; this one issues 6 (ND = 2)
// or 3 (ND = 1) kind::i8 MMAs at four times the per-instruction K rate and stages 5 (4) bytes per K element and tile
// row instead of 16 — DESIGN.md §7.2(a).
//
// One CTA per 128 frames x 128 components, the whole K range:
//   warp 0      TMA producer: ND frame planes (128 x 64 B) and 3 digit planes (128 x 64 B) per stage
//   warp 1      TMEM allocation (512 columns) and the MMA thread: 2 x ND x 3 instructions per stage; K is cut into
//               chains of <= 16 384 (2 * 255^2 * 16 384 < 2^31: two u8 x u8 products per accumulator stay exact)
//   warps 2..9  after each chain: tcgen05.ld of the class accumulators, Horner-combined into one int64 per element
//               kept in registers (64 columns per thread); at the end  V = alpha * 2^(ex - 23) * acc - bias  in float64
#include "i8_mma.cuh"

namespace fm {
namespace pi8 {

using namespace fm::i8;

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int NE = 3;                    // digits per mask entry
constexpr int NUM_EPI_WARPS = 8;
constexpr int THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int COLS_PER_THREAD = BN / 2;  // two epilogue warps share a TMEM lane quarter, half the columns each
constexpr int BAR_BYTES = 256;
constexpr int KCHAIN = 16384;
constexpr int TMEM_COLS = 512;

template <int ND>
struct Cfg {
  static constexpr int NC = ND + NE - 1;                       // weight classes = TMEM accumulators
  static constexpr int STAGE_BYTES = (ND * BM + NE * BN) * BKB;
  static constexpr int STAGES = ND == 1 ? 6 : 5;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static_assert(NC * BN <= TMEM_COLS, "accumulators exceed TMEM");
  static_assert(SMEM_BYTES <= 232448, "stage ring exceeds shared memory");
};

struct Params {
  int M, N, K;
  float* V;
  int64_t ldv;
  const int32_t* exps;
  const double* bias;
  double alpha;
  uint32_t idesc_s8, idesc_u8;           // B operand signed (digit 0) / unsigned (digits 1, 2)
  int kb_shift;                          // log2 of the K-block width of the operand layout (31: plain row-major matrices)
};

#define FM_I8_TMEM_LD_X8(taddr, r)                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"           \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(taddr)                                                                              \
               : "memory")

// one chain's class accumulators -> the thread's int64 sums (Horner weights 256^(NC-1-j))
template <int NC>
__device__ __forceinline__ void drain_classes(uint32_t tmem_acc, int q, int half, long long (&acc)[COLS_PER_THREAD]) {
#pragma unroll
  for (int cc = 0; cc < COLS_PER_THREAD; cc += 8) {
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * COLS_PER_THREAD + cc);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      uint32_t r[8];
      FM_I8_TMEM_LD_X8(taddr + (uint32_t)(j * BN), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[cc + t] += (long long)(int32_t)r[t] * (1ll << (8 * (NC - 1 - j)));
    }
  }
}

// V = alpha * 2^(ex - 7) * 256^-(NE-1) * acc - bias for the thread's row m and its 64 columns from nbase on
__device__ __forceinline__ void store_row(const long long (&acc)[COLS_PER_THREAD], int m, int nbase, const Params& p) {
  if (m >= p.M) return;
  float* dst_row = p.V + (size_t)m * p.ldv;
  const bool vec_ok = (p.ldv % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.V) & 15) == 0);
#pragma unroll
  for (int t = 0; t < COLS_PER_THREAD; t += 4) {
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = nbase + t + u;
      double v = 0.0;
      if (n < p.N) {
        // 2^(ex - 23) built from its exponent field (|ex| <= 126 by fm::slice_exponent: always normal); separately
        // rounded multiply and subtract (no FMA contraction) so that a host model reproduces every bit
        const double sc = __dmul_rn(p.alpha, __hiloint2double((1023 + p.exps[n] - 7 - 8 * (NE - 1)) << 20, 0));
        v = __dmul_rn((double)acc[t + u], sc);
        if (p.bias) v = __dsub_rn(v, p.bias[n]);
      }
      o[u] = (float)v;
    }
    const int n = nbase + t;
    if (vec_ok && n + 4 <= p.N) {
      *reinterpret_cast<float4*>(dst_row + n) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (n + u < p.N) dst_row[n + u] = o[u];
    }
  }
}

template <int ND>
__global__ void __launch_bounds__(THREADS, 1)
project_i8_kernel(const __grid_constant__ CUtensorMap tmD0, const __grid_constant__ CUtensorMap tmD1,
                  const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmQ1,
                  const __grid_constant__ CUtensorMap tmQ2, const Params p) {
  using C = Cfg<ND>;
  constexpr int STAGES = C::STAGES;
  constexpr int NC = C::NC;
  const int n0 = blockIdx.x * BN;        // the component tiles of one frame tile are neighbours: D is shared through L2
  const int m0 = blockIdx.y * BM;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t accempty_bar = bar_base + 8u * (2 * STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * C::STAGE_BYTES + 8 * (2 * STAGES + 2));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = (p.K + BKB - 1) / BKB;
  constexpr int CHAIN_KB = KCHAIN / BKB;
  const int nchains = (nkb + CHAIN_KB - 1) / CHAIN_KB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accfull_bar, 1);
    mbar_init(accempty_bar, NUM_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t stage = smem_base + s * C::STAGE_BYTES;
        const uint32_t qbase = stage + ND * BM * BKB;
        mbar_expect_tx(full_bar(s), (uint32_t)C::STAGE_BYTES);
        const int kblk = (i * BKB) >> p.kb_shift, kin = (i * BKB) - (kblk << p.kb_shift);
        tma_load_3d(stage, &tmD0, full_bar(s), kin, m0, kblk);
        if constexpr (ND == 2) tma_load_3d(stage + BM * BKB, &tmD1, full_bar(s), kin, m0, kblk);
        tma_load_3d(qbase, &tmQ0, full_bar(s), kin, n0, kblk);
        tma_load_3d(qbase + BN * BKB, &tmQ1, full_bar(s), kin, n0, kblk);
        tma_load_3d(qbase + 2 * BN * BKB, &tmQ2, full_bar(s), kin, n0, kblk);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int i = 0;
      for (int c = 0; c < nchains; ++c) {
        if (c > 0) {                                       // the epilogue has drained the previous chain
          mbar_wait(accempty_bar, (uint32_t)(c - 1) & 1u);
          tcgen05_fence_after();
        }
        const int len = min(CHAIN_KB, nkb - c * CHAIN_KB);
        for (int j = 0; j < len; ++j, ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t stage = smem_base + s * C::STAGE_BYTES;
          const uint32_t qbase = stage + ND * BM * BKB;
#pragma unroll
          for (int k = 0; k < BKB / 32; ++k) {             // K = 32 per instruction = 32 bytes = 2 descriptor units
            const uint64_t adv = (uint64_t)(k * 2);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
              const uint64_t dd = make_desc(stage + d * BM * BKB) + adv;
#pragma unroll
              for (int e = 0; e < NE; ++e) {
                const uint64_t dq = make_desc(qbase + e * BN * BKB) + adv;
                // issue order (d outer, e inner): class d + e is first written by (0, e) or by (d, NE - 1)
                const bool opens = (d == 0 || e == NE - 1);
                umma_i8(tmem_acc + (uint32_t)((d + e) * BN), dd, dq, e == 0 ? p.idesc_s8 : p.idesc_u8,
                        (opens && j == 0 && k == 0) ? 0u : 1u);
              }
            }
          }
          tcgen05_commit(empty_bar(s));
        }
        tcgen05_commit(accfull_bar);
      }
    }
  } else {
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                      // which 64 of the 128 columns of every class
    long long acc[COLS_PER_THREAD];
#pragma unroll
    for (int t = 0; t < COLS_PER_THREAD; ++t) acc[t] = 0;
    for (int c = 0; c < nchains; ++c) {
      mbar_wait(accfull_bar, (uint32_t)c & 1u);
      tcgen05_fence_after();
      drain_classes<NC>(tmem_acc, q, half, acc);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty_bar);
    }
    store_row(acc, m0 + q * 32 + lane, n0 + half * COLS_PER_THREAD, p);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant (impl 1, cta_group::2): two CTAs of a cluster project 256 frames x 128 components.  Each CTA stages
// its own 128 frame rows and HALF of every digit plane (64 component rows), so per SM the TMA fill drops from 40 to
// 28 KB per stage and the tensor core reads 6 instead of 8 KB of shared memory per instruction — the two limits a
// 128 x 128 kind::i8 tile sits on.  The protocol is gemm_tf32x3_pair_kernel's, which has run
// on B200: the leader's MMA thread issues for both tensor cores; a relay warp in each CTA forwards "my stage has
// landed" to the LEADER's barrier (where that kernel's splitter warps arrived); tcgen05.commit multicasts "stage
// free" and "chain complete" to both CTAs; between chains both CTAs' epilogue warps arrive on the leader's barrier.
// ---------------------------------------------------------------------------------------------
constexpr int PAIR_THREADS = 32 * (3 + NUM_EPI_WARPS);     // TMA, MMA, relay, 8 epilogue warps

template <int ND>
struct PairCfg {
  static constexpr int NC = ND + NE - 1;
  static constexpr int STAGE_BYTES = (ND * BM + NE * (BN / 2)) * BKB;
  static constexpr int STAGES = ND == 1 ? 8 : 7;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "stage ring exceeds shared memory");
  static_assert(8 * (3 * STAGES + 3) <= BAR_BYTES, "barrier block too small");
};

template <int ND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
project_i8_pair_kernel(const __grid_constant__ CUtensorMap tmD0, const __grid_constant__ CUtensorMap tmD1,
                       const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmQ1,
                       const __grid_constant__ CUtensorMap tmQ2, const Params p) {
  using C = PairCfg<ND>;
  constexpr int STAGES = C::STAGES;
  constexpr int NC = C::NC;
  constexpr int QROWS = BN / 2;                            // digit-plane rows staged by each CTA
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs)
  const int m0 = (int)(blockIdx.x >> 1) * (2 * BM) + (int)rank * BM;   // the pair = 2 consecutive CTAs along x
  const int n0 = blockIdx.y * BN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                  // local: TMA landed
  auto ready_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };      // leader's: both CTAs' stages landed
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };  // local: MMAs done with the stage
  const uint32_t accfull_bar = bar_base + 8u * (3 * STAGES);                 // local: chain complete (multicast)
  const uint32_t accempty_bar = bar_base + 8u * (3 * STAGES + 1);            // leader's: both CTAs drained
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 2);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * C::STAGE_BYTES + 8 * (3 * STAGES + 2));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = (p.K + BKB - 1) / BKB;
  constexpr int CHAIN_KB = KCHAIN / BKB;
  const int nchains = (nkb + CHAIN_KB - 1) / CHAIN_KB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(ready_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accfull_bar, 1);
    mbar_init(accempty_bar, 2 * NUM_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // peer barriers initialised, TMEM allocated
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t stage = smem_base + s * C::STAGE_BYTES;
        const uint32_t qbase = stage + ND * BM * BKB;
        const int nq = n0 + (int)rank * QROWS;
        mbar_expect_tx(full_bar(s), (uint32_t)C::STAGE_BYTES);
        const int kblk = (i * BKB) >> p.kb_shift, kin = (i * BKB) - (kblk << p.kb_shift);
        tma_load_3d(stage, &tmD0, full_bar(s), kin, m0, kblk);
        if constexpr (ND == 2) tma_load_3d(stage + BM * BKB, &tmD1, full_bar(s), kin, m0, kblk);
        tma_load_3d(qbase, &tmQ0, full_bar(s), kin, nq, kblk);
        tma_load_3d(qbase + QROWS * BKB, &tmQ1, full_bar(s), kin, nq, kblk);
        tma_load_3d(qbase + 2 * QROWS * BKB, &tmQ2, full_bar(s), kin, nq, kblk);
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      int i = 0;
      for (int c = 0; c < nchains; ++c) {
        if (c > 0) {                                       // both CTAs have drained the previous chain
          mbar_wait_cluster(accempty_bar, (uint32_t)(c - 1) & 1u);
          tcgen05_fence_after();
        }
        const int len = min(CHAIN_KB, nkb - c * CHAIN_KB);
        for (int j = 0; j < len; ++j, ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
          mbar_wait_cluster(ready_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t stage = smem_base + s * C::STAGE_BYTES;
          const uint32_t qbase = stage + ND * BM * BKB;
#pragma unroll
          for (int k = 0; k < BKB / 32; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
              const uint64_t dd = make_desc(stage + d * BM * BKB) + adv;
#pragma unroll
              for (int e = 0; e < NE; ++e) {
                const uint64_t dq = make_desc(qbase + e * QROWS * BKB) + adv;
                const bool opens = (d == 0 || e == NE - 1);
                umma_i8_pair(tmem_acc + (uint32_t)((d + e) * BN), dd, dq, e == 0 ? p.idesc_s8 : p.idesc_u8,
                             (opens && j == 0 && k == 0) ? 0u : 1u);
              }
            }
          }
          tcgen05_commit_pair(empty_bar(s));
        }
        tcgen05_commit_pair(accfull_bar);
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {                                       // relay: my stage has landed -> the leader's barrier
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        mbar_arrive_cluster(ready_bar(s), 0);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 3) >> 2;
    long long acc[COLS_PER_THREAD];
#pragma unroll
    for (int t = 0; t < COLS_PER_THREAD; ++t) acc[t] = 0;
    for (int c = 0; c < nchains; ++c) {
      mbar_wait(accfull_bar, (uint32_t)c & 1u);
      tcgen05_fence_after();
      drain_classes<NC>(tmem_acc, q, half, acc);           // every CTA drains its own 128 TMEM lanes
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(accempty_bar, 0);
    }
    store_row(acc, m0 + q * 32 + lane, n0 + half * COLS_PER_THREAD, p);
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer's tensor core may still be reading this CTA's shared memory / TMEM
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}

// kb == 0: plain row-major operands of pitch ldd / ldq; kb > 0 (a power of two): K-blocked operands, ldd / ldq are
// then the block strides
template <int ND>
static int launch(const uint8_t* D_hi, const uint8_t* D_lo, int64_t ldd, const int8_t* Q0, const uint8_t* Q1,
                  const uint8_t* Q2, int64_t ldq, Params p, int impl, cudaStream_t st, int64_t kb = 0) {
  const bool pair = impl == 1;
  CUtensorMap td0, td1, tq0, tq1, tq2;
  int rc;
  // plane 0 is the most significant one; a pair CTA stages half of every digit tile
  const int qrows = pair ? BN / 2 : BN;
  p.kb_shift = 31;
  if (kb > 0)
    for (p.kb_shift = 0; ((int64_t)1 << p.kb_shift) < kb; ++p.kb_shift) {}
  if ((rc = make_map_u8_3d(&td0, ND == 2 ? D_hi : D_lo, ldd, kb, ldd, p.M, p.K, BM)) != FM_OK) return rc;
  td1 = td0;
  if (ND == 2 && (rc = make_map_u8_3d(&td1, D_lo, ldd, kb, ldd, p.M, p.K, BM)) != FM_OK) return rc;
  if ((rc = make_map_u8_3d(&tq0, reinterpret_cast<const uint8_t*>(Q0), ldq, kb, ldq, p.N, p.K, qrows)) != FM_OK) return rc;
  if ((rc = make_map_u8_3d(&tq1, Q1, ldq, kb, ldq, p.N, p.K, qrows)) != FM_OK) return rc;
  if ((rc = make_map_u8_3d(&tq2, Q2, ldq, kb, ldq, p.N, p.K, qrows)) != FM_OK) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess, attr_err_pair = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(project_i8_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg<ND>::SMEM_BYTES);
    attr_err_pair = cudaFuncSetAttribute(project_i8_pair_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         PairCfg<ND>::SMEM_BYTES);
  });
  if (pair) {
    FM_CUDA_TRY(attr_err_pair);
    // instruction M spans both CTAs
    const uint32_t m256 = ((uint32_t)((2 * BM) >> 4) << 24), mmask = ~(0x1Fu << 24);
    p.idesc_s8 = (p.idesc_s8 & mmask) | m256;
    p.idesc_u8 = (p.idesc_u8 & mmask) | m256;
    dim3 grid((unsigned)(2 * cdiv(p.M, 2 * BM)), (unsigned)cdiv(p.N, BN), 1);
    project_i8_pair_kernel<ND><<<grid, PAIR_THREADS, PairCfg<ND>::SMEM_BYTES, st>>>(td0, td1, tq0, tq1, tq2, p);
  } else {
    FM_CUDA_TRY(attr_err);
    dim3 grid((unsigned)cdiv(p.N, BN), (unsigned)cdiv(p.M, BM), 1);
    project_i8_kernel<ND><<<grid, THREADS, Cfg<ND>::SMEM_BYTES, st>>>(td0, td1, tq0, tq1, tq2, p);
  }
  FM_LAUNCH_CHECK();
  return FM_OK;
}

}  // namespace pi8
}  // namespace fm

extern "C" int fm_project_i8(const uint8_t* D_hi, const uint8_t* D_lo, int64_t ldd, const int8_t* Q0, const uint8_t* Q1,
                             const uint8_t* Q2, int64_t ldq, const int32_t* exps, int M, int N, int K, double alpha,
                             const double* bias, float* V, int64_t ldv, int impl, void* stream) {
  using namespace fm;
  using namespace fm::pi8;
  FM_REQUIRE(D_lo && Q0 && Q1 && Q2 && exps && V && M >= 1 && N >= 1 && K >= 1 && ldv >= N,
             "fm_project_i8: bad args M=%d N=%d K=%d", M, N, K);
  FM_REQUIRE(ldd >= K && ldq >= K && ldd % 16 == 0 && ldq % 16 == 0 &&
                 ((reinterpret_cast<uintptr_t>(D_hi) | reinterpret_cast<uintptr_t>(D_lo) |
                   reinterpret_cast<uintptr_t>(Q0) | reinterpret_cast<uintptr_t>(Q1) | reinterpret_cast<uintptr_t>(Q2)) & 15) == 0,
             "fm_project_i8: TMA operands need 16-byte aligned bases and row pitches (ldd=%lld ldq=%lld)",
             (long long)ldd, (long long)ldq);
  FM_REQUIRE(cdiv(M, BM) <= 65535, "fm_project_i8: more than 65535 frame tiles; project in batches");
  FM_REQUIRE(impl == 0 || impl == 1, "fm_project_i8: unknown impl %d (0 = one CTA per tile, 1 = CTA pairs)", impl);
  FM_REQUIRE(K <= (1 << 22), "fm_project_i8: K = %d exceeds 2^22 (int64 accumulation over 256 chains)", K);
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp layout): c_format S32 = 2 at bit 4, a/b format (0 = u8,
  // 1 = s8) at bits 7 / 10, K-major operands, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc_u8 = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  Params p{M, N, K, V, ldv, exps, bias, alpha, idesc_u8 | (1u << 10), idesc_u8, 31};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return D_hi ? launch<2>(D_hi, D_lo, ldd, Q0, Q1, Q2, ldq, p, impl, st)
              : launch<1>(nullptr, D_lo, ldd, Q0, Q1, Q2, ldq, p, impl, st);
}

// The same projection on K-BLOCKED operands: element (row, k) of a plane lives at base + (k / kb) * block_stride +
// row * kb + k % kb (fm_retile_u8 makes that layout).  With the plain layout a 128-row operand tile of a wide view
// (row pitch 1.3 MB at 1280x1024, sbin 1) touches one 2 MB page per row and the projection runs at a third of its rate
// (profiles/r02_summary.md: 803 -> 278 TFLOP/s from the pitch alone); blocked, a tile is 128 * kb contiguous bytes.
// kb: a power of two >= 64; block strides multiples of 16, >= rows * kb; bytes between K and the end of the last block
// must be zero in the digit planes.
extern "C" int fm_project_i8_blocked(const uint8_t* D_hi, const uint8_t* D_lo, int64_t d_block_stride, const int8_t* Q0,
                                     const uint8_t* Q1, const uint8_t* Q2, int64_t q_block_stride, int64_t kb,
                                     const int32_t* exps, int M, int N, int K, double alpha, const double* bias, float* V,
                                     int64_t ldv, int impl, void* stream) {
  using namespace fm;
  using namespace fm::pi8;
  FM_REQUIRE(D_lo && Q0 && Q1 && Q2 && exps && V && M >= 1 && N >= 1 && K >= 1 && ldv >= N,
             "fm_project_i8_blocked: bad args M=%d N=%d K=%d", M, N, K);
  FM_REQUIRE(kb >= BKB && (kb & (kb - 1)) == 0 && d_block_stride >= (int64_t)M * kb && q_block_stride >= (int64_t)N * kb &&
                 d_block_stride % 16 == 0 && q_block_stride % 16 == 0 &&
                 ((reinterpret_cast<uintptr_t>(D_hi) | reinterpret_cast<uintptr_t>(D_lo) | reinterpret_cast<uintptr_t>(Q0) |
                   reinterpret_cast<uintptr_t>(Q1) | reinterpret_cast<uintptr_t>(Q2)) & 15) == 0,
             "fm_project_i8_blocked: kb must be a power of two >= %d, block strides >= rows * kb and multiples of 16, bases "
             "16-byte aligned (kb=%lld strides %lld %lld)", BKB, (long long)kb, (long long)d_block_stride, (long long)q_block_stride);
  FM_REQUIRE(cdiv(M, BM) <= 65535, "fm_project_i8_blocked: more than 65535 frame tiles; project in batches");
  FM_REQUIRE(impl == 0 || impl == 1, "fm_project_i8_blocked: unknown impl %d", impl);
  FM_REQUIRE(K <= (1 << 22), "fm_project_i8_blocked: K = %d exceeds 2^22", K);
  const uint32_t idesc_u8 = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  Params p{M, N, K, V, ldv, exps, bias, alpha, idesc_u8 | (1u << 10), idesc_u8, 31};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return D_hi ? launch<2>(D_hi, D_lo, d_block_stride, Q0, Q1, Q2, q_block_stride, p, impl, st, kb)
              : launch<1>(nullptr, D_lo, d_block_stride, Q0, Q1, Q2, q_block_stride, p, impl, st, kb);
}
