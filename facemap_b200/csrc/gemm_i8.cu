// Integer tensor-core GEMM: C[M,N] (int32) = A[M,K] . B[N,K]^T with 8-bit operands (u8 or s8 each), K-contiguous,
// tcgen05.mma kind::i8 with the int32 accumulator in TMEM, operands staged by TMA (64-byte swizzle).
//
// EXPERIMENTAL — written at the end of round 1 without GPU time left: it compiles for sm_100a (ptxas emits UTCIMMA) but
// has not run yet, nothing in the product path calls it, and its GPU test is opt-in (tests/test_gemm_i8_gpu.py).
//
// Why it exists: the operands of this path are integers — the motion energy per binned
// pixel |dS| <= 255 sbin^2 and the block sums themselves (facemap/process.py:34-43, 201-212) — so Gram matrices and
// projections can be formed from 8-bit digit planes with EXACT int32 accumulation (no truncating fp32 accumulator, no
// float64 partial stacks), moving a quarter of the bytes of a TF32 product per K element.  One product per launch;
// the caller combines digit planes with their weights (256^d, 2^-8e).
//
// Structure (the one-product form of gemm_tf32x3_chain_kernel): one CTA per 128 x 256 tile and K range,
//   warp 0      TMA producer: A (128 x 64 B) and B (256 x 64 B) tiles per stage, 6 stages
//   warp 1      TMEM allocation (256 columns) and the MMA thread: two K = 32 instructions per stage
//   warps 2..9  epilogue: tcgen05.ld -> int32 -> global (or split-K partials)
// Overflow: |a b| <= 255^2 (u8 u8) or 255 * 128, so one accumulator holds K <= 33 000 (u8 u8) / 65 000 (u8 s8)
// terms; longer contractions are split over blockIdx.z and summed in int64 by the caller.
#include "i8_mma.cuh"

namespace fm {
namespace i8 {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int STAGES = 6;
constexpr int STAGE_BYTES = (BM + BN) * BKB;
constexpr int NUM_EPI_WARPS = 8;
constexpr int THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
constexpr int KMAX_PER_ACC = 32768;     // 255^2 * 32768 < 2^31

struct Params {
  int M, N, K;
  int32_t* C;
  int64_t ldc;
  int ksplit;
  int32_t* workspace;
  int64_t ldw;
  int upper_only;
  uint32_t idesc;
};


__global__ void __launch_bounds__(THREADS, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  if (p.upper_only && n0 + BN - 1 < m0) return;            // tile entirely below the diagonal

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = (p.K + BKB - 1) / BKB;
  const int kb0 = (int)(((long long)blockIdx.z * KB) / p.ksplit);
  const int kb1 = (int)(((long long)(blockIdx.z + 1) * KB) / p.ksplit);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(BN) : "memory");
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
        const uint32_t stage = smem_base + s * STAGE_BYTES;
        mbar_expect_tx(full_bar(s), (uint32_t)STAGE_BYTES);
        tma_load_2d(stage, &tmA, full_bar(s), (kb0 + i) * BKB, m0);
        tma_load_2d(stage + BM * BKB, &tmB, full_bar(s), (kb0 + i) * BKB, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tcgen05_fence_after();
        const uint32_t stage = smem_base + s * STAGE_BYTES;
        const uint64_t da = make_desc(stage);
        const uint64_t db = make_desc(stage + BM * BKB);
#pragma unroll
        for (int k = 0; k < BKB / 32; ++k) {               // K = 32 per instruction = 32 bytes = 2 descriptor units
          const uint64_t adv = (uint64_t)(k * 2);
          umma_i8(tmem_acc, da + adv, db + adv, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        tcgen05_commit(empty_bar(s));
      }
      tcgen05_commit(acc_bar);
    }
  } else {
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                      // which 128 of the 256 accumulator columns
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    int32_t* dst_row = (p.ksplit > 1)
                           ? p.workspace + ((size_t)blockIdx.z * p.M + (size_t)(row_ok ? m : 0)) * p.ldw
                           : p.C + (size_t)(row_ok ? m : 0) * p.ldc;
    const bool vec_ok = ((p.ksplit > 1 ? p.ldw : p.ldc) % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.ksplit > 1 ? p.workspace : p.C) & 15) == 0);
    if (nkb > 0) {
      mbar_wait(acc_bar, 0u);
      tcgen05_fence_after();
    }
#pragma unroll 1
    for (int cc = 0; cc < BN / 2; cc += 32) {
      uint32_t r[32];
      if (nkb > 0) {
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2) + cc);
        FM_I8_TMEM_LD_X32(taddr, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;            // an empty K range contributes zeros
      }
      if (row_ok) {
        const int nbase = n0 + half * (BN / 2) + cc;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = nbase + j;
          if (vec_ok && n + 4 <= p.N) {
            *reinterpret_cast<int4*>(dst_row + n) = make_int4((int)r[j], (int)r[j + 1], (int)r[j + 2], (int)r[j + 3]);
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (n + t < p.N) dst_row[n + t] = (int32_t)r[j + t];
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(BN) : "memory");
  }
}


}  // namespace i8
}  // namespace fm

extern "C" int fm_gemm_i8_tn(const uint8_t* A, int64_t lda, int a_signed, const uint8_t* B, int64_t ldb, int b_signed,
                             int M, int N, int K, int32_t* C, int64_t ldc, int ksplit, int32_t* workspace, int64_t ldw,
                             int upper_only, void* stream) {
  using namespace fm;
  using namespace fm::i8;
  FM_REQUIRE(A && B && M >= 1 && N >= 1 && K >= 1, "fm_gemm_i8_tn: bad shape M=%d N=%d K=%d", M, N, K);
  FM_REQUIRE(lda >= K && ldb >= K && lda % 16 == 0 && ldb % 16 == 0 &&
                 ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
             "fm_gemm_i8_tn: TMA operands need 16-byte aligned bases and row pitches (lda=%lld ldb=%lld)",
             (long long)lda, (long long)ldb);
  FM_REQUIRE(ksplit >= 1, "fm_gemm_i8_tn: ksplit must be >= 1");
  const int KB = (int)cdiv(K, BKB);
  if (ksplit > KB) ksplit = KB;
  FM_REQUIRE((int64_t)cdiv(KB, ksplit) * BKB <= KMAX_PER_ACC,
             "fm_gemm_i8_tn: %d K per partial would overflow the int32 accumulator; raise ksplit", (int)(cdiv(KB, ksplit) * BKB));
  if (ksplit > 1) {
    FM_REQUIRE(workspace && ldw >= N, "fm_gemm_i8_tn: split-K needs a workspace with ldw >= N");
  } else {
    FM_REQUIRE(C && ldc >= N, "fm_gemm_i8_tn: C missing or ldc < N");
  }
  FM_REQUIRE(ksplit <= 65535 && cdiv(M, BM) <= 65535, "fm_gemm_i8_tn: grid too large");
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp layout): c_format S32 = 2 at bit 4, a/b format (0 = u8,
  // 1 = s8) at bits 7 / 10, K-major operands, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  Params p{M, N, K, C, ldc, ksplit, workspace, ldw, upper_only, idesc};
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_map_u8(&ta, A, lda, M, K, BM)) != FM_OK) return rc;
  if ((rc = make_map_u8(&tb, B, ldb, N, K, BN)) != FM_OK) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  FM_CUDA_TRY(attr_err);
  dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(M, BM), (unsigned)ksplit);
  gemm_i8_kernel<<<grid, THREADS, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(ta, tb, p);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
