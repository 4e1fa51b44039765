// K2/K3/K4 — C = A . B^T with both operands K-contiguous float32, evaluated as a 3xTF32 split on
// the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged
// by TMA with the 128-byte swizzle).
//
// Replaces the dense contractions of the reference path: `X @ U` (facemap/process.py:485, 495,
// 503, 510) and the products inside utils.svdecon (facemap/utils.py:751-776), here in the Gram
// formulation (matlab/utils/svdecon.m:16-43).
//
// Kernels in this file (fm_gemm_tn's `impl`):
//   gemm_tf32x3_chain_kernel   impl 0, the product configuration: 128 x 256 tiles, two TMEM accumulators used
//                              alternately in chains of 128 K, drained into float32 register accumulators;
//                              impl 10 = the same kernel issuing ONE product per K step (exact-in-TF32 operands)
//   gemm_tf32x3_kernel         one TMEM chain per launch (impl 9; also serves outputs <= 128 wide), tile variants 2..6
//   gemm_tf32x3_pair_kernel    CTA pair, cta_group::2 (impl 7/8)
//   gemm_fp32_ref_kernel       plain FFMA reference for the tests (impl 1)
// Common structure — one CTA computes a 128 x BN tile over one K range (split-K over blockIdx.z):
//   warp 0        TMA producer: raw fp32 tiles of A (128x32) and B (BNx32) per stage
//   warps 2..9    splitter: lo = x - (x & 0xffffe000) into a second buffer (kind::tf32 reads only the
//                 top 19 bits of an operand word, so the raw tile itself is the hi operand), then
//                 fence.proxy.async + arrive; afterwards the same warps run the epilogue
//   warp 1        TMEM allocation and the single-thread MMA issue: per 8-wide K step
//                 D += A_lo.B_hi ; D += A_hi.B_lo ; D += A_hi.B_hi   (fp32 accumulate in TMEM)
//   epilogue      tcgen05.ld -> rscale[m]*acc + rbias[m] -> global (or raw split-K partials)
// The lo.lo term (~2^-22 relative) is dropped, as in every 3xTF32 scheme.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include "common.cuh"

namespace fm {

constexpr int BM = 128;
constexpr int REF_BK = 32;             // K granularity of the split-K partition (all variants)
constexpr int NUM_SPLIT_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (2 + NUM_SPLIT_WARPS);

struct GemmParams {
  int M, N, K;
  float* C;
  int64_t ldc;
  const float* rscale;
  const float* rbias;
  int ksplit;
  float* workspace;
  int64_t ldw;
  int upper_only;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    const uint64_t now = globaltimer_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) {
      printf("facemap_b200 gemm: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major swizzled shared-memory matrix descriptor (sm_100 format); rows are BK floats wide
// (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (= 1, unused for swizzled K-major)
//   [32,46) SBO >> 4 (8 rows x row bytes) | [46,48) version = 1
//   [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
template <int BK>
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t saddr) {
  constexpr uint64_t SBO = (8 * BK * 4) >> 4;
  constexpr uint64_t LAYOUT = (BK == 32) ? 2 : 4;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (SBO << 32) | (1ull << 46) | (LAYOUT << 61);
}

#define TMEM_LD_32x32b_X32(taddr, r)                                                                     \
  asm volatile(                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                          \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),         \
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),       \
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),       \
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                            \
      : "r"(taddr)                                                                                       \
      : "memory")

template <int BN, int BK>
struct GemmCfg {
  static_assert(BK == 32 || BK == 16, "BK must be one 128-byte or 64-byte swizzle row");
  static constexpr int HI_BYTES = (BM + BN) * BK * 4;     // raw/hi tiles of A then B
  static constexpr int STAGE_BYTES = 2 * HI_BYTES;        // + lo tiles
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

// HI_INPLACE: overwrite the raw tile with its TF32-truncated value.  kind::tf32 reads only the top
// 19 bits of each operand word, so the raw tile already IS the hi operand; the rewrite is kept as
// a template option (impl 5 in fm_gemm_tn) to let the tests prove the two are bit-identical.
// B_PRESPLIT: the lo part of B comes pre-computed from global memory (fm_split_lo; the projection's
// B operand is the mask matrix, constant over all frames and L2-resident), so TMA fills both B tiles and
// the splitter warps only touch the A tile — a third of the LDS/STS traffic on the shared-memory pipe
// that ncu shows to be the limiter of the in-kernel split.
// A_PRESPLIT (needs B_PRESPLIT): A's lo plane comes from global memory too (the Gram matrices, where
// A = B = X and one fm_split_lo pass serves both operands): no splitter stage at all, the MMA thread
// consumes the TMA barrier directly — a plain TMA -> tcgen05 pipeline.
template <int BN, int BK, bool HI_INPLACE, bool B_PRESPLIT, bool A_PRESPLIT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
                   const GemmParams p) {
  static_assert(!A_PRESPLIT || B_PRESPLIT, "A_PRESPLIT requires B_PRESPLIT");
  using Cfg = GemmCfg<BN, BK>;
  constexpr int STAGES = Cfg::STAGES;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  if (p.upper_only && n0 + BN - 1 < m0) return;  // tile entirely below the diagonal

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barriers: full[s], split[s], empty[s], tmem_full ; then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 1);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int KB = (p.K + BK - 1) / BK;
  const int kb0 = (int)(((long long)blockIdx.z * KB) / p.ksplit);
  const int kb1 = (int)(((long long)(blockIdx.z + 1) * KB) / p.ksplit);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), NUM_SPLIT_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(full_bar(s), (uint32_t)(Cfg::HI_BYTES + (B_PRESPLIT ? BN * BK * 4 : 0) +
                                               (A_PRESPLIT ? BM * BK * 4 : 0)));
        tma_load_2d(stage, &tmA, full_bar(s), (kb0 + i) * BK, m0);
        tma_load_2d(stage + BM * BK * 4, &tmB, full_bar(s), (kb0 + i) * BK, n0);
        if constexpr (A_PRESPLIT) tma_load_2d(stage + Cfg::HI_BYTES, &tmAlo, full_bar(s), (kb0 + i) * BK, m0);
        if constexpr (B_PRESPLIT)
          tma_load_2d(stage + Cfg::HI_BYTES + BM * BK * 4, &tmBlo, full_bar(s), (kb0 + i) * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor: D=F32 (bit 4), A=B=TF32 (2 at bits 7, 10), K-major both, N>>3 @17, M>>4 @24
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(A_PRESPLIT ? full_bar(s) : split_bar(s), ph);
        tcgen05_fence_after();
        const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
        const uint64_t a_hi = make_desc_kmajor<BK>(stage);
        const uint64_t b_hi = make_desc_kmajor<BK>(stage + BM * BK * 4);
        const uint64_t a_lo = make_desc_kmajor<BK>(stage + Cfg::HI_BYTES);
        const uint64_t b_lo = make_desc_kmajor<BK>(stage + Cfg::HI_BYTES + BM * BK * 4);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);  // 8 floats = 32 bytes = 2 x 16-byte units
          umma_tf32(tmem_acc, a_lo + adv, b_hi + adv, IDESC, (i > 0 || k > 0) ? 1u : 0u);
          umma_tf32(tmem_acc, a_hi + adv, b_lo + adv, IDESC, 1u);
          umma_tf32(tmem_acc, a_hi + adv, b_hi + adv, IDESC, 1u);
        }
        tcgen05_commit(empty_bar(s));  // frees the stage once these MMAs have read it
      }
      tcgen05_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // ===== splitter, then epilogue =====
    const int st = threadIdx.x - 64;  // 0 .. 255
    for (int i = 0; i < (A_PRESPLIT ? 0 : nkb); ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      float4* hi = reinterpret_cast<float4*>(smem_gen + s * Cfg::STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(smem_gen + s * Cfg::STAGE_BYTES + Cfg::HI_BYTES);
      constexpr int NV = (B_PRESPLIT ? BM * BK * 4 : Cfg::HI_BYTES) / 16;   // A tile only when B is pre-split
#pragma unroll 4
      for (int v = st; v < NV; v += 32 * NUM_SPLIT_WARPS) {
        const float4 x = hi[v];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
        if constexpr (HI_INPLACE) hi[v] = h;
        lo[v] = l;
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(split_bar(s));
    }

    // epilogue: warp w may only touch TMEM lanes [32*(w%4), +32)
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // 0 / 1: which half of the BN columns
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    float scale = 1.f, bias = 0.f;
    float* dst_row;
    if (p.ksplit > 1) {
      dst_row = p.workspace + ((size_t)blockIdx.z * p.M + (size_t)(row_ok ? m : 0)) * p.ldw;
    } else {
      dst_row = p.C + (size_t)(row_ok ? m : 0) * p.ldc;
      if (row_ok && p.rscale) scale = p.rscale[m];
      if (row_ok && p.rbias) bias = p.rbias[m];
    }
    const bool vec_ok = ((p.ksplit > 1 ? p.ldw : p.ldc) % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.ksplit > 1 ? p.workspace : p.C) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 2; c += 32) {
      const int col = half * (BN / 2) + c;
      uint32_t r[32];
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
      TMEM_LD_32x32b_X32(taddr, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int n = n0 + col;
      if (row_ok && n < p.N) {
        if (vec_ok && n + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = fmaf(__uint_as_float(r[j + 0]), scale, bias);
            o.y = fmaf(__uint_as_float(r[j + 1]), scale, bias);
            o.z = fmaf(__uint_as_float(r[j + 2]), scale, bias);
            o.w = fmaf(__uint_as_float(r[j + 3]), scale, bias);
            *reinterpret_cast<float4*>(dst_row + n + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < p.N) dst_row[n + j] = fmaf(__uint_as_float(r[j]), scale, bias);
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

// ---------------------------------------------------------------------------------------------
// Chained-accumulation variant: one CTA walks its whole K range in CHAINS of CH k-blocks.  The MMA
// thread alternates between two 256-column TMEM accumulators (all 512 columns); while chain c+1 is
// being accumulated, the splitter/epilogue warps drain chain c with tcgen05.ld and add it into
// float32 REGISTER accumulators (128 per thread) with round-to-nearest.  The tensor cores truncate
// the fp32 accumulator on every MMA (-5.4e-8 relative, always toward zero); bounding every TMEM chain
// to CH*BK/8*3 = 48 accumulations caps that bias at ~2.6e-6 of an entry without any split-K workspace,
// partial-sum traffic or reduction kernel.
//   SPLIT_MODE 0: A and B split in the kernel; 1: B pre-split (TMA stages B_lo), A split in the kernel;
//              2: both pre-split (no split work; the warps only drain);
//              3: single product — the operands are taken as they are (top 19 bits of every word), ONE MMA per
//                 K step, no lo planes staged.  For operands that are exact in TF32 (the 8-bit slices of
//                 fm_slice_rows, integer planes): with K partials short enough every sum is exact.
// ---------------------------------------------------------------------------------------------
template <int BK, int SPLIT_MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32x3_chain_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
                         const GemmParams p) {
  constexpr int BN = 256;
  constexpr int CH = 8;                       // k-blocks per TMEM chain
  using Cfg = GemmCfg<BN, BK>;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(CH > STAGES, "a chain must outlast the stage ring (drain scheduling)");
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  if (p.upper_only && n0 + BN - 1 < m0) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto accfull_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + b); };
  auto accempty_bar = [&](int b) { return bar_base + 8u * (3 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = (p.K + BK - 1) / BK;
  const int kb0 = (int)(((long long)blockIdx.z * KB) / p.ksplit);
  const int kb1 = (int)(((long long)(blockIdx.z + 1) * KB) / p.ksplit);
  const int nkb = kb1 - kb0;
  const int nchains = (nkb + CH - 1) / CH;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), NUM_SPLIT_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accfull_bar(b), 1);
      mbar_init(accempty_bar(b), NUM_SPLIT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BN)
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
        const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(full_bar(s), (uint32_t)(Cfg::HI_BYTES + ((SPLIT_MODE == 1 || SPLIT_MODE == 2) ? BN * BK * 4 : 0) +
                                               (SPLIT_MODE == 2 ? BM * BK * 4 : 0)));
        tma_load_2d(stage, &tmA, full_bar(s), (kb0 + i) * BK, m0);
        tma_load_2d(stage + BM * BK * 4, &tmB, full_bar(s), (kb0 + i) * BK, n0);
        if constexpr (SPLIT_MODE == 2) tma_load_2d(stage + Cfg::HI_BYTES, &tmAlo, full_bar(s), (kb0 + i) * BK, m0);
        if constexpr (SPLIT_MODE == 1 || SPLIT_MODE == 2)
          tma_load_2d(stage + Cfg::HI_BYTES + BM * BK * 4, &tmBlo, full_bar(s), (kb0 + i) * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      int i = 0;
      for (int c = 0; c < nchains; ++c) {
        const int buf = c & 1;
        mbar_wait(accempty_bar(buf), (((uint32_t)(c >> 1)) & 1u) ^ 1u);   // drained by the epilogue warps
        tcgen05_fence_after();
        const uint32_t acc = tmem_acc + (uint32_t)(buf * BN);
        const int len = min(CH, nkb - c * CH);
        for (int j = 0; j < len; ++j, ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
          mbar_wait(SPLIT_MODE >= 2 ? full_bar(s) : split_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
          const uint64_t a_hi = make_desc_kmajor<BK>(stage);
          const uint64_t b_hi = make_desc_kmajor<BK>(stage + BM * BK * 4);
          const uint64_t a_lo = make_desc_kmajor<BK>(stage + Cfg::HI_BYTES);
          const uint64_t b_lo = make_desc_kmajor<BK>(stage + Cfg::HI_BYTES + BM * BK * 4);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            if constexpr (SPLIT_MODE == 3) {
              umma_tf32(acc, a_hi + adv, b_hi + adv, IDESC, (j > 0 || k > 0) ? 1u : 0u);
            } else {
              umma_tf32(acc, a_lo + adv, b_hi + adv, IDESC, (j > 0 || k > 0) ? 1u : 0u);
              umma_tf32(acc, a_hi + adv, b_lo + adv, IDESC, 1u);
              umma_tf32(acc, a_hi + adv, b_hi + adv, IDESC, 1u);
            }
          }
          tcgen05_commit(empty_bar(s));
        }
        tcgen05_commit(accfull_bar(buf));
      }
    }
  } else {
    const int st = threadIdx.x - 64;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float acc[BN / 2];
#pragma unroll
    for (int j = 0; j < BN / 2; ++j) acc[j] = 0.f;
    int drained = 0;
    auto drain = [&](int c) {
      const int buf = c & 1;
      mbar_wait(accfull_bar(buf), ((uint32_t)(c >> 1)) & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int cc = 0; cc < BN / 2; cc += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2) + cc);
        TMEM_LD_32x32b_X32(taddr, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cc + j] += __uint_as_float(r[j]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty_bar(buf));
    };
    if constexpr (SPLIT_MODE >= 2) {
      for (int c = 0; c < nchains; ++c) drain(c);
    } else {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        const float4* hi = reinterpret_cast<const float4*>(smem_gen + s * Cfg::STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem_gen + s * Cfg::STAGE_BYTES + Cfg::HI_BYTES);
        constexpr int NV = (SPLIT_MODE == 1 ? BM * BK * 4 : Cfg::HI_BYTES) / 16;
#pragma unroll 4
        for (int v = st; v < NV; v += 32 * NUM_SPLIT_WARPS) {
          const float4 x = hi[v];
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          lo[v] = l;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(s));
        // chain `drained` ended at k-block (drained+1)*CH-1; once the ring has wrapped past it its MMAs are
        // complete (their stage was re-filled), so the wait below is (almost) free
        if (drained < nchains && (drained + 1) * CH - 1 + STAGES <= i) {
          drain(drained);
          ++drained;
        }
      }
      for (; drained < nchains; ++drained) drain(drained);
    }

    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    float scale = 1.f, bias = 0.f;
    float* dst_row;
    if (p.ksplit > 1) {
      dst_row = p.workspace + ((size_t)blockIdx.z * p.M + (size_t)(row_ok ? m : 0)) * p.ldw;
    } else {
      dst_row = p.C + (size_t)(row_ok ? m : 0) * p.ldc;
      if (row_ok && p.rscale) scale = p.rscale[m];
      if (row_ok && p.rbias) bias = p.rbias[m];
    }
    const bool vec_ok = ((p.ksplit > 1 ? p.ldw : p.ldc) % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.ksplit > 1 ? p.workspace : p.C) & 15) == 0);
    const int nbase = n0 + half * (BN / 2);
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < BN / 2; j += 4) {
        const int n = nbase + j;
        if (vec_ok && n + 4 <= p.N) {
          float4 o;
          o.x = fmaf(acc[j + 0], scale, bias);
          o.y = fmaf(acc[j + 1], scale, bias);
          o.z = fmaf(acc[j + 2], scale, bias);
          o.w = fmaf(acc[j + 3], scale, bias);
          *reinterpret_cast<float4*>(dst_row + n) = o;
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (n + t < p.N) dst_row[n + t] = fmaf(acc[j + t], scale, bias);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(2 * BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster on one TPC compute a 256 x 256 tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 rows), so per-SM shared-memory
// traffic for B (TMA fill, lo split, tensor-core operand reads) is halved — shared-memory
// bandwidth is what caps the single-CTA kernel at ~60 % tensor-pipe utilisation (ncu, round 1).
// The leader CTA's MMA thread issues tcgen05.mma.cta_group::2; both CTAs' splitter warps arrive on
// the LEADER's barrier (remote mbarrier arrive); tcgen05.commit multicasts "stage free" and
// "accumulator ready" to both CTAs.  Every CTA drains its own 128 TMEM lanes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    const uint64_t now = globaltimer_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) {
      printf("facemap_b200 gemm(pair): mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BK>
struct PairCfg {
  static constexpr int BN = 256;                           // UMMA N; each CTA stages BN/2 rows of B
  static constexpr int HI_BYTES = (BM + BN / 2) * BK * 4;  // raw A tile then raw half-B tile
  static constexpr int STAGE_BYTES = 2 * HI_BYTES;         // + lo tiles
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
};

template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmParams p) {
  using Cfg = PairCfg<BK>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = Cfg::BN;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs)
  const int pair_m0 = (blockIdx.x >> 1) * (2 * BM);          // the pair = 2 consecutive CTAs along x
  const int m0 = pair_m0 + (int)rank * BM;
  const int n0 = blockIdx.y * BN;
  if (p.upper_only && n0 + BN - 1 < pair_m0) return;       // whole pair below the diagonal (cluster-uniform)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // local: TMA landed
  auto split_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };     // leader's: both CTAs split
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); }; // local: MMAs done with the stage
  const uint32_t tmem_full_bar = bar_base + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 1);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = (p.K + BK - 1) / BK;
  const int kb0 = (int)(((long long)blockIdx.z * KB) / p.ksplit);
  const int kb1 = (int)(((long long)(blockIdx.z + 1) * KB) / p.ksplit);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), 2 * NUM_SPLIT_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // peer barriers initialised, TMEM allocated
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp == 0) {
    // ===== TMA producer (each CTA fills its own stage buffers) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(full_bar(s), (uint32_t)Cfg::HI_BYTES);
        tma_load_2d(stage, &tmA, full_bar(s), (kb0 + i) * BK, m0);
        tma_load_2d(stage + BM * BK * 4, &tmB, full_bar(s), (kb0 + i) * BK, n0 + (int)rank * (BN / 2));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both tensor cores =====
    if (rank == 0 && lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((2 * BM) >> 4) << 24);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait_cluster(split_bar(s), ph);
        tcgen05_fence_after();
        const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
        const uint64_t a_hi = make_desc_kmajor<BK>(stage);
        const uint64_t b_hi = make_desc_kmajor<BK>(stage + BM * BK * 4);
        const uint64_t a_lo = make_desc_kmajor<BK>(stage + Cfg::HI_BYTES);
        const uint64_t b_lo = make_desc_kmajor<BK>(stage + Cfg::HI_BYTES + BM * BK * 4);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);
          umma_tf32_pair(tmem_acc, a_lo + adv, b_hi + adv, IDESC, (i > 0 || k > 0) ? 1u : 0u);
          umma_tf32_pair(tmem_acc, a_hi + adv, b_lo + adv, IDESC, 1u);
          umma_tf32_pair(tmem_acc, a_hi + adv, b_hi + adv, IDESC, 1u);
        }
        tcgen05_commit_pair(empty_bar(s));
      }
      tcgen05_commit_pair(tmem_full_bar);
    }
  } else {
    // ===== splitter (both CTAs arrive on the leader's barrier), then epilogue =====
    const int st = threadIdx.x - 64;
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      const float4* hi = reinterpret_cast<const float4*>(smem_gen + s * Cfg::STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(smem_gen + s * Cfg::STAGE_BYTES + Cfg::HI_BYTES);
      constexpr int NV = Cfg::HI_BYTES / 16;
#pragma unroll 4
      for (int v = st; v < NV; v += 32 * NUM_SPLIT_WARPS) {
        const float4 x = hi[v];
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        lo[v] = l;
      }
      asm volatile("fence.proxy.async;" ::: "memory");   // generic writes -> async proxy, all state spaces
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(split_bar(s), 0);
    }

    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    float scale = 1.f, bias = 0.f;
    float* dst_row;
    if (p.ksplit > 1) {
      dst_row = p.workspace + ((size_t)blockIdx.z * p.M + (size_t)(row_ok ? m : 0)) * p.ldw;
    } else {
      dst_row = p.C + (size_t)(row_ok ? m : 0) * p.ldc;
      if (row_ok && p.rscale) scale = p.rscale[m];
      if (row_ok && p.rbias) bias = p.rbias[m];
    }
    const bool vec_ok = ((p.ksplit > 1 ? p.ldw : p.ldc) % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.ksplit > 1 ? p.workspace : p.C) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 2; c += 32) {
      const int col = half * (BN / 2) + c;
      uint32_t r[32];
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
      TMEM_LD_32x32b_X32(taddr, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int n = n0 + col;
      if (row_ok && n < p.N) {
        if (vec_ok && n + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = fmaf(__uint_as_float(r[j + 0]), scale, bias);
            o.y = fmaf(__uint_as_float(r[j + 1]), scale, bias);
            o.z = fmaf(__uint_as_float(r[j + 2]), scale, bias);
            o.w = fmaf(__uint_as_float(r[j + 3]), scale, bias);
            *reinterpret_cast<float4*>(dst_row + n + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < p.N) dst_row[n + j] = fmaf(__uint_as_float(r[j]), scale, bias);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer's tensor core may still be reading this CTA's shared memory / TMEM
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Plain FP32 FFMA kernel with the same interface: the on-device numerical reference the tests
// hold the tensor-core kernel against (never used by the pipeline).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_fp32_ref_kernel(const float* __restrict__ A, int64_t lda,
                                                            const float* __restrict__ B, int64_t ldb,
                                                            const GemmParams p) {
  __shared__ float sA[16][65];
  __shared__ float sB[16][65];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int KB = (p.K + REF_BK - 1) / REF_BK;
  const int k_begin = (int)(((long long)blockIdx.z * KB) / p.ksplit) * REF_BK;
  const int k_end = min(p.K, (int)(((long long)(blockIdx.z + 1) * KB) / p.ksplit) * REF_BK);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      const int k = k0 + c;
      sA[c][r] = (m0 + r < p.M && k < k_end) ? A[(size_t)(m0 + r) * lda + k] : 0.f;
      sB[c][r] = (n0 + r < p.N && k < k_end) ? B[(size_t)(n0 + r) * ldb + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sB[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      if (p.ksplit > 1) {
        p.workspace[((size_t)blockIdx.z * p.M + m) * p.ldw + n] = acc[i][j];
      } else {
        const float sc = p.rscale ? p.rscale[m] : 1.f;
        const float bi = p.rbias ? p.rbias[m] : 0.f;
        p.C[(size_t)m * p.ldc + n] = fmaf(acc[i][j], sc, bi);
      }
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

static int make_map(CUtensorMap* map, const float* base, int64_t ld, int rows, int K, int box_rows, int bk) {
  auto enc = tensor_map_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable (driver too old?)");
    return FM_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p ld %lld rows %d K %d)", (int)r, base,
              (long long)ld, rows, K);
    return FM_ERR_CUDA;
  }
  return FM_OK;
}

template <int BN, int BK, bool HI_INPLACE, bool B_PRESPLIT = false, bool A_PRESPLIT = false>
static int launch_tc(const float* A, int64_t lda, const float* B, int64_t ldb, const GemmParams& p, cudaStream_t st,
                     const float* B_lo = nullptr, int64_t ldb_lo = 0, const float* A_lo = nullptr,
                     int64_t lda_lo = 0) {
  CUtensorMap ta, tb, talo, tblo;
  int rc;
  if ((rc = make_map(&ta, A, lda, p.M, p.K, BM, BK)) != FM_OK) return rc;
  if ((rc = make_map(&tb, B, ldb, p.N, p.K, BN, BK)) != FM_OK) return rc;
  if (B_PRESPLIT) {
    if ((rc = make_map(&tblo, B_lo, ldb_lo, p.N, p.K, BN, BK)) != FM_OK) return rc;
  } else {
    tblo = tb;
  }
  if (A_PRESPLIT) {
    if ((rc = make_map(&talo, A_lo, lda_lo, p.M, p.K, BM, BK)) != FM_OK) return rc;
  } else {
    talo = ta;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, BK, HI_INPLACE, B_PRESPLIT, A_PRESPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GemmCfg<BN, BK>::SMEM_BYTES);
  });
  FM_CUDA_TRY(attr_err);
  dim3 grid((unsigned)cdiv(p.N, BN), (unsigned)cdiv(p.M, BM), (unsigned)p.ksplit);
  gemm_tf32x3_kernel<BN, BK, HI_INPLACE, B_PRESPLIT, A_PRESPLIT><<<grid, GEMM_THREADS, GemmCfg<BN, BK>::SMEM_BYTES, st>>>(ta, tb, talo, tblo, p);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

template <int BK, int SPLIT_MODE>
static int launch_chain(const float* A, int64_t lda, const float* B, int64_t ldb, const GemmParams& p, cudaStream_t st,
                        const float* B_lo, int64_t ldb_lo, const float* A_lo, int64_t lda_lo) {
  constexpr int BN = 256;
  CUtensorMap ta, tb, talo, tblo;
  int rc;
  if ((rc = make_map(&ta, A, lda, p.M, p.K, BM, BK)) != FM_OK) return rc;
  if ((rc = make_map(&tb, B, ldb, p.N, p.K, BN, BK)) != FM_OK) return rc;
  tblo = tb;
  talo = ta;
  if ((SPLIT_MODE == 1 || SPLIT_MODE == 2) && (rc = make_map(&tblo, B_lo, ldb_lo, p.N, p.K, BN, BK)) != FM_OK) return rc;
  if (SPLIT_MODE == 2 && (rc = make_map(&talo, A_lo, lda_lo, p.M, p.K, BM, BK)) != FM_OK) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tf32x3_chain_kernel<BK, SPLIT_MODE>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN, BK>::SMEM_BYTES);
  });
  FM_CUDA_TRY(attr_err);
  dim3 grid((unsigned)cdiv(p.N, BN), (unsigned)cdiv(p.M, BM), (unsigned)p.ksplit);
  gemm_tf32x3_chain_kernel<BK, SPLIT_MODE><<<grid, GEMM_THREADS, GemmCfg<BN, BK>::SMEM_BYTES, st>>>(ta, tb, talo, tblo, p);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

template <int BK>
static int launch_pair(const float* A, int64_t lda, const float* B, int64_t ldb, const GemmParams& p, cudaStream_t st) {
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_map(&ta, A, lda, p.M, p.K, BM, BK)) != FM_OK) return rc;
  if ((rc = make_map(&tb, B, ldb, p.N, p.K, PairCfg<BK>::BN / 2, BK)) != FM_OK) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tf32x3_pair_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    PairCfg<BK>::SMEM_BYTES);
  });
  FM_CUDA_TRY(attr_err);
  // cluster shape (2,1,1) is compiled into the kernel (__cluster_dims__); grid.x = 2 CTAs per 256-row pair
  dim3 grid((unsigned)(2 * cdiv(p.M, 2 * BM)), (unsigned)cdiv(p.N, PairCfg<BK>::BN), (unsigned)p.ksplit);
  FM_REQUIRE(grid.y <= 65535, "fm_gemm_tn(pair): N too large");
  gemm_tf32x3_pair_kernel<BK><<<grid, GEMM_THREADS, PairCfg<BK>::SMEM_BYTES, st>>>(ta, tb, p);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

}  // namespace fm

extern "C" int64_t fm_gemm_workspace_bytes(int M, int64_t ldw, int ksplit) {
  return (int64_t)ksplit * M * ldw * 4;
}

extern "C" int fm_gemm_tn(const float* A, int64_t lda, const float* A_lo, int64_t lda_lo, const float* B,
                          int64_t ldb, const float* B_lo, int64_t ldb_lo, int M, int N, int K, float* C,
                          int64_t ldc, const float* rscale,
                          const float* rbias, int ksplit, float* workspace, int64_t ldw, int upper_only, int impl,
                          void* stream) {
  using namespace fm;
  FM_REQUIRE(A && B && M >= 1 && N >= 1 && K >= 1, "fm_gemm_tn: bad shape M=%d N=%d K=%d", M, N, K);
  FM_REQUIRE(lda >= K && ldb >= K, "fm_gemm_tn: leading dimensions smaller than K");
  FM_REQUIRE(ksplit >= 1, "fm_gemm_tn: ksplit must be >= 1");
  const int KB = (int)cdiv(K, REF_BK);
  if (ksplit > KB) ksplit = KB;
  if (ksplit > 1) {
    FM_REQUIRE(workspace && ldw >= N, "fm_gemm_tn: split-K needs a workspace with ldw >= N");
  } else {
    FM_REQUIRE(C && ldc >= N, "fm_gemm_tn: C missing or ldc < N");
  }
  FM_REQUIRE(ksplit <= 65535 && cdiv(M, 64) <= 65535, "fm_gemm_tn: grid too large");
  GemmParams p{M, N, K, C, ldc, rscale, rbias, ksplit, workspace, ldw, upper_only};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == 1) {
    dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 64), (unsigned)ksplit);
    gemm_fp32_ref_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, p);
    FM_LAUNCH_CHECK();
    return FM_OK;
  }
  FM_REQUIRE(impl == 0 || (impl >= 2 && impl <= 10), "fm_gemm_tn: unknown impl %d", impl);
  FM_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
             "fm_gemm_tn: TMA operands need 16-byte aligned bases and ld %% 4 == 0 (lda=%lld ldb=%lld)",
             (long long)lda, (long long)ldb);
  if (B_lo != nullptr)
    FM_REQUIRE(ldb_lo >= K && ldb_lo % 4 == 0 && (reinterpret_cast<uintptr_t>(B_lo) & 15) == 0,
               "fm_gemm_tn: B_lo needs a 16-byte aligned base and ldb_lo %% 4 == 0");
  if (A_lo != nullptr)
    FM_REQUIRE(B_lo != nullptr && lda_lo >= K && lda_lo % 4 == 0 && (reinterpret_cast<uintptr_t>(A_lo) & 15) == 0,
               "fm_gemm_tn: A_lo needs B_lo, a 16-byte aligned base and lda_lo %% 4 == 0");
  // impl 0 = product configuration: chained accumulation (TMEM chains of 128 K drained into register
  // accumulators) on 128 x 256 tiles; outputs at most 128 wide use the 128 x 128 single-accumulator kernel
  if (impl == 0 && N > 128) {
    if (B_lo != nullptr && A_lo != nullptr) return launch_chain<16, 2>(A, lda, B, ldb, p, st, B_lo, ldb_lo, A_lo, lda_lo);
    if (B_lo != nullptr) return launch_chain<16, 1>(A, lda, B, ldb, p, st, B_lo, ldb_lo, nullptr, 0);
    return launch_chain<16, 0>(A, lda, B, ldb, p, st, nullptr, 0, nullptr, 0);
  }
  // 10: the chained kernel with ONE product per K step (operands exact in TF32: slices, integer planes)
  if (impl == 10) {
    FM_REQUIRE(N > 128 && A_lo == nullptr && B_lo == nullptr,
               "fm_gemm_tn: impl 10 (single product) serves outputs wider than 128 and takes no lo planes");
    return launch_chain<16, 3>(A, lda, B, ldb, p, st, nullptr, 0, nullptr, 0);
  }
  // 2..8: tile-shape / CTA-pair variants, 9: the single-accumulator kernel (one TMEM chain per launch);
  // kept for the tests and tuning runs
  switch (impl) {
    case 2: return launch_tc<256, 32, false>(A, lda, B, ldb, p, st);
    case 3: return launch_tc<128, 32, false>(A, lda, B, ldb, p, st);
    case 4: return launch_tc<128, 16, false>(A, lda, B, ldb, p, st);
    case 5: return launch_tc<256, 16, true>(A, lda, B, ldb, p, st);   // explicit hi rewrite (== impl 9 bit for bit)
    case 6: return launch_tc<128, 32, true>(A, lda, B, ldb, p, st);
    case 7: return launch_pair<16>(A, lda, B, ldb, p, st);            // CTA pair, cta_group::2
    case 8: return launch_pair<32>(A, lda, B, ldb, p, st);
    default: break;
  }
  if (B_lo != nullptr) {
    if (A_lo != nullptr)
      return (N > 128) ? launch_tc<256, 16, false, true, true>(A, lda, B, ldb, p, st, B_lo, ldb_lo, A_lo, lda_lo)
                       : launch_tc<128, 32, false, true, true>(A, lda, B, ldb, p, st, B_lo, ldb_lo, A_lo, lda_lo);
    return (N > 128) ? launch_tc<256, 16, false, true>(A, lda, B, ldb, p, st, B_lo, ldb_lo)
                     : launch_tc<128, 32, false, true>(A, lda, B, ldb, p, st, B_lo, ldb_lo);
  }
  return (N > 128) ? launch_tc<256, 16, false>(A, lda, B, ldb, p, st) : launch_tc<128, 32, false>(A, lda, B, ldb, p, st);
}
