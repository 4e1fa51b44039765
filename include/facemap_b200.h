/*
 * facemap_b200 — C-ABI of the B200-native motion/movie-SVD hot path.
 *
 * The reference (MouseLand/facemap) has no FFI: its seams are NumPy expressions inside
 * facemap/process.py and facemap/utils.py.  Each entry point below replaces one such seam; the
 * reference lines it stands in for are cited on every declaration (paths relative to the
 * reference checkout).  The Python host (facemap_b200/process.py) binds these with ctypes — see
 * INTEGRATION.md for the stub a facemap maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all launches go to
 *     that stream, nothing synchronises unless stated;
 *   - matrices are row-major float32 with an explicit leading dimension `ld*` counted in
 *     elements; GEMM operands need ld % 4 == 0 and 16-byte aligned bases (TMA requirement);
 *   - return value: 0 on success, non-zero on error; fm_last_error() gives the text of the
 *     last error raised on the calling thread;
 *   - there is no CPU fallback: without a Blackwell (sm_100) device every compute entry point
 *     fails with FM_ERR_NO_DEVICE.
 */
#ifndef FACEMAP_B200_H
#define FACEMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_OK 0
#define FM_ERR_INVALID 1
#define FM_ERR_CUDA 2
#define FM_ERR_NO_DEVICE 3
#define FM_ERR_SOLVER 4

#define FM_MAX_ROIS 8

/* library / device ----------------------------------------------------------------------- */
int fm_version(void);
const char* fm_last_error(void);
/* 0 if device `dev` is an sm_100 part this library was built for. */
int fm_check_device(int dev);
/* number of kernels this library launched on the calling process since load (all threads). */
uint64_t fm_launch_count(void);

/* Page-lock / unlock caller-owned host memory in place (HostArraySource: frames stay where the
 * caller put them and are DMA'd from there).  A refusal returns non-zero and leaves no sticky error. */
int fm_host_register(void* ptr_host, size_t bytes);
int fm_host_unregister(void* ptr_host);

/* ROI rectangles in BINNED coordinates of one view, half-open [y0,y1) x [x0,x1).
 * facemap/process.py:716-723 (yrange_bin / xrange_bin), :464-500 (use in the projection loop) */
typedef struct fm_rois {
  int32_t n;
  int32_t y0[FM_MAX_ROIS], y1[FM_MAX_ROIS], x0[FM_MAX_ROIS], x1[FM_MAX_ROIS];
} fm_rois;

/*
 * K1 — fused spatial_bin + frame diff + |.| + average subtract (+ motion energy, + running sums).
 * Replaces, for one view: spatial_bin (facemap/process.py:34-43), np.abs(np.diff(.)) and the
 * avg subtraction (:201-212 stage 2, :448-463 stage 3), imbin_mot.sum(-1) (:460, :480) and the
 * per-segment sums of subsampled_mean (:77-87).
 *
 *   frames      uint8 [T, H, W] contiguous
 *   prev_sum    int32 [Lyb*Lxb] integer block sums of the frame preceding frames[0] (the carried
 *               `imend`, process.py:449-453) or NULL.  With prev_sum: rows = T, row r <-> frame r.
 *               Without: rows = T-1, row r <-> frame r+1 (first chunk, process.py:455-457).
 *   avgmotion / avgframe   float32 [Lyb*Lxb] or NULL (NULL = subtract nothing)
 *   x_mot       float32 [rows, ldx]; written at columns [col0, col0 + Lyb*Lxb):
 *               f32( |b_t - b_{t-1}| - avgmotion ), b = blocksum / sbin^2 evaluated in float64
 *   x_mov       float32 [rows, ldx]: f32( b_t - avgframe )                     (either may be NULL)
 *   energy      uint64 [rows]: += sum over the view of |S_t - S_{t-1}| (integer block sums S);
 *               caller zeroes.  motion = energy / sbin^2.
 *   roi_energy  uint64 [rois->n, rows]: the same restricted to each ROI rectangle (or NULL)
 *   last_sum    int32 [Lyb*Lxb]: block sums of frames[T-1] (next call's prev_sum) or NULL
 *   tsum_bin    int64 [Lyb*Lxb]: += sum_t S_t over ALL T frames          (stage 1) or NULL
 *   tsum_adiff  int64 [Lyb*Lxb]: += sum over rows of |S_t - S_{t-1}|     (stage 1) or NULL
 */
int fm_bin_motion(const uint8_t* frames, int T, int H, int W, int sbin,
                  const int32_t* prev_sum, const float* avgmotion, const float* avgframe,
                  float* x_mot, float* x_mov, int64_t ldx, int64_t col0,
                  unsigned long long* energy, const fm_rois* rois, unsigned long long* roi_energy,
                  int32_t* last_sum, long long* tsum_bin, long long* tsum_adiff, void* stream);

/*
 * The same on frames that pupil / blink ROIs have painted (facemap/process.py:543, 567 write 255 into the
 * chunk buffer that stage 3 bins afterwards): every frame byte is OR-ed with `ormask` (uint8 [H, W], 0x00 or
 * 0xFF) as it is loaded, so no painted copy of the frames is needed.  All other arguments as fm_bin_motion.
 */
int fm_bin_motion_painted(const uint8_t* frames, int T, int H, int W, int sbin,
                          const int32_t* prev_sum, const float* avgmotion, const float* avgframe,
                          float* x_mot, float* x_mov, int64_t ldx, int64_t col0,
                          unsigned long long* energy, const fm_rois* rois, unsigned long long* roi_energy,
                          int32_t* last_sum, long long* tsum_bin, long long* tsum_adiff,
                          const uint8_t* ormask, void* stream);

/*
 * K1 writing its integer results instead of the float32 operands (the stage-3 default) — the motion
 * energy |S_t - S_(t-1)| (q_mot_*) and / or the block sum S_t (q_mov_*) of every binned pixel (facemap/process.py:34-43,
 * 455-461 before the division and the subtraction of the average) as two uint8 planes, value = 256 * hi + lo, rows and
 * columns like x_mot / x_mov of fm_bin_motion, row pitch ldq bytes.  These are the frame operands of fm_project_i8.
 * sbin <= 16; the high planes may be NULL for sbin == 1; either pair may be NULL; ormask may be NULL.  energy,
 * roi_energy, last_sum as in fm_bin_motion.
 */
int fm_bin_motion_planes(const uint8_t* frames, int T, int H, int W, int sbin, const int32_t* prev_sum,
                         uint8_t* q_mot_hi, uint8_t* q_mot_lo, uint8_t* q_mov_hi, uint8_t* q_mov_lo, int64_t ldq,
                         int64_t col0, unsigned long long* energy, const fm_rois* rois,
                         unsigned long long* roi_energy, int32_t* last_sum, const uint8_t* ormask, void* stream);

/*
 * Motion energy in the reference's own floating-point arithmetic (facemap/process.py:34-43, 455-460): per row
 * the sum over the view's binned pixels of |b_t - b_{t-1}|, b = float64 mean of means (sbin > 1) or float32
 * pixel value (sbin == 1), summed in NumPy's pairwise order — bit-identical to `imbin_mot.sum(-1)`.  Needed
 * where K1's exact integer energy cannot decide the reference's float32 accumulator: sbin not a power of two
 * (float32 rounding ties between views) and sbin == 1 on views of more than 65 793 pixels (float32 sums
 * beyond 2^24).  Row convention as fm_bin_motion: with prev_frame (uint8 [H, W], the frame before frames[0])
 * rows = T, else rows = T - 1.  leaf_*: fm_pairwise_plan(Lyb*Lxb) uploaded to the device.  out: float64 [rows].
 */
int fm_pairwise_plan(int n, int cap, int32_t* lo, int32_t* len, uint8_t* comb);   /* host arrays; returns #leaves */
int fm_motion_ref_scratch_bytes(int H, int W, int sbin, int nleaves, int blocks, size_t* bytes);
int fm_motion_ref(const uint8_t* frames, int T, int H, int W, int sbin, const uint8_t* prev_frame,
                  const uint8_t* ormask, const int32_t* leaf_lo, const int32_t* leaf_len,
                  const uint8_t* leaf_comb, int nleaves, void* scratch, int blocks, double* out, void* stream);
/* sbin == 1, everything stage 3 needs of a view in ONE pass over its frames: `out` as fm_motion_ref, the |difference|
 * bytes as the projection's operand plane (plane [rows, ldq] at columns col0 ..: what fm_bin_motion_planes writes at sbin 1)
 * and the exact integer energy (energy [rows] +=, may be NULL).  grp_* (ngroups <= 256, or 0): the maximal subtrees of the
 * pairwise tree holding at most 65 793 byte values — their float32 sums never round, so they are summed as integers by
 * whole warps and only the merges above them are folded in order (kernels.exact_groups); without them one thread folds
 * all leaves.  facemap/process.py:448-461 on full-resolution views. */
int fm_motion_ref_planes(const uint8_t* frames, int T, int H, int W, const uint8_t* prev_frame, const uint8_t* ormask,
                         const int32_t* leaf_lo, const int32_t* leaf_len, const uint8_t* leaf_comb, int nleaves,
                         const int32_t* grp_first, const int32_t* grp_n, const uint8_t* grp_comb, int ngroups,
                         void* scratch, int blocks, double* out, uint8_t* plane, int64_t ldq, int64_t col0,
                         uint64_t* energy, void* stream);

/*
 * Stage-1 segment means in the reference's float64 arithmetic for a non-power-of-two sbin (process.py:77-87):
 * mean_bin[p] = mean over the T frames of the binned value, mean_adiff[p] = mean over the T-1 absolute
 * differences, both accumulated in frame order like `imbin.mean(axis=0)`.  The caller adds them into the float32
 * averages (`avg32 = f32(f64(avg32) + mean)`).  Power-of-two sbin uses K1's integer time sums (exact).
 */
int fm_mean_ref(const uint8_t* frames, int T, int H, int W, int sbin, double* mean_bin, double* mean_adiff,
                void* stream);

/*
 * Stage-1 accumulator update, bit-compatible with subsampled_mean (facemap/process.py:82-86,
 * 95-96): acc32 = f32( f64(acc32) + (f64(tsum) / sbin^2) / count ) for n pixels.
 * `finalize_div` > 0 additionally divides by that segment count in float32 (process.py:95-96).
 */
int fm_mean_update(float* acc, const long long* tsum, int64_t n, int sbin, int count,
                   int finalize_div, void* stream);

/*
 * K2/K3/K4 — C[M,N] = rscale[m] * (A[M,K] . B[N,K]^T) + rbias[m]      (3xTF32 on tcgen05/TMEM)
 * A and B are both K-contiguous (row-major [rows, K]).  Replaces the dense contractions of the
 * path: X @ U (facemap/process.py:485,495,503,510), and the GEMMs inside utils.svdecon
 * (facemap/utils.py:751-776 -> sklearn randomized_svd; here Gram + eigensolve, the formulation
 * of matlab/utils/svdecon.m:16-43).
 *   B_lo            NULL, or the pre-computed lo plane of B (fm_split_lo): TMA then stages both B
 *                   tiles and only A is split in the kernel (impl 0 only)
 *   A_lo            NULL, or the lo plane of A (only together with B_lo): no in-kernel split at all
 *   rscale, rbias   float32 [M] or NULL
 *   ksplit >= 1: if > 1, `C` is ignored and raw partial products are written to
 *                workspace[ksplit, M, ldw] (no scale/bias); reduce with fm_gram_assemble.
 *   upper_only != 0: tiles entirely below the diagonal are skipped (symmetric Gram).
 *   impl: 0 = tcgen05 3xTF32 (product path), 1 = plain FP32 FFMA reference kernel (validation);
 *         2..8 = tile-shape / CTA-pair (cta_group::2) variants of the tensor-core kernel kept for the
 *         tests and tuning runs (same results, see csrc/gemm_tf32x3.cu); 10 = one product per K step on the
 *         operands as they are (no split: for operands exact in TF32 such as fm_slice_rows planes; N > 128).
 *   ksplit > 1 with ldw: reduce with fm_splitk_reduce (general) or fm_gram_assemble (Gram).
 */
int fm_gemm_tn(const float* A, int64_t lda, const float* A_lo, int64_t lda_lo, const float* B, int64_t ldb,
               const float* B_lo, int64_t ldb_lo, int M, int N, int K, float* C, int64_t ldc,
               const float* rscale, const float* rbias,
               int ksplit, float* workspace, int64_t ldw, int upper_only, int impl, void* stream);

/*
 * The digit-plane product of the Gram matrices (svd.i8_gram, the default):
 * C[M,N] (int32) = A[M,K] . B[N,K]^T with 8-bit operands, K-contiguous, on tcgen05.mma kind::i8 with exact int32
 * accumulation.  a_signed / b_signed: 0 = uint8, 1 = int8.  Row pitches in bytes, multiples of 16; 16-byte aligned
 * bases.  ksplit > 1 writes raw int32 partials to workspace[ksplit, M, ldw] (sum them in int64); at most 32 768 K
 * per partial (int32 range).  For the integer operands of this path — motion energies and block sums
 * (facemap/process.py:34-43, 201-212) as u8 digit planes, masks as s8 slices — see DESIGN.md §7.2.
 */
int fm_gemm_i8_tn(const uint8_t* A, int64_t lda, int a_signed, const uint8_t* B, int64_t ldb, int b_signed,
                  int M, int N, int K, int32_t* C, int64_t ldc, int ksplit, int32_t* workspace, int64_t ldw,
                  int upper_only, void* stream);

/*
 * Companions of fm_gemm_i8_tn (svd.i8_gram, the default way the Gram matrices of both SVD levels are formed):
 * fm_slice_rows_i8: X[r,c] ~ 2^(exps[r] - 7) * (Q0 + Q1 / 2^8 + Q2 / 2^16), Q0 int8 (floor), Q1 uint8 (floor), Q2 uint8
 *   (nearest, carries propagated), 2^exps[r] bounding row r: 23 bits below the row bound, unbiased.  Plane row pitch
 *   ldq in bytes (a multiple of 16 for the GEMM); pad bytes zeroed.
 * fm_gram_assemble_i8: P = int32 [6][ksplit][n][ldp], the plane products 00, 11, 22 (upper triangles), 01, 02, 12
 *   (full) from fm_gemm_i8_tn  ->  G = X X^T - ncols * mean mean^T in float64 (mean may be NULL), exact up to the
 *   2^-24 representation error of the digits.  The Gram form of utils.svdecon (facemap/utils.py:751-776).
 */
int fm_slice_rows_i8(const float* X, int64_t ldx, int rows, int64_t ncols, int8_t* Q0, uint8_t* Q1, uint8_t* Q2,
                     int64_t ldq, int32_t* exps, void* stream);
int fm_gram_assemble_i8(const int32_t* P, int ksplit, int n, int64_t ldp, const int32_t* exps, const double* mean,
                        int64_t ncols, double* G, void* stream);
/* G += the weighted plane products of a further block of columns of the same operand (its own exps; no mean term). */
int fm_gram_accumulate_i8(const int32_t* P, int ksplit, int n, int64_t ldp, const int32_t* exps, double* G,
                          void* stream);
/* The same for the row-side Gram of a wide matrix Y [n, c] centred over its columns (fm_gram_assemble_rowside):
 * G = Y Y^T - q 1^T - 1 q^T + (mu.mu) 1 1^T with q = Y mu (fm_row_dot), mu[nmu] the column means. */
int fm_gram_assemble_i8_rowside(const int32_t* P, int ksplit, int n, int64_t ldp, const int32_t* exps,
                                const double* q, const double* mu, int nmu, double* G, void* stream);

/*
 * The projection of process_ROIs (facemap/process.py:485, 495, 503, 510) on the
 * integer tensor cores.  V[t,c] (float32) = alpha * sum_p D[t,p] U[c,p] - bias[c] with the frame operand as integers
 * D = 256 * D_hi + D_lo (u8 planes [M, ldd]; D_hi may be NULL for operands below 256: sbin 1) — the motion energy
 * |S_t - S_(t-1)| or the block sum S_t that fm_bin_motion forms, alpha = 1 / sbin^2 — and the masks as the digit planes
 * of fm_slice_rows_i8 (Q0 int8, Q1 / Q2 uint8 [N, ldq], exps[N]).  bias[c] = sum_p avg[p] U[c,p] (float64, fm_row_dot;
 * may be NULL) carries the subtraction of avgmotion / avgframe.  Exact int32 accumulation in chains of 16 384 K,
 * int64 across chains, one float64 scaling and one rounding to float32 per output.  Pitches in bytes, multiples of 16;
 * 16-byte aligned plane bases; K <= 2^22.  impl: 0 = one CTA per 128 x 128 tile, 1 = CTA pairs (cta_group::2, 256 x 128
 * per pair, half of every digit tile staged per CTA); same results.  DESIGN.md section 7.2(a).
 */
int fm_project_i8(const uint8_t* D_hi, const uint8_t* D_lo, int64_t ldd, const int8_t* Q0, const uint8_t* Q1,
                  const uint8_t* Q2, int64_t ldq, const int32_t* exps, int M, int N, int K, double alpha,
                  const double* bias, float* V, int64_t ldv, int impl, void* stream);

/* lo[r,c] = X[r,c] - tf32_trunc(X[r,c]): pre-computed second plane of a B operand that many GEMMs
 * reuse (pass it as B_lo above). */
int fm_split_lo(const float* X, int64_t ldx, int rows, int cols, float* lo, int64_t ldlo, void* stream);

/*
 * Row-scaled 8-bit slices of a Gram operand: X[r,c] = S0 + S1 + S2 + rest, S_d = k_d * 2^(e_r - 8(d+1)) with k_d
 * the nearest integer (signed digits, |k_0| <= 256, |k_1|, |k_2| <= 128), |rest| <= 2^(e_r - 25), 2^e_r bounding
 * row r.  Slices are exact TF32 operands and products of two slices add up exactly in float32 over 240 terms, so
 * fm_gemm_tn on (S0, S0) with K partials of <= 240 yields the leading part of X X^T without any rounding; the
 * cross terms are 2^-8 and 2^-16 of it.  Used for the merge Gram of utils.svdecon (facemap/utils.py:751-776) when
 * its small eigenvalues are wanted to float64 accuracy (svd.sliced_gram).
 * All planes have row pitch lds >= ncols; pad columns are zeroed.
 */
int fm_slice_rows(const float* X, int64_t ldx, int rows, int64_t ncols, float* S0, float* S1, float* S2,
                  int64_t lds, void* stream);

/*
 * Reduction of split-K partials: C[m,n] = rscale[m] * sum_s part[s,m,n] + rbias[m], summed in
 * float64.  tcgen05 accumulates in float32 with truncation, so long contractions are cut into
 * short K chains (fm_gemm_tn with ksplit > 1) and reduced exactly here.
 */
int fm_splitk_reduce(const float* part, int ksplit, int M, int N, int64_t ldw, const float* rscale,
                     const float* rbias, float* C, int64_t ldc, void* stream);

/* bytes of workspace fm_gemm_tn needs for (M, ldw, ksplit). */
int64_t fm_gemm_workspace_bytes(int M, int64_t ldw, int ksplit);

/*
 * PCA centring pieces (sklearn/decomposition/_pca.py:733-736 called from facemap/utils.py:773):
 *   fm_row_means : mean[r] = (1/ncols) sum_c X[r, c]   (float64)
 *   fm_gram_assemble : G[i,j] (float64, n x n) = sum_s part[s,i,j] - ncols * mean[i] * mean[j]
 *                      taking part from the upper triangle when upper_only and mirroring.
 */
int fm_row_means(const float* X, int64_t ldx, int rows, int64_t ncols, double* mean, void* stream);
int fm_gram_assemble(const float* part, int ksplit, int n, int64_t ldw, const double* mean,
                     int64_t ncols, int upper_only, double* G, void* stream);

/*
 * Row-side form of the same PCA for wide matrices (merge of a small ROI: p pixels < c stacked
 * components, facemap/process.py:276-295): with mu = column means of Y [p, c] and q = Y mu,
 *   fm_row_dot : q[r] = sum_c Y[r,c] * w[c]
 *   fm_gram_assemble_rowside : G = sum_s part[s] - q 1^T - 1 q^T + (mu.mu) 1 1^T   (= Yc Yc^T, p x p)
 *   fm_sign_from_right : the sign rule needs the RIGHT vectors, V S = Y^T U - mu (1^T U); given
 *       T[k, c] = Ut Y (fm_gemm_tn) it flips the rows of Ut whose largest-|.| right entry is negative.
 */
int fm_row_dot(const float* Y, int64_t ldy, int rows, int64_t ncols, const double* w, double* q, void* stream);
int fm_gram_assemble_rowside(const float* part, int ksplit, int n, int64_t ldw, const double* q,
                             const double* mu, int nmu, int upper_only, double* G, void* stream);
int fm_sign_from_right(const float* T, int64_t ldt, int k, int ncols, const double* mu, float* Ut,
                       int64_t ldu, int p, void* stream);

/*
 * Symmetric eigensolve (cuSOLVER syevd — the one library call on the path; timed separately,
 * not claimed as a kernel).  G is float64 n x n, overwritten with eigenvectors (row j =
 * eigenvector of the j-th smallest eigenvalue); lam[n] ascending.  `use_fp32` solves in float32
 * (G is converted in place to/from).  The solves are asynchronous like every other entry point:
 * cuSOLVER's devInfo words are kept in a device ring owned by the solver and read back — one copy, one
 * stream synchronisation — by fm_solver_check, which fails on the first non-zero entry since the
 * previous check (the host calls it once per pass, when it collects the results anyway).
 */
typedef struct fm_solver fm_solver;
int fm_solver_create(fm_solver** out);
int fm_solver_destroy(fm_solver* s);
int fm_solver_check(fm_solver* s, void* stream);
int fm_syevd(fm_solver* s, double* G, int n, double* lam, int use_fp32, void* stream);
/* Only the k largest eigenpairs (cusolverDnXsyevdx): rows 0..k-1 of G and lam[0..k) in ascending order. */
int fm_syevdx_top(fm_solver* s, double* G, int n, int k, double* lam, void* stream);
/* `batch` matrices stored back to back (cusolverDnXsyevBatched); lam is [batch, n]. */
int fm_syevd_batched(fm_solver* s, double* G, int n, int batch, double* lam, void* stream);

/*
 * Hand-written eigensolver for batches of small Gram matrices (the per-chunk svdecon of facemap/process.py:229-259 in
 * its Gram form, matlab/utils/svdecon.m:16-43): the k largest eigenpairs of `batch` symmetric n x n float64 matrices
 * stored back to back, 2 <= n <= 1024, entirely on the device — Householder tridiagonalisation by one thread-block
 * cluster per matrix, Sturm bisection, inverse iteration, back-transformation (csrc/eigh_cluster.cu).  G is destroyed.
 * lam [batch, k] and W [batch, k, n] come back in ASCENDING order of eigenvalue (row k-1 = the largest): what
 * fm_topk_components reads with nev = k.  status [batch] (int32) receives self-check flags (bit 0: neighbouring
 * eigenvectors not orthogonal — eigenvalues too close for independent inverse iterations; bit 1: non-finite or
 * non-unit vector); 0 = good.  `scratch` needs fm_eigh_topk_scratch_bytes(n, batch, k) bytes.  Asynchronous.
 */
size_t fm_eigh_topk_scratch_bytes(int n, int batch, int k);
int fm_eigh_topk_batched(double* G, int n, int batch, int k, double* lam, double* W, void* scratch,
                         size_t scratch_bytes, int* status, void* stream);

/*
 * Top-k extraction with the reference's sign rule (sklearn/utils/extmath.py:924-982,
 * u_based_decision=False: the largest-|.| entry of each right singular vector is positive):
 * evec/lam hold `nev` eigenpairs in ascending order (nev = n after fm_syevd, nev = k' after
 * fm_syevdx_top), evec row-major with row length n:
 *   Wt[c, j]  = sgn_c * evec[nev-1-c][j]                  float32 [k, ldwt]  (c = 0..k-1)
 *   sval[c]   = sqrt(max(lam[nev-1-c], 0))                float32 [k]
 *   rscale[c] = 1 / sval[c] (0 if sval == 0)              float32 [k]   (may be NULL)
 *   rbias[c]  = - sum_j Wt[c, j] * mean[j]                float32 [k]   (centring of X^T W; may
 *               be NULL); when rscale is requested the bias is pre-multiplied by rscale[c] so
 *               that fm_gemm_tn's epilogue rscale*acc + rbias gives (acc - w.mean) / sval.
 */
int fm_topk_components(const double* evec, const double* lam, int n, int nev, int k,
                       const double* mean, float* Wt, int64_t ldwt, float* sval, float* rbias,
                       float* rscale, void* stream);

/* out[c, r] = in[r, c]  (rows x cols float32 -> cols x rows). */
int fm_transpose(const float* in, int64_t ldin, int rows, int cols, float* out, int64_t ldout,
                 void* stream);

/* ROI column gather (facemap/process.py:218-222, 479-483): out[r, (y-y0)*(x1-x0) + (x-x0)] =
 * X[r, col0 + y*Lxb + x] for the binned rectangle. */
int fm_gather_roi(const float* X, int64_t ldx, int rows, int64_t col0, int Lxb,
                  int y0, int y1, int x0, int x1, float* out, int64_t ldout, void* stream);
/*
 * K-blocked operands for wide views.  fm_retile_u8 copies a byte plane [rows, ld] into the layout
 * dst[(c / kb) * block_stride + r * kb + c % kb] (zero-filled up to the end of the last block); fm_project_i8_blocked is
 * fm_project_i8 on planes in that layout (block strides instead of row pitches; kb a power of two >= 64).  With row pitches
 * beyond ~256 KB (1280x1024 at sbin 1: 1.3 MB) every row of an operand tile lies on its own 2 MB page and the projection
 * drops to a third of its rate; blocked, a 128-row tile is 128 * kb contiguous bytes (facemap/process.py:485, 503 at
 * BASELINE configs[3]).
 */
int fm_retile_u8(const uint8_t* src, int64_t ld, int rows, int64_t cols, uint8_t* dst, int64_t kb, int64_t block_stride,
                 void* stream);
int fm_project_i8_blocked(const uint8_t* D_hi, const uint8_t* D_lo, int64_t d_block_stride, const int8_t* Q0,
                          const uint8_t* Q1, const uint8_t* Q2, int64_t q_block_stride, int64_t kb, const int32_t* exps,
                          int M, int N, int K, double alpha, const double* bias, float* V, int64_t ldv, int impl,
                          void* stream);

/* The same gather on one byte plane of fm_bin_motion_planes (pitches in bytes). */
int fm_gather_roi_u8(const uint8_t* X, int64_t ldx, int rows, int64_t col0, int Lxb,
                     int y0, int y1, int x0, int x1, uint8_t* out, int64_t ldout, void* stream);

/* ---- full-resolution ROI processors riding on the stage-3 frame sweep (facemap/process.py:410-425) ----
 * ROI rectangles are [y0, y1) x [x0, x1) in frame pixels (the reference's yrange[0]..yrange[-1] inclusive,
 * process.py:538-542).  `mask` / `ormask` are uint8 [ny, nx] (ROI crop) or [H, W] (fm_paint): non-zero =
 * the pixel reads as 255 because an earlier ROI painted it (process.py:543, 567 write into the chunk). */

/* out[t] = frames[t] | ormask   (ormask bytes 0x00 / 0xFF): the frames stage 3 bins after painting. */
int fm_paint(const uint8_t* frames, int T, int H, int W, const uint8_t* ormask, uint8_t* out, void* stream);

/* blink (process.py:557-572): out[t] = sum over the painted ROI of term[pixel value]; term = the reference's
 * elementwise expression `maximum(0, 255 - v - (255 - saturation))` tabulated for v = 0..255 as float64. */
int fm_blink(const uint8_t* frames, int T, int H, int W, int y0, int y1, int x0, int x1,
             const uint8_t* mask, const double* term, double* out, void* stream);

/* running (running.py:91-133 as executed — its fft2/ifft2 results are discarded): for frame t and its
 * predecessor, the first argmax over the (2l+1)^2 corner window (l = min(floor(.4 ny), floor(.4 nx))) of
 * sign(x_t - mean_t) * sign(x_{t-1} - mean_{t-1}) * g2, g2 = float32 square of the float32 Fourier-domain
 * Gaussian [ny, nx].  out[t] = (dy, dx) int32; out[0] needs prev_crop (painted uint8 [ny, nx]) and
 * prev_mean (float32 [1]) of the frame before the chunk, else it is left untouched.  means: float32 [T]
 * (written: the NumPy-rounded ROI mean of every frame; means[T-1] is the next call's prev_mean). */
int fm_running(const uint8_t* frames, int T, int H, int W, int y0, int y1, int x0, int x1,
               const uint8_t* ormask, const uint8_t* prev_crop, const float* prev_mean, const float* g2,
               float* means, int32_t* out, void* stream);

/* pupil (pupil.py:8-132): Gaussian fit of darkness = max(0, (255 - v) - dark_offset) (float32; dark_offset =
 * float32(255 - saturation)) over the painted ROI, 5 refinement rounds at `sigma`, optional reflector map
 * (uint8 [ny, nx], non-zero = filled in from the median, utils.py:666-700).  out: float64 [T, 9] =
 * com(2), area, axdir(2x2 row-major), axlen(2); a frame the reference leaves NaN is all NaN.
 * scratch: fm_pupil_scratch_bytes(ny, nx, blocks) bytes; `blocks` persistent thread blocks share the
 * frames, one frame per block at a time. */
int fm_pupil_scratch_bytes(int ny, int nx, int blocks, size_t* bytes);
int fm_pupil(const uint8_t* frames, int T, int H, int W, int y0, int y1, int x0, int x1,
             const uint8_t* mask, const uint8_t* reflector, float dark_offset, double sigma,
             void* scratch, int blocks, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FACEMAP_B200_H */
