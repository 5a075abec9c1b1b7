/*
 * mepp_b200.h -- C ABI of the B200-native MEPP profiling hot path.
 *
 * Plain pointers and sizes only; no torch / C++ types.  Every entry point
 * replaces one piece of the reference's (npdeloss/mepp) per-task profiling path;
 * the reference interface each one stands in for is cited as file:line into the
 * reference tree.  All "dev" pointers are CUDA device pointers on the current
 * device; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *
 * Return value: 0 on success, non-zero on error; mepp_last_error() returns a
 * thread-local message.  Nothing here falls back to the CPU: device entry points
 * fail when no CUDA device / sm_100a is present.
 *
 * Data layouts (all little-endian):
 *   N        number of sequences, L sequence length (all equal, utils.py:85-88)
 *   n_pad    N rounded up to a multiple of 32 (one "k-block" = 32 sequences)
 *   sym_T    uint32 [W3][n_pad], W3 = ceil(3L/32)+8 : per sequence a continuous bit stream with
 *            3 bits per base -- the 2-bit code A=0 C=1 G=2 T=3 (onehot_dna.py:3) plus an N flag
 *            (symbol 4 = not an upper-case ACGT, onehot_dna.py:23-30); base i sits at bits
 *            [3i, 3i+3) of the stream, word w of sequence n is sym_T[w][n] (transposed so that
 *            lanes = sequences read coalesced)
 *   X planes the match matrix as GEMM operand A, fp16 hi/lo split, tiled:
 *            [m_tile][k_block][plane 2][kc 4][row 128][8 seq] halves, row = task*L + position, followed
 *            by a bit array uint32 [m_tile][FW]: bit (2*k_block + h) = "sequences [16h, 16h+16) of the
 *            k-block hold a non-zero value in this m-tile" (all-zero blocks are skipped by the GEMM)
 *   Y planes the (1+P) score columns as GEMM operand B, same tiling with BN columns:
 *            [n_tile][k_block][plane 2][kc 4][col BN][8 seq] halves
 */
#ifndef MEPP_B200_H
#define MEPP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEPP_B200_ABI_VERSION 1
#define MEPP_MAX_MOTIF_LEN 32   /* taps held in a 64-bit code window */
#define MEPP_MAX_MARGIN 16
#define MEPP_ENTRY_LISTS 64     /* sub-lists of the sparse form of X (power of two) */
#define MEPP_BP_WORDS 72        /* uint32 words of a task record of the bit-parallel scan stage */
#define MEPP_X_SCALE 256.0f     /* X is stored as 256*x in the fp16 planes (exact) */

int         mepp_abi_version(void);
const char* mepp_last_error(void);
/* 1 when a CUDA device with compute capability 10.x is present. */
int         mepp_device_ok(void);

/* ------------------------------------------------------------------ host side
 * MOODS.tools.log_odds / threshold_from_p as called at motif_layers.py:37-48.
 * pwm, W: row-major 4 x m (rows A,C,G,T).  bg: 4 doubles.  Pure host code. */
int mepp_log_odds(const double* pwm, int m, const double* bg, double pseudocount, double* W);
int mepp_threshold_from_p(const double* W, int m, const double* bg, double pvalue,
                          double dp_precision, double* thr);

/* Per-task scan table from a 4 x m PWM and an orientation ('+', '-', anything else =
 * dual), replacing motif_matrix_to_conv1d_layer (motif_layers.py:12-78):
 *   lut   float [MEPP_MAX_MOTIF_LEN][4][2]  (tap, base, filter) float32 weights, zero padded
 *   bias  float32(-thr)  ; n_filters 1 or 2 ; thr double (for reporting)
 *   qtab  uint32 [16][64], qthr uint32 [2]: conservative 16-bit integer pre-filter of the scan
 *         kernel (pairs of taps, both filters packed low/high); may be NULL
 * use_flags: 0 = reproduce the reference (pseudocount/pvalue/bg never reach MOODS,
 * core.py:513-534, so 1e-4 / 1e-4 / flat are used); 1 = honour the arguments. */
int mepp_motif_table(const double* pwm, int m, int orientation_char,
                     double pseudocount, double pvalue, const double* bg, int use_flags,
                     double dp_precision,
                     float* lut, float* bias, int* n_filters, double* thr,
                     uint32_t* qtab, uint32_t* qthr);
/* Same table for a threshold that is already known: the '+' and '+/-' orientations of a motif share
 * filter 0 and its threshold (motif_layers.py:36-76), so the DP runs once per motif. */
int mepp_motif_table_with_threshold(const double* pwm, int m, int orientation_char, double pseudocount,
                                    const double* bg_in, int use_flags, double thr, float* lut, float* bias,
                                    int* n_filters, uint32_t* qtab, uint32_t* qthr);
/* Thresholds on the device (the host DP above is the longest host-side piece of an end-to-end step).
 * mepp_prepare_tables (host, no DP): for T tasks given as zero-padded (4, 32) matrices pwms[T][4][32] with lengths
 * mlen[T] and orientations ochar[T] ('+', '-', anything else = both strands): lut float [T][32][4][2], nfilt [T],
 * qtab uint32 [T][1024], qpar double [T][4] (what the pass bounds need besides the threshold), and for the
 * threshold DP of filter 0: smat int32 [T][32][4] (integer scores shifted to >= 0), dp_par double [T][8] =
 * {bg[4], p, precision, m * min score, score range R}.
 * mepp_thresholds_dp (device): threshold_from_p for U units (any subset of the tasks: upload their smat / mlen /
 * dp_par rows); scratch_dev holds two arrays of R_u + 1 doubles per unit at element offset scratch_off_dev[u].
 * mepp_task_thresholds (device): bias = float(-thr), qthr pass bounds and the threshold of every task from the
 * thresholds of the units (unit_of_task: the '+' and '+/-' orientations of a motif share one). */
int mepp_prepare_tables(const double* pwms, const int32_t* mlen, const int32_t* ochar, int T, double pseudocount,
                        double pvalue, const double* bg_in, int use_flags, double dp_precision, float* lut,
                        int32_t* nfilt, uint32_t* qtab, double* qpar, int32_t* smat, double* dp_par);
int mepp_thresholds_dp(const int32_t* smat_dev, const int32_t* mlen_dev, const double* par_dev,
                       const long long* scratch_off_dev, double* scratch_dev, double* thr_dev, int U, void* stream);
int mepp_task_thresholds(const double* thr_unit_dev, const int32_t* unit_of_task_dev, const double* qpar_dev,
                         const int32_t* mlen_dev, const int32_t* nfilt_dev, int T, float* bias_dev, uint32_t* qthr_dev,
                         double* thr_task_dev, void* stream);

/* ------------------------------------------------------------------ layout sizes */
int64_t mepp_sym_words(int L);              /* W3 */
int64_t mepp_pad32(int64_t n);              /* n_pad */
int64_t mepp_x_planes_bytes(int64_t m_rows, int64_t n_pad);
int     mepp_choose_bn(int C, int* n_tiles_n);  /* column tile width (multiple of 32, <=256) */
int64_t mepp_y_planes_bytes(int C, int64_t n_pad);   /* tiles + n_pad uint32 of scratch for mepp_build_y */

/* ------------------------------------------------------------------ K0: pack
 * Replaces seq_to_mat / scored_fasta_df_to_dataset one-hot (onehot_dna.py:61-70,
 * utils.py:141-194) and append_gc_ratios_to_dataset('global') (utils.py:275-308).
 * ascii_dev: uint8 [N][L].  gc_dev: float [n_pad] (z, 0 for padding). */
int mepp_pack_dna(const uint8_t* ascii_dev, int64_t N, int L,
                  uint32_t* sym_T_dev, float* gc_dev, void* stream);

/* ------------------------------------------------------------------ Y operand
 * Replaces add_permuted_scores_to_dataset (permutations.py:5-58).
 * yhat_dev: float [N] centred+scaled scores (host does the float64 centring);
 * perm_dev: int32 [P][N] permutation indices (explicit: the reference's shuffle is
 * unseeded); zhat_dev: float [n_pad] centred GC covariate or NULL.
 * Outputs: y_planes (tiled fp16 hi/lo, zero padded), col_syz double [C] = sum_n y_c z,
 * ysum double[2] = {sum y, sum y^2} of the split-rounded values (same for every column). */
int mepp_build_y(const float* yhat_dev, const int32_t* perm_dev, const float* zhat_dev,
                 int64_t N, int P, void* y_planes_dev, double* col_syz_dev, double* ysum_dev,
                 void* stream);

/* ------------------------------------------------------------------ K1: scan
 * Replaces the conv model + ReLU + margin blur + counts/density
 * (motif_layers.py:80-165, core.py:25-145, core.py:415-470) for a group of T tasks.
 *   lut_dev   float [T][MEPP_MAX_MOTIF_LEN][4][2]   bias_dev float [T]
 *   mlen_dev  int32 [T]   nfilt_dev int32 [T]   qtab_dev uint32 [T][1024]   qthr_dev uint32 [T][2]
 *   units_dev optional int32 [n_units][2] = (task scanned, twin task or -1): the reference's default
 *             orientations '+' and '+/-' of one motif (mepp.py:67-90) share filter 0 and its threshold
 *             (motif_layers.py:36-76), so one pass over the sequences with the two-filter table of the
 *             '+/-' task also yields the '+' task (x = relu of filter 0 alone).  Every task must appear
 *             exactly once, as a scanned task or as a twin; a twin's own tables are not read.
 *             NULL: every task is scanned on its own.
 *   entries_dev   optional sparse form of X for mepp_profile_sparse: {int32 row, int32 seq, float x, pad}
 *                 records, one per non-zero of the blurred score matrix, in MEPP_ENTRY_LISTS sub-lists of
 *                 entry_cap / MEPP_ENTRY_LISTS records each (sub-list = k-block % MEPP_ENTRY_LISTS);
 *                 entry_count_dev uint64[16 * MEPP_ENTRY_LISTS]: record count of sub-list j at [16 j] (a count
 *                 may exceed the sub-list capacity: then use the tiled path),
 *                 row_nnz_dev int32 [T*L + 1] non-zeros per row.  x_planes_dev may be NULL when given.
 *   max_motif_len  upper bound of mlen over the scanned tasks (sizes the shared-memory pair tables;
 *             a longer task traps); 0 = MEPP_MAX_MOTIF_LEN
 *   x_planes_dev  A operand (rows t*L+l).  Must be ALL-ZERO on entry (tiles and flag words): X is ~99.9 % zeros,
 *                 so the scan stores only the 16-byte pieces that hold a non-zero value and sets the flag bit
 *                 of their 128 x 16 block.  mepp_x_planes_reset restores the all-zero state after the GEMM.
 *   rowstats_dev  double [T*L][3]: sum_n x, sum_n x^2, sum_n x*zhat  (x = blurred score); zeroed by the call
 *   count_dev     int32 [T*L]  : #sequences with an unsmoothed match at each position; zeroed by the call
 *   density_dev   int32 [T][n_pad]: #matches per sequence
 *   hits_dev      optional: {int32 task, int32 seq, int32 pos, float x0} records, capacity hits_cap,
 *                 hit_count_dev int64[1] total number of matches (may exceed cap); NULL to skip
 *   xscale_dev    optional float [T]: per-task power-of-two scale of the fp16 hi + lo pieces of x_planes_dev
 *                 (mepp_x_scales; pass the same array to mepp_profile_corr_gemm); NULL: MEPP_X_SCALE for every task */
int mepp_x_scales(const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev, const int32_t* nfilt_dev, int T,
                  float* xscale_dev, void* stream);
int mepp_scan(const uint32_t* sym_T_dev, const float* zhat_dev,
              int64_t N, int L, int T, int margin,
              const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev, const int32_t* nfilt_dev,
              const uint32_t* qtab_dev, const uint32_t* qthr_dev, const int32_t* units_dev, int n_units,
              int max_motif_len,
              void* x_planes_dev, double* rowstats_dev, int32_t* count_dev, int32_t* density_dev,
              void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
              void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev, int32_t* row_nnz_dev,
              const float* xscale_dev, void* stream);

/* ---- bit-parallel first stage of the scan (same reference code as mepp_scan: motif_layers.py:80-165) ----
 * Matches are rare, so most windows can be discarded by counting "bad" letters: with a cut d0 on the per-tap
 * penalty max_c W[c][j] - W[c][j], a window with K or more letters above the cut (at taps with one or two good
 * letters) cannot reach the threshold (K, d0 chosen per task so that this is certain; N bases never count).  One
 * lane = one sequence, one register word = 32 window starts, a 2-bit saturating counter per filter updated with
 * logic ops on bit planes of the sequences.  The survivors are evaluated exactly as in mepp_scan, so the outputs
 * are identical.
 *   mepp_plane_words(L)   words per sequence and letter of the plane stream (word 0 and the tail are zero padding)
 *   mepp_pack_planes      sym_T (mepp_pack_dna) -> planes_T uint32 [mepp_plane_words(L)][4][n_pad]:
 *                         bit k of word w >= 1 of plane c = [base 32 (w - 1) + k is a letter other than c];
 *                         N bases, positions past L and padding sequences are zero in every plane
 *   mepp_scan_bitpar_tables   per task: counted taps, K and whether the task's unit takes the bit-parallel kernel
 *                         (modelled cost <= cost_ratio x that of the pair tables: 0 = only units that can never
 *                         match, 1 = the model's break-even, huge = every unit with a sound choice);
 *                         bp_dev uint32 [T][MEPP_BP_WORDS]; needs the final bias (threshold) of every task
 *   mepp_scan_bitpar      mepp_scan with the first stage chosen per unit; bp_scratch_dev int32 [4 * U + 2],
 *                         U = n_units (or T without a unit list) */
int64_t mepp_plane_words(int L);
int mepp_pack_planes(const uint32_t* sym_T_dev, int64_t N, int L, uint32_t* planes_T_dev, void* stream);
int mepp_scan_bitpar_tables(const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev,
                            const int32_t* nfilt_dev, int T, double cost_ratio, uint32_t* bp_dev, void* stream);
int mepp_scan_bitpar(const uint32_t* sym_T_dev, const float* zhat_dev,
                     int64_t N, int L, int T, int margin,
                     const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev, const int32_t* nfilt_dev,
                     const uint32_t* qtab_dev, const uint32_t* qthr_dev, const int32_t* units_dev, int n_units,
                     int max_motif_len,
                     void* x_planes_dev, double* rowstats_dev, int32_t* count_dev, int32_t* density_dev,
                     void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                     void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev, int32_t* row_nnz_dev,
                     const uint32_t* planes_T_dev, const uint32_t* bp_dev, int32_t* bp_scratch_dev,
                     const float* xscale_dev, void* stream);

/* ---- lean scan for the sparse path (optional, MEPP_SCAN_LEAN=1; same reference code as mepp_scan) ----
 * The scan only finds and publishes matches (count, density, match list) and drops them into fixed-capacity buckets
 * per (task, k-block of 32 sequences); a second kernel sorts every bucket by (sequence, position) and builds the
 * entries of the blurred score matrix, row_nnz and rowstats from it (a row belongs to the first match inside its
 * window; float32 sum in position order, IEEE division: the values mepp_scan produces).  No staging rows in shared
 * memory, no per-chunk blur.  buckets_dev: mepp_scan_lean_bucket_bytes(T, N) bytes; bucket_count_dev int32
 * [T * n_pad / 32].  If a bucket overflows (more than 64 matches of a task in 32 sequences) entry_count_dev[1] is
 * non-zero and the group has to be redone with mepp_scan. */
int64_t mepp_scan_lean_bucket_bytes(int T, int64_t N);
int mepp_scan_lean(const uint32_t* sym_T_dev, const float* zhat_dev,
                   int64_t N, int L, int T, int margin,
                   const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev, const int32_t* nfilt_dev,
                   const uint32_t* qtab_dev, const uint32_t* qthr_dev, const int32_t* units_dev, int n_units,
                   int max_motif_len,
                   double* rowstats_dev, int32_t* count_dev, int32_t* density_dev,
                   void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                   void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev, int32_t* row_nnz_dev,
                   void* buckets_dev, int32_t* bucket_count_dev, void* stream);

/* ---- tensor-core first stage of the lean scan (default on the sparse path; MEPP_SCAN_TC=0 switches back) ----
 * Replaces the same reference code as mepp_scan (the Keras Conv1D of mepp/motif_layers.py:50-76, 80-165): the
 * convolution over the one-hot sequence runs as tcgen05.mma.kind::i8 products of the flat one-hot stream
 * (mepp_pack_onehot: 4 bytes per base, overlapping 16-byte-strided rows = windows four bases apart) with conservatively
 * quantised int8 weights (4 window phases x 2 strands x units as columns); only the sign of the int32 accumulators is
 * read back from TMEM.  The filter has no false negatives; its survivors are scored in the canonical float32 order and
 * published exactly like mepp_scan_lean's matches (same outputs, same bucket / entry / rowstats contract, same
 * overflow flag in entry_count_dev[1]).
 * onehot_dev: mepp_onehot_words(N, L) uint32.  units_host / mlen_host: host copies of the unit list (null = one unit per
 * task, like units_dev) and of the motif lengths of all T tasks (the N-tile plan is made on the host).
 * work_dev: mepp_scan_tc_work_bytes(N, L, units) bytes of scratch (weights, candidate lists).
 * entries_dev may be NULL (match-driven profile, mepp_profile_matches: only the row sums are built from the buckets).
 * dbg_acc_dev (optional, int32 [128][dbg_ld]): raw accumulators of the first 128 windows, for bring-up tests (such calls
 * run the two-issuer kernel tc_scan4_kernel, which carries the dump; the default is tc_scan4i_kernel). */
int64_t mepp_onehot_words(int64_t N, int L);
int mepp_pack_onehot(const uint8_t* ascii_dev, int64_t N, int L, uint32_t* onehot_dev, void* stream);
int64_t mepp_scan_tc_work_bytes(int64_t N, int L, int n_units);
int mepp_scan_tc(const uint32_t* sym_T_dev, const uint32_t* onehot_dev, const float* zhat_dev,
                 int64_t N, int L, int T, int margin,
                 const float* lut_dev, const float* bias_dev, const int32_t* mlen_dev, const int32_t* nfilt_dev,
                 const int32_t* units_dev, const int32_t* units_host, int n_units, const int32_t* mlen_host,
                 double* rowstats_dev, int32_t* count_dev, int32_t* density_dev,
                 void* hits_dev, int64_t hits_cap, unsigned long long* hit_count_dev,
                 void* entries_dev, int64_t entry_cap, unsigned long long* entry_count_dev, int32_t* row_nnz_dev,
                 void* buckets_dev, int32_t* bucket_count_dev, void* work_dev, int64_t work_bytes,
                 int32_t* dbg_acc_dev, int dbg_ld, void* stream);

/* ------------------------------------------------------------------ heat-map reduction
 * Replaces mepp/plot.py:57-66 (clip at mean + sigma * std of the non-zero scores) and plot.py:32-55,106-145 (block
 * maximum down to the figure's pixel grid) on the dense (N, L) score matrix that core.py:568-573 returns: computed
 * from the match list {task, seq, pos, x0} (hits_dev / hit_count_dev as mepp_scan leaves them), the dense matrix never
 * exists.  out_dev: float [T][H][W], H = ceil(N / block_rows), W = ceil(L / block_cols) (the blocks of
 * get_aspect_tensor, plot.py:32-40); sigma < 0: no clip.  work_dev: mepp_heatmap_work_bytes(T) bytes. */
int64_t mepp_heatmap_work_bytes(int T);
int mepp_heatmap(const void* hits_dev, const unsigned long long* hit_count_dev, int64_t hits_cap, int T, int64_t N, int L,
                 int block_rows, int block_cols, float sigma, float* out_dev, void* work_dev, void* stream);

/* Measurement hook (bench.py's roofline block): while enabled, every tensor-core scan launch of mepp_scan_tc is bracketed
 * by CUDA events on its stream.  A call returns what was gathered since the previous call -- summed kernel time (it
 * waits for the recorded events), int8 multiply-adds issued (128 x columns x 32 per K-step and 512-base tile, padding
 * included), bytes of the one-hot stream read, launches -- clears it and sets the new state.  Outputs may be NULL. */
int mepp_scan_tc_timing(int enable, double* ms_out, double* macs_out, double* stream_bytes_out, int* launches_out);

/* Restore the all-zero state of an A operand that mepp_scan(..., N, L, T, margin, ...) wrote, from the
 * match list of that call (every match dirties rows pos-margin..pos+margin of one 8-column group); a NULL
 * or overflowed list clears the whole operand instead.  Call after the operand's last consumer. */
int mepp_x_planes_reset(void* x_planes_dev, const void* hits_dev, int64_t hits_cap,
                        const unsigned long long* hit_count_dev, int64_t N, int L, int T, int margin,
                        void* stream);

/* ------------------------------------------------------------------ K2s: sparse profile
 * r = correlation(S = X . Y) like mepp_profile_corr_gemm, for X given as the scan's entry list: the
 * records are binned by row and every row of S is gathered from the rows of Y (float32, row-major
 * [n_pad][mepp_y_rows_ld(C)], built once per dataset by mepp_build_y_rows), float64 accumulation.
 * Cost is proportional to the number of non-zeros of X (4 KB of L2 traffic each at P = 1000); the
 * tensor-core GEMM below costs the same whatever X holds.  Nothing is written if the entry list
 * overflowed (*entry_count_dev > entry_cap): the caller then takes the tiled path.
 * work_dev: mepp_profile_sparse_work_bytes(m_rows, entry_cap) bytes, beginning with int32 offsets[m_rows + 2]
 * (CSR row offsets; offsets[m_rows] = number of non-zeros, offsets[m_rows + 1] = 1 if the list overflowed);
 * row_nnz_dev is consumed. */
/* Column statistics of the exact float32 score operand (what the sparse profile path multiplies: mepp_build_y_rows):
 * ysum_dev[2] = sum and sum of squares of yhat (float64; the same for every permuted column), col_syz_dev[c] =
 * sum_n y_c[n] * zhat[n] (zeros when zhat_dev is NULL).  mepp_build_y computes the same from the fp16 hi + lo split the
 * tiled operand of the tensor path holds. */
int mepp_y_stats(const float* yhat_dev, const int32_t* perm_dev, const float* zhat_dev, int64_t N, int P,
                 double* col_syz_dev, double* ysum_dev, void* stream);
/* Chunked form of mepp_build_y_rows + mepp_y_stats, for permutation matrices that are staged a few hundred rows at a
 * time (cfg 5: 1e4 x 1e6 indices = 40 GB): columns col0 .. col0 + n_cols - 1 of y_rows / col_syz from perm_chunk_dev
 * (n_cols x N; NULL with col0 = 0, n_cols = 1: column 0 = the true scores).  The call with col0 = 0 also clears y_rows
 * and fills ysum, so it comes first; C = 1 + P of the whole operand. */
int mepp_build_y_cols(const float* yhat_dev, const int32_t* perm_chunk_dev, const float* zhat_dev, int64_t N, int col0,
                      int n_cols, int C, float* y_rows_dev, double* col_syz_dev, double* ysum_dev, void* stream);
int64_t mepp_y_rows_ld(int C);
int mepp_build_y_rows(const float* yhat_dev, const int32_t* perm_dev, int64_t N, int P, float* y_rows_dev,
                      void* stream);
/* Match-driven form of the sparse profile (default): the blur commutes with the contraction, so r is built from ONE
 * gathered row of Y per match -- hits_dev / hit_count_dev / count_dev as mepp_scan leaves them, rowstats_dev the
 * float32-blur row sums -- instead of one per non-zero of the blurred matrix (2 margin + 1 times fewer bytes).  Two
 * passes: per-position products S0 (float64, m_rows x mepp_y_rows_ld(C), inside work_dev), then their window sums
 * and the correlation epilogue.  A match enters as float32(x0 / k), the blurred value of a window it is alone in, so the
 * products are the ones mepp_profile_sparse sums; only windows holding two matches of one sequence differ (by the
 * rounding of one float32 addition).  n_pad_rows = rows of y_rows_dev (mepp_pad32(N)): an operand beyond the L2 is
 * gathered in 128-column slabs, slab by slab.  A match list longer than hits_cap leaves r untouched (the caller sees
 * hit_count > hits_cap and redoes the group). */
int64_t mepp_profile_matches_work_bytes(int64_t m_rows, int64_t hits_cap, int C);
int mepp_profile_matches(const void* hits_dev, const unsigned long long* hit_count_dev, int64_t hits_cap,
                         const int32_t* count_dev, int T, int L, int margin, const float* y_rows_dev, int64_t n_pad_rows,
                         const double* rowstats_dev, const double* colconst_dev, float* r_dev, int64_t ldr, int C,
                         void* work_dev, void* stream);
int64_t mepp_profile_sparse_work_bytes(int64_t m_rows, int64_t entry_cap);
int mepp_profile_sparse(const void* entries_dev, const unsigned long long* entry_count_dev, int64_t entry_cap,
                        int32_t* row_nnz_dev, const float* y_rows_dev, const double* rowstats_dev,
                        const double* colconst_dev, float* r_dev, int64_t ldr, int64_t m_rows, int C, void* work_dev,
                        void* stream);

/* ------------------------------------------------------------------ K2: profile GEMM
 * S[r][c] = sum_n X[r][n] * Y[n][c] on tcgen05 tensor cores (fp16 hi/lo split, 3 products,
 * fp32 TMEM accumulators); replaces the (B,L,1+P) x*y reduce_sum of core.py:250.
 * S_dev: float [m_rows][ldS], ldS = n_tiles_n*BN. */
int mepp_profile_gemm(const void* x_planes_dev, const void* y_planes_dev, float* S_dev,
                      int64_t m_rows, int C, int64_t n_pad, void* stream);
/* Instrumentation: out2 = {K=16 MMA steps issued, K=16 steps total} accumulated over all
 * mepp_profile_gemm launches of this process since the last reset (all-zero steps of X are skipped;
 * synchronises the device). */
int mepp_gemm_counters(unsigned long long* out2, int reset);
/* CUDA-core cross-check of the same contraction from the same planes (float64 accumulate). */
int mepp_profile_gemm_check(const void* x_planes_dev, const void* y_planes_dev, double* S_dev,
                            int64_t m_rows, int C, int64_t n_pad,
                            const int32_t* rows_dev, int n_rows_sel, void* stream);

/* ------------------------------------------------------------------ correlation epilogue
 * core.py:182-267 (Pearson / partial correlation) + np.nan_to_num (core.py:602).
 * r_dev: float [m_rows][ldr] (ldr >= C).  use_z: 1 = partial correlation controlling z. */
int mepp_correlation(const float* S_dev, int64_t ldS, const double* rowstats_dev,
                     const double* col_syz_dev, const double* ysum_dev, const double* zsum_host2,
                     int64_t N, int64_t m_rows, int C, int use_z, float* r_dev, int64_t ldr, void* stream);
/* Fused form (what the engine runs): per-dataset column constants once, then the GEMM writes r directly
 * (the same formulas, evaluated in float64 by the GEMM's epilogue warps; S never reaches HBM).
 * colconst_dev: double [8 + 2*C].  r_dev: float [m_rows][ldr], ldr = n_tiles_n*BN. */
int mepp_col_constants(const double* col_syz_dev, const double* ysum_dev, const double* zsum_host2,
                       int64_t N, int C, int use_z, double* colconst_dev, void* stream);
/* Same, with {sum zhat, sum zhat^2} read from device memory (no host synchronisation). */
int mepp_col_constants_dev(const double* col_syz_dev, const double* ysum_dev, const double* zsum_dev, int64_t N, int C,
                           int use_z, double* colconst_dev, void* stream);
/* xscale_dev / L: the per-task scales the scan wrote x_planes_dev with (float [m_rows / L]) or NULL. */
int mepp_profile_corr_gemm(const void* x_planes_dev, const void* y_planes_dev, const double* rowstats_dev,
                           const double* colconst_dev, float* r_dev, int64_t m_rows, int C, int64_t n_pad,
                           const float* xscale_dev, int L, void* stream);

/* ------------------------------------------------------------------ K3: permutation statistics
 * core.py:269-376 for T tasks of L positions and C = 1+P columns (r_dev: float [T*L][ldr]).
 *   k_lo, k_hi: floor of the numpy "linear" virtual index for the two percentiles over P values
 *   ks_lo, ks_hi: same over L*P values (the pooled "semi" null)
 * Outputs (per task):
 *   full_os   float [T][L][4]: order statistics a[k_lo], a[k_lo+1], a[k_hi], a[k_hi+1] of r[l,1:]
 *   full_cnt  int32 [T][L]   : #{p : |r[l,p]| >= |r[l,0]|}
 *   semi_os   float [T][4]   : pooled order statistics
 *   semi_cnt  int64 [T][L]   : #{(l',p) : |r[l',p]| >= |r[l,0]|}
 *   integral  double [T][C]  : trapezoid over positions of |r[:,c]|
 * work_dev: scratch of mepp_perm_stats_work_bytes(T, L, C) bytes. */
int64_t mepp_perm_stats_work_bytes(int T, int L, int C);
int mepp_perm_stats(const float* r_dev, int64_t ldr, int T, int L, int C,
                    int k_lo, int k_hi, int64_t ks_lo, int64_t ks_hi,
                    float* full_os_dev, int32_t* full_cnt_dev, float* semi_os_dev,
                    long long* semi_cnt_dev, double* integral_dev, void* work_dev, void* stream);

/* ------------------------------------------------------------------ K4: output columns
 * The per-position / per-task columns of core.get_positional_r_df (core.py:286-374), the parametric
 * Fisher-z interval and t-test (core.py:380-413, float32 arithmetic) and the smoothed counts
 * (core.py:438-454) of one task group, from the outputs of mepp_perm_stats / mepp_scan.  g_* are the
 * interpolation weights of numpy's 'linear' percentile at k_* (per position over P values, pooled over
 * L*P values); z_margin = float32(norm.ppf(1 - alpha/2) * sqrt(1 / (n_seq - 3))).
 * out_dev (mepp_finalize_out_bytes): float64 [full_lower, full_upper, full_pval, semi_pval, pval, count,
 * smoothed_count][T*L], float32 [positional_r, lower, upper][T*L], then (8-byte aligned) float64
 * [semi_lower, semi_upper, integral_lower, integral_upper, integral_pval][T], float32 [integral_r][T].
 * With C == 1 (no permutations) the permutation columns are left untouched. */
int64_t mepp_finalize_out_bytes(int T, int L);
int mepp_finalize_positions(const float* r_dev, int64_t ldr, int T, int L, int C, const float* full_os_dev,
                            const int32_t* full_cnt_dev, const float* semi_os_dev, const long long* semi_cnt_dev,
                            const double* integral_dev, const int32_t* count_dev, int64_t n_seq, int margin,
                            int k_lo, int k_hi, double g_lo, double g_hi, double gs_lo, double gs_hi, float z_margin,
                            void* out_dev, void* stream);

/* density [n] int32 (matches per (task, sequence), mepp_scan) -> uint16 (saturating; a count is <= L) for the copy to
 * the host: what core.get_motif_density_by_rank returns per task (core.py:456-470) needs 2 bytes per sequence.
 * n must be a multiple of 4 (rows are padded to 32 sequences). */
int mepp_density_u16(const int32_t* density_dev, int64_t n, uint16_t* out_dev, void* stream);

/* ------------------------------------------------------------------ host: output tables
 * Tab-separated table of numeric columns, byte-identical to pandas.DataFrame.to_csv(path, sep='\t', index=False)
 * (how mepp/single.py:223-241 writes positions_df.tsv / ranks_df.tsv): header_line = the tab-joined column names,
 * col_kind[c]: 0 float64, 1 float32, 2 int64, 3 int32, 4 uint16 (as an integer), 5 uint16 printed as float64;
 * col_ptr[c] / col_stride_bytes[c]: first element and byte stride of column c (a stride of 0 repeats one value, for
 * the columns the reference broadcasts over all rows).  Floats are written as numpy's str(): shortest digits that
 * round-trip in the column's precision, positional for 1e-4 <= |x| < 1e16, NaN as the empty string.  Host only. */
int mepp_write_tsv(const char* path, const char* header_line, int n_cols, const int32_t* col_kind,
                   const void* const* col_ptr, const int64_t* col_stride_bytes, int64_t n_rows);

#ifdef __cplusplus
}
#endif
#endif /* MEPP_B200_H */
