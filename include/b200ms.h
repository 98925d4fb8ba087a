/*
 * b200ms.h -- C-ABI of libb200ms.so: the B200 (sm_100a) implementation of Morphik's ColPali
 * late-interaction (MaxSim) scoring path.  Plain pointers and sizes only; no C++/torch types.
 *
 * Drop-in boundary (reference paths relative to /root/reference):
 *   The reference has no FFI on this path; its scorers are reached through the Python plugin
 *   interface core/vector_store/base_vector_store.py:7-65 (BaseVectorStore).  The host mirror of that
 *   interface lives in morphik-core_b200/store.py and calls this library through ctypes
 *   (INTEGRATION.md shows the binding).  Each entry point below names what it replaces:
 *
 *   b200ms_sign_pack       <- morphik_rust binary_quantize_batch_packed  (morphik_rust/src/binary_ops.rs:147-222,
 *                             core/utils/fast_ops.py:191-227, called from multi_vector_store.py:329-345)
 *   b200ms_pack_pages      <- store_embeddings' per-page quantise + layout (multi_vector_store.py:681-703;
 *                             fast_multivector_store.py:673-707 writes fp32 .npy; values are bf16-exact :674)
 *   b200ms_set_corpus      <- the table scan source: multi_vector_embeddings rows (multi_vector_store.py:240-251)
 *   b200ms_pack_queries    <- _binary_quantize(query) (multi_vector_store.py:731) / torch.from_numpy(q).float()
 *                             (fast_multivector_store.py:554)
 *   b200ms_score           <- SQL public.max_sim over every row (multi_vector_store.py:287-311,746-763) for
 *                             B200MS_B1, and colpali_engine score_multi_vector (fast_multivector_store.py:553-555)
 *                             for B200MS_BF16; B200MS_I8 is new (BASELINE config 3)
 *   b200ms_topk            <- ORDER BY similarity DESC LIMIT k (multi_vector_store.py:759) / torch.topk
 *                             (fast_multivector_store.py:556), plus the WHERE document_id IN (...) filter (:752-757)
 *   b200ms_merge_topk      <- (new) merge of per-GPU-shard top-k lists after the NCCL all-gather
 *   b200ms_search_host     <- MultiVectorStore.query_similar's scoring section with HOST query buffers
 *                             (multi_vector_store.py:721-763): quantise/convert, scan, top-k, results to host
 *   b200ms_search_device   <- same with device-resident inputs/outputs, asynchronous on the caller's stream
 *
 * Conventions: every function returns 0 on success or a negative B200MS_E* code; b200ms_last_error()
 * returns a message for the last failure on that handle (or the last global failure for NULL).
 * A handle is bound to one CUDA device and is NOT thread-safe: the caller serialises calls per handle
 * (the Python store holds a lock and calls through asyncio.to_thread, SURVEY 8b "Threading").
 * Device buffers passed in are caller-owned (torch tensors); the library only owns small scratch.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with B200MS_ECUDA.
 */
#ifndef B200MS_H_
#define B200MS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MS_VERSION 200 /* 0.2.0: + multi-GPU entry points (b200ms_comm_*, b200ms_allgather_topk, b200ms_sharded_search_*),
                              b200ms_rerank_batch_device, b200ms_fde_configure_ex, B200MS_F8 corpora, option "zero_pad_compat";
                              options "a_in_tmem" / "split4" / "epi_w4" removed with the kernels they selected */

/* element types */
#define B200MS_F32 0  /* float32 source rows (ingest / query side only)            */
#define B200MS_BF16 1 /* bfloat16 rows, 256 B per 128-d patch vector                */
#define B200MS_I8 2   /* int8 rows, 128 B per patch vector (global symmetric scale) */
#define B200MS_B1 3   /* sign bits, MSB-first, 16 B per patch vector                */
#define B200MS_I32 4  /* int32 scores (I8 / B1 corpora)                             */
#define B200MS_F8 5   /* fp8 e4m3 rows, 128 B per patch vector (global power-of-two scale), f32 scores */

/* error codes */
#define B200MS_OK 0
#define B200MS_EINVAL (-1) /* bad argument                                  */
#define B200MS_ECUDA (-2)  /* CUDA runtime/driver failure or no device       */
#define B200MS_ESTATE (-3) /* call order violated (e.g. score before corpus) */
#define B200MS_ENOMEM (-4) /* scratch allocation failed                      */

#define B200MS_DIM 128        /* embedding dimension of ColPali / ColQwen patch vectors */
#define B200MS_ROW_GROUP 32   /* pages and queries are padded to multiples of 32 rows   */
#define B200MS_MAX_K 4096     /* largest k accepted by b200ms_topk / b200ms_merge_topk  */

typedef struct b200ms b200ms_t;

int b200ms_version(void);
/* Number of CUDA devices visible (0 without a driver); never fails. */
int b200ms_device_count(void);
int b200ms_create(int device, b200ms_t** out);
int b200ms_destroy(b200ms_t* h);
const char* b200ms_last_error(const b200ms_t* h);

/* ---- layout helpers (host, pure arithmetic) ------------------------------------------------- */
/* Rows a page of `len` true rows occupies in the padded corpus layout: roundup(len, 32); 0 stays 0. */
int64_t b200ms_padded_len(int64_t len);
/* Sum of b200ms_padded_len over page_lens[0..n_pages). */
int64_t b200ms_padded_rows(const int32_t* page_lens, int64_t n_pages);
/* Bytes per row for a corpus dtype (BF16 256 / I8 and F8 128 / B1 16), or 0. */
int64_t b200ms_row_bytes(int dtype);
/* Number of 32-row query groups for queries of the given lengths: sum ceil(len/32) (a 0-length query -> 0). */
int64_t b200ms_query_groups(const int32_t* q_lens, int n_q);

/* ---- quantise / layout (device) ------------------------------------------------------------- */
/* x: device [rows,128] F32|BF16 -> out: device [rows,16] sign bits, bit = (float32(x) > 0), MSB first. */
int b200ms_sign_pack(b200ms_t* h, const void* x, int src_dtype, int64_t rows, uint8_t* out, void* stream);

/* out[i] = popcount(q_bits XOR cand_bits[i]) for one packed 16-byte query row against n packed candidate rows (device,
 * 16-byte aligned): fast_ops.hamming_distance_batch (core/utils/fast_ops.py:242-248, binary_ops.rs:266-292). */
int b200ms_hamming_batch(b200ms_t* h, const uint8_t* q_bits, const uint8_t* cand_bits, int64_t n, uint32_t* out, void* stream);

/* Convert n_pages pages stored back to back at src (device, [sum len,128] F32|BF16) into the padded
 * corpus layout at dst (device): page i occupies padded_len(len_i) rows, the padding rows repeat the
 * page's last true row (max-/min-invariant, so scores are unchanged).  dst_dtype BF16: round-to-nearest-even;
 * I8: rint(x * i8_scale) clamped to [-127,127]; F8: e4m3(x * i8_scale), round-to-nearest-even, saturating (use a power-of-two
 * scale, 64 for unit-norm rows); B1: sign bits.  page_lens is a HOST array. */
int b200ms_pack_pages(b200ms_t* h, const void* src, int src_dtype, const int32_t* page_lens, int64_t n_pages,
                      void* dst, int dst_dtype, float i8_scale, void* stream);

/* Attach a packed corpus (device memory, caller-owned, must stay valid until the next set_corpus/destroy).
 * rows: [b200ms_padded_rows(page_lens), 128] of dtype BF16|I8|F8|B1, 1024-byte aligned.  page_lens (HOST) are the
 * TRUE lengths.  Builds the chunk->page table, the work-unit plan and the TMA descriptor. */
int b200ms_set_corpus(b200ms_t* h, const void* rows, int dtype, const int32_t* page_lens, int64_t n_pages);
int64_t b200ms_corpus_pages(const b200ms_t* h);
int64_t b200ms_corpus_rows(const b200ms_t* h);

/* Pack n_q queries stored back to back at q (device, [sum len,128] F32|BF16) into 32-row groups of the
 * corpus dtype at q_packed (device, capacity >= roundup(groups,4)*32 rows; tail rows are zeroed).
 * group_offsets_out (HOST, [n_q+1]) receives the group range of every query.  Returns group count in
 * *n_groups_out. */
int b200ms_pack_queries(b200ms_t* h, const void* q, int src_dtype, const int32_t* q_lens, int n_q, void* q_packed,
                        int dst_dtype, float i8_scale, int32_t* group_offsets_out, int* n_groups_out, void* stream);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* group_scores[g, p] = sum over the (<=32) tokens of group g of max over the rows of page p of <q_t, d_r>
 *   BF16 / F8 corpus: float32 (fp32 accumulate on tcgen05);  I8: int32 exact;  B1: int32 sum_t max_r (128 - hamming).
 * q_packed: device [roundup(n_groups,4)*32, 128] of the corpus dtype; q_lens/group_offsets (HOST) as produced by
 * b200ms_pack_queries (B1 needs the true token counts; may be NULL for BF16/I8).
 * group_scores: device [n_groups_padded, ld] with ld >= n_pages, n_groups_padded = roundup(n_groups,4).
 * Pages with zero rows score 0 (COALESCE(...,0.0), multi_vector_store.py:308). */
int b200ms_score(b200ms_t* h, const void* q_packed, int n_groups, const int32_t* q_lens, const int32_t* group_offsets,
                 int n_q, void* group_scores, int64_t ld, void* stream);

/* Per query: score[p] = scale * sum_{g in [group_offsets[q], group_offsets[q+1])} group_scores[g, p]; keep the k
 * best pages whose bit is set in allow_mask (device, bit p&31 of word p>>5; NULL = all), ordered by score DESC
 * then page id ASC (the reference leaves ties unspecified).  Outputs are device arrays: top_scores/top_ids
 * [n_q, k] (unused slots: -inf / -1), top_counts [n_q].  id_base is added to every page id (global ids of a shard). */
int b200ms_topk(b200ms_t* h, const void* group_scores, int score_dtype, int64_t n_pages, int64_t ld,
                const int32_t* group_offsets, int n_q, const uint32_t* allow_mask, int k, float scale, int64_t id_base,
                float* top_scores, int64_t* top_ids, int32_t* top_counts, void* stream);

/* Merge candidate lists (e.g. the all-gathered per-shard top-k): cand_scores/cand_ids device [n_q, m]
 * (entries with id < 0 are ignored) -> best k by (score DESC, id ASC).  m <= 2*B200MS_MAX_K. */
int b200ms_merge_topk(b200ms_t* h, const float* cand_scores, const int64_t* cand_ids, int n_q, int m, int k,
                      float* top_scores, int64_t* top_ids, int32_t* top_counts, void* stream);

/* Whole query step with HOST buffers (pageable or pinned): q_host [sum q_lens,128] float32.  Copies the
 * queries to the device, packs, scores the attached corpus, selects top-k and copies results back;
 * synchronous.  allow_mask_host may be NULL.  score scale: BF16 1.0; B1 1/128; I8 1/(q_scale*corpus_scale). */
int b200ms_search_host(b200ms_t* h, const float* q_host, const int32_t* q_lens, int n_q, int k,
                       const uint32_t* allow_mask_host, float i8_q_scale, float score_scale, int64_t id_base,
                       float* top_scores_host, int64_t* top_ids_host, int32_t* top_counts_host);

/* b200ms_search_host with ONE allow-mask PER QUERY: allow_masks_host is a [n_masks, ceil(n_pages/32)] matrix of page
 * bitmasks, mask_index_host[i] the row that filters query i (-1 = unfiltered); n_masks = 0 means no filtering at all.
 * The production caller always passes the requesting user's authorised doc_ids (core/services/document_service.py:408-417),
 * so concurrent query_similar calls carry DIFFERENT filters; with per-query masks they still share one pass over the
 * corpus (four 32-token queries fit one 128-row MMA tile and cost the same as one). */
int b200ms_search_host_masked(b200ms_t* h, const float* q_host, const int32_t* q_lens, int n_q, int k,
                              const uint32_t* allow_masks_host, int n_masks, const int32_t* mask_index_host,
                              float i8_q_scale, float score_scale, int64_t id_base, float* top_scores_host,
                              int64_t* top_ids_host, int32_t* top_counts_host);

/* Same with everything on the device and no synchronisation (enqueued on `stream`).  q_dev: F32|BF16. */
int b200ms_search_device(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q, int k,
                         const uint32_t* allow_mask_dev, float i8_q_scale, float score_scale, int64_t id_base,
                         float* top_scores_dev, int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream);

/* b200ms_search_device with one allow-mask per query: allow_masks_dev [n_masks, ceil(n_pages/32)] and mask_index_dev
 * [n_q] (-1 = unfiltered) are DEVICE arrays (see b200ms_search_host_masked). */
int b200ms_search_device_masked(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q, int k,
                                const uint32_t* allow_masks_dev, int n_masks, const int32_t* mask_index_dev,
                                float i8_q_scale, float score_scale, int64_t id_base, float* top_scores_dev,
                                int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream);

/* MaxSim over an explicit candidate list (the "rerank" half of a two-stage search, fast_multivector_store.py:545-557):
 * cand_ids_dev = device array of n_cand page ids (-1 = unused slot).  Only those pages are scanned; ties break towards
 * the earlier candidate slot (= better first-stage rank).  Outputs as b200ms_search_device. */
int b200ms_rerank_device(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q,
                         const int64_t* cand_ids_dev, int n_cand, int k, float i8_q_scale, float score_scale,
                         float* top_scores_dev, int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream);

/* b200ms_rerank_device with one candidate list PER QUERY (the two-stage search of a query batch: every query reranks its own
 * first-stage candidates, fast_multivector_store.py:526-557): cand_ids_dev = device [n_q, n_cand] page ids (-1 = unused slot).
 * One launch per 128-token query tile scores that tile's queries against their own lists only; outputs as above. */
int b200ms_rerank_batch_device(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q,
                               const int64_t* cand_ids_dev, int n_cand, int k, float i8_q_scale, float score_scale,
                               float* top_scores_dev, int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream);

/* ---- multi-GPU: document shards, one NCCL all-gather of top-k lists (new relative to the reference) -------------
 * One process per GPU; every rank owns a handle over its shard.  NCCL is bound at run time (dlopen libnccl.so.2, or the
 * path in B200MS_NCCL_LIB); b200ms_comm_available() says whether that worked.
 *   rank 0: b200ms_comm_unique_id(id) -> ship the 128 bytes to every rank by any means (file, socket, MPI, torch) ->
 *   every rank: b200ms_comm_init(h, id, rank, world)      (collective; creates the handle's own communicator)
 *   or b200ms_comm_adopt(h, ncclComm_t, rank, world) to use a communicator the host already has.
 * Exchange layout of one rank's list ("xchg"): [n_q*k int64 global page ids][n_q*k float32 scores] padded to a multiple of
 * 16 bytes = b200ms_xchg_bytes(n_q,k) bytes (so that every rank's block of the gathered buffer is 8-byte aligned), unused
 * entries id -1 / score -inf -- b200ms_search_device can write it in place (top_ids_dev = xchg,
 * top_scores_dev = xchg + n_q*k*8). */
int b200ms_comm_available(void);
int b200ms_comm_unique_id(uint8_t* id128);
int b200ms_comm_init(b200ms_t* h, const uint8_t* id128, int rank, int world);
int b200ms_comm_adopt(b200ms_t* h, void* nccl_comm, int rank, int world);
int b200ms_comm_destroy(b200ms_t* h);
int b200ms_comm_rank(const b200ms_t* h);
int b200ms_comm_world(const b200ms_t* h);
int64_t b200ms_xchg_bytes(int n_q, int k);
/* ncclBroadcast of a device buffer from `root` (e.g. the query rows of a batch) on `stream`. */
int b200ms_bcast_device(b200ms_t* h, void* buf, int64_t bytes, int root, void* stream);
/* ncclSend / ncclRecv of a device buffer to / from rank `peer` (e.g. ingest rows from the coordinating rank to the owner). */
int b200ms_send_device(b200ms_t* h, const void* buf, int64_t bytes, int peer, void* stream);
int b200ms_recv_device(b200ms_t* h, void* buf, int64_t bytes, int peer, void* stream);
/* The one collective of the path: all-gather every rank's xchg block (nccl_comm NULL = the handle's communicator) and merge
 * the world*k candidates per query -> identical top-k on every rank (score DESC, id ASC).  world * k <= 8192. */
int b200ms_allgather_topk(b200ms_t* h, void* nccl_comm, const void* xchg_local_dev, int n_q, int k, float* top_scores_dev,
                          int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream);
/* Pipelined sharded search.  _begin enqueues the local scan + top-k (global ids = id_base + local page) on `stream` and the
 * all-gather + merge on the handle's communication stream behind an event, then returns a ticket (>= 0; negative = error);
 * the caller's stream is NOT made to wait for the other ranks, so the next _begin may follow at once (two exchange slots: at
 * most two searches in flight).  _end makes `stream` wait for that ticket's merged result in the output buffers given to
 * _begin.  Masks as in b200ms_search_device_masked (n_masks = 0: unfiltered).  Works with world == 1 (no collective). */
int64_t b200ms_sharded_search_begin(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q, int k,
                                    const uint32_t* allow_masks_dev, int n_masks, const int32_t* mask_index_dev,
                                    float i8_q_scale, float score_scale, int64_t id_base, float* top_scores_dev,
                                    int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream);
int b200ms_sharded_search_end(b200ms_t* h, int64_t ticket, void* stream);
/* The same pipeline with HOST query rows (pinned or pageable float32) and host results: _host_begin copies the rows to the
 * device on the handle's stream and returns at once; _host_end blocks until that ticket's merged top-k has been copied back
 * and stores it in the caller's arrays.  Collect ticket t before beginning t+2. */
int64_t b200ms_sharded_search_host_begin(b200ms_t* h, const float* q_host, const int32_t* q_lens, int n_q, int k,
                                         const uint32_t* allow_masks_dev, int n_masks, const int32_t* mask_index_dev,
                                         float i8_q_scale, float score_scale, int64_t id_base);
int b200ms_sharded_search_host_end(b200ms_t* h, int64_t ticket, float* top_scores_host, int64_t* top_ids_host,
                                   int32_t* top_counts_host);

/* ---- fixed-dimensional encodings (MUVERA FDE): candidate generation in front of the scorer ---------------------------
 * Replaces the `fixed_dimensional_encoding` C++ extension (fast_multivector_store.py:325-331,447-449,521) and the
 * Turbopuffer ANN over its output (:526-532) with an exact encoder and an exhaustive cosine scan.
 * configure: reps repetitions x 2^ksim SimHash partitions x proj_dim AMS-sketch dims = fde_dim floats per item.
 *   simhash [reps,128,ksim] Gaussian, ams_index [reps,128] in [0,proj_dim), ams_sign [reps,128] = +-1 (HOST arrays).
 * encode: items stored back to back at rows (device F32|BF16, [sum len,128]); is_document 0: SUM per partition (query),
 *   1: AVERAGE per partition, empty partition -> 0 (document); out device float32 [n_items, fde_dim].
 * finalize: fp32 FDEs -> bf16 rows of the FDE corpus + 1/||row||.
 * scan: scores[q,p] = <q_fde[q], F[p]> * inv_norm[p]  (cosine ranking);  feed to b200ms_topk (score dtype F32,
 *   group_offsets = 0..n_q) to get the candidate list. */
int b200ms_fde_configure(b200ms_t* h, int reps, int ksim, int proj_dim, float scale, const float* simhash,
                         const int32_t* ams_index, const float* ams_sign);
/* configure + the upstream knobs the reference leaves at their defaults (fast_multivector_store.py:325-331 sets neither):
 *   fill_empty_partitions != 0: an empty document partition takes the projection of the point whose SimHash sign bits are
 *     nearest (Hamming distance, first minimum) instead of zeros;
 *   final_dim > 0: a final count sketch of the whole encoding, out[j] = sum_{i: final_index[i]==j} final_sign[i] * v[i]
 *     (final_index / final_sign: HOST arrays of reps*2^ksim*proj_dim entries; final_dim a multiple of 8) -- fde_dim becomes
 *     final_dim.  The matrices are INPUTS: a caller that holds the upstream extension's matrices passes them here and gets
 *     encodings interchangeable with vectors already stored by the reference. */
int b200ms_fde_configure_ex(b200ms_t* h, int reps, int ksim, int proj_dim, float scale, const float* simhash,
                            const int32_t* ams_index, const float* ams_sign, int fill_empty_partitions, int final_dim,
                            const int32_t* final_index, const float* final_sign);
int64_t b200ms_fde_dim(const b200ms_t* h);
int b200ms_fde_encode(b200ms_t* h, const void* rows, int src_dtype, const int32_t* item_lens, int64_t n_items,
                      int is_document, float* out, void* stream);
/* Document FDEs of pages [first_page, first_page + n_pages) of the attached BF16 corpus, computed from its packed rows
 * (rebuilds the FDE matrix of a corpus loaded from a shard file; at most 65535 pages per call). */
int b200ms_fde_encode_corpus(b200ms_t* h, int64_t first_page, int64_t n_pages, float* out, void* stream);
int b200ms_fde_finalize(b200ms_t* h, const float* fde, int64_t n, void* out_rows, float* inv_norm, void* stream);
int b200ms_fde_scan(b200ms_t* h, const void* fde_rows, const float* inv_norm, int64_t n_pages, const float* q_fde, int n_q,
                    float* scores, int64_t ld, void* stream);

/* ---- instrumentation ------------------------------------------------------------------------- */
/* Kernels launched by this handle since creation (bench.py's gpu_launches claim). */
int64_t b200ms_launch_count(const b200ms_t* h);
/* Device time (ms, CUDA events on the launching stream) of the last b200ms_score call's scoring kernels only;
 * synchronises.  Returns a negative error code on failure. */
float b200ms_last_score_ms(b200ms_t* h);
/* Number of scoring calls (b200ms_score or search) recorded so far, and the device times (ms) of the most recent n of
 * them (n <= 256, oldest first; synchronises; returns n or a negative error).  bench.py derives the roofline's
 * `achieved` from these: CUDA events on the launching stream, recorded inside the timed region. */
int64_t b200ms_score_call_count(const b200ms_t* h);
int b200ms_score_times_ms(b200ms_t* h, float* out_ms, int n);
/* Tuning knobs (0 keeps the default): rows per work unit, CTAs to launch. */
int b200ms_set_tuning(b200ms_t* h, int64_t unit_rows, int max_ctas);
/* Named options:
 *   "pair_cta"        CTA pairs with tcgen05 cta_group::2, M = N = 256: 0 = off, 1 = passes of >= 3 query tiles, 2 = also
 *                     exactly 2 tiles (default)
 *   "b1_tensor"       1 = score 1-bit corpora on tcgen05 with in-kernel bit expansion, 0 = POPC kernel, 2 = auto by group count
 *   "rowm"            1 (default) = full scans of <= 2 query groups (<= 64 query tokens; sign-bit corpora: 1 group) run on
 *                     maxsim_rowm_kernel: patch rows are the tcgen05 M operand, the query tokens N = 32 or 64, four lane-quadrant
 *                     epilogue warps in parallel; 0 = the query-as-M kernels (same results bit for bit)
 *   "zero_pad_compat" N > 0: reproduce colpali_engine score_multi_vector's zero-padding quirk with batch size N (128 upstream,
 *                     processing_colpali.py:350-362): a page shorter than the longest page of its batch scores
 *                     sum_t max(max_r <q_t,d_r>, 0).  Rerank calls batch the candidates in first-stage order (exactly the
 *                     reference, fast_multivector_store.py:553-555); full scans batch pages in page-id order.  0 = clean MaxSim
 *                     (default).  Float and int8 corpora only -- SQL max_sim has no padding.
 *   "zero_copy"       1 (default): host entry points move small calls (<= 256 query rows) through mapped pinned memory
 *   "host_graph"      1 (default): small unmasked host searches (pack -> score -> top-k) are captured into a CUDA graph on their
 *                     second use and replayed as one launch; 0 = three plain launches (same results)
 *   "fde_gemm"        1 (default): FDE scan on tcgen05 when fde_dim % 64 == 0, 0 = SIMT scan
 *   "unit_rows", "max_ctas"   work-unit size (needs a new b200ms_set_corpus) and launch width. */
int b200ms_set_option(b200ms_t* h, const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* B200MS_H_ */
