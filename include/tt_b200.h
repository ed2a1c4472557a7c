/*
 * tt_b200.h -- C ABI of the B200-native retrieval hot path (libtt_b200.so).
 *
 * One index = one row-major matrix of leaf embeddings resident in HBM plus four int32
 * relation arrays describing the hierarchical node tree.  The library replaces, for that
 * index, the numeric work behind the two objects the reference constructs at
 *   /root/reference/src/tensortruth/rag_engine.py:639-645   (and again :674-679)
 *       base         = index.as_retriever(similarity_top_k=k)      -> tt_scan_* + tt_rescore_topk
 *       am_retriever = AutoMergingRetriever(base, storage_context) -> tt_automerge
 * and, for a corpus row-sharded over several GPUs (not in the reference),
 * the k-way merge after the all-gather                              -> tt_merge_topk.
 *
 * Conventions
 *   - extern "C"; every entry point returns 0 (TT_OK) or a negative TT_ERR_* code and never
 *     throws; tt_last_error() gives the message for the calling thread.
 *   - All pointers are DEVICE pointers unless the name ends in _host.  The library never
 *     allocates or frees device memory: the caller (PyTorch) owns corpus, outputs, workspace.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous;
 *     nothing here synchronises.  The calls are CUDA-graph capturable.
 *   - Ordering rule everywhere: key descending, ties -> smaller id first.  ids are global row
 *     ordinals (id_base + local row) and must stay below 2^32.
 *   - There is no CPU implementation behind this ABI.
 */
#ifndef TT_B200_H
#define TT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_OK 0
#define TT_ERR_INVALID (-1)     /* bad argument */
#define TT_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define TT_ERR_UNSUPPORTED (-3) /* shape/alignment this build has no kernel for */
#define TT_ERR_WORKSPACE (-4)   /* workspace too small */
#define TT_ERR_TIMEOUT (-5)     /* a device-side wait ran out of time (tt_status_read says which) */

/*
 * Device-side status.  No kernel of this library traps or hangs: a wait that runs out of time -- an mbarrier of a
 * TMA/MMA ring (a protocol bug or a lost copy), a peer's exchange flag (a rank that died, raised or never made the
 * matching call) -- records a code and lets the launch finish on whatever it has, so the CUDA context, the other
 * indexes of the process and the peers' mappings of this GPU stay usable.  The results of such a launch are
 * garbage: a caller must look at the status before it trusts them (the Python binding does, and raises TTError
 * with TT_ERR_TIMEOUT on every rank concerned; errors past MultiIndexRetriever are logged and skipped per index
 * by the reference, rag_engine.py:453-455, and must never take the process down).
 *
 *   tt_status_configure(mapped_word_host, timeout_ms)   for the CURRENT device: `mapped_word_host` is an optional
 *       uint32 in pinned (device-mapped) HOST memory that the kernels OR the code into as well, so that the caller
 *       can poll it after its usual stream synchronisation without an extra device read; timeout_ms bounds one
 *       ring wait (exchange waits get four times that).  0 = keep the default (about 4 s).
 *   tt_status_read(out_host, clear)   synchronous read of the device-resident words (OR of all of them).
 */
#define TT_STATUS_RING_TIMEOUT 1u
#define TT_STATUS_EXCHANGE_TIMEOUT 2u /* bits 8..15: the source rank that was waited for */
/* cudaStreamSynchronize(stream) for hosts that hold nothing but the raw handle (the Python binding: one foreign call
 * that releases the interpreter lock, instead of an event record + an event wait through PyTorch). */
int tt_stream_synchronize(void* stream);
int tt_status_configure(uint32_t* mapped_word_host, int timeout_ms);
int tt_status_read(uint32_t* out_host, int clear);

#define TT_DTYPE_BF16 0
#define TT_DTYPE_F32 1

/* score_mode: what ChromaVectorStore would surface for the row. */
#define TT_SCORE_COSINE 0        /* key = score = f32(<q,c>/(|q||c|))                         */
#define TT_SCORE_CHROMA_L2_EXP 1 /* d = f32(|q|^2+|c|^2-2<q,c>); key = -d; score = f32(exp(-d)) */

/* scan variants */
#define TT_SCAN_AUTO 0
#define TT_SCAN_SIMT 1    /* CUDA-core fp32 streaming scan (any dim % 8 == 0)                  */
#define TT_SCAN_TCGEN05 2 /* TMA + tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), dim % 64 == 0   */

int tt_version(void);
const char* tt_last_error(void);

/* Number of shortlists (= persistent CTAs) a scan on `device` emits per query. */
int tt_scan_num_lists(int device);

/* Largest kprime (shortlist length per CTA) the scan kernels support. */
int tt_scan_max_kprime(void);

/*
 * Query preparation (replaces nothing in the reference -- the embedder hands over fp32):
 * q_hat = q/|q| in fp32, split into q_hi = bf16(q_hat), q_lo = bf16(q_hat - q_hi), so that two
 * bf16 MMA columns per query carry 16 mantissa bits of the query through the tensor cores.
 * q_hi/q_lo: bf16 [n_q, dim].
 */
int tt_prepare_queries(const float* q_f32, int n_q, int dim, void* q_hi_bf16, void* q_lo_bf16, void* stream);

/*
 * The same, and out_rho[q] = |q/|q| - q_hi|_2 rounded up: the part of the query a hi-only scan (q_lo not used) never
 * sees.  By Cauchy-Schwarz it bounds that scan's score error for EVERY row, so it replaces the worst-case 2^-8 the
 * hi-only certificate budgets for it (typically 2.5x smaller).
 */
int tt_prepare_queries_rho(const float* q_f32, int n_q, int dim, void* q_hi_bf16, void* q_lo_bf16, float* out_rho,
                           void* stream);

/*
 * Hands a hi-only scan's per-query slack max(0, eps_hi_only - rho[q]) to the certificate by lowering the query's
 * stage-1 thresholds (cand_thresh [n_q, n_lists], between stage 1 and stage 2): margin = s_k - max thresh grows by it,
 * and every consumer keeps comparing the margin with the nominal eps.  -inf / +inf thresholds keep their meaning.
 */
int tt_certificate_credit(float* cand_thresh, int n_q, int n_lists, const float* rho, float eps_hi_only, void* stream);

/*
 * Stage 1 -- dense scan + fused on-chip shortlist.  Replaces the vector-store query behind
 * `index.as_retriever(similarity_top_k=k).retrieve()` (rag_engine.py:639; ChromaVectorStore.query
 * -> collection.query(query_embeddings, n_results=k)).
 *
 * Streams corpus[n_rows, dim] (bf16, row stride row_stride_elems) exactly once per query tile,
 * scores every row approximately, a(r) = (<q_hi,c_r> + <q_lo,c_r>) * inv_norm[r] (fp32
 * accumulate; inv_norm may be NULL = 1), and keeps, per persistent CTA l and query b, the
 * kprime best rows it saw.  The score matrix never reaches HBM.
 *
 *   out_ids    int64 [n_q, n_lists*kprime]  global ids (id_base + row), -1 = empty slot
 *   out_approx float [n_q, n_lists*kprime]  approximate scores of those rows
 *   out_thresh float [n_q, n_lists]         every row CTA l saw and did NOT emit has
 *                                           a(r) <= out_thresh[b,l]  (-inf: l emitted all it saw)
 * q_lo_bf16 may be NULL (hi only: half the tensor work, wider certificate).
 * variant: TT_SCAN_AUTO | TT_SCAN_SIMT | TT_SCAN_TCGEN05.
 * ws: tt_scan_workspace_bytes() bytes of device memory, zeroed ONCE by the caller when it allocates
 *     them; the kernel leaves them zero.  Holds the dynamic tile scheduler's counters, so one
 *     workspace must not be shared by launches that can overlap.  NULL = static tile interleave.
 */
size_t tt_scan_workspace_bytes(void);
int tt_scan_topk_bf16(const void* corpus_bf16, int64_t n_rows, int dim, int64_t row_stride_elems,
                      const float* inv_norm, const void* q_hi_bf16, const void* q_lo_bf16, int n_q,
                      int kprime, int64_t id_base, int variant,
                      int64_t* out_ids, float* out_approx, float* out_thresh,
                      void* ws, size_t ws_bytes, void* stream);

/*
 * Stage 1 over a SEGMENTED corpus: several indexes concatenated row-wise in one matrix, segment s = rows
 * [seg_end_host[s-1], seg_end_host[s]) (seg_end_host[n_seg-1] == n_rows).  Replaces the per-index fan-out of
 * MultiIndexRetriever._retrieve_impl (rag_engine.py:416-461: one vector-store query per index on a thread pool)
 * by ONE corpus pass in which every (segment, query) pair keeps its own shortlists -- a "virtual query"
 * v = s * n_q + q in the outputs, which stage 2 and tt_automerge then treat like any query:
 *
 *   out_ids / out_approx  [n_seg * n_q, n_lists * kprime]      out_thresh  [n_seg * n_q, n_lists]
 *
 * Same contract per virtual query as tt_scan_topk_bf16, with "every row" read as "every row of segment s".
 * tcgen05 variant only (dim % 128 == 0), hi+lo queries (q_lo_bf16 required), n_seg <= TT_MAX_SEGMENTS.
 * seg_end_host is a HOST array, read during the call.
 */
#define TT_MAX_SEGMENTS 16
int tt_scan_topk_bf16_segmented(const void* corpus_bf16, int64_t n_rows, int dim, int64_t row_stride_elems,
                                const float* inv_norm, const void* q_hi_bf16, const void* q_lo_bf16, int n_q,
                                int kprime, int64_t id_base, const int64_t* seg_end_host, int n_seg,
                                int64_t* out_ids, float* out_approx, float* out_thresh,
                                void* ws, size_t ws_bytes, void* stream);

/*
 * Stage 1 for hi-only batches: from a few dozen to tens of thousands of concurrent queries (BASELINE configs[1]'s
 * batch-64 -- HBM-bound, one 64-column pass -- up to config C4, the tensor-bound regime).  Same role and same output contract as tt_scan_topk_bf16 with ONE list per
 * query (n_lists = 1), queries as bf16 hi halves only:
 *
 *   out_ids    int64 [n_q, kprime]   the kprime rows with the best approximate score, best first; -1 = empty
 *   out_approx float [n_q, kprime]
 *   out_thresh float [n_q]           every row NOT emitted has a(r) <= out_thresh[b];  -inf: nothing was
 *                                    left out;  +inf: the query's candidate buffer overflowed (the
 *                                    certificate of tt_rescore_topk then fails and the caller re-runs it)
 *
 * A GEMM-shaped tcgen05 kernel (256 x 256 output tiles per CTA pair, 256 x 128 / 64 when the batch fits one
 * narrower block; corpus and query tiles both TMA-streamed) visits the corpus in phases of geometrically growing size; between phases each query's
 * candidate buffer is cut to its kprime best and the K'-th score becomes the threshold the next phase's
 * epilogue filters with.  The score matrix never reaches HBM.
 * kprime in {128, 256, 512}; dim % 64 == 0; ws: tt_scan_gemm_workspace_bytes(n_q, kprime) bytes
 * (no initialisation needed; 16 * kprime * 8 B per query).
 * q_lo_bf16 != NULL (n_q <= 32): the hi+lo pass -- every query occupies two of the 64 MMA columns (hi half, lo half), the
 * epilogue adds the two accumulators, a(r) carries 16 mantissa bits of the query like tt_scan_topk_bf16 with q_lo: the
 * tight certificate bound at the bytes-in-flight of the streaming pipeline (batches of 17-32 queries).
 */
size_t tt_scan_gemm_workspace_bytes(int n_q, int kprime);
int tt_scan_gemm_topk_bf16(const void* corpus_bf16, int64_t n_rows, int dim, int64_t row_stride_elems,
                           const float* inv_norm, const void* q_hi_bf16, const void* q_lo_bf16 /* nullable */, int n_q,
                           int kprime, int64_t id_base,
                           int64_t* out_ids, float* out_approx, float* out_thresh,
                           void* ws, size_t ws_bytes, void* stream);

/*
 * Stage 2 -- exact re-score of the shortlist + exact top-k.  Together with stage 1 this is the
 * exact brute-force result the reference's (approximate, HNSW) query targets.
 *
 * For every candidate: dot, |c|^2 in fp64 over the stored values, key/score as TT_SCORE_*
 * defines, then the k best by (key desc, id asc).
 *
 *   cand_ids    int64 [n_q, n_cand]   from stage 1 (-1 entries ignored)
 *   cand_thresh float [n_q, n_lists]  from stage 1, or NULL
 *   out_keys    float [n_q, k]        ordering keys (what tt_merge_topk consumes); may be NULL
 *   out_scores  float [n_q, k]        reported scores; -inf padding
 *   out_ids     int64 [n_q, k]        -1 padding
 *   out_margin  float [n_q]           certificate: (k-th exact cosine) - max_l cand_thresh[b,l];
 *                                     the top-k is proven exact iff margin > the stage-1 error
 *                                     bound.  +inf when nothing was left out.  NULL ok.
 *   ws          >= tt_rescore_workspace_bytes(n_q, n_cand) bytes
 */
size_t tt_rescore_workspace_bytes(int n_q, int n_cand);
int tt_rescore_topk(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                    int64_t id_base, const float* q_f32, int n_q,
                    const int64_t* cand_ids, int n_cand, const float* cand_thresh, int n_lists,
                    int k, int score_mode,
                    float* out_keys, float* out_scores, int64_t* out_ids, float* out_margin,
                    void* ws, size_t ws_bytes, void* stream);

/*
 * Exact brute-force scan (fp64 scoring of EVERY row, CUDA cores).  The certificate-failure
 * fallback and the on-GPU secondary oracle for corpora too large for the CPU oracle.
 * Same outputs as tt_rescore_topk (no margin: the result is exact by construction).
 * ws >= tt_scan_exact_workspace_bytes(device, n_q, k).
 */
/*
 * ROW GATE (metadata filters; reference: _build_metadata_filters, rag_engine.py:301-365 -> the
 * `filters=` of index.as_retriever): a search restricted to a subset of the rows passes an inv_norm array in which the
 * excluded rows hold NaN.  Every stage-1 variant then never shortlists them (a NaN score passes no comparison; the
 * certificate's thresholds bound the dropped ELIGIBLE rows only), and tt_scan_exact_f64_gated -- the exact scan computes
 * its own norms -- reads the same array as `row_gate` just for that test.  NULL = every row is eligible.
 */
size_t tt_scan_exact_workspace_bytes(int device, int n_q, int k);
int tt_scan_exact_f64(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                      int64_t id_base, const float* q_f32, int n_q, int k, int score_mode,
                      float* out_keys, float* out_scores, int64_t* out_ids,
                      void* ws, size_t ws_bytes, void* stream);

int tt_scan_exact_f64_gated(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                            int64_t id_base, const float* q_f32, int n_q, int k, int score_mode,
                            const float* row_gate /* [n_rows], NaN = excluded; nullable */,
                            float* out_keys, float* out_scores, int64_t* out_ids,
                            void* ws, size_t ws_bytes, void* stream);

/*
 * k-way merge of per-shard top-k lists after the all-gather (row-sharded corpus).
 *   keys float [n_lists, n_q, k_in], ids int64 [n_lists, n_q, k_in]  (rank-major, as all_gather lays them)
 *   keys_list_stride / ids_list_stride: elements between consecutive lists (0 = dense, n_q*k_in), so
 *   that keys and ids can live in one all-gathered record per rank.
 *   out_scores float [n_q, k_out] (score_mode applied to the merged keys), out_ids int64 [n_q, k_out]
 */
int tt_merge_topk(const float* keys, const int64_t* ids, int n_lists, int64_t keys_list_stride,
                  int64_t ids_list_stride, int n_q, int k_in, int k_out,
                  int score_mode, float* out_scores, int64_t* out_ids, void* stream);

/*
 * Peer exchange for the row-sharded corpus: the all-gather fused into the kernels on either side of it.
 * Every rank owns, per pipeline slot, a receive region [world][record] and a flag array uint32[world] in
 * memory that all ranks have mapped (CUDA peer / symmetric memory; tensor_truth_b200/sharded.py).
 *   tt_rescore_topk_push   = tt_rescore_topk whose selecting kernel ALSO stores this rank's (keys, ids)
 *                            record into every peer's region (slot [rank]) and then raises flag[rank]
 *                            = epoch on every peer with a system-scope release.
 *   tt_exchange_push       = the same publication of an already finished record (after a host-side repair).
 *   tt_merge_topk_pulled   = tt_merge_topk whose kernel first waits (acquire) until all `world` flags of
 *                            this rank's slot have reached `epoch`.
 * record layout: keys float[n_q*k] at offset 0, ids int64[n_q*k] at ids_off_bytes and, when margins_off_bytes
 * != 0, this rank's certificate margins float[n_q] at margins_off_bytes (so that every rank can see, in its own
 * receive region, whether EVERY shard's top-k was proven exact -- no second host round trip); records of
 * consecutive source ranks are rec_stride_bytes apart.  A slot may be reused once every rank has merged it (sharded.py
 * keeps a ring of 4 slots per stream pair, which the data dependencies of the pipeline make sufficient).
 */
#define TT_MAX_PEERS 16
typedef struct tt_exchange {
    int world, rank;
    uint32_t epoch;
    uint64_t rec_stride_bytes;
    uint64_t ids_off_bytes;
    void* peer_recv[TT_MAX_PEERS];      /* peer p: base of ITS receive region for this slot, as mapped here */
    uint32_t* peer_flags[TT_MAX_PEERS]; /* peer p: ITS flag array for this slot, as mapped here             */
    uint32_t* ticket;                   /* local device word, zero between calls (one per lane)             */
    uint64_t margins_off_bytes;         /* 0: margins are not exchanged                                     */
    /* Device-resident epoch (NULL: `epoch` above is used and peer_recv / peer_flags address the slot directly).
     * *epoch_dev counts the pushes this LANE (one stream of a pipelined caller) has completed; a pushing launch
     * opens epoch *epoch_dev + 1 and stores it back when its flags go up, the waiting launch that follows it in
     * stream order reads it.  The slot of the receive ring is epoch % n_slots: peer_recv[p] + slot *
     * slot_stride_bytes, peer_flags[p] + slot * flag_slot_stride.  Nothing in the descriptor changes from step to
     * step, so a captured CUDA graph of the step replays as it is.  All ranks must push in lockstep per lane;
     * 2 slots per lane suffice (slot e % 2 is rewritten at e + 2, after this rank merged e + 1, which needed every
     * peer's push of e + 1, which that peer issued after ITS merge of e). */
    uint32_t* epoch_dev;
    uint32_t n_slots;                   /* 0 or 1: a single slot                                            */
    uint64_t slot_stride_bytes;
    uint64_t flag_slot_stride;          /* in uint32 elements, >= world                                     */
} tt_exchange_t;

/*
 * Certificate for TT_SCORE_CHROMA_L2_EXP.  Stage 1 orders rows by cosine; when every row norm of the shard lies
 * in [row_norm_min, row_norm_max] (unit-norm embeddings: both ~1) a dropped row's squared-L2 key is bounded from
 * above through its cosine bound, and out_margin becomes (k-th exact key) - (that bound, eps included):
 * the top-k is proven exact iff out_margin > 0.  Without it (NULL) L2-mode margins are -inf unless nothing was dropped.
 */
typedef struct tt_l2_cert {
    float row_norm_min, row_norm_max; /* bounds on |c_r| over the rows of this shard, as stored */
    float eps;                        /* bound on |stage-1 score - exact cosine| (what cosine mode compares its margin to) */
} tt_l2_cert_t;

int tt_rescore_topk_push(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                         int64_t id_base, const float* q_f32, int n_q,
                         const int64_t* cand_ids, int n_cand, const float* cand_thresh, int n_lists,
                         int k, int score_mode,
                         float* out_keys, float* out_scores, int64_t* out_ids, float* out_margin,
                         void* ws, size_t ws_bytes, const tt_exchange_t* xchg, const tt_l2_cert_t* l2_cert,
                         void* stream);
int tt_exchange_push(const void* record, size_t nbytes, const tt_exchange_t* xchg, void* stream);
/* Device-side rendezvous of all ranks over the flags of `xchg` (its own descriptor: flags + epoch_dev, no receive
 * region needed -- peer_recv may repeat peer_flags): returns, in stream order, once every rank has reached it. */
int tt_peer_barrier(const tt_exchange_t* xchg, void* stream);
int tt_merge_topk_pulled(const float* keys, const int64_t* ids, int n_lists, int64_t keys_list_stride,
                         int64_t ids_list_stride, int n_q, int k_in, int k_out, int score_mode,
                         float* out_scores, int64_t* out_ids, const tt_exchange_t* xchg, void* stream);

/*
 * Stage 3 -- auto-merge.  Replaces AutoMergingRetriever._retrieve after the base retriever
 * returned (rag_engine.py:641-643; upstream _fill_in_nodes / _get_parents_and_merge /
 * _try_merging loop, then a stable sort by score).  One CTA per query; float64 scores.
 *
 *   ids int64 [n_q, k] (-1 padding at the tail), scores float [n_q, k]
 *   parent_of, child_count, prev_id, next_id: int32 [n_nodes] (tensor_truth_b200/tree.py)
 *   out_ids int64 [n_q, max_out], out_scores double [n_q, max_out], out_len int32 [n_q]
 *   out_len[b] = -1 if the merged list did not fit max_out (or k > tt_automerge_max_k()).
 */
int tt_automerge_max_k(void);
int tt_automerge(const int64_t* ids, const float* scores, int n_q, int k,
                 const int32_t* parent_of, const int32_t* child_count,
                 const int32_t* prev_id, const int32_t* next_id, int64_t n_nodes,
                 double ratio_thresh, int max_rounds,
                 int64_t* out_ids, double* out_scores, int32_t* out_len, int max_out, void* stream);

/*
 * Fused tails: the same arithmetic in fewer launches on the latency chain of a query.
 *
 *   tt_rescore_topk_fused   = tt_rescore_topk_push in ONE launch: the blocks of a query re-score its candidates, the
 *                             block that finishes last (per-query ticket) selects -- and then either pushes the
 *                             record to the peers (xchg != NULL) or runs stage 3 on it (am != NULL, xchg == NULL).
 *                             ws: tt_rescore_fused_workspace_bytes(n_q, n_cand) bytes, ZEROED ONCE by the caller
 *                             when it allocates them (the tickets; the kernel leaves them zero).
 *                             cand_approx (stage 1's out_approx) + prefilter_window > 0 (cosine mode, k <= 32,
 *                             n_cand <= 5120): only candidates whose approximate score is within the window of the
 *                             k-th best approximate score are re-scored -- exact as long as window >= 2 x the
 *                             stage-1 error bound the certificate is checked against (a candidate further below
 *                             cannot reach the exact top-k); one block per query.  A query with more than 1024
 *                             candidates inside the window comes back unproven (out_margin = -inf).
 *   tt_merge_topk_fused     = tt_merge_topk_pulled + (optionally) the certificate margins every source pushed,
 *                             copied out of the receive region into out_all_margins float [n_lists, n_q], +
 *                             (optionally) stage 3 on the merged list in the same block.  out_scores / out_ids
 *                             are required when am != NULL (stage 3 reads the merged list back from them).
 */
typedef struct tt_automerge_args {
    const int32_t* parent_of;
    const int32_t* child_count;
    const int32_t* prev_id;
    const int32_t* next_id;
    int64_t n_nodes;
    double ratio_thresh;
    int max_rounds;
    int64_t* out_ids;    /* [n_q, max_out] */
    double* out_scores;  /* [n_q, max_out] */
    int32_t* out_len;    /* [n_q] */
    int max_out;
} tt_automerge_args_t;

size_t tt_rescore_fused_workspace_bytes(int n_q, int n_cand);
int tt_rescore_topk_fused(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                          int64_t id_base, const float* q_f32, int n_q,
                          const int64_t* cand_ids, int n_cand, const float* cand_thresh, int n_lists,
                          int k, int score_mode,
                          float* out_keys, float* out_scores, int64_t* out_ids, float* out_margin,
                          void* ws, size_t ws_bytes, const tt_exchange_t* xchg, const tt_l2_cert_t* l2_cert,
                          const tt_automerge_args_t* am, const float* cand_approx, float prefilter_window, void* stream);
int tt_merge_topk_fused(const float* keys, const int64_t* ids, int n_lists, int64_t keys_list_stride,
                        int64_t ids_list_stride, int n_q, int k_in, int k_out, int score_mode,
                        float* out_scores, int64_t* out_ids, const tt_exchange_t* xchg, float* out_all_margins,
                        const tt_automerge_args_t* am, void* stream);

/*
 * The stage AFTER the retrieval path: the dense layers of the cross-encoder reranker
 * (SentenceTransformerRerank.postprocess_nodes; services/model_manager.py:333-337, services/rag_service.py:343-346).
 * tensor_truth_b200/rerank.py strings them into an XLM-RoBERTa encoder.
 *
 * tt_linear_bf16:  y[T, n_out] = act(x[T, k_in] W[n_out, k_in]^T + bias) (+ residual[T, n_out]), all bf16 except
 *                  bias (fp32, nullable); fp32 accumulation on tcgen05 (256 x 256 tiles per CTA pair), the layer tail
 *                  fused into the epilogue.  activation: 0 none, 1 exact (erf) GELU.  residual nullable.
 *                  k_in % 128 == 0, n_out % 256 == 0, 16-byte aligned pointers.
 * tt_layernorm_bf16: y = LayerNorm(x) over the last dimension (fp32 statistics), dim % 8 == 0, dim <= 2048.
 * tt_embed_layernorm_bf16: y[t] = LayerNorm(word_emb[word_ids[t]] + pos_emb[pos_ids[t]] (+ type_emb[0])).
 */
#define TT_ACT_NONE 0
#define TT_ACT_GELU 1
int tt_linear_bf16(const void* x_bf16, int64_t n_rows, int k_in, const void* w_bf16, int n_out, const float* bias,
                   const void* residual_bf16, int activation, void* y_bf16, void* stream);
int tt_layernorm_bf16(const void* x_bf16, int64_t n_rows, int dim, const float* gamma, const float* beta, float eps,
                      void* y_bf16, void* stream);
int tt_embed_layernorm_bf16(const int32_t* word_ids, const int32_t* pos_ids, int64_t n_rows, int dim,
                            const void* word_emb_bf16, const void* pos_emb_bf16, const void* type_emb_bf16,
                            const float* gamma, const float* beta, float eps, void* y_bf16, void* stream);

/*
 * tt_attention_varlen_bf16: bidirectional multi-head self-attention over PACKED sequences (no padding), the attention
 *                  of every encoder layer.  qkv bf16 [n_tokens, 3 * n_heads * head_dim] as one fused Q|K|V projection
 *                  writes it; cu_seqlens int32 [n_seq + 1] token offsets; out bf16 [n_tokens, n_heads * head_dim].
 *                  One CTA per (128-row query tile, head): S = Q K^T and O += P V on tcgen05 (accumulators in TMEM),
 *                  softmax in registers, V consumed as an MN-major operand as it lies in memory.  head_dim == 64,
 *                  every sequence <= max_len <= 512; max_tiles >= sum_i ceil(len_i / 128) sizes the grid (surplus
 *                  CTAs exit), e.g. n_tokens / 128 + n_seq.  scale: softmax scale (1 / sqrt(head_dim)).
 * tt_cls_head_f32: logits[i] = w2 . tanh(W1 x[cu_seqlens[i]] + b1) + b2 on the first (<s>) token of every sequence --
 *                  RobertaClassificationHead with one output unit; x bf16 [n_tokens, hidden], fp32 weights.
 */
int tt_attention_varlen_bf16(const void* qkv_bf16, int64_t n_tokens, int n_heads, int head_dim, const int32_t* cu_seqlens,
                             int n_seq, int max_len, int max_tiles, float scale, void* out_bf16, void* stream);
int tt_cls_head_f32(const void* x_bf16, const int32_t* cu_seqlens, int n_seq, int hidden, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* logits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H */
