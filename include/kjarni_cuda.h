/* kjarni_cuda.h -- C ABI of libkjarni_cuda.so, the B200 (sm_100a) backend for Kjarni's
 * data-parallel hot path: the batched BERT-family encoder forward behind
 * Embedder / Classifier / Reranker, and the brute-force cosine top-k scan behind
 * Searcher.  This is the inner FFI a Rust `kjarni-cuda-sys` crate would bind
 * (see INTEGRATION.md); it fills the slot that `Device::Wgpu` / `KJARNI_DEVICE_GPU`
 * selects today.  Plain pointers and sizes only; caller-allocated outputs; opaque
 * handles; no allocation crosses the boundary.
 *
 * All `file:line` citations are into the reference tree (olafurjohannsson/kjarni),
 * KT = crates/kjarni-transformers/src, KM = crates/kjarni-models/src,
 * KS = crates/kjarni-search/src, KR = crates/kjarni-rag/src, KF = crates/kjarni-ffi/src.
 *
 * Threading: handles may be used from several host threads; calls on one handle are
 * serialised internally (the reference's wgpu path serialises on a pool mutex too,
 * KT/cpu/encoder/traits.rs:108-112).  Every call returns after its stream has
 * synchronised unless the name ends in `_async`.
 */
#ifndef KJARNI_CUDA_H
#define KJARNI_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numbering as KjarniErrorCode (KF/src/error.rs:14-40). */
typedef enum KjcStatus {
    KJC_OK = 0,
    KJC_NULL_POINTER = 1,
    KJC_INVALID_UTF8 = 2,
    KJC_MODEL_NOT_FOUND = 3,
    KJC_LOAD_FAILED = 4,
    KJC_INFERENCE_FAILED = 5,
    KJC_GPU_UNAVAILABLE = 6,
    KJC_INVALID_CONFIG = 7,
    KJC_CANCELLED = 8,
    KJC_TIMEOUT = 9,
    KJC_STREAM_ENDED = 10,
    KJC_UNKNOWN = 255
} KjcStatus;

/* Thread-local message of the last failing call on this thread; valid until the next
 * failure on the same thread (KF/src/error.rs:7-9,57-101). */
const char* kjc_last_error_message(void);
void kjc_clear_error(void);
const char* kjc_error_name(int status);
const char* kjc_version(void);
/* Number of visible CUDA devices; a negative value means no usable driver. */
int kjc_device_count(void);

/* ------------------------------------------------------------------ encoder */

typedef struct KjcEncoder KjcEncoder;

/* Layout selected from config.json (KM/models/sentence_encoder/model.rs:41-54,
 * KM/models/sentence_encoder/configs.rs:271-364,638-687). */
typedef enum KjcArch {
    KJC_ARCH_BERT = 0,
    KJC_ARCH_BERT_PREFIXED = 1,
    KJC_ARCH_DISTILBERT = 2,
    KJC_ARCH_ROBERTA = 3, /* `roberta.` prefix, positions from row 2 (KM/models/sequence_classifier/configs.rs:147-283) */
    KJC_ARCH_MPNET = 4    /* attention.attn.{q,k,v,o}, tanh-GELU, positions from row 2 (KM/models/sentence_encoder/configs.rs:370-468) */
} KjcArch;
/* Head auto-detected from tensor names, first match wins (KT/cpu/encoder/classifier.rs:113-206). */
typedef enum KjcHeadKind {
    KJC_HEAD_ABSENT = 0,      /* sentence encoder: no classifier tensors */
    KJC_HEAD_DENSE_TANH = 1,  /* classifier.dense + tanh + classifier.out_proj */
    KJC_HEAD_PRE_RELU = 2,    /* pre_classifier + ReLU + classifier (DistilBERT) */
    KJC_HEAD_POOLER_TANH = 3, /* bert.pooler.dense + tanh + classifier (cross-encoder) */
    KJC_HEAD_LINEAR = 4       /* classifier only */
} KjcHeadKind;

typedef struct KjcEncoderInfo {
    int32_t arch;         /* KjcArch */
    int32_t hidden_size;
    int32_t num_layers;
    int32_t num_heads;
    int32_t intermediate_size; /* read from the fc1 weight shape, never from metadata */
    int32_t vocab_size;
    int32_t max_position_embeddings;
    int32_t type_vocab_size; /* 0: no token-type table (DistilBERT) */
    int32_t position_offset; /* extra_pos_embeddings: 0 BERT/DistilBERT, 2 RoBERTa/MPNet */
    int32_t head_kind;       /* KjcHeadKind */
    int32_t num_labels;      /* rows of the classifier weight; 0 without a head */
    int32_t device;
    float layer_norm_eps;
} KjcEncoderInfo;

/* Loads `<model_dir>/config.json` + `<model_dir>/model.safetensors` (F32/F16/BF16 tensors),
 * resolves the tensor-name layout, pre-packs GEMM weights to bf16 (Q/K/V fused to [3H,H])
 * and uploads everything to `device` once.  Mirrors EncoderLoader::load_from_pretrained
 * (KT/pipeline/encoder/loader.rs:82-141) + CpuTransformerEncoder::new
 * (KT/cpu/encoder/transformer_encoder.rs:30-235).  `tokenizer.json` is not read here
 * (token ids come in through the API).
 * Errors: KJC_MODEL_NOT_FOUND (missing dir/files), KJC_LOAD_FAILED (bad safetensors, missing
 * tensor, shape mismatch), KJC_INVALID_CONFIG (bad config.json / unsupported sizes),
 * KJC_GPU_UNAVAILABLE (no device / not sm_100). */
int kjc_encoder_create(const char* model_dir, int device, KjcEncoder** out);
/* The same handle type backed by ONE WEIGHT REPLICA PER DEVICE of `device_ids[0..n_devices)`, inside this process (the Rust / C# /
 * Go hosts have no torchrun): one persistent host thread, stream and pinned staging slice per GPU.  kjc_encoder_forward splits a
 * host batch by contiguous rows over the replicas -- sequences are independent (KT/cpu/encoder/traits.rs:66-139), so there is NO
 * collective and the rows are bit-identical to a one-GPU forward of the same batch.  The device-pointer entry
 * (kjc_encoder_forward_device_async), profiling and the debug hooks address replica 0.
 * A device may be listed more than once (several replicas on one GPU; only useful for testing the split on a one-GPU box).
 * Errors: as kjc_encoder_create, plus KJC_INVALID_CONFIG for an empty list. */
int kjc_encoder_create_multi(const char* model_dir, const int* device_ids, int n_devices, KjcEncoder** out);
int kjc_encoder_device_count(const KjcEncoder* enc);
void kjc_encoder_destroy(KjcEncoder* enc);
int kjc_encoder_info(const KjcEncoder* enc, KjcEncoderInfo* out);
/* Label `i` of config.json:id2label (sorted by numeric key), or NULL. Borrowed; valid until destroy.
 * (KF/src/classifier.rs:253-300 labels / num_labels.) */
const char* kjc_encoder_label(const KjcEncoder* enc, int i);

typedef enum KjcOutput {
    KJC_OUT_HIDDEN = 0,     /* out: f32 [B,S,H] last hidden state -- get_hidden_states_batch_from_ids, KT/cpu/encoder/traits.rs:66-139 */
    KJC_OUT_POOLED = 1,     /* out: f32 [B,H]   pooled (+L2)      -- encode_batch, KT/cpu/encoder/traits.rs:204-225,529-536 */
    KJC_OUT_LOGITS = 2      /* out: f32 [B,C]   head logits       -- predict_logits, KM/models/sequence_classifier/mod.rs:266-346;
                                                                     predict_pairs, KM/models/cross_encoder/model.rs:170-240 */
} KjcOutput;
typedef enum KjcPooling { KJC_POOL_MEAN = 0, KJC_POOL_CLS = 1, KJC_POOL_MAX = 2, KJC_POOL_LAST = 3 } KjcPooling;
/* Which reference CPU variant's padding convention to reproduce (they differ only for a fully
 * padded sequence): ALLOC = masked score := -1e9 (KT/utils/masks.rs:4-36; Classifier/Reranker always),
 * NOALLOC = -inf (KT/cpu/encoder/encoder_self_attention.rs:311-325; Embedder when B*S >= 1000 or <= 1),
 * AUTO = ComputeStrategy::select (KT/cpu/strategy.rs:29-47) for POOLED/HIDDEN, ALLOC for LOGITS. */
typedef enum KjcMaskConvention { KJC_MASK_AUTO = 0, KJC_MASK_ALLOC = 1, KJC_MASK_NOALLOC = 2 } KjcMaskConvention;

typedef struct KjcForwardOptions {
    int32_t output;          /* KjcOutput */
    int32_t pooling;         /* KjcPooling (POOLED only) */
    int32_t normalize;       /* L2-normalise pooled rows (POOLED only) */
    int32_t mask_convention; /* KjcMaskConvention */
} KjcForwardOptions;

/* One encoder forward over a batch of token ids, HOST buffers in and out.
 *   ids       u32 [B*S]  token ids (id >= vocab -> zero embedding row, KT/cpu/embeddings/mod.rs:227-246)
 *   mask      f32 [B*S]  1 = token, 0 = padding; NULL = all ones
 *   type_ids  u32 [B*S]  or NULL (row 0 of the type table is added to every token, KT/cpu/embeddings/mod.rs:216-223)
 *   out       f32, caller-allocated, size per `opts->output`
 * Host->device copies of ids/mask/type_ids and the device->host copy of `out` are part of the call.
 * Errors: KJC_NULL_POINTER, KJC_INVALID_CONFIG (S > max positions is allowed as in the reference: positions
 * beyond the table add nothing; S > 512 or B*S == 0 is rejected), KJC_INFERENCE_FAILED (CUDA error, or a
 * token-type id out of range -- the reference panics there, KT/cpu/embeddings/mod.rs:312-317). */
int kjc_encoder_forward(KjcEncoder* enc, const uint32_t* ids, const float* mask, const uint32_t* type_ids, int batch,
                        int seq_len, const KjcForwardOptions* opts, float* out);

/* Same forward with DEVICE pointers (all on the encoder's device), enqueued on `stream`
 * (a cudaStream_t; NULL = the encoder's own stream) without synchronising. */
int kjc_encoder_forward_device_async(KjcEncoder* enc, const uint32_t* d_ids, const float* d_mask, const uint32_t* d_type_ids,
                                     int batch, int seq_len, const KjcForwardOptions* opts, float* d_out, void* stream);
/* Largest number of sequences processed per internal micro-batch for `seq_len` (activations are
 * sized to stay L2-resident); larger batches are looped internally. */
int kjc_encoder_micro_batch(const KjcEncoder* enc, int seq_len);
/* 1 when the handle runs out-proj + LN1 -> FFN-up and FFN-down + LN2 -> next QKV as one launch each (hidden size 384); the
 * per-kernel-class profile then books those launches under KJC_K_GEMM_FFN_UP / KJC_K_GEMM_FFN_DOWN. */
int kjc_encoder_chained(const KjcEncoder* enc);
/* Precision of the residual stream between kernels (every replica of the handle):
 *   0  bf16 for every output -- the fused / chained kernels everywhere (fastest; hidden states then carry one bf16 rounding per
 *      LayerNorm output: max-abs error up to ~6e-2 at |x| ~ 5 against the fp32 reference);
 *   1  (default) fp32 for KJC_OUT_HIDDEN, bf16 for pooled embeddings and logits: the hidden states -- the tensor
 *      EncoderLanguageModel::get_hidden_states_batch_from_ids returns (KT/cpu/encoder/traits.rs:66-139) -- keep an fp32 residual
 *      stream and only the GEMM operands are rounded to bf16 (max-abs error < 2e-2, measured ~8e-3);
 *   2  fp32 for every output.
 * The fp32 mode runs projection -> fp32 sums (+ fp32 residual) -> LayerNorm kernel instead of the fused kernels. */
int kjc_encoder_set_fp32_residual(KjcEncoder* enc, int mode);
/* Number of kernel launches the last forward on this handle enqueued (for bench bookkeeping). */
int64_t kjc_encoder_last_launch_count(const KjcEncoder* enc);

/* Per-kernel-class device timing (CUDA events around every launch of the forwards issued while profiling
 * is on).  Used by bench.py for the roofline numbers; perturbs overlap slightly, so never on in a timed region. */
typedef enum KjcKernelClass {
    KJC_K_EMBED_LN = 0,      /* gather + LayerNorm */
    KJC_K_GEMM_QKV = 1,
    KJC_K_ATTENTION = 2,
    KJC_K_GEMM_OUT = 3,      /* out-proj + bias + residual */
    KJC_K_LAYERNORM = 4,
    KJC_K_GEMM_FFN_UP = 5,   /* + bias + GELU */
    KJC_K_GEMM_FFN_DOWN = 6, /* + bias + residual */
    KJC_K_OUTPUT = 7,        /* pool+L2 / head / hidden copy */
    KJC_NUM_KERNEL_CLASSES = 8
} KjcKernelClass;
int kjc_encoder_set_profiling(KjcEncoder* enc, int on);
/* ms[KJC_NUM_KERNEL_CLASSES], launches[KJC_NUM_KERNEL_CLASSES]: totals since profiling was switched on. */
int kjc_encoder_get_profile(KjcEncoder* enc, double* ms, int64_t* launches);

/* softmax over the label axis in place (SequenceClassifier::classify_scores_batch,
 * KM/models/sequence_classifier/mod.rs:248-263 -> KT/activations.rs:223-242). Host-side helper. */
void kjc_softmax_rows(float* logits, int rows, int cols);
/* sigmoid per logit in place: ClassificationMode::MultiLabel scores (kjarni/src/classifier/model.rs:313-335,528-531).
 * Host-side helper. */
void kjc_sigmoid_rows(float* logits, int rows, int cols);

/* ---------------------------------------------------------------- tokenizer
 * Text -> ids for the WordPiece / WordLevel tokenizer.json files of the BERT-family checkpoints, with the configuration
 * EncoderLoader applies (KT/pipeline/encoder/loader.rs:99-115): truncation to `max_length` (LongestFirst, right) and
 * BatchLongest padding with id 0; pairs are encoded as `[CLS] a [SEP] b [SEP]` with type ids 0/1 by the file's
 * post-processor (KM/models/cross_encoder/model.rs:176-194).  Host-only; restates the `tokenizers` 0.22 crate's published
 * algorithm (kjarni_b200/csrc/tokenizer.hpp), pinned by goldens from that crate's Python binding. */
typedef struct KjcTokenizer KjcTokenizer;
/* max_length <= 0: no truncation.  Errors: KJC_MODEL_NOT_FOUND, KJC_LOAD_FAILED (bad JSON), KJC_INVALID_CONFIG (BPE /
 * Unigram models, unsupported normalizer / pre-tokenizer / post-processor). */
int kjc_tokenizer_create(const char* tokenizer_json_path, int max_length, KjcTokenizer** out);
void kjc_tokenizer_destroy(KjcTokenizer* tok);
/* Tokenizer::token_to_id: 1 and *out_id set when the token is in the vocabulary (or an added token), else 0. */
int kjc_tokenizer_token_to_id(const KjcTokenizer* tok, const char* token, uint32_t* out_id);
/* encode_batch(texts, add_special_tokens): `pairs` NULL or one second segment per text.  *out_seq_len = padded length.
 * ids NULL: only the length is computed; otherwise ids / mask / type_ids are [n, cap_seq] (mask, type_ids may be NULL)
 * and cap_seq >= *out_seq_len is required (KJC_INVALID_CONFIG otherwise); columns beyond the length are zero. */
int kjc_tokenizer_encode_batch(const KjcTokenizer* tok, const char* const* texts, const char* const* pairs, int n, int add_special_tokens,
                               uint32_t* ids, float* mask, uint32_t* type_ids, int cap_seq, int* out_seq_len);

/* ---------------------------------------------------------------- index scan */

typedef struct KjcIndex KjcIndex;
typedef enum KjcScanMode {
    KJC_SCAN_SEGMENT = 0,     /* Segment::search_vectors semantics, KR/segment.rs:307-337,355-370 */
    KJC_SCAN_VECTORSTORE = 1  /* VectorStore::search semantics, KS/vector.rs:131-165 */
} KjcScanMode;

/* An index shard: `capacity_rows` x `dim` fp32 rows resident in HBM on `device`.
 * `id_base` is the global id of local row 0 (IndexReader::local_to_global, KR/index_reader.rs:313-319). */
int kjc_index_create(int dim, uint64_t capacity_rows, uint64_t id_base, int device, KjcIndex** out);
void kjc_index_destroy(KjcIndex* idx);
uint64_t kjc_index_len(const KjcIndex* idx);
int kjc_index_dim(const KjcIndex* idx);
/* Append `n` rows from host memory (raw little-endian f32, the `vectors.bin` layout, KR/segment.rs:87-123). */
int kjc_index_add_rows(KjcIndex* idx, const float* rows, uint64_t n);
/* Append the rows of one on-disk segment file `<segment_dir>/vectors.bin` (KR/segment.rs:195-262);
 * the row count is file_size / (4*dim). */
int kjc_index_load_vectors_bin(KjcIndex* idx, const char* path);
/* Append `n` rows of the counter-based synthetic generator shared with the oracle
 * (bench / test input only): x[r,c] = (hash32(seed, row0+r, c) >> 8) * 2^-24 - 0.5. */
int kjc_index_append_synthetic(KjcIndex* idx, uint32_t seed, uint64_t row0, uint64_t n);
/* ---- on-disk index directory (row f1 of SURVEY 8f): what IndexReader::open reads, KR/index_reader.rs:161-204 ----
 *   <root>/config.json                      IndexConfig {dimension, max_docs_per_segment, ...}   KR/config.rs:5-27
 *   <root>/segments/seg_%06d/segment.json   SegmentMeta {id, doc_count, dimension, ...}          KR/segment.rs:10-19
 *   <root>/segments/seg_%06d/vectors.bin    raw LE f32 [doc_count, dimension]                    KR/segment.rs:87-123
 * Segment directories are taken in file-name order; one that Segment::open would reject (segment.json, vectors.bin,
 * docs.idx or bm25.bin missing, unparsable segment.json) is skipped as the reference does (it logs a warning), so global
 * ids = sum of preceding segment lengths + local id (KR/index_reader.rs:313-319) agree with the host-side IndexReader,
 * which keeps serving text / metadata / BM25 from the same directory. */
typedef struct KjcIndexDirInfo {
    int32_t dimension;
    int32_t n_segments;           /* segments that load */
    int32_t n_skipped;            /* segment directories skipped (see above) */
    int32_t reserved;
    uint64_t total_rows;          /* IndexReader::len */
    uint64_t max_docs_per_segment;
} KjcIndexDirInfo;
/* Host-only (no GPU needed). Errors: KJC_MODEL_NOT_FOUND (no config.json), KJC_LOAD_FAILED (bad config.json, a segment whose
 * dimension differs from the index's), KJC_INVALID_CONFIG. */
int kjc_index_dir_info(const char* root, KjcIndexDirInfo* out);
/* doc_count of each loadable segment in order (up to `cap` entries); returns the number of segments, or -1 on error. */
int kjc_index_dir_segment_lens(const char* root, uint64_t* out_lens, int cap);
/* Global-id range [*lo, *hi) of part `part` of `parts` over `total_rows` rows: contiguous blocks whose sizes differ by at
 * most one row -- the row-sharding of SURVEY 8e (rank r of a world of `parts` loads part r). */
int kjc_index_part_range(uint64_t total_rows, int part, int parts, uint64_t* lo, uint64_t* hi);
/* New shard on `device` holding rows [lo, hi) of part `part` of `parts` of the on-disk index; id_base = lo; rows are copied
 * from the mmap'ed vectors.bin files through pinned staging.  parts = 1 loads the whole index on one GPU. */
int kjc_index_open_dir(const char* root, int device, int part, int parts, KjcIndex** out);
/* Global id of local row 0 of this shard. */
uint64_t kjc_index_id_base(const KjcIndex* idx);

/* Copy local rows [row, row+n) back to the host (Segment::get_embedding, KR/segment.rs:240-262). */
int kjc_index_get_rows(const KjcIndex* idx, uint64_t row, uint64_t n, float* out);

/* Top-k cosine search of `nq` HOST queries [nq, dim] against this shard.
 *   out_ids    u64 [nq*k]  global ids, order (score desc, id asc); UINT64_MAX where fewer than k results
 *   out_scores f32 [nq*k]  -inf where empty
 *   out_counts i32 [nq] or NULL: number of valid results per query (0 for a zero-norm query in SEGMENT mode)
 * Errors: KJC_NULL_POINTER, KJC_INVALID_CONFIG (k < 1 or k > 256, nq < 1), KJC_INFERENCE_FAILED. */
int kjc_index_search(KjcIndex* idx, const float* queries, int nq, int k, int mode, uint64_t* out_ids, float* out_scores,
                     int32_t* out_counts);
/* Device-pointer variant on `stream` (NULL = the index's stream): this shard's sorted candidates -- the per-segment top-k of
 * IndexReader::search_semantic -- left in device memory for the gather + merge.  ALWAYS exact: it reads one 4-byte counter back
 * (one stream synchronise) and re-runs on the exact scan every query the tensor-core filter could not prove, like kjc_index_search.
 * This is the entry the multi-GPU paths use (kjarni_b200/distributed.py, kjc_sharded_index_search). */
int kjc_index_search_device(KjcIndex* idx, const float* d_queries, int nq, int k, int mode, uint64_t* d_out_ids, float* d_out_scores,
                            int32_t* d_out_counts, void* stream);
/* The same without any synchronise (graph capture, latency-critical pipelines): queries the filter could not prove are NOT
 * re-run, only counted in kjc_index_unverified_count -- the caller must poll that counter (or use the entry above). */
int kjc_index_search_device_async(KjcIndex* idx, const float* d_queries, int nq, int k, int mode, uint64_t* d_out_ids,
                                  float* d_out_scores, int32_t* d_out_counts, void* stream);
/* Merge `n_lists` sorted candidate lists (e.g. one per shard/rank after the NVLink gather) into the final
 * top-k per query: concat -> stable sort desc -> truncate of KR/index_reader.rs:212-224 with the canonical
 * tie-break (score desc, global id asc).  Device pointers: cand_* are [n_lists, nq, k]. */
int kjc_topk_merge_device_async(int device, const uint64_t* d_cand_ids, const float* d_cand_scores, int n_lists, int nq, int k,
                                uint64_t* d_out_ids, float* d_out_scores, int32_t* d_out_counts, void* stream);
/* The same for candidates gathered with ONE collective: `d_packed` holds n_lists records of kjc_packed_record_bytes(nq, k) bytes
 * (12*nq*k rounded up to 16), each = one rank's [nq,k] u64 ids followed by its [nq,k] f32 scores (the rank points
 * kjc_index_search_device's two outputs into one buffer and all-gathers that buffer). */
size_t kjc_packed_record_bytes(int nq, int k);
int kjc_topk_merge_packed_device_async(int device, const void* d_packed, int n_lists, int nq, int k, uint64_t* d_out_ids, float* d_out_scores,
                                       int32_t* d_out_counts, void* stream);
int64_t kjc_index_last_launch_count(const KjcIndex* idx);
/* Searches with k <= 64 on shards whose dim is a multiple of 64 and <= 1024 run as a tensor-core similarity GEMM over a
 * bf16 shadow of the shard that only FILTERS 32 (k <= 16) or 128 (k <= 64) candidates per query; the candidates are re-scored in exact fp32 and a
 * per-query bound proves the exact top-k (see kjarni_b200/csrc/scan_gemm.cuh).  A query whose bound does not hold is re-run
 * on the exact scan by kjc_index_search (which may synchronise); kjc_index_search_device_async cannot synchronise and
 * counts such queries here instead (synchronises the device; 0 = every async result so far was proven exact). */
int64_t kjc_index_unverified_count(KjcIndex* idx);

/* ---- row-sharded index over several devices of one process -------------------------------------------------------------
 * Contiguous row shards, shard p on device_ids[p] holds global rows [lo_p, hi_p) (kjc_index_part_range; global id = shard base +
 * local id, KR/index_reader.rs:313-319).  A search runs the same query batch on every shard concurrently (one host thread per
 * GPU; per-shard results are always exact, see kjc_index_search_device), gathers the per-shard [nq,k] candidates on
 * device_ids[0] with peer copies over NVLink (12 B per candidate) and merges them with the merge kernel: the per-segment top-k ->
 * concat -> sort -> truncate of IndexReader::search_semantic (KR/index_reader.rs:207-228).  Ids, scores, order (score desc, id
 * asc) and counts are those of one shard holding all the rows. */
typedef struct KjcShardedIndex KjcShardedIndex;
/* Empty index of `capacity_rows` rows in total; rows arrive in global-id order (add_rows / append_synthetic). */
int kjc_sharded_index_create(int dim, uint64_t capacity_rows, const int* device_ids, int n_devices, KjcShardedIndex** out);
/* IndexReader::open (KR/index_reader.rs:161-204) with the rows of the on-disk index split over the devices. */
int kjc_sharded_index_open_dir(const char* root, const int* device_ids, int n_devices, KjcShardedIndex** out);
void kjc_sharded_index_destroy(KjcShardedIndex* idx);
uint64_t kjc_sharded_index_len(const KjcShardedIndex* idx);
int kjc_sharded_index_dim(const KjcShardedIndex* idx);
int kjc_sharded_index_shards(const KjcShardedIndex* idx);
uint64_t kjc_sharded_index_shard_len(const KjcShardedIndex* idx, int shard);
int kjc_sharded_index_add_rows(KjcShardedIndex* idx, const float* rows, uint64_t n);
/* Appends the next `n` rows of the deterministic synthetic sequence of `seed` (row g of the index = synthetic row g). */
int kjc_sharded_index_append_synthetic(KjcShardedIndex* idx, uint32_t seed, uint64_t n);
/* Same outputs as kjc_index_search, over all shards. */
int kjc_sharded_index_search(KjcShardedIndex* idx, const float* queries, int nq, int k, int mode, uint64_t* out_ids, float* out_scores,
                             int32_t* out_counts);
/* Launches of the last search summed over the shards + the merge. */
int64_t kjc_sharded_index_last_launch_count(const KjcShardedIndex* idx);

/* cosine of two host vectors (kjarni_cosine_similarity, KF/src/lib.rs:177-188 -> KS/vector.rs:131-148). */
float kjc_cosine_similarity(const float* a, const float* b, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* KJARNI_CUDA_H */
