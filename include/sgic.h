/*
 * sgic.h — C ABI of the B200-native exact inner-product k-NN path.
 *
 * This is the drop-in boundary for the retrieval hot path of
 * lionl1106/Searchable-Generative-Image-Compression.  The reference has no FFI of its own:
 * it reaches the arithmetic through the Python `faiss` module (IndexFlatIP / add / search /
 * ntotal / d / read_index / write_index).  Each entry point below names the reference call
 * site it replaces (paths relative to the reference root).  Plain pointers and sizes only; no
 * torch / numpy types.  Every function returns 0 on success, non-zero on failure;
 * sgic_last_error() returns the calling thread's last message (what the Python layer raises
 * as RuntimeError, mirroring how faiss surfaces C++ exceptions).
 *
 * Buffers named host_* are ordinary host memory (pageable or pinned).  Buffers named dev_*
 * are device pointers on the index's device; `stream` is a cudaStream_t passed as void*
 * (NULL = the index's own stream).  *_dev calls enqueue work and do not synchronise.
 *
 * Streams.  An index owns one non-blocking stream; host-buffer calls (add_f32, add_u8, add_c2df, search,
 * reconstruct, write, write_v2) run on it and return when their work is done.  The *_dev calls run on the
 * stream the caller passes.  The library orders consecutive calls on one index across streams by itself: a
 * call first makes its stream wait (cudaStreamWaitEvent) for whatever the previous call on a DIFFERENT stream
 * enqueued, so add_f32_dev(stream A) followed by search() / write() / search_dev(stream B) sees the new rows
 * without a synchronise in between, and the shared search workspaces are never used by two streams at once.
 * The caller's part: buffers passed to a *_dev call must stay valid until that stream has run the work, and
 * calls on one index are serialised by its mutex (concurrent searches from several host threads take turns; the
 * resident service batches them into one search instead, see service.py).
 */
#ifndef SGIC_H_
#define SGIC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgic_index sgic_index;

enum { SGIC_F16 = 0, SGIC_BF16 = 1 };  /* storage type of the HBM-resident database */

enum {
  SGIC_RETAIN_F32 = 1, /* keep the fp32 rows given to add_f32 on the host so that
                          sgic_index_write is bit-identical to faiss.write_index */
  SGIC_RETAIN_U8 = 2,  /* keep the u8 clip_stream codes of rows added through add_u8 / add_c2df on the
                          host (1 byte per element) so that sgic_index_write regenerates the fp32
                          rows of dequantize_clip_u8 + l2n (src/build.py:18-24) and the file is
                          byte-identical to the one src/build.py:95,99 writes */
};

/* per-file status codes of the batched .c2df ingest; they mirror the exceptions of
 * decode_clip_from_c2df (src/search.py:24-41 == src/build.py:26-43) and unpack_c2df
 * (src/filemaker.py:137-173) that build.py:87-88 turns into "[SKIP] name: err". */
enum {
  SGIC_C2DF_OK = 0,
  SGIC_C2DF_BAD_MAGIC = 1,     /* filemaker.py:145  assert data[:4] == b"C2DF"            */
  SGIC_C2DF_TRUNCATED = 2,     /* struct.error / short slice while walking the TLV entries */
  SGIC_C2DF_NO_CLIP = 3,       /* search.py:26-27   ValueError: no clip_stream / clip_meta */
  SGIC_C2DF_BAD_DIM = 4,       /* search.py:31-33   ValueError: invalid clip_meta.dim      */
  SGIC_C2DF_ZSTD = 5,          /* search.py:35      zstd.ZstdError                         */
  SGIC_C2DF_DIM_MISMATCH = 6,  /* search.py:37-38   ValueError: q.size != dim              */
  SGIC_C2DF_WRONG_D = 7,       /* row dimension differs from the index's d (np.concatenate
                                  would raise in build.py:91)                              */
  SGIC_C2DF_BAD_TYPE = 8,      /* filemaker.py:135  ValueError: unknown type code          */
  SGIC_C2DF_BAD_ENTRY = 9,     /* filemaker.py:102-135,149: the header or an entry does not
                                  load (json.JSONDecodeError, UnicodeDecodeError, numpy's
                                  TypeError / ValueError for an array entry)               */
};

const char* sgic_last_error(void);
int sgic_version(void);

/* faiss.IndexFlatIP(d) — src/build.py:93,232; src/compress.py:97.
 * dtype: SGIC_F16 | SGIC_BF16.  device: CUDA ordinal.  capacity_rows: rows to reserve up
 * front (0 = grow on demand).  d must be a multiple of 8 and <= 2048. */
int sgic_index_create(int d, int dtype, int device, int64_t capacity_rows, int flags, sgic_index** out);
int sgic_index_destroy(sgic_index* h);

/* The same index row-sharded over n_dev GPUs of this box, behind ONE handle and driven by ONE process — the
 * reference's caller is a single process (src/search.py:149-162, webapp.py:246-248), so faiss.IndexFlatIP(d) has
 * to be able to stand for all the GPUs.  Every entry point of this header takes the handle:
 *   add*        large blocks are cut into n_dev contiguous slices (one per GPU), small ones go whole to one GPU;
 *               add_c2df cuts the FILE list into n_dev slices that are walked / decoded concurrently, and a
 *               file's row number is the count of good files before it, exactly as build.py:80-88 numbers them;
 *   search      queries replicated, every GPU scans its rows on its own stream, the shards' final writers store
 *               their (nq, k) answers with global row numbers straight into the home GPU's memory over NVLink
 *               (peer access, no collective), one merge ordered (score desc, row asc) -> the answer is identical
 *               to the single-GPU answer on the same rows;
 *   *_dev       device buffers live on the HOME GPU = dev_ids[0] (sgic_index_device), `stream` is a stream of it;
 *   write       IxFI in global row order; sgic_index_save_shards / sgic_index_load_shards keep one SGI2 file per GPU.
 * dev_ids may name a GPU twice (two shards on one GPU: how the sharded logic is tested on a one-GPU box).
 * capacity_rows is split evenly. */
int sgic_index_create_sharded(int d, int dtype, int n_dev, const int* dev_ids, int64_t capacity_rows, int flags,
                              sgic_index** out);
/* shards behind a handle (1 for an ordinary index); the g-th shard as a BORROWED single-GPU handle — rows may be
 * appended to the shards directly (bulk loaders, generators), after which sgic_index_adopt_shards makes them the
 * index's rows: shard 0's rows first, then shard 1's, ... (one contiguous global range per GPU). */
int sgic_index_n_shards(const sgic_index* h);
int sgic_index_shard(sgic_index* h, int g, sgic_index** out);
int sgic_index_adopt_shards(sgic_index* h);
/* one "shard-%05d-of-%05d.sgi2" file per GPU under dir (written / read concurrently, rows as stored in HBM);
 * load needs as many devices as there are files. */
int sgic_index_save_shards(sgic_index* h, const char* dir);
int sgic_index_load_shards(const char* dir, int n_dev, const int* dev_ids, int flags, sgic_index** out);

/* index.ntotal / index.d — src/search.py:78,114; src/build.py:101,120,240. */
int64_t sgic_index_ntotal(const sgic_index* h);
int sgic_index_d(const sgic_index* h);
int sgic_index_dtype(const sgic_index* h);
int sgic_index_device(const sgic_index* h);

int sgic_index_reserve(sgic_index* h, int64_t rows);
int sgic_index_reset(sgic_index* h);

/* index.add(X) — src/build.py:94,233; src/compress.py:107.  X is (n,d) fp32 C-contiguous;
 * rows are appended as given (no normalisation), rounded to the storage type on device. */
int sgic_index_add_f32(sgic_index* h, int64_t n, const float* host_x);
int sgic_index_add_f32_dev(sgic_index* h, int64_t n, const float* dev_x, void* stream);
/* rows already in the storage type on the device (synthetic generators, shard loaders) */
int sgic_index_add_packed_dev(sgic_index* h, int64_t n, const void* dev_rows, void* stream);

/* Device-side loader for the u8 payload of clip_stream: dequantize_clip_u8 + l2n
 * (src/search.py:16-22) run as one kernel, 1 byte/element over PCIe. */
int sgic_index_add_u8(sgic_index* h, int64_t n, const uint8_t* host_q);
int sgic_index_add_u8_dev(sgic_index* h, int64_t n, const uint8_t* dev_q, void* stream);

/* Batched .c2df ingest — the loop of build_index_from_c2df_dir (src/build.py:80-88).
 * blob holds n files back to back; file i is blob[offsets[i] .. offsets[i+1]) (n+1 offsets).
 * status_out[i] gets a SGIC_C2DF_* code; files with status != 0 are skipped (as [SKIP] does)
 * and the good ones are appended in order.  Returns the number of rows added in *n_added. */
int sgic_index_add_c2df(sgic_index* h, const uint8_t* blob, const int64_t* offsets, int64_t n,
                        int32_t* status_out, int64_t* n_added, int n_threads);

/* Host-only half of the above: walk the TLV container, find clip_stream / clip_meta.dim,
 * zstd-decode into out_u8 (n rows of `dim` bytes; rows of failed files are left untouched).
 * dim_out[i] receives clip_meta.dim (or 0).  No GPU needed. */
int sgic_c2df_parse(const uint8_t* blob, const int64_t* offsets, int64_t n, int dim, uint8_t* out_u8,
                    int32_t* status_out, int32_t* dim_out, int n_threads);

/* index.search(q, k) — src/search.py:115 (faiss search_c(n, x, k, D, I)).
 * host_q (nq,d) fp32; D (nq,k) fp32 and I (nq,k) int64 are caller-owned host buffers.
 * Per query: the k largest inner products sorted descending; ties ordered by ascending row;
 * missing slots are I=-1, D=-FLT_MAX.  k >= 1. */
int sgic_index_search(sgic_index* h, int64_t nq, const float* host_q, int64_t k, float* host_D,
                      int64_t* host_I);
/* Same with device-resident queries and outputs; id_base is added to every returned row
 * number (a shard's first global row). */
int sgic_index_search_dev(sgic_index* h, int64_t nq, const float* dev_q, int64_t k, float* dev_D,
                          int64_t* dev_I, int64_t id_base, void* stream);

/* Merge of per-shard answers after the all-gather: lists laid out
 * [n_lists][nq][k], each sorted (score desc, id asc), ids global, -1 = empty slot.
 * Output (nq,k) ordered (score desc, id asc) — identical to a single-GPU search.
 * tie_by_position = 0: ids < 2^32 (exact for any assignment of rows to shards);
 * tie_by_position = 1: any int64 ids, but shard g must hold rows below shard g+1. */
int sgic_merge_topk_dev(int device, int64_t nq, int n_lists, int64_t k, const float* dev_D_lists,
                        const int64_t* dev_I_lists, float* dev_D, int64_t* dev_I, int tie_by_position,
                        void* stream);

/* K5x — the same exchange + merge WITHOUT a collective library call (SURVEY.md §8e "optional fusion: peer-store
 * epilogue + flag instead of NCCL").  Every rank owns one exchange buffer; the ranks swap its 64-byte CUDA IPC
 * handle once (any transport: the Python host uses torch.distributed.all_gather_object) and map each other's
 * buffers.  sgic_xchg_merge_dev then (1) stores this rank's (nq,k) answer into its slot of EVERY rank's buffer
 * over NVLink and raises a system-scope release flag per destination, (2) merges the `world` lists of its own
 * buffer as soon as their flags carry this step's epoch — two launches on `stream`, identical result to
 * all-gather + sgic_merge_topk_dev.  Collective: every rank calls it once per search, in the same order.
 * nq*k <= max_cands; world <= 16.  sgic_xchg_error() != 0 after a rank waited 20 s for a peer in vain. */
typedef struct sgic_xchg sgic_xchg;
int sgic_xchg_create(int device, int world, int rank, int64_t max_cands, sgic_xchg** out);
int sgic_xchg_export(sgic_xchg* x, uint8_t* handle64);
int sgic_xchg_open(sgic_xchg* x, const uint8_t* handles /* world * 64 bytes, rank order */);
int sgic_xchg_merge_dev(sgic_xchg* x, int64_t nq, int64_t k, const float* dev_D_local, const int64_t* dev_I_local,
                        float* dev_D, int64_t* dev_I, int tie_by_position, void* stream);
/* A rank's whole search step with HOST buffers in one call (what sgic_index_search is for one GPU): H2D of the
 * queries, local scan with id_base added to the row numbers, sgic_xchg_merge_dev, D2H of the merged answer.
 * Collective like sgic_xchg_merge_dev; every rank receives the same (D, I). */
int sgic_xchg_search(sgic_xchg* x, sgic_index* h, int64_t nq, const float* host_q, int64_t k, float* host_D,
                     int64_t* host_I, int64_t id_base, int tie_by_position);
int sgic_xchg_error(sgic_xchg* x);
int sgic_xchg_destroy(sgic_xchg* x);

/* faiss.write_index / faiss.read_index — src/build.py:95,99,235,238; src/compress.py:95,111;
 * src/search.py:69,76.  File layout "IxFI" (SURVEY.md §8a F3): fp32 little-endian rows. */
int sgic_index_write(sgic_index* h, const char* path);
int sgic_index_read(const char* path, int dtype, int device, int flags, sgic_index** out);

/* On-disk format v2 "SGI2" (additive): a 4 KB header page followed by the rows exactly as
 * they sit in HBM (fp16 / bf16).  Half the bytes of the fp32 IxFI file that src/search.py:69,76 re-reads for
 * every query process, no conversion on load.  sgic_index_read recognises both formats by their magic, so
 * load_index (src/search.py:65-88) works on either file.  row_start / total_rows / shard / n_shards describe
 * where the rows sit in a row-sharded logical index (single file: 0, -1, 0, 1). */
int sgic_index_write_v2(sgic_index* h, const char* path, int64_t row_start, int64_t total_rows, int shard,
                        int n_shards);
/* out4 = {row_start, total_rows, shard, n_shards} of an index loaded from an SGI2 file. */
int sgic_index_shard_info(const sgic_index* h, int64_t* out4);

/* The retained u8 codes of rows [i0, i0+n) (SGIC_RETAIN_U8; fails unless every row came in as codes). */
int sgic_index_codes(sgic_index* h, int64_t i0, int64_t n, uint8_t* host_out);
/* dequantize_clip_u8 + l2n (src/search.py:16-22 == src/build.py:18-24) on the host for n rows of d codes, with
 * numpy's operation order (fp32 division by 255, *2, -1; pairwise sum of squares; division by max(norm, 1e-9)):
 * bit-identical to the reference's fp32 rows.  What sgic_index_write uses for an index that retains its codes. */
int sgic_codes_to_f32(const uint8_t* host_q, int64_t n, int d, float* host_out);

/* rows [i0, i0+n) up-cast to fp32 (faiss reconstruct_n); host buffer. */
int sgic_index_reconstruct(sgic_index* h, int64_t i0, int64_t n, float* host_out);

/* Device pointer of the packed database and timing of the last search (kernel-only, ms,
 * measured with CUDA events on the index's stream when enabled). */
const void* sgic_index_data_dev(const sgic_index* h);
/* Option "timing" = 2 brackets the scan kernel of every search with a pair of CUDA events on the search's stream
 * and does NOT synchronise; this call waits for the most recent of them and returns the durations (ms) of the
 * last min(max_n, 256, searches since the previous call) scan kernels, oldest first, then forgets them. */
int sgic_index_scan_times(sgic_index* h, int max_n, float* ms_out, int* n_out);
int sgic_index_set_option(sgic_index* h, const char* name, int64_t value);
int64_t sgic_index_get_stat(const sgic_index* h, const char* name);

#ifdef __cplusplus
}
#endif
#endif /* SGIC_H_ */
