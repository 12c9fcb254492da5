/*
 * sourmash.h -- the drop-in boundary of the B200 build.
 *
 * This header declares, with identical names, argument order, types, ownership and error
 * conventions, every symbol the reference exports through src/ffi.rs, src/utils.rs and
 * src/errors.rs (its own header is generated from those files by cbindgen, Makefile:14-15).
 * A host that loads the reference's libsourmash (Python via cffi/milksnake, README.md:28-31)
 * can load this build's libsourmash.so instead.  Each declaration cites the reference
 * definition it replaces as `file:line` relative to the reference tree.
 *
 * Conventions (reference src/utils.rs:14-45,154-166):
 *   - handles are opaque heap objects; whoever receives one from a *_new / *_first_mh /
 *     *_get_mhs / *_load_* call frees it with the matching *_free, exactly once;
 *   - fallible calls record their error in a THREAD-LOCAL slot and return an all-zero value
 *     (0, 0.0, NULL, empty string); the slot is not cleared on success, so callers clear it
 *     with sourmash_err_clear() and read it with sourmash_err_get_last_code() around a call;
 *   - SourmashStr is not NUL-terminated; free it with sourmash_str_free() (a no-op unless
 *     `owned`);
 *   - a handle may be used from any thread, but not from two threads at once.
 *
 * Where the work happens: sketch state lives in B200 HBM and every sketching / merging /
 * comparing entry point runs CUDA kernels (sourmash_rust_b200/csrc).  There is no CPU
 * fallback: without a usable sm_100 device those entry points fail with
 * SOURMASH_ERROR_CODE_INTERNAL.  Batch entry points for whole read sets and sketch
 * collections, which this per-object ABI cannot express, are in sourmash_b200.h.
 */
#ifndef SOURMASH_B200_SOURMASH_H
#define SOURMASH_B200_SOURMASH_H

#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes: src/errors.rs:28-50 ------------------------------------------------ */
/* (C keeps tags and typedef names apart, so the reference header reuses the name for both; a
 * C++ translation unit needs a distinct tag) */
#ifdef __cplusplus
enum SourmashErrorCodeValue {
#else
enum SourmashErrorCode {
#endif
  SOURMASH_ERROR_CODE_NO_ERROR = 0,
  SOURMASH_ERROR_CODE_PANIC = 1,          /* unexpected C++ exception (reference: Rust panic) */
  SOURMASH_ERROR_CODE_INTERNAL = 2,       /* includes "no CUDA device" */
  SOURMASH_ERROR_CODE_MSG = 3,
  SOURMASH_ERROR_CODE_UNKNOWN = 4,        /* io / utf-8 / JSON errors land here, as errors.rs:52-73 */
  SOURMASH_ERROR_CODE_MISMATCH_K_SIZES = 101,
  SOURMASH_ERROR_CODE_MISMATCH_D_N_A_PROT = 102,
  SOURMASH_ERROR_CODE_MISMATCH_MAX_HASH = 103,
  SOURMASH_ERROR_CODE_MISMATCH_SEED = 104,
  SOURMASH_ERROR_CODE_INVALID_D_N_A = 1101,
  SOURMASH_ERROR_CODE_INVALID_PROT = 1102,
  SOURMASH_ERROR_CODE_IO = 100001,
  SOURMASH_ERROR_CODE_UTF8_ERROR = 100002,
  SOURMASH_ERROR_CODE_PARSE_INT = 100003,
  SOURMASH_ERROR_CODE_SERDE_ERROR = 100004,
};
typedef uint32_t SourmashErrorCode;

/* ---- opaque handles -------------------------------------------------------------------- */
typedef struct KmerMinHash KmerMinHash; /* src/lib.rs:37-46 */
typedef struct Signature Signature;     /* src/lib.rs:546-565 */

/* src/utils.rs:169-174 */
typedef struct {
  char *data;
  uintptr_t len;
  bool owned;
} SourmashStr;

/* ---- hashing ---------------------------------------------------------------------------- */
/* MurmurHash3 x64_128 of the NUL-terminated bytes, first 64-bit half.  src/ffi.rs:15-24 */
uint64_t hash_murmur(const char *kmer, uint64_t seed);

/* ---- KmerMinHash: lifecycle ---------------------------------------------------------------- */
/* src/ffi.rs:26-44 -> KmerMinHash::new, src/lib.rs:142-174 */
KmerMinHash *kmerminhash_new(uint32_t n, uint32_t k, bool prot, uint64_t seed, uint64_t mx,
                             bool track_abundance);
/* NULL is accepted.  src/ffi.rs:46-53 */
void kmerminhash_free(KmerMinHash *ptr);

/* ---- KmerMinHash: ingest ------------------------------------------------------------------ */
/* NUL-terminated DNA; with force == false the first k-mer holding a non-ACGT byte sets
 * SOURMASH_ERROR_CODE_INVALID_D_N_A after the k-mers before it were added; with force == true
 * such k-mers are skipped.  src/ffi.rs:55-70 -> src/lib.rs:252-274 */
void kmerminhash_add_sequence(KmerMinHash *ptr, const char *sequence, bool force);
/* src/ffi.rs:72-80 -> src/lib.rs:192-245 */
void kmerminhash_add_hash(KmerMinHash *ptr, uint64_t h);
/* hash the NUL-terminated word with the sketch's seed and add it.  src/ffi.rs:82-95 */
void kmerminhash_add_word(KmerMinHash *ptr, const char *word);
/* add_hash of every min of `other`.  src/ffi.rs:260-274 -> src/lib.rs:405-410 */
void kmerminhash_add_from(KmerMinHash *ptr, const KmerMinHash *other);
/* raw appends that bypass ordering (used to reload a stored sketch).  src/ffi.rs:143-150,179-188 */
void kmerminhash_mins_push(KmerMinHash *ptr, uint64_t val);
void kmerminhash_abunds_push(KmerMinHash *ptr, uint64_t val);

/* ---- KmerMinHash: combine and compare ------------------------------------------------------- */
/* set union, abundances summed, truncated to num.  src/ffi.rs:244-258 -> src/lib.rs:307-403 */
void kmerminhash_merge(KmerMinHash *ptr, const KmerMinHash *other);
/* Jaccard over bottom-num of the union.  src/ffi.rs:311-325 -> src/lib.rs:501-508 */
double kmerminhash_compare(KmerMinHash *ptr, const KmerMinHash *other);
/* |A n B|.  src/ffi.rs:276-290 -> src/lib.rs:428-436 */
uint64_t kmerminhash_count_common(KmerMinHash *ptr, const KmerMinHash *other);
/* returns the size of the (truncated) union; 0 and NO error when incompatible.
 * src/ffi.rs:292-309 -> src/lib.rs:438-468 */
uint64_t kmerminhash_intersection(KmerMinHash *ptr, const KmerMinHash *other);

/* ---- KmerMinHash: read-out ----------------------------------------------------------------- */
/* fresh copies owned by the caller (the reference exports no matching free; this build adds
 * kmerminhash_slice_free in sourmash_b200.h).  src/ffi.rs:97-122 */
const uint64_t *kmerminhash_get_mins(KmerMinHash *ptr);
const uint64_t *kmerminhash_get_abunds(KmerMinHash *ptr); /* NULL when not tracking */
uintptr_t kmerminhash_get_mins_size(KmerMinHash *ptr);   /* src/ffi.rs:134-141 */
uintptr_t kmerminhash_get_abunds_size(KmerMinHash *ptr); /* src/ffi.rs:166-177 */
/* out-of-range idx: returns 0 and records SOURMASH_ERROR_CODE_PANIC.  src/ffi.rs:124-132,152-164 */
uint64_t kmerminhash_get_min_idx(KmerMinHash *ptr, uint64_t idx);
uint64_t kmerminhash_get_abund_idx(KmerMinHash *ptr, uint64_t idx);
/* src/ffi.rs:190-242 */
bool kmerminhash_is_protein(KmerMinHash *ptr);
uint64_t kmerminhash_seed(KmerMinHash *ptr);
bool kmerminhash_track_abundance(KmerMinHash *ptr);
uint32_t kmerminhash_num(KmerMinHash *ptr);
uint32_t kmerminhash_ksize(KmerMinHash *ptr);
uint64_t kmerminhash_max_hash(KmerMinHash *ptr);

/* ---- Signature -------------------------------------------------------------------------------- */
Signature *signature_new(void);                                  /* src/ffi.rs:329-332 */
void signature_free(Signature *ptr);                             /* src/ffi.rs:334-342 */
void signature_set_name(Signature *ptr, const char *name);       /* src/ffi.rs:344-362 */
void signature_set_filename(Signature *ptr, const char *name);   /* src/ffi.rs:364-382 */
void signature_push_mh(Signature *ptr, const KmerMinHash *other); /* clones; src/ffi.rs:384-399 */
void signature_set_mh(Signature *ptr, const KmerMinHash *other);  /* clones; src/ffi.rs:401-416 */
SourmashStr signature_get_name(Signature *ptr);                  /* "" when unset; src/ffi.rs:418-431 */
SourmashStr signature_get_filename(Signature *ptr);              /* src/ffi.rs:433-446 */
SourmashStr signature_get_license(Signature *ptr);               /* src/ffi.rs:448-457 */
/* new owned clone of the first sketch (a default sketch if there is none).  src/ffi.rs:459-473 */
KmerMinHash *signature_first_mh(Signature *ptr);
/* owned array of owned clones; *size receives the count.  src/ffi.rs:504-522 */
KmerMinHash **signature_get_mhs(Signature *ptr, uintptr_t *size);
/* metadata and FIRST sketch equal.  src/ffi.rs:475-489 -> src/lib.rs:663-675 */
bool signature_eq(Signature *ptr, Signature *other);
/* compact JSON object, field order of src/lib.rs:546-565 and 79-100.  src/ffi.rs:491-501 */
SourmashStr signature_save_json(Signature *ptr);
/* compact JSON array of the signatures.  src/ffi.rs:524-534 */
SourmashStr signatures_save_buffer(Signature **ptr, uintptr_t size);
/* one Signature per stored sketch that passes the ksize (0 = any) / moltype (NULL = any, "dna" or
 * "protein", case-insensitive) filter; ignore_md5sum is accepted and ignored as in the
 * reference.  The path form reads `-` as stdin and inflates gzip input (src/file.rs:47-77).
 * src/ffi.rs:536-604 -> src/lib.rs:593-645 */
Signature **signatures_load_path(const char *ptr, bool ignore_md5sum, uintptr_t ksize,
                                 const char *select_moltype, uintptr_t *size);
Signature **signatures_load_buffer(const char *ptr, uintptr_t insize, bool ignore_md5sum,
                                   uintptr_t ksize, const char *select_moltype, uintptr_t *size);

/* ---- errors and strings: src/utils.rs:52-245 --------------------------------------------- */
void sourmash_init(void);
void sourmash_err_clear(void);
SourmashErrorCode sourmash_err_get_last_code(void);
SourmashStr sourmash_err_get_last_message(void); /* owned; "" when there is no error */
SourmashStr sourmash_err_get_backtrace(void);    /* always empty in this build */
void sourmash_str_free(SourmashStr *s);
/* the reference borrows `s` yet marks the result owned (utils.rs:226-233); this build returns an
 * owned COPY, so sourmash_str_free on it is always safe */
SourmashStr sourmash_str_from_cstr(const char *s);

#ifdef __cplusplus
}
#endif
#endif /* SOURMASH_B200_SOURMASH_H */
