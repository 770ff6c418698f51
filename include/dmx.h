/*
 * dmx.h -- C ABI of libdmx: the B200 (sm_100a) implementation of demuxlet's per-barcode
 * genotype-likelihood / doublet-pair scan.
 *
 * The reference has no plugin or FFI boundary for this path: the whole computation is a code
 * region of main() (/root/reference/cmd_cram_demuxlet.cpp:386-875) that consumes the in-memory
 * pileup `sc_dropseq_lib_t scl` (/root/reference/sc_drop_seq.h:34-58), the per-SNP genotype
 * probabilities `scl.snps[i].gps`, `gridAlpha`, `doublet_prior` and the three min-filters, and
 * prints four TSV files.  This header is the boundary a maintainer would cut there
 * (INTEGRATION.md shows the call site): flat arrays in, fixed-size records out, plain pointers and
 * sizes, no C++/CUDA/torch types.  Each entry point cites the reference lines it replaces.
 *
 * Conventions: every function returns 0 (DMX_OK) or a negative dmx_status and never throws or
 * aborts; the caller owns all host buffers (the library copies what it keeps unless
 * DMX_BORROW_INPUTS is set); a dmx_ctx is not thread-safe.  A context made by dmx_create is bound to
 * one CUDA device; dmx_create_sharded makes one that drives several GPUs of the host over disjoint
 * barcode ranges -- cells are independent in the reference (:576-875), cf. its --group-list sharding
 * (:68).  Processes that own one GPU each (torchrun) use one dmx_create context per process instead.
 */
#ifndef DMX_H
#define DMX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMX_ABI_VERSION 2
#define DMX_MAX_ALPHA 16

typedef enum {
  DMX_OK = 0,
  DMX_E_INVALID = -1,      /* bad argument (nv < 2, n_alpha < 2, null pointer, unsorted CSR ...) */
  DMX_E_CUDA = -2,         /* CUDA runtime error; dmx_last_error() has the text                   */
  DMX_E_NODEVICE = -3,     /* no usable sm_100 device: the library has NO CPU fallback           */
  DMX_E_STATE = -4,        /* call order: gp/pileup not set, run not done                        */
  DMX_E_NOMEM = -5,
  DMX_E_IO = -6,           /* output file cannot be created (:409, :535)                         */
  DMX_E_UNSUPPORTED = -7
} dmx_status;

/* dmx_config.flags */
#define DMX_EXACT_TIES    0x1u  /* re-score near-tied candidates with the reference's exact double
                                   arithmetic so BEST / SNG.1ST / DBL.1ST are bit-identical (default on
                                   through dmx_config_default) */
#define DMX_BORROW_INPUTS 0x2u  /* dmx_set_gp/dmx_set_pileup keep the caller's pointers (which must
                                   stay valid and unchanged until the ctx is destroyed or re-set)
                                   instead of copying; use with dmx_alloc_host for pinned transfers */
#define DMX_SKIP_SINGLE   0x8u  /* do not compute pass 1 (.single, :412-461); dmx_get_single then fails   */
#define DMX_TRUST_INPUTS  0x10u /* skip the host-side range check of ent_snp against nsnps in dmx_upload  */
#define DMX_HOST_FINALISE 0x20u /* re-score near ties on the host (libm log) even when the device re-scorer
                                   has verified itself bit-identical to the host libm                    */
#define DMX_FP32_PAIRS    0x4u  /* pair scan accumulates in fp32 (faster, ~1e-6 LLK error) instead
                                   of fp64; singlet, pass-1 and all re-scored values stay fp64      */

typedef struct {
  int32_t nv;                  /* number of VCF samples, >= 2 (:180)                   */
  int32_t n_alpha;             /* alpha grid size, >= 2, <= DMX_MAX_ALPHA (:348)                   */
  const double* alpha;         /* gridAlpha (:54, :78-90); alpha[0] is the "singlet" slot (Q3)     */
  double doublet_prior;        /* --doublet-prior (:56)                                            */
  int32_t min_total, min_uniq, min_snp;   /* --min-total/--min-uniq/--min-snp (:69-71, :581)       */
  int32_t device;              /* CUDA device ordinal                                              */
  uint32_t flags;
  int32_t n_host_threads;      /* threads of the tie finaliser; 0 = hardware concurrency           */
} dmx_config;

/* one row of <out>.best (:816-874), fixed size POD */
typedef struct {
  int32_t status;              /* 0 row present; 1 dropped by the min-filters (:581); 2 no SNP (:592) */
  int32_t n_snp;               /* N.SNP = covered SNPs incl. all-allele-2 ones (:835)               */
  int32_t sng1, sng2;          /* iSing1, iSing2 (:746-758)                                         */
  int32_t dbl1, dbl2, alpha_idx; /* jBest, kBest, alphaBest (:799-814)                              */
  int32_t call;                /* 0 DBL, 1 SNG, 2 AMB (:837-857)                                    */
  double sng_llk1, sng_llk2, sng_llk0;                 /* :816-820 */
  double llk12, llk1, llk2, llk10, llk20, llk00;       /* :821-826 */
  double prb_dbl, prb_sng1;                            /* :827-828 */
  double max_llk, sum_single, sum_double;              /* :713-734 */
  int32_t n_rescored;          /* hypotheses the tie finaliser re-scored exactly for this cell      */
  uint32_t exact_mask;         /* DMX_X_* : which LLK fields carry reference-exact bits             */
} dmx_best_rec;

#define DMX_X_SNG1  0x01u
#define DMX_X_SNG2  0x02u
#define DMX_X_LLK12 0x04u
#define DMX_X_LLK1  0x08u
#define DMX_X_LLK2  0x10u
#define DMX_X_FULLCELL 0x80u   /* whole cell recomputed exactly (candidate list overflowed)         */

typedef struct {
  double h2d_ms, prep_ms, kernel_ms, d2h_ms, finalise_ms, total_ms;  /* last dmx_run / dmx_compute */
  int64_t h2d_bytes, d2h_bytes;
  int64_t n_entries, n_reads, n_tasks;
  int64_t gpu_launches;        /* kernels launched by the last run                                 */
  int64_t n_rescored;          /* hypotheses re-scored exactly                                     */
  int64_t n_fullcell;          /* cells whose candidate list overflowed (dense re-scan)            */
  int64_t n_rescored_gpu;      /* ... of n_rescored, by the device re-scorer                       */
  int32_t gpu_exact;           /* 1: device log() verified bit-identical to the host libm          */
  int32_t pad_;
  double algo_bytes;           /* SURVEY 8(d) gather-model bytes of the pair kernel, last run      */
  double pair_kernel_ms;       /* CUDA-event time of the pair-scan kernel(s) alone                 */
} dmx_stats;

typedef struct dmx_ctx dmx_ctx;

/* defaults of the reference CLI: doublet_prior 0.5 (:29), filters 0 (:31-33), alpha {0, 0.5}
 * must still be supplied by the caller; flags = DMX_EXACT_TIES */
void dmx_config_default(dmx_config* cfg);

/* number of CUDA devices this library can run on (compute capability 10.x); 0 if none */
int dmx_device_count(void);

int dmx_create(dmx_ctx** out, const dmx_config* cfg);
void dmx_destroy(dmx_ctx* ctx);

/* ---- multi-GPU (the reference's own parallel mode is --group-list over one barcode set, :68, :92-99) ----
 * One context for n_devices GPUs of this host (cfg->device is ignored; a device may be listed more than once).
 * dmx_set_pileup splits the barcodes into n_devices contiguous cell-id ranges with (nearly) equal numbers of pileup
 * entries; every device holds the whole genotype-probability matrix; dmx_run / dmx_upload / dmx_compute drive the
 * devices from one host thread each (no data-path collective: barcodes are independent), and the per-barcode records
 * of all shards land in one set of host arrays, so every getter and dmx_write_outputs behave exactly as for a
 * single-device context and produce the same bytes.  cfg->n_host_threads is the total, shared between the shards. */
int dmx_create_sharded(dmx_ctx** out, const dmx_config* cfg, int32_t n_devices, const int32_t* devices);
int dmx_shard_count(const dmx_ctx* ctx);             /* 1 for a dmx_create context */
/* barcode range [cell_lo, cell_hi) and device of one shard (valid after dmx_set_pileup) */
int dmx_shard_range(const dmx_ctx* ctx, int32_t shard, int64_t* cell_lo, int64_t* cell_hi, int32_t* device);
/* dmx_get_stats of one shard (times of that device alone; the context's own stats carry max times / summed counts) */
int dmx_shard_stats(dmx_ctx* ctx, int32_t shard, dmx_stats* out);

/* gp[nsnps][nv][3], genotype fastest: exactly the floats the reference widens into
 * scl.snps[i].gps (:181-184, :227-231).  Also computes the cohort mean gp0s (:391-401) on device. */
int dmx_set_gp(dmx_ctx* ctx, const float* gp, int64_t nsnps);

/* CSR flattening of sc_dropseq_lib_t (sc_drop_seq.h:34-58):
 *   cell_ptr[ncells+1]  into entries, cells in cell-id (first-seen) order
 *   ent_snp[E]          SNP id, strictly ascending inside a cell (std::map<int32_t,..> order)
 *   ent_ptr[E+1]        into reads
 *   reads[R]            allele<<8 | bq, in UMI-string order inside an entry (std::map<std::string,..>),
 *                       allele 0 ref / 1 alt / 2 other (:322), bq already capped (:323)
 *   totl/pass/uniq[ncells]  cell_totl_reads / cell_pass_reads / cell_uniq_reads */
int dmx_set_pileup(dmx_ctx* ctx, int64_t ncells, const int64_t* cell_ptr, const int32_t* ent_snp,
                   const int64_t* ent_ptr, const uint16_t* reads, const int32_t* totl,
                   const int32_t* pass, const int32_t* uniq);

/* dmx_run = dmx_upload (host->device copies) + dmx_compute (kernels, record read-back, tie finaliser).
 * Replaces cmd_cram_demuxlet.cpp:412-461 (pass 1) and :539-734, :746-758, :799-828 (pass 2). */
int dmx_upload(dmx_ctx* ctx);
int dmx_compute(dmx_ctx* ctx);
int dmx_run(dmx_ctx* ctx);

/* results, valid after dmx_run/dmx_compute; arrays are in cell-id order */
int dmx_get_best(dmx_ctx* ctx, dmx_best_rec* out /* [ncells] */);
/* pass 1 (:455-459): llk[ncells][nv] = llks, llk0[ncells] = llk0s */
int dmx_get_single(dmx_ctx* ctx, double* llk, double* llk0);
/* :760-769: llk_j00[ncells][nv] = llksAB[j][0][0]; llk00[ncells][n_alpha] = llks00 */
int dmx_get_sing2(dmx_ctx* ctx, double* llk_j00, double* llk00);
/* dense llksAB[cell][j][k][n] for cells [cell_lo, cell_hi) (what --write-pair prints, :772-797);
 * re-runs the pair scan for that range with dense stores */
int dmx_get_pair(dmx_ctx* ctx, int64_t cell_lo, int64_t cell_hi, double* llkAB);

/* exact (reference-arithmetic) value of one hypothesis, as the tie finaliser computes it */
int dmx_rescore(dmx_ctx* ctx, int64_t cell, int32_t j, int32_t k, int32_t n, double* out);

/* writers with the reference's formats: <prefix>.single (:470,:506), .sing2 (:533,:760),
 * .best (:571,:830-874) and, if write_pair, .pair (:570,:774-792).  Rows in barcode
 * lexicographic order (std::map iteration, :472,:576). */
int dmx_write_outputs(dmx_ctx* ctx, const char* const* barcodes /* [ncells] */,
                      const char* const* samples /* [nv] */, const char* prefix, int write_pair);

/* host-only utility (needs no device, no ctx): llksAB[j][k][n] of one cell in the reference's exact arithmetic
 * (cmd_cram_demuxlet.cpp:597-684), from the same flat arrays dmx_set_gp / dmx_set_pileup take */
int dmx_exact_llk(int32_t nv, int32_t n_alpha, const double* alpha, const float* gp, const int64_t* cell_ptr,
                  const int32_t* ent_snp, const int64_t* ent_ptr, const uint16_t* reads, int64_t cell, int32_t j,
                  int32_t k, int32_t n, double* out);

int dmx_get_stats(dmx_ctx* ctx, dmx_stats* out);

/* ---- pileup packer (host only) -------------------------------------------------------------------
 * Sort-based replacement of sc_dropseq_lib_t::add_read (sc_drop_seq.cpp:34-77) followed by the flattening of the
 * container: n reads in arrival order -- cell id (add_cell order, sc_drop_seq.cpp:21-32), SNP id, UMI string
 * (umi_buf[umi_off[i] .. umi_off[i+1])), allele, base quality -- in; the CSR arrays dmx_set_pileup takes plus
 * cell_pass_reads / cell_uniq_reads out.  The first read of a (snp, cell, umi) keeps its allele and quality, later
 * ones only count; SNPs ascending inside a cell, UMIs in std::string order inside an entry. */
typedef struct dmx_pileup dmx_pileup;
int dmx_pack_reads(int64_t n, const int32_t* cell, const int32_t* snp, const uint8_t* allele, const uint8_t* bq,
                   const char* umi_buf, const int64_t* umi_off, int64_t ncells, int32_t n_threads, dmx_pileup** out);
int dmx_pileup_sizes(const dmx_pileup* p, int64_t* ncells, int64_t* n_entries, int64_t* n_reads);
/* copies into caller buffers sized from dmx_pileup_sizes: cell_ptr[ncells+1], ent_snp[E], ent_ptr[E+1], reads[R],
 * pass[ncells], uniq[ncells]; any pointer may be null */
int dmx_pileup_copy(const dmx_pileup* p, int64_t* cell_ptr, int32_t* ent_snp, int64_t* ent_ptr, uint16_t* reads,
                    int32_t* pass, int32_t* uniq);
void dmx_pileup_free(dmx_pileup* p);

/* ---- VCF field -> genotype-probability matrix (device) ----------------------------------------------
 * Replaces, per biallelic diploid VCF record, BCFFilteredReader::parse_posteriors
 * (bcf_filtered_reader.cpp:360-454) with parse_genotypes (:186-242) for --field GT and the ten-iteration EM of
 * parse_likelihoods (:244-320) for --field PL, i.e. what cmd_cram_demuxlet.cpp:153,:211 call and :227-231 widen
 * into scl.snps[i].gps.  The output is bit-identical to the reference's floats.
 *   DMX_FIELD_GT  in = int8[nsnps][nv][2]   allele indices as bcf_gt_allele() decodes them: 0 ref, 1 alt, < 0 missing
 *                                           (minDP / minGQ masking, :203-226, applied by the caller)
 *   DMX_FIELD_GP  in = float[nsnps][nv][3]  the FORMAT floats as bcf_get_format_float returns them
 *   DMX_FIELD_PL  in = int32[nsnps][nv][3]  the FORMAT integers (missing = INT32_MIN reads as phred 255, PhredHelper.h:40)
 * gt_error is --geno-error (cmd_cram_demuxlet.cpp:30, :153); PL ignores it, as the reference does.
 * All nv samples are the selected ones (sm_icols), ploidy 2 (bcf_filtered_reader.cpp:152-153). */
#define DMX_FIELD_GT 0
#define DMX_FIELD_GP 1
#define DMX_FIELD_PL 2

typedef struct {
  double h2d_ms, kernel_ms, d2h_ms, total_ms;   /* kernel_ms: CUDA events around the field kernel(s) */
  int64_t h2d_bytes, d2h_bytes, gpu_launches;
  double algo_bytes;                            /* bytes every launch must move: input + 12 * nv per SNP */
} dmx_field_stats;

/* host buffers in and out: gp_out[nsnps][nv][3] is what dmx_set_gp takes; af_out[nsnps] (nullable) = get_af(1)
 * (bcf_filtered_reader.h:177-179; GT and PL).  SNP chunks are copied, transformed and read back on two streams;
 * pass dmx_alloc_host buffers for overlapped copies.  stats may be null. */
int dmx_build_gp(int32_t device, int32_t field, int32_t nv, int64_t nsnps, const void* in, double gt_error,
                 float* gp_out, double* af_out, dmx_field_stats* stats);
/* the same on DEVICE pointers (input, output rows of out_stride floats, nullable af), enqueued on `stream`
 * (a cudaStream_t, or null); returns after the launch */
int dmx_build_gp_device(int32_t device, int32_t field, int32_t nv, int64_t nsnps, const void* d_in, double gt_error,
                        float* d_out, int64_t out_stride, double* d_af, void* stream);
/* dmx_set_gp from a VCF field: builds the matrix into the context's device copy (no second upload) and keeps the
 * read-back on the host for the tie finaliser */
int dmx_set_field(dmx_ctx* ctx, int32_t field, const void* in, int64_t nsnps, double gt_error);

/* the reference's per-record variant filter on genotype counts (BCFFilteredReader::passed_vfilter,
 * bcf_filtered_reader.cpp:540-573, which demuxlet always runs on GT: SURVEY Q9) with the two thresholds its command
 * line exposes, --min-mac (default 1, cmd_cram_demuxlet.cpp:24,:47) and --min-callrate (default 0.5, :25,:48):
 * pass_out[s] = 1 when an / (2 nv) >= min_callrate and both allele counts are >= min_mac.  gt as DMX_FIELD_GT. */
int dmx_site_filter(int32_t device, int32_t nv, int64_t nsnps, const int8_t* gt, int32_t min_mac, double min_callrate,
                    uint8_t* pass_out);

/* ---- VCF / BCF reader (host only, no htslib; SURVEY 8f N4, VCF half) --------------------------------------------
 * Reads a VCF (plain, gzip or bgzip text) or a BCF2 file and delivers per biallelic record what the reference obtains from htslib's
 * bcf_read / vcf_parse1 (bcf_chunked_reader.cpp:265-305) and bcf_get_genotypes / bcf_get_format_float /
 * bcf_get_format_int32 (bcf_filtered_reader.cpp:194, :250, :408): the arrays dmx_site_filter and dmx_build_gp take.
 * Records with an allele count other than 2 are skipped and counted (maxAlleles = 2, cmd_cram_demuxlet.cpp:26).
 * Missing values: GT -1, GP NaN, PL INT32_MIN.  All samples are selected. */
typedef struct dmx_vcf dmx_vcf;
int dmx_vcf_read(const char* path, int32_t field /* DMX_FIELD_*: which FORMAT key besides GT */, dmx_vcf** out);
/* the same with the FORMAT key named: the reference reads ANY --field other than GT and PL as genotype posteriors
 * like GP (parse_posteriors' last branch, bcf_filtered_reader.cpp:408-452): field = DMX_FIELD_GP, key = that name */
int dmx_vcf_read_key(const char* path, int32_t field, const char* key /* null: "GP" / "PL" */, dmx_vcf** out);
/* --sm / --sm-list (bcf_filtered_reader.cpp:95-121): keep only the named samples, in the byte-lexicographic order of
 * their names (the reference's std::set), duplicates merged; DMX_E_INVALID if a name is not in the header */
int dmx_vcf_select(dmx_vcf* v, int32_t n, const char* const* names);
int dmx_vcf_sizes(const dmx_vcf* v, int32_t* nv, int64_t* nrec, int64_t* n_skipped, int32_t* n_chrom);
const char* dmx_vcf_sample(const dmx_vcf* v, int32_t i);        /* sample id of column i (get_sample_id_at)       */
const char* dmx_vcf_chrom(const dmx_vcf* v, int32_t rid);       /* chromosome name of rid (first-seen order)       */
/* any pointer may be null: gt int8[nrec][nv][2]; field_values = gt again (GT), float[nrec][nv][3] (GP) or
 * int32[nrec][nv][3] (PL); rid / pos (0-based) [nrec]; first base of REF / ALT [nrec] (add_snp, :230) */
int dmx_vcf_copy(const dmx_vcf* v, int8_t* gt, void* field_values, int32_t* rid, int32_t* pos, char* ref, char* alt);
/* flags[nrec]: 1 where the record's FORMAT does not have the requested key (its values are all missing).  The reference
 * fails on such a record only when it reaches parse_posteriors, i.e. after the genotype filter and inside the span of
 * the BAM (cmd_cram_demuxlet.cpp:153,:211-212): check the records that survive dmx_site_filter */
int dmx_vcf_no_field(const dmx_vcf* v, uint8_t* flags);
void dmx_vcf_free(dmx_vcf* v);
const char* dmx_vcf_last_error(void);

/* ---- BAM sweep (host only, no htslib; SURVEY 8f N4, BAM half; `path` may also be SAM text with its header) --------
 * The reference's walk over the position-sorted BAM next to the SNP list (cmd_cram_demuxlet.cpp:186-338): alignment
 * filter (sam_filtered_reader.cpp:284-296), barcode / UMI tags, RD.TOTL, and per SNP of the reference's SNP buffer
 * (:202-237, restated exactly: in practice the SNPs under the read) the base and quality at the SNP position through
 * the CIGAR (hts_utils.cpp:279-359), thresholds and the allele code (:304-325).  Emits one
 * tuple per scl.add_read call, in call order -- exactly the arrays dmx_pack_reads takes --, the barcodes in add_cell
 * order and cell_totl_reads.  SNPs: sorted by (rid, pos), ids = their index (the order of scl.add_snp); chroms[rid] are
 * the VCF's chromosome names. */
typedef struct {
  const char* tag_group;       /* --tag-group, "CB" (:12); "" = one group "." (:241-243)             */
  const char* tag_umi;         /* --tag-UMI, "UB" (:13); "": "." + hex(rand()) per read like :274-276 (the reference
                                  never seeds rand(): glibc's default sequence, reproduced from a private state) */
  int32_t min_mq;              /* --min-MQ 20 (:19)                                                   */
  uint32_t exclude_flag;       /* --excl-flag 0x0f04 (:18)                                            */
  int32_t min_bq, cap_bq, min_td;   /* --min-BQ 13, --cap-BQ 40, --min-TD 0 (:14-16)                 */
  int64_t n_groups;            /* --group-list (:68, :131-140): barcodes to keep, 0 = all             */
  const char* const* groups;
} dmx_sweep_opts;
typedef struct dmx_sweep dmx_sweep;
void dmx_sweep_opts_default(dmx_sweep_opts* o);
int dmx_sweep_bam(const char* path, const dmx_sweep_opts* o, int32_t n_chrom, const char* const* chroms, int64_t nsnps,
                  const int32_t* snp_rid, const int32_t* snp_pos /* 0-based */, const char* snp_ref, const char* snp_alt,
                  dmx_sweep** out);
int dmx_sweep_sizes(const dmx_sweep* s, int64_t* ncells, int64_t* nreads, int64_t* umi_bytes);
int dmx_sweep_counts(const dmx_sweep* s, int64_t* n_records, int64_t* n_filtered, int64_t* n_off_vcf, int64_t* n_skip_bcd);
/* cell / snp / allele / bq [nreads], umi_buf[umi_bytes], umi_off[nreads+1], totl[ncells]; any pointer may be null */
int dmx_sweep_copy(const dmx_sweep* s, int32_t* cell, int32_t* snp, uint8_t* allele, uint8_t* bq, char* umi_buf,
                   int64_t* umi_off, int32_t* totl);
const char* dmx_sweep_barcode(const dmx_sweep* s, int64_t i);
void dmx_sweep_free(dmx_sweep* s);
const char* dmx_sweep_last_error(void);

/* pinned host memory for DMX_BORROW_INPUTS callers */
void* dmx_alloc_host(int64_t bytes);
void dmx_free_host(void* p);

const char* dmx_strerror(int status);
const char* dmx_last_error(void);      /* text of the last DMX_E_CUDA on this thread */
int dmx_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
