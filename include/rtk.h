/* rtk.h — C ABI of librtk_b200.so: the B200-native replacement for Ratatosk's per-read
 * correction hot path.
 *
 * The reference (DecodeGenetics/Ratatosk) has no FFI; its seam is a set of C++ free
 * functions called from the worker loop of search() (src/Ratatosk.cpp:808-867).  Each entry
 * point below names the reference interface it replaces.  Plain pointers and sizes only;
 * every function returns 0 on success or a negative RTK_E* code (never exits, unlike the
 * reference's exit(1)); rtk_last_error() gives the message for the calling thread.
 *
 * Threading: a context serialises work on its own CUDA stream; distinct contexts (one per
 * GPU) may be used concurrently from distinct host threads.
 */
#ifndef RTK_H
#define RTK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTK_OK 0
#define RTK_EINVAL (-1)
#define RTK_ECUDA (-2)
#define RTK_EIO (-3)
#define RTK_ENOMEM (-4)
#define RTK_ENOGRAPH (-5)
#define RTK_EUNSUPPORTED (-6)

typedef struct rtk_ctx rtk_ctx;          /* one GPU: stream, device graph, scratch */
typedef struct rtk_host_graph rtk_host_graph; /* flat graph slab in host memory */

/* POD mirror of the hot-path fields of Correct_Opt (src/Common.hpp:16-158). */
typedef struct rtk_opt {
    uint32_t k;                    /* Correct_Opt::k for this pass (31 pass 1 / 63 pass 2) */
    uint32_t insert_sz;            /* 500 */
    uint32_t min_cov_vertices;     /* 2 */
    uint32_t max_km_cov;           /* 128 (max'ed with the graph's top-0.1% coverage) */
    uint32_t max_len_weak_region1; /* 1000 */
    uint32_t max_len_weak_region2; /* 5000 */
    uint32_t nb_correction_rounds; /* 1 */
    int32_t out_qual;              /* 1 */
    int32_t max_qual;              /* 40 */
    int32_t trim_qual;             /* 0 */
    double weak_region_len_factor; /* 0.25 */
    double large_k_factor;         /* 1.5 */
    double min_score;              /* 0.0 */
    double min_confidence_snp_corr;/* 0.9 */
    uint32_t force_unres_snp_corr; /* 0 */
    uint32_t reserved;             /* bit 0: rtk_color_long_reads keeps all reads where the reference would subsample */
} rtk_opt;

void rtk_opt_default(rtk_opt* opt, int pass); /* pass 1: k=31, pass 2: k=63 */

/* A k-mer hit: const_UnitigMap<UnitigData> with len==1 (Bifrost/src/UnitigMap.hpp:58-66). */
typedef struct rtk_hit {
    uint32_t pos;     /* position in the query read */
    uint32_t unitig;  /* unitig id (order of the index FASTA) */
    uint32_t dist;    /* k-mer offset in the unitig, forward orientation */
    uint32_t strand;  /* 1: read k-mer equals the unitig's forward k-mer */
} rtk_hit;

typedef struct rtk_graph_info {
    uint32_t k;
    uint64_t n_unitigs, n_kmers, pool_bases, n_buckets, n_gsets, slab_bytes, max_km_cov_graph;
} rtk_graph_info;

int rtk_version(void);
const char* rtk_last_error(void);
void rtk_free(void* p); /* releases any buffer handed out by this library */

/* ---- graph (replaces CompactedDBG::read + readGraphData, src/Ratatosk.cpp:1087-1089) ---- */
int rtk_graph_load(const char* fasta_path, const char* rtsk_path, int k, rtk_host_graph** out);
/* unitigs only (no colours / flags): synthetic graphs for benchmarks and tests */
int rtk_graph_from_unitigs(int k, uint64_t n, const char* const* seqs, rtk_host_graph** out);
void rtk_graph_free(rtk_host_graph* g);
int rtk_graph_get_info(const rtk_host_graph* g, rtk_graph_info* info);
const void* rtk_graph_slab(const rtk_host_graph* g, uint64_t* bytes);
/* Flat cache file = the slab itself (versioned header + 256-byte aligned sections, offsets relative to the file start):
 * rtk_graph_open maps it read-only and uses it in place (no parse, no flatten; the page cache is shared by the processes
 * of a node).  Replaces readGraphData's per-unitig find + Roaring deserialisation (src/Graph.cpp:722-801) on every run.
 * rtk_graph_save writes atomically (temporary file + rename).  rtk_graph_load_cached = rtk_graph_load through the cache
 * file cache_path (NULL: "<rtsk_path>.k<k>.rtkflat"): opened when it is at least as recent as the index files, else the
 * index is parsed and the cache (re)written; *from_cache (optional) tells which happened. */
int rtk_graph_save(const rtk_host_graph* g, const char* path);
int rtk_graph_open(const char* path, rtk_host_graph** out);
int rtk_graph_load_cached(const char* fasta_path, const char* rtsk_path, int k, const char* cache_path, rtk_host_graph** out,
                          int* from_cache);
/* per-unitig accessors over the host slab (tests / host-side callers) */
int rtk_graph_unitig_seq(const rtk_host_graph* g, uint32_t unitig, char* buf, uint64_t cap, uint64_t* len);
int rtk_graph_unitig_words(const rtk_host_graph* g, uint32_t unitig, uint64_t* kmcov, uint64_t* shared, uint32_t adj[8]);
int rtk_graph_unitig_colors(const rtk_host_graph* g, uint32_t unitig, const uint32_t** gids, uint64_t* n_g,
                            const uint32_t** lids, uint64_t* n_l);

/* ---- device context ---- */
int rtk_ctx_create(int device, rtk_ctx** out);
void rtk_ctx_destroy(rtk_ctx* ctx);
/* A second context on the same device that SHARES the parent's resident graph (own stream and scratch).  One
 * context serves one host thread at a time; forks let several host threads issue batches concurrently, the way
 * the reference's correction threads share one CompactedDBG (src/Correction.cpp worker pool).  The parent must
 * outlive its forks; re-fork after rtk_graph_upload / rtk_graph_adopt_device on the parent. */
int rtk_ctx_fork(const rtk_ctx* parent, rtk_ctx** out);
int rtk_graph_upload(rtk_ctx* ctx, const rtk_host_graph* g);               /* H2D copy of the slab */
/* slab already resident on this device (e.g. after an NCCL broadcast done by the caller) */
int rtk_graph_adopt_device(rtk_ctx* ctx, const void* dev_slab, uint64_t bytes);
int rtk_ctx_sync(rtk_ctx* ctx);

/* ---- K1: CompactedDBG::searchSequence (Bifrost/src/Search.tcc:526-768) ----
 * Batch of n_reads upper-case reads, read i = seq_pool[seq_off[i], seq_off[i+1]).
 * flags: bit0 exact, bit1 insertion, bit2 deletion, bit3 substitution, bit4 or_exclusive_match.
 * Supported combinations are the two the reference uses (src/Graph.cpp:97,193): exact only,
 * or any inexact subset without exact.  Output: for read i the hits hits[hit_off[i],hit_off[i+1])
 * in the order the reference's v_um holds them.  *hits / *hit_off are library-allocated
 * (rtk_free).  stats (optional, 4 x u64): probes, raw hits, kernel ns, total ns. */
#define RTK_SEARCH_EXACT 1u
#define RTK_SEARCH_INS 2u
#define RTK_SEARCH_DEL 4u
#define RTK_SEARCH_SUBST 8u
#define RTK_SEARCH_OR_EXCL 16u
int rtk_search_sequence(rtk_ctx* ctx, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                        uint32_t flags, rtk_hit** hits, uint64_t** hit_off, uint64_t* stats);

/* Device-resident variant for measurement: reads already in HBM (dev_seq/dev_seq_off), raw
 * labelled hits stay in HBM; returns the number of probes and raw hits.  Timed by bench.py. */
int rtk_k1_sweep_device(rtk_ctx* ctx, uint32_t n_reads, const char* dev_seq, const uint64_t* dev_seq_off,
                        const uint64_t* host_seq_off, uint32_t flags, uint64_t* n_probes, uint64_t* n_raw_hits,
                        float* kernel_ms);

/* ---- getSeeds (src/Graph.cpp:3-482): solid and weak anchors of each read ---- */
typedef struct rtk_seeds {
    rtk_hit* solid;        /* concatenated per read */
    uint64_t* solid_off;   /* n_reads+1 */
    rtk_hit* weak;
    uint64_t* weak_off;    /* n_reads+1 */
} rtk_seeds;
int rtk_get_seeds(rtk_ctx* ctx, const rtk_opt* opt, int pass, uint32_t n_reads, const char* seq_pool,
                  const uint64_t* seq_off, rtk_seeds* out, uint64_t* stats);
void rtk_seeds_free(rtk_seeds* s);

/* ---- K4: edlibAlign distance modes (src/edlib.cpp:141-296) with the IUPAC equality table
 * (src/Common.hpp:262-276).  Pair i: query q_pool[q_off[i],q_off[i+1]), target likewise.
 * mode[i]: 0 NW, 1 SHW, 2 HW.  kmax[i]: edlib's k (-1 = unbounded).
 * Out: dist[i] (-1 if > kmax), and all end locations with that distance:
 * end_loc[end_off[i], end_off[i+1]) ascending (library-allocated). */
int rtk_edlib_batch(rtk_ctx* ctx, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool,
                    const uint64_t* t_off, const uint8_t* mode, const int32_t* kmax, int32_t* dist,
                    int32_t** end_loc, uint64_t** end_off, uint64_t* stats);

/* ---- K5: edlibAlign with EDLIB_TASK_PATH (src/edlib.cpp:262-279, obtainAlignmentTraceback :945-1134) ----
 * mode[i]: 0 NW, 1 SHW.  Out per pair: dist, the (first) end column, and the edit operations of the path
 * (0 match, 1 query base unaligned, 2 target base unaligned, 3 mismatch) in ops[ops_off[i], ops_off[i+1]).
 * Pairs at or above edlib's 1 MiB switch go through its divide-and-conquer (split row and size switch
 * reproduced, src/edlib.cpp:1191-1356).  flags[i] is reserved (always 0). */
int rtk_edlib_path_batch(rtk_ctx* ctx, uint32_t n, const char* q_pool, const uint64_t* q_off, const char* t_pool,
                         const uint64_t* t_off, const uint8_t* mode, int32_t* dist, int32_t* end_loc, uint8_t** ops,
                         uint64_t** ops_off, uint8_t* flags, uint64_t* stats);

/* ---- K2/K3 + K4: exploreSubGraph / exploreSubGraphLong (src/GraphTraversal.cpp:456-587, :589-720) ----
 * Bounded DFS from a start unitig towards an optional target unitig with colour-set threshold intersection
 * (getNumberSharedPairID >= min_cov against `pids` = WeightsPairID::all_pids), edge-flag test, and scoring of
 * every candidate by edit distance against the read window (getScorePath :867-909); returns, per call, the
 * equal-best terminal and non-terminal paths in discovery order and the best / second-best scores. */
typedef struct rtk_subgraph_call_t {
    uint32_t start_unitig, start_strand;
    uint32_t end_unitig, end_strand, end_dist; /* end_unitig = 0xFFFFFFFF: no target */
    uint32_t level;                            /* DFS depth below the start; the reference passes level-1 = 3 */
    uint32_t max_len_path;
    uint32_t ref_len;
    uint64_t ref_off;                          /* read window = ref_pool[ref_off, ref_off+ref_len) */
    uint64_t pid_off;                          /* pids = pid_pool[pid_off, pid_off+pid_len), sorted */
    uint32_t pid_len;                          /* 0 = no colour filter (all_pids.isEmpty()) */
    uint32_t min_cov;                          /* Correct_Opt::min_cov_vertices */
    uint32_t max_len_subpath;                  /* 0: exploreSubGraph, bursts bounded by `level` unitigs (pass 1).  > 0: exploreSubGraphLong
                                                  (src/GraphTraversal.cpp:589-720, pass 2): a partial path is expanded while it spells fewer than
                                                  max_len_subpath = k * large_k_factor bases; `level` is ignored */
    uint32_t reserved;
} rtk_subgraph_call_t;

typedef struct rtk_path_node {
    uint32_t unitig, strand, dist, len;        /* const_UnitigMap of one path vertex (Path::toVector) */
} rtk_path_node;

typedef struct rtk_subgraph_out {
    double* scores;        /* 4 per call: best terminal, 2nd terminal, best non-terminal, 2nd non-terminal */
    uint64_t* path_off;    /* n_calls+1: paths of call i = [path_off[i], path_off[i+1]) */
    uint32_t* n_terminal;  /* per call: the first n_terminal of its paths are terminal */
    uint64_t* node_off;    /* n_paths+1 */
    rtk_path_node* nodes;
    uint32_t* path_len;    /* per path: spelled length */
    int32_t* path_ed;      /* per path: edit distance behind its score */
} rtk_subgraph_out;

int rtk_explore_subgraph_batch(rtk_ctx* ctx, uint32_t n_calls, const rtk_subgraph_call_t* calls, const char* ref_pool,
                               uint64_t ref_bytes, const uint32_t* pid_pool, uint64_t n_pids, double weak_region_len_factor,
                               rtk_subgraph_out* out, uint64_t* stats);
void rtk_subgraph_out_free(rtk_subgraph_out* out);

/* ---- explorePathsBFS2 / explorePathsBFS (src/GraphTraversal.cpp:212-454, :3-210) ----
 * Best path between two anchors (um_e != NULL) or from one anchor into the open end of a read window
 * (um_e == NULL), within +-weak_region_len_factor of the window length: queue of partial paths, one
 * exploreSubGraph burst (K2/K3/K4) per pop, selectors (K4), path qualities (K5), fixRepeats.  Anchors are
 * getSeeds hits (len 1).  Output: the path's vertices (0 = none found), its per-base quality string and
 * spelled length; buffers are library-allocated (rtk_free). */
int rtk_explore_paths(rtk_ctx* ctx, const rtk_opt* opt, const rtk_hit* um_s, const rtk_hit* um_e, const char* ref,
                      uint32_t ref_len, const uint32_t* pids, uint32_t n_pids, rtk_path_node** nodes, uint32_t* n_nodes,
                      char** qual, uint32_t* path_len);

/* ---- device-resident region engine: extractSemiWeakPaths (src/Correction.cpp:3-157) with explorePathsBFS2 / explorePathsBFS
 * (src/GraphTraversal.cpp:212-454, :3-210), exploreSubGraph(Long) (:456-720), getScorePath (:722-772, :867-909) and the
 * selectors (src/Alignment.cpp:3-147, :967-1015) chained on the GPU, one warp per call, no host round trips.  A call is the
 * path search of one weak region: from the left anchor `start` over the weak anchors of the region to the right anchor
 * `end` (or to the open end of the read).  Positions are read positions as in the reference.  Output per call: status 0 =
 * the path reached the end (paths.first), 1 = it dead-ended at a weak anchor (paths.second), 2 = the engine declined the
 * call (`bail` says why: fixRepeats on short-cycle unitigs, the 512 / 1024 list collapses, alignments above edlib's 1 MiB
 * traceback switch, scratch capacity) and the caller must use rtk_explore_paths-level entries instead.  rtk_correct_batch
 * uses this engine internally and falls back per call. */
typedef struct rtk_region_call_t {
    uint64_t win_off;        /* window = win_pool[win_off, +win_len) = s[start_pos, pos_um_solid2 + k) */
    uint64_t weak_off;       /* weak anchors from i_weak on: weak_pool[weak_off, +n_weak), ascending pos */
    uint64_t pid_off;        /* WeightsPairID::all_pids, sorted: pid_pool[pid_off, +pid_len) */
    uint32_t win_len;
    uint32_t n_weak;
    uint32_t pid_len;
    uint32_t start_pos;      /* um_start.first */
    uint32_t start_unitig, start_dist, start_strand;
    uint32_t has_end;
    uint32_t end_pos;        /* pos_um_solid2 when has_end */
    uint32_t end_unitig, end_dist, end_strand;
    uint32_t s_len;          /* read length (open end: pos_um_solid2 = s_len - k) */
    uint32_t reserved;       /* bit 0: follow dead ends like the `correct` lambda (src/Correction.cpp:619-651): best prefix alignment of
                                the dead-end path, restart from the next weak anchor behind it; one segment per restart */
} rtk_region_call_t;

/* one extractSemiWeakPaths call of a region (a region has several when dead ends are followed) */
typedef struct rtk_region_seg_t {
    uint32_t status;         /* 0 complete path, 1 dead-end path */
    uint32_t start_weak;     /* index (in the call's weak list) of the weak anchor this segment starts from; 0xFFFFFFFF: the left anchor */
    uint32_t n_nodes, len;
    uint64_t node_off, str_off;
    int32_t shw_dist, shw_first_end;   /* dead ends followed: SHW distance / first end location of the path against the rest of the window */
} rtk_region_seg_t;

typedef struct rtk_region_result_t {
    uint32_t status, bail;
    uint32_t n_nodes, len;   /* path vertices / spelled length */
    uint64_t node_off;       /* nodes[node_off, +n_nodes) */
    uint64_t str_off;        /* chars[str_off, +len) = spelled path; chars[str_off + pad8(len), +len) = its quality string */
    uint32_t n_hops, n_pops, n_cands, n_aligns;   /* work done: BFS calls, queue pops, candidates scored, alignments */
    uint64_t seg_off;        /* segs[seg_off, +n_segs): every extractSemiWeakPaths call of the region, in order; the fields above */
    uint32_t n_segs;         /* describe the last one */
    uint32_t reserved;       /* work done, continued: DP cells swept by the call's alignments / 1024 (saturating) */
} rtk_region_result_t;

typedef struct rtk_region_out {
    rtk_region_result_t* results;   /* n_calls */
    rtk_path_node* nodes;
    char* chars;
    rtk_region_seg_t* segs;
    uint64_t n_nodes, n_chars, n_segs;
} rtk_region_out;

int rtk_region_paths_batch(rtk_ctx* ctx, const rtk_opt* opt, int pass, uint32_t n_calls, const rtk_region_call_t* calls,
                           const char* win_pool, uint64_t win_bytes, const rtk_hit* weak_pool, uint64_t n_weak,
                           const uint32_t* pid_pool, uint64_t n_pids, rtk_region_out* out, uint64_t* stats);
void rtk_region_out_free(rtk_region_out* out);

/* ---- the per-read body of search() (src/Ratatosk.cpp:808-867): getSeeds + correctSequence for a ticket of reads ----
 * Pass 1 (k = 31 graph coloured by short reads).  Inputs: reads and their qualities (qual_pool may be NULL).
 * Output: corrected read i = out_seq_pool[out_off[i], out_off[i+1]) with its quality string at the same
 * offsets of out_qual_pool - the exact bytes the reference writes to the FASTQ (library-allocated).
 * stats (optional, 24 x u64, accumulated): [0] K1 probes, [1] K1 raw hits, [2] K1 kernel ns, [3] K1 stage ns (copies +
 * host replay), [5] batched GPU service calls, [6] GPU requests served, [7] K4 / [8] K5 / [9] K2+K3+K4 kernel ns (CUDA
 * events on the launching streams; the services overlap), [10] getSeeds stage ns, [11] region stage ns, [12] bytes copied host->device, [13] device->host, [14] kernels launched
 * (process-wide tallies: exact when one batch runs at a time), [16] region-engine kernel ns, [17] region-engine calls, [18] calls
 * the engine declined (served by the request-at-a-time path instead), [19] DP cells swept inside the engine / 1024. */
int rtk_correct_batch(rtk_ctx* ctx, const rtk_opt* opt, int pass, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                      const char* qual_pool, const uint64_t* qual_off, char** out_seq_pool, char** out_qual_pool,
                      uint64_t** out_off, uint64_t* stats);
/* Same, with the reads ALSO resident in HBM already (dev_seq_pool / dev_seq_off: device copies of seq_pool / the
 * offsets rebased to 0, upper-case): the exact k-mer sweep reads them in place instead of uploading.  Used by bench.py
 * to separate the device-resident rate from the end-to-end rate; the host copies are still needed because the
 * region logic runs on the host. */
int rtk_correct_batch_resident(rtk_ctx* ctx, const rtk_opt* opt, int pass, uint32_t n_reads, const char* seq_pool,
                               const uint64_t* seq_off, const char* dev_seq_pool, const uint64_t* dev_seq_off,
                               const char* qual_pool, const uint64_t* qual_off, char** out_seq_pool, char** out_qual_pool,
                               uint64_t** out_off, uint64_t* stats);

/* Measurement aid like rtk_correct_batch_resident: tells ctx that the reads of its NEXT rtk_correct_batch / rtk_correct_two_pass_batch
 * call (same n_reads, same bytes, upper-case) are already in HBM at dev_seq_pool with offsets dev_seq_off (rebased to 0). */
int rtk_ctx_resident_reads(rtk_ctx* ctx, const char* dev_seq_pool, const uint64_t* dev_seq_off, uint32_t n_reads, uint64_t total_bases);

/* ---- phasing (src/Graph.hpp:53, src/Graph.cpp:869-1097): the step the multi-thread branch of search() runs on every read
 * of the SECOND pass before getSeeds (src/Ratatosk.cpp:832).  raw = the uncorrected read, corr / qual = its pass-1 correction.
 * Stretches of corr whose long-read colours are compatible with no other stretch of the read are reverted to raw; output
 * = the pair phasing() returns, to be passed to rtk_correct_batch(pass 2).  Reads upper-case.  K1 exact sweeps and one
 * whole-read NW path alignment (K5) per read, batched over the call. */
int rtk_phasing_batch(rtk_ctx* ctx, const rtk_opt* opt, uint32_t n_reads, const char* raw_pool, const uint64_t* raw_off,
                      const char* corr_pool, const uint64_t* corr_off, const char* qual_pool, const uint64_t* qual_off,
                      char** out_seq_pool, char** out_qual_pool, uint64_t** out_off);

/* ---- both passes of a ticket as one pipeline: pass 1 on ctx1 (k1 graph), then phasing and pass 2 on ctx2 (k2 graph; same device),
 * i.e. what `Ratatosk correct -1` followed by `correct -2 -O` from the two indexes compute for these reads (src/Ratatosk.cpp:808-867
 * run twice).  Every stage of a read depends only on that read and on the read-only graphs, so the library cuts the ticket into
 * gangs of reads that flow through the three stages independently: the stages of different gangs overlap on the device and on the
 * host instead of each waiting for the slowest read of the previous stage.  Same bytes out as the three separate calls.
 * p1_* (all three or none): the pass-1 output (what <out>.2.fastq holds).  stats1 / stats2: as rtk_correct_batch, per pass;
 * stage_ns (optional, 3 x u64, accumulated): mean over the gangs of the time spent in pass 1, phasing, pass 2. */
int rtk_correct_two_pass_batch(rtk_ctx* ctx1, rtk_ctx* ctx2, const rtk_opt* opt1, const rtk_opt* opt2, uint32_t n_reads,
                               const char* seq_pool, const uint64_t* seq_off, const char* qual_pool, const uint64_t* qual_off,
                               char** out_seq_pool, char** out_qual_pool, uint64_t** out_off, char** p1_seq_pool, char** p1_qual_pool,
                               uint64_t** p1_off, uint64_t* stats1, uint64_t* stats2, uint64_t* stage_ns);

/* ---- fixSNPs (src/Alignment.cpp:846-964; `-f` / Correct_Opt::force_unres_snp_corr, called on the pass-1 read before phasing and
 * getSeeds of the second pass, src/Ratatosk.cpp:672 / :828): an IUPAC code left by pass 1 is replaced by a base when exactly one
 * of its bases puts a k-mer of the graph over it.  One warp per read (the scan is order-dependent within a read), K1 lookups.
 * Reads upper-case.  Output: the reads at the same offsets (rebased to seq_off[0]), library-allocated (rtk_free); *n_fixed
 * (optional) = codes replaced.  rtk_phasing_batch runs this step itself when opt->force_unres_snp_corr is set; callers that
 * skip phasing (the reference's single-thread branch) call it before rtk_correct_batch(pass 2). */
int rtk_fix_snps_batch(rtk_ctx* ctx, const rtk_opt* opt, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                       char** out_seq_pool, uint64_t* n_fixed);

/* ---- graph annotation: detectSNPs (src/Graph.hpp:24, body src/Graph.cpp:484-720 with isValidSNPcandidate,
 * src/GraphTraversal.cpp:1057-1147) and detectShortCycles (src/Graph.hpp:25, body src/Graph.cpp:4660-4854), the two steps the
 * reference runs on the coloured graph after addCoverage when it builds an index (src/Ratatosk.cpp:1124-1134, :1230-1240).  Both
 * read only the colours, edge flags and sequences of the graph resident on ctx and return what the reference stores per unitig:
 *   rtk_detect_snps: UnitigData::ambiguity_ids - amb_ids[amb_off[u], amb_off[u+1]) = (position << 4 | IUPAC base set A1 C2 G4 T8),
 *     ascending; K1 substitution sweep of every coloured unitig against the graph + one warp per unitig replaying its ordered
 *     candidates with the two local traversals of isValidSNPcandidate.
 *   rtk_detect_short_cycles: UnitigData::isShortCycle() as is_cycle[u] and the compactedCycles blob
 *     cyc_pool[cyc_off[u], cyc_off[u+1]) (NUL-terminated strings, discovery order); one warp per unitig.
 * Only opt->min_cov_vertices is read (opt may be NULL: 2).  Buffers are library-allocated (rtk_free).
 * stats (optional, 10 x u64, accumulated): [0] K1 probes, [1] K1 raw hits, [2] K1 kernel ns, [3] K1 stage ns, [4] candidates,
 * [5] unitigs with candidates, [6] local traversals started, [7] annotation kernel ns, [8] unitigs re-run with the large arena. */
int rtk_detect_snps(rtk_ctx* ctx, const rtk_opt* opt, uint64_t** amb_off, uint32_t** amb_ids, uint64_t* stats);
int rtk_detect_short_cycles(rtk_ctx* ctx, const rtk_opt* opt, uint8_t** is_cycle, uint64_t** cyc_off, char** cyc_pool, uint64_t* stats);
/* ---- colouring a graph with long reads: the long_read_correct branch of addCoverage (src/Graph.hpp, body src/Graph.cpp:1561-3366) as
 * run by `Ratatosk index -2` on the k2 graph with the pass-1 corrected reads (src/Ratatosk.cpp:1213-1221).  ctx holds the graph (no
 * colours needed: rtk_graph_load with rtsk_path NULL).  Reads shorter than min_len (Correct_Opt::min_len_2nd_pass, 3000) or k are
 * skipped, bases whose quality is below getQual(min_conf) (Correct_Opt::min_confidence_2nd_pass, 0.0 -> every base pass 1 left at
 * '!') are masked, reads of the same name (name_pool / name_off, optional) share an id.  K1 exact sweep of all reads + per-unitig
 * aggregation + one warp per unitig for the edge flags.  Out, per unitig u (library-allocated, rtk_free): kmcov[u] =
 * UnitigData::kmCov_cardBranches (unphased coverage in bits 31..61, isBranching in bit 63), shared[u] bits 0..7 = the edge flags
 * (UnitigData::shared_pids), colours col_ids[col_off[u], col_off[u+1]) ascending; read_id (optional) = the id each input read
 * received (0xFFFFFFFF: none).  The reference deals the ids in an order that depends on thread timing (src/Graph.cpp:1655-1663), so
 * its colouring is reproduced up to a relabelling of the ids.  Fails when the estimated haplotype coverage reaches 10 (the
 * reference then subsamples reads at random, :2312) unless opt->reserved bit 0 is set: then every read is kept (the colouring is a
 * superset of any subsample the reference could draw; larger index, same meaning of every word).  stats (optional, 10 x u64): [0..3] K1 as above, [4] (unitig, read) pairs,
 * [5] ids dealt, [7] flag kernel ns, [8] estimated haplotype coverage. */
int rtk_color_long_reads(rtk_ctx* ctx, const rtk_opt* opt, uint32_t n_reads, const char* seq_pool, const uint64_t* seq_off,
                         const char* qual_pool, const uint64_t* qual_off, const char* name_pool, const uint64_t* name_off,
                         uint32_t min_len, double min_conf, uint64_t** kmcov, uint64_t** shared, uint64_t** col_off, uint32_t** col_ids,
                         uint32_t** read_id, uint64_t* stats);

/* The graph of g with the words and colours rtk_color_long_reads returned (a new host graph; no annotations yet): upload it and run
 * rtk_detect_snps / rtk_detect_short_cycles on it, like the reference runs them on the graph addCoverage has just coloured. */
int rtk_graph_recolor(const rtk_host_graph* g, const uint64_t* kmcov, const uint64_t* shared, const uint64_t* col_off,
                      const uint32_t* col_ids, rtk_host_graph** out);
/* writeGraphData (src/Graph.cpp:786-801) for a whole graph: head k-mer + UnitigData::write (src/UnitigData.hpp:493-517) per unitig, from
 * the words and colours of g and the given annotations; the file `Ratatosk correct -d` reads.  Written atomically. */
int rtk_rtsk_write(const rtk_host_graph* g, const char* path, const uint64_t* amb_off, const uint32_t* amb_ids, const uint8_t* is_cycle,
                   const uint64_t* cyc_off, const char* cyc_pool);

/* Index file with these annotations: a copy of rtsk_in (the .rtsk g was loaded from) in which, per unitig, the short-cycle flag, the
 * ambiguity ids and the compacted-cycles blob are replaced by the given ones and everything else is kept byte for byte - the part of
 * writeGraphData (src/Graph.cpp:786-801, UnitigData::write src/UnitigData.hpp:493-517) that detectSNPs / detectShortCycles own.  The
 * reference reads the result (readGraphData).  Written atomically (temporary file + rename); rtsk_out may equal rtsk_in. */
int rtk_rtsk_write_annotations(const rtk_host_graph* g, const char* rtsk_in, const char* rtsk_out, const uint64_t* amb_off,
                               const uint32_t* amb_ids, const uint8_t* is_cycle, const uint64_t* cyc_off, const char* cyc_pool);
/* The same for the unitigs [first_unitig, first_unitig + n_unitigs) only (clamped to the graph): unitigs are independent, so the
 * ranks of a multi-GPU job, each holding a replica of the graph, take one range each and no collective is needed; the outputs
 * still span the whole graph (empty outside the range), so partial results merge by concatenating each unitig's list. */
int rtk_detect_snps_range(rtk_ctx* ctx, const rtk_opt* opt, uint64_t first_unitig, uint64_t n_unitigs, uint64_t** amb_off,
                          uint32_t** amb_ids, uint64_t* stats);
int rtk_detect_short_cycles_range(rtk_ctx* ctx, const rtk_opt* opt, uint64_t first_unitig, uint64_t n_unitigs, uint8_t** is_cycle,
                                  uint64_t** cyc_off, char** cyc_pool, uint64_t* stats);
/* what the loaded index stores for a unitig (host slab): its ambiguity ids and its compacted-cycles blob */
int rtk_graph_unitig_annotations(const rtk_host_graph* g, uint32_t unitig, const uint32_t** amb_ids, uint64_t* n_amb,
                                 const char** cyc, uint64_t* cyc_bytes);

#ifdef __cplusplus
}
#endif
#endif /* RTK_H */
