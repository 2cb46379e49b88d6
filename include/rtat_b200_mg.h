/*
 * rtat_b200_mg.h — column-split GEMM across the GPUs of one node (one process per GPU).
 *
 * The reference has no multi-GPU code; this layer is new and sits beside the
 * single-GPU ABI of rtat_b200.h.  Scheme (BASELINE.json config 5): C's columns are split into
 * contiguous blocks, rank r owns C[:, J_r] and op(B)[:, J_r] (zero-copy views in column-major
 * storage), and A — which every rank needs in full — lives on a root rank and is replicated.
 * No reduction is needed because K is not split.
 *
 * Replication is pipelined with the math: A is cut into K-panels; while the DMMA kernel multiplies
 * panel p, the copy engines pull panel p+1 from the root's memory over NVLink (CUDA IPC peer
 * mapping, no SMs used), so only the first panel's transfer is exposed.  The process group that
 * exchanges the 64-byte IPC handle and provides barriers is the caller's (torch.distributed / NCCL
 * in bench.py, gloo in the CPU tests); an NCCL broadcast of the panels is the alternative transport
 * and is what bench.py falls back to when peer mapping is unavailable.
 */
#ifndef RTAT_B200_MG_H
#define RTAT_B200_MG_H
#include "rtat_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtat_b200_mg_context *rtat_b200_mg_t;
#define RTAT_B200_MG_HANDLE_BYTES 64

/* pure host helper: rank's contiguous column block [first, first + count) of n columns */
int rtat_b200_mg_partition(int n, int world, int rank, int *first, int *count);
/* pure host helper: K-panel p of `panels` over extent k: [first, first + count), multiples of 16 but the last */
int rtat_b200_mg_panel(int k, int panels, int p, int *first, int *count);

/* pure host helper: panel p (0..3) of the automatic schedule — 1/16, 3/16, 4/16, 8/16 of k */
int rtat_b200_mg_auto_panel(int k, int p, int *first, int *count);

int rtat_b200_mg_create(rtat_b200_mg_t *mg, rtat_b200_handle_t handle, int rank, int world);
int rtat_b200_mg_destroy(rtat_b200_mg_t mg);

/* root: allocate `bytes` of device memory other processes can map; handle_out travels through the
 * caller's process group */
int rtat_b200_mg_alloc_shared(rtat_b200_mg_t mg, size_t bytes, void **dptr, unsigned char handle_out[RTAT_B200_MG_HANDLE_BYTES]);
int rtat_b200_mg_free_shared(rtat_b200_mg_t mg, void *dptr);
/* non-root: map the root's allocation into this process (peer access over NVLink) */
int rtat_b200_mg_open_shared(rtat_b200_mg_t mg, const unsigned char handle[RTAT_B200_MG_HANDLE_BYTES], void **peer_dptr);
int rtat_b200_mg_close_shared(rtat_b200_mg_t mg, void *peer_dptr);

/* C_loc (m x n_loc) = alpha * op(A) * op(B_loc) + beta * C_loc.
 * A_src: the full A — local memory on the root, the mapped peer pointer elsewhere.
 * A_stage: local buffer with A's shape and leading dimension that receives the panels (may equal
 * A_src, or be NULL, when A_src is local: then nothing is copied).
 * panels: number of equal K-panels the transfer is pipelined over, or 0 for the automatic schedule
 * (a small first panel, whose transfer is the only exposed one, then three growing ones).
 * Work is enqueued on the handle's stream (math) and an internal copy stream; returns immediately. */
int rtat_b200_mg_dgemm(rtat_b200_mg_t mg, int transa, int transb, int m, int n_loc, int k, const double *alpha, const double *A_src,
                       int lda, double *A_stage, const double *B_loc, int ldb, const double *beta, double *C_loc, int ldc, int panels);

/* ---- host-buffer column split: A, B_loc, C_loc in (pinned) host memory -----------------------------------------------
 * The multi-GPU counterpart of rtat_b200_dgemm_host (rtat_b200.h).  Every rank passes the same full A and its own column
 * block of op(B) / C.  A crosses PCIe ONCE per step in total: for every K-panel each rank uploads only its 1/world share
 * into its replica of A, publishes a flag, and pulls the other shares from the peers' replicas over NVLink with the copy
 * engines (replicas are IPC-mapped into every process).  A flag is a generation number in the WRITER's own page, set by a
 * 4-byte H2D copy queued behind the upload; a reader polls it over NVLink with a one-thread kernel (30 s timeout -> error
 * status), so nothing is ever pushed across GPUs and the step needs no process-group collective.  (Option "mg.signal" = 0,
 * read by _host_setup, selects the first design instead — flags pushed with 4-byte peer copies, waits as stream memory
 * operations — which measured 20 % slower: the pushed words land late.)  Around that exchange the schedule is the single-GPU host
 * pipeline: K-panels of A against the leading column blocks while A is still arriving, then whole-K column blocks with
 * their C blocks streaming back to the host.
 *
 * Setup (once): every rank calls _host_setup, the 64-byte handles are all-gathered through the caller's process group,
 * every rank calls _host_connect with the world x 64 bytes in rank order.  _dgemm_host is collective: every rank must
 * make the same sequence of calls (same m, k, transa, lda; n_loc > 0 on every rank).  It returns when C_loc has landed. */
#define RTAT_B200_MG_MAX_WORLD 16
int rtat_b200_mg_host_setup(rtat_b200_mg_t mg, size_t bytes_a, unsigned char handle_out[RTAT_B200_MG_HANDLE_BYTES]);
int rtat_b200_mg_host_connect(rtat_b200_mg_t mg, const unsigned char *handles /* world x RTAT_B200_MG_HANDLE_BYTES */);
int rtat_b200_mg_host_teardown(rtat_b200_mg_t mg);
int rtat_b200_mg_dgemm_host(rtat_b200_mg_t mg, int transa, int transb, int m, int n_loc, int k, const double *alpha, const double *A,
                            int lda, const double *B_loc, int ldb, const double *beta, double *C_loc, int ldc);

#ifdef __cplusplus
}
#endif
#endif
