/*
 * rtat_b200_host.h — C entry points over the C++ host layer (rtatblas_b200/include/rtat: planner,
 * executor, plan algebra), built as rtatblas_b200/lib/librtat_host.so.  They exist so that callers
 * without a C++ toolchain drive the same tuned-GEMM path a C++ user gets from rtat.h:
 *
 *   reference interface (file:line)                              entry point
 *   ----------------------------------------------------------   -----------------------------------
 *   Planning_System<GEMM_Executor[_Pad]<T>> ctor                 rtat_host_planner_create
 *     (src/planning/planning_system.h:38-63, src/rtat.h:20-37)
 *   Planning_System::create_plan + ::execute (:65-114)           rtat_host_planner_gemm
 *   Planning_System::calculate_workspace (:116-118)              rtat_host_planner_workspace
 *   Planning_System::make_statistics + Planner_Statistics::json  rtat_host_planner_statistics
 *     (:120-130, src/planning/planner_statistics.h:50-67)
 *   (new) JSON plan cache, in memory and on disk                 rtat_host_planner_{save,load}_plans[_file]
 *
 * Matrices are device pointers, column-major; the work runs on the stream bound to `handle`.
 * Functions returning char* hand out malloc'd text to be released with rtat_host_free.
 * Errors: -1 / NULL, with the message available from rtat_host_last_error().
 */
#ifndef RTAT_B200_HOST_H
#define RTAT_B200_HOST_H
#include "rtat_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef void *rtat_host_planner_t;

/* method: "gemm" (8 plans) or "gemm_pad" (64 plans); data_type: "double" or "float" */
rtat_host_planner_t rtat_host_planner_create(const char *method, const char *data_type);
void rtat_host_planner_destroy(rtat_host_planner_t planner);

/* One tuned GEMM: plan = forced_plan if given (e.g. "TNN"), else create_plan(key); then execute with
 * the caller's workspace (a plan that does not fit degrades to the default plan, as in the
 * reference).  alpha/beta point to host scalars of the planner's data type.  plan_out (>= 8 bytes,
 * optional) receives the plan that was requested. */
int rtat_host_planner_gemm(rtat_host_planner_t planner, rtat_b200_handle_t handle, int transa, int transb, int m, int n, int k,
                           const void *alpha, const void *A, int lda, const void *B, int ldb, const void *beta, void *C, int ldc,
                           void *workspace, size_t workspace_bytes, const char *forced_plan, char *plan_out);
/* bytes of workspace `plan` needs for this problem */
long long rtat_host_planner_workspace(rtat_host_planner_t planner, int transa, int transb, int m, int n, int k, int lda, int ldb,
                                      int ldc, const char *plan);
/* SYRK / TRSM through their planners (reference src/rtat.h:23-26 syrk_planner / trsm_planner; plan strings are the
 * two Bool_Op letters of SYRK_Options "ac" / TRSM_Options "sa", e.g. "FT").  The planner must have been created with
 * method "syrk" / "trsm"; other planners fail with -1.  Arguments as in rtat_b200_{d,s}syrk / _trsm; alpha/beta point
 * to host scalars of the planner's data type. */
int rtat_host_planner_syrk(rtat_host_planner_t planner, rtat_b200_handle_t handle, int uplo, int trans, int n, int k, const void *alpha,
                           const void *A, int lda, const void *beta, void *C, int ldc, void *workspace, size_t workspace_bytes,
                           const char *forced_plan, char *plan_out);
long long rtat_host_planner_syrk_workspace(rtat_host_planner_t planner, int uplo, int trans, int n, int k, int lda, int ldc,
                                           const char *plan);
int rtat_host_planner_trsm(rtat_host_planner_t planner, rtat_b200_handle_t handle, int side, int uplo, int trans, int diag, int m, int n,
                           const void *alpha, const void *A, int lda, void *B, int ldb, void *workspace, size_t workspace_bytes,
                           const char *forced_plan, char *plan_out);
long long rtat_host_planner_trsm_workspace(rtat_host_planner_t planner, int side, int uplo, int trans, int diag, int m, int n, int lda,
                                           int ldb, const char *plan);
char *rtat_host_planner_statistics(rtat_host_planner_t planner);
char *rtat_host_planner_save_plans(rtat_host_planner_t planner);
long long rtat_host_planner_load_plans(rtat_host_planner_t planner, const char *json_text);
int rtat_host_planner_set_tests_until_converge(rtat_host_planner_t planner, int n);
/* On-disk plan cache (reference src/app/runner.h:81-82 keeps its statistics in a JSON file the same way): save overwrites
 * `path` with the converged plans; load returns how many plans were taken over (0 when the file does not exist), -1 when
 * the file is malformed. */
int rtat_host_planner_save_plans_file(rtat_host_planner_t planner, const char *path);
long long rtat_host_planner_load_plans_file(rtat_host_planner_t planner, const char *path);
/* Planning_System::set_default_preference: relative margin within which the default plan is kept over a faster-looking
 * one (0, the default, = the reference's strict minimum of the means). */
int rtat_host_planner_set_default_preference(rtat_host_planner_t planner, double margin);
/* statistics with the derived columns "tflops" / "gbs" next to "times" (what run_tests prints) */
char *rtat_host_planner_statistics_rates(rtat_host_planner_t planner);
int rtat_host_planner_num_plans(rtat_host_planner_t planner);
char *rtat_host_planner_plan_name(rtat_host_planner_t planner, int index);
void rtat_host_free(char *text);
const char *rtat_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
