/* oracle.c — see oracle.h.  TEST INFRASTRUCTURE ONLY: CPU restatement of the reference algorithm. */
#include "oracle.h"
#include <string.h>
#include <math.h>

static int roundup(int x, int pad) { return ((x + pad - 1) / pad) * pad; }

struct plan_flags { int ta, tb, tc, pa, pb, pc; };

/* GEMM_Options string "xyz" (gemm.cpp:147-157, >> at :170-181) or GEMM_Options_Pad "xyzabc"
 * (gemm.cpp:41-53, >> at :64-80: transa,transb,transc,pada,padb,padc). */
static int parse_plan(const char *plan, struct plan_flags *f) {
  size_t len = plan ? strlen(plan) : 0;
  if (len != 3 && len != 6) return 1;
  int *t[3] = {&f->ta, &f->tb, &f->tc};
  int *p[3] = {&f->pa, &f->pb, &f->pc};
  for (int i = 0; i < 3; i++) {
    if (plan[i] == 'T') *t[i] = 1; else if (plan[i] == 'N') *t[i] = 0; else return 1;
    *p[i] = 0;
    if (len == 6) { if (plan[3 + i] == 'P') *p[i] = 1; else if (plan[3 + i] == 'N') *p[i] = 0; else return 1; }
  }
  return 0;
}

/* MatrixOp::workspace_req (matrixop.h:172-185) specialised to the trees form_operation builds:
 * ws = [ta|pa] fp(A') + [tb|pb] fp(B') + [tc|pc] fp(S), fp = ld'*cols. */
size_t oracle_gemm_plan_workspace(const char *plan, int transa, int transb, int m, int n, int k,
                                  size_t elem_size) {
  struct plan_flags f;
  if (parse_plan(plan, &f)) return (size_t)-1;
  int a_rows = transa ? k : m, a_cols = transa ? m : k;
  int b_rows = transb ? n : k, b_cols = transb ? k : n;
  size_t ws = 0;
  if (f.ta || f.pa) {
    int r = f.ta ? a_cols : a_rows, c = f.ta ? a_rows : a_cols;
    ws += (size_t)roundup(r, f.pa ? 32 : 1) * c;
  }
  if (f.tb || f.pb) {
    int r = f.tb ? b_cols : b_rows, c = f.tb ? b_rows : b_cols;
    ws += (size_t)roundup(r, f.pb ? 32 : 1) * c;
  }
  if (f.tc) ws += (size_t)roundup(n, f.pc ? 32 : 1) * m;   /* S is n x m (gemm.cpp:108-113, :205-210) */
  else if (f.pc) ws += (size_t)roundup(m, 32) * n;         /* gemm.cpp:118-123 */
  return ws * elem_size;
}

static int parse_bool_plan(const char *plan) {
  if (!plan || strlen(plan) != 2) return 1;
  for (int i = 0; i < 2; i++)
    if (plan[i] != 'F' && plan[i] != 'T') return 1;
  return 0;
}

/* workspace_req of the SYRK trees (syrk.cpp:72-104): [transpose_C] n*n + [transpose_A] footprint(A^T) */
size_t oracle_syrk_plan_workspace(const char *plan, int trans, int n, int k, size_t elem_size) {
  (void)trans;
  if (parse_bool_plan(plan)) return (size_t)-1;
  size_t ws = 0;
  if (plan[0] == 'T') ws += (size_t)n * k;
  if (plan[1] == 'T') ws += (size_t)n * n;
  return ws * elem_size;
}

/* workspace_req of the TRSM trees (trsm.cpp:72-113): [swap_side] n*m (B^T) + [transpose_A] na*na */
size_t oracle_trsm_plan_workspace(const char *plan, int left, int m, int n, size_t elem_size) {
  if (parse_bool_plan(plan)) return (size_t)-1;
  size_t na = left ? m : n, ws = 0;
  if (plan[1] == 'T') ws += na * na;
  if (plan[0] == 'T') ws += (size_t)n * m;
  return ws * elem_size;
}

#define T double
#define SUF d
#define WIDE long double
#define FMA fma
#include "oracle_impl.inc"
#undef T
#undef SUF
#undef WIDE
#undef FMA

#define T float
#define SUF s
#define WIDE double
#define FMA fmaf
#include "oracle_impl.inc"
#undef T
#undef SUF
#undef WIDE
#undef FMA

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

void oracle_dfill_uniform(double *x, size_t n, uint64_t seed, double lo, double hi) {
  for (size_t i = 0; i < n; i++) {
    uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + i);
    double u = (double)((r >> 11) + 1) * (1.0 / 9007199254740992.0);
    x[i] = lo + (hi - lo) * u;
  }
}
void oracle_sfill_uniform(float *x, size_t n, uint64_t seed, float lo, float hi) {
  for (size_t i = 0; i < n; i++) {
    uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + i);
    double u = (double)((r >> 11) + 1) * (1.0 / 9007199254740992.0);
    x[i] = (float)((double)lo + ((double)hi - (double)lo) * u);
  }
}
void oracle_fill_uniform_d(double *x, size_t n, uint64_t seed, double lo, double hi) { oracle_dfill_uniform(x, n, seed, lo, hi); }
void oracle_fill_uniform_s(float *x, size_t n, uint64_t seed, float lo, float hi) { oracle_sfill_uniform(x, n, seed, lo, hi); }
