// Host-only tests of the C++ layer (no GPU needed): restated from the reference's
// tests/workspace_test.cpp and tests/json_test.cpp, plus the plan algebra / key ordering /
// option enumeration / op-tree workspace accounting that the reference only exercises on a GPU.
#include <unistd.h>
#include <cstdlib>
#include <fstream>
#include <set>
#include <sstream>
#include <rtat/rtat.h>
#include "mini_test.h"

using namespace rtat;

// ---- tests/workspace_test.cpp -------------------------------------------------------------------
TEST(Workspace_Test, Default_Constructor) { Workspace w; ASSERT_EQ(w.size<char>(), 0u); }

TEST(Workspace_Test, Size) {
  size_t N = 8192;
  std::vector<char> bytes(N);
  Workspace space(bytes.data(), N);
  ASSERT_EQ(space.size<char>(), N);
  ASSERT_EQ(space.size<double>(), N / sizeof(double));
  ASSERT_EQ(space.size<int>(), N / sizeof(int));
  ASSERT_EQ(space.size<float>(), N / sizeof(float));
}

TEST(Workspace_Test, Offset_Constructor) {
  size_t N = 1024;
  std::vector<char> bytes(N);
  Workspace space(bytes.data(), N);
  size_t offset = 301, size = 123;
  Workspace sub(space, offset, size);
  ASSERT_EQ((char *)sub, &(((char *)space)[offset]));
  ASSERT_EQ(sub.size<char>(), size);
  ASSERT_DEATH(Workspace(space, N - 2, 10));  // stepping outside the parent ends the process
}

TEST(Workspace_Test, Peel) {
  size_t N = 1024;
  std::vector<char> bytes(N);
  Workspace space(bytes.data(), N);
  size_t s1 = 18, s2 = 27, s3 = 48;
  Workspace p1 = space.peel<double>(s1), p2 = space.peel<float>(s2), p3 = space.peel<char>(s3);
  size_t used = p1.size<char>() + p2.size<char>() + p3.size<char>();
  ASSERT_EQ(p1.size<char>(), s1 * sizeof(double));
  ASSERT_EQ((char *)p1, &bytes[0]);
  ASSERT_EQ(p2.size<char>(), s2 * sizeof(float));
  ASSERT_EQ((char *)p2, &bytes[s1 * sizeof(double)]);
  ASSERT_EQ(p3.size<char>(), s3);
  ASSERT_EQ((char *)p3, &bytes[s1 * sizeof(double) + s2 * sizeof(float)]);
  ASSERT_EQ(space.size<char>(), N - used);
  ASSERT_EQ((char *)space, &bytes[used]);
}

TEST(Matrix_Test, Dims_And_Views) {
  ASSERT_THROWS(MatrixDims(10, 3, 9));  // ld < m
  MatrixDims d(10, 3, 12);
  ASSERT_EQ(d.footprint(), 36u);
  std::vector<double> buf(36);
  Workspace w(buf.data(), 36);
  Matrix<double> M(w, d);
  ASSERT_EQ(M.ptr(), buf.data());
  ASSERT_THROWS(Matrix<double>(Workspace(buf.data(), 35), d));  // span smaller than the footprint
  Matrix<double> blk = M.column_block(1, 2);
  ASSERT_EQ(blk.ptr(), buf.data() + 12);
  ASSERT_EQ(blk.dims().n, 2u);
  ASSERT_EQ(blk.dims().ld, 12u);
  ASSERT_THROWS(M.column_block(2, 2));
}

// ---- tests/json_test.cpp (GEMM parts) + plain GEMM_Options, which the reference leaves untested --
TEST(JSON_Test, GEMM_Key) {
  for (auto ta : {"N", "T"})
    for (auto tb : {"N", "T"}) {
      Json j = Json::object();
      j["transA"] = ta; j["transB"] = tb; j["m"] = 6741; j["n"] = 2463; j["k"] = 155;
      GEMM_Key key = from_json<GEMM_Key>(j);
      ASSERT_EQ(to_json(key), j);
      ASSERT_EQ(key.m, 6741); ASSERT_EQ(key.n, 2463); ASSERT_EQ(key.k, 155);
      GEMM_Key again = from_json<GEMM_Key>(to_json(key));
      ASSERT_TRUE(!(again < key) && !(key < again));
    }
}

TEST(JSON_Test, GEMM_Options_And_Pad) {
  for (auto &o : GEMM_Options::enumerate()) {
    GEMM_Options back = from_json<GEMM_Options>(to_json(o));
    ASSERT_TRUE(!(back < o) && !(o < back));
    ASSERT_EQ(to_json(back), to_json(o));
  }
  for (auto &o : GEMM_Options_Pad::enumerate()) {
    GEMM_Options_Pad back = from_json<GEMM_Options_Pad>(to_json(o));
    ASSERT_TRUE(!(back < o) && !(o < back));
    Json j = to_json(o);
    ASSERT_EQ(j["padA"].get<std::string>(), std::string(o.pada));
    ASSERT_EQ(j["transC"].get<std::string>(), std::string(o.transc));
  }
  Json bad = Json::object();
  bad["transA"] = "X"; bad["transB"] = "N"; bad["transC"] = "N";
  ASSERT_THROWS(from_json<GEMM_Options>(bad));
}

TEST(JSON_Test, Parse_Dump_Round_Trip) {
  std::string text = "{\"keywords\": {\"run_type\": \"exhaustive\", \"repetitions\": 3, \"x\": 1.5e-3, \"flag\": true, \"n\": null},\n"
                     " \"problems\": [{\"m\": 1, \"transA\": \"N\"}, [], {}], \"s\": \"a\\\"b\\n\"}";
  Json j = Json::parse(text);
  ASSERT_EQ(j["keywords"]["repetitions"].get<int>(), 3);
  ASSERT_NEAR(j["keywords"]["x"].get<double>(), 1.5e-3, 1e-18);
  ASSERT_EQ(j["problems"].size(), 3u);
  ASSERT_EQ(j["s"].get<std::string>(), std::string("a\"b\n"));
  ASSERT_EQ(Json::parse(j.dump(2)), j);
  ASSERT_EQ(Json::parse(j.dump()), j);
  // object keys come out sorted, floats in shortest round-trip form
  Json o = Json::object();
  o["b"] = 1; o["a"] = 0.1f;
  ASSERT_EQ(o.dump(), std::string("{\"a\":0.10000000149011612,\"b\":1}"));
  ASSERT_THROWS(Json::parse("{\"a\": }"));
  ASSERT_THROWS(Json::parse("[1, 2"));
}

// ---- keys, options ----------------------------------------------------------------------------------
TEST(GEMM_Key_Test, Text_Form_And_Ordering) {
  GEMM_Key k(BLAS_Operation("N"), BLAS_Operation("T"), /*m*/ 4097, /*k*/ 2049, /*n*/ 3001);
  ASSERT_EQ(std::string(k), std::string("N,T,,4097,2049,3001"));  // double comma; m,k,n order (gemm.cpp:6-15)
  ASSERT_EQ(k.n, 3001); ASSERT_EQ(k.k, 2049);
  // ordering is lexicographic on that text, not numeric: "1024" < "128"
  GEMM_Key a(gpu::BLAS_OP_N, gpu::BLAS_OP_N, 1024, 8, 8), b(gpu::BLAS_OP_N, gpu::BLAS_OP_N, 128, 8, 8);
  ASSERT_TRUE(a < b); ASSERT_FALSE(b < a);
  GEMM_Key c(gpu::BLAS_OP_T, gpu::BLAS_OP_N, 1, 1, 1);
  ASSERT_TRUE(b < c);
  ASSERT_NEAR(k.flopcount(), 2.0 * 4097 * 3001 * 2049, 1);
}

TEST(Options_Test, Enumeration_Order_And_Text) {
  std::vector<std::string> names;
  for (auto &o : GEMM_Options::enumerate()) names.push_back(o);
  std::vector<std::string> want = {"NNN", "NNT", "NTN", "NTT", "TNN", "TNT", "TTN", "TTT"};
  ASSERT_TRUE(names == want);
  auto pads = GEMM_Options_Pad::enumerate();
  ASSERT_EQ(pads.size(), 64u);
  ASSERT_EQ(std::string(pads[0]), std::string("NNNNNN"));
  ASSERT_EQ(std::string(pads[1]), std::string("NNNNNP"));   // pad C varies fastest
  ASSERT_EQ(std::string(pads[8]), std::string("NNTNNN"));   // then trans C
  ASSERT_EQ(std::string(pads[63]), std::string("TTTPPP"));
  std::set<std::string> uniq;
  for (auto &o : pads) uniq.insert(o);
  ASSERT_EQ(uniq.size(), 64u);
  ASSERT_EQ(std::string(GEMM_Options::default_opts()), std::string("NNN"));
  GEMM_Options_Pad p;
  std::istringstream in("TNTPNP");
  ASSERT_TRUE(bool(in >> p));
  ASSERT_TRUE(p.transa == BLAS_Op::TRANS && p.transb == BLAS_Op::NOTRANS && p.transc == BLAS_Op::TRANS);
  ASSERT_TRUE(p.pada == Pad_Op::PAD && p.padb == Pad_Op::NOPAD && p.padc == Pad_Op::PAD);
  GEMM_Options o;
  std::istringstream bad("NNNN");
  ASSERT_FALSE(bool(bad >> o));
  ASSERT_EQ(std::string(!BLAS_Op(BLAS_Op::TRANS)), std::string("N"));
  ASSERT_EQ(std::string(!BLAS_Operation("N")), std::string("T"));
  ASSERT_THROWS(BLAS_Operation("Q"));
  ASSERT_THROWS(Pad_Op(std::string("T")));
}

TEST(Option_Filter_Test, Predicates) {
  using P = std::pair<GEMM_Options, GEMM_Key>;
  Predicate<P> no_tc = [](P p) { return !(p.first.transc == BLAS_Op::TRANS); };
  Predicate<P> small = [](P p) { return p.second.m < 100; };
  GEMM_Key k1(gpu::BLAS_OP_N, gpu::BLAS_OP_N, 10, 10, 10), k2(gpu::BLAS_OP_N, gpu::BLAS_OP_N, 1000, 10, 10);
  Option_Filter<GEMM_Key, GEMM_Options> all;
  ASSERT_EQ(all.apply(k1).size(), 8u);
  Option_Filter<GEMM_Key, GEMM_Options> both(conjunction<P>({no_tc, small}));
  ASSERT_EQ(both.apply(k1).size(), 4u);
  ASSERT_EQ(both.apply(k2).size(), 0u);
  Option_Filter<GEMM_Key, GEMM_Options> either(disjunction<P>({no_tc, small}));
  ASSERT_EQ(either.apply(k2).size(), 4u);
  ASSERT_EQ(either.apply(k1).size(), 8u);
}

// ---- plan algebra: workspace accounting of the operation trees -------------------------
static size_t expected_ws(const std::string &plan, bool ta, bool tb, size_t m, size_t n, size_t k, size_t elem) {
  std::string pad = plan.size() == 6 ? plan.substr(3) : "NNN";
  auto up = [](size_t x, size_t p) { return (x + p - 1) / p * p; };
  size_t ar = ta ? k : m, ac = ta ? m : k, br = tb ? n : k, bc = tb ? k : n, ws = 0;
  if (plan[0] == 'T' || pad[0] == 'P') { size_t r = plan[0] == 'T' ? ac : ar, c = plan[0] == 'T' ? ar : ac; ws += up(r, pad[0] == 'P' ? 32 : 1) * c; }
  if (plan[1] == 'T' || pad[1] == 'P') { size_t r = plan[1] == 'T' ? bc : br, c = plan[1] == 'T' ? br : bc; ws += up(r, pad[1] == 'P' ? 32 : 1) * c; }
  if (plan[2] == 'T') ws += up(n, pad[2] == 'P' ? 32 : 1) * m;
  else if (pad[2] == 'P') ws += up(m, 32) * n;
  return ws * elem;
}

template <typename T, typename Opts>
static void check_plan_space(bool ta, bool tb, size_t m, size_t n, size_t k) {
  size_t ar = ta ? k : m, ac = ta ? m : k, br = tb ? n : k, bc = tb ? k : n;
  // never dereferenced: only shapes matter for the accounting
  Matrix<T> A(Workspace((T *)nullptr, ar * ac), ar, ac, ar), B(Workspace((T *)nullptr, br * bc), br, bc, br), C(Workspace((T *)nullptr, m * n), m, n, m);
  GEMM_Inputs<T> in(nullptr, ta ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, tb ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, A, B, C, T(1), T(0));
  for (auto &o : Opts::enumerate()) {
    auto op = Opts(o).form_operation(in);
    ASSERT_EQ(op->workspace_req_bytes(), expected_ws(o, ta, tb, m, n, k, sizeof(T)));
    ASSERT_EQ(op->output_space_req(), 0u);       // every plan writes into the caller's C
    ASSERT_EQ(op->dims().m, m);
    ASSERT_EQ(op->dims().n, n);
  }
}

TEST(Plan_Algebra_Test, Workspace_Accounting_All_Plans) {
  for (bool ta : {false, true})
    for (bool tb : {false, true}) {
      check_plan_space<double, GEMM_Options>(ta, tb, 70, 45, 62);
      check_plan_space<float, GEMM_Options>(ta, tb, 23, 16, 35);
      check_plan_space<double, GEMM_Options_Pad>(ta, tb, 4097, 3001, 2049);
      check_plan_space<float, GEMM_Options_Pad>(ta, tb, 33, 64, 1);
    }
}

TEST(Plan_Algebra_Test, Config2_Scratch_Sizes) {
  // DGEMM 8192^3 op TN: explicit-transpose plan TNN needs one 512 MiB scratch, TTT needs three
  size_t n = 8192;
  Matrix<double> A(Workspace((double *)nullptr, n * n), n, n, n);
  GEMM_Inputs<double> in(nullptr, gpu::BLAS_OP_T, gpu::BLAS_OP_N, A, A, A, 1.0, 0.0);
  ASSERT_EQ(GEMM_Options(BLAS_Op::NOTRANS, BLAS_Op::NOTRANS, BLAS_Op::NOTRANS).form_operation(in)->workspace_req_bytes(), 0u);
  ASSERT_EQ(GEMM_Options(BLAS_Op::TRANS, BLAS_Op::NOTRANS, BLAS_Op::NOTRANS).form_operation(in)->workspace_req_bytes(), size_t(512) << 20);
  ASSERT_EQ(GEMM_Options(BLAS_Op::TRANS, BLAS_Op::TRANS, BLAS_Op::TRANS).form_operation(in)->workspace_req_bytes(), size_t(3) * (size_t(512) << 20));
}

TEST(MatrixOp_Test, Shape_Checks_Are_Fatal) {
  Matrix<double> A(Workspace((double *)nullptr, 12), 3, 4, 3), B(Workspace((double *)nullptr, 10), 5, 2, 5), C(Workspace((double *)nullptr, 6), 3, 2, 3);
  auto mk = [](Matrix<double> M) { return std::unique_ptr<MatrixOp<double>>(std::make_unique<NoOp<double>>(M)); };
  ASSERT_DEATH(MatrixMult<double>(mk(A), mk(B), mk(C), false, false, 1.0, 0.0));       // k mismatch 4 vs 5
  ASSERT_DEATH(MatrixAccumulate<double>(mk(A), mk(C), 1.0, 0.0, false));               // 3x4 into 3x2
  MatrixMove<double> mv(mk(A), 1.0, true, 32);
  ASSERT_EQ(mv.dims().m, 4u); ASSERT_EQ(mv.dims().n, 3u); ASSERT_EQ(mv.dims().ld, 32u);
  ASSERT_EQ(mv.output_space_req(), 96u);
  ASSERT_EQ(mv.scratch_space_req(), 0u);
}

// ---- SYRK / TRSM host pieces (reference src/methods/syrk.{h,cpp}, trsm.{h,cpp}, json_encoding.h:110-223) ----------
TEST(SYRK_TRSM_Test, Key_Text_Ordering_And_Json) {
  SYRK_Key sk(gpu::BLAS_FILL_MODE_LOWER, gpu::BLAS_OP_T, 213, 74);
  ASSERT_EQ(std::string(sk), std::string("Lower,T,213,74"));
  ASSERT_TRUE(SYRK_Key(gpu::BLAS_FILL_MODE_LOWER, gpu::BLAS_OP_N, 9, 9) < SYRK_Key(gpu::BLAS_FILL_MODE_UPPER, gpu::BLAS_OP_N, 1, 1));
  ASSERT_TRUE(SYRK_Key(gpu::BLAS_FILL_MODE_LOWER, gpu::BLAS_OP_N, 1024, 8) < SYRK_Key(gpu::BLAS_FILL_MODE_LOWER, gpu::BLAS_OP_N, 128, 8));  // text order
  ASSERT_EQ(to_json(sk).dump(), std::string("{\"k\":74,\"n\":213,\"trans\":\"T\",\"uplo\":\"Lower\"}"));
  ASSERT_EQ(std::string(from_json<SYRK_Key>(to_json(sk))), std::string(sk));
  ASSERT_EQ(sk.flopcount(), 213.0 * 214.0 * 74.0);

  TRSM_Key tk(gpu::BLAS_SIDE_LEFT, gpu::BLAS_FILL_MODE_LOWER, gpu::BLAS_OP_N, gpu::BLAS_DIAG_NON_UNIT, 435, 527);
  ASSERT_EQ(std::string(tk), std::string("Left,Lower,N,Non-Unit,435,527"));
  TRSM_Key tk2(gpu::BLAS_SIDE_RIGHT, gpu::BLAS_FILL_MODE_UPPER, gpu::BLAS_OP_T, gpu::BLAS_DIAG_UNIT, 3, 5);
  ASSERT_EQ(std::string(tk2), std::string("Right,Upper,T,Unit,3,5"));
  ASSERT_TRUE(tk < tk2);
  ASSERT_EQ(to_json(tk2).dump(), std::string("{\"diag\":\"Unit\",\"m\":3,\"n\":5,\"side\":\"Right\",\"trans\":\"T\",\"uplo\":\"Upper\"}"));
  ASSERT_EQ(std::string(from_json<TRSM_Key>(to_json(tk))), std::string(tk));
  ASSERT_EQ(tk.flopcount(), 435.0 * 435.0 * 527.0);
  ASSERT_EQ(tk2.flopcount(), 5.0 * 5.0 * 3.0);
}

template <typename Opts>
static void check_bool_plan_space(const char *json_first, const char *json_second) {
  auto all = Opts::enumerate();
  ASSERT_EQ(all.size(), 4u);
  const char *names[4] = {"FF", "FT", "TF", "TT"};
  for (int i = 0; i < 4; i++) {
    ASSERT_EQ(std::string(all[i]), std::string(names[i]));
    Opts parsed;
    std::istringstream in(names[i]);
    ASSERT_TRUE(bool(in >> parsed));
    ASSERT_EQ(std::string(parsed), std::string(names[i]));  // the reference's >> always yields "TT" (syrk.cpp:64-65)
    Json j = to_json(all[i]);
    ASSERT_EQ(j[json_first].template get<std::string>(), std::string(1, names[i][0]));
    ASSERT_EQ(j[json_second].template get<std::string>(), std::string(1, names[i][1]));
    ASSERT_EQ(std::string(from_json<Opts>(j)), std::string(names[i]));
  }
  ASSERT_EQ(std::string(Opts::default_opts()), std::string("FF"));
  Opts bad;
  std::istringstream three("FFF");
  ASSERT_TRUE(!(three >> bad));
}
TEST(SYRK_TRSM_Test, Options_Enumeration_Text_And_Json) {
  check_bool_plan_space<SYRK_Options>("transA", "transC");
  check_bool_plan_space<TRSM_Options>("swap_side", "transA");
}

TEST(SYRK_TRSM_Test, Workspace_Accounting) {
  // SYRK: [transpose_C] n*n + [transpose_A] n*k elements (syrk.cpp:72-104 through matrixop.h:172-185)
  for (bool trans : {false, true}) {
    const size_t n = 213, k = 74, ar = trans ? k : n, ac = trans ? n : k;
    Matrix<double> A(Workspace((double *)nullptr, ar * ac), ar, ac, ar), C(Workspace((double *)nullptr, n * n), n, n, n);
    SYRK_Inputs<double> in(nullptr, gpu::BLAS_FILL_MODE_UPPER, trans ? gpu::BLAS_OP_T : gpu::BLAS_OP_N, A, C, 1.0, 0.0);
    ASSERT_EQ(in.n(), n);
    ASSERT_EQ(in.k(), k);
    for (auto o : SYRK_Options::enumerate()) {
      auto op = o.form_operation(in);
      ASSERT_EQ(op->workspace_req_bytes(), sizeof(double) * ((o.transpose_A ? n * k : 0) + (o.transpose_C ? n * n : 0)));
      ASSERT_EQ(op->output_space_req(), 0u);
      ASSERT_EQ(op->dims().m, n);
    }
  }
  // TRSM: [swap_side] n*m + [transpose_A] na*na elements (trsm.cpp:72-113)
  for (bool left : {false, true}) {
    const size_t m = 435, n = 527, na = left ? m : n;
    Matrix<float> A(Workspace((float *)nullptr, na * na), na, na, na), B(Workspace((float *)nullptr, m * n), m, n, m);
    TRSM_Inputs<float> in(nullptr, left ? gpu::BLAS_SIDE_LEFT : gpu::BLAS_SIDE_RIGHT, gpu::BLAS_FILL_MODE_LOWER, gpu::BLAS_OP_N,
                          gpu::BLAS_DIAG_UNIT, A, B, 1.f);
    for (auto o : TRSM_Options::enumerate()) {
      auto op = o.form_operation(in);
      ASSERT_EQ(op->workspace_req_bytes(), sizeof(float) * ((o.swap_side ? m * n : 0) + (o.transpose_A ? na * na : 0)));
      ASSERT_EQ(op->dims().m, m);
      ASSERT_EQ(op->dims().n, n);
    }
  }
  // shape checks are fatal, like the reference's bare `throw;`
  auto mk = [](size_t r, size_t c) {
    return std::unique_ptr<MatrixOp<double>>(std::make_unique<NoOp<double>>(Matrix<double>(Workspace((double *)nullptr, r * c), r, c, r)));
  };
  ASSERT_DEATH(MatrixSyrk<double>(mk(5, 3), mk(4, 4), true, false, 1.0, 0.0));          // n mismatch
  ASSERT_DEATH(MatrixSyrk<double>(mk(4, 3), mk(4, 5), true, false, 1.0, 0.0));          // C not square
  ASSERT_DEATH(MatrixTrs<double>(mk(4, 4), mk(5, 3), true, true, false, false, 1.0));   // left: mB != mA
  ASSERT_DEATH(MatrixTrs<double>(mk(4, 5), mk(4, 3), true, true, false, false, 1.0));   // A not square
  MatrixSyrkAlloc<double> sa(mk(3, 7), false, true, 1.0, 4);
  ASSERT_EQ(sa.dims().m, 7u); ASSERT_EQ(sa.dims().ld, 8u); ASSERT_EQ(sa.output_space_req(), 56u);
}

TEST(SYRK_TRSM_Test, Facade_Planners_And_Plan_Cache) {
  ::rtat::rtat r;
  auto &sp = r.syrk_planner<double>();
  ASSERT_EQ(&sp, &r.syrk_planner<double>());
  auto &tp = r.trsm_planner<float>();
  ASSERT_NE((void *)&tp, (void *)&r.trsm_planner<double>());
  Json cache = Json::parse("[{\"key\": {\"uplo\":\"Upper\",\"trans\":\"N\",\"n\":4096,\"k\":512}, \"option\": {\"transA\":\"T\",\"transC\":\"F\"}}]");
  ASSERT_EQ(sp.load_plans(cache), 1u);
  ASSERT_EQ(std::string(sp.create_plan(SYRK_Key(gpu::BLAS_FILL_MODE_UPPER, gpu::BLAS_OP_N, 4096, 512))), std::string("TF"));
  ASSERT_EQ(sp.save_plans(), cache);
  Json tcache = Json::parse("[{\"key\": {\"side\":\"Right\",\"uplo\":\"Lower\",\"trans\":\"T\",\"diag\":\"Unit\",\"m\":100,\"n\":200}, \"option\": {\"swap_side\":\"T\",\"transA\":\"T\"}}]");
  ASSERT_EQ(tp.load_plans(tcache), 1u);
  ASSERT_EQ(std::string(tp.create_plan(TRSM_Key(gpu::BLAS_SIDE_RIGHT, gpu::BLAS_FILL_MODE_LOWER, gpu::BLAS_OP_T, gpu::BLAS_DIAG_UNIT, 100, 200))),
            std::string("TT"));
}

TEST(Facade_Test, Lazy_Planners) {
  ::rtat::rtat r;
  auto &d1 = r.gemm_planner<double>();
  auto &d2 = r.gemm_planner<double>();
  ASSERT_EQ(&d1, &d2);
  ASSERT_NE((void *)&r.gemm_planner<float>(), (void *)&d1);
  ASSERT_EQ(r.gemm_planner<double>().save_plans().dump(), std::string("[]"));
  // plan cache round trip through JSON without touching a device
  Json cache = Json::parse("[{\"key\": {\"transA\":\"T\",\"transB\":\"N\",\"m\":8192,\"n\":8192,\"k\":8192}, \"option\": {\"transA\":\"N\",\"transB\":\"N\",\"transC\":\"N\"}}]");
  ASSERT_EQ(d1.load_plans(cache), 1u);
  GEMM_Key key(gpu::BLAS_OP_T, gpu::BLAS_OP_N, 8192, 8192, 8192);
  ASSERT_EQ(std::string(d1.create_plan(key)), std::string("NNN"));  // served from the cache, no exploration
  ASSERT_EQ(d1.save_plans(), cache);
}

// The plan cache on disk (reference src/app/runner.h:81-82 keeps its planner output in a JSON file): what one process
// tuned, the next one loads, creates plans from without exploring, and writes back unchanged.
TEST(Plan_Cache_Test, File_Round_Trip) {
  char path[] = "/tmp/rtat_plan_cache_XXXXXX";
  int fd = mkstemp(path);
  ASSERT_TRUE(fd >= 0);
  close(fd);
  std::remove(path);
  GEMM_Key c2(gpu::BLAS_OP_T, gpu::BLAS_OP_N, 8192, 8192, 8192), c1(gpu::BLAS_OP_N, gpu::BLAS_OP_N, 1024, 1024, 1024);
  {
    GEMM_Planner first;
    ASSERT_EQ(first.load_plans_file(path), 0u);  // a missing file is an empty cache
    Json cache = Json::parse(
        "[{\"key\": {\"transA\":\"T\",\"transB\":\"N\",\"m\":8192,\"n\":8192,\"k\":8192}, \"option\": {\"transA\":\"T\",\"transB\":\"N\",\"transC\":\"N\"}},"
        " {\"key\": {\"transA\":\"N\",\"transB\":\"N\",\"m\":1024,\"n\":1024,\"k\":1024}, \"option\": {\"transA\":\"N\",\"transB\":\"N\",\"transC\":\"N\"}}]");
    ASSERT_EQ(first.load_plans(cache), 2u);
    first.save_plans_file(path);
  }
  GEMM_Planner second;
  ASSERT_EQ(second.load_plans_file(path), 2u);
  ASSERT_EQ(std::string(second.create_plan(c2)), std::string("TNN"));  // served from the file, no exploration
  ASSERT_EQ(std::string(second.create_plan(c1)), std::string("NNN"));
  second.save_plans_file(path);
  GEMM_Planner third;
  ASSERT_EQ(third.load_plans_file(path), 2u);
  ASSERT_EQ(third.save_plans(), second.save_plans());
  // a malformed file is an error, not an empty cache
  { std::ofstream bad(path, std::ios::trunc); bad << "[{\"key\": "; }
  bool threw = false;
  try { third.load_plans_file(path); } catch (const std::exception &) { threw = true; }
  ASSERT_TRUE(threw);
  std::remove(path);
}

// Statistics with the derived columns run_tests prints: TFLOP/s from Key::flopcount, GB/s from Key::bytecount; and the
// reference's null for an empty table / a key without plans (nlohmann default-constructed values, planner_statistics.h:50-67).
TEST(Statistics_Test, Rates_And_Empty_Forms) {
  using Stats = Planner_Statistics<GEMM_Key, GEMM_Options>;
  ASSERT_EQ(Stats({}).json().dump(), std::string("null"));
  std::map<GEMM_Key, std::map<GEMM_Options, std::vector<float>>> t;
  GEMM_Key key(gpu::BLAS_OP_N, gpu::BLAS_OP_N, 1024, 1024, 1024);
  t[key];
  ASSERT_EQ(Stats(t).json()[size_t(0)]["options"].dump(), std::string("null"));
  t[key][GEMM_Options()] = {0.5f, 1.0f};
  Json j = Stats(t).json(sizeof(double));
  const Json &plan = j[size_t(0)]["options"][size_t(0)];
  ASSERT_EQ(plan["times"].size(), 2u);
  ASSERT_NEAR(plan["tflops"][size_t(1)].get<double>(), 2.0 * 1024 * 1024 * 1024 / 1e-3 * 1e-12, 1e-9);
  ASSERT_NEAR(plan["gbs"][size_t(0)].get<double>(), 3.0 * 1024 * 1024 * 8 / 0.5e-3 * 1e-9, 1e-6);
  ASSERT_TRUE(!Stats(t).json()[size_t(0)]["options"][size_t(0)].contains("tflops"));  // plain form = the reference's schema
  // the reference's default: strict minimum of the means unless a margin is asked for
  auto rates = Stats(t).get_floprates();
  ASSERT_EQ(rates.at(key).at(GEMM_Options()).size(), 2u);
}

int main(int argc, char **argv) { return mini::run_all(argc, argv); }
