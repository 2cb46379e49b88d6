// mini_test.h — a few macros standing in for GoogleTest (not available in this image), so the
// restated reference suites keep their shape: TEST(suite, name) { ASSERT_*/EXPECT_* ... }.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <string>
#include <vector>
#include <sys/wait.h>
#include <unistd.h>

namespace mini {
struct Case { std::string name; std::function<void()> fn; };
inline std::vector<Case> &registry() { static std::vector<Case> r; return r; }
inline int &failures() { static int f = 0; return f; }
struct Registrar { Registrar(const char *n, std::function<void()> f) { registry().push_back({n, std::move(f)}); } };
struct Abort {};
inline void fail(const char *file, int line, const std::string &what, bool fatal) {
  std::printf("  FAILED %s:%d: %s\n", file, line, what.c_str());
  failures()++;
  if (fatal) throw Abort();
}
// runs f in a forked child; true when the child died abnormally (signal or non-zero exit)
template <typename F>
bool dies(F f) {
  fflush(stdout);
  pid_t pid = fork();
  if (pid == 0) {
    if (!freopen("/dev/null", "w", stdout) || !freopen("/dev/null", "w", stderr)) _exit(4);
    try { f(); } catch (...) { _exit(3); }
    _exit(0);
  }
  int status = 0;
  waitpid(pid, &status, 0);
  return !(WIFEXITED(status) && WEXITSTATUS(status) == 0);
}
inline int run_all(int argc, char **argv) {
  const char *filter = argc > 1 ? argv[1] : nullptr;
  int ran = 0;
  for (auto &c : registry()) {
    if (filter && c.name.find(filter) == std::string::npos) continue;
    int before = failures();
    std::printf("[ RUN  ] %s\n", c.name.c_str());
    fflush(stdout);
    try { c.fn(); } catch (Abort &) {} catch (const std::exception &e) { fail("?", 0, std::string("exception: ") + e.what(), false); } catch (const char *m) { fail("?", 0, std::string("thrown: ") + m, false); }
    std::printf("[ %s ] %s\n", failures() == before ? " OK " : "FAIL", c.name.c_str());
    ran++;
  }
  std::printf("%d tests, %d failures\n", ran, failures());
  return failures() ? 1 : 0;
}
}  // namespace mini

#define TEST(suite, name)                                                   \
  static void suite##_##name();                                             \
  static mini::Registrar reg_##suite##_##name(#suite "." #name, suite##_##name); \
  static void suite##_##name()
#define MINI_CHECK(cond, text, fatal) do { if (!(cond)) mini::fail(__FILE__, __LINE__, text, fatal); } while (0)
#define ASSERT_TRUE(x) MINI_CHECK((x), #x, true)
#define EXPECT_TRUE(x) MINI_CHECK((x), #x, false)
#define ASSERT_FALSE(x) MINI_CHECK(!(x), "!(" #x ")", true)
#define ASSERT_EQ(a, b) MINI_CHECK((a) == (b), #a " == " #b, true)
#define EXPECT_EQ(a, b) MINI_CHECK((a) == (b), #a " == " #b, false)
#define ASSERT_NE(a, b) MINI_CHECK(!((a) == (b)), #a " != " #b, true)
#define ASSERT_NEAR(a, b, tol) MINI_CHECK(std::fabs(double(a) - double(b)) <= (tol), #a " ~= " #b, true)
#define ASSERT_DEATH(stmt) MINI_CHECK(mini::dies([&] { stmt; }), "expected death: " #stmt, true)
#define ASSERT_THROWS(stmt) do { bool threw_ = false; try { stmt; } catch (...) { threw_ = true; } MINI_CHECK(threw_, "expected throw: " #stmt, true); } while (0)
