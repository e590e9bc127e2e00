// Minimal stand-in for the part of GoogleTest that the reference's test sources use
// (TEST, TEST_F, ::testing::Test with SetUp, EXPECT_TRUE / EXPECT_NEAR / EXPECT_DOUBLE_EQ,
// InitGoogleTest, RUN_ALL_TESTS), so that tests/src/long_term_planner_tests.cc of
// yannickBurkhardt/LongTermPlanner compiles UNCHANGED against this repository's drop-in
// header (GoogleTest is not installed in the build container). Test infrastructure only.
#ifndef LTP_GTEST_SHIM_H
#define LTP_GTEST_SHIM_H

#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

namespace testing {

class Test {
 public:
  virtual ~Test() {}
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};

struct Registry {
  struct Entry {
    std::string name;
    std::function<Test*()> make;
  };
  std::vector<Entry> tests;
  long failures = 0;        // failed assertions in the current test
  long total_failures = 0;
  long checks = 0;
  static Registry& get() {
    static Registry r;
    return r;
  }
};

struct Registrar {
  Registrar(const char* suite, const char* name, std::function<Test*()> make) {
    Registry::get().tests.push_back({std::string(suite) + "." + name, make});
  }
};

inline void InitGoogleTest(int*, char**) {}

inline void report(const char* file, int line, const std::string& msg) {
  Registry& r = Registry::get();
  r.failures++;
  if (r.failures <= 5) std::printf("%s:%d: Failure\n  %s\n", file, line, msg.c_str());
}

// EXPECT_* are usable as statements and swallow a trailing `<< ...`
struct Sink {
  template <class T>
  Sink& operator<<(const T&) { return *this; }
};

}  // namespace testing

inline int RUN_ALL_TESTS() {
  testing::Registry& r = testing::Registry::get();
  int failed_tests = 0;
  const char* filter = std::getenv("LTP_GTEST_FILTER");
  for (auto& e : r.tests) {
    if (filter && e.name.find(filter) == std::string::npos) continue;
    std::printf("[ RUN      ] %s\n", e.name.c_str());
    std::fflush(stdout);
    r.failures = 0;
    testing::Test* t = e.make();
    t->SetUp();
    t->TestBody();
    t->TearDown();
    delete t;
    r.total_failures += r.failures;
    if (r.failures) {
      failed_tests++;
      std::printf("[  FAILED  ] %s (%ld failed expectations)\n", e.name.c_str(), r.failures);
    } else {
      std::printf("[       OK ] %s\n", e.name.c_str());
    }
  }
  std::printf("[==========] %zu tests, %ld expectations checked, %d tests failed\n", r.tests.size(), r.checks,
              failed_tests);
  return failed_tests ? 1 : 0;
}

#define LTP_GTEST_CLASS(suite, name) suite##_##name##_Test

#define LTP_GTEST_DEFINE(suite, name, base)                                                      \
  class LTP_GTEST_CLASS(suite, name) : public base {                                             \
   public:                                                                                       \
    void TestBody() override;                                                                    \
  };                                                                                             \
  static ::testing::Registrar suite##_##name##_registrar(                                        \
      #suite, #name, [] { return static_cast<::testing::Test*>(new LTP_GTEST_CLASS(suite, name)); }); \
  void LTP_GTEST_CLASS(suite, name)::TestBody()

#define TEST(suite, name) LTP_GTEST_DEFINE(suite, name, ::testing::Test)
#define TEST_F(fixture, name) LTP_GTEST_DEFINE(fixture, name, fixture)

#define EXPECT_TRUE(cond)                                                                        \
  do {                                                                                           \
    ::testing::Registry::get().checks++;                                                         \
    if (!(cond)) ::testing::report(__FILE__, __LINE__, std::string("Expected true: ") + #cond);  \
  } while (0)

#define EXPECT_NEAR(a, b, tol)                                                                   \
  do {                                                                                           \
    ::testing::Registry::get().checks++;                                                         \
    const double a_ = (a), b_ = (b), t_ = (tol);                                                 \
    if (!(std::fabs(a_ - b_) <= t_))                                                             \
      ::testing::report(__FILE__, __LINE__,                                                      \
                        std::string(#a " vs " #b ": ") + std::to_string(a_) + " vs " +           \
                            std::to_string(b_) + " (tol " + std::to_string(t_) + ")");           \
  } while (0)

#define EXPECT_DOUBLE_EQ(a, b) EXPECT_NEAR(a, b, 4 * 2.220446049250313e-16 * std::fabs((double)(b)))

#endif  // LTP_GTEST_SHIM_H
