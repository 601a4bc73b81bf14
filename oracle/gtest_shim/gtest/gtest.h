// TEST INFRASTRUCTURE.  A small stand-in for GoogleTest (the reference fetches gtest 1.14 from the network at
// configure time, reference CMakeLists.txt:72-79; there is no network here).  It implements exactly the
// subset the reference's tests/*.cu use, so those files compile UNCHANGED against this repository's headers
// and libqsim_b200.so (oracle/Makefile target `reftests`): TEST, TEST_F, ::testing::Test,
// EXPECT_/ASSERT_ {EQ,NE,LT,LE,GT,GE,TRUE,FALSE,NEAR,DOUBLE_EQ,FLOAT_EQ,THROW,NO_THROW}, SCOPED_TRACE,
// ::testing::AssertionResult, message streaming, and a main().
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
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
    struct Entry { std::string suite, name; std::function<Test*()> make; };
    std::vector<Entry> tests;
    std::vector<std::string> traces;
    int failures_in_current = 0;
    bool fatal = false;
    static Registry& get() { static Registry r; return r; }
};

struct Registrar {
    Registrar(const char* suite, const char* name, std::function<Test*()> make) {
        Registry::get().tests.push_back({suite, name, std::move(make)});
    }
};

class AssertionResult {
public:
    explicit AssertionResult(bool ok) : ok_(ok) {}
    AssertionResult(const AssertionResult& o) : ok_(o.ok_), msg_(o.msg_.str()) {}
    explicit operator bool() const { return ok_; }
    template <class T> AssertionResult& operator<<(const T& v) { msg_ << v; return *this; }
    std::string message() const { return msg_.str(); }
private:
    bool ok_;
    std::ostringstream msg_;
};
inline AssertionResult AssertionSuccess() { return AssertionResult(true); }
inline AssertionResult AssertionFailure() { return AssertionResult(false); }

// Reports a failure when destroyed; collects the user's streamed message.
class Reporter {
public:
    Reporter(const char* file, int line, const std::string& what, bool fatal) : fatal_(fatal) {
        head_ << file << ":" << line << ": Failure\n" << what;
    }
    Reporter(const Reporter& o) : fatal_(o.fatal_) { head_ << o.head_.str(); user_ << o.user_.str(); }
    ~Reporter() {
        Registry& r = Registry::get();
        std::cout << head_.str();
        if (!user_.str().empty()) std::cout << "\n" << user_.str();
        for (const auto& t : r.traces) std::cout << "\n  trace: " << t;
        std::cout << std::endl;
        r.failures_in_current++;
        if (fatal_) r.fatal = true;
    }
    template <class T> Reporter& operator<<(const T& v) { user_ << v; return *this; }
private:
    bool fatal_;
    std::ostringstream head_, user_;
};

// `ASSERT_x(...) << msg;` must be able to `return` from the test body: the classic void-assignment trick.
struct Voidify { void operator=(const Reporter&) const {} };

struct ScopedTrace {
    explicit ScopedTrace(const std::string& s) { Registry::get().traces.push_back(s); }
    ~ScopedTrace() { Registry::get().traces.pop_back(); }
};

template <class T>
auto print_value(std::ostream& os, const T& v, int) -> decltype(os << v, void()) { os << v; }
template <class T>
void print_value(std::ostream& os, const T&, long) { os << "<" << sizeof(T) << "-byte object>"; }

template <class A, class B>
std::string cmp_text(const char* ea, const char* eb, const A& a, const B& b, const char* op) {
    std::ostringstream os;
    os << "Expected: (" << ea << ") " << op << " (" << eb << "), actual: ";
    print_value(os, a, 0);
    os << " vs ";
    print_value(os, b, 0);
    return os.str();
}

inline bool almost_equal(double a, double b) {   // 4-ULP comparison like gtest's EXPECT_DOUBLE_EQ
    if (a == b) return true;
    if (std::isnan(a) || std::isnan(b)) return false;
    long long ia, ib;
    std::memcpy(&ia, &a, 8); std::memcpy(&ib, &b, 8);
    if (ia < 0) ia = (long long)0x8000000000000000ULL - ia;
    if (ib < 0) ib = (long long)0x8000000000000000ULL - ib;
    const long long d = ia > ib ? ia - ib : ib - ia;
    return d <= 4;
}
inline bool almost_equal_f(float a, float b) {
    if (a == b) return true;
    int ia, ib;
    std::memcpy(&ia, &a, 4); std::memcpy(&ib, &b, 4);
    if (ia < 0) ia = (int)0x80000000u - ia;
    if (ib < 0) ib = (int)0x80000000u - ib;
    const long long d = (long long)ia - ib;
    return (d < 0 ? -d : d) <= 4;
}

inline int RunAllTests() {
    Registry& r = Registry::get();
    int failed = 0, ran = 0;
    std::vector<std::string> failed_names;
    for (auto& e : r.tests) {
        std::cout << "[ RUN      ] " << e.suite << "." << e.name << std::endl;
        r.failures_in_current = 0;
        r.fatal = false;
        try {
            Test* t = e.make();
            t->SetUp();
            if (!r.fatal) t->TestBody();
            t->TearDown();
            delete t;
        } catch (const std::exception& ex) {
            std::cout << "unexpected exception: " << ex.what() << std::endl;
            r.failures_in_current++;
        } catch (...) {
            std::cout << "unexpected non-standard exception" << std::endl;
            r.failures_in_current++;
        }
        ++ran;
        if (r.failures_in_current) { ++failed; failed_names.push_back(e.suite + "." + e.name); std::cout << "[  FAILED  ] "; }
        else std::cout << "[       OK ] ";
        std::cout << e.suite << "." << e.name << std::endl;
    }
    std::cout << "[==========] " << ran << " tests ran.\n[  PASSED  ] " << (ran - failed) << " tests." << std::endl;
    if (failed) {
        std::cout << "[  FAILED  ] " << failed << " tests:" << std::endl;
        for (auto& n : failed_names) std::cout << "[  FAILED  ] " << n << std::endl;
    }
    return failed ? 1 : 0;
}

inline void InitGoogleTest(int*, char**) {}

}  // namespace testing

#define GTEST_CLASS_(suite, name) suite##_##name##_Test
#define GTEST_DEFINE_(suite, name, base)                                                                     \
    class GTEST_CLASS_(suite, name) : public base {                                                          \
    public:                                                                                                  \
        void TestBody() override;                                                                            \
    };                                                                                                       \
    static ::testing::Registrar suite##_##name##_registrar(#suite, #name,                                    \
                                                           [] { return static_cast<::testing::Test*>(new GTEST_CLASS_(suite, name)); }); \
    void GTEST_CLASS_(suite, name)::TestBody()
#define TEST(suite, name) GTEST_DEFINE_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTEST_DEFINE_(fixture, name, fixture)

#define GTEST_NONFATAL_(cond, text) \
    if (cond) ; else ::testing::Reporter(__FILE__, __LINE__, text, false)
#define GTEST_FATAL_(cond, text) \
    if (cond) ; else return ::testing::Voidify() = ::testing::Reporter(__FILE__, __LINE__, text, true)

#define GTEST_CMP_(macro, a, b, op) macro((a)op(b), ::testing::cmp_text(#a, #b, (a), (b), #op))
#define EXPECT_EQ(a, b) GTEST_CMP_(GTEST_NONFATAL_, a, b, ==)
#define EXPECT_NE(a, b) GTEST_CMP_(GTEST_NONFATAL_, a, b, !=)
#define EXPECT_LT(a, b) GTEST_CMP_(GTEST_NONFATAL_, a, b, <)
#define EXPECT_LE(a, b) GTEST_CMP_(GTEST_NONFATAL_, a, b, <=)
#define EXPECT_GT(a, b) GTEST_CMP_(GTEST_NONFATAL_, a, b, >)
#define EXPECT_GE(a, b) GTEST_CMP_(GTEST_NONFATAL_, a, b, >=)
#define ASSERT_EQ(a, b) GTEST_CMP_(GTEST_FATAL_, a, b, ==)
#define ASSERT_NE(a, b) GTEST_CMP_(GTEST_FATAL_, a, b, !=)
#define ASSERT_LT(a, b) GTEST_CMP_(GTEST_FATAL_, a, b, <)
#define ASSERT_LE(a, b) GTEST_CMP_(GTEST_FATAL_, a, b, <=)
#define ASSERT_GT(a, b) GTEST_CMP_(GTEST_FATAL_, a, b, >)
#define ASSERT_GE(a, b) GTEST_CMP_(GTEST_FATAL_, a, b, >=)

#define GTEST_BOOL_(macro, expr, want)                                                        \
    macro(static_cast<bool>(expr) == want, std::string("Value of: " #expr "\n  Expected: ") + (want ? "true" : "false"))
#define EXPECT_TRUE(e) GTEST_BOOL_(GTEST_NONFATAL_, e, true)
#define EXPECT_FALSE(e) GTEST_BOOL_(GTEST_NONFATAL_, e, false)
#define ASSERT_TRUE(e) GTEST_BOOL_(GTEST_FATAL_, e, true)
#define ASSERT_FALSE(e) GTEST_BOOL_(GTEST_FATAL_, e, false)

#define GTEST_NEAR_(macro, a, b, tol) \
    macro(std::fabs(double(a) - double(b)) <= double(tol), ::testing::cmp_text(#a, #b, (a), (b), "~="))
#define EXPECT_NEAR(a, b, tol) GTEST_NEAR_(GTEST_NONFATAL_, a, b, tol)
#define ASSERT_NEAR(a, b, tol) GTEST_NEAR_(GTEST_FATAL_, a, b, tol)
#define EXPECT_DOUBLE_EQ(a, b) GTEST_NONFATAL_(::testing::almost_equal((a), (b)), ::testing::cmp_text(#a, #b, (a), (b), "=="))
#define ASSERT_DOUBLE_EQ(a, b) GTEST_FATAL_(::testing::almost_equal((a), (b)), ::testing::cmp_text(#a, #b, (a), (b), "=="))
#define EXPECT_FLOAT_EQ(a, b) GTEST_NONFATAL_(::testing::almost_equal_f((a), (b)), ::testing::cmp_text(#a, #b, (a), (b), "=="))
#define ASSERT_FLOAT_EQ(a, b) GTEST_FATAL_(::testing::almost_equal_f((a), (b)), ::testing::cmp_text(#a, #b, (a), (b), "=="))

#define GTEST_THROW_(macro, stmt, extype)                                                  \
    macro(([&]() -> bool { try { stmt; } catch (const extype&) { return true; } catch (...) { return false; } return false; })(), \
          "Expected: " #stmt " throws " #extype)
#define EXPECT_THROW(stmt, extype) GTEST_THROW_(GTEST_NONFATAL_, stmt, extype)
#define ASSERT_THROW(stmt, extype) GTEST_THROW_(GTEST_FATAL_, stmt, extype)
#define GTEST_NO_THROW_(macro, stmt) \
    macro(([&]() -> bool { try { stmt; } catch (...) { return false; } return true; })(), "Expected: " #stmt " does not throw")
#define EXPECT_NO_THROW(stmt) GTEST_NO_THROW_(GTEST_NONFATAL_, stmt)
#define ASSERT_NO_THROW(stmt) GTEST_NO_THROW_(GTEST_FATAL_, stmt)

#define GTEST_CAT2_(a, b) a##b
#define GTEST_CAT_(a, b) GTEST_CAT2_(a, b)
#define SCOPED_TRACE(msg)                                                     \
    std::ostringstream GTEST_CAT_(gtest_trace_os_, __LINE__);                 \
    GTEST_CAT_(gtest_trace_os_, __LINE__) << msg;                             \
    ::testing::ScopedTrace GTEST_CAT_(gtest_trace_, __LINE__)(GTEST_CAT_(gtest_trace_os_, __LINE__).str())

#define RUN_ALL_TESTS() ::testing::RunAllTests()
