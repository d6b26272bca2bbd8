// TEST INFRASTRUCTURE — not part of the product path.
//
// Link-time stand-ins for the three pieces of libvvhd that cannot be built in
// this image (no Lua, no HDF5, no system LAPACK), so that the reference's own
// hot-path translation units
//   libvvhd/src/{TSortedTree,MEpsilonFast,MConvectiveFast,MDiffusiveFast,
//                MFlowmove,TBody,TMatrix,TSpace}.cpp
// can be compiled UNMODIFIED from /root/reference and linked into
// oracle/_ref/libvvref.so (see oracle/Makefile). Nothing here restates any
// hot-path arithmetic.
//
//  * TEval  (reference: libvvhd/headers/TEval.hpp:5-29, src/TEval.cpp) —
//    the reference evaluates "f(t)" strings with an embedded Lua state. Every
//    BASELINE configuration uses constant expressions ("1", "0", ""), so this
//    stand-in parses the string with strtod once; empty string -> 0, exactly
//    as TEval.cpp:167-168 returns 0 for an empty expression.
//  * HDF5 entry points used by TSpace.cpp:46-82 and the three Space
//    load/save members defined in TSpace_{save,load}_hdf.cpp/TSpace_load_v13.cpp:
//    abort() — the oracle drivers never do file I/O through Space.
//  * cblas_dgemv/dgesv_/dgetrf_/dgetri_ (TMatrix.cpp:13-15,176,207,243,250):
//    forwarded to scipy's bundled OpenBLAS, looked up with dlopen the first
//    time the body SLAE is solved (only the cylinder fixtures need it).

#include "TEval.hpp"
#include "TSpace.hpp"
#include "elementary.h"
#include "cblas.h"
#include <hdf5.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <dlfcn.h>
#include <glob.h>

const char* libvvhd_gitrev = "oracle-shim";
const char* libvvhd_gitinfo = "oracle-shim";
const char* libvvhd_gitdiff = "";

/* ------------------------------- TEval ---------------------------------- */

static double parse_const(const std::string& s) {
    if (s.empty()) return 0;
    char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    while (end && *end == ' ') end++;
    if (!end || *end != 0)
        throw std::invalid_argument("oracle TEval stand-in: only constant expressions are supported: '" + s + "'");
    return v;
}

TEval::TEval(): lua_state(nullptr), expr(), cacheTime1(), cacheTime2(), cacheValue1(), cacheValue2() {}
TEval::TEval(const std::string& str): TEval() { *this = str; }
TEval::TEval(const TEval& copy): TEval() { *this = copy; }
TEval& TEval::operator=(const std::string& str) { cacheValue1 = parse_const(str); expr = str; return *this; }
TEval& TEval::operator=(const TEval& copy) { expr = copy.expr; cacheValue1 = copy.cacheValue1; return *this; }
TEval::~TEval() {}
double TEval::eval(double) const { return expr.empty() ? 0 : cacheValue1; }

/* ------------------------------- HDF5 ----------------------------------- */

[[noreturn]] static void no_hdf5(const char* what) {
    fprintf(stderr, "oracle/_ref: %s needs HDF5, which this image does not have\n", what);
    abort();
}
extern "C" {
herr_t H5Eset_auto(hid_t, void*, void*) { no_hdf5("H5Eset_auto"); }
htri_t H5Fis_hdf5(const char*) { no_hdf5("H5Fis_hdf5"); }
hid_t H5Fopen(const char*, unsigned, hid_t) { no_hdf5("H5Fopen"); }
hid_t H5Fcreate(const char*, unsigned, hid_t, hid_t) { no_hdf5("H5Fcreate"); }
herr_t H5Fclose(hid_t) { no_hdf5("H5Fclose"); }
}
void h5_throw(std::string fn, std::string arg) { throw std::runtime_error(fn + "(" + arg + ")"); }
void Space::save_hdf(int64_t) { no_hdf5("Space::save_hdf"); }
void Space::load_hdf(int64_t, metainfo_t*) { no_hdf5("Space::load_hdf"); }
void Space::load_v13(const char*) { no_hdf5("Space::load_v13"); }

/* --------------------------- BLAS / LAPACK ------------------------------ */

typedef void (*dgemv_fn)(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, double, const double*, int,
                         const double*, int, double, double*, int);
typedef void (*dgesv_fn)(int*, const int*, double*, int*, int*, double*, int*, int*);
typedef void (*dgetrf_fn)(int*, int*, double*, int*, int*, int*);
typedef void (*dgetri_fn)(int*, double*, int*, int*, double*, int*, int*);

static void* blas_handle() {
    static void* h = nullptr;
    if (h) return h;
    const char* env = getenv("VVREF_OPENBLAS");
    if (env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        const char* patterns[] = {
            "/opt/prime-rl/.venv/lib/python3*/site-packages/scipy.libs/libscipy_openblas*.so",
            "/usr/lib/python3*/site-packages/scipy.libs/libscipy_openblas*.so",
        };
        for (const char* pat : patterns) {
            glob_t g;
            if (glob(pat, 0, nullptr, &g) == 0) {
                for (size_t i = 0; i < g.gl_pathc && !h; i++) h = dlopen(g.gl_pathv[i], RTLD_NOW | RTLD_LOCAL);
            }
            globfree(&g);
            if (h) break;
        }
    }
    if (!h) {
        fprintf(stderr, "oracle/_ref: cannot find scipy's OpenBLAS for the body SLAE (set VVREF_OPENBLAS)\n");
        abort();
    }
    return h;
}
static void* blas_sym(const char* a, const char* b) {
    void* s = dlsym(blas_handle(), a);
    if (!s) s = dlsym(blas_handle(), b);
    if (!s) { fprintf(stderr, "oracle/_ref: OpenBLAS lacks %s\n", a); abort(); }
    return s;
}
extern "C" {
void cblas_dgemv(CBLAS_ORDER o, CBLAS_TRANSPOSE t, int M, int N, double alpha, const double* A, int lda,
                 const double* X, int incX, double beta, double* Y, int incY) {
    static dgemv_fn f = (dgemv_fn)blas_sym("scipy_cblas_dgemv", "cblas_dgemv");
    f(o, t, M, N, alpha, A, lda, X, incX, beta, Y, incY);
}
void dgesv_(int* n, const int* nrhs, double* a, int* lda, int* ipiv, double* x, int* incx, int* info) {
    static dgesv_fn f = (dgesv_fn)blas_sym("scipy_dgesv_", "dgesv_");
    f(n, nrhs, a, lda, ipiv, x, incx, info);
}
void dgetrf_(int* M, int* N, double* A, int* lda, int* IPIV, int* INFO) {
    static dgetrf_fn f = (dgetrf_fn)blas_sym("scipy_dgetrf_", "dgetrf_");
    f(M, N, A, lda, IPIV, INFO);
}
void dgetri_(int* N, double* A, int* lda, int* IPIV, double* WORK, int* lwork, int* INFO) {
    static dgetri_fn f = (dgetri_fn)blas_sym("scipy_dgetri_", "dgetri_");
    f(N, A, lda, IPIV, WORK, lwork, INFO);
}
}
