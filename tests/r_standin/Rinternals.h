/* Stand-in for R's <Rinternals.h>: TEST INFRASTRUCTURE ONLY (this image has no R).
 *
 * Just enough of R's C API for r/bsfg_glue.c to compile UNMODIFIED and run: an SEXP is a pointer to a tagged struct
 * (type, length, dim, data, names), numeric vectors/matrices are column-major double arrays exactly as in R, `.Call`
 * is rs_call() (which plays the role of R's error context: Rf_error() long-jumps back into it).  The Python harness
 * (tests/test_r_glue*.py) builds the argument objects through the rs_* helpers at the bottom, wrapping NumPy buffers
 * without copying, the way R hands its own vectors to .Call.  Semantics follow "Writing R Extensions" section 5.9
 * for the entry points listed here; nothing else of R is modelled.
 */
#ifndef RSTANDIN_RINTERNALS_H
#define RSTANDIN_RINTERNALS_H
#include <stddef.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

enum { NILSXP = 0, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22 };

typedef struct rs_obj {
    int type;
    R_xlen_t len;
    int nrow, ncol;            /* -1: no dim attribute */
    void* data;                /* double* / int* / SEXP* / char* / external address */
    struct rs_obj* names;      /* STRSXP or NULL */
    void (*finalizer)(struct rs_obj*);
    int owns;                  /* data allocated by the stand-in (freed by rs_reset) */
} *SEXP;

extern SEXP R_NilValue;
extern SEXP R_NamesSymbol;

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))

double* REAL(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
const char* CHAR(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int LENGTH(SEXP x);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
SEXP Rf_getAttrib(SEXP x, SEXP what);
int Rf_isNull(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
double Rf_asReal(SEXP x);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_mkNamed(int type, const char** names);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarReal(double v);
char* R_alloc(size_t n, int size);
#if defined(__GNUC__)
__attribute__((noreturn, format(printf, 1, 2)))
#endif
void Rf_error(const char* fmt, ...);

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);

/* ---- harness side (called from Python through ctypes) ---- */
SEXP rs_nil(void);
SEXP rs_real_view(double* p, R_xlen_t len, int nrow, int ncol);   /* wraps caller-owned memory; nrow = -1: plain vector */
SEXP rs_int_vector(const int* v, R_xlen_t len);
SEXP rs_lgl_scalar(int v);
SEXP rs_real_scalar(double v);
SEXP rs_list(int n, const char** names, SEXP* elems);             /* named VECSXP */
int rs_type(SEXP x);
void rs_dims(SEXP x, long long* len, int* nrow, int* ncol);
SEXP rs_call(void* fn, int nargs, SEXP* args);                    /* .Call: NULL + rs_last_error() on Rf_error */
const char* rs_last_error(void);
void rs_run_finalizer(SEXP extptr);                               /* what gc() does to an unreachable external pointer */
void rs_reset(void);                                              /* frees everything the stand-in allocated */
void rs_poison_stack(void);                                       /* fills ~64 KB of stack below the caller with 0xAB */

#ifdef __cplusplus
}
#endif
#endif
