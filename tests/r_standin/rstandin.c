/* Implementation of the R API stand-in declared in Rinternals.h (test infrastructure only). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "Rinternals.h"

static struct rs_obj nil_obj = {NILSXP, 0, -1, -1, NULL, NULL, NULL, 0};
static struct rs_obj names_sym = {NILSXP, 0, -1, -1, NULL, NULL, NULL, 0};
SEXP R_NilValue = &nil_obj;
SEXP R_NamesSymbol = &names_sym;

static jmp_buf rs_jmp;
static int rs_jmp_armed = 0;
static char rs_err[1024];

/* every allocation is remembered so rs_reset() can release it */
static void** rs_allocs = NULL;
static size_t rs_nalloc = 0, rs_cap = 0;
static void* rs_malloc(size_t n) {
    void* p = calloc(1, n ? n : 1);
    if (!p) { fprintf(stderr, "rstandin: out of memory\n"); abort(); }
    if (rs_nalloc == rs_cap) {
        rs_cap = rs_cap ? 2 * rs_cap : 256;
        rs_allocs = (void**)realloc(rs_allocs, rs_cap * sizeof(void*));
    }
    rs_allocs[rs_nalloc++] = p;
    return p;
}
void rs_reset(void) {
    for (size_t i = 0; i < rs_nalloc; i++) free(rs_allocs[i]);
    rs_nalloc = 0;
}

static SEXP rs_new(int type, R_xlen_t len, int nrow, int ncol, size_t elem) {
    SEXP x = (SEXP)rs_malloc(sizeof(struct rs_obj));
    x->type = type; x->len = len; x->nrow = nrow; x->ncol = ncol;
    x->data = elem ? rs_malloc(elem * (size_t)(len > 0 ? len : 1)) : NULL;
    x->owns = 1;
    return x;
}
static void need(SEXP x, int type, const char* what) {
    if (!x || x->type != type) Rf_error("%s applied to an object of type %d", what, x ? x->type : -1);
}

double* REAL(SEXP x) { need(x, REALSXP, "REAL()"); return (double*)x->data; }
int* INTEGER(SEXP x) {
    if (!x || (x->type != INTSXP && x->type != LGLSXP)) Rf_error("INTEGER() applied to an object of type %d", x ? x->type : -1);
    return (int*)x->data;
}
int* LOGICAL(SEXP x) { need(x, LGLSXP, "LOGICAL()"); return (int*)x->data; }
const char* CHAR(SEXP x) { need(x, CHARSXP, "CHAR()"); return (const char*)x->data; }
R_xlen_t XLENGTH(SEXP x) { return x ? x->len : 0; }
int LENGTH(SEXP x) { return (int)XLENGTH(x); }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) {
    need(x, VECSXP, "VECTOR_ELT()");
    if (i < 0 || i >= x->len) Rf_error("VECTOR_ELT: index out of range");
    SEXP e = ((SEXP*)x->data)[i];
    return e ? e : R_NilValue;
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) {
    need(x, VECSXP, "SET_VECTOR_ELT()");
    if (i < 0 || i >= x->len) Rf_error("SET_VECTOR_ELT: index out of range");
    ((SEXP*)x->data)[i] = v;
    return v;
}
SEXP STRING_ELT(SEXP x, R_xlen_t i) {
    need(x, STRSXP, "STRING_ELT()");
    if (i < 0 || i >= x->len) Rf_error("STRING_ELT: index out of range");
    return ((SEXP*)x->data)[i];
}
SEXP Rf_getAttrib(SEXP x, SEXP what) {
    if (what == R_NamesSymbol && x && x->names) return x->names;
    return R_NilValue;
}
int Rf_isNull(SEXP x) { return !x || x->type == NILSXP; }
int Rf_nrows(SEXP x) {
    if (!x || x->type == NILSXP) Rf_error("object is not a matrix");
    return x->nrow >= 0 ? x->nrow : (int)x->len;      /* R: a plain vector has nrows = length */
}
int Rf_ncols(SEXP x) {
    if (!x || x->type == NILSXP) Rf_error("object is not a matrix");
    return x->ncol >= 0 ? x->ncol : 1;
}
double Rf_asReal(SEXP x) {
    if (x && x->len >= 1) {
        if (x->type == REALSXP) return ((double*)x->data)[0];
        if (x->type == INTSXP || x->type == LGLSXP) return (double)((int*)x->data)[0];
    }
    Rf_error("asReal: not a numeric scalar");
}
int Rf_asInteger(SEXP x) {
    if (x && x->len >= 1) {
        if (x->type == REALSXP) return (int)((double*)x->data)[0];
        if (x->type == INTSXP || x->type == LGLSXP) return ((int*)x->data)[0];
    }
    Rf_error("asInteger: not a numeric scalar");
}
int Rf_asLogical(SEXP x) { return Rf_asInteger(x) != 0; }
SEXP Rf_allocVector(int type, R_xlen_t n) {
    switch (type) {
        case REALSXP: return rs_new(type, n, -1, -1, sizeof(double));
        case INTSXP: case LGLSXP: return rs_new(type, n, -1, -1, sizeof(int));
        case VECSXP: case STRSXP: return rs_new(type, n, -1, -1, sizeof(SEXP));
        default: Rf_error("allocVector: type %d not modelled", type);
    }
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
    SEXP x = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    x->nrow = nrow; x->ncol = ncol;
    return x;
}
static SEXP rs_mkchar(const char* s) {
    SEXP c = rs_new(CHARSXP, (R_xlen_t)strlen(s), -1, -1, 0);
    c->data = rs_malloc(strlen(s) + 1);
    strcpy((char*)c->data, s);
    return c;
}
SEXP Rf_mkNamed(int type, const char** names) {
    int n = 0;
    while (names[n][0]) n++;
    SEXP x = Rf_allocVector(type, n);
    x->names = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; i++) ((SEXP*)x->names->data)[i] = rs_mkchar(names[i]);
    return x;
}
SEXP Rf_ScalarInteger(int v) { SEXP x = Rf_allocVector(INTSXP, 1); ((int*)x->data)[0] = v; return x; }
SEXP Rf_ScalarReal(double v) { SEXP x = Rf_allocVector(REALSXP, 1); ((double*)x->data)[0] = v; return x; }
char* R_alloc(size_t n, int size) { return (char*)rs_malloc(n * (size_t)size); }

void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(rs_err, sizeof rs_err, fmt, ap);
    va_end(ap);
    if (!rs_jmp_armed) { fprintf(stderr, "rstandin: Rf_error outside rs_call: %s\n", rs_err); abort(); }
    longjmp(rs_jmp, 1);
}

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP x = rs_new(EXTPTRSXP, 0, -1, -1, 0);
    x->data = p;
    return x;
}
void* R_ExternalPtrAddr(SEXP s) { need(s, EXTPTRSXP, "R_ExternalPtrAddr()"); return s->data; }
void R_ClearExternalPtr(SEXP s) { need(s, EXTPTRSXP, "R_ClearExternalPtr()"); s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) {
    (void)onexit;
    need(s, EXTPTRSXP, "R_RegisterCFinalizerEx()");
    s->finalizer = fun;
}

/* ---- harness side ---- */
SEXP rs_nil(void) { return R_NilValue; }
SEXP rs_real_view(double* p, R_xlen_t len, int nrow, int ncol) {
    SEXP x = rs_new(REALSXP, len, nrow, ncol, 0);
    x->data = p; x->owns = 0;
    return x;
}
SEXP rs_int_vector(const int* v, R_xlen_t len) {
    SEXP x = Rf_allocVector(INTSXP, len);
    memcpy(x->data, v, sizeof(int) * (size_t)len);
    return x;
}
SEXP rs_lgl_scalar(int v) { SEXP x = Rf_allocVector(LGLSXP, 1); ((int*)x->data)[0] = v; return x; }
SEXP rs_real_scalar(double v) { return Rf_ScalarReal(v); }
SEXP rs_list(int n, const char** names, SEXP* elems) {
    SEXP x = Rf_allocVector(VECSXP, n);
    x->names = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; i++) {
        ((SEXP*)x->names->data)[i] = rs_mkchar(names[i]);
        ((SEXP*)x->data)[i] = elems[i];
    }
    return x;
}
int rs_type(SEXP x) { return x ? x->type : NILSXP; }
void rs_dims(SEXP x, long long* len, int* nrow, int* ncol) { *len = x->len; *nrow = x->nrow; *ncol = x->ncol; }
const char* rs_last_error(void) { return rs_err; }
void rs_run_finalizer(SEXP s) { if (s && s->type == EXTPTRSXP && s->finalizer) s->finalizer(s); }

typedef SEXP (*f0)(void);
#define A(i) args[i]
SEXP rs_call(void* fn, int nargs, SEXP* args) {
    SEXP out = NULL;
    rs_err[0] = 0;
    rs_jmp_armed = 1;
    if (setjmp(rs_jmp) == 0) {
        switch (nargs) {
            case 0: out = ((SEXP(*)(void))fn)(); break;
            case 1: out = ((SEXP(*)(SEXP))fn)(A(0)); break;
            case 2: out = ((SEXP(*)(SEXP, SEXP))fn)(A(0), A(1)); break;
            case 3: out = ((SEXP(*)(SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2)); break;
            case 4: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3)); break;
            case 5: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4)); break;
            case 6: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4), A(5)); break;
            case 7: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6)); break;
            case 8: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7)); break;
            case 9: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8)); break;
            case 10: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9)); break;
            case 14: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(
                               A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13)); break;
            default: snprintf(rs_err, sizeof rs_err, "rs_call: %d arguments not supported", nargs); out = NULL;
        }
    } else {
        out = NULL;                      /* Rf_error() was raised */
    }
    rs_jmp_armed = 0;
    return out;
}

#if defined(__GNUC__)
__attribute__((noinline))
#endif
void rs_poison_stack(void) {
    volatile unsigned char junk[65536];
    for (size_t i = 0; i < sizeof junk; i++) junk[i] = 0xAB;
    (void)junk[12345];
}
