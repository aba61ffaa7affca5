/* A recording mock of libbsfg_b200 for the CPU-only glue test (test infrastructure only): the glue is linked against
 * this instead of the CUDA library, every entry point returns BSFG_OK, and bsfg_set_state keeps a byte copy of the
 * bsfg_state it was handed so the test can check that the glue zeroes the members it does not fill
 * (px_factor / F_px were stack garbage). */
#include <string.h>
#include "bsfg_b200.h"

static bsfg_state last_state;
static int n_set_state = 0;
static int dummy_ctx;

const bsfg_state* mock_last_state(void) { return &last_state; }
int mock_set_state_calls(void) { return n_set_state; }
unsigned long mock_sizeof_state(void) { return (unsigned long)sizeof(bsfg_state); }

const char* bsfg_last_error(void) { return "mock"; }
int bsfg_create(const bsfg_config* cfg, bsfg_ctx** out) { (void)cfg; *out = (bsfg_ctx*)&dummy_ctx; return BSFG_OK; }
int bsfg_destroy(bsfg_ctx* c) { (void)c; return BSFG_OK; }
int bsfg_set_priors(bsfg_ctx* c, const bsfg_priors* p) { (void)c; (void)p; return BSFG_OK; }
int bsfg_upload_constants(bsfg_ctx* c, const bsfg_constants* k) { (void)c; (void)k; return BSFG_OK; }
int bsfg_set_state(bsfg_ctx* c, const bsfg_state* st) { (void)c; memcpy(&last_state, st, sizeof last_state); n_set_state++; return BSFG_OK; }
int bsfg_get_state(bsfg_ctx* c, bsfg_state* st) { (void)c; st->k = 3; st->nrun = 7; return BSFG_OK; }
int bsfg_sweep(bsfg_ctx* c, int n) { (void)c; return n >= 0 ? BSFG_OK : BSFG_EARG; }
