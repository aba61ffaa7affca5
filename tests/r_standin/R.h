/* Stand-in for R's <R.h> (test infrastructure only; see Rinternals.h in this directory). */
#ifndef RSTANDIN_R_H
#define RSTANDIN_R_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#endif
