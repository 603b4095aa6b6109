/* MINIMAL STAND-IN for R's <R.h>, for compile checks only (tests/test_r_glue.py).
 * R is not installed in the build image; these are the declarations of R's public C API that
 * r/src/glmma_R.c uses, with the signatures documented in "Writing R Extensions" (section 5 and 6).
 * Never shipped and never linked: with a real R installation the real headers are used instead. */
#ifndef GLMMA_R_STUB_R_H
#define GLMMA_R_STUB_R_H
#include <stddef.h>
#include <string.h>
void Rf_error(const char*, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char*, ...) __attribute__((format(printf, 1, 2)));
#endif
