/*
 * shim.c — ORACLE build infrastructure.  The handful of backend services the
 * reference leaf objects call: palloc/pfree onto malloc, ereport(ERROR) onto a
 * siglongjmp back into refwrap.c.  Everything else the objects reference but
 * never reach on these paths is auto-stubbed by the Makefile (stubs.c).
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

sigjmp_buf *ref_err_jmp = NULL;
char ref_err_msg[256];
static int cur_elevel;

/* data symbols the objects reference */
int DateStyle = 1;
void *session_timezone = NULL;
void *CurrentMemoryContext = NULL;
struct { int a, b, c, d; } GpIdentity;
unsigned int magic_hash_stash = 0;

void *palloc(size_t n) { return malloc(n ? n : 1); }
void *palloc0(size_t n) { return calloc(1, n ? n : 1); }
void pfree(void *p) { free(p); }
void *pg_detoast_datum(void *d) { return d; }
void *pg_detoast_datum_packed(void *d) { return d; }
int AggCheckCallContext(void *fcinfo, void **ctx) { (void) fcinfo; if (ctx) *ctx = NULL; return 1; /* AGG_CONTEXT_AGGREGATE */ }

static void
raise(void)
{
	if (cur_elevel >= 20)			/* ERROR */
	{
		if (ref_err_jmp)
			siglongjmp(*ref_err_jmp, 1);
		fprintf(stderr, "reference ereport(ERROR) outside REF_TRY: %s\n", ref_err_msg);
		abort();
	}
}

int errstart(int elevel, const char *f, int l, const char *fn, const char *dom)
{ (void) f; (void) l; (void) fn; (void) dom; cur_elevel = elevel; return elevel >= 20; }
void errfinish(int dummy, ...) { (void) dummy; raise(); }
int errcode(int c) { (void) c; return 0; }
int errmsg(const char *fmt, ...)
{ va_list ap; va_start(ap, fmt); vsnprintf(ref_err_msg, sizeof ref_err_msg, fmt, ap); va_end(ap); return 0; }
int errhint(const char *fmt, ...) { (void) fmt; return 0; }
int errdetail(const char *fmt, ...) { (void) fmt; return 0; }
int errdetail_internal(const char *fmt, ...) { (void) fmt; return 0; }
int errprintstack(int on) { (void) on; return 0; }
void elog_start(const char *f, int l, const char *fn) { (void) f; (void) l; (void) fn; }
void elog_finish(int elevel, const char *fmt, ...)
{ va_list ap; va_start(ap, fmt); vsnprintf(ref_err_msg, sizeof ref_err_msg, fmt, ap); va_end(ap); cur_elevel = elevel; raise(); }

void ref_unreachable(const char *name)
{ fprintf(stderr, "reference object called unstubbed backend function %s\n", name); abort(); }

/* timestamp.c is not among the leaf objects (it needs the int128 configure probe); date.o calls only this
 * comparison, which with integer datetimes is the plain ordering of two int64 (timestamp.c:2536-2539) */
int timestamp_cmp_internal(long dt1, long dt2) { return (dt1 < dt2) ? -1 : ((dt1 > dt2) ? 1 : 0); }

/* numeric_in asks whether its input spells "NaN" */
int pg_strncasecmp(const char *a, const char *b, size_t n)
{
	while (n-- > 0)
	{
		unsigned char x = (unsigned char) *a++, y = (unsigned char) *b++;
		if (x >= 'A' && x <= 'Z') x += 'a' - 'A';
		if (y >= 'A' && y <= 'Z') y += 'a' - 'A';
		if (x != y) return (int) x - (int) y;
		if (x == 0) break;
	}
	return 0;
}
