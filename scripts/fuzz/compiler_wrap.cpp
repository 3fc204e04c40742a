/* Test/fuzz harness: the plan compiler's entry points without the CUDA library (scripts/fuzz/run_compiler_fuzz.sh). */
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <vector>
#include "../../greengage_b200/csrc/gg_program.h"
#include "../../include/ggb200.h"
static char lasterr[512];
void gg_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(lasterr, sizeof lasterr, fmt, ap); va_end(ap); }
extern "C" const char *fz_last_error() { return lasterr; }
extern "C" int fz_scanagg(const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool, char *buf, int cap)
{
	ggp_program prog; ggp_aggmap aggmap[GG_MAX_AGGS]; char msg[256];
	int rc = ggp_compile_scanagg(scan, agg, pool, &prog, aggmap, msg, sizeof msg);
	if (rc != GG_OK) { gg_set_error("%s", msg); return rc; }
	return ggp_disasm(&prog, buf, cap);
}
extern "C" int fz_join(const gg_scan *outer, const gg_scan *inner, const gg_hashjoin *hj, const gg_agg *agg, const gg_exprpool *pool, char *buf, int cap)
{
	std::vector<ggp_joinprog> jpbuf(1); ggp_joinprog &jp = jpbuf[0]; ggp_aggmap aggmap[GG_MAX_AGGS]; char msg[256];
	int rc = ggp_compile_join(outer, inner, hj, agg, pool, &jp, aggmap, msg, sizeof msg);
	if (rc != GG_OK) { gg_set_error("%s", msg); return rc; }
	int n = ggp_disasm(&jp.build, buf, cap);
	n += ggp_disasm(&jp.probe, buf + n, cap - n);
	return n;
}
extern "C" int fz_motion(const gg_scan *scan, const gg_exprpool *pool, const int32_t *hk, int nkeys, const int32_t *pl, int npl, char *buf, int cap)
{
	ggp_program prog; char msg[256]; uint8_t ht[GG_MAX_KEYS];
	int rc = ggp_compile_motion(scan, pool, hk, nkeys, pl, npl, &prog, ht, msg, sizeof msg);
	if (rc != GG_OK) { gg_set_error("%s", msg); return rc; }
	return ggp_disasm(&prog, buf, cap);
}
/* the whole compiled program and aggregate map as bytes (comparing two builds of the compiler) */
extern "C" int fz_scanagg_raw(const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool, unsigned char *out, int cap)
{
	ggp_program prog; ggp_aggmap aggmap[GG_MAX_AGGS]; char msg[256];
	memset(aggmap, 0, sizeof aggmap);
	int rc = ggp_compile_scanagg(scan, agg, pool, &prog, aggmap, msg, sizeof msg);
	if (rc != GG_OK) return rc;
	if ((int) (sizeof prog + sizeof aggmap) > cap) return -8;
	memcpy(out, &prog, sizeof prog); memcpy(out + sizeof prog, aggmap, sizeof aggmap);
	return (int) (sizeof prog + sizeof aggmap);
}
extern "C" int fz_join_raw(const gg_scan *outer, const gg_scan *inner, const gg_hashjoin *hj, const gg_agg *agg, const gg_exprpool *pool, unsigned char *out, int cap)
{
	std::vector<ggp_joinprog> jpbuf(1); ggp_aggmap aggmap[GG_MAX_AGGS]; char msg[256];
	memset(aggmap, 0, sizeof aggmap);
	int rc = ggp_compile_join(outer, inner, hj, agg, pool, &jpbuf[0], aggmap, msg, sizeof msg);
	if (rc != GG_OK) return rc;
	if ((int) (sizeof(ggp_joinprog) + sizeof aggmap) > cap) return -8;
	memcpy(out, &jpbuf[0], sizeof(ggp_joinprog)); memcpy(out + sizeof(ggp_joinprog), aggmap, sizeof aggmap);
	return (int) (sizeof(ggp_joinprog) + sizeof aggmap);
}
