/* TEST INFRASTRUCTURE: C entry points over the product's plan compiler for tests/mock/ggb200_mock.c */
#include <vector>
#include "../../greengage_b200/csrc/gg_program.h"
#include "../../include/ggb200.h"

extern "C" int mock_compile_scanagg(const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool, char *err, int errlen)
{
	ggp_program prog;
	ggp_aggmap aggmap[GG_MAX_AGGS];
	return ggp_compile_scanagg(scan, agg, pool, &prog, aggmap, err, errlen);
}

extern "C" int mock_compile_join(const gg_scan *outer, const gg_scan *inner, const gg_hashjoin *hj, const gg_agg *agg,
                                 const gg_exprpool *pool, char *err, int errlen)
{
	std::vector<ggp_joinprog> jp(1);
	ggp_aggmap aggmap[GG_MAX_AGGS];
	return ggp_compile_join(outer, inner, hj, agg, pool, &jp[0], aggmap, err, errlen);
}
