/*
 * gg_program.h — the per-plan "device program" the host compiles from gg_plan.h
 * structures and hands to the kernels as a __grid_constant__ parameter.
 *
 * The reference evaluates Expr trees with one fmgr call per node per row
 * (execQual.c:2169,6260; SURVEY §8a rows 5,7).  Here the tree is flattened once
 * per plan into an accumulator machine: one 64-bit accumulator + null flag and
 * four temporaries per lane, every lane of a warp running the same op on its
 * own tuple.  Each op is a single IEEE/integer operation applied in the same
 * order as the tree, so per-row values are bit-identical to the reference's.
 */
#ifndef GG_PROGRAM_H
#define GG_PROGRAM_H

#include <stdint.h>
#include "../../include/gg_plan.h"

#define GGP_MAX_COLS    16     /* distinct referenced columns per side */
#define GGP_MAX_CONSTS  24
#define GGP_MAX_CODE    224
#define GGP_MAX_ACCS    16     /* accumulator columns (deduplicated aggregate arguments) */
#define GGP_MAX_PAIRS   128    /* fast path: groups x accumulator columns held in registers */
#define GGP_FAST_GROUPS 32     /* fast path: groups per block table */

/* how a referenced column is loaded into the 64-bit accumulator */
enum ggp_loadtype {
	GGP_LD_I4 = 1,       /* int4/date: sign-extended */
	GGP_LD_I8 = 2,       /* int8/timestamp/float8 bits */
	GGP_LD_BPCHAR = 3,   /* short string, trailing blanks stripped (bcTruelen), <= 8 bytes packed LSB-first */
	GGP_LD_VARCHAR = 4,  /* short string, not stripped */
	GGP_LD_BOOL = 5
};

enum ggp_src {           /* operand kinds */
	GGP_SRC_NONE = 0,
	GGP_SRC_COL = 1,     /* column slot of the outer (scan) tuple */
	GGP_SRC_CONST = 2,
	GGP_SRC_TEMP = 3,
	GGP_SRC_ICOL = 4     /* column slot of the inner tuple (joins) */
};

enum ggp_opcode {
	GGP_LOAD = 1,        /* acc = src */
	GGP_STORE,           /* temp[idx] = acc */
	GGP_F8ADD, GGP_F8SUB, GGP_F8RSUB, GGP_F8MUL, GGP_F8DIV, GGP_F8RDIV,   /* float.c:782-850 incl. CHECKFLOATVAL */
	GGP_CMPF8,           /* acc = float8_cmp_internal(acc, src) cc  (float.c:964) ; cc in aux */
	GGP_CMPI,            /* acc = (int64)acc cc (int64)src */
	GGP_CMPSTR,          /* acc = packed strings equal / not equal (bpchareq on stripped bytes) */
	GGP_DATE2TS,         /* acc = date2timestamp(acc)  (date.c:457) */
	GGP_I2F8,            /* acc = (double)(int64)acc   (i4tod / i8tod) */
	GGP_AND, GGP_OR,     /* 3-valued, src = temp/col/const (execQual.c:3404,3455) */
	GGP_NOT, GGP_ISNULL, GGP_ISNOTNULL
};

enum ggp_cc { GGP_LT = 0, GGP_LE, GGP_EQ, GGP_NE, GGP_GT, GGP_GE };

typedef struct ggp_op {
	uint8_t op;
	uint8_t src;         /* ggp_src */
	uint8_t idx;         /* slot / const / temp index */
	uint8_t aux;         /* cc for compares */
} ggp_op;

/* accumulator column kinds */
enum ggp_acckind {
	GGP_ACC_F8SUM = 1,   /* sum and (if sq) sum of squares of a float8 expression + non-null count */
	GGP_ACC_F8MIN, GGP_ACC_F8MAX,
	GGP_ACC_I8SUM, GGP_ACC_I8MIN, GGP_ACC_I8MAX,
	GGP_ACC_COUNT        /* non-null count only (count(expr)); count(*) needs no column */
};

typedef struct ggp_attr {
	int16_t attlen;
	int8_t  attalign;    /* 'c','s','i','d' */
	int8_t  slot;        /* column slot if referenced, else -1 */
	int16_t cacheoff;    /* attcacheoff: constant offset while no NULL/varlena precedes, else -1 (heaptuple.c:1160) */
	int8_t  notnull;
	int8_t  pad;
} ggp_attr;

/* One side (scan tuple layout + the columns the program touches) */
typedef struct ggp_side {
	int32_t natts;           /* attributes in the descriptor */
	int32_t natts_walk;      /* walk attributes [0, natts_walk) : highest referenced attno */
	int32_t first_walk;      /* first attribute whose offset is not a constant (no-NULL tuples start walking here) */
	int32_t ncols;
	ggp_attr att[GG_MAX_ATTS];
	uint8_t  coltype[GGP_MAX_COLS];   /* ggp_loadtype per slot */
	uint8_t  colatt[GGP_MAX_COLS];    /* 0-based attribute per slot */
} ggp_side;

typedef struct ggp_span { int16_t start, len; } ggp_span;

typedef struct ggp_program {
	ggp_side outer;
	int32_t  nconst;
	int32_t  nullable;       /* 1: some referenced value can be NULL => null-tracking kernel variant */
	int64_t  consts[GGP_MAX_CONSTS];
	int32_t  constnull;      /* bit i: const i is NULL */
	int32_t  ncode;
	ggp_op   code[GGP_MAX_CODE];
	ggp_span qual;           /* len 0: no qual */
	int32_t  nkeys;
	ggp_span key[GG_MAX_KEYS];
	uint8_t  keytype[GG_MAX_KEYS];    /* 1 int, 2 float8 (normalise -0/NaN), 3 string */
	int32_t  nacc;
	ggp_span acc[GGP_MAX_ACCS];
	uint8_t  acckind[GGP_MAX_ACCS];
	uint8_t  accsq[GGP_MAX_ACCS];     /* float8 sum also needs sum of squares (avg's float8_accum state) */
} ggp_program;

/* One partial group record: what a block (or a segment, for the FINAL stage) knows about one group.
 * The merge kernel folds records with equal keys in a fixed order, so results are deterministic. */
typedef struct ggp_grec {
	uint64_t key[GG_MAX_KEYS];
	uint32_t keynull;        /* bit i: key i is NULL */
	uint32_t valid;
	uint64_t count;          /* rows of the group (count(*)) */
	double   sum[GGP_MAX_ACCS];      /* F8SUM: sum ; F8MIN/MAX: value ; I8*: int64 bits */
	double   sumsq[GGP_MAX_ACCS];
	uint64_t n[GGP_MAX_ACCS];        /* non-null inputs */
} ggp_grec;

/* error flags raised by kernels (bit mask in a device word) */
#define GGP_EF_FLOAT_OVERFLOW   0x01
#define GGP_EF_FLOAT_UNDERFLOW  0x02
#define GGP_EF_DIV_ZERO         0x04
#define GGP_EF_VISIBILITY       0x08
#define GGP_EF_BADPAGE          0x10
#define GGP_EF_GROUP_OVERFLOW   0x20   /* more groups than the fast path holds: rerun on the general path */
#define GGP_EF_STRING_TOO_LONG  0x40
#define GGP_EF_DATE_RANGE       0x80
#define GGP_EF_NOTNULL_VIOLATED 0x100
#define GGP_EF_INT_OVERFLOW     0x200
#define GGP_EF_TABLE_FULL       0x400
#define GGP_EF_SAW_INF          0x800   /* informational: an aggregate input was +-Inf/NaN */
#define GGP_EF_INFO_MASK        (GGP_EF_SAW_INF)

typedef struct ggp_acckinds { uint8_t k[GGP_MAX_ACCS]; } ggp_acckinds;

#ifdef __cplusplus
/* host-side compiler (gg_compile.cpp) */
struct ggp_aggmap {          /* how each Aggref reads the accumulator columns */
	int32_t col;             /* accumulator column, -1 for count(*) */
};
int ggp_compile_scanagg(const gg_scan *scan, const gg_agg *agg, const gg_exprpool *pool,
                        ggp_program *prog, ggp_aggmap *aggmap, char *err, int errlen);
#endif

#endif /* GG_PROGRAM_H */
