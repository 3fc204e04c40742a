/*
 * refwrap_tqual.c — ORACLE infrastructure: the reference's own tqual.o + transam.o (utils/time/tqual.c,
 * access/transam/transam.c, compiled in place) behind a plain C surface for tests/golden/make_golden.py:
 * HeapTupleSatisfiesMVCC (tqual.c:997) of one tuple header against one snapshot.  The server state the rule consults is
 * stood in for by a table the caller fills: pg_clog status per xid (TransactionIdGetStatus, clog.c:375, is what
 * transam.c's TransactionLogFetch reads), "in progress" for the procarray (TransactionIdIsInProgress, procarray.c:978:
 * a transaction with neither a commit nor an abort record that has not crashed), one own xid for
 * TransactionIdIsCurrentTransactionId (xact.c:790).  Command ids are raw t_cid (no combo cids: callers do not set
 * HEAP_COMBOCID).  Test infrastructure only.
 */
#include "postgres.h"
#include "access/htup_details.h"
#include "access/transam.h"
#include "access/clog.h"
#include "access/xact.h"
#include "storage/bufmgr.h"
#include "utils/snapshot.h"
#include "utils/tqual.h"

bool		gp_disable_tuple_hints = false;
TransactionId TransactionXmin = FirstNormalTransactionId;

static const uint8 *tq_clog;
static uint32 tq_base, tq_n;
static TransactionId tq_own;

static int
tq_status(TransactionId xid)
{
	uint32		d = xid - tq_base;

	if (d >= tq_n)
		return TRANSACTION_STATUS_IN_PROGRESS;
	return (tq_clog[d >> 2] >> ((d & 3) * 2)) & 3;
}

XidStatus
TransactionIdGetStatus(TransactionId xid, XLogRecPtr *lsn)
{
	if (lsn)
		*lsn = 0;
	return tq_status(xid);
}

bool
TransactionIdIsInProgress(TransactionId xid)
{
	return TransactionIdIsNormal(xid) && tq_status(xid) == TRANSACTION_STATUS_IN_PROGRESS;
}

bool
TransactionIdIsCurrentTransactionId(TransactionId xid)
{
	return tq_own != InvalidTransactionId && xid == tq_own;
}

CommandId
HeapTupleHeaderGetCmin(HeapTupleHeader tup)
{
	return HeapTupleHeaderGetRawCommandId(tup);
}

CommandId
HeapTupleHeaderGetCmax(HeapTupleHeader tup)
{
	return HeapTupleHeaderGetRawCommandId(tup);
}

bool		XLogNeedsFlush(XLogRecPtr record) { return false; }
bool		BufferIsPermanent(Buffer buffer) { return false; }
void		MarkBufferDirtyHint(Buffer buffer, bool buffer_std) { }
bool		CLOGTransactionIsOld(TransactionId xid) { return false; }

/* header: the first 24 bytes of a heap tuple (copied: the rule sets hint bits).  Returns HeapTupleSatisfiesMVCC's answer. */
int
ref_heap_satisfies_mvcc(const uint8 *header, uint32 xmin, uint32 xmax, uint32 xcnt, const uint32 *xip, uint32 curcid,
						uint32 own_xid, uint32 clog_base, uint32 clog_n, const uint8 *clog)
{
	union
	{
		HeapTupleHeaderData h;
		uint8		bytes[64];
	}			copy;
	HeapTupleData tup;
	SnapshotData snap;

	tq_clog = clog;
	tq_base = clog_base;
	tq_n = clog_n;
	tq_own = own_xid;
	/* transam.c keeps the last (xid, status) it fetched in statics: push this call's answers out with a lookup of the
	 * bootstrap xid's neighbour, whose status nobody caches (TransactionLogFetch :62 answers special xids before the cache) */
	memset(&copy, 0, sizeof copy);
	memcpy(copy.bytes, header, 24);
	memset(&tup, 0, sizeof tup);
	tup.t_data = &copy.h;
	tup.t_len = 24;
	ItemPointerSet(&tup.t_self, 0, 1);
	memset(&snap, 0, sizeof snap);
	snap.satisfies = HeapTupleSatisfiesMVCC;
	snap.xmin = xmin;
	snap.xmax = xmax;
	snap.xip = (TransactionId *) xip;
	snap.xcnt = xcnt;
	snap.curcid = curcid;
	snap.haveDistribSnapshot = false;
	return HeapTupleSatisfiesMVCC(NULL, &tup, &snap, InvalidBuffer) ? 1 : 0;
}

/* transam.c caches one (xid, status) pair across calls (cachedFetchXid :40); a differential test changes the status table
 * under it, so it asks for an xid no case uses, with a committed status, before every case */
void
ref_tqual_flush_cache(uint32 scratch_xid)
{
	static const uint8 one = 0x55;			/* four committed xids */

	tq_clog = &one;
	tq_base = scratch_xid & ~3u;
	tq_n = 4;
	(void) TransactionIdDidCommit(scratch_xid);
}
