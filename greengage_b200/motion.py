"""Motion for host-resident rows: Redistribute / Gather of the (few) partial-aggregate rows a slice emits.

Replaces execMotionSender / execMotionUnsortedReceiver over the UDP interconnect
(src/backend/executor/nodeMotion.c:270-374,378; src/backend/cdb/motion/ic_udpifc.c) with collectives on
one communicator (torch.distributed: NCCL over NVLink on the GPU box, gloo in CPU tests): a count
exchange followed by an all-to-all-v of fixed-width row records.  Routing is libgghost's gg_cdbhash_route:
cdbhash + jump consistent hash, bit-exact with src/backend/cdb/cdbhash.c:191-287.  EOS is implicit in
collective completion (SendEndOfStream, cdbmotion.c:532, has nothing left to do).
Bulk redistribution of scanned rows is done on the device (csrc/gg_motion.cu).
"""
import ctypes as C

import numpy as np

from . import capi

ROW_BYTES = C.sizeof(capi.gg_aggrow)


def route_aggrow(row, key_typids, nsegs):
    n = len(key_typids)
    t = (C.c_int32 * n)(*key_typids)
    v = (C.c_int64 * n)(*[row.key[i] for i in range(n)])
    ln = (C.c_int32 * n)(*[row.keylen[i] for i in range(n)])
    nu = (C.c_int32 * n)(*[row.keyisnull[i] for i in range(n)])
    return capi.host_lib().gg_cdbhash_route(t, v, ln, nu, n, nsegs)


def _rows_to_bytes(rows):
    buf = np.zeros(len(rows) * ROW_BYTES, dtype=np.uint8)
    for i, r in enumerate(rows):
        buf[i * ROW_BYTES:(i + 1) * ROW_BYTES] = np.frombuffer(bytes(r), dtype=np.uint8)
    return buf


def _bytes_to_rows(buf):
    n = buf.size // ROW_BYTES
    out = []
    raw = buf.tobytes()
    for i in range(n):
        out.append(capi.gg_aggrow.from_buffer_copy(raw[i * ROW_BYTES:(i + 1) * ROW_BYTES]))
    return out


def _tensor(arr, device):
    import torch
    t = torch.from_numpy(arr)
    return t.to(device) if device is not None else t


def redistribute_aggrows(rows, key_typids, device=None, group=None):
    """Redistribute Motion (MOTIONTYPE_HASH) of aggregate rows on their grouping keys.
    Returns the rows routed to this rank."""
    import torch
    import torch.distributed as dist
    nsegs = dist.get_world_size(group)
    dest = [route_aggrow(r, key_typids, nsegs) for r in rows]
    order = sorted(range(len(rows)), key=lambda i: dest[i])
    send_counts = np.zeros(nsegs, dtype=np.int64)
    for d in dest:
        send_counts[d] += 1
    sendbuf = _rows_to_bytes([rows[i] for i in order])
    sc = _tensor(send_counts.copy(), device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.cpu().numpy()
    st = _tensor(sendbuf if sendbuf.size else np.zeros(0, dtype=np.uint8), device)
    rt = torch.empty(int(recv_counts.sum()) * ROW_BYTES, dtype=torch.uint8, device=st.device)
    dist.all_to_all_single(rt, st, output_split_sizes=[int(c) * ROW_BYTES for c in recv_counts],
                           input_split_sizes=[int(c) * ROW_BYTES for c in send_counts], group=group)
    return _bytes_to_rows(rt.cpu().numpy())


SMALL_MOTION_ROWS = 64      # rows per segment one fixed-size record carries


def _allgather_rows(rows, device, group):
    """One collective: every rank contributes [row count | up to SMALL_MOTION_ROWS rows].  Returns (per-rank row
    lists, per-rank true counts); a rank with more rows than the record holds contributes its count only."""
    import torch
    import torch.distributed as dist
    nsegs = dist.get_world_size(group)
    buf = np.zeros(8 + SMALL_MOTION_ROWS * ROW_BYTES, dtype=np.uint8)
    buf[:8] = np.frombuffer(np.int64(len(rows)).tobytes(), dtype=np.uint8)
    if len(rows) <= SMALL_MOTION_ROWS:
        b = _rows_to_bytes(rows)
        buf[8:8 + b.size] = b
    t = _tensor(buf, device)
    out = torch.empty(nsegs * buf.size, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    host = out.cpu().numpy().reshape(nsegs, buf.size)
    counts = [int(np.frombuffer(host[r, :8].tobytes(), dtype=np.int64)[0]) for r in range(nsegs)]
    per = [_bytes_to_rows(host[r, 8:8 + counts[r] * ROW_BYTES]) if counts[r] <= SMALL_MOTION_ROWS else [] for r in range(nsegs)]
    return per, counts


def redistribute_small(rows, key_typids, device=None, group=None):
    """Redistribute Motion for the handful of rows a partial aggregate emits: ONE all-gather of fixed-size records,
    each segment keeps the rows cdbhash routes to it (sender order).  When any segment has more rows than a record
    holds — every rank sees that in the gathered counts — all ranks take the all-to-all-v path instead."""
    import torch.distributed as dist
    nsegs = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per, counts = _allgather_rows(rows, device, group)
    if max(counts) > SMALL_MOTION_ROWS:
        return redistribute_aggrows(rows, key_typids, device=device, group=group)
    return [r for sender in per for r in sender if route_aggrow(r, key_typids, nsegs) == rank]


def gather_small(rows, dst=0, device=None, group=None):
    """Gather Motion for a handful of rows per segment: one all-gather, the receiver keeps everything."""
    import torch.distributed as dist
    per, counts = _allgather_rows(rows, device, group)
    if max(counts) > SMALL_MOTION_ROWS:
        return gather_aggrows(rows, dst, device=device, group=group)
    if dist.get_rank(group) != dst:
        return []
    return [r for sender in per for r in sender]


# ---- the same two Motions on raw row buffers (numpy uint8 views of gg_aggrow arrays): no per-row Python objects ----

def route_rows_raw(buf, n, key_typids, nsegs):
    """dest[i] for the n gg_aggrow records in buf (one C call)"""
    dest = np.empty(n, dtype=np.int32)
    if n:
        t = (C.c_int32 * len(key_typids))(*key_typids)
        capi.host_lib().gg_cdbhash_route_aggrows(buf.ctypes.data, n, t, len(key_typids), nsegs, dest.ctypes.data)
    return dest


def _allgather_raw(buf, n, device, group):
    import torch
    import torch.distributed as dist
    nsegs = dist.get_world_size(group)
    rec = np.zeros(8 + SMALL_MOTION_ROWS * ROW_BYTES, dtype=np.uint8)
    rec[:8].view(np.int64)[0] = n
    if n <= SMALL_MOTION_ROWS:
        rec[8:8 + n * ROW_BYTES] = buf[:n * ROW_BYTES]
    t = _tensor(rec, device)
    out = torch.empty(nsegs * rec.size, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    host = out.cpu().numpy().reshape(nsegs, rec.size)
    counts = host[:, :8].copy().view(np.int64).reshape(nsegs)
    return host, counts


def redistribute_small_raw(buf, n, key_typids, device=None, group=None):
    """redistribute_small on a raw buffer; returns (buffer of the rows routed here, their number)"""
    import torch.distributed as dist
    nsegs, rank = dist.get_world_size(group), dist.get_rank(group)
    host, counts = _allgather_raw(buf, n, device, group)
    if counts.max() > SMALL_MOTION_ROWS:
        rows = redistribute_aggrows(_bytes_to_rows(buf[:n * ROW_BYTES]), key_typids, device=device, group=group)
        return _rows_to_bytes(rows), len(rows)
    allrows = np.concatenate([host[r, 8:8 + int(counts[r]) * ROW_BYTES] for r in range(nsegs)])
    tot = int(counts.sum())
    dest = route_rows_raw(allrows, tot, key_typids, nsegs)
    keep = np.nonzero(dest == rank)[0]
    mine = allrows.reshape(tot, ROW_BYTES)[keep].reshape(-1) if tot else allrows
    return np.ascontiguousarray(mine), len(keep)


def gather_small_raw(buf, n, dst=0, device=None, group=None):
    import torch.distributed as dist
    nsegs, rank = dist.get_world_size(group), dist.get_rank(group)
    host, counts = _allgather_raw(buf, n, device, group)
    if counts.max() > SMALL_MOTION_ROWS:
        rows = gather_aggrows(_bytes_to_rows(buf[:n * ROW_BYTES]), dst, device=device, group=group)
        return _rows_to_bytes(rows), len(rows)
    if rank != dst:
        return np.zeros(0, dtype=np.uint8), 0
    return np.concatenate([host[r, 8:8 + int(counts[r]) * ROW_BYTES] for r in range(nsegs)]), int(counts.sum())


def gather_aggrows(rows, dst=0, device=None, group=None):
    """Gather Motion (MOTIONTYPE_FIXED to one receiver): rows of all ranks on `dst`, in sender order."""
    import torch
    import torch.distributed as dist
    nsegs = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cnt = _tensor(np.array([len(rows)], dtype=np.int64), device)
    counts = [torch.empty_like(cnt) for _ in range(nsegs)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0
    pad = np.zeros(mx * ROW_BYTES, dtype=np.uint8)
    b = _rows_to_bytes(rows)
    pad[:b.size] = b
    t = _tensor(pad, device)
    bufs = [torch.empty_like(t) for _ in range(nsegs)]
    dist.all_gather(bufs, t, group=group)
    if rank != dst:
        return []
    out = []
    for r in range(nsegs):
        out.extend(_bytes_to_rows(bufs[r].cpu().numpy()[:counts[r] * ROW_BYTES]))
    return out
