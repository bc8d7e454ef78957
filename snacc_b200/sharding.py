"""Sharding of the ordered-pair job matrix across GPUs.

Jobs are independent: there is no data-path collective.  The corpus is replicated (one broadcast), rank r owns the
contiguous COLUMN band ``[r*N/W, (r+1)*N/W)`` of S -- all x against its share of the y -- and the int64 blocks are
gathered at the end.  Columns, not rows, because a tile of the LZ4 kernel is "one y, up to 104 x": with all N rows on
every rank the tiles stay full however many ranks there are (row bands leave them 62 % full at 8 ranks), and the
per-sequence work every rank then repeats (C(x) and the prefix checkpoint of every x) runs one sequence per CTA in
parallel, so it costs the same wall time for N sequences as for N/W.  With ``torch.distributed`` uninitialised this
is the single-GPU path.  The same code runs under the ``gloo`` backend on CPU for the host-logic tests, with
``size_fn`` standing in for the GPU engine.
"""
import numpy as np

from . import fasta
from .engine import GETSIZEOF_BIAS


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def owned_cols(n, rank, world):
    """contiguous column band of rank `rank`: a rectangle of the job matrix, which the library runs without
    per-job arrays (snacc_tile_sizes)"""
    return np.arange(rank * n // world, (rank + 1) * n // world, dtype=np.int64)


owned_rows = owned_cols          # the bands of gather_rows (which gathers the TRANSPOSED column blocks) are the same


def broadcast_corpus(files, dist, device=None):
    """rank 0 parses the FASTA files once; everybody receives (data, seq_offsets, rec_offsets)."""
    import torch
    rank = dist.get_rank()
    if rank == 0:
        data, so, ro = fasta.load_corpus(files)
        meta = [int(data.size), so.tolist(), ro.tolist()]
    else:
        data, meta = None, None
    box = [meta]
    dist.broadcast_object_list(box, src=0)
    nbytes, so, ro = box[0]
    if device is not None:
        t = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if rank == 0:
            t.copy_(torch.from_numpy(data.copy()))
    else:
        t = torch.from_numpy(data.copy()) if rank == 0 else torch.empty(nbytes, dtype=torch.uint8)
    dist.broadcast(t, src=0)
    return t, np.asarray(so, dtype=np.uint64), np.asarray(ro, dtype=np.uint64)


def gather_rows(local_block, n, dist):
    """all ranks get the full n x width int64 matrix from per-rank blocks of consecutive rows (owned_rows)"""
    import torch
    world = dist.get_world_size()
    width = local_block.shape[1]
    rows_max = (n + world - 1) // world
    pad = np.full((rows_max, width), -1, dtype=np.int64)
    pad[:local_block.shape[0]] = local_block
    t = torch.from_numpy(pad)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.zeros((n, width), dtype=np.int64)
    for r in range(world):
        rows = owned_rows(n, r, world)
        full[rows] = outs[r].cpu().numpy()[:rows.size]
    return full


def ncd_host(C, S, fast_mode=False, bias=GETSIZEOF_BIAS):
    """float64 epilogue on the host (identical arithmetic to snacc_ncd on the device)."""
    C = np.asarray(C, dtype=np.int64) + bias
    S = np.asarray(S, dtype=np.int64) + bias
    lo = np.minimum(C[:, None], C[None, :]).astype(np.float64)
    hi = np.maximum(C[:, None], C[None, :]).astype(np.float64)
    d1 = (S.astype(np.float64) - lo) / hi
    if fast_mode:
        return d1
    d2 = (S.T.astype(np.float64) - lo) / hi
    return np.minimum(d1, d2)


def all_pairs(files, algorithm, reverse_complement, fast_mode, engine=None, rows_per_call=None, size_fn=None):
    """Returns (labels, C, S, D).  ``size_fn(data, so, ro, rc, cols) -> (C_all, S[:, cols])`` replaces the GPU
    engine in CPU tests."""
    dist = _dist()
    n = len(files)
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    cols = owned_cols(n, rank, world)
    own_engine = None
    if size_fn is None:
        from .engine import Engine
        if engine is None:
            import os
            own_engine = engine = Engine(int(os.environ.get("LOCAL_RANK", "0")) if dist else 0)
        if dist:
            import torch
            dev = torch.device("cuda", engine.device) if dist.get_backend() == "nccl" else None
            t, so, ro = broadcast_corpus(files, dist, dev)
            if dev is not None:
                torch.cuda.synchronize(dev)
                engine.upload_device(t.data_ptr(), so, ro, reverse_complement)
            else:
                engine.upload(t.numpy(), so, ro, reverse_complement)
        else:
            data, so, ro = fasta.load_corpus(files)
            engine.upload(data, so, ro, reverse_complement)
        C = engine.single_sizes(algorithm)                  # every rank: all x (also leaves their prefix checkpoints)
        S_cols = np.zeros((n, cols.size), dtype=np.int64)
        step = rows_per_call or max(1, n)
        for a in range(0, n, step):
            nr = min(step, n - a)
            if not cols.size:
                break
            if fast_mode:
                # upper triangle only: rows a..a+nr-1 against the owned columns at or right of the row
                xs, ys = [], []
                for r in range(a, a + nr):
                    cc = cols[cols >= r]
                    xs.append(np.full(cc.size, r, dtype=np.int32)); ys.append(cc.astype(np.int32))
                xs = np.concatenate(xs) if xs else np.zeros(0, np.int32)
                ys = np.concatenate(ys) if ys else np.zeros(0, np.int32)
                if xs.size:
                    vals = engine.pair_sizes(algorithm, xs, ys)
                    S_cols[xs, ys - int(cols[0])] = vals
            else:
                S_cols[a:a + nr] = engine.tile_sizes(algorithm, a, nr, int(cols[0]), int(cols.size))
    else:
        if dist:
            t, so, ro = broadcast_corpus(files, dist, None)
            data = t.numpy()
        else:
            data, so, ro = fasta.load_corpus(files)
        C, S_cols = size_fn(data, so, ro, reverse_complement, cols)
    C = np.asarray(C, dtype=np.int64)
    if dist:
        # one gather of the (transposed) column blocks; no collective on the data path before this point
        S = np.ascontiguousarray(gather_rows(np.ascontiguousarray(S_cols.T), n, dist).T)
    else:
        S = S_cols
    if fast_mode:
        iu = np.triu_indices(n, 1)
        S = S.copy()
        S[(iu[1], iu[0])] = S[iu]
    D = ncd_host(C, S, fast_mode)
    if own_engine is not None:
        own_engine.close()
    return [str(f) for f in files], C, S, D
