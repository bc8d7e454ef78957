"""Row sharding of the ordered-pair job matrix across GPUs.

Jobs are independent: there is no data-path collective.  The corpus is replicated (one broadcast), rank
r owns the contiguous row band ``[r*N/W, (r+1)*N/W)`` of S (so every x prefix state is built by exactly one
rank), and the int64 row blocks are gathered at the end.  With ``torch.distributed`` uninitialised this is the
single-GPU path.  The same code runs under the ``gloo`` backend on CPU for the host-logic tests, with
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


def owned_rows(n, rank, world):
    """contiguous row band of rank `rank`: a rectangle of the job matrix, which the library runs without
    per-job arrays (snacc_tile_sizes)"""
    return np.arange(rank * n // world, (rank + 1) * n // world, dtype=np.int64)


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
    """all ranks get the full n x width int64 matrix from per-rank row blocks (rows r, r+W, ...)"""
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
    """Returns (labels, C, S, D).  ``size_fn(data, so, ro, rc, rows) -> (C_rows, S_rows)`` replaces the GPU
    engine in CPU tests."""
    dist = _dist()
    n = len(files)
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    rows = owned_rows(n, rank, world)
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
        C_rows = engine.single_sizes(algorithm, rows.astype(np.int32)) if rows.size else np.zeros(0, np.int64)
        S_rows = np.zeros((rows.size, n), dtype=np.int64)
        step = rows_per_call or max(1, rows.size)
        for a in range(0, rows.size, step):
            rr = rows[a:a + step]
            if fast_mode:
                xs = np.concatenate([np.full(n - r, r, dtype=np.int32) for r in rr]) if rr.size else np.zeros(0, np.int32)
                ys = np.concatenate([np.arange(r, n, dtype=np.int32) for r in rr]) if rr.size else np.zeros(0, np.int32)
                vals = engine.pair_sizes(algorithm, xs, ys)
                k = 0
                for i, r in enumerate(rr):
                    S_rows[a + i, r:] = vals[k:k + n - r]
                    k += n - r
            elif rr.size:
                S_rows[a:a + rr.size] = engine.tile_sizes(algorithm, int(rr[0]), int(rr.size), 0, n)
    else:
        if dist:
            t, so, ro = broadcast_corpus(files, dist, None)
            data = t.numpy()
        else:
            data, so, ro = fasta.load_corpus(files)
        C_rows, S_rows = size_fn(data, so, ro, reverse_complement, rows)
    if dist:
        # one gather of [S rows | C] per rank; no collective on the data path before this point
        full = gather_rows(np.concatenate([S_rows, np.asarray(C_rows, dtype=np.int64)[:, None]], axis=1), n, dist)
        S, C = np.ascontiguousarray(full[:, :n]), np.ascontiguousarray(full[:, n])
    else:
        S, C = S_rows, C_rows
    if fast_mode:
        iu = np.triu_indices(n, 1)
        S = S.copy()
        S[(iu[1], iu[0])] = S[iu]
    D = ncd_host(C, S, fast_mode)
    if own_engine is not None:
        own_engine.close()
    return [str(f) for f in files], C, S, D
