"""Sharding of the ordered-pair job matrix across GPUs -- the product's multi-GPU path.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).  Jobs are independent, so the only
exchanges are at the two ends:

* corpus: rank r reads and parses its contiguous band of the FASTA files, uploads it to its own GPU (1/W of the
  bytes over PCIe per rank) and the bands are all-gathered over NVLink (``exchange_corpus``) -- every rank then
  holds the whole corpus;
* results: the int64 blocks are all-gathered at the end (``gather_cols`` / ``_gather_values``).

Reference semantics (both orders, cli.py:120-129): rank r owns a contiguous COLUMN band of S -- all x against its
share of the y, bands cut by cumulative length so every rank gets the same number of bytes.  Columns, not rows,
because a tile of the LZ4 kernel is "one y, up to 104 x": with all N rows on every rank the tiles stay full however
many ranks there are.  ``--fast-mode`` (upper triangle only): column j costs j+1 jobs, so whole columns are dealt
to the ranks longest-processing-time first by their byte count and each rank runs its job list in one library
call.  With ``torch.distributed`` uninitialised this is the single-GPU path.
"""
import os
import time

import weakref

import numpy as np

from . import fasta
from .engine import GETSIZEOF_BIAS

_OWN_GROUP = False


class _nvtx:
    """NVTX range around the collectives (no-op without CUDA)"""
    def __init__(self, name):
        self.name, self.on = name, False

    def __enter__(self):
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.nvtx.range_push(self.name)
                self.on = True
        except Exception:
            pass

    def __exit__(self, *exc):
        if self.on:
            import torch
            torch.cuda.nvtx.range_pop()


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def init_distributed():
    """Join the process group described by the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*),
    binding this process to GPU LOCAL_RANK.  No-op when there is no such environment or a group already exists.
    Returns the local device index (0 outside torchrun)."""
    global _OWN_GROUP
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world <= 1 or "RANK" not in os.environ:
        return 0
    import torch
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if torch.cuda.is_available():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
        _OWN_GROUP = True
    return local


def shutdown_distributed():
    global _OWN_GROUP
    dist = _dist()
    if dist is not None and _OWN_GROUP:
        dist.destroy_process_group()
        _OWN_GROUP = False


def local_device():
    return int(os.environ.get("LOCAL_RANK", "0")) if _dist() else 0


def is_rank0():
    dist = _dist()
    return dist is None or dist.get_rank() == 0


# ---- partitions ---------------------------------------------------------------------------------------------------
def band_bounds(weights, world):
    """W+1 boundaries of contiguous bands with near-equal total weight (every band non-empty while n >= W)."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    if n == 0:
        return np.zeros(world + 1, dtype=np.int64)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    target = cum[-1] * np.arange(world + 1) / world
    b = np.searchsorted(cum, target, side="left").astype(np.int64)
    b = np.clip(b, 1, n)
    b = np.where(np.abs(cum[b - 1] - target) <= np.abs(cum[b] - target), b - 1, b)     # the nearer cut
    b[0], b[-1] = 0, n
    for r in range(1, world):                      # keep the bands non-empty and ordered
        b[r] = min(max(b[r], b[r - 1] + (1 if n >= world else 0)), n - (world - r) * (1 if n >= world else 0))
    return b


def owned_cols(n, rank, world, lengths=None):
    """contiguous column band of rank `rank`: a rectangle of the job matrix, which the library runs without
    per-job arrays (snacc_tile_sizes).  Bands hold equal bytes when the lengths are given, else equal counts."""
    b = band_bounds(np.ones(n) if lengths is None else lengths, world)
    return np.arange(b[rank], b[rank + 1], dtype=np.int64)


def fast_mode_cols(lengths, world):
    """--fast-mode: column j carries the jobs (i, j), i <= j.  Whole columns are dealt to ranks greedily, heaviest
    first (cost = bytes the compressor reads: sum over i <= j of len_i + len_j).  Returns a list of sorted column
    arrays, one per rank; deterministic, so every rank derives everybody's share."""
    L = np.asarray(lengths, dtype=np.float64)
    n = L.size
    cost = np.cumsum(L) + L * np.arange(1, n + 1)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world)
    owner = np.zeros(n, dtype=np.int64)
    for j in order:
        r = int(np.argmin(load))
        owner[j] = r
        load[r] += cost[j]
    return [np.nonzero(owner == r)[0].astype(np.int64) for r in range(world)]


def triangle_jobs(cols):
    """job list (xs, ys) of the upper-triangle columns `cols`: all (i, j) with i <= j, built without a Python loop"""
    cols = np.asarray(cols, dtype=np.int64)
    cnt = cols + 1
    ys = np.repeat(cols, cnt)
    start = np.cumsum(cnt) - cnt
    xs = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(start, cnt)
    return xs.astype(np.int32), ys.astype(np.int32)


# ---- corpus exchange ----------------------------------------------------------------------------------------------
def exchange_corpus(local_data, local_seq_lens, local_rec_lens, dist, device=None):
    """All-gather of per-rank corpus bands.  ``local_data``: uint8 array/tensor of this rank's sequences
    (concatenated), with their lengths and FASTA record lengths.  Returns (tensor of the whole corpus -- on
    ``device`` when given --, seq_offsets, rec_offsets).  The bytes move GPU to GPU (NCCL over NVLink) when a device
    is given: each rank's H2D copy is its own band only."""
    import torch
    world = dist.get_world_size()
    meta = [None] * world
    dist.all_gather_object(meta, ([int(v) for v in local_seq_lens], [int(v) for v in local_rec_lens]))
    nbytes = [int(sum(m[0])) for m in meta]
    cap = max(max(nbytes), 1)
    t = local_data if isinstance(local_data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_data, dtype=np.uint8))
    buf = torch.empty(cap, dtype=torch.uint8, device=device if device is not None else "cpu")
    buf[:t.numel()].copy_(t, non_blocking=True)           # the H2D copy of this rank's band
    allb = torch.empty(world * cap, dtype=torch.uint8, device=buf.device)
    with _nvtx("snacc_b200: corpus all-gather"):
        dist.all_gather_into_tensor(allb, buf)
    if all(b == cap for b in nbytes):
        full = allb
    else:
        full = torch.cat([allb[r * cap:r * cap + nbytes[r]] for r in range(world)])
    seq_lens = [l for m in meta for l in m[0]]
    rec_lens = [l for m in meta for l in m[1]]
    so = np.zeros(len(seq_lens) + 1, dtype=np.uint64)
    so[1:] = np.cumsum(np.asarray(seq_lens, dtype=np.uint64))
    ro = np.zeros(len(rec_lens) + 1, dtype=np.uint64)
    ro[1:] = np.cumsum(np.asarray(rec_lens, dtype=np.uint64))
    return full, so, ro


def load_corpus_sharded(files, dist, device=None):
    """rank r parses files [r*n/W, (r+1)*n/W) (the reference parses every file 2N+1 times, pairwise_ncd.py:29-36),
    then the bands are exchanged.  Returns what fasta.load_corpus returns, the bytes as a tensor."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(files)
    mine = files[rank * n // world:(rank + 1) * n // world]
    err = None
    data, seq_lens, rec_lens = np.zeros(0, np.uint8), [], []
    try:
        if mine:
            data, so, ro = fasta.load_corpus(mine)
            seq_lens, rec_lens = np.diff(so.astype(np.int64)).tolist(), np.diff(ro.astype(np.int64)).tolist()
    except Exception as e:                                 # every rank must reach the collectives: raise afterwards
        err = e
    errs = [None] * world
    dist.all_gather_object(errs, None if err is None else f"{type(err).__name__}: {err}")
    if err is not None:
        raise err
    bad = [e for e in errs if e]
    if bad:
        raise ValueError(bad[0]) if bad[0].startswith("ValueError") else RuntimeError(bad[0])
    return exchange_corpus(data, seq_lens, rec_lens, dist, device)


# ---- result gathers -----------------------------------------------------------------------------------------------
def _all_gather_padded(local, dist, device=None):
    """list (one per rank) of the 1-d int64 arrays every rank contributed"""
    import torch
    world = dist.get_world_size()
    local = np.ascontiguousarray(local, dtype=np.int64).ravel()
    sizes = [None] * world
    dist.all_gather_object(sizes, int(local.size))
    cap = max(max(sizes), 1)
    pad = np.zeros(cap, dtype=np.int64)
    pad[:local.size] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * cap, dtype=torch.int64, device=t.device)
    with _nvtx("snacc_b200: result / record all-gather"):
        dist.all_gather_into_tensor(out, t)
    out = out.cpu().numpy()
    return [out[r * cap:r * cap + sizes[r]] for r in range(world)]


def gather_cols(S_cols, bounds, n, dist, device=None, out=None):
    """all ranks get the full n x n matrix from per-rank column bands [bounds[r], bounds[r+1]).  With a CUDA device (NCCL)
    the bands are gathered and laid side by side on the device and the matrix comes back in one copy into page-locked
    memory (the c3 matrix is 800 MB: transposing it on the host cost more than computing it on 8 GPUs)."""
    if device is not None and getattr(device, "type", None) == "cuda":
        import torch
        world = dist.get_world_size()
        widths = [int(bounds[r + 1]) - int(bounds[r]) for r in range(world)]
        wmax = max(max(widths), 1)
        local = _tensor_of(np.ascontiguousarray(S_cols, dtype=np.int64).reshape(n, -1))
        # the device buffers are kept between calls (one shape at a time): a job that is repeated -- bench.py -- must not
        # depend on what the caching allocator can reuse
        key = (str(device), world, n, wmax)
        if _GATHER_BUF.get("key") != key:
            _GATHER_BUF.clear()
            _GATHER_BUF.update(key=key, pad=torch.zeros((n, wmax), dtype=torch.int64, device=device),
                               out=torch.empty((world, n, wmax), dtype=torch.int64, device=device),
                               S=torch.empty((n, n), dtype=torch.int64, device=device))
        pad, gathered, S_dev = _GATHER_BUF["pad"], _GATHER_BUF["out"], _GATHER_BUF["S"]
        if local.shape[1]:
            pad[:, :local.shape[1]].copy_(local, non_blocking=True)
        with _nvtx("snacc_b200: result all-gather"):
            dist.all_gather_into_tensor(gathered, pad)
        torch.cat([gathered[r, :, :widths[r]] for r in range(world) if widths[r]], dim=1, out=S_dev)
        S = out if out is not None and out.shape == (n, n) and out.dtype == np.int64 else _result_buffer(n, n)
        _tensor_of(S).copy_(S_dev)
        return S
    parts = _all_gather_padded(np.ascontiguousarray(S_cols.T), dist, device)     # transposed: a band is contiguous
    S = np.zeros((n, n), dtype=np.int64)
    for r, p in enumerate(parts):
        a, b = int(bounds[r]), int(bounds[r + 1])
        if b > a:
            S[:, a:b] = p.reshape(b - a, n).T
    return S


_GATHER_BUF = {}                               # gather_cols: device buffers of the last shape
_PINNED = weakref.WeakValueDictionary()        # data pointer of a page-locked numpy array -> the torch tensor that owns it


def pinned_array(shape, dtype=np.int64):
    """page-locked host array (numpy view of a pinned torch tensor); torch copies to / from it go by DMA because the
    tensor is remembered (``torch.from_numpy`` alone would treat the memory as pageable and stage the copy)"""
    import torch
    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    a = t.numpy()
    _PINNED[a.ctypes.data] = t
    return a


def _tensor_of(a):
    import torch
    t = _PINNED.get(a.ctypes.data)
    if t is not None and tuple(t.shape) == tuple(a.shape) and t.numpy().dtype == a.dtype:
        return t
    return torch.from_numpy(a)


def _result_buffer(rows, cols):
    """int64 (rows, cols) host array the library copies sizes into: page-locked when torch + CUDA are there (the c3 matrix
    is 800 MB; a pageable destination halves the copy rate)"""
    try:
        import torch
        if torch.cuda.is_available() and rows * cols >= (1 << 20):
            return pinned_array((rows, cols), np.int64)
    except Exception:
        pass
    return np.zeros((rows, cols), dtype=np.int64)


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


# ---- the sharded job -----------------------------------------------------------------------------------------------
MAX_JOBS_PER_CALL = 1 << 24


def sizes_matrix(engine, algorithm, fast_mode=False, rows_per_call=None, stats=None, band_out=None, full_out=None):
    """C (all singles) and S (all ordered pairs, or the mirrored upper triangle in fast mode) for the corpus already
    uploaded to ``engine``, sharded over the process group when there is one.  Every rank returns the full result.
    ``stats`` (a dict) receives the library's kernel times and launch counts of this rank; ``band_out``: an int64
    (n, width of this rank's band) array to receive the rank's column band (a caller that repeats the job -- bench.py --
    passes the same page-locked buffer every time instead of having a new one pinned per call); ``full_out``: the same
    for the gathered (n, n) matrix under a process group."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    n = engine.n_seqs
    lengths = np.asarray(engine.lengths, dtype=np.int64)
    dev = None
    if dist and dist.get_backend() == "nccl":
        import torch
        dev = torch.device("cuda", engine.device)

    def note(main=True):
        if stats is not None:
            stats["kernel_ms"] = stats.get("kernel_ms", 0.0) + engine.stat("total_kernel_ms")
            stats["launches"] = stats.get("launches", 0) + int(engine.stat("launches"))
            if main:
                stats["main_ms"] = stats.get("main_ms", 0.0) + engine.stat("main_kernel_ms")
                stats["packed_jobs"] = stats.get("packed_jobs", 0) + int(engine.stat("packed_jobs"))
                stats["bytewise_jobs"] = stats.get("bytewise_jobs", 0) + int(engine.stat("bytewise_jobs"))

    bounds = band_bounds(lengths, world)
    shares = fast_mode_cols(lengths, world) if fast_mode else [np.arange(bounds[r], bounds[r + 1]) for r in range(world)]
    if dist and engine.prefix_record_bytes(algorithm):
        # deflate codecs: the per-sequence preparation (index, match table, parse of the sequence alone) is what does
        # not shrink with the number of ranks, so each rank prepares only the sequences it owns as y -- its columns --
        # and the x-side products (checkpoint + size of x alone, a 680-byte record) are all-gathered
        mine = np.asarray(shares[rank], dtype=np.int32)
        if mine.size:
            engine.single_sizes(algorithm, mine)
            note(main=False)
        recs = engine.export_prefix(algorithm, mine)
        parts = _all_gather_padded(recs.view(np.int64).ravel() if recs.size else np.zeros(0, np.int64), dist, dev)
        for r, part in enumerate(parts):
            if r != rank and part.size:
                engine.import_prefix(algorithm, np.asarray(shares[r], dtype=np.int32), part.view(np.uint8))
    # every rank: all x (LZ4: the same pass leaves the prefix checkpoint every pair stream x.y resumes from; deflate
    # under a process group: served from the records just exchanged)
    C = engine.single_sizes(algorithm)
    note(main=False)
    if not fast_mode:
        a, b = int(bounds[rank]), int(bounds[rank + 1])
        S_cols = band_out if band_out is not None and band_out.shape == (n, b - a) else _result_buffer(n, b - a)
        step = rows_per_call or max(1, n)
        if b > a:
            for r0 in range(0, n, step):
                nr = min(step, n - r0)
                engine.tile_sizes(algorithm, r0, nr, a, b - a, out=S_cols[r0:r0 + nr])     # D2H straight into the band
                note()
        if stats is not None:
            stats["jobs"] = n * (b - a)
            stats["bytes"] = float(n * lengths[a:b].sum() + (b - a) * lengths.sum())
        t_g = time.perf_counter()
        S = gather_cols(S_cols, bounds, n, dist, dev, out=full_out) if dist else S_cols
        if stats is not None:
            stats["gather_s"] = time.perf_counter() - t_g
    else:
        xs, ys = triangle_jobs(shares[rank])
        vals = np.zeros(xs.size, dtype=np.int64)
        for k in range(0, xs.size, MAX_JOBS_PER_CALL):
            vals[k:k + MAX_JOBS_PER_CALL] = engine.pair_sizes(algorithm, xs[k:k + MAX_JOBS_PER_CALL],
                                                              ys[k:k + MAX_JOBS_PER_CALL])
            note()
        if stats is not None:
            stats["jobs"] = int(xs.size)
            stats["bytes"] = float(lengths[xs].sum() + lengths[ys].sum())
        parts = _all_gather_padded(vals, dist, dev) if dist else [vals]
        S = np.zeros((n, n), dtype=np.int64)
        for r, p in enumerate(parts):
            rx, ry = triangle_jobs(shares[r])
            S[rx, ry] = p
            S[ry, rx] = p                                  # mirrored: fast mode never computes the other order
    return np.asarray(C, dtype=np.int64), S


def all_pairs(files, algorithm, reverse_complement, fast_mode, engine=None, rows_per_call=None, _size_fn=None):
    """Returns (labels, C, S, D).  ``_size_fn(data, so, ro, rc, cols) -> (C_all, S[:, cols])`` is a test hook (the
    gloo tests exercise the partition / exchange / gather logic without a GPU); the product never passes it."""
    dist = _dist()
    n = len(files)
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    own_engine = None
    if _size_fn is None:
        from .engine import Engine
        if engine is None:
            own_engine = engine = Engine(local_device())
        try:
            if dist:
                import torch
                dev = torch.device("cuda", engine.device) if dist.get_backend() == "nccl" else None
                t, so, ro = load_corpus_sharded(files, dist, dev)
                if dev is not None:
                    torch.cuda.synchronize(dev)
                    engine.upload_device(t.data_ptr(), so, ro, reverse_complement)
                else:
                    engine.upload(t.numpy(), so, ro, reverse_complement)
                del t
            else:
                data, so, ro = fasta.load_corpus(files)
                engine.upload(data, so, ro, reverse_complement)
            C, S = sizes_matrix(engine, algorithm, fast_mode, rows_per_call)
            D = engine.ncd(C, S, formula=1 if fast_mode else 0)        # K4: float64 epilogue on the device
        finally:
            if own_engine is not None:
                own_engine.close()
    else:
        if dist:
            t, so, ro = load_corpus_sharded(files, dist, None)
            data = t.numpy()
        else:
            data, so, ro = fasta.load_corpus(files)
        lengths = np.diff(so.astype(np.int64))
        if not fast_mode:
            bounds = band_bounds(lengths, world)
            cols = np.arange(bounds[rank], bounds[rank + 1], dtype=np.int64)
            C, S_cols = _size_fn(data, so, ro, reverse_complement, cols)
            S = gather_cols(np.asarray(S_cols, dtype=np.int64).reshape(n, cols.size), bounds, n, dist) if dist else S_cols
        else:
            shares = fast_mode_cols(lengths, world)
            cols = shares[rank]
            C, S_cols = _size_fn(data, so, ro, reverse_complement, cols)
            xs, ys = triangle_jobs(cols)
            pos = {int(c): k for k, c in enumerate(cols)}
            vals = np.array([S_cols[x, pos[int(y)]] for x, y in zip(xs, ys)], dtype=np.int64)
            parts = _all_gather_padded(vals, dist) if dist else [vals]
            S = np.zeros((n, n), dtype=np.int64)
            for r, p in enumerate(parts):
                rx, ry = triangle_jobs(shares[r])
                S[rx, ry] = p
                S[ry, rx] = p
        C = np.asarray(C, dtype=np.int64)
        D = ncd_host(C, S, fast_mode)
    return [str(f) for f in files], C, np.asarray(S, dtype=np.int64), D


# ---- in-library multi-GPU entry: spawn one rank per GPU ------------------------------------------------------------
def run_multi_gpu(files, algorithm, reverse_complement, fast_mode, gpus):
    """``ncd_matrix(..., gpus=N)`` from a plain Python process: launches N ranks of ``snacc_b200._rank_worker`` with
    ``torch.distributed.run`` (one process per GPU, NCCL) and reads rank 0's result back."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory(prefix="snacc_b200_") as tmp:
        job = os.path.join(tmp, "job.npz")
        out = os.path.join(tmp, "out.npz")
        np.savez(job, files=np.array([str(f) for f in files]), algorithm=algorithm,
                 reverse_complement=bool(reverse_complement), fast_mode=bool(fast_mode))
        env = dict(os.environ)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={int(gpus)}",
               "--standalone", "--local-addr", "127.0.0.1", "-m", "snacc_b200._rank_worker", job, out]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(out):
            raise RuntimeError(f"multi-GPU run failed (exit {r.returncode}):\n{r.stderr[-4000:]}")
        z = np.load(out)
        return [str(f) for f in files], z["C"], z["S"], z["D"]
