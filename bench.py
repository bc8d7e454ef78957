#!/usr/bin/env python
"""bench.py -- headline benchmark of the snacc all-pairs NCD hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--codec lz4|gzip|zlib] [--config c4|c3|c5]

Workload (BASELINE.json configs[3], "c4"): 512 synthetic E. coli-sized (5 Mbp) mutated-phylogeny genomes,
LZ4-frame NCD.  A *step* is the WHOLE job: C(i) for all 512 genomes, C(i.j) for all 512 x 512 ordered
pairs (the reference's semantics, cli.py:104-136) and the float64 NCD matrix.  With N GPUs the job runs through the
product's own multi-GPU path (snacc_b200/sharding.py, the code `snacc --gpus N` / torchrun runs): the columns of the
job matrix are split into one contiguous band per rank -- all x against the rank's share of the y -- (strong
scaling; no data-path collective, the bands are all-gathered at the end of the step).  Metric: NCD pairs/s, where
-- as in SURVEY.md 8d -- a pair is an unordered {i,j} entry of the finished matrix and costs two ordered compressor
jobs, so pairs = ordered pair jobs / 2 (131072 per step).  `value` is timed with the corpus resident in HBM; `e2e`
starts from pinned HOST memory every step: each rank copies its 1/N band of the corpus to its GPU, the bands are
all-gathered over NVLink (NCCL), the corpus is re-packed, and sizes and distances are read back.  Every step
recomputes all per-genome prefix state (`invalidate_caches`): nothing is reused across steps.

The same JSON line carries a `gzip` object: the same corpus through the deflate (gzip level 9) kernels, timed the same
way (BASELINE.json's metric is quoted on "lz4, gzip"), and `parity`: the sizes of the timed step compared with the real
liblz4 / zlib on the sample of jobs the `cpu_baseline` leg compresses anyway -- a mismatch makes the run fail.
`host_s` reports the host-side stages that are outside the step (FASTA parsing of the whole configuration, CSV
writing).  One JSON line on stdout (rank 0).

`--impl reference` times the reference's own CPU compressor calls (system liblz4 / zlib through oracle/ref_codecs.c,
all host threads) on a bounded sample of the same workload, plus a `reference_verbatim` sub-leg: the reference's
per-job path as written (cli.py:104-129 -- a thread pool over jobs, every job re-reads and re-parses its FASTA
files, pairwise_ncd.py:29-36,59-90) on a small sample of files.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GENOMES = 512
GENOME_LEN = 5_000_000
SEED = 4            # config index 3 -> seed 4 (1-based), stated in config
CPU_SAMPLE_JOBS = {"lz4": 192, "gzip": 32, "zlib": 64}   # ordered pair jobs of the CPU sample (the deflate codecs are ~250x slower)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="snacc_b200")
    ap.add_argument("--codec", default="lz4", choices=["lz4", "gzip", "zlib"])
    ap.add_argument("--genomes", type=int, default=N_GENOMES)
    ap.add_argument("--length", type=int, default=GENOME_LEN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gzip-leg", action="store_true", help="skip the gzip sub-leg of the default run")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the c3 / c5 sub-legs of the default run")
    ap.add_argument("--no-host-stages", action="store_true", help="skip the FASTA-parse / CSV-write timings")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--fast-mode", action="store_true", help="upper triangle only (README --fast-mode True)")
    ap.add_argument("--exceptions", type=float, default=0.0,
                    help="fraction of bases replaced by N / IUPAC / lower-case bytes (real-assembly stand-in)")
    ap.add_argument("--config", default="c4", choices=["c4", "c3", "c5"],
                    help="c4 (default, the headline): 512 x 5 Mbp lz4; c3: 10,000 x ~11 kbp viral genomes (BASELINE.json configs[2]); "
                         "c5: 2,048 x 5 Mbp, gzip, the full ordered matrix C(xy) and C(yx) (BASELINE.json configs[4])")
    ap.add_argument("--opt", action="append", default=[], help="library tunable name=value (snacc_set_option), repeatable")
    ap.add_argument("--band", type=int, default=0, help="rows per library call (0 = the rank's whole band, 1024 for c3)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index
        self.nvml = None

    def _nvml_loop(self):
        # the same counters nvidia-smi prints, read through NVML in this process every 200 ms
        import pynvml as nv
        h = self.nvml
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
        while not self.quit.wait(0.2 if self.lines else 0.0):
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.lines.append(",".join([str(sm), str(mx), f"{pw:.1f}"] + ["Active" if rs & bit else "Not Active" for _, bit in names]))
            except Exception:
                break

    def start(self):
        # NVML in-process when the bindings are there (no second process attaching to the driver during the timed region);
        # nvidia-smi otherwise, or when SNACC_BENCH_NVIDIA_SMI is set
        if not os.environ.get("SNACC_BENCH_NVIDIA_SMI"):
            try:
                import pynvml as nv
                nv.nvmlInit()
                self.nvml = nv.nvmlDeviceGetHandleByIndex(self.index)
                self.quit = threading.Event()
                self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
                self.thread.start()
                self.source = "NVML (pynvml), 200 ms"
                return
            except Exception:
                self.nvml = None
        self.source = "nvidia-smi -lms 200"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.quit.set()
            self.thread.join(timeout=2)
        elif not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": getattr(self, "source", "nvidia-smi")}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- CPU legs (the only places that touch oracle/) -----------------------------------------------------------------
def cpu_sample(genomes_np, codec, n_jobs, threads, n_singles=0):
    """Reference compressor calls (system liblz4/zlib, all host threads) on pre-loaded sequences: jobs (0, j) and (1, j)
    of the workload, then `n_singles` singles (untimed share reported separately).  Returns a dict with the sizes."""
    import numpy as np
    from oracle import lib as olib
    need = min(len(genomes_np), max(2, n_jobs // 2))
    seqs = genomes_np[:need]
    so = np.zeros(len(seqs) + 1, dtype=np.uint64)
    so[1:] = np.cumsum([s.size for s in seqs])
    corpus = np.concatenate(seqs)
    xs = np.repeat(np.arange(2, dtype=np.int32), need)[:n_jobs]
    ys = np.tile(np.arange(need, dtype=np.int32), 2)[:n_jobs]
    olib.load()
    t = time.perf_counter()
    sizes = olib.ref_batch_sizes(corpus, so, xs, ys, codec, threads)
    dt = time.perf_counter() - t
    lens = np.diff(so.astype(np.int64))
    singles = None
    if n_singles:
        sx = np.arange(min(n_singles, need), dtype=np.int32)
        singles = olib.ref_batch_sizes(corpus, so, sx, np.full(sx.size, -1, np.int32), codec, threads)
    return {"seconds": dt, "jobs": int(xs.size), "bytes": float(lens[xs].sum() + lens[ys].sum()), "xs": xs, "ys": ys,
            "sizes": sizes, "singles": singles, "need": need}


def fast_fasta(path, name, seq, width=70):
    """one-record FASTA file, written without a Python loop over lines"""
    import numpy as np
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    full = seq.size // width
    body = np.empty((full, width + 1), dtype=np.uint8)
    body[:, :width] = seq[:full * width].reshape(full, width)
    body[:, width] = 10
    with open(path, "wb") as fh:
        fh.write(b">" + name.encode() + b"\n")
        body.tofile(fh)
        if seq.size > full * width:
            fh.write(seq[full * width:].tobytes() + b"\n")


def verbatim_leg(genomes_np, codec, threads, n_files):
    """The reference's own per-job path on FASTA files: thread pool over N singles + N^2 ordered pairs, every job
    re-reads and re-parses its files (cli.py:104-129, pairwise_ncd.py:29-36,59-90).  Runs the UNMODIFIED reference
    module when /root/reference is present (build container), its restatement oracle/snacc_oracle.py otherwise (the
    GPU box has no reference tree)."""
    from concurrent.futures import ThreadPoolExecutor
    from itertools import product
    from pathlib import Path
    from oracle import ref_loader, snacc_oracle
    tmp = tempfile.mkdtemp(prefix="snacc_verbatim_")
    try:
        files = []
        for i in range(min(n_files, len(genomes_np))):
            p = Path(tmp) / f"mysteryGenome_{i + 1}.fasta"
            fast_fasta(p, f"g{i}", genomes_np[i])
            files.append(p)
        if ref_loader.available():
            ref = ref_loader.load_reference_pairwise_ncd()
            fn, kind = (lambda s: ref.compressed_size(s, codec)), "reference"
        else:
            fn, kind = (lambda s: snacc_oracle.compressed_size(s, codec, False, backend="system")), "port"
        t = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            singles = list(ex.map(fn, files))
            pairs = list(ex.map(fn, product(files, repeat=2)))
        dt = time.perf_counter() - t
        n = len(files)
        return {"value": (n * n / 2) / dt, "unit": "pairs/s", "cores": threads, "kind": kind, "seconds": dt,
                "sample": f"{n} FASTA files of {genomes_np[0].size / 1e6:g} Mbp: {n} singles + {n * n} ordered pair jobs through "
                          f"compressed_size under ThreadPoolExecutor(max_workers={threads}), every job re-parsing its files "
                          f"({'unmodified /root/reference/snacc/pairwise_ncd.py' if kind == 'reference' else 'oracle/snacc_oracle.py restatement'}"
                          ", FASTA parser = oracle/fasta_shim.py since Biopython is not installed)",
                "checksum": int(sum(s for _, s in singles) + sum(s for _, s in pairs))}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_arm(args, config, codec, n, L):
    import numpy as np
    from snacc_b200 import synth
    threads = host_cores()

    def leg(cdc, steps, warmup):
        sample_jobs = CPU_SAMPLE_JOBS[cdc]
        need = min(n, max(2, sample_jobs // 2))
        genomes = synth.phylogeny(need, L, seed=SEED)
        for _ in range(warmup):
            cpu_sample(genomes, cdc, min(sample_jobs, 2 * threads), threads)
        tot_t, tot_jobs, tot_bytes = 0.0, 0, 0.0
        for _ in range(steps):
            r = cpu_sample(genomes, cdc, sample_jobs, threads)
            tot_t += r["seconds"]; tot_jobs += r["jobs"]; tot_bytes += r["bytes"]
        value = (tot_jobs / 2) / tot_t
        return {"value": value, "unit": "pairs/s", "ms_per_step": 1e3 * tot_t / steps, "steps": steps,
                "algorithmic_GBps": tot_bytes / tot_t / 1e9,
                "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "codec_only",
                                 "sample": f"{sample_jobs} ordered pair jobs (rows 0-1 x first {need} genomes) per step, system "
                                           f"{'liblz4 1.9.4' if cdc == 'lz4' else 'zlib 1.3'} via oracle/ref_codecs.c, sequences "
                                           "pre-loaded (no FASTA parsing), one thread per core"}}, genomes

    main, genomes = leg(codec, args.steps, args.warmup if codec == "lz4" else min(args.warmup, 1))
    line = {"impl": "reference", "metric": "ncd_pairs_per_s", "value": main["value"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config, "algorithmic_GBps": main["algorithmic_GBps"], "cpu_baseline": main["cpu_baseline"],
            "e2e": {"value": main["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if codec == "lz4" and not args.no_gzip_leg and args.config == "c4":
        g, _ = leg("gzip", max(1, min(args.steps, 2)), 0)
        line["gzip"] = g
    if not args.no_host_stages:
        line["reference_verbatim"] = verbatim_leg(genomes, codec, threads, 4 if L > 1_000_000 else 16)
    print(json.dumps(line))
    return 0


# ---- the GPU arm --------------------------------------------------------------------------------------------------
def inject_exceptions(corpus_dev, rate, seed):
    """replace a fraction `rate` of the bases by N, IUPAC codes and lower-case letters (what real assemblies contain)"""
    import torch
    g = torch.Generator(device=corpus_dev.device)
    g.manual_seed(seed + 1000)
    m = torch.rand(corpus_dev.numel(), generator=g, device=corpus_dev.device) < rate
    alt = torch.tensor(list(b"NNNNRYKMSWacgtn"), dtype=torch.uint8, device=corpus_dev.device)
    pick = alt[torch.randint(0, alt.numel(), (corpus_dev.numel(),), generator=g, device=corpus_dev.device)]
    return torch.where(m, pick, corpus_dev)


def describe(cfg_name, n, L, codec, world, fast_mode, seed):
    mode = "upper triangle only (--fast-mode: N + N(N+1)/2 jobs)" if fast_mode else \
        "ordered-pair matrix (N + N^2 compressor jobs, reference semantics cli.py:104-136)"
    workload = (f"{cfg_name}: {n} x {L / 1e6:g} Mbp synthetic mutated-phylogeny genomes, {codec} NCD; step = the whole "
                f"{n} x {n} {mode} + float64 NCD")
    return {"workload": workload, "n_genomes": n, "genome_len": L, "codec": codec,
            "seed": seed, "pair_unit": "ordered pair jobs / 2 (an unordered {i,j} costs two ordered jobs, cli.py:120-136); N^2/2 per step",
            "l2_policy": "inputs larger than L2 (corpus %.2f GB per GPU, replicated)" % (n * L / 1e9),
            "sharding": f"product path snacc_b200/sharding.py: contiguous column bands of the job matrix cut by bytes, one per rank "
                        f"({world} rank(s): all x against the rank's y), no data-path collective; bands all-gathered at the end "
                        "of the step; deflate codecs: per-sequence preparation sharded, 680-byte prefix records all-gathered; "
                        "e2e: each rank uploads 1/N of the corpus, NCCL all-gather over NVLink"}


def run_workload(args, cfg_name, n, L, seed, legs, env, want_host_stages):
    """One synthetic corpus of the named shape on the GPU(s); `legs` = [(codec, steps, warmup, with_cpu)] timed one after
    the other on it.  Returns (reports per codec -- rank 0 only --, corpus generation seconds, host stage timings,
    parity failures)."""
    import numpy as np
    import torch
    from snacc_b200 import sharding, synth
    from snacc_b200.engine import Engine
    rank, local_rank, world, dev, ddist = env["rank"], env["local_rank"], env["world"], env["dev"], env["ddist"]
    peak, peak_src, traffic_tab = env["peak"], env["peak_src"], env["traffic_tab"]
    SEED = seed
    # ---- synthetic corpus: generated on the device (same seed on every rank), then mirrored to pinned host ----
    t0 = time.perf_counter()
    genomes = synth.phylogeny_torch(n, L, SEED, dev)
    lengths = np.array([g.numel() for g in genomes], dtype=np.int64)
    so = np.zeros(n + 1, dtype=np.uint64)
    so[1:] = np.cumsum(lengths)
    corpus_dev = torch.cat(genomes)
    del genomes
    if args.exceptions:
        corpus_dev = inject_exceptions(corpus_dev, args.exceptions, SEED)
    corpus_host = torch.empty(corpus_dev.numel(), dtype=torch.uint8, pin_memory=True)
    corpus_host.copy_(corpus_dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    eng = Engine(local_rank)
    for kv in args.opt:
        name, val = kv.split("=")
        eng.set_option(name, int(val))
    eng.upload_device(corpus_dev.data_ptr(), so)
    del corpus_dev
    torch.cuda.empty_cache()
    band_lo, band_hi = rank * n // world, (rank + 1) * n // world         # this rank's share of the corpus upload (e2e)
    band_bytes = corpus_host[int(so[band_lo]):int(so[band_hi])]
    band_lens = lengths[band_lo:band_hi].tolist()

    def barrier():
        torch.cuda.synchronize()
        if ddist:
            ddist.barrier()
        torch.cuda.synchronize()

    # page-locked result buffers, reused by every step (pinning 800 MB per step would cost more than the c3 kernels)
    bounds = sharding.band_bounds(lengths, world)
    my_w = int(bounds[rank + 1] - bounds[rank])
    S_band = sharding.pinned_array((n, my_w), np.int64) if not args.fast_mode else None
    D_buf = torch.empty((n, n), dtype=torch.float64, pin_memory=True).numpy()
    S_full = sharding.pinned_array((n, n), np.int64) if world > 1 and not args.fast_mode else None

    def run_step(cdc, e2e):
        """one full pass through the product path: [e2e: corpus from pinned host memory,] C(i) for every sequence, S for
        the rank's share of the job matrix, gather, float64 NCD (device kernel, read back)"""
        h2d = 0
        if e2e:
            if ddist:
                t, so2, ro2 = sharding.exchange_corpus(band_bytes, band_lens, band_lens, ddist, dev)   # H2D of 1/N + all-gather
                torch.cuda.synchronize(dev)
                eng.upload_device(t.data_ptr(), so2)
                del t
            else:
                eng.upload(corpus_host.numpy(), so)      # H2D of the step's inputs from pinned memory (+ repack)
            h2d = int(band_bytes.numel())
        else:
            eng.set_option("invalidate_caches", 1)       # nothing (prefix checkpoints ...) survives from the last step
        st = {}
        band = args.band or (1024 if cfg_name == "c3" else None)
        t0 = time.perf_counter()
        C, S = sharding.sizes_matrix(eng, cdc, args.fast_mode, band, st, band_out=S_band, full_out=S_full)
        t1 = time.perf_counter()
        D = eng.ncd(C, S, formula=1 if args.fast_mode else 0, out=D_buf)   # K4: float64 epilogue kernel, result read back
        t2 = time.perf_counter()
        st["launches"] = st.get("launches", 0) + 1
        st["check"] = int(S[::7, ::5].sum() + C.sum()) ^ int(np.float64(D[::7, ::5].sum()).view(np.int64) & 0xffff)
        if os.environ.get("SNACC_BENCH_TRACE"):
            print(f"[trace rank {rank}] {cfg_name} {cdc} e2e={int(e2e)}: sizes_matrix {1e3 * (t1 - t0):.0f} ms (kernels "
                  f"{st.get('kernel_ms', 0):.0f}, gather {1e3 * st.get('gather_s', 0):.0f}), ncd {1e3 * (t2 - t1):.0f} ms, "
                  f"checksum {1e3 * (time.perf_counter() - t2):.0f} ms", file=sys.stderr, flush=True)
        st["h2d"] = h2d
        st["C"], st["S"] = C, S
        return st

    def time_leg(cdc, steps, warmup, with_e2e):
        for _ in range(warmup):
            run_step(cdc, False)
        sampler = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            sampler.start()
        t_start = time.perf_counter()
        stats = [run_step(cdc, False) for _ in range(steps)]
        barrier()
        wall = time.perf_counter() - t_start
        clocks = sampler.stop() if rank == 0 else None
        e2e_wall = float("nan")
        if with_e2e:
            run_step(cdc, True)
            barrier()
            t_e = time.perf_counter()
            for _ in range(steps):
                run_step(cdc, True)
            barrier()
            e2e_wall = time.perf_counter() - t_e
        dev_ms = sum(s["kernel_ms"] for s in stats)
        t = torch.tensor([dev_ms * 1e-3, wall, e2e_wall if with_e2e else 0.0], dtype=torch.float64, device=dev)
        if ddist:
            ddist.all_reduce(t, op=ddist.ReduceOp.MAX)
        dev_s, wall_s, e2e_s = [float(v) for v in t.tolist()]
        return stats, clocks, dev_s, wall_s, (e2e_s if with_e2e else None)

    host_np = corpus_host.numpy()
    parity_failed = []

    def leg_report(cdc, steps, warmup, with_cpu, cfg_name):
        stats, clocks, dev_s, wall_s, e2e_s = time_leg(cdc, steps, warmup, not args.no_e2e)
        if rank != 0:
            return None
        n_pair_jobs = n * (n + 1) // 2 if args.fast_mode else n * n
        pairs_per_step = n_pair_jobs / 2 if not args.fast_mode else n * (n + 1) / 2
        bytes_step = (float(np.sum(lengths)) * 2 * n if not args.fast_mode else
                      float(np.sum(np.cumsum(lengths) + lengths * np.arange(1, n + 1))))
        main_ms = sum(s["main_ms"] for s in stats) / len(stats)
        per_launch_bytes = stats[0]["bytes"]             # rank 0's dominant-kernel launches: sum of len(x)+len(y) over its jobs
        achieved = per_launch_bytes / (main_ms * 1e-3) / 1e9 if main_ms > 0 else None
        kernel = (("lz4_pk_pair_kernel<linked>" if cfg_name != "c3" else "lz4_pk_pair_kernel<single-block>")
                  if cdc == "lz4" else "dfl_junction_kernel (sum over the step's batches)")
        rep = {"value": pairs_per_step * steps / wall_s, "unit": "pairs/s", "steps": steps, "warmup": warmup,
               "ms_per_step": 1e3 * wall_s / steps, "algorithmic_GBps": bytes_step * steps / wall_s / 1e9,
               "device_ms_per_step": 1e3 * max(dev_s, 1e-9) / steps,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak if achieved else None,
                            "traffic": traffic_tab.get(f"{cdc}:{cfg_name}:{n}x{L}:n{world}"), "kernel": kernel,
                            "peak_source": peak_src,
                            "note": "achieved = algorithmic bytes of the kernel's launches in one step (sum of len(x)+len(y) over "
                                    "their pair jobs) / their CUDA-event duration; the path is latency/integer bound, not HBM "
                                    "bound; traffic (when not null) = DRAM bytes of those launches measured by ncu "
                                    "(profiles/traffic.json): far below the algorithmic bytes because y is staged once per "
                                    "CTA and x only enters through its checkpoint"},
               "gpu_launches": int(sum(s["launches"] for s in stats)),
               "packed_jobs_per_step": stats[0].get("packed_jobs", 0), "bytewise_jobs_per_step": stats[0].get("bytewise_jobs", 0),
               "clocks": clocks, "checksum": stats[-1]["check"]}
        if e2e_s is not None:
            d2h = 8 * (stats[0]["jobs"] + n) + (8 * n * n if world > 1 else 0) + 8 * n * n
            rep["e2e"] = {"value": pairs_per_step * steps / e2e_s, "unit": "pairs/s",
                          "h2d_bytes_per_step": int(corpus_host.numel() + (so.nbytes + 4 * n) * world + 8 * (n * n + n)),
                          "d2h_bytes_per_step": int(d2h),
                          "note": "per step: every rank copies its 1/N band of the corpus from pinned host memory, NCCL all-gather, "
                                  "re-pack, all kernels, sizes and D read back; bytes summed over ranks" if world > 1 else
                                  "per step: corpus H2D from pinned host memory through snacc_upload, re-pack, all kernels, sizes and D read back"}
        # ---- parity gate: the timed step's sizes against the real library on a sample (also the cpu_baseline timing) ----
        C, S = stats[-1]["C"], stats[-1]["S"]
        threads = host_cores()
        jobs = CPU_SAMPLE_JOBS[cdc] if (with_cpu or cdc == "lz4") else 8
        need = max(2, jobs // 2)
        host_genomes = [host_np[int(so[i]):int(so[i + 1])] for i in range(min(n, need))]
        r = cpu_sample(host_genomes, cdc, jobs, threads, n_singles=min(need, 96 if cdc == "lz4" else 2))
        wb = {"lz4": 0, "gzip": 0, "zlib": 0}[cdc]      # both sides report the full compressed length
        got = S[r["xs"], r["ys"]]
        keep = (r["xs"] <= r["ys"]) if args.fast_mode else np.ones(got.size, dtype=bool)
        mism = int(np.sum((got != r["sizes"] + wb) & keep))
        checked = int(keep.sum())
        if r["singles"] is not None:
            mism += int(np.sum(C[:r["singles"].size] != r["singles"] + wb))
            checked += int(r["singles"].size)
        rep["parity"] = {"checked": checked, "mismatches": mism,
                         "against": f"system {'liblz4 1.9.4' if cdc == 'lz4' else 'zlib 1.3'} (oracle/ref_codecs.c) on rows 0-1 x first "
                                    f"{r['need']} genomes + {0 if r['singles'] is None else r['singles'].size} singles of the last step"}
        if mism:
            parity_failed.append((cdc, mism, checked))
        if with_cpu:
            rep["cpu_baseline"] = {"value": (r["jobs"] / 2) / r["seconds"], "unit": "pairs/s", "cores": threads,
                                   "kind": "codec_only", "algorithmic_GBps": r["bytes"] / r["seconds"] / 1e9,
                                   "sample": f"{r['jobs']} ordered pair jobs (rows 0-1 x first {r['need']} genomes), system "
                                             f"{'liblz4 1.9.4' if cdc == 'lz4' else 'zlib 1.3'} via oracle/ref_codecs.c, "
                                             "sequences pre-loaded (no FASTA parsing: the reference's own per-job path is the "
                                             "`reference_verbatim` leg of --impl reference), one thread per core"}
        return rep

    reports = {}
    for cdc, steps, warmup, with_cpu in legs:
        reports[cdc] = leg_report(cdc, steps, warmup, with_cpu, cfg_name)
    host_s = None
    if want_host_stages and rank == 0:
        host_s = host_stages(host_np, so, n, None)
    eng.close()
    del corpus_host, band_bytes, host_np, S_band, D_buf, S_full
    torch.cuda.empty_cache()
    return reports, gen_s, host_s, parity_failed


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    codec = args.codec
    if args.config == "c3" and args.genomes == N_GENOMES and args.length == GENOME_LEN:
        args.genomes, args.length = 10_000, 10_700
    if args.config == "c5":
        codec = args.codec = "gzip"
        if args.genomes == N_GENOMES:
            args.genomes = 2048
    n, L = args.genomes, args.length
    config = describe(args.config, n, L, codec, world, args.fast_mode, SEED)
    if args.exceptions:
        config["exceptions"] = f"{args.exceptions:g} of the bases replaced by N / IUPAC / lower-case bytes"

    import numpy as np

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, config, codec, n, L)

    import torch
    import torch.distributed as dist
    from snacc_b200 import sharding, synth
    from snacc_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"                # keep NCCL's version banner off stdout: ONE JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(tempfile.gettempdir(), "snacc_bench_nccl.%h.%p.log"))
    sharding.init_distributed()                          # the product's own entry: NCCL group + GPU LOCAL_RANK
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ddist = sharding._dist()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic_tab = {}
    env = {"rank": rank, "local_rank": local_rank, "world": world, "dev": dev, "ddist": ddist, "peak": peak,
           "peak_src": peak_src, "traffic_tab": traffic_tab}

    with_cpu = (not args.no_cpu_baseline) and world == 1
    default_run = codec == "lz4" and args.config == "c4" and not args.fast_mode and n == N_GENOMES and L == GENOME_LEN
    legs = [(codec, args.steps, args.warmup, with_cpu)]
    if codec == "lz4" and args.config == "c4" and not args.no_gzip_leg and not args.fast_mode:
        legs.append(("gzip", max(1, min(args.steps, 5)), min(args.warmup, 3), with_cpu))
    reports, gen_s, host_s, failed = run_workload(args, args.config, n, L, SEED, legs, env,
                                                  world == 1 and not args.no_host_stages)
    extra = {}
    if default_run and not args.no_extra_legs:
        # the other configurations the metric is quoted on (BASELINE.json configs[2] and configs[4]), a few steps each
        r3, g3, _, f3 = run_workload(args, "c3", 10_000, 10_700, 3, [("lz4", max(1, min(args.steps, 3)), 1, with_cpu)], env, False)
        r5, g5, _, f5 = run_workload(args, "c5", 2048, GENOME_LEN, 5, [("gzip", 1, 1, False)], env, False)
        failed = failed + f3 + f5
        if rank == 0:
            r3["lz4"]["config"] = describe("c3", 10_000, 10_700, "lz4", world, False, 3)
            r5["gzip"]["config"] = describe("c5", 2048, GENOME_LEN, "gzip", world, False, 5)
            extra = {"c3": r3["lz4"], "c5": r5["gzip"]}
    if rank != 0:
        sharding.shutdown_distributed()
        return 0
    line = {"metric": "ncd_pairs_per_s", "n_gpus": world, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config, "corpus_gen_s": gen_s}
    line.update(reports[codec])
    if "gzip" in reports and codec != "gzip":
        reports["gzip"]["config"] = {"workload": config["workload"].replace(f"{codec} NCD", "gzip NCD"), "codec": "gzip",
                                     "note": "same corpus, same step definition, deflate level 9 kernels (gzip.compress, pairwise_ncd.py:74)"}
        line["gzip"] = reports["gzip"]
    line.update(extra)
    if host_s is not None:
        line["host_s"] = host_s
    print(json.dumps(line))
    sharding.shutdown_distributed()
    if failed:
        print(f"PARITY FAILURE: {failed}", file=sys.stderr)
        return 3
    return 0


def host_stages(host_np, so, n, rep):
    """Host-side stages of the configuration that sit outside the step: parsing the N FASTA files once
    (snacc_b200.fasta.load_corpus: native parser, 16 threads) and writing the N x N CSV (cli.write_distance_csv)."""
    import numpy as np
    from snacc_b200 import fasta
    from snacc_b200.cli import write_distance_csv
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    need = int(so[n]) * 1.02 + (1 << 20)
    m = n
    try:
        free = shutil.disk_usage(base).free
        while m > 8 and int(so[m]) * 1.02 > 0.5 * free:
            m //= 2
    except Exception:
        m = min(n, 64)
    tmp = tempfile.mkdtemp(prefix="snacc_bench_", dir=base)
    try:
        t = time.perf_counter()
        files = []
        for i in range(m):
            p = os.path.join(tmp, f"genome_{i:05d}.fasta")
            fast_fasta(p, f"genome_{i}", host_np[int(so[i]):int(so[i + 1])])
            files.append(p)
        write_s = time.perf_counter() - t
        t = time.perf_counter()
        data, so2, ro2 = fasta.load_corpus(files)
        parse_s = time.perf_counter() - t
        ok = bool(data.size == int(so[m]) and np.array_equal(data[:1 << 20], host_np[:1 << 20]))
        del data
        D = np.random.default_rng(0).random((n, n))
        t = time.perf_counter()
        write_distance_csv([os.path.join(tmp, f"genome_{i:05d}.fasta") for i in range(n)], D, os.path.join(tmp, "d.csv"))
        csv_s = time.perf_counter() - t
        return {"fasta_parse": parse_s * (n / m), "fasta_files_parsed": m, "fasta_files_total": n,
                "fasta_bytes": int(so[m]), "fasta_roundtrip_ok": ok, "fasta_write_untimed": write_s, "csv_write": csv_s,
                "note": "outside the step: the product parses every FASTA file ONCE per run (the reference re-parses per job); "
                        "fasta_parse is scaled from the files parsed when the scratch disk could not hold all of them"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
