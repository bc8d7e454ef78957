#!/usr/bin/env python
"""bench.py -- headline benchmark of the snacc all-pairs NCD hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--codec lz4|gzip]

Workload (BASELINE.json configs[3], "c4"): 512 synthetic E. coli-sized (5 Mbp) mutated-phylogeny genomes,
LZ4-frame NCD.  A *step* is the WHOLE job: C(i) for all 512 genomes, C(i.j) for all 512 x 512 ordered
pairs (the reference's semantics, cli.py:104-136) and the float64 NCD matrix.  With N GPUs the columns of
the job matrix are split into one contiguous band per rank -- all x against the rank's share of the y --
(strong scaling; no data-path collective, the bands are all-gathered at the end of the step).  Metric: NCD pairs/s, where -- as in SURVEY.md 8d -- a
pair is an unordered {i,j} entry of the finished matrix and costs two ordered compressor jobs, so
pairs = ordered pair jobs / 2 (131072 per step).  `value` is timed with the corpus resident in HBM;
`e2e` re-uploads the corpus from pinned host memory through the C ABI (and re-packs it) and reads the
sizes back, every step.  Every step recomputes all per-genome prefix state (`invalidate_caches`):
nothing is reused across steps.  One JSON line on stdout (rank 0).

`--impl reference` times the reference's own CPU compressor calls (system liblz4 / zlib through
oracle/ref_codecs.c, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GENOMES = 512
GENOME_LEN = 5_000_000
SEED = 4            # config index 3 -> seed 4 (1-based), stated in config
CPU_SAMPLE_JOBS = 192          # lz4; the deflate codecs are ~250x slower on the CPU: 32 jobs


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
    ap.add_argument("--config", default="c4", choices=["c4", "c3", "c5"],
                    help="c4 (default, the headline): 512 x 5 Mbp lz4; c3: 10,000 x ~11 kbp viral genomes (BASELINE.json configs[2]); "
                         "c5: 2,048 x 5 Mbp, gzip, the full ordered matrix C(xy) and C(yx) (BASELINE.json configs[4])")
    ap.add_argument("--band", type=int, default=0, help="rows per library call (0 = the rank's whole band, 1024 for c3)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
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
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def job_bytes(lengths, xs, ys):
    import numpy as np
    return float(np.sum(lengths[xs]) + np.sum(lengths[ys]))


def cpu_sample_jobs(codec):
    return CPU_SAMPLE_JOBS if codec == "lz4" else 32


def cpu_reference_sample(genomes_np, codec, n_jobs, threads):
    """Reference compressor calls (system liblz4/zlib, all host threads) on pre-loaded sequences:
    jobs (0, j) and (1, j) of the workload.  Returns (seconds, jobs, algorithmic bytes)."""
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
    olib.ref_batch_sizes(corpus, so, xs, ys, codec, threads)
    dt = time.perf_counter() - t
    lens = np.diff(so.astype(np.int64))
    return dt, int(xs.size), job_bytes(lens, xs, ys)


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
    workload = (f"{args.config}: {n} x {L / 1e6:g} Mbp synthetic mutated-phylogeny genomes, {codec} NCD; step = the whole "
                f"{n} x {n} ordered-pair matrix (N + N^2 compressor jobs, reference semantics cli.py:104-136) + float64 NCD")
    config = {"workload": workload, "n_genomes": n, "genome_len": L, "codec": codec,
              "seed": SEED, "pair_unit": "ordered pair jobs / 2 (an unordered {i,j} costs two ordered jobs, cli.py:120-136); N^2/2 per step",
              "l2_policy": "inputs larger than L2 (corpus %.2f GB per GPU, replicated)" % (n * L / 1e9),
              "sharding": f"contiguous column band [r*N/{world}, (r+1)*N/{world}) of the job matrix per rank (all x against the "
                          "rank's y), no data-path collective; bands gathered with all_gather at the end of the step"}

    import numpy as np

    if args.impl == "reference":
        if rank != 0:
            return 0
        from snacc_b200 import synth
        threads = host_cores()
        sample_jobs = cpu_sample_jobs(codec)
        need = max(2, sample_jobs // 2)
        genomes = synth.phylogeny(min(n, need), L, seed=SEED)
        for _ in range(args.warmup if codec == "lz4" else min(args.warmup, 1)):
            cpu_reference_sample(genomes, codec, min(sample_jobs, 2 * threads), threads)
        tot_t, tot_jobs, tot_bytes = 0.0, 0, 0.0
        for _ in range(args.steps):
            dt, jobs, nbytes = cpu_reference_sample(genomes, codec, sample_jobs, threads)
            tot_t += dt; tot_jobs += jobs; tot_bytes += nbytes
        value = (tot_jobs / 2) / tot_t
        sample = f"{sample_jobs} ordered pair jobs (rows 0-1 x first {need} genomes) per step, sequences pre-loaded"
        line = {"impl": "reference", "metric": "ncd_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config, "algorithmic_GBps": tot_bytes / tot_t / 1e9,
                "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference",
                                 "sample": sample},
                "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from snacc_b200 import synth
    from snacc_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"                # keep NCCL's version banner off stdout: ONE JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic corpus: generated on the device (same seed on every rank), then mirrored to pinned host ----
    t0 = time.perf_counter()
    genomes = synth.phylogeny_torch(n, L, SEED, dev)
    lengths = np.array([g.numel() for g in genomes], dtype=np.int64)
    so = np.zeros(n + 1, dtype=np.uint64)
    so[1:] = np.cumsum(lengths)
    corpus_dev = torch.cat(genomes)
    del genomes
    corpus_host = torch.empty(corpus_dev.numel(), dtype=torch.uint8, pin_memory=True)
    corpus_host.copy_(corpus_dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    eng = Engine(local_rank)
    eng.upload_device(corpus_dev.data_ptr(), so)
    del corpus_dev
    torch.cuda.empty_cache()

    # contiguous COLUMN band per rank (all x against the rank's share of the y): a tile of the LZ4 kernel is "one y,
    # up to 104 x", so with every row on every rank the tiles stay full for any number of ranks
    my_cols = np.arange(rank * n // world, (rank + 1) * n // world, dtype=np.int32)
    all_rows = np.arange(n, dtype=np.int32)
    step_jobs = int(my_cols.size) * n
    step_bytes = float(n * np.sum(lengths[my_cols]) + my_cols.size * np.sum(lengths))

    def run_step(e2e):
        """one full pass: C(i) for every sequence, S(i,j) for all rows x the rank's columns, gather, NCD on rank 0"""
        if e2e:
            eng.upload(corpus_host.numpy(), so)          # H2D of the step's inputs from pinned memory (+ repack)
        else:
            eng.set_option("invalidate_caches", 1)       # nothing (prefix checkpoints ...) survives from the last step
        c = eng.single_sizes(codec, all_rows)            # every rank: the same pass leaves the checkpoints of every x
        ms1, l1 = eng.stat("total_kernel_ms"), eng.stat("launches")
        band = args.band or (1024 if args.config == "c3" else n)
        s = np.empty((n, my_cols.size), dtype=np.int64)
        ms2 = l2 = main_ms = packed = 0
        for a in range(0, n, band):                      # sizes come back to the host inside (D2H)
            nb = min(band, n - a)
            s[a:a + nb] = eng.tile_sizes(codec, a, nb, int(my_cols[0]), int(my_cols.size))
            ms2 += eng.stat("total_kernel_ms"); l2 += eng.stat("launches"); main_ms += eng.stat("main_kernel_ms")
            packed += eng.stat("packed_jobs") if codec == "lz4" else 0
        if world > 1:
            # gather the column bands (transposed: rank r owns columns [r*n/world, (r+1)*n/world); 2 MiB of int64 at n = 512)
            per = (n + world - 1) // world
            sbuf = torch.zeros(per * n, dtype=torch.int64, device=dev)
            sbuf[:s.size] = torch.from_numpy(np.ascontiguousarray(s.T).ravel()).to(dev)
            sg = [torch.empty_like(sbuf) for _ in range(world)]
            dist.all_gather(sg, sbuf)
            C = c; S = np.zeros((n, n), dtype=np.int64)
            for r in range(world):
                cc = np.arange(r * n // world, (r + 1) * n // world)
                S[:, cc] = sg[r][:cc.size * n].cpu().numpy().reshape(cc.size, n).T
        else:
            C, S = c, s.reshape(n, n)
        launches = int(l1 + l2)
        check = 0
        if rank == 0:
            D = eng.ncd(C, S)                            # float64 epilogue kernel, result read back
            launches += 1
            check = int(S.sum() + C.sum()) ^ int(np.float64(D.sum()).view(np.int64) & 0xffff)
        return {"kernel_ms": ms1 + ms2, "main_ms": main_ms, "launches": launches, "jobs": step_jobs,
                "bytes": step_bytes, "check": check, "packed": int(packed)}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        run_step(False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_start = time.perf_counter()
    stats = [run_step(False) for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t_start
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(s["kernel_ms"] for s in stats)
    # ---- e2e: host buffers, H2D + D2H inside the timed region ----
    run_step(True)
    barrier()
    t_e = time.perf_counter()
    for k in range(args.steps):
        run_step(True)
    barrier()
    e2e_wall = time.perf_counter() - t_e

    t = torch.tensor([dev_ms * 1e-3, wall, e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s, e2e_s = [float(v) for v in t.tolist()]
    pairs_per_step = n * n / 2                           # ordered pair jobs / 2 (same unit as the reference arm)
    bytes_total = float(np.sum(lengths)) * 2 * n * args.steps
    timed_s = max(dev_s, 1e-9)
    value = pairs_per_step * args.steps / wall_s
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    main_ms = sum(s["main_ms"] for s in stats) / len(stats)
    traffic = None                                       # DRAM bytes per launch of the dominant kernel, from a committed ncu capture
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{codec}:{args.config}:{n}x{L}:n{world}")
    except Exception:
        pass
    per_launch_bytes = stats[0]["bytes"]                 # rank 0's pair-kernel launch: sum of len(x)+len(y) over its jobs
    achieved = per_launch_bytes / (main_ms * 1e-3) / 1e9
    line = {"metric": "ncd_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall_s / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "algorithmic_GBps": bytes_total / wall_s / 1e9,
            "device_ms_per_step": 1e3 * timed_s / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": ("lz4_pk_pair_kernel<linked>" if args.config != "c3" else "lz4_pk_pair_kernel<single-block>")
                                   if codec == "lz4" else "dfl_junction_kernel (sum over the step's batches)",
                         "peak_source": peak_src,
                         "note": "achieved = algorithmic bytes of one launch (sum of len(x)+len(y) over its pair jobs) / "
                                 "its CUDA-event duration; the path is latency/integer bound, not HBM bound; traffic (when not "
                                 "null) = DRAM bytes of that launch measured by ncu (profiles/traffic.json): far below the "
                                 "algorithmic bytes because y is staged once per CTA and x only enters through its checkpoint"},
            "e2e": {"value": pairs_per_step * args.steps / e2e_s, "unit": "pairs/s",
                    "h2d_bytes_per_step": int(corpus_host.numel() + so.nbytes + 4 * n + 8 * (n * n + n)),
                    "d2h_bytes_per_step": int(8 * (stats[0]["jobs"] + n) + 8 * n * n)},
            "packed_jobs_per_step": stats[0]["packed"],
            "gpu_launches": int(sum(s["launches"] for s in stats)),
            "clocks": clocks, "corpus_gen_s": gen_s, "checksum": stats[-1]["check"]}
    if not args.no_cpu_baseline and world == 1:
        threads = host_cores()
        need = max(2, cpu_sample_jobs(codec) // 2)
        host_genomes = [corpus_host.numpy()[int(so[i]):int(so[i + 1])] for i in range(min(n, need))]
        dt, jobs, nbytes = cpu_reference_sample(host_genomes, codec, cpu_sample_jobs(codec), threads)
        line["cpu_baseline"] = {"value": (jobs / 2) / dt, "unit": "pairs/s", "cores": threads, "kind": "reference",
                                "algorithmic_GBps": nbytes / dt / 1e9,
                                "sample": f"{jobs} ordered pair jobs (rows 0-1 x first {need} genomes), system "
                                          f"{'liblz4 1.9.4' if codec == 'lz4' else 'zlib 1.3'} via oracle/ref_codecs.c, "
                                          "sequences pre-loaded, one thread per core"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
