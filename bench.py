#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path built so far: the getSeeds k-mer lookup sweep.

Metric (BASELINE.json, second half): k-mer lookups/s vs the HBM roofline.  A *lookup* is one
variant-string k-mer window of CompactedDBG::searchSequence as getSeeds drives it in pass 1
(src/Graph.cpp:97,193): the exact sweep plus the 9k+1 one-edit sweeps = 1 + ~250.5 windows per read
base at k=31.  The count is a closed-form function of the read lengths (`nominal_lookups`), identical
for both arms.  (The first half of the metric, corrected bases/s, needs the traversal + alignment
stages that are not wired end to end yet; it is reported as soon as they are.)

Workload = BASELINE.json configs[1] shape: E. coli-like 4.64 Mbp genome, k=31, ONT-like reads at 10 %
error, synthetic (seeded numpy; no network).  A step = one batch of reads through the sweep.
  value : K1 exact + inexact kernels over a batch already resident in HBM
  e2e   : rtk_get_seeds through the C ABI from pinned host buffers (H2D of reads, kernels, D2H of hits,
          host-side anchor extraction) - anchors identical to the reference's getSeeds
  --impl reference : the reference's own getSeeds (oracle/_ref/libref_seams.so, unmodified objects) on
          all host threads, on a bounded sample of the same reads
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 31
GENOME_LEN = 4_640_000
LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


# ----------------------------------------------------------------------------- synthetic workload
def make_genome(seed=4640):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=GENOME_LEN, dtype=np.uint8)


def genome_unitigs(genome, seed=4641):
    """cut the genome into unitigs overlapping by k-1 (what SNP / error bubbles do to a real graph)"""
    rng = np.random.default_rng(seed)
    cuts = [0]
    while cuts[-1] < len(genome) - 40:
        cuts.append(min(len(genome) - (K - 1), cuts[-1] + int(rng.integers(40, 6000))))
    seqs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        seqs.append(LUT[genome[a:b + K - 1]].tobytes())
    return seqs


def make_reads(genome, total_bases, seed):
    """ONT-like reads: lognormal(9.0, 0.6) >= 1 kb, 3 % sub / 2.5 % ins / 4.5 % del, random strand"""
    rng = np.random.default_rng(seed)
    pool, offs, tot = [], [0], 0
    while tot < total_bases:
        ln = int(min(max(1000, rng.lognormal(9.0, 0.6)), 200_000))
        p = int(rng.integers(0, len(genome) - ln))
        s = genome[p:p + ln]
        if rng.random() < 0.5:
            s = (3 - s)[::-1]
        r = rng.random(ln)
        keep = r >= 0.045
        s = s[keep].copy()
        sub = r[keep] < 0.075
        s[sub] = (s[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
        ins = rng.random(len(s)) < 0.025
        cnt = 1 + ins.astype(np.int64)
        out = np.repeat(s, cnt)
        ipos = np.cumsum(cnt)[ins] - 1
        out[ipos] = rng.integers(0, 4, size=len(ipos), dtype=np.uint8)
        pool.append(LUT[out])
        tot += len(out)
        offs.append(tot)
    return np.concatenate(pool), np.asarray(offs, dtype=np.uint64)


def nominal_lookups(read_lens, k=K):
    """windows of the exact sweep + the 9k+1 variant strings of searchSequence (Search.tcc:685-765)"""
    total = 0
    L = np.asarray(read_lens, dtype=np.int64)
    L = L[L >= k]
    total += int((L - k + 1).sum())                       # exact
    total += int((3 * k * (L - k + 1)).sum())             # substitution: 3 letters x k slot offsets per window
    for i in range(k):                                    # insertion strings: 4 letters each
        li = L + (L - i + k - 2) // (k - 1)
        total += int((4 * np.maximum(li - k + 1, 0)).sum())
    Ld = L[L >= k + 1]
    for i in range(k + 1):                                # deletion strings
        li = Ld - (np.maximum(Ld - i, 0) + k) // (k + 1)
        total += int(np.maximum(li - k + 1, 0).sum())
    return total


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """the reference's own getSeeds (unmodified objects) on all host threads, bounded sample per step"""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    import refseams as R
    cores = os.cpu_count() or 1
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_seams.so not built (needs /root/reference at build time)"}))
        return
    genome = make_genome()
    unitigs = genome_unitigs(genome)
    tmp = tempfile.mkdtemp(prefix="rtk_bench_")
    fa = os.path.join(tmp, "graph.fasta")
    with open(fa, "w") as f:
        for i, u in enumerate(unitigs):
            f.write(">%d\n%s\n" % (i, u.decode()))
    g = R.RefGraph(fa, "", K, threads=min(cores, 16))
    sample_bases = args.ref_sample_bases
    times, looks = [], []
    for step in range(args.warmup + args.steps):
        pool, off = make_reads(genome, sample_bases, seed=1000 + step)
        reads = [pool[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(lambda s: g.get_seeds(s, "", False), reads))
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
            looks.append(nominal_lookups(np.diff(off.astype(np.int64))))
    total_t, total_l = sum(times), sum(looks)
    val = total_l / total_t
    line = {"metric": "kmer_lookups_per_s", "value": val, "unit": "lookups/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, sample_bases),
            "cpu_baseline": {"value": val, "unit": "lookups/s", "cores": cores, "kind": "reference",
                             "sample": "%d bases of reads per step through the reference getSeeds (libref_seams.so), %d threads" % (sample_bases, cores)},
            "e2e": {"value": val, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, bases_per_step):
    return {"workload": "E. coli-like 4.64 Mbp genome, k=31, pass-1 getSeeds lookup sweep (exact + 9k+1 one-edit passes), "
                        "ONT-like reads lognormal(9.0,0.6) >= 1 kb at 10% error",
            "bases_per_step_per_gpu": int(bases_per_step), "k": K,
            "lookup_definition": "variant-string k-mer windows of searchSequence: 1 + ~250.5 per read base",
            "l2_policy": "256 MiB buffer written between timed iterations (L2 flush)"}


# ----------------------------------------------------------------------------- product arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases-per-step", type=int, default=32 << 20)
    ap.add_argument("--ref-sample-bases", type=int, default=4_000_000)
    ap.add_argument("--cpu-baseline-bases", type=int, default=16_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import ratatosk_b200 as rb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path exists in the product)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = rb.load_library()

    # ---- graph: built on rank 0, one H2D, ONE NCCL broadcast of the slab, adopted in place by every rank
    genome = make_genome()
    ctx = rb.Context(local_rank)
    if rank == 0:
        g = rb.Graph.from_unitigs(genome_unitigs(genome), K)
        slab = torch.from_numpy(g.slab())
        nbytes = torch.tensor([slab.numel()], dtype=torch.int64, device="cuda")
    else:
        g, slab, nbytes = None, None, torch.zeros(1, dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(nbytes, src=0)
    d_slab = torch.empty(int(nbytes.item()), dtype=torch.uint8, device="cuda")
    if rank == 0:
        d_slab.copy_(slab)
    t_bcast = 0.0
    if world > 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.broadcast(d_slab, src=0)
        torch.cuda.synchronize()
        t_bcast = time.perf_counter() - t0
    ctx.adopt_device_slab(d_slab.data_ptr(), d_slab.numel())

    # ---- reads: each rank its own shard (weak scaling), pinned on the host + resident copy in HBM
    n_batches = 2
    batches = []
    for b in range(n_batches):
        pool, off = make_reads(genome, args.bases_per_step, seed=7 + 101 * rank + b)
        h_pool = torch.from_numpy(pool).pin_memory()
        h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
        batches.append({"h_pool": h_pool, "h_off": h_off, "d_pool": h_pool.cuda(), "d_off": h_off.cuda(),
                        "off_np": off, "n": len(off) - 1, "bases": int(off[-1]),
                        "lookups": nominal_lookups(np.diff(off.astype(np.int64)))})
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    INEX = rb.api.SEARCH_INS | rb.api.SEARCH_DEL | rb.api.SEARCH_SUBST | rb.api.SEARCH_OR_EXCL

    def resident_step(bt):
        """K1 over a batch resident in HBM: exact sweep + the 9k+1 one-edit sweeps -> labelled hits in HBM"""
        pr, nh, ms = C.c_uint64(), C.c_uint64(), C.c_float()
        out = {}
        for name, flags in (("exact", rb.api.SEARCH_EXACT), ("inexact", INEX)):
            rc = L.rtk_k1_sweep_device(ctx.h, bt["n"], C.c_void_p(bt["d_pool"].data_ptr()), C.c_void_p(bt["d_off"].data_ptr()),
                                       bt["off_np"].ctypes.data_as(C.POINTER(C.c_uint64)), flags, C.byref(pr), C.byref(nh), C.byref(ms))
            if rc != 0:
                raise RuntimeError(L.rtk_last_error().decode())
            out[name] = (pr.value, nh.value, ms.value)
        return out

    opt = rb.default_opt(1)
    seeds = rb.api.RtkSeeds()

    def e2e_step(bt):
        st = (C.c_uint64 * 8)()
        rc = L.rtk_get_seeds(ctx.h, C.byref(opt), 1, bt["n"], C.cast(bt["h_pool"].data_ptr(), C.c_char_p),
                             C.cast(bt["h_off"].data_ptr(), C.POINTER(C.c_uint64)), C.byref(seeds), st)
        if rc != 0:
            raise RuntimeError(L.rtk_last_error().decode())
        n_anchor = int(seeds.solid_off[bt["n"]]) + int(seeds.weak_off[bt["n"]])
        L.rtk_seeds_free(C.byref(seeds))
        return list(st), n_anchor

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for w in range(args.warmup):
        resident_step(batches[w % n_batches])
    # ---- timed: EXACTLY K steps, L2 flushed between iterations, device time = max over ranks
    step_s, kernel_ms, probes, raw_hits = [], [], 0, 0
    barrier()
    with ClockSampler(local_rank) as clk:
        for s in range(args.steps):
            bt = batches[s % n_batches]
            flush.fill_(s & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = resident_step(bt)
            torch.cuda.synchronize()
            step_s.append(time.perf_counter() - t0)
            kernel_ms.append(r["inexact"][2])
            probes += r["exact"][0] + r["inexact"][0]
            raw_hits += r["exact"][1] + r["inexact"][1]
            last = r
    barrier()
    clocks = clk.summary()
    my_t = sum(step_s)
    my_l = sum(batches[s % n_batches]["lookups"] for s in range(args.steps))

    # ---- e2e through the C ABI from pinned host buffers
    for w in range(min(args.warmup, 2)):
        e2e_step(batches[w % n_batches])
    barrier()
    e2e_s, h2d, d2h = [], 0, 0
    for s in range(args.steps):
        bt = batches[s % n_batches]
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st, n_anchor = e2e_step(bt)
        e2e_s.append(time.perf_counter() - t0)
        h2d += 2 * bt["bases"] + 2 * 8 * (bt["n"] + 1)   # reads + the masked copy for the one-edit sweep, offsets twice
        d2h += 16 * st[1]                                  # labelled raw hits
    barrier()
    my_e2e_t = sum(e2e_s)

    if dist is not None:
        tt = torch.tensor([my_t, my_e2e_t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ll = torch.tensor([float(my_l)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        tot_t, tot_e2e_t, tot_l = float(tt[0]), float(tt[1]), float(ll[0])
    else:
        tot_t, tot_e2e_t, tot_l = my_t, my_e2e_t, float(my_l)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        # dominant kernel = the one-edit sweep; algorithmic bytes per SURVEY.md §8d: 16 B per miss, 24 B per hit
        pr_i, nh_i, _ = last["inexact"]
        kms = float(np.mean(kernel_ms))
        alg_bytes = 16.0 * (pr_i - nh_i) + 24.0 * nh_i
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_inexact_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {"metric": "kmer_lookups_per_s", "value": tot_l / tot_t, "unit": "lookups/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": dict(workload_config(args, args.bases_per_step), graph_broadcast_s=t_bcast,
                               slab_bytes=int(d_slab.numel()), kernel_probes_per_step=int(pr_i),
                               nominal_lookups_per_step=int(batches[0]["lookups"]),
                               seeded_bases_per_s=float(world * args.steps * args.bases_per_step / tot_t)),
                "clocks": clocks,
                "e2e": {"value": tot_l / tot_e2e_t, "unit": "lookups/s", "h2d_bytes_per_step": int(h2d / args.steps),
                        "d2h_bytes_per_step": int(d2h / args.steps), "ms_per_step": 1e3 * tot_e2e_t / args.steps,
                        "bases_per_s": float(world * args.steps * args.bases_per_step / tot_e2e_t)},
                "gpu_launches": 2 * args.steps,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "rtk_k1_inexact_kernel<u64>", "kernel_ms": kms,
                             "peak_source": peak_src,
                             "note": "graph of this config (E. coli: 47 MB index) is L2-resident, so DRAM traffic is far below the algorithmic bytes"}}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, genome)
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args, genome):
    """reference getSeeds on the host cores (unmodified objects when built here, else the oracle port)"""
    import refseams as R
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    pool, off = make_reads(genome, args.cpu_baseline_bases, seed=4242)
    reads = [pool[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    looks = nominal_lookups(np.diff(off.astype(np.int64)))
    if R.available():
        tmp = tempfile.mkdtemp(prefix="rtk_bench_")
        fa = os.path.join(tmp, "graph.fasta")
        with open(fa, "w") as f:
            for i, u in enumerate(genome_unitigs(genome)):
                f.write(">%d\n%s\n" % (i, u.decode()))
        g = R.RefGraph(fa, "", K, threads=min(cores, 16))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(lambda s: g.get_seeds(s, "", False), reads))
        dt = time.perf_counter() - t0
        return {"value": looks / dt, "unit": "lookups/s", "cores": cores, "kind": "reference",
                "sample": "%d bases (%d reads) of the same workload through the reference getSeeds, %d threads, %.1f s" %
                          (int(off[-1]), len(reads), cores, dt)}
    from common import OracleGraph
    og = OracleGraph([u.decode() for u in genome_unitigs(genome)], K)
    reads = reads[:8]
    looks = nominal_lookups([len(r) for r in reads])
    t0 = time.perf_counter()
    for s in reads:
        og.search(s, True, False, False, False, False)
        og.search(s, False, True, True, True, True)
    dt = time.perf_counter() - t0
    return {"value": looks / dt, "unit": "lookups/s", "cores": 1, "kind": "port",
            "sample": "%d reads through oracle/rtk_oracle.cpp searchSequence (exact + inexact), 1 thread, %.1f s" % (len(reads), dt)}


if __name__ == "__main__":
    main()
