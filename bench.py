#!/usr/bin/env python3
"""bench.py — headline benchmark: corrected long-read bases per second (BASELINE.json `metric`), pass 1, k = 31.

Workload = BASELINE.json configs[1] ("E. coli: 30x Illumina + 30x ONT R9.4, k=31 pass-1 on 1xB200"), synthetic
because there is no network: bench_data/F3 holds the index the UNMODIFIED reference built (`Ratatosk index -1`) from
the seeded F3 recipe of tests/golden/make_fixtures.py (4.64 Mbp diploid genome with repeat families, tandem repeats and
homopolymers, 30x PE150 short reads) and the genome itself; the long reads are drawn here from that genome with numpy
(ONT-like: lognormal(9.0, 0.6) >= 1 kb, 3 % substitutions / 2.5 % insertions / 4.5 % deletions, Q5-29), a different
seeded batch per rank and step.  A step = one batch of reads through the per-read body of the reference's search()
(getSeeds + correctSequence, src/Ratatosk.cpp:808-867); "corrected bases" = input read bases of the batch.

  value : rtk_correct_batch_resident — reads already resident in HBM for the k-mer sweep when the clock starts
  e2e   : rtk_correct_batch through the C ABI from pinned HOST buffers; the corrected reads come back in host memory
          (every H2D / D2H copy the library makes is inside the timed region and counted by the library)
  roofline : the k-mer lookup sweep (K1, the kernel the metric's second half names) against the measured HBM peak;
          `kernels` lists the GPU time of every kernel family of the step so its share can be checked against
          profiles/
  --impl reference : the reference's own getSeeds + correctSequence (oracle/_ref/libref_seams.so = the unmodified
          reference objects behind extern "C" probes) on all host threads, bounded sample of the same workload
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# the library's service contexts own ~20 CUDA streams: give them their own hardware work queues (default 8); must be set
# before CUDA initialises in this process (torch does that first here)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout (one JSON line is the contract); an explicit setting wins

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 31
F3 = os.path.join(ROOT, "bench_data", "F3")
F3_FASTA = os.path.join(F3, "index.k31.fasta.gz")
F3_RTSK = os.path.join(F3, "index.k31.rtsk")
LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


# ----------------------------------------------------------------------------- synthetic workload
def load_haplotypes():
    z = np.load(os.path.join(F3, "genome.npz"))
    n = int(z["n"])
    p = z["hap0_2bit"]
    h0 = np.empty(len(p) * 4, dtype=np.uint8)
    h0[0::4], h0[1::4], h0[2::4], h0[3::4] = p >> 6, (p >> 4) & 3, (p >> 2) & 3, p & 3
    h0 = h0[:n]
    h1 = h0.copy()
    h1[z["snp_pos"]] = z["snp_base"]
    return [h0, h1]


def make_reads(haps, total_bases, seed):
    """ONT-like reads: lognormal(9.0, 0.6) >= 1 kb, 3 % sub / 2.5 % ins / 4.5 % del, random haplotype and strand,
    qualities uniform in Q5..Q29 (the F3 recipe of tests/golden/make_fixtures.py, vectorised)"""
    rng = np.random.default_rng(seed)
    pool, offs, tot = [], [0], 0
    glen = len(haps[0])
    while tot < total_bases:
        ln = int(min(max(1000, rng.lognormal(9.0, 0.6)), 200_000))
        h = haps[int(rng.integers(0, len(haps)))]
        p = int(rng.integers(0, glen - ln))
        s = h[p:p + ln]
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
    seq = np.concatenate(pool)
    qual = (33 + rng.integers(5, 30, size=len(seq))).astype(np.uint8)
    return seq, qual, np.asarray(offs, dtype=np.uint64)


def workload_config(bases_per_step, extra=None):
    cfg = {"workload": "BASELINE configs[1] shape: E. coli-like 4.64 Mbp genome, index built by the unmodified reference "
                       "from 30x PE150 (bench_data/F3, k=31), pass-1 correction of ONT-like reads lognormal(9.0,0.6) >= 1 kb "
                       "at 10% error, Q5-29",
           "bases_per_step_per_gpu": int(bases_per_step), "k": K, "pass": 1,
           "corrected_bases_definition": "input long-read bases through getSeeds + correctSequence",
           "l2_policy": "256 MiB buffer written between timed iterations (L2 flush)"}
    if extra:
        cfg.update(extra)
    return cfg


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


# ----------------------------------------------------------------------------- reference (CPU) arm
class ReferenceRunner:
    """the reference's own per-read correction (unmodified objects) on all host threads"""

    def __init__(self):
        import refseams as R
        self.R = R
        self.cores = os.cpu_count() or 1
        self.kind = "reference"
        if not R.available():
            raise RuntimeError("oracle/_ref/libref_seams.so not built (needs /root/reference at build time)")
        t0 = time.perf_counter()
        self.g = R.RefGraph(F3_FASTA, F3_RTSK, K, threads=min(self.cores, 16))
        self.load_s = time.perf_counter() - t0

    def run(self, seq, qual, off):
        from concurrent.futures import ThreadPoolExecutor
        reads = [(seq[int(off[i]):int(off[i + 1])].tobytes().decode(), qual[int(off[i]):int(off[i + 1])].tobytes().decode())
                 for i in range(len(off) - 1)]
        reads.sort(key=lambda r: -len(r[0]))   # longest first: the pool drains evenly
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=self.cores) as ex:
            out = list(ex.map(lambda r: self.g.correct_read(r[0], r[1], False), reads))
        dt = time.perf_counter() - t0
        assert len(out) == len(reads)
        return dt


def run_reference(args, rank):
    if rank != 0:
        return
    try:
        ref = ReferenceRunner()
    except Exception as e:   # cannot happen on a box that received oracle/_ref; kept so the driver always gets a line
        print(json.dumps({"impl": "reference", "unavailable": str(e).splitlines()[0]}))
        return
    haps = load_haplotypes()
    times, bases = [], []
    for step in range(args.warmup + args.steps):
        seq, qual, off = make_reads(haps, args.ref_sample_bases, seed=1000 + step)
        dt = ref.run(seq, qual, off)
        if step >= args.warmup:
            times.append(dt)
            bases.append(int(off[-1]))
    val = sum(bases) / sum(times)
    sample = "%d bases of reads per step through the reference getSeeds + correctSequence (libref_seams.so), %d threads" % (
        args.ref_sample_bases, ref.cores)
    line = {"metric": "corrected_bases_per_s", "value": val, "unit": "bases/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "impl": "reference",
            "config": workload_config(args.ref_sample_bases, {"graph_load_s": ref.load_s}),
            "cpu_baseline": {"value": val, "unit": "bases/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(args, haps):
    ref = ReferenceRunner()
    seq, qual, off = make_reads(haps, args.cpu_baseline_bases, seed=4242)
    dt = ref.run(seq, qual, off)
    return {"value": int(off[-1]) / dt, "unit": "bases/s", "cores": ref.cores, "kind": ref.kind,
            "sample": "%d bases (%d reads) of the same workload through the reference getSeeds + correctSequence "
                      "(libref_seams.so), %d threads, %.1f s" % (int(off[-1]), len(off) - 1, ref.cores, dt)}


# ----------------------------------------------------------------------------- product arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases-per-step", type=int, default=64 << 20)
    ap.add_argument("--ref-sample-bases", type=int, default=3_000_000)
    ap.add_argument("--cpu-baseline-bases", type=int, default=8_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--check", action="store_true", help="compare one small batch with the reference's output (needs oracle/_ref)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # exactly ONE line on stdout: NCCL / driver banners written to file descriptor 1 by native code go to stderr instead;
    # the descriptor is restored just before the JSON line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

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

    # ---- graph: loaded + flattened on rank 0, one H2D, ONE NCCL broadcast of the slab, adopted in place by every rank
    ctx = rb.Context(local_rank)
    t0 = time.perf_counter()
    if rank == 0:
        g = rb.Graph.load(F3_FASTA, F3_RTSK, K)
        slab = torch.from_numpy(g.slab())
        nbytes = torch.tensor([slab.numel()], dtype=torch.int64, device="cuda")
    else:
        g, slab, nbytes = None, None, torch.zeros(1, dtype=torch.int64, device="cuda")
    t_load = time.perf_counter() - t0
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

    # ---- reads: each rank its own shard (weak scaling), pinned on the host + a resident copy of the bases in HBM
    haps = load_haplotypes()
    n_batches = 2
    batches = []
    for b in range(n_batches):
        seq, qual, off = make_reads(haps, args.bases_per_step, seed=7 + 101 * rank + b)
        h_seq = torch.from_numpy(seq).pin_memory()
        h_qual = torch.from_numpy(qual).pin_memory()
        h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
        batches.append({"h_seq": h_seq, "h_qual": h_qual, "h_off": h_off, "d_seq": h_seq.cuda(), "d_off": h_off.cuda(),
                        "n": len(off) - 1, "bases": int(off[-1])})
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    opt = rb.default_opt(1)
    u64p = C.POINTER(C.c_uint64)

    def correct_step(bt, resident):
        """one batch through getSeeds + correctSequence; returns (library stats, corrected bases out)"""
        os_, oq_, oo = C.c_void_p(), C.c_void_p(), u64p()
        st = (C.c_uint64 * 24)()
        seq_p = C.cast(bt["h_seq"].data_ptr(), C.c_char_p)
        qual_p = C.cast(bt["h_qual"].data_ptr(), C.c_char_p)
        off_p = C.cast(bt["h_off"].data_ptr(), u64p)
        if resident:
            rc = L.rtk_correct_batch_resident(ctx.h, C.byref(opt), 1, bt["n"], seq_p, off_p, C.c_void_p(bt["d_seq"].data_ptr()),
                                              C.c_void_p(bt["d_off"].data_ptr()), qual_p, off_p, C.byref(os_), C.byref(oq_),
                                              C.byref(oo), st)
        else:
            rc = L.rtk_correct_batch(ctx.h, C.byref(opt), 1, bt["n"], seq_p, off_p, qual_p, off_p, C.byref(os_), C.byref(oq_),
                                     C.byref(oo), st)
        if rc != 0:
            raise RuntimeError(L.rtk_last_error().decode())
        out_bases = int(oo[bt["n"]])
        res = (C.string_at(os_, min(out_bases, 64)), out_bases)   # the corrected reads are host memory already
        L.rtk_free(os_)
        L.rtk_free(oq_)
        L.rtk_free(C.cast(oo, C.c_void_p))
        return list(st), res

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.check and rank == 0:
        check_against_reference(ctx, rb, haps)

    def timed_loop(resident, warmup):
        for w in range(warmup):
            correct_step(batches[w % n_batches], resident)
        barrier()
        step_s, stats, ev_ms, seeds_ms = [], np.zeros(24, dtype=np.float64), 0.0, []
        with ClockSampler(local_rank) as clk:
            for s in range(args.steps):
                bt = batches[s % n_batches]
                flush.fill_(s & 0xFF)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                t0 = time.perf_counter()
                st, _ = correct_step(bt, resident)
                e1.record()
                torch.cuda.synchronize()
                step_s.append(time.perf_counter() - t0)
                ev_ms += e0.elapsed_time(e1)
                stats += np.asarray(st, dtype=np.float64)
                seeds_ms.append(round(st[10] / 1e6, 1))
        barrier()
        return sum(step_s), stats, ev_ms, clk.summary(), {"step": [round(1e3 * x, 1) for x in step_s], "getSeeds_stage": seeds_ms}

    my_t, st_res, ev_ms, clocks, step_ms_res = timed_loop(True, args.warmup)
    my_e2e_t, st_e2e, ev_e2e_ms, _, step_ms_e2e = timed_loop(False, args.warmup)
    my_b = sum(batches[s % n_batches]["bases"] for s in range(args.steps))

    if dist is not None:
        tt = torch.tensor([my_t, my_e2e_t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        bb = torch.tensor([float(my_b)], dtype=torch.float64, device="cuda")
        dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        tot_t, tot_e2e_t, tot_b = float(tt[0]), float(tt[1]), float(bb[0])
    else:
        tot_t, tot_e2e_t, tot_b = my_t, my_e2e_t, float(my_b)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        K_ = args.steps
        probes, raw_hits, k1_ns = st_res[0] / K_, st_res[1] / K_, st_res[2] / K_
        # K1 algorithmic bytes: 16 B per probe rejected at the index, 24 B per hit (k = 31)
        alg_bytes = 16.0 * (probes - raw_hits) + 24.0 * raw_hits
        achieved = alg_bytes / (k1_ns * 1e-9) / 1e9 if k1_ns else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_inexact_traffic.json")))["dram_bytes_per_probe"] * probes
        except Exception:
            pass
        fam = {"K1 lookup sweep (rtk_k1_exact/inexact_kernel)": st_res[2] / K_ / 1e6,
               "K4 edit distance (rtk_myers_kernel<G>, selectors / prefixes / repeats)": st_res[7] / K_ / 1e6,
               "K5 alignment paths (rtk_myers_fill_kernel + rtk_traceback_kernel)": st_res[8] / K_ / 1e6,
               "K2/K3+K4 graph bursts (rtk_dfs_kernel + leaf rtk_myers_kernel)": st_res[9] / K_ / 1e6}
        tot_fam = sum(fam.values()) or 1.0
        bases_per_step = tot_b / K_ / world
        line = {"metric": "corrected_bases_per_s", "value": tot_b / tot_t, "unit": "bases/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": workload_config(args.bases_per_step, {
                    "reads_per_step_per_gpu": batches[0]["n"], "graph_load_flatten_s": t_load, "graph_broadcast_s": t_bcast,
                    "slab_bytes": int(d_slab.numel()), "host_cores": os.cpu_count(),
                    "value_arm": "reads resident in HBM for the k-mer sweep; the region logic runs on host fibers and ships "
                                 "candidate strings to the K2-K5 services in both arms",
                    "kmer_lookups_per_s_in_kernel": probes / (k1_ns * 1e-9) if k1_ns else None,
                    "kmer_lookups_per_step": probes,
                    "ms_per_step_cuda_events": ev_ms / args.steps, "step_ms_rank0": step_ms_res,
                    "stage_ms_per_step": {"getSeeds (K1 + host anchor logic)": st_res[10] / K_ / 1e6,
                                          "regions (host fibers + K2-K5 services)": st_res[11] / K_ / 1e6},
                    "gpu_service_calls_per_step": st_res[5] / K_, "gpu_requests_per_step": st_res[6] / K_}),
                "clocks": clocks,
                "e2e": {"value": tot_b / tot_e2e_t, "unit": "bases/s", "h2d_bytes_per_step": int(st_e2e[12] / K_),
                        "d2h_bytes_per_step": int(st_e2e[13] / K_), "ms_per_step": 1e3 * tot_e2e_t / args.steps, "step_ms_rank0": step_ms_e2e,
                        "note": "byte counts are the library's own tally of every cudaMemcpyAsync it issued in the step; the "
                                "corrected reads are assembled in host memory (the D2H traffic is the services' answers)"},
                "gpu_launches": int(st_res[14]),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "rtk_k1_inexact_kernel<u64> (+ the exact sweep, ~3 % of its time)",
                             "kernel_ms_per_step": k1_ns / 1e6, "peak_source": peak_src,
                             "algorithmic_bytes_per_step": alg_bytes,
                             "share_of_step_gpu_time": (st_res[2] / K_ / 1e6) / tot_fam,
                             "note": "the metric names k-mer lookups vs the HBM roofline, so K1 is the kernel reported here; by GPU "
                                     "time the step is dominated by the bit-parallel Myers kernels (latency-bound launches of a few CTAs, HBM "
                                     "traffic negligible) - see `kernels` and profiles/.  traffic = measured DRAM bytes per probe (ncu, "
                                     "profiles/k1_inexact_traffic.json) x the probes of this launch"},
                "kernels": {"gpu_ms_per_step": fam, "share": {k: v / tot_fam for k, v in fam.items()},
                            "note": "CUDA-event time on each service's launching stream; the services run concurrently, so the "
                                    "sum can exceed the step's wall time"}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, haps)
            except Exception as e:
                line["cpu_baseline"] = {"unavailable": str(e).splitlines()[0]}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def check_against_reference(ctx, rb, haps):
    """checker leg (not timed): a small fresh batch must come out byte-identical to the reference's own correction.  The
    reference breaks ties between colour sets of equal cardinality by pointer hash (tests/ref_worker.py), so its output is
    taken from three processes and a read may match any of them."""
    import pickle
    import tempfile
    seq, qual, off = make_reads(haps, 300_000, seed=99)
    reads = [(seq[int(off[i]):int(off[i + 1])].tobytes().decode(), qual[int(off[i]):int(off[i + 1])].tobytes().decode())
             for i in range(len(off) - 1)]
    tmp = tempfile.mkdtemp(prefix="rtk_check_")
    pickle.dump(reads, open(os.path.join(tmp, "reads.pkl"), "wb"))
    variants = []
    for i in range(3):
        dst = os.path.join(tmp, "out%d.pkl" % i)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "ref_worker.py"), F3_FASTA, F3_RTSK, str(K),
                               os.path.join(tmp, "reads.pkl"), dst])
        variants.append(pickle.load(open(dst, "rb")))
    ours = ctx.correct([r[0] for r in reads], [r[1] for r in reads])
    bad = [i for i in range(len(reads)) if all(ours[i] != v[i] for v in variants)]
    sys.stderr.write("[check] %d reads (%d bases): %s\n" % (len(reads), int(off[-1]), "identical to the reference" if not bad else "DIFFER: %r" % bad[:10]))
    if bad:
        raise SystemExit("bench.py --check: output differs from the reference")


if __name__ == "__main__":
    main()
