"""N>1 host logic on CPU: world_size-2 gloo run of the ticket dealer + ordered merge (ratatosk_b200/shard.py)."""
import os
import random
import subprocess
import sys
import textwrap

from ratatosk_b200 import shard

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_tickets_cover_reads_in_order():
    rng = random.Random(5)
    lens = [rng.randint(500, 400000) for _ in range(300)]
    tk = shard.make_tickets(lens)
    assert tk[0][0] == 0 and tk[-1][1] == len(lens)
    for (a, b), (c, d) in zip(tk[:-1], tk[1:]):
        assert b == c and a < b
    for a, b in tk[:-1]:
        assert sum(lens[a:b]) >= shard.BUFFER_SZ and sum(lens[a:b - 1]) < shard.BUFFER_SZ
    dealt = shard.deal(tk, 3)
    assert sorted(t for r in dealt for t in r) == list(range(len(tk)))


def test_two_ranks_gloo_ordered_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, random
        sys.path.insert(0, %r)
        import torch.distributed as dist
        from ratatosk_b200 import shard
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        rng = random.Random(11)
        reads = ["".join(rng.choice("ACGT") for _ in range(rng.randint(200, 3000))) for _ in range(200)]
        tickets = shard.make_tickets([len(r) for r in reads], buffer_sz=20000)
        mine = shard.deal(tickets, world)[rank]
        # stand-in for the per-read hot path: any pure per-read function
        blocks = {t: [r[::-1].lower() for r in reads[tickets[t][0]:tickets[t][1]]] for t in mine}
        gathered = [None] * world
        dist.gather_object(blocks, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            out = shard.merge_ordered(gathered)
            assert out == [r[::-1].lower() for r in reads], "order not restored"
            assert len(tickets) > 4 and all(len(b) > 0 for b in gathered)
            print("MERGE_OK", len(tickets))
        dist.barrier()
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MERGE_OK" in r.stdout


def test_two_ranks_gloo_annotation_by_unitig_ranges(tmp_path):
    """The index steps shard like the reads do (replicas of the graph, no collective in the data path): two ranks, each runs
    rtk_detect_snps_range / rtk_detect_short_cycles_range on its half of the unitigs of the F2 k = 31 graph (kernel sources on the CPU
    simulator); the gathered halves concatenate to what the reference stored in the index for every unitig."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        sys.path.insert(0, os.path.join(%r, "tests"))
        import torch.distributed as dist
        import ratatosk_b200 as rb
        from common import GOLDEN, ensure_built
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        lib = os.path.join(%r, "tests", "hostsim", "_build", "librtk_hostsim.so")
        d = os.path.join(GOLDEN, "F2")
        g = rb.Graph.load(os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk"), 31, lib=lib)
        ctx = rb.Context(0, lib=lib)
        ctx.upload(g)
        n = g.info()["n_unitigs"]
        per = (n + world - 1) // world
        first = rank * per
        off, ids = ctx.detect_snps(first=first, count=per)
        flags, coff, pool = ctx.detect_short_cycles(first=first, count=per)
        mine = {}
        for u in range(n):
            a = [int(x) for x in ids[int(off[u]):int(off[u + 1])]]
            b = pool[int(coff[u]):int(coff[u + 1])]
            if a or b or flags[u]:
                assert first <= u < first + per, "annotation outside the rank's range"
                mine[u] = (a, int(flags[u]), b)
        gathered = [None] * world
        dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            merged = {}
            for part in gathered:
                assert not (set(part) & set(merged))
                merged.update(part)
            want = {}
            for u in range(n):
                a, b = g.unitig_annotations(u)
                f = (g.unitig_words(u)[1] >> 8) & 1
                if a or b or f:
                    want[u] = (a, f, b)
            assert merged == want and all(len(p) > 100 for p in gathered)
            print("ANNOTATION_OK", len(want))
        dist.barrier()
        dist.destroy_process_group()
    """ % (ROOT, ROOT, ROOT)))
    from common import ensure_built
    ensure_built()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-2000:])
    assert "ANNOTATION_OK" in r.stdout
