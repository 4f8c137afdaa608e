"""TEST INFRASTRUCTURE: corrects a pickled list of (seq, qual) with the UNMODIFIED reference library in a fresh process.

The reference is not deterministic from process to process: chooseColors (src/Correction.cpp:215-429) orders colour sets of
equal cardinality by POINTER hash, so address-space randomisation can flip a tie (observed on 1 of 154 E. coli-scale reads,
1 run in 6).  Tests that compare against freshly computed reference output therefore collect the outputs of a few processes
and accept any of them for such a read (tests/test_correct.py)."""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refseams as R  # noqa: E402

if __name__ == "__main__":
    fa, rt, k, src, dst = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    reads = pickle.load(open(src, "rb"))
    g = R.RefGraph(fa, rt, k, threads=8)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        out = list(ex.map(lambda r: g.correct_read(r[0], r[1], False), reads))
    pickle.dump(out, open(dst, "wb"))
