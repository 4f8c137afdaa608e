#!/usr/bin/env python
"""DEVELOPMENT AID (needs oracle/_ref, i.e. the build container): wall-clock of the reference's detectSNPs and detectShortCycles
steps inside `Ratatosk index -1 -v`, read off the time stamps of its progress lines (src/Ratatosk.cpp:1124-1140), as the CPU
baseline of SURVEY 8(f)3.  Usage: time_ref_annotate.py <work_dir> <genome_len> <threads>.  Leaves the index in <work_dir>/i.*"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work, glen, threads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
os.makedirs(work, exist_ok=True)
if not os.path.exists(os.path.join(work, "sr.fastq")):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "make_f4.py"), "--out-dir", work, "--genome-len", str(glen),
                           "--long-bases", "3"])
p = subprocess.Popen([os.path.join(ROOT, "oracle", "_ref", "Ratatosk"), "index", "-1", "-v", "-c", str(threads), "-s", "sr.fastq",
                      "-l", "lr_sample.fastq", "-o", "i"], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
t0 = time.time()
marks = {}
for line in p.stdout:
    t = time.time() - t0
    for key, tag in (("Adding colors and coverage", "addCoverage"), ("Adding SNPs candidates", "detectSNPs"),
                     ("Adding micro/mini-satellites", "detectShortCycles"), ("Writing index", "write")):
        if key in line and (tag not in marks or tag == "write"):   # "Writing index" is printed for the graph too: keep the last
            marks[tag] = t
p.wait()
out = {"genome_len": glen, "threads": threads, "total_s": time.time() - t0, "marks_s": marks}
if "detectSNPs" in marks and "detectShortCycles" in marks and "write" in marks:
    out["detectSNPs_s"] = marks["detectShortCycles"] - marks["detectSNPs"]
    out["detectShortCycles_s"] = marks["write"] - marks["detectShortCycles"]
    if "addCoverage" in marks:
        out["addCoverage_s"] = marks["detectSNPs"] - marks["addCoverage"]
print(json.dumps(out))
