#!/bin/bash
# TEST INFRASTRUCTURE — regenerates the pass-2 golden files of tests/golden/F1 and F2 with the UNMODIFIED reference
# binary (oracle/_ref/Ratatosk, built by oracle/Makefile from /root/reference).  Run in the build container only.
#
#   index.k63.fasta.gz            k2 = 63 graph of the short reads, written by `Ratatosk index -1` next to the k31 index
#   index.k63.rtsk                its colouring by the pass-1 corrected long reads, `Ratatosk index -2`
#   corrected_pass2_nophasing.fastq.gz   `Ratatosk correct -2 -c 1`: the single-thread branch of search() runs
#                                 getSeeds + correctSequence only (src/Ratatosk.cpp:670-703) = the seam rtk_correct_batch(pass 2) replaces
#   corrected_pass2.fastq.gz      `Ratatosk correct -2 -O -c 8`: the multi-thread branch adds phasing() (src/Ratatosk.cpp:832)
#
# Inputs: the committed pass-1 corrected reads (corrected_pass1.fastq.gz) and their raw reads (reads.fastq.gz); the short
# reads are regenerated from the seeded recipe (make_fixtures.py).  Index building is not reproducible run to run
# (std::random_device in src/Graph.cpp), which is why the index files are committed.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
R=$HERE/../../oracle/_ref/Ratatosk
W=${1:-/tmp/rtk_golden_pass2}
mkdir -p $W
for F in F1 F2; do
  python $HERE/make_fixtures.py --recipe $F --out $W/$F > $W/$F.gen.log
  $R index -1 -c 8 -s $W/$F.sr.fastq -l $W/$F.lr.fastq -o $W/$F.i > $W/$F.index1.log 2>&1
  zcat $HERE/$F/corrected_pass1.fastq.gz > $W/$F.p1.fastq
  zcat $HERE/$F/reads.fastq.gz > $W/$F.raw.fastq
  $R index -2 -c 8 -g $W/$F.i.index.k63.fasta.gz -l $W/$F.p1.fastq -o $W/$F.j > $W/$F.index2.log 2>&1
  $R correct -2 -O -c 8 -g $W/$F.i.index.k63.fasta.gz -d $W/$F.j.index.k63.rtsk -l $W/$F.p1.fastq -L $W/$F.raw.fastq -o $W/$F.o > $W/$F.c8.log 2>&1
  $R correct -2 -O -c 1 -g $W/$F.i.index.k63.fasta.gz -d $W/$F.j.index.k63.rtsk -l $W/$F.p1.fastq -L $W/$F.raw.fastq -o $W/$F.o1 > $W/$F.c1.log 2>&1
  cp $W/$F.i.index.k63.fasta.gz $HERE/$F/index.k63.fasta.gz
  cp $W/$F.j.index.k63.rtsk $HERE/$F/index.k63.rtsk
  gzip -9 -c $W/$F.o.fastq > $HERE/$F/corrected_pass2.fastq.gz
  gzip -9 -c $W/$F.o1.fastq > $HERE/$F/corrected_pass2_nophasing.fastq.gz
done
