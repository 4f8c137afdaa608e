#!/bin/bash
# F4 (chr20-scale, BASELINE.json configs[2]) bench / parity data: genome + reads by scripts/make_f4.py, index and golden
# outputs by the UNMODIFIED reference binary (oracle/_ref/Ratatosk).  TEST / BENCH DATA generator; run once in the build
# container (needs /root/reference for oracle/_ref), outputs under bench_data/F4 (git-ignored: ~200 MB, travels to the GPU box
# with the working tree).  Usage: scripts/make_f4.sh [work_dir] [threads] [genome_len]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
WORK="${1:-/tmp/f4}"; THREADS="${2:-6}"; GLEN="${3:-64000000}"
REF="$ROOT/oracle/_ref/Ratatosk"
OUT="$ROOT/bench_data/F4"
mkdir -p "$WORK" "$OUT"
cd "$WORK"
if [ ! -f i.index.k31.rtsk ]; then
  [ -f sr.fastq ] || python "$ROOT/scripts/make_f4.py" --out-dir "$WORK" --genome-len "$GLEN" --long-bases 3
  "$REF" index -1 -v -c "$THREADS" -s sr.fastq -l lr_sample.fastq -o i > index1.log 2>&1
  rm -f sr.fastq
fi
cp genome.npz "$OUT/genome.npz"
cp i.index.k31.fasta.gz "$OUT/index.k31.fasta.gz"; cp i.index.k31.rtsk "$OUT/index.k31.rtsk"
cp i.index.k63.fasta.gz "$OUT/index.k63.fasta.gz"
# golden sample: the first 200 long reads through the reference's pass 1
head -n 800 lr_sample.fastq > reads200.fastq
"$REF" correct -1 -v -c "$THREADS" -g i.index.k31.fasta.gz -d i.index.k31.rtsk -l reads200.fastq -o g200 > correct1_200.log 2>&1
gzip -c reads200.fastq > "$OUT/reads200.fastq.gz"; gzip -c g200.2.fastq > "$OUT/corrected200_pass1.fastq.gz"
# pass-1 correction of the whole long-read sample (colours of the k = 63 graph), then the pass-2 index and goldens
"$REF" correct -1 -v -c "$THREADS" -g i.index.k31.fasta.gz -d i.index.k31.rtsk -l lr_sample.fastq -o p1 > correct1_all.log 2>&1
"$REF" index -2 -v -c "$THREADS" -g i.index.k63.fasta.gz -l p1.2.fastq -o j > index2.log 2>&1
cp j.index.k63.rtsk "$OUT/index.k63.rtsk"
"$REF" correct -2 -O -v -c "$THREADS" -g i.index.k63.fasta.gz -d j.index.k63.rtsk -l g200.2.fastq -L reads200.fastq -o g200b > correct2_200.log 2>&1
gzip -c g200b.fastq > "$OUT/corrected200_pass2.fastq.gz"
echo "F4 done" > "$WORK/DONE"
