#!/usr/bin/env bash
# make_reference_goldens.sh — on ANY box with `go` (>= 1.12) and `samtools` (>= 1.13) on PATH, build the unmodified
# reference (brentp/goleft v0.2.6) and write the outputs that pin this repo's oracle to the real thing:
#
#   tests/golden/ref_t_<W>.{depth,callable}.bed          goleft depth  -w W        on depth/test/t.bam   (.fai mode)
#   tests/golden/ref_t_bed_<W>.{depth,callable}.bed      goleft depth  -w W --bed  on depth/test/t.bam   (BED mode, windows.bed)
#   tests/golden/ref_hla_<W>.{depth,callable}.bed        goleft depth  -w W        on depth/test/hla.bam
#   tests/golden/ref_empty_<W>.{depth,callable}.bed      goleft depth  -w W        on a BAM without reads on the contig
#   tests/golden/ref_indexcov_sample27.{bed,roc,ped}     goleft indexcov           on indexcov/test-data/sample_issue_27.bam(.bai)
#   tests/golden/ref_depthwed.tsv                        goleft depthwed -s 1000   on two of the depth.bed files above
#
# This image has neither toolchain (`go version`, `samtools --version`: command not found), so the files cannot be made
# here; until they exist every report says "bit-exact vs the restated oracle, parity unpinned".  tests/test_reference_goldens.py
# compares BOTH the oracle and the GPU output with these files when they are present and skips, loudly, when they are not.
#
#   usage: tools/make_reference_goldens.sh /path/to/goleft-checkout [outdir]
set -euo pipefail
REF=${1:?path to a brentp/goleft checkout (v0.2.6)}
OUT=${2:-$(cd "$(dirname "$0")/.." && pwd)/tests/golden}
command -v go >/dev/null || { echo "go not found" >&2; exit 2; }
command -v samtools >/dev/null || { echo "samtools not found" >&2; exit 2; }
ver=$(samtools --version | head -1 | awk '{print $2}')
case "$ver" in 0.*|1.[0-9]|1.[0-9].*|1.1[0-2]|1.1[0-2].*) echo "samtools $ver < 1.13: it caps depth at -d and prints zero-depth lines (SURVEY.md 8a D0); use >= 1.13" >&2; exit 2;; esac
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
( cd "$REF" && CGO_ENABLED=0 go build -o "$TMP/goleft" ./cmd/goleft )
G="$TMP/goleft"
T="$REF/depth/test"

# reference FASTA index: goleft depth only reads <ref>.fai (depth.go:125-128); write one from the BAM header
fai_of() { samtools view -H "$1" | awk -F'\t' '$1=="@SQ"{n="";l=0;for(i=2;i<=NF;i++){if($i~/^SN:/)n=substr($i,4);if($i~/^LN:/)l=substr($i,4)}print n"\t"l"\t0\t60\t61"}' > "$2.fai"; }

# the reference's own tests run against test/hg19.fa(.fai) (depth/functional-test.sh:45); fall back to a .fai made from the header
if [[ -e "$T/hg19.fa.fai" ]]; then cp "$T/hg19.fa.fai" "$TMP/t.fa.fai"; else fai_of "$T/t.bam" "$TMP/t.fa"; fi
for W in 100 1000000000 55 60 71 13 2001 250; do      # depth/functional-test.sh:45-70 (+ the default 250)
  "$G" depth -Q 1 --ordered --windowsize $W --prefix "$TMP/t_$W" -r "$TMP/t.fa" "$T/t.bam"
  cp "$TMP/t_$W.depth.bed" "$OUT/ref_t_$W.depth.bed"; cp "$TMP/t_$W.callable.bed" "$OUT/ref_t_$W.callable.bed"
done
if [[ -e "$T/windows.bed" ]]; then
  for W in 10 1000000 50 55 60 71 13 2002; do         # depth/functional-test.sh:73-97
    "$G" depth -Q 1 --ordered --windowsize $W --bed "$T/windows.bed" --prefix "$TMP/tb_$W" -r "$TMP/t.fa" "$T/t.bam"
    cp "$TMP/tb_$W.depth.bed" "$OUT/ref_t_bed_$W.depth.bed"; cp "$TMP/tb_$W.callable.bed" "$OUT/ref_t_bed_$W.callable.bed"
  done
fi
if [[ -e "$T/hla.bam" ]]; then
  fai_of "$T/hla.bam" "$TMP/hla.fa"
  for W in 250 1000; do
    "$G" depth -Q 1 --ordered --windowsize $W --prefix "$TMP/hla_$W" -r "$TMP/hla.fa" "$T/hla.bam"
    cp "$TMP/hla_$W.depth.bed" "$OUT/ref_hla_$W.depth.bed"; cp "$TMP/hla_$W.callable.bed" "$OUT/ref_hla_$W.callable.bed"
  done
fi
# a contig without a single read: .fai lists one the BAM does not cover
printf 'nothing_here\t12345\t0\t60\t61\n' >> "$TMP/t.fa.fai"
"$G" depth -Q 1 --ordered -c nothing_here --windowsize 500 --prefix "$TMP/empty_500" -r "$TMP/t.fa" "$T/t.bam" || true
for f in depth callable; do [[ -e "$TMP/empty_500.nothing_here.$f.bed" ]] && cp "$TMP/empty_500.nothing_here.$f.bed" "$OUT/ref_empty_500.$f.bed"; done

cp "$TMP/t.fa.fai" "$OUT/ref_t.fai"
IC="$REF/indexcov/test-data"
if [[ -e "$IC/sample_issue_27_0001.bam" ]]; then
  "$G" indexcov -d "$TMP/ic27" "$IC/sample_issue_27_0001.bam"
  zcat "$TMP/ic27/ic27-indexcov.bed.gz" > "$OUT/ref_indexcov_sample27.bed"
  cp "$TMP/ic27/ic27-indexcov.roc" "$OUT/ref_indexcov_sample27.roc"
  cut -f1-6,10- "$TMP/ic27/ic27-indexcov.ped" > "$OUT/ref_indexcov_sample27.ped" || cp "$TMP/ic27/ic27-indexcov.ped" "$OUT/ref_indexcov_sample27.ped"
fi
"$G" depthwed -s 1000 "$OUT/ref_t_100.depth.bed" "$OUT/ref_t_55.depth.bed" > "$OUT/ref_depthwed.tsv" || true
samtools --version | head -1 > "$OUT/ref_PROVENANCE.txt"; go version >> "$OUT/ref_PROVENANCE.txt"; (cd "$REF" && git rev-parse HEAD 2>/dev/null || true) >> "$OUT/ref_PROVENANCE.txt"
echo "wrote $(ls "$OUT"/ref_* | wc -l) files to $OUT"
