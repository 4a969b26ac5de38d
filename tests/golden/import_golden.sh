#!/bin/bash
# Provenance of tests/golden/: data fixtures (no source code) copied verbatim from the
# reference's own test suites so the parity tests can run where /root/reference is absent.
#   htscodecs/tests/dat/{q4,q8,q40+dir,qvar,qsimd,u32}       raw inputs
#   htscodecs/tests/dat/{r4x16,r4x8,arith,fqzcomp}/*         reference-compressed golden streams
#   htscodecs/tests/names/*.names + names/tok3/*             name sets + tok3 golden streams
#   test/bgziptest.txt(.gz,.gz.gzi), test/bgzf_boundaries/*  BGZF fixtures
set -e
REF=${REF:-/root/reference}
D=$(cd "$(dirname "$0")" && pwd)
mkdir -p $D/htscodecs $D/htslib
cp -r $REF/htscodecs/tests/dat $D/htscodecs/
cp -r $REF/htscodecs/tests/names $D/htscodecs/
cp $REF/test/bgziptest.txt $REF/test/bgziptest.txt.gz $REF/test/bgziptest.txt.gz.gzi $D/htslib/
cp -r $REF/test/bgzf_boundaries $D/htslib/
cp $REF/test/ce#1.sam $REF/test/ce#1000.sam $REF/test/ce.fa $REF/test/ce.fa.fai $D/htslib/ 2>/dev/null || true
cp $REF/test/range.bam $REF/test/colons.bam $D/htslib/ 2>/dev/null || true
cp $REF/test/range.cram $REF/test/*_java.cram $D/htslib/ 2>/dev/null || true   # foreign-writer CRAM 3.0 files
cp $REF/test/xx.fa $REF/test/xx.fa.fai $REF/test/auxf.fa $REF/test/auxf.fa.fai $D/htslib/ 2>/dev/null || true   # their references
chmod -R u+w $D
