#!/usr/bin/env bash
# Regenerates the fixtures in this directory from the reference checkout (run in the build
# container only; /root/reference does not exist on the GPU box).  Databases are built by the
# UNMODIFIED reference binary (oracle/_ref/kmer-db, see oracle/build_ref.sh) from the
# reference's own test inputs; the *.csv / a2a* files are the reference's committed golden
# outputs for those inputs (test/virus, test/synth), copied verbatim as test vectors.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
BIN="$HERE/../../oracle/_ref/kmer-db"
cd "$REF"
"$BIN" build test/virus/seqs.list "$HERE/virus.k18.db"
"$BIN" build -f 0.1 test/virus/seqs.list "$HERE/virus.k18.f01.db"
"$BIN" build -k 24 test/virus/seqs.list "$HERE/virus.k24.db"
gzip -9 -f "$HERE/virus.k24.db"      # 65536 mostly empty hashtables: 5 MB -> small
"$BIN" build -multisample-fasta -k 21 test/synth/synth.fa "$HERE/synth.k21.db"
cp test/virus/k18.csv "$HERE/virus.k18.csv"
cp test/virus/k18.sparse.csv "$HERE/virus.k18.sparse.csv"
cp test/virus/k18.frac.csv "$HERE/virus.k18.f01.csv"
cp test/virus/k24.csv "$HERE/virus.k24.csv"
cp test/synth/a2a "$HERE/synth.k21.csv"
cp test/synth/a2a-sparse "$HERE/synth.k21.sparse.csv"
# database of the first 100 virus genomes (the CI's k18.parts.db, .github/workflows/main.yml:73) for new2all
"$BIN" build test/virus/seqs.part1.list "$HERE/virus.k18.part1.db"
# all2all-parts with output filters on the virus genomes split into three parts (60 / 60 / 45 samples, in list order); without
# filters the mode reproduces test/virus/k18.sparse.csv (the reference's own CI check, .github/workflows/self-hosted.yml:357-363)
P3="$(mktemp -d)"
sed -n 1,60p test/virus/seqs.list > "$P3/a.list"; sed -n 61,120p test/virus/seqs.list > "$P3/b.list"; sed -n '121,$p' test/virus/seqs.list > "$P3/c.list"
for x in a b c; do "$BIN" build "$P3/$x.list" "$P3/$x.db"; done
printf '%s\n' "$P3/a.db" "$P3/b.db" "$P3/c.db" > "$P3/db.list"
"$BIN" all2all-parts "$P3/db.list" "$P3/plain.csv" && cmp "$P3/plain.csv" test/virus/k18.sparse.csv
"$BIN" all2all-parts -min jaccard:0.985 -max num-kmers:29700 -min ani:0.9995 "$P3/db.list" "$HERE/virus.k18.parts.filtered.csv"
rm -rf "$P3"
# the reference's test INPUTS (FASTA, lists) and every golden OUTPUT of test/virus, test/synth and the
# amino-acid part of test/protein, verbatim, for the CLI tests (build / new2all / distance / all2all-sp)
tar cJf "$HERE/reference_fixtures.tar.xz" test/virus test/synth test/protein/aa_100x1000.fasta test/protein/*.a2a
# -sample-rows <criterion>:<count> of all2all-sp on the virus database (the best <count> neighbours of every sample);
# all2all-parts with the same option writes the same bytes for the genomes split into parts (checked here, not stored)
for c in jaccard:3 num-kmers:5 ani:2 mash-query:4; do
  "$BIN" all2all-sp -sample-rows "$c" "$HERE/virus.k18.db" "$HERE/virus.k18.sampled.${c/:/_}.csv"
done
"$BIN" all2all-sp -min jaccard:0.99 -max num-kmers:29800 -sample-rows cosine:2 "$HERE/virus.k18.db" "$HERE/virus.k18.sampled.filtered.cosine_2.csv"
P2="$(mktemp -d)"
sed -n 1,100p test/virus/seqs.list > "$P2/a.list"; sed -n '101,$p' test/virus/seqs.list > "$P2/b.list"
for x in a b; do "$BIN" build "$P2/$x.list" "$P2/$x.db"; done
printf '%s\n' "$P2/a.db" "$P2/b.db" > "$P2/db.list"
"$BIN" all2all-parts -sample-rows jaccard:3 "$P2/db.list" "$P2/s.csv" && cmp "$P2/s.csv" "$HERE/virus.k18.sampled.jaccard_3.csv"
"$BIN" all2all-parts -min jaccard:0.99 -max num-kmers:29800 -sample-rows cosine:2 "$P2/db.list" "$P2/f.csv" && cmp "$P2/f.csv" "$HERE/virus.k18.sampled.filtered.cosine_2.csv"
rm -rf "$P2"
# `minhash -f 0.1` on the virus genomes: digests of the 165 <sample>.minhash files the reference writes (the CI then builds
# from them and expects test/virus/k18.frac.csv, .github/workflows/main.yml:143-148)
M="$(mktemp -d)"; cp -r test/virus "$M/virus"
( cd "$M" && sed 's#\./test/virus/#./virus/#' virus/seqs.list > seqs.list && "$BIN" minhash -f 0.1 seqs.list \
  && sha256sum virus/data/*.minhash | sed 's#  virus/data/#  #' > "$HERE/virus.minhash.f01.sha256" )
rm -rf "$M"
