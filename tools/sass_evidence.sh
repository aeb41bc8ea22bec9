#!/bin/bash
# SASS / ptxas evidence for profiles/: per kernel the counts of the mnemonics that prove the bulk-copy engine (UBLKCP),
# mbarriers (SYNCS), warp reductions (REDUX), shared / global atomics (ATOMS / ATOMG / RED), plus ptxas' register / spill table.
# usage: tools/sass_evidence.sh > profiles/rNN_sass.txt   (run after `make -C deflate_compression_b200/csrc`)
SO=deflate_compression_b200/libdeflate_b200.so
echo "# cuobjdump -lelf $SO"; cuobjdump -lelf $SO
echo; echo "# ptxas -v (deflate_compression_b200/csrc/build/ptxas.log): registers, spills, shared memory per kernel"
grep -E "Compiling entry|Used [0-9]+ registers|spill" deflate_compression_b200/csrc/build/ptxas.log | sed -e 's/^ptxas info    : //' |
  awk '/Compiling entry/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name); next} /spill/ {sp=$0; next} /Used/ {print name": "$0" | "sp}'
echo; echo "# cuobjdump -sass: mnemonic counts per kernel"
cuobjdump -sass $SO | awk '
  /Function :/ {f=$3}
  /UBLKCP/ {c[f,"UBLKCP"]++} /SYNCS/ {c[f,"SYNCS"]++} /REDUX/ {c[f,"REDUX"]++} /ATOMS/ {c[f,"ATOMS"]++}
  /ATOMG/ {c[f,"ATOMG"]++} / RED\./ {c[f,"RED"]++} /LDS/ {c[f,"LDS"]++} /STS/ {c[f,"STS"]++} /SHFL/ {c[f,"SHFL"]++} /VOTE/ {c[f,"VOTE"]++}
  /^[ \t]+\/\*[0-9a-f]+\*\// {n[f]++}
  END {for (f in n) printf "%s: %d SASS lines, UBLKCP %d, SYNCS %d, REDUX %d, ATOMS %d, ATOMG %d, RED %d, LDS %d, STS %d, SHFL %d, VOTE %d\n", f, n[f], c[f,"UBLKCP"], c[f,"SYNCS"], c[f,"REDUX"], c[f,"ATOMS"], c[f,"ATOMG"], c[f,"RED"], c[f,"LDS"], c[f,"STS"], c[f,"SHFL"], c[f,"VOTE"]}' | sort
echo; echo "# the bulk-copy / mbarrier instructions themselves (first 12)"
cuobjdump -sass $SO | grep -E "UBLKCP|SYNCS" | head -12
