#!/bin/bash
# SASS opcode evidence per object file (run in the build container after __graft_entry__.build()):
#   UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor (TMA), HMMA = mma.sync,
#   LDGMC = multimem.ld_reduce (NVLS), REDG = red.global (weight-gradient reductions), LDGSTS = cp.async,
#   F*2 = FADD2 / FMUL2 / FFMA2 (Blackwell packed fp32), ATOMS = shared-memory atomics (fp32 add = CAS loop)
out=${1:-profiles/sass_histogram.txt}
{
  echo "# cuobjdump -sass build/*.o | opcode counts ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
  printf "%-16s %8s %6s %6s %8s %8s %8s %8s %8s %8s %8s %6s %6s\n" object UTCxMMA LDTM STTM UTMALDG HMMA.tf32 HMMA.f16 LDGSTS LDGMC REDG SYNCS "F*2" ATOMS
  for o in build/*.o; do
    s=$(cuobjdump -sass $o 2>/dev/null)
    c() { echo "$s" | grep -c -E "$1"; }
    printf "%-16s %8d %6d %6d %8d %8d %8d %8d %8d %8d %8d %6d %6d\n" $(basename $o) "$(c 'UTC[A-Z]*MMA')" "$(c 'LDTM')" "$(c 'STTM')" \
      "$(c 'UTMALDG')" "$(c 'HMMA\.1688\.F32\.TF32')" "$(c 'HMMA\.16816')" "$(c 'LDGSTS')" "$(c 'LDGMC')" "$(c 'REDG')" "$(c 'SYNCS')" "$(c 'F(ADD|MUL|FMA)2')" "$(c 'ATOMS')"
  done
} > $out
cat $out
