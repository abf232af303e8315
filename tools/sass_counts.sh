#!/bin/sh
# SASS evidence of the sm_100a tensor-core / TMA path, per object file of libt2p.so:
#   UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor load / store,
#   UTCBAR = tcgen05.commit, HMMA.16816 = mma.sync (kept only for d = 16 heads and the final 128 -> 5 convolution)
cd "$(dirname "$0")/.."
printf "%-22s %8s %6s %6s %8s %8s %7s %11s %6s\n" object UTCHMMA LDTM STTM UTMALDG UTMASTG UTCBAR HMMA.16816 MUFU
for o in build/obj/*.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { printf "%s" "$s" | grep -c "$1"; }
  printf "%-22s %8s %6s %6s %8s %8s %7s %11s %6s\n" "$(basename "$o")" "$(c UTCHMMA)" "$(c LDTM)" "$(c STTM)" "$(c UTMALDG)" \
    "$(c UTMASTG)" "$(c UTCBAR)" "$(c 'HMMA.16816')" "$(c MUFU)"
done
echo
echo "per kernel (functions containing UTCHMMA):"
cuobjdump -sass text2protein_b200/libt2p.so 2>/dev/null | awk '/Function : /{name=$3} /UTCHMMA/{n[name]++} END{for (k in n) print n[k], k}' | sort -rn | c++filt | sed 's/t2p::(anonymous namespace):://' | cut -c1-110
