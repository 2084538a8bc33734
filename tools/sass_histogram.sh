#!/bin/bash
# SASS evidence for profiles/: opcode histogram of the in-tree libagb.so (runs without a GPU), overall and per kernel for
# the Blackwell-specific mnemonics (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce,
# LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync).
cd "$(dirname "$0")/.."
OUT=${1:-profiles/r02_sass_opcodes.txt}
SO=affectgpt_b200/libagb.so
{
  echo "# cuobjdump -sass $SO ($(date -u +%Y-%m-%d)), sm_100a; sha256 $(sha256sum $SO | cut -c1-16)"
  echo "# ---- Blackwell / tensor mnemonics, whole library"
  cuobjdump -sass $SO | grep -oE '\b(UTCHMMA(\.2CTA)?|UTCQMMA|UTMALDG(\.[0-9A-Z.]+)?|UTMASTG(\.[0-9A-Z.]+)?|UTMAREDG(\.[0-9A-Z.]+)?|UTMAPF|LDTM(\.[0-9a-zA-Z.]+)?|STTM(\.[0-9a-zA-Z.]+)?|UTCBAR(\.[0-9A-Z.]+)?|UTCATOMSWS(\.[0-9A-Z.]+)?|HMMA\.[0-9A-Z.]+|SYNCS\.[A-Z.0-9]+|MUFU\.[A-Z0-9]+|FFMA2|FADD2|FMUL2)\b' | sort | uniq -c | sort -rn
  echo "# ---- per kernel: tcgen05.mma / TMA / TMEM instruction counts"
  cuobjdump -sass $SO | awk '
    /Function :/ { fn=$3; next }
    /UTCHMMA/ { mma[fn]++ } /UTMALDG/ { ld[fn]++ } /UTMASTG/ { st[fn]++ } /UTMAREDG/ { red[fn]++ } /LDTM/ { ldtm[fn]++ } /STTM/ { sttm[fn]++ } / HMMA\./ { hmma[fn]++ }
    END { for (f in mma) seen[f]=1; for (f in ld) seen[f]=1; for (f in hmma) seen[f]=1; for (f in ldtm) seen[f]=1;
          for (f in seen) printf "UTCHMMA=%-3d UTMALDG=%-3d UTMASTG=%-3d UTMAREDG=%-3d LDTM=%-3d STTM=%-3d HMMA=%-3d %s\n", mma[f], ld[f], st[f], red[f], ldtm[f], sttm[f], hmma[f], f }' | c++filt | sed -E 's/\(CUtensorMap_st.*//; s/\(float const\*.*//; s/\(__nv_bfloat16 const\*.*//' | sort -k8
  echo "# ---- top 30 opcodes, whole library"
  cuobjdump -sass $SO | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn | head -30
} > "$OUT"
wc -l "$OUT"
