#!/bin/bash
# SASS evidence of the Blackwell-native instructions in the shipped library (no GPU needed):
#   tools/sass_summary.sh > profiles/sass_summary_r2.txt
LIB=smart-tree_b200/libst_b200.so
echo "# cuobjdump -sass $LIB: instruction counts per kernel (only kernels with at least one of them are listed)"
echo "# UTC*MMA = tcgen05.mma, STTM/LDTM = tcgen05.st/ld (TMEM), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 = fma.rn.f32x2"
cuobjdump -sass $LIB | awk '
/Function :/ { fn=$3 }
/UTC[A-Z]*MMA/ { mma[fn]++ } /STTM/ { st[fn]++ } /LDTM/ { ld[fn]++ } /UBLKCP/ { blk[fn]++ } /UTCBAR/ { bar[fn]++ } /SYNCS/ { sy[fn]++ } /LDGSTS/ { lg[fn]++ } /FFMA2/ { f2[fn]++ } /UCGABAR|CGABAR/ { cga[fn]++ } /MEMBAR.ALL.GPU/ { mb[fn]++ } /CCTL.IVALL/ { iv[fn]++ }
END { printf "%-10s %-6s %-6s %-7s %-7s %-6s %-7s %-6s %-7s %-10s %-7s %s\n","UTCxMMA","STTM","LDTM","UBLKCP","UTCBAR","SYNCS","LDGSTS","FFMA2","CGABAR","MEMBAR.ALL","IVALL","kernel";
  for (f in mma) seen[f]=1; for (f in blk) seen[f]=1; for (f in f2) seen[f]=1; for (f in lg) seen[f]=1; for (f in cga) seen[f]=1; for (f in mb) seen[f]=1;
  for (f in seen) printf "%-10d %-6d %-6d %-7d %-7d %-6d %-7d %-6d %-7d %-10d %-7d %s\n", mma[f],st[f],ld[f],blk[f],bar[f],sy[f],lg[f],f2[f],cga[f],mb[f],iv[f],f }' | (read -r h; echo "$h"; sort -k12 | c++filt | cut -c1-230)
