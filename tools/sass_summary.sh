#!/bin/bash
# SASS opcode evidence + ptxas resource table of the in-tree objects (runs in the build container, no GPU):
#   bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
cd "$(dirname "$0")/../scene-aware-3d-multi-human_b200/csrc" || exit 1
make -s > /dev/null 2>&1
echo "# SASS opcode counts per object (cuobjdump -sass, sm_100a) -- tensor-core / TMEM / TMA / atomics evidence"
echo "# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops,"
echo "# REDG.E.ADD.64 = 64-bit fixed-point gradient reductions, ATOMS.CAST.SPIN = compare-and-swap loops on shared memory"
printf "%-18s %8s %7s %6s %7s %7s %6s %10s %10s %14s %8s %8s\n" object UTCHMMA UTCBAR LDTM UBLKCP UTMALDG SYNCS REDG.ADD64 REDG.F32 ATOMS.CAST.SPIN ATOMS ATOMG
for o in *.o; do
  s=$(cuobjdump -sass "$o")
  c() { echo "$s" | grep -c "$1"; }
  printf "%-18s %8d %7d %6d %7d %7d %6d %10d %10d %14d %8d %8d\n" "$o" $(c UTCHMMA) $(c UTCBAR) $(c LDTM) $(c UBLKCP) $(c UTMALDG) $(c "SYNCS") $(c "REDG.E.ADD.64") $(c "REDG.E.ADD.F32") $(c "ATOMS.CAST.SPIN") $(c "ATOMS") $(c "ATOMG")
done
echo
echo "# ptxas resource usage per kernel (registers / stack frame / spills / static shared memory)"
for l in *.ptxas.log; do
  awk -v f="${l%.ptxas.log}" '/Compiling entry function/ {k=$7} /bytes stack frame/ {st=$1; sp=$5; sl=$9} /Used [0-9]+ registers/ {r=$5; sm="0"; for(i=1;i<=NF;i++) if($i=="bytes" && $(i+1) ~ /smem/) sm=$(i-1); gsub(/\x27/,"",k); printf "%-12s %-60s regs %3d stack %4s spill-st %4s spill-ld %4s smem %s\n", f, substr(k,1,60), r, st, sp, sl, sm}' "$l"
done
echo
echo "# dynamic linkage of libmhopt.so (no cuBLAS / cuDNN / NCCL at link time; NCCL is bound with dlopen in mh_comm.cu)"
ldd ../libmhopt.so | awk '{print $1}' | tr '\n' ' '; echo
