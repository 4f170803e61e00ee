#!/bin/bash
# SASS opcode digest of the shipped library: the Blackwell-native instructions per kernel
# (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTMA* = tensor-map TMA)
cd "$(dirname "$0")/.."
cuobjdump -sass gordo_b200/lib/libgordo_b200.so 2>/dev/null | awk '
/Function :/ {fn=$3}
/UTCHMMA|UTCQMMA|UTCMMA/ {mma[fn]++}
/LDTM/ {ldtm[fn]++}
/STTM/ {sttm[fn]++}
/UBLKCP/ {blk[fn]++}
/UTMALDG|UTMASTG/ {tma[fn]++}
/MUFU/ {mu[fn]++}
/ FFMA/ {ff[fn]++}
/SYNCS|ARRIVES/ {sy[fn]++}
{n[fn]++}
END {for (f in n) printf "%-70s instr=%6d UTC*MMA=%3d LDTM=%3d STTM=%3d UBLKCP=%2d UTMA=%2d MUFU=%4d FFMA=%5d mbarrier-ops=%3d\n", f, n[f], mma[f], ldtm[f], sttm[f], blk[f], tma[f], mu[f], ff[f], sy[f]}' \
 | sed -E 's/_ZN[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_0-9]+_cu_[0-9a-f]+[0-9]+//; s/^_Z[0-9]*//' | sort
